"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the PnP-OVSS mask-extraction hot path, written
line-for-line after the reference sources cited in each function.  It exists to CHECK the CUDA path.

Rules (enforced by tests/test_boundary.py):
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
    this package;
  * the product package pnp_ovss_b200 never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle pinning"):
  * a1-a7, a9, a10: pinned against outputs of the reference's OWN functions, extracted from
    /root/reference with ast and executed in the build container by tests/golden/make_golden.py
    (fixtures committed under tests/golden/*.npz).
  * a8 (dense CRF): PARITY UNPINNED -- pydensecrf is neither vendored in the reference, nor installed,
    nor buildable offline; oracle/densecrf.c restates the published algorithm and is validated by
    property tests only.
"""
