"""One command to pin the dense-CRF restatement the day a pydensecrf build is available:

    python tests/golden/diff_pydensecrf.py

runs the REAL pydensecrf (lucasb-eyer/pydensecrf, the library the reference calls at DRV:1031-1032, 1063-1071) on the inputs
stored in tests/golden/crf_restatement.npz and prints its distance from the stored outputs of oracle/densecrf.c.
Exit code 0 iff every marginal agrees within 1e-3 and every MAP label is equal."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from crf_cases import CASES  # noqa: E402

try:
    import pydensecrf.densecrf as dcrf
except ImportError:
    print("pydensecrf is not installed here: nothing to diff (the stored outputs remain unpinned)")
    sys.exit(2)

g = np.load(os.path.join(HERE, "crf_restatement.npz"))
worst, ok = 0.0, True
for name, (_, H, W, C) in CASES.items():
    img, U = np.ascontiguousarray(g[name + "_image"]), np.ascontiguousarray(g[name + "_unary"])
    for it in (1, 3, 10):
        d = dcrf.DenseCRF2D(W, H, C)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(sxy=3, compat=7)
        d.addPairwiseBilateral(sxy=50, srgb=5, rgbim=img, compat=10)
        Q = np.array(d.inference(it)).reshape(C, -1)
        err = float(np.abs(Q - g["%s_Q%d" % (name, it)]).max())
        worst = max(worst, err)
        line = "%s, %2d iterations: max |Q_pydensecrf - Q_restatement| = %.3e" % (name, it, err)
        if it == 10:
            same = bool((Q.argmax(0) == g[name + "_map"]).all())
            ok = ok and same
            line += "   MAP labels %s" % ("equal" if same else "DIFFER at %d pixels" % int((Q.argmax(0) != g[name + "_map"]).sum()))
        print(line)
ok = ok and worst <= 1e-3
print("restatement %s (worst marginal difference %.3e)" % ("PINNED" if ok else "DIFFERS from pydensecrf", worst))
sys.exit(0 if ok else 1)
