"""Writes tests/golden/crf_restatement.npz: for each case of crf_cases.py the inputs (guide image, unary) and what the C
restatement of pydensecrf (oracle/densecrf.c) makes of them through the reference's call sequence (DRV:1063-1072):
Q after 1, 3 and 10 mean-field iterations and the MAP labels.  The dense CRF is the one stage whose parity is UNPINNED
(pydensecrf is not installable offline); the day a pydensecrf build exists,

    python tests/golden/diff_pydensecrf.py

runs the real library on the stored inputs and prints its distance from these stored outputs.

    python tests/golden/make_crf_restatement.py        # regenerate (deterministic)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from crf_cases import CASES, make_case  # noqa: E402
from oracle import densecrf as D  # noqa: E402

ITERS = (1, 3, 10)


def run_restatement(img, U, C, H, W, n_iter):
    d = D.DenseCRF2D(W, H, C)
    d.setUnaryEnergy(np.ascontiguousarray(U))
    d.addPairwiseGaussian(sxy=3, compat=7)
    d.addPairwiseBilateral(sxy=50, srgb=5, rgbim=img, compat=10)
    return np.array(d.inference(n_iter), dtype=np.float32)


if __name__ == "__main__":
    out = {}
    for name, (_, H, W, C) in CASES.items():
        img, p = make_case(name)
        U = D.unary_from_softmax(p)
        out[name + "_image"], out[name + "_unary"] = img, U.astype(np.float32)
        for it in ITERS:
            out["%s_Q%d" % (name, it)] = run_restatement(img, U, C, H, W, it)
        out[name + "_map"] = out[name + "_Q10"].argmax(0).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "crf_restatement.npz"), **out)
    print("wrote", os.path.join(HERE, "crf_restatement.npz"), sorted(out))
