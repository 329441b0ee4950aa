"""oracle/exact_meanfield.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second opinion on the dense-CRF restatement (oracle/densecrf.c) that shares no code with it: mean-field inference of the
fully connected CRF of Kraehenbuehl & Koltun (NIPS 2011, eq. 1-4 / Algorithm 1) with the TRUE Gaussian kernels evaluated pair
by pair in float64 -- O(N^2), for images of at most ~32x32.

    k_g(i,j) = exp(-|p_i - p_j|^2 / (2 sxy_g^2))                          addPairwiseGaussian(sxy=3, compat=7)   (DRV:1068)
    k_b(i,j) = exp(-|p_i - p_j|^2 / (2 sxy_b^2) - |I_i - I_j|^2 / (2 srgb^2))   addPairwiseBilateral(50, 5, img, 10) (DRV:1069)

with what pydensecrf's defaults add: NORMALIZE_SYMMETRIC (K~ = D^-1/2 K D^-1/2, D = diag(K 1)), the self term j = i kept
(the permutohedral filter cannot leave it out), Potts compatibility mu(l,l') = -w [l = l'] folded as a +w K~ Q message, and
Q <- softmax(-U + sum_k w_k K~_k Q) for n iterations starting from softmax(-U).

The permutohedral lattice only APPROXIMATES these kernels (Adams et al. 2010), so the restatement -- and the CUDA path,
which matches the restatement to 1e-3 -- are expected to track this exact result to a few 1e-2 in the marginals, not better.
A restatement error in the parts that matter (feature scaling, the alpha normalisation of the lattice, the symmetric
normalisation, the sign or weight of the Potts message, the softmax) moves the marginals by far more than that."""
import numpy as np


def exact_dense_crf(image, unary, n_iter=10, pos_w=7.0, pos_xy_std=3.0, bi_w=10.0, bi_xy_std=50.0, bi_rgb_std=5.0,
                    normalize="symmetric", msg_scale=1.0):
    """image uint8 [H,W,3]; unary float [C,H*W] (= -log p).  Returns Q float64 [C,H*W].  pos_w / bi_w = 0 leaves that kernel out.

    normalize ("symmetric" | "row" | None) and msg_scale exist so that tests can build deliberately WRONG models (no
    normalisation, the lattice's alpha = 1/(1+2^-d) forgotten, the Potts sign flipped, ...) and show that the lattice
    restatement is far from them while it is close to the right one."""
    H, W, _ = image.shape
    N = H * W
    yy, xx = np.mgrid[0:H, 0:W]
    pos = np.stack([xx.reshape(-1), yy.reshape(-1)], 1).astype(np.float64)
    col = image.reshape(N, 3).astype(np.float64)
    d_pos = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    d_col = ((col[:, None, :] - col[None, :, :]) ** 2).sum(-1)
    kernels = []
    if pos_w:
        kernels.append((pos_w, np.exp(-d_pos / (2.0 * pos_xy_std ** 2))))
    if bi_w:
        kernels.append((bi_w, np.exp(-d_pos / (2.0 * bi_xy_std ** 2) - d_col / (2.0 * bi_rgb_std ** 2))))
    normed = []
    for w, K in kernels:
        if normalize == "symmetric":
            s = 1.0 / np.sqrt(K.sum(1) + 1e-20)
            K = K * s[:, None] * s[None, :]
        elif normalize == "row":
            K = K / K.sum(1, keepdims=True)
        normed.append((w * msg_scale, K))
    U = np.asarray(unary, np.float64)

    def softmax(t):
        t = t - t.max(0, keepdims=True)
        e = np.exp(t)
        return e / e.sum(0, keepdims=True)

    Q = softmax(-U)
    for _ in range(n_iter):
        t = -U
        for w, K in normed:
            t = t + w * (Q @ K.T)
        Q = softmax(t)
    return Q
