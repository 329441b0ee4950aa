"""CPU: property tests of the dense-CRF restatement (oracle/densecrf.c).  pydensecrf is not available to diff against
(PARITY UNPINNED, see the file header), so the oracle is held to the properties the published algorithm has."""
import numpy as np
import torch

import synth
from oracle import densecrf as D


def _crf(H, W, C, img):
    d = D.DenseCRF2D(W, H, C)
    d.addPairwiseGaussian(3, 7)
    d.addPairwiseBilateral(50, 5, img, 10)
    return d


def test_marginals_are_distributions_and_iter0_is_softmax():
    H, W, C = 40, 36, 5
    img = synth.guide_image(1, H, W)
    rng = np.random.default_rng(0)
    p = rng.random((C, H, W)).astype(np.float32)
    p /= p.sum(0, keepdims=True)
    U = D.unary_from_softmax(p)
    d = _crf(H, W, C, img)
    d.setUnaryEnergy(np.ascontiguousarray(U))
    Q0 = d.inference(0)
    assert np.allclose(Q0, p.reshape(C, -1), rtol=1e-5, atol=1e-7)      # softmax(-U) = softmax(log p) = p
    Q = d.inference(10)
    assert Q.min() >= 0 and np.allclose(Q.sum(0), 1.0, atol=1e-5)


def test_filter_is_linear_and_symmetric():
    H, W, C = 32, 32, 3
    d = _crf(H, W, C, synth.guide_image(2, H, W))
    rng = np.random.default_rng(1)
    x, y = rng.random((C, H * W)).astype(np.float32), rng.random((C, H * W)).astype(np.float32)
    for k in (0, 1):
        Kx, Ky = d.kernel_apply(k, x), d.kernel_apply(k, y)
        assert np.allclose(d.kernel_apply(k, 2 * x + 0.5 * y), 2 * Kx + 0.5 * Ky, rtol=1e-4, atol=1e-5)
        a, b = float((y.astype(np.float64) * Kx).sum()), float((Ky.astype(np.float64) * x).sum())
        # splat and slice are transposes and each axis blur is symmetric; the axis blurs only commute approximately,
        # so K is symmetric to ~1e-3 (densecrf's `reverse` order exists for exactly this reason)
        assert abs(a - b) <= 1e-2 * abs(a)


def test_norm_is_inverse_sqrt_of_K_ones():
    H, W = 24, 30
    lat = D.Lattice(np.stack(np.meshgrid(np.arange(W) / 3.0, np.arange(H) / 3.0), -1).reshape(-1, 2).astype(np.float32))
    K1 = lat.compute(np.ones((H * W, 1), np.float32))[:, 0]
    d = D.DenseCRF2D(W, H, 2)
    d.addPairwiseGaussian(3, 1)
    assert np.allclose(d.kernel_norm(0), 1.0 / np.sqrt(K1 + 1e-20), rtol=1e-6)
    assert K1.min() > 0


def test_spatial_filter_approximates_a_gaussian():
    """The d=2 lattice filter is an approximation of the Gaussian exp(-|p-q|^2 / (2 sxy^2)) (Adams et al. 2010)."""
    H = W = 32
    sxy = 3.0
    yy, xx = np.mgrid[0:H, 0:W]
    feat = np.stack([xx / sxy, yy / sxy], -1).reshape(-1, 2).astype(np.float32)
    lat = D.Lattice(feat)
    x = np.zeros((H * W, 1), np.float32)
    x[16 * W + 16] = 1.0                                           # impulse in the middle
    got = lat.compute(x)[:, 0].reshape(H, W)
    want = np.exp(-((xx - 16) ** 2 + (yy - 16) ** 2) / (2 * sxy ** 2))
    got, want = got / got.sum(), want / want.sum()
    assert np.abs(got - want).sum() < 0.25                          # total-variation distance of the two kernels
    cy, cx = (got * yy).sum(), (got * xx).sum()
    assert abs(cy - 16) < 0.5 and abs(cx - 16) < 0.5               # centred


def test_strong_pairwise_term_smooths_the_labelling():
    H, W, C = 48, 48, 2
    img = np.full((H, W, 3), 128, np.uint8)
    rng = np.random.default_rng(3)
    p = np.full((C, H, W), 0.5, np.float32)
    p[0, :, :24] = 0.7
    p[0, :, 24:] = 0.3
    p[0] += rng.normal(0, 0.25, (H, W)).astype(np.float32)          # noisy unary around a clean left/right split
    p[0] = np.clip(p[0], 0.02, 0.98)
    p[1] = 1 - p[0]
    noisy = np.argmax(p, 0)
    lab = D.densecrf(img, torch.log(torch.from_numpy(p)))
    truth = np.zeros((H, W), np.int64)
    truth[:, 24:] = 1
    assert (lab != truth).mean() < 0.5 * (noisy != truth).mean()


def test_bilateral_filter_approximates_the_bilateral_gaussian():
    """The d=5 lattice over (x/sxy, y/sxy, r/srgb, g/srgb, b/srgb): filtering a random field must track the brute-force
    bilateral Gaussian sum_j exp(-|f_i - f_j|^2 / 2) x_j, and must respect a colour edge (no leakage across it)."""
    H = W = 24
    sxy, srgb = 6.0, 8.0
    rng = np.random.default_rng(3)
    img = np.zeros((H, W, 3), np.float32)
    img[:, : W // 2] = (40, 60, 80)
    img[:, W // 2:] = (200, 180, 160)                              # a hard vertical colour edge
    img += rng.normal(0, 2.0, img.shape).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    feat = np.concatenate([np.stack([xx / sxy, yy / sxy], -1), img / srgb], -1).reshape(-1, 5).astype(np.float32)
    lat = D.Lattice(feat)
    x = rng.random((H * W, 1)).astype(np.float32)
    got = lat.compute(x)[:, 0]
    d2 = ((feat[:, None, :] - feat[None, :, :]) ** 2).sum(-1)
    want = (np.exp(-0.5 * d2) @ x)[:, 0]
    got, want = got / got.mean(), want / want.mean()               # the lattice kernel is Gaussian up to a constant
    assert np.corrcoef(got, want)[0, 1] > 0.97
    assert np.abs(got - want).mean() < 0.06                       # measured 0.035 with a field std of 0.22
    # impulse on the left side of the edge: (almost) nothing arrives on the right side
    e = np.zeros((H * W, 1), np.float32)
    e[12 * W + 8] = 1.0
    r = lat.compute(e)[:, 0].reshape(H, W)
    assert r[:, W // 2:].sum() < 1e-3 * r[:, : W // 2].sum()


# ------------------------------------------------------------------------------------------------ second opinion: exact O(N^2) mean-field
def _restatement(img, U, C, H, W, n_iter, pos_w=7, bi_w=10):
    d = D.DenseCRF2D(W, H, C)
    d.setUnaryEnergy(np.ascontiguousarray(U))
    if pos_w:
        d.addPairwiseGaussian(sxy=3, compat=pos_w)
    if bi_w:
        d.addPairwiseBilateral(sxy=50, srgb=5, rgbim=img, compat=bi_w)
    return np.array(d.inference(n_iter))


def test_restatement_tracks_the_exact_dense_crf_and_not_its_wrong_variants():
    """oracle/densecrf.c against an independent O(N^2) mean-field with the TRUE Gaussian kernels (oracle/exact_meanfield.py,
    no shared code) on <= 32x32 images.  The lattice is an approximation, so the bound is loose (a few 1e-2) -- but every
    deliberately wrong model (no normalisation, row instead of symmetric normalisation, the lattice's alpha forgotten, the
    Potts sign flipped, kernel widths off, weights swapped) is several times further away than the right one."""
    import crf_cases
    from oracle.exact_meanfield import exact_dense_crf
    for name, (_, H, W, C) in crf_cases.CASES.items():
        img, p = crf_cases.make_case(name)
        U = D.unary_from_softmax(p)
        for pos_w, bi_w in ((7, 10), (7, 0), (0, 10)):
            Ql = _restatement(img, U, C, H, W, 1, pos_w, bi_w)
            kw = dict(pos_w=pos_w, bi_w=bi_w)
            err = lambda **k: float(np.abs(Ql - exact_dense_crf(img, U, 1, **dict(kw, **k))).mean())
            right = err()
            assert right <= 0.006 and float(np.abs(Ql - exact_dense_crf(img, U, 1, **kw)).max()) <= 0.06, (name, pos_w, bi_w, right)
            assert err(normalize=None) >= 10 * right, "an unnormalised kernel must be far away"
            assert err(msg_scale=-1.0) >= 10 * right, "the Potts message has the wrong sign"
            assert err(normalize="row") >= 1.5 * right
            if pos_w and not bi_w:   # d = 2: alpha = 0.8 (the d = 5 lattice's 1/(1+2^-5) = 0.97 is too close to 1 to see)
                assert err(msg_scale=1.0 + 2.0 ** -2) >= 2.5 * right, "alpha = 1/(1+2^-d) of the lattice is missing"
            if pos_w:
                assert err(pos_xy_std=4.5) >= 3 * right, "spatial kernel width"
            if bi_w and not pos_w:
                assert err(bi_rgb_std=10.0) >= 1.3 * right, "colour kernel width"
            if pos_w and bi_w:
                assert err(pos_w=bi_w, bi_w=pos_w) >= 4 * right, "kernel weights swapped"
        # three iterations in: still close (the error grows as the marginals saturate), labels agree
        Q3, E3 = _restatement(img, U, C, H, W, 3), exact_dense_crf(img, U, 3)
        assert float(np.abs(Q3 - E3).mean()) <= 0.012 and float((Q3.argmax(0) == E3.argmax(0)).mean()) >= 0.97


def test_crf_restatement_fixture_is_current():
    """tests/golden/crf_restatement.npz holds inputs + outputs of oracle/densecrf.c (made by make_crf_restatement.py) so that a
    pydensecrf build can be diffed in one command (tests/golden/diff_pydensecrf.py); it must describe today's restatement."""
    import os
    import crf_cases
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crf_restatement.npz"))
    for name, (_, H, W, C) in crf_cases.CASES.items():
        img, p = crf_cases.make_case(name)
        assert np.array_equal(g[name + "_image"], img)
        U = D.unary_from_softmax(p)
        assert np.allclose(g[name + "_unary"], U, rtol=0, atol=1e-6)
        for it in (1, 3, 10):
            assert np.allclose(g["%s_Q%d" % (name, it)], _restatement(img, g[name + "_unary"], C, H, W, it), rtol=0, atol=2e-6)
        assert np.array_equal(g[name + "_map"], g[name + "_Q10"].argmax(0))
