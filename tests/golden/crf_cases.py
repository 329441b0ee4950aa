"""Small dense-CRF cases shared by the exact O(N^2) second-opinion tests and by the committed fixture
tests/golden/crf_restatement.npz (inputs + outputs of oracle/densecrf.c, ready to be diffed against a real pydensecrf build):
piecewise-constant guide images with mild noise and smooth, slightly noisy class probabilities -- the regime in which the
permutohedral lattice approximates the true Gaussian kernels well."""
import numpy as np

CASES = {"blocks_28x28_3": (1, 28, 28, 3), "blocks_24x32_4": (2, 24, 32, 4), "blocks_32x32_5": (3, 32, 32, 5)}


def make_case(name):
    """-> (image uint8 [H,W,3], probabilities float32 [C,H,W] summing to one over C)."""
    seed, H, W, C = CASES[name]
    rng = np.random.default_rng(seed)
    img = np.zeros((H, W, 3), np.float64)
    img[:, : W // 2] = (60, 90, 120)
    img[:, W // 2:] = (200, 160, 90)
    img[H // 2:, : W // 3] = (20, 200, 40)
    img = np.clip(img + rng.normal(0, 2.0, img.shape), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    p = np.stack([0.5 + 0.4 * np.sin(2 * np.pi * (xx / W * 1.3 + yy / H * 0.7 + k / C)) for k in range(C)]).astype(np.float32) ** 2 + 0.1
    p += rng.random((C, H, W)).astype(np.float32) * 0.15
    p /= p.sum(0, keepdims=True)
    return img, p
