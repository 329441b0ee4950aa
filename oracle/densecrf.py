"""oracle/densecrf.py -- TEST INFRASTRUCTURE.  ctypes face of oracle/densecrf.c with the pydensecrf
surface the reference uses (PnP_OVSS_0514_updated_segmentation.py:1063-1072).  PARITY UNPINNED (see
oracle/densecrf.c header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/densecrf.c -> oracle/libdensecrf_oracle.so (gcc, seconds)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libdensecrf_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.dcrf_create.restype = vp
        L.dcrf_create.argtypes = [ci, ci, ci]
        L.dcrf_free.argtypes = [vp]
        L.dcrf_set_unary.argtypes = [vp, vp]
        L.dcrf_add_pairwise_gaussian.argtypes = [vp, cf, cf, cf]
        L.dcrf_add_pairwise_bilateral.argtypes = [vp, cf, cf, cf, cf, cf, vp, cf]
        L.dcrf_inference.argtypes = [vp, ci, vp]
        L.dcrf_kernel_M.argtypes = [vp, ci]
        L.dcrf_kernel_M.restype = ci
        L.dcrf_kernel_norm.argtypes = [vp, ci]
        L.dcrf_kernel_norm.restype = vp
        L.dcrf_kernel_lattice.argtypes = [vp, ci]
        L.dcrf_kernel_lattice.restype = vp
        L.dcrf_kernel_apply.argtypes = [vp, ci, vp, vp, ci]
        L.pl_create.restype = vp
        L.pl_create.argtypes = [vp, ci, ci]
        L.pl_free.argtypes = [vp]
        L.pl_compute.argtypes = [vp, vp, vp, ci, ci]
        for name in ("pl_M", "pl_N", "pl_d"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = ci
        for name in ("pl_offset", "pl_barycentric", "pl_n1", "pl_n2", "pl_keys"):
            getattr(L, name).argtypes = [vp]
            getattr(L, name).restype = vp
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def unary_from_softmax(sm, scale=None, clip=1e-5):
    """pydensecrf.utils.unary_from_softmax: -log(clip(p, 1e-5, 1)) as float32 [C, N]."""
    num_cls = sm.shape[0]
    if scale is not None:
        uniform = np.ones(sm.shape) / num_cls
        sm = scale * sm + (1 - scale) * uniform
    if clip is not None:
        sm = np.clip(sm, clip, 1.0)
    return -np.log(sm).reshape([num_cls, -1]).astype(np.float32)


class Lattice:
    """Permutohedral lattice over features [N, d] (float32)."""

    def __init__(self, feature):
        feature = np.ascontiguousarray(feature, dtype=np.float32)
        self.N, self.d = feature.shape
        self._h = lib().pl_create(_ptr(feature), self.N, self.d)
        self.M = lib().pl_M(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().pl_free(self._h)
            self._h = None

    def _view(self, fn, shape, dtype):
        addr = getattr(lib(), fn)(self._h)
        n = int(np.prod(shape))
        buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    @property
    def offset(self):
        return self._view("pl_offset", (self.N, self.d + 1), np.int32)

    @property
    def barycentric(self):
        return self._view("pl_barycentric", (self.N, self.d + 1), np.float32)

    @property
    def n1(self):
        return self._view("pl_n1", (self.d + 1, self.M), np.int32)

    @property
    def n2(self):
        return self._view("pl_n2", (self.d + 1, self.M), np.int32)

    @property
    def keys(self):
        return self._view("pl_keys", (self.M, self.d), np.int16)

    def compute(self, x, reverse=False):
        """x [N, vs] -> filtered [N, vs] (un-normalised splat/blur/slice)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        lib().pl_compute(self._h, _ptr(out), _ptr(x), x.shape[1], int(reverse))
        return out


class DenseCRF2D:
    """Same call surface as pydensecrf.densecrf.DenseCRF2D for the calls the reference makes."""

    def __init__(self, w, h, c):
        self.w, self.h, self.c = int(w), int(h), int(c)
        self._h = lib().dcrf_create(self.w, self.h, self.c)
        self._n_kernels = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dcrf_free(self._h)
            self._h = None

    def setUnaryEnergy(self, U):
        U = np.asarray(U)
        if U.dtype != np.float32 or not U.flags.c_contiguous:
            raise ValueError("Buffer dtype mismatch / ndarray is not C-contiguous")  # as pydensecrf
        if U.shape != (self.c, self.w * self.h):
            raise ValueError("Bad shape for unary energy (Need (%d, %d), got %s)" % (self.c, self.w * self.h, U.shape))
        lib().dcrf_set_unary(self._h, _ptr(U))

    def addPairwiseGaussian(self, sxy, compat):
        sx, sy = (sxy, sxy) if np.isscalar(sxy) else sxy
        lib().dcrf_add_pairwise_gaussian(self._h, float(sx), float(sy), float(compat))
        self._n_kernels += 1

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat):
        sx, sy = (sxy, sxy) if np.isscalar(sxy) else sxy
        sr, sg, sb = (srgb, srgb, srgb) if np.isscalar(srgb) else srgb
        rgbim = np.asarray(rgbim)
        if rgbim.dtype != np.uint8 or not rgbim.flags.c_contiguous or rgbim.shape != (self.h, self.w, 3):
            raise ValueError("Bad shape for pairwise bilateral (Need (%d, %d, 3) uint8 C-contiguous)" % (self.h, self.w))
        lib().dcrf_add_pairwise_bilateral(self._h, float(sx), float(sy), float(sr), float(sg), float(sb), _ptr(rgbim), float(compat))
        self._n_kernels += 1

    def inference(self, n):
        Q = np.empty((self.c, self.w * self.h), dtype=np.float32)
        lib().dcrf_inference(self._h, int(n), _ptr(Q))
        return Q

    # -- test helpers (not part of pydensecrf) --
    def kernel_M(self, k):
        return lib().dcrf_kernel_M(self._h, k)

    def kernel_norm(self, k):
        addr = lib().dcrf_kernel_norm(self._h, k)
        n = self.w * self.h
        return np.frombuffer((ctypes.c_char * (4 * n)).from_address(addr), dtype=np.float32).copy()

    def kernel_apply(self, k, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        lib().dcrf_kernel_apply(self._h, k, _ptr(x), _ptr(out), x.shape[0])
        return out


def densecrf(image, mask, n_iter=10, pos_w=7, pos_xy_std=3, bi_w=10, bi_xy_std=50, bi_rgb_std=5, return_q=False):
    """Restates densecrf() of PnP_OVSS_0514_updated_segmentation.py:1030-1074.

    image uint8 [H,W,3]; mask float [C',H,W] (torch tensor or ndarray).  Returns float32 [H,W] argmax map
    (and Q [C',H,W] when return_q)."""
    import torch
    import torch.nn.functional as F

    image = np.ascontiguousarray(image).copy()
    output_logits = torch.as_tensor(mask)
    output_probs = F.softmax(output_logits, dim=0).cpu().numpy()
    c, h, w = output_probs.shape
    U = unary_from_softmax(output_probs)
    U = np.ascontiguousarray(U)
    d = DenseCRF2D(w, h, c)
    d.setUnaryEnergy(U)
    d.addPairwiseGaussian(sxy=pos_xy_std, compat=pos_w)
    d.addPairwiseBilateral(sxy=bi_xy_std, srgb=bi_rgb_std, rgbim=image, compat=bi_w)
    Q = d.inference(n_iter)
    Q = np.array(Q).reshape((c, h, w))
    MAP = np.argmax(Q, axis=0).reshape((h, w)).astype(np.float32)
    if return_q:
        return MAP, Q
    return MAP
