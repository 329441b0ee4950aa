"""CPU stand-ins for pnp_ovss_b200.ops used by the dry-run tests: same signatures and output shapes/dtypes, trivially
simple arithmetic (torch CPU).  They let the HOST composition (pipeline.batch_confusion, bench bookkeeping) execute on
a box without a GPU so that Python-level mistakes surface in the CPU suite.  They are NOT a fallback: nothing in the
product imports this file, and the numbers they produce mean nothing."""
import types

import torch


class _Lattice:
    def __init__(self, M):
        self.M = M
        self.struct = types.SimpleNamespace(max_row=1, n_vertices=M)
        self.storage = torch.zeros(1)


def make(ops_module):
    """A namespace with every function pipeline.py calls, shaped like `ops_module`'s."""
    ns = types.SimpleNamespace()
    ns.calls = []

    def log(name):
        ns.calls.append(name)

    def token_merge(gradcam, seg_start, seg_len, seg_div, row_offset=3, max_end=None):
        log("token_merge")
        assert max_end is not None, "the pipeline must hand the host-side bound over (no device read-back)"
        B, C = seg_start.shape
        out = torch.zeros((B, C) + tuple(gradcam.shape[2:]))
        for b in range(B):
            for c in range(C):
                s, l = int(seg_start[b, c]) + row_offset, int(seg_len[b, c])
                if l:
                    out[b, c] = gradcam[b, s:s + l].sum(0) / seg_div[b, c]
        return out

    def salience_dropout_round(gradcam, agg, chosen, n_prev, imgs, norm_imgs, P, patch, row_lo, row_hi, save_len, round_idx, ensemble_r=None):
        log("salience_dropout_round")
        if ensemble_r is not None:
            ensemble_r.copy_(gradcam)
        if round_idx == 0:
            agg.copy_(gradcam * 2)
        else:
            agg.add_(gradcam)
        score = gradcam[:, row_lo:row_hi].sum(1).flatten(1)
        chosen[:, n_prev:n_prev + save_len] = score.argsort(1)[:, -save_len:].to(torch.int32)

    def threshold_upsample(class_maps, H, W, threshold, rescale, with_background):
        log("threshold_upsample")
        x = torch.nn.functional.interpolate(class_maps, size=(H, W), mode="bilinear", align_corners=True)
        if with_background:
            x = torch.cat([(x.max(1, keepdim=True)[0] == 0).float(), x], 1)
        return x.contiguous()

    def crf_pad_channels(C):
        return (C + 3) // 4 * 4

    def lowrank_blur_unary(class_maps, H, W, threshold, rescale, with_background, sigma, unary=True, labels=False, maps=False, minmax=False,
                           n_classes=None):
        log("lowrank_blur_unary")
        x = threshold_upsample(class_maps, H, W, threshold, rescale, with_background)
        ns.calls.pop()
        B, Cc = x.shape[:2]
        if n_classes is not None:   # the channels an image lacks are dead
            assert n_classes.dtype == torch.int32 and n_classes.shape == (B,)
            for b in range(B):
                x[b, int(n_classes[b]) + (1 if with_background else 0):] = -float("inf")
        out = {}
        if unary:
            out["unary"] = crf_unary_from_maps(x.view(B, Cc, H * W))
            ns.calls.pop()
            if n_classes is not None:
                out["unary"][:, :, Cc:] = float("inf")
        if labels:
            out["labels"] = x.view(B, Cc, H * W).argmax(1).to(torch.int32)
        if maps:
            out["maps"] = x
        return out

    def gaussian_blur(maps, sigma, normalize=True):
        log("gaussian_blur")
        n = maps.numel() // (maps.shape[-1] * maps.shape[-2])
        return maps.clone(), torch.stack([maps.reshape(n, -1).min(1)[0], maps.reshape(n, -1).max(1)[0]], 1)

    def crf_unary_from_maps(maps, minmax=None):
        log("crf_unary_from_maps")
        B, C, N = maps.shape
        Cp = (C + 3) // 4 * 4
        U = torch.zeros(B, N, Cp)
        U[:, :, :C] = -torch.log_softmax(maps, 1).transpose(1, 2)
        return U

    def build_lattice(H, W, sxy, rgb=None, srgb=None, device=None):
        log("build_lattice")
        return _Lattice(H * W // 4)

    def build_lattice_begin(H, W, sxy, rgb=None, srgb=None, device=None):
        log("build_lattice")
        return _Lattice(H * W // 4), None

    def build_lattice_finish(lat, ws):
        return lat

    def crf_inference(lattices, weights, unary, C, n_iter, want_labels=True, scratch=None):
        log("crf_inference")
        return unary, unary[:, :, :C].argmin(2).to(torch.int32)

    def argmax_channels(maps):
        log("argmax_channels")
        return maps.argmax(1).to(torch.int32)

    def confusion_accumulate(labels, gt, n_class, hist, lut=None, pred_out=None, bad_count=None):
        log("confusion_accumulate")
        pred = labels.long() if lut is None else torch.gather(lut.long(), 1, labels.long())
        if pred_out is not None:
            pred_out.copy_(pred.float())
        ok = (gt >= 0) & (gt < n_class)
        idx = n_class * gt[ok].long() + pred[ok]
        hist += torch.bincount(idx, minlength=n_class * n_class).view(n_class, n_class)
        return hist

    for k, v in list(locals().items()):
        if callable(v) and k not in ("make", "log"):
            assert hasattr(ops_module, k), k
            setattr(ns, k, v)
    return ns
