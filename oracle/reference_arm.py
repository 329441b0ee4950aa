"""oracle/reference_arm.py -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's own way of running the path on the host CPU, used by bench.py's `--impl reference` arm and its
`cpu_baseline` leg: torch-CPU model pass exactly as BITM:386-457 drives it (capture in all 12 cross-attention
blocks, loss.backward() through the whole model, 12x12 GradCAM maps built, one used), then the per-image CPU
post-processing of DRV:424-481 through oracle.hotpath (torch CPU interpolate, the real scipy gaussian_filter, the
C restatement of pydensecrf, numpy argmax / bincount).  The model architecture/weights come from
pnp_ovss_b200.blip_itm (random init; the checker may import the product, never the other way round); its
cross-attention modules are swapped for the torch restatement of MED:191-311 below, so no CUDA kernel of the
product is on this path."""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import hotpath as O


class ReferenceCrossAttention(nn.Module):
    """BertSelfAttention(is_cross_attention=True) as the reference runs it (MED:191-311), torch only."""

    def __init__(self, src):
        super().__init__()
        self.query, self.key, self.value, self.heads = src.query, src.key, src.value, src.heads
        self.save_attention = False
        self.attention_map = None
        self.attn_gradients = None

    def save_attn_gradients(self, g):      # MED:164-165
        self.attn_gradients = g

    def get_attention_map(self):           # MED:176-177
        return self.attention_map

    def get_attn_gradients(self):          # MED:167-168
        return self.attn_gradients

    def _split(self, x):
        B, L, D = x.shape
        return x.view(B, L, self.heads, D // self.heads).permute(0, 2, 1, 3)

    def forward(self, hidden, enc, enc_mask=None, kv=None, lin=None):
        q, k, v = self._split(self.query(hidden)), self._split(self.key(enc)), self._split(self.value(enc))
        scores = torch.matmul(q, k.transpose(-1, -2))          # MED:228
        scores = scores / math.sqrt(q.shape[-1])               # MED:267
        if enc_mask is not None:
            scores = scores + enc_mask[:, None, None, :]       # MED:269-271
        probs = nn.Softmax(dim=-1)(scores)                     # MED:274
        if self.save_attention:                                # MED:280-283
            self.attention_map = probs
            self.attn_vo = torch.matmul(probs, v)              # the extra (unused) product of MED:282
            probs.register_hook(self.save_attn_gradients)
        ctx = torch.matmul(probs, v)                           # MED:300
        B, h, T, d = ctx.shape
        return ctx.permute(0, 2, 1, 3).reshape(B, T, h * d)


def install_reference_capture(model):
    """Swap every cross-attention of a pnp_ovss_b200.blip_itm.BlipITM for the torch restatement above."""
    for lyr in model.layer:
        lyr.crossattention.self = ReferenceCrossAttention(lyr.crossattention.self)
    return model


def compute_gradcam_ensemble_reference(model, visual_input, text_input, tokenized_text):
    """BITM:386-457, faithfully wasteful: all 12 blocks capture, full backward, 12x12 maps materialised."""
    for lyr in model.layer:
        lyr.crossattention.self.save_attention = True                     # BITM:388-392
    with torch.enable_grad():
        output = model(visual_input, text_input)                           # BITM:395
        loss = output[:, 1].sum()                                          # BITM:399
        model.zero_grad()
        loss.backward()                                                    # BITM:404
    P = model.patch_num
    blocklist = []
    with torch.no_grad():
        for lyr in model.layer:                                            # BITM:411-435
            xa = lyr.crossattention.self
            cams, grads = xa.get_attention_map(), xa.get_attn_gradients()
            gradcams = O.gradcam_from_capture(cams, grads, tokenized_text.attention_mask, P)
            blocklist.append([gradcams[:, h, 1:, :, :].cpu().detach().clone() for h in range(gradcams.shape[1])])
    for lyr in model.layer:
        lyr.crossattention.self.save_attention = False
        lyr.crossattention.self.attention_map = lyr.crossattention.self.attn_gradients = None
    return blocklist, [], output.detach()


# ----------------------------------------------------------------------------------------------- per-image CPU post-processing
def _post_one(args):
    (cm, threshold, gt_shape, guide, data_type, ids, mode, rescale) = args
    torch.set_num_threads(1)
    with np.errstate(all="ignore"):
        return O.image_to_labels(torch.from_numpy(cm), threshold, gt_shape, guide, data_type, list(ids), mode, rescale)


def reference_batch_confusion(model, imgs, captions, tokens, decode, class_lists, dataset_ids, gts, guides, *, drop_iter,
                              layer, head, threshold, data_type, mode, n_class, coco=False, pool=None, timings=None, device=None):
    """One batch of save_img_union_attention on the CPU (DRV:290-521).  Images' post-processing -- every (image, pass) job of
    the batch in ONE map call -- is fanned over `pool` (a multiprocessing pool) when given; the reference itself does it in
    one Python loop.  device: run the model pass on that CUDA device the reference's way (imgs .to(rank) at DRV:608, the
    12x12 GradCAM maps .cpu()-ed one by one at BITM:431-433); everything after it stays on the CPU like in the reference."""
    P = model.patch_num

    def gradcam_fn(x):
        if device is not None:
            tk = type(tokens)(tokens.input_ids.to(device), tokens.attention_mask.to(device))
            return compute_gradcam_ensemble_reference(model, x.to(device), captions, tk)[0][layer][head]
        return compute_gradcam_ensemble_reference(model, x, captions, tokens)[0][layer][head]

    import time
    t0 = time.perf_counter()
    g0, agg, chosen, _ = O.salience_dropout(gradcam_fn, imgs, drop_iter, P)
    t_model = time.perf_counter() - t0
    B = imgs.shape[0]
    passes = []
    if not coco or drop_iter < 3:
        passes.append((g0, True))
    if agg is not None:
        passes.append((agg, coco))
    jobs = []
    for gmaps, rescale in passes:
        for b in range(B):
            toks = O.token_strings(tokens.input_ids[b], decode)
            cm = O.mean_over_filtered_label_tokens(toks, gmaps[b], len(class_lists[b]))
            jobs.append((cm.numpy().copy(), threshold, tuple(gts[b].shape), guides[b], data_type, dataset_ids[b], mode, rescale))
    preds = pool.map(_post_one, jobs, chunksize=1) if pool is not None else [_post_one(j) for j in jobs]
    hists = [O.scores(gts, preds[i * B:(i + 1) * B], n_class)[1] for i in range(len(passes))]
    if timings is not None:
        timings["model_s"] = timings.get("model_s", 0.0) + t_model
        timings["post_s"] = timings.get("post_s", 0.0) + (time.perf_counter() - t0 - t_model)
    return hists


def make_pool(n_workers):
    """Spawned (not forked: the parent has live OpenMP threads) worker pool for the per-image post-processing."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    return ctx.Pool(n_workers, initializer=_pool_init)


def _pool_init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    torch.set_num_threads(1)
    from . import densecrf
    densecrf.lib()
