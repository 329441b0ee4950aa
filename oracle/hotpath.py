"""oracle/hotpath.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (torch-CPU / numpy / scipy) of rows a1-a7, a9, a10 of SURVEY.md section 8a.  Every
function cites the reference lines it follows.  Shorthands:
  DRV  = /root/reference/PnP_OVSS_0514_updated_segmentation.py
  DRVC = /root/reference/PnP_OVSS_0514_updated_segmentation_coco.py
  BITM = /root/reference/Files to replace for BLIP/blip_image_text_matching.py
  MED  = /root/reference/Files to replace for BLIP/med.py
Quirks of the reference are reproduced on purpose (see DESIGN.md "Quirks kept").
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy.ndimage import gaussian_filter

SEP_ID = 102  # hard-coded at DRV:814


# ----------------------------------------------------------------------------------------------
# a1  cross-attention softmax + capture (MED:267-283)
# ----------------------------------------------------------------------------------------------
def cross_attention_probs(scores, attention_mask=None, head_size=64):
    """scores [B,h,T,K] raw QK^T -> probs = softmax(scores/sqrt(head_size) + mask, -1)  (MED:267-274)."""
    s = scores / math.sqrt(head_size)
    if attention_mask is not None:
        s = s + attention_mask
    return torch.nn.Softmax(dim=-1)(s)


def softmax_backward(probs, dprobs, head_size=64):
    """Autograd of MED:267-274: dL/dscores given dL/dprobs (what torch autograd computes)."""
    inner = (probs * dprobs).sum(-1, keepdim=True)
    return probs * (dprobs - inner) / math.sqrt(head_size)


# ----------------------------------------------------------------------------------------------
# a2  GradCAM from captured probs and their gradient (BITM:415-433)
# ----------------------------------------------------------------------------------------------
def gradcam_from_capture(cams, grads, attention_mask_500, patch_num):
    """cams, grads [B,12,T,K]; attention_mask_500 [B,500] (the max_length-padded tokens, DRV:608-610).

    Returns gradcams [B,12,T,P,P] exactly as BITM:427-429; the caller slices [:, head, 1:] (BITM:431-433)."""
    B = cams.shape[0]
    nh = cams.shape[1]
    mask = attention_mask_500.view(attention_mask_500.size(0), 1, -1, 1, 1)
    gradcams = cams[:, :, :, 1:].reshape(B, nh, -1, patch_num, patch_num) * grads[:, :, :, 1:].clamp(0).reshape(
        B, nh, -1, patch_num, patch_num) * mask[:, :, :cams.shape[2], :, :]
    gradcams[gradcams < 0] = 0
    return gradcams


def gradcam_head(cams, grads, attention_mask_500, patch_num, head):
    """The one entry the driver reads: blocklist[layer][head] -> [B,T-1,P,P]  (BITM:431-433, DRV:572-574)."""
    return gradcam_from_capture(cams, grads, attention_mask_500, patch_num)[:, head, 1:, :, :].detach().clone()


# ----------------------------------------------------------------------------------------------
# a3  token -> class merge (DRV:810-853; inline twin DRV:656-701)
# ----------------------------------------------------------------------------------------------
def token_strings(input_ids_row, decode):
    """DRV:811-818: decode tokens after position 0 up to (excluding) SEP=102, then drop 'a picture of'."""
    out = []
    for token_id in input_ids_row[1:]:
        word = decode([int(token_id)])
        if int(token_id) == SEP_ID:
            break
        out.append(word)
    return out[3:]


def mean_over_filtered_label_tokens(token_strs, gradcam_filtered, n_classes):
    """token_strs: list[str] from token_strings(); gradcam_filtered [T-1,P,P]; returns [C,P,P] (DRV:819-853)."""
    special = '##'
    g = gradcam_filtered[3:-1]
    L = token_strs
    if len(L) != n_classes:
        out = torch.zeros((n_classes, g.shape[1], g.shape[2]), dtype=g.dtype)
        ind_token = 0
        ind_classes = 0
        word_length = 1
        while ind_token < len(L):
            if not L[ind_token].startswith(special):
                out[ind_classes, :, :] = g[ind_token, :, :].detach().clone()
                if ind_token + 1 < len(L) and not L[ind_token + 1].startswith(special):
                    ind_classes += 1
                ind_token += 1
                word_length = 1
            else:
                word_length += 1
                out[ind_classes, :, :] = out[ind_classes, :, :].detach().clone() + g[ind_token, :, :].detach().clone()
                if ind_token + 1 < len(L) and not L[ind_token + 1].startswith(special):
                    out[ind_classes, :, :] /= word_length
                    ind_classes += 1
                ind_token += 1
        return out
    return g[:n_classes]


# ----------------------------------------------------------------------------------------------
# a4  Salience DropOut loop (DRV:564-722)
# ----------------------------------------------------------------------------------------------
def salience_dropout(gradcam_fn, imgs_in, drop_iter, patch_num, save_len=10, argsort_kind=None):
    """gradcam_fn(imgs [B,3,S,S]) -> [B,T-1,P,P] stands for compute_gradcam_ensemble(...)[layer][head].

    Returns (gradcam_0, gradcam_agg_or_None, chosen: list[list[int]], imgs_dropped_per_round).
    argsort_kind=None uses numpy's default like DRV:646; 'stable' fixes the tie order."""
    if drop_iter == 1:  # DRV:565-575
        g = gradcam_fn(imgs_in)
        return g.detach().clone(), None, [[] for _ in range(imgs_in.shape[0])], [imgs_in]
    imgs = imgs_in.detach().clone()  # DRV:578
    B = imgs.shape[0]
    chosen = [[] for _ in range(B)]  # DRV:580-582
    ensemble = []
    dropped_inputs = []
    for _ in range(drop_iter):
        for b in range(B):  # DRV:589-603
            for max_patch in chosen[b]:
                mx = (max_patch // patch_num) * 16
                my = (max_patch % patch_num) * 16
                imgs[b, :, mx:mx + 16, my:my + 16] = 0
        dropped_inputs.append(imgs.detach().clone())
        g = gradcam_fn(imgs).detach().clone()  # DRV:611-621
        g_pred = g.detach().clone()  # DRV:623-635
        for b in range(B):
            for max_patch in chosen[b]:
                g_pred[b][:, max_patch // patch_num, max_patch % patch_num] = 0
        ensemble.append(g_pred)
        for b in range(B):  # DRV:638-647
            sum_cam = g[b][3:-1, :, :].sum(dim=0)
            sort_union = sum_cam.flatten().cpu().numpy().copy()
            for idx in chosen[b]:
                sort_union[idx] = 0
            if argsort_kind is None:
                top = np.argsort(sort_union.flatten())[-save_len:]
            else:
                top = np.argsort(sort_union.flatten(), kind=argsort_kind)[-save_len:]
            chosen[b].extend(int(t) for t in top)
    g0 = ensemble[0].detach().clone()  # DRV:716-721 (m0 is counted twice)
    agg = ensemble[0].detach().clone()
    for r in range(drop_iter):
        agg += ensemble[r].detach().clone()
    return g0, agg, chosen, dropped_inputs


# ----------------------------------------------------------------------------------------------
# a5 + a7  threshold, upsample, rescale, background rule (DRV:348-380, 424-455; DRVC:512-569)
# ----------------------------------------------------------------------------------------------
def scale_0_1(AA):
    """DRV:1078-1094 (in place for 3-D/4-D)."""
    if len(AA.shape) == 2:
        return AA
    elif len(AA.shape) == 3:
        c, h, w = AA.shape
        AA = AA.view(AA.size(0), -1)
        AA -= AA.min(-1, keepdim=True)[0]
        AA /= AA.max(-1, keepdim=True)[0]
        AA = AA.view(c, h, w)
    elif len(AA.shape) == 4:
        b, c, h, w = AA.shape
        AA = AA.view(AA.size(0), AA.size(1), -1)
        AA -= AA.min(-1, keepdim=True)[0]
        AA /= AA.max(-1, keepdim=True)[0]
        AA = AA.view(b, c, h, w)
    return AA


def add_background_rule(data_type, n_classes):
    """a7: does this image get a background channel?  DRV:373-379/449-455, DRVC:538-541/566-569."""
    if data_type in ("voc", "coco_object"):
        return True
    if data_type in ("psc", "ade20k", "coco_stuff"):
        return n_classes < 3
    raise ValueError(data_type)


def threshold_upsample(pred_map, threshold, out_hw, rescale, with_background):
    """pred_map [C,P,P] -> [C',H,W] float32.  DRV:349-379 (rescale=True) / DRV:425-455 (rescale=False)."""
    th = pred_map.clone().detach()
    for i in range(pred_map.shape[0]):
        th[i] = (pred_map[i] - pred_map[i].min()) / (pred_map[i].max() - pred_map[i].min())
    th = (th >= threshold).type(torch.bool)
    pred = pred_map * th
    pred = F.interpolate(pred.unsqueeze(0), size=(int(out_hw[0]), int(out_hw[1])), mode='bilinear',
                         align_corners=True).squeeze()
    if rescale:
        pred = scale_0_1(pred)
    if len(pred.shape) < 3:
        max_map = pred
        pred = pred.unsqueeze(0)
    else:
        max_map = torch.max(pred, dim=0)[0]
    background = (max_map == 0).unsqueeze(0)
    if with_background:
        return torch.cat((background, pred), dim=0)
    return pred


# ----------------------------------------------------------------------------------------------
# a6  Gaussian blur + min-max (DRV:1149-1153, called from DRV:1005-1011)
# ----------------------------------------------------------------------------------------------
def blurring(att_resize, img_shape, scale=0.05):
    att_resize = gaussian_filter(np.asarray(att_resize), scale * max(img_shape))
    att_resize = att_resize - att_resize.min()
    att_resize = att_resize / att_resize.max()
    return att_resize


def blur_channels(final_pred_wbackground, img_shape, scale=0.05):
    """DRV:1005-1011: per-channel blurring, stacked."""
    outs = []
    for i in range(final_pred_wbackground.shape[0]):
        outs.append(torch.from_numpy(blurring(final_pred_wbackground[i], img_shape, scale=scale)))
    return torch.stack(outs, axis=0)


def postprocess(mode, final_pred_wbackground, org_img, img_shape, crf_fn=None):
    """DRV:1002-1028.  crf_fn(image, mask) defaults to the C restatement in oracle.densecrf."""
    if crf_fn is None:
        from .densecrf import densecrf as crf_fn
    if "blur" in mode and "crf" in mode:
        x = blur_channels(final_pred_wbackground, img_shape)
        return crf_fn(org_img, x)
    elif "crf" in mode:
        return crf_fn(org_img, final_pred_wbackground)
    elif "blur" in mode:
        x = blur_channels(final_pred_wbackground, img_shape)
        return torch.argmax(x, dim=0).numpy()
    raise ValueError(mode)


# ----------------------------------------------------------------------------------------------
# a9  relabel local -> dataset ids, sequential and aliasing (DRV:390-399, 468-480; DRVC:549-584)
# ----------------------------------------------------------------------------------------------
def relabel_sequential(argmax_map, dataset_ids, with_background):
    """dataset_ids[i] is the id written for local class i (e.g. best_class_idx[i]+1 for VOC, DRV:392).

    In place, from the last class down, exactly like the reference loop (aliasing included)."""
    shift = 1 if with_background else 0
    for i in range(len(dataset_ids) - 1, -1, -1):
        argmax_map[argmax_map == int(i + shift)] = dataset_ids[i]
    return argmax_map


# ----------------------------------------------------------------------------------------------
# a10  confusion matrix and scores (DRV:1106-1146)
# ----------------------------------------------------------------------------------------------
def fast_hist(label_true, label_pred, n_class):
    mask = (label_true >= 0) & (label_true < n_class)
    hist = np.bincount(n_class * label_true[mask].astype(int) + label_pred[mask].astype(int),
                       minlength=n_class ** 2).reshape(n_class, n_class)
    return hist


def scores(label_trues, label_preds, n_class):
    hist = np.zeros((n_class, n_class))
    for lt, lp in zip(label_trues, label_preds):
        hist += fast_hist(lt.flatten(), lp.flatten(), n_class)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        valid = hist.sum(axis=1) > 0
        mean_iu = np.nanmean(iu[valid])
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return {"Pixel Accuracy": acc, "Mean Accuracy": acc_cls, "Frequency Weighted IoU": fwavacc,
            "Mean IoU": mean_iu, "Class IoU": iu}, hist


# ----------------------------------------------------------------------------------------------
# composition of one image's post-processing, as save_img_union_attention does it (DRV:424-481)
# ----------------------------------------------------------------------------------------------
def image_to_labels(class_maps, threshold, gt_hw, org_img, data_type, dataset_ids, mode, rescale, crf_fn=None):
    """class_maps [C,P,P] (after a3) -> relabelled [H,W] map."""
    C = class_maps.shape[0]
    with_bg = add_background_rule(data_type, C)
    x = threshold_upsample(class_maps, threshold, gt_hw, rescale, with_bg)
    if mode:
        lab = postprocess(mode, x, org_img, gt_hw, crf_fn)
    else:
        lab = torch.argmax(x, dim=0).numpy()
    lab = np.array(lab)
    return relabel_sequential(lab, dataset_ids, with_bg)


# ----------------------------------------------------------------------------------------------
# composition of one batch, as save_img_union_attention does it (DRV:290-521 / DRVC:337-640)
# ----------------------------------------------------------------------------------------------
def batch_confusion(gradcam_fn, imgs_in, token_ids, decode, class_lists, dataset_ids, gts, guides, *, drop_iter,
                    patch_num, threshold, data_type, mode, n_class, coco=False, crf_fn=None, argsort_kind=None):
    """Returns (hist_round0 or None, hist_all_drop or None) as float64 [n,n] like DRV:495-520.

    token_ids [B,>=T] input ids (row b of the max_length-padded tokens); dataset_ids[b][i] = id written for
    local class i.  `coco` selects the COCO driver's deltas: Scale_0_1 also on the N-round path (DRVC:527) and
    the round-0 pass only when drop_iter < 3 (DRVC:420, 602)."""
    g0, agg, chosen, _ = salience_dropout(gradcam_fn, imgs_in, drop_iter, patch_num, argsort_kind=argsort_kind)
    B = imgs_in.shape[0]

    def labels_from(gmaps, rescale):
        preds = []
        for b in range(B):
            toks = token_strings(token_ids[b], decode)
            cm = mean_over_filtered_label_tokens(toks, gmaps[b], len(class_lists[b]))
            preds.append(image_to_labels(cm, threshold, gts[b].shape, guides[b], data_type, dataset_ids[b], mode,
                                         rescale, crf_fn))
        return preds

    hist0 = hist_agg = None
    if not coco or drop_iter < 3:
        _, hist0 = scores(gts, labels_from(g0, True), n_class)
    if agg is not None:
        _, hist_agg = scores(gts, labels_from(agg, coco), n_class)
    return hist0, hist_agg, chosen
