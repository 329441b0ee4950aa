"""Drop-in replacements for the Python callables of the reference's mask-extraction path: same names, same
arguments, same return types, computed by libpnp_ovss_b200.so on the current CUDA device.

    from pnp_ovss_b200.reference_api import (compute_gradcam_ensemble, Inference_BLIP_filteredcaption,
        Mean_over_filtered_label_tokens, postprocess, blurring, densecrf, Scale_0_1, _fast_hist, scores)
    from pnp_ovss_b200.reference_api import DenseCRF2D, unary_from_softmax      # pydensecrf surface

Inputs may live on the host (numpy / CPU tensors, as the reference passes them): they are copied to the GPU, the
kernels run there, and results come back in the container type the reference returns.  There is no CPU code path:
without a CUDA device or without the built library these functions raise.

DRV = PnP_OVSS_0514_updated_segmentation.py, DRVC = its _coco twin, BITM = blip_image_text_matching.py."""
import numpy as np
import torch

from . import host, ops
from ._lib import PnpError


def _device():
    if not torch.cuda.is_available():
        raise PnpError("pnp_ovss_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(x, dtype=torch.float32):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device=_device(), dtype=dtype, non_blocking=True).contiguous()


# ------------------------------------------------------------------------------------------------- a5: Scale_0_1
def Scale_0_1(AA):
    """DRV:1078-1094.  Per-channel min-max rescale, in place like the reference.  Computed on the GPU (host tensors
    are uploaded and the result is written back into AA); on the fused path it is folded into
    pnp_threshold_upsample(rescale=1), this entry point exists for callers that use it standalone and is tensor glue
    around device reductions, not a kernel of its own."""
    if AA.dim() == 2:
        return AA
    x = AA if AA.is_cuda else _to_dev(AA, AA.dtype)
    shape = x.shape
    flat = x.reshape(*shape[:-2], -1)
    flat = flat - flat.min(-1, keepdim=True)[0]
    flat = flat / flat.max(-1, keepdim=True)[0]
    AA.copy_(flat.view(shape))
    return AA


# ------------------------------------------------------------------------------------------------- a6: blurring
def blurring(att_resize, img_shape, scale=0.05):
    """DRV:1149-1153: gaussian_filter(sigma = scale*max(img_shape)) then min-max.  Returns a numpy [H,W] array."""
    x = _to_dev(att_resize)
    if x.dim() != 2:
        raise PnpError("blurring expects one [H,W] map")
    out, _ = ops.gaussian_blur(x.unsqueeze(0), scale * max(img_shape), normalize=True)
    return out[0].cpu().numpy()


# ------------------------------------------------------------------------------------------------- a8: pydensecrf surface
def unary_from_softmax(sm, scale=None, clip=1e-5):
    """pydensecrf.utils.unary_from_softmax: -log(clip(p)) as float32 [C, N] (numpy in, numpy out, computed on the GPU)."""
    x = _to_dev(np.asarray(sm), torch.float64 if np.asarray(sm).dtype == np.float64 else torch.float32)
    num_cls = x.shape[0]
    if scale is not None:
        x = scale * x + (1 - scale) / num_cls
    if clip is not None:
        x = x.clamp(clip, 1.0)
    return (-x.log()).reshape(num_cls, -1).to(torch.float32).cpu().numpy()


def _spatial_lattice(H, W, sxy, device):
    """The Gaussian-kernel lattice depends on (H, W, sxy) only: built once per shape (shared LRU of pipeline.py)."""
    from .pipeline import _SPATIAL
    return _SPATIAL.get(H, W, sxy, device)


class DenseCRF2D:
    """pydensecrf.densecrf.DenseCRF2D for the calls the reference makes (DRV:1066-1071):
    DenseCRF2D(w,h,c), setUnaryEnergy(float32 [c, w*h]), addPairwiseGaussian(sxy, compat),
    addPairwiseBilateral(sxy, srgb, rgbim uint8 [h,w,3], compat), inference(n) -> Q [c, w*h] (numpy)."""

    def __init__(self, w, h, c):
        self.w, self.h, self.c = int(w), int(h), int(c)
        self._dev = _device()
        self._unary = None
        self._lattices = []
        self._weights = []

    def setUnaryEnergy(self, U):
        U = np.asarray(U)
        if U.dtype != np.float32 or not U.flags.c_contiguous:
            raise ValueError("Buffer dtype mismatch / ndarray is not C-contiguous")
        if U.shape != (self.c, self.w * self.h):
            raise ValueError("Bad shape for unary energy (Need (%d, %d), got %s)" % (self.c, self.w * self.h, U.shape))
        self._unary = ops.crf_pack(_to_dev(U).unsqueeze(0))

    def addPairwiseGaussian(self, sxy, compat, kernel=None, normalization=None):
        self._lattices.append(_spatial_lattice(self.h, self.w, sxy, self._dev))
        self._weights.append(float(compat))

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat, kernel=None, normalization=None):
        rgbim = np.asarray(rgbim)
        if rgbim.dtype != np.uint8 or not rgbim.flags.c_contiguous or rgbim.shape != (self.h, self.w, 3):
            raise ValueError("Bad shape for pairwise bilateral (Need (%d, %d, 3) uint8 C-contiguous)" % (self.h, self.w))
        rgb = torch.from_numpy(rgbim).to(self._dev).unsqueeze(0).contiguous()
        self._lattices.append(ops.build_lattice(self.h, self.w, sxy, rgb=rgb, srgb=srgb))
        self._weights.append(float(compat))

    def inference(self, n):
        if self._unary is None:
            raise RuntimeError("setUnaryEnergy was not called")
        Q, _ = ops.crf_inference(self._lattices, self._weights, self._unary, self.c, int(n), want_labels=False)
        return ops.crf_unpack(Q, self.c)[0].cpu().numpy()


def densecrf(image, mask, n_iter=10, pos_w=7, pos_xy_std=3, bi_w=10, bi_xy_std=50, bi_rgb_std=5, return_q=False):
    """DRV:1030-1074.  image uint8 [H,W,3]; mask [C',H,W] (tensor or ndarray, raw channel scores).
    Returns the float32 [H,W] argmax map (and Q [C',H,W] when return_q), as numpy like the reference."""
    dev = _device()
    m = _to_dev(mask)
    C, H, W = m.shape
    image = np.ascontiguousarray(image)
    if image.dtype != np.uint8 or image.shape != (H, W, 3):
        raise ValueError("image must be uint8 [H,W,3] matching the mask")
    rgb = torch.from_numpy(image).to(dev).unsqueeze(0).contiguous()
    U = ops.crf_unary_from_maps(m.view(1, C, H * W))  # softmax over channels + unary_from_softmax, fused
    lat_s = _spatial_lattice(H, W, pos_xy_std, dev)
    lat_b = ops.build_lattice(H, W, bi_xy_std, rgb=rgb, srgb=bi_rgb_std)
    Q, labels = ops.crf_inference([lat_s, lat_b], [pos_w, bi_w], U, C, n_iter, want_labels=True)
    MAP = labels.view(H, W).to(torch.float32).cpu().numpy()
    if return_q:
        return MAP, ops.crf_unpack(Q, C)[0].view(C, H, W).cpu().numpy()
    return MAP


# ------------------------------------------------------------------------------------------------- postprocess
def postprocess(args, final_pred_wbackground, org_img_list, label_trues, img):
    """DRV:1002-1028: `args.postprocess` substring tests select blur, crf or blur+crf; returns a numpy label map."""
    mode = args.postprocess
    x = _to_dev(final_pred_wbackground)
    C, H, W = x.shape
    img_shape = (label_trues[img].shape[0], label_trues[img].shape[1])
    if "blur" in mode:
        x, _ = ops.gaussian_blur(x, 0.05 * max(img_shape), normalize=True)
    if "crf" in mode:
        return densecrf(org_img_list[img], x)
    if "blur" in mode:
        return ops.argmax_channels(x.view(1, C, H * W)).view(H, W).to(torch.int64).cpu().numpy()
    raise PnpError("postprocess mode %r selects neither blur nor crf" % (mode,))


# ------------------------------------------------------------------------------------------------- a10: confusion matrix
def _fast_hist(label_true, label_pred, n_class):
    """DRV:1106-1112 on flat arrays; returns an int64 [n,n] numpy array."""
    gt = _to_dev(np.asarray(label_true, dtype=np.float32).reshape(1, -1))
    pred = _to_dev(np.asarray(label_pred).reshape(1, -1), dtype=torch.int32)
    hist = torch.zeros((n_class, n_class), dtype=torch.int64, device=gt.device)
    bad = torch.zeros(1, dtype=torch.int32, device=gt.device)
    ops.confusion_accumulate(pred, gt, n_class, hist, bad_count=bad)
    if int(bad.item()):
        raise ValueError("predicted label outside [0, n_class) (np.bincount(...).reshape would fail in the reference)")
    return hist.cpu().numpy()


def scores(label_trues, label_preds, cats=None, n_class=None):
    """DRV:1115-1146.  Returns ({...metrics...}, hist float64 [n,n]); the histogram is accumulated on the GPU.  With `cats`
    (id -> name, as the drivers pass it) 'Class IoU' is the reference's dict keyed by class name; without it, the bare array."""
    dev = _device()
    hist = torch.zeros((n_class, n_class), dtype=torch.int64, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    for lt, lp in zip(label_trues, label_preds):
        gt = _to_dev(np.asarray(lt, dtype=np.float32).reshape(1, -1))
        pred = _to_dev(np.asarray(lp).reshape(1, -1), dtype=torch.int32)
        ops.confusion_accumulate(pred, gt, n_class, hist, bad_count=bad)
    if int(bad.item()):
        raise ValueError("predicted label outside [0, n_class)")
    table, h = metrics_from_hist(hist.cpu().numpy().astype(np.float64))
    if cats is not None:   # DRV:1130-1137: per-class IoU keyed by name, "Background" first, then cats[class_id]
        names = ["Background"] + [cats[int(i)] for i in range(1, n_class)]
        table["Class IoU"] = dict(zip(names, table["Class IoU"]))
    return table, h


def metrics_from_hist(hist):
    """The derived statistics of DRV:1121-1146 / Calculate_mIoU.py:221-256 (host arithmetic on the [n,n] matrix)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        valid = hist.sum(axis=1) > 0
        mean_iu = np.nanmean(iu[valid])
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return {"Pixel Accuracy": acc, "Mean Accuracy": acc_cls, "Frequency Weighted IoU": fwavacc, "Mean IoU": mean_iu,
            "Class IoU": iu}, hist


# ------------------------------------------------------------------------------------------------- a3: token merge
def Mean_over_filtered_label_tokens(model_textloc, txt_tokens_filtered, gradcam_filtered, class_filtered_list, img_num):
    """DRV:810-853.  gradcam_filtered [T-1,P,P] of image img_num -> [C,P,P] (on the GPU; .cpu() it if needed)."""
    tokenizer = model_textloc.module.tokenizer if hasattr(model_textloc, "module") else model_textloc.tokenizer
    toks = host.token_strings(txt_tokens_filtered.input_ids[img_num].tolist(), tokenizer.decode)
    n_classes = len(class_filtered_list[img_num])
    segs = host.build_token_segments(toks, n_classes)
    g = _to_dev(gradcam_filtered).unsqueeze(0)
    dev = g.device
    start = torch.tensor([[s[0] for s in segs]], dtype=torch.int32, device=dev)
    length = torch.tensor([[s[1] for s in segs]], dtype=torch.int32, device=dev)
    div = torch.tensor([[s[2] for s in segs]], dtype=torch.float32, device=dev)
    return ops.token_merge(g, start, length, div, row_offset=3)[0]


# ------------------------------------------------------------------------------------------------- a1/a2/a4
def compute_gradcam_ensemble(args, model, visual_input, text_input, tokenized_text, drop_iter=0):
    """BITM:386-457.  Returns (blocklist, [], itm_output) where blocklist[layer][head] -> [B,T-1,P,P].

    The reference builds all 12x12 maps and the drivers read exactly one, blocklist[max_att_block_num-1][prune_att_head]
    (DRV:572-574, 619-621).  Here that entry is computed eagerly by the fused softmax-backward/GradCAM kernel of the
    model's cross-attention (pnp_ovss_b200.blip_itm.BlipITM.gradcam); any other [layer][head] is computed on first
    access by one more model pass, so the full 12x12 surface stays readable without paying for it.  `model` may also
    be the reference's own LAVIS BlipITM object (see lavis_compat.gradcam_from_lavis_model)."""
    layer = int(args.max_att_block_num) - 1
    head = int(args.prune_att_head)
    if hasattr(model, "gradcam"):            # pnp_ovss_b200.blip_itm.BlipITM: trimmed pass, fused kernels (a)+(b)
        def compute(l, h):
            return model.gradcam(visual_input, text_input, tokenized_text, layer=l, head=h)
    else:                                    # a live LAVIS BlipITM (BITM:388-425 attribute protocol), kernel (a) patched in
        from . import lavis_compat
        P = int(int(args.img_size) / 16)

        def compute(l, h):
            return lavis_compat.gradcam_from_lavis_model(model, visual_input, text_input, tokenized_text, l, h, P,
                                                         fused=getattr(args, "fused_xattn", True))
    gradcam, output = compute(layer, head)
    xattn = _cross_attention_modules(model)
    return _LazyBlocklist(lambda l, h: compute(l, h)[0], {(layer, head): gradcam}, len(xattn),
                          getattr(xattn[0], "heads", None) or xattn[0].num_attention_heads), [], output


def _cross_attention_modules(model):
    from .lavis_compat import cross_attention_modules
    return cross_attention_modules(model)


class _LazyBlocklist:
    """blocklist[layer][head], materialised on demand."""

    def __init__(self, compute, cache, n_layers, n_heads):
        self._compute, self._cache, self._n_layers, self._n_heads = compute, cache, n_layers, n_heads

    def __len__(self):
        return self._n_layers

    def __getitem__(self, layer):
        if not -self._n_layers <= layer < self._n_layers:
            raise IndexError(layer)
        return _LazyHeadlist(self, layer % self._n_layers)


class _LazyHeadlist:
    def __init__(self, parent, layer):
        self._p, self._layer = parent, layer

    def __len__(self):
        return self._p._n_heads

    def __getitem__(self, head):
        if not -self._p._n_heads <= head < self._p._n_heads:
            raise IndexError(head)
        key = (self._layer, head % self._p._n_heads)
        if key not in self._p._cache:
            self._p._cache[key] = self._p._compute(*key)
        return self._p._cache[key]


def Inference_BLIP_filteredcaption(args, model_textloc, txt_tokens_filtered, imgs_in, norm_imgs, img_ids,
                                   caption_filtered_list, class_filtered_list, rank):
    """DRV:564-722.  Returns (gradcam_0 [B,T-1,P,P], gradcam_agg or None), both CUDA tensors."""
    from .pipeline import salience_dropout_loop
    model = model_textloc.module if hasattr(model_textloc, "module") else model_textloc
    imgs = _to_dev(imgs_in).clone()
    norm = _to_dev(norm_imgs).clone() if norm_imgs is not None else None
    tokens = txt_tokens_filtered

    def gradcam_fn(x):
        return compute_gradcam_ensemble(args, model, x, caption_filtered_list, tokens)[0][int(args.max_att_block_num) - 1][
            int(args.prune_att_head)]

    g0, agg, _ = salience_dropout_loop(gradcam_fn, imgs, norm, int(args.drop_iter), int(int(args.img_size) / 16))
    return g0, agg


# ---------------------------------------------------------------------------------------------- inputs of the path (host side)
def _gt_png(path):
    from PIL import Image
    return np.float32(Image.open(path))


def Load_GroundTruth(args, img_ids, coco_thing=None):
    """DRV:901-928 / DRVC:1095-1125: ground-truth label maps as float32 [H,W] arrays, one per image id.

    voc: VOCdevkit/VOC2012/SegmentationClass/<id>.png with the 255 border folded into background 0; psc: the
    59-class Context PNGs as they are; ade20k: ADE_val_<id padded to 8>.png; coco_stuff: stuff-164k PNGs with
    255 -> 0 and every other value + 1 (the reference's per-pixel Python loop, vectorised).  coco_object needs the
    pycocotools handle the reference passes (`coco_thing`) and paints instances in annotation order, first one wins."""
    import os
    dt = args.data_type
    out = []
    for img_id in img_ids:
        if dt == "voc":
            m = _gt_png(os.path.join("%s/VOCdevkit/VOC2012/SegmentationClass" % args.home_dir, img_id + ".png"))
            m[m == 255] = 0
        elif dt == "psc":
            m = _gt_png(os.path.join("%s/mmsegmentation/data/VOCdevkit/VOC2010/SegmentationClassContext/" % args.home_dir,
                                     img_id + ".png"))
        elif dt == "ade20k":
            m = _gt_png(os.path.join("%s/ADEChallengeData2016/annotations/validation/" % args.home_dir,
                                     "ADE_val_" + str(img_id).rjust(8, "0") + ".png"))
        elif dt == "coco_stuff":
            m = _gt_png(os.path.join("%s/coco_stuff164k/annotations/val2017/" % args.home_dir, "{:012d}".format(int(img_id)) + ".png"))
            m = np.where(m == 255, np.float32(0), m + np.float32(1)).astype(np.float32)
        elif dt == "coco_object":
            if coco_thing is None:
                raise ValueError("coco_object ground truth needs the pycocotools COCO handle (coco_thing)")
            info = coco_thing.loadImgs(coco_thing.getImgIds(imgIds=[int(img_id)]))[0]
            m = np.zeros((info["height"], info["width"]))
            for ann in coco_thing.loadAnns(coco_thing.getAnnIds(imgIds=info["id"], iscrowd=None)):
                m[np.logical_and(coco_thing.annToMask(ann), m == 0)] = ann["category_id"]
        else:
            raise ValueError("unknown data_type %r" % (dt,))
        out.append(m)
    return out


def load_OrgImage(args, img_ids, coco_thing=None):
    """DRV:930-955 / DRVC:1127-1138: the RGB uint8 [H,W,3] images that guide the dense CRF."""
    import os
    from PIL import Image
    dt = args.data_type
    out = []
    for img_id in img_ids:
        if dt in ("voc", "psc"):
            path = os.path.join("%s/VOCdevkit/VOC2012/JPEGImages/" % args.home_dir, img_id + ".jpg")
        elif dt == "ade20k":
            path = os.path.join("%s/ADEChallengeData2016/images/validation/" % args.home_dir,
                                "ADE_val_" + str(img_id).rjust(8, "0") + ".jpg")
        elif dt in ("coco_object", "coco_stuff"):
            if coco_thing is not None:
                name = coco_thing.loadImgs(coco_thing.getImgIds(imgIds=[int(img_id)]))[0]["file_name"]
            else:
                name = "{:012d}.jpg".format(int(img_id))      # val2017 file names are the zero-padded image id
            path = os.path.join("%s/coco/images/val2017/" % args.home_dir, name)
        else:
            raise ValueError("unknown data_type %r" % (dt,))
        out.append(np.asarray(Image.open(path).convert("RGB")))
    return out


_GPT4O_CACHE = {}


def Load_predicted_classes(args, nms, *rest):
    """DRV:726-787 and its COCO twin DRVC:855-963 (which takes `cats` after `nms`): appends image `img`'s GPT-4o
    classes (probability > 70) to the three running lists and returns them.

    VOC/Context/ADE driver:  (args, nms, best_class_idx_list, class_filtered_list, caption_filtered_list,
                              gt_class_name_list, img_ids, img, pred_path)
    COCO driver:             (args, nms, cats, best_class_idx_list, class_filtered_list, caption_filtered_list,
                              gt_class_name_list, img_ids, img, pred_path)
    The answers are read from {home_dir}/GPT4o_classification/<data_type>_classification_noboundary.json (parsed once
    per file, not once per image like the reference)."""
    import json
    coco = args.data_type in ("coco_object", "coco_stuff")
    if coco:
        cats, best_list, class_list, caption_list, _gt_names, img_ids, img = rest[:7]
    else:
        best_list, class_list, caption_list, _gt_names, img_ids, img = rest[:6]
    if args.data_type not in ("voc", "psc", "ade20k", "coco_object", "coco_stuff"):
        raise ValueError("unknown data_type %r" % (args.data_type,))
    path = "%s/GPT4o_classification/%s_classification_noboundary.json" % (args.home_dir, args.data_type)
    if path not in _GPT4O_CACHE:
        with open(path, "r") as f:
            _GPT4O_CACHE[path] = json.load(f)
    answers = _GPT4O_CACHE[path]
    if coco:
        answer = answers[str(int(img_ids[img])).rjust(12, "0")]
        best, cls, caption = host.parse_gpt4o_classes_coco(answer, [c["id"] for c in cats], nms, args.data_type)
    else:
        key = "ADE_val_" + img_ids[img].rjust(8, "0") if args.data_type == "ade20k" else img_ids[img]
        best, cls, caption = host.parse_gpt4o_classes(answers[key], nms)
    best_list.append(best)
    class_list.append(cls)
    caption_list.append(caption)
    return best_list, class_list, caption_list


def save_img_union_attention(*a, max_block_num=None, cam_type="gradcam"):
    """One batch of the reference's evaluation loop, DRV:290-521, and its COCO twin DRVC:338-640 (same positional
    arguments after a leading pycocotools handle `coco_thing`):

        (model_textloc, imgs_in, org_img_sizes, args, gt_class_name_list, img_ids, drop_iter, norm_imgs, img_shape, cats,
         nms, txt_tokens, rank, att_head, max_block_num=None, cam_type="gradcam")

    Reads the guide images, ground truth and GPT-4o class lists the way the reference does (load_OrgImage,
    Load_GroundTruth, Load_predicted_classes), tokenises the filtered captions, then runs the whole batch on the GPU
    (pipeline.batch_confusion: DropOut rounds through compute_gradcam_ensemble, token merge, threshold/upsample, blur,
    CRF, argmax + relabel, confusion matrix) and writes the matrices where the reference writes them
    (`hist_withfiltered_caption/`, `all_drop_hist_with_filtered_caption/`, float64 .npy named after the first image id).
    Visualisation side outputs (getAttMap JPEGs, Draw_Segmentation_map) are not produced.  Returns None like the
    reference; the two matrices are also kept on `save_img_union_attention.last` as int64 CUDA tensors."""
    from . import pipeline
    a = list(a)
    coco_thing = None
    if not (hasattr(a[0], "module") or hasattr(a[0], "parameters")):      # the COCO driver passes coco_thing first
        coco_thing = a.pop(0)
    if len(a) > 14 and max_block_num is None:
        max_block_num = a[14]
    (model_textloc, imgs_in, _org_img_sizes, args, gt_class_name_list, img_ids, _drop_iter, norm_imgs, _img_shape, cats, nms,
     _txt_tokens, _rank, att_head) = a[:14]
    model = model_textloc.module if hasattr(model_textloc, "module") else model_textloc
    coco = args.data_type in ("coco_object", "coco_stuff")
    org_img_list = load_OrgImage(args, img_ids, coco_thing)
    label_trues = Load_GroundTruth(args, img_ids, coco_thing)
    best_class_idx_list, class_filtered_list, caption_filtered_list = [], [], []
    for img in range(len(img_ids)):
        if coco:
            Load_predicted_classes(args, nms, cats, best_class_idx_list, class_filtered_list, caption_filtered_list,
                                   gt_class_name_list, img_ids, img, None)
        else:
            Load_predicted_classes(args, nms, best_class_idx_list, class_filtered_list, caption_filtered_list,
                                   gt_class_name_list, img_ids, img, None)
    dev = _device()
    tokens = model.tokenizer(caption_filtered_list, padding="max_length", max_length=500, return_tensors="pt").to(dev)  # DRV:317-319
    layer, head = int(args.max_att_block_num) - 1, int(args.prune_att_head)

    def gradcam_fn(x):
        return compute_gradcam_ensemble(args, model, x, caption_filtered_list, tokens)[0][layer][head]

    if coco:
        dataset_ids = [[int(cats[i]["id"]) for i in best] for best in best_class_idx_list]       # DRVC:549-556
        n_class = 91 if args.data_type == "coco_object" else 183                                   # DRVC:597-600
    else:
        dataset_ids = [[i + 1 for i in best] for best in best_class_idx_list]                      # DRV:468-480
        n_class = len(cats) + 1                                                                    # DRV:496
    imgs = _to_dev(imgs_in).clone()
    norm = _to_dev(norm_imgs).clone() if isinstance(norm_imgs, torch.Tensor) else None
    hist0, hist_all, _ = pipeline.batch_confusion(
        gradcam_fn, imgs, tokens.input_ids.tolist(), model.tokenizer.decode, class_filtered_list, dataset_ids, label_trues,
        org_img_list, drop_iter=int(args.drop_iter), patch_num=int(int(args.img_size) / 16), threshold=float(args.threshold),
        data_type=args.data_type, mode=args.postprocess, n_class=n_class, coco=coco, norm_imgs=norm)
    if hist0 is not None:
        print(img_ids, "miou filtered_caption", metrics_from_hist(hist0.cpu().numpy().astype(np.float64))[0]["Mean IoU"])
        pipeline.save_hist_npy(hist0, args.save_path, "hist_withfiltered_caption", img_ids[0], max_block_num, att_head)
    if hist_all is not None:
        print(img_ids, "miou all_drop_with_filtered caption",
              metrics_from_hist(hist_all.cpu().numpy().astype(np.float64))[0]["Mean IoU"])
        pipeline.save_hist_npy(hist_all, args.save_path, "all_drop_hist_with_filtered_caption", img_ids[0], max_block_num, att_head)
    save_img_union_attention.last = (hist0, hist_all)
    return None
