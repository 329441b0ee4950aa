"""GPU: rows a1+a2 through a real (tiny) model pass, and the reference-named entry points of reference_api, against
the CPU restatement of the reference's own procedure (oracle/reference_arm.py: torch softmax + hooks + full backward)."""
import copy

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _tiny_model(tok):
    from pnp_ovss_b200.blip_itm import BlipITM
    torch.manual_seed(11)
    m = BlipITM(img_size=96, tokenizer=tok, vocab=30524, hidden=128, layers=3, heads=2, inter=256, vit_dim=64, vit_depth=2,
                vit_heads=2, max_pos=64).eval()
    with torch.no_grad():  # random init at std 0.02 gives near-uniform attention; widen it so the maps have structure
        for p in m.parameters():
            p.mul_(2.0)
    return m


class _Args:
    max_att_block_num, prune_att_head, drop_iter, img_size = 2, 1, 3, 96


def test_gradcam_through_model_matches_reference_procedure(dev):
    from oracle import reference_arm as RA
    from pnp_ovss_b200 import reference_api as R
    tok = synth.SyntheticWordPieceTokenizer()
    caps = ["A picture of cat aeroplane", "A picture of dog"]
    tokens = tok(caps, padding="max_length", max_length=500)
    g = torch.Generator().manual_seed(3)
    imgs = torch.randn(2, 3, 96, 96, generator=g)
    model = _tiny_model(tok)
    ref_model = RA.install_reference_capture(copy.deepcopy(model))
    blocks, _, out_ref = RA.compute_gradcam_ensemble_reference(ref_model, imgs, caps, tokens)
    want = blocks[1][1]                                           # [layer][head] as DRV:572-574 reads it
    gm = model.to(dev)
    for full in (False, True):
        if full:
            gm.requires_grad_(True)
        else:
            gm.requires_grad_(False)
        cam, out = gm.gradcam(imgs.to(dev), caps, tokens.to(dev), layer=1, head=1, full_backward=full)
        assert cam.shape == want.shape
        scale = float(want.abs().max())
        assert scale > 0
        err = (cam.cpu() - want).abs().max().item()
        assert err <= 1e-3 * scale, "GradCAM differs: %g (scale %g, full=%s)" % (err, scale, full)
        assert torch.allclose(out.cpu(), out_ref, rtol=1e-3, atol=1e-5)
    # the reference-named entry point, lazily materialising only [layer][head]
    gm.requires_grad_(False)
    lst, empty, out = R.compute_gradcam_ensemble(_Args, gm, imgs.to(dev), caps, tokens.to(dev))
    assert empty == [] and torch.allclose(lst[1][1].cpu(), want, atol=1e-3 * scale)
    # any other [layer][head] of the reference's 12x12 surface is computed on first access
    assert len(lst) == 3 and len(lst[0]) == 2
    other = lst[0][0]
    assert other.shape == want.shape and torch.allclose(other.cpu(), blocks[0][0], atol=1e-3 * float(blocks[0][0].abs().max()))
    with pytest.raises(IndexError):
        lst[5]


def test_inference_blip_filteredcaption_matches_oracle_loop(dev):
    """Inference_BLIP_filteredcaption (DRV:564-722) with the tiny model vs the oracle's DropOut loop around the
    reference-procedure GradCAM (same weights, CPU)."""
    from oracle import hotpath as O
    from oracle import reference_arm as RA
    from pnp_ovss_b200 import reference_api as R
    tok = synth.SyntheticWordPieceTokenizer()
    caps = ["A picture of cat aeroplane", "A picture of dog"]
    tokens = tok(caps, padding="max_length", max_length=500)
    g = torch.Generator().manual_seed(5)
    imgs = torch.randn(2, 3, 96, 96, generator=g)
    model = _tiny_model(tok)
    ref_model = RA.install_reference_capture(copy.deepcopy(model))
    g0_o, agg_o, chosen_o, _ = O.salience_dropout(
        lambda x: RA.compute_gradcam_ensemble_reference(ref_model, x, caps, tokens)[0][1][1], imgs, 3, 6, argsort_kind="stable")
    gm = model.to(dev).requires_grad_(False)

    class Wrap:  # the drivers pass a DDP-wrapped model (model_textloc.module)
        module = gm

    g0, agg = R.Inference_BLIP_filteredcaption(_Args, Wrap, tokens.to(dev), imgs, None, ["a", "b"], caps, [["cat", "aeroplane"], ["dog"]], 0)
    scale = float(agg_o.abs().max())
    assert (g0.cpu() - g0_o).abs().max().item() <= 1e-3 * float(g0_o.abs().max())
    assert (agg.cpu() - agg_o).abs().max().item() <= 1e-3 * scale
    # token merge entry point on the accumulated map of image 0
    cm = R.Mean_over_filtered_label_tokens(Wrap, tokens, agg[0], [["cat", "aeroplane"], ["dog"]], 0)
    toks = O.token_strings(tokens.input_ids[0], tok.decode)
    want = O.mean_over_filtered_label_tokens(toks, agg[0].cpu(), 2)
    assert torch.equal(cm.cpu(), want)


def test_live_lavis_protocol_model_is_a_drop_in(dev):
    """compute_gradcam_ensemble / Inference_BLIP_filteredcaption handed a LAVIS-shaped model object (the reference's
    module tree, call form and capture protocol; tests/lavis_shaped_model.py) instead of the native BlipITM: the fused
    softmax is patched into the requested block for the call (fused_xattn=True) or the model's own hooks are read on
    the device (False).  Oracle: the CPU reference procedure on the same weights."""
    from lavis_shaped_model import LavisShapedBlipITM
    from oracle import hotpath as O
    from oracle import reference_arm as RA
    from pnp_ovss_b200 import reference_api as R
    from pnp_ovss_b200.blip_itm import BlipITM
    from pnp_ovss_b200.lavis_compat import cross_attention_modules
    tok = synth.SyntheticWordPieceTokenizer()
    caps = ["A picture of cat aeroplane", "A picture of dog"]
    tokens = tok(caps, padding="max_length", max_length=500)
    imgs = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(3))
    dims = dict(hidden=128, layers=3, heads=2, inter=256, vit_dim=64, vit_depth=2, vit_heads=2, max_pos=64)
    torch.manual_seed(21)
    lavis = LavisShapedBlipITM(tok, img_size=96, **dims).eval()
    native = BlipITM(img_size=96, tokenizer=tok, vocab=30524, **dims).eval()
    native.load_lavis_checkpoint(lavis.state_dict())
    ref_model = RA.install_reference_capture(copy.deepcopy(native))
    blocks, _, out_ref = RA.compute_gradcam_ensemble_reference(ref_model, imgs, caps, tokens)
    want, scale = blocks[1][1], float(blocks[1][1].abs().max())
    assert scale > 0
    lavis = lavis.to(dev)

    class A(_Args):
        fused_xattn = True

    for fused in (True, False):
        A.fused_xattn = fused
        lst, empty, out = R.compute_gradcam_ensemble(A, lavis, imgs.to(dev), caps, tokens.to(dev))
        assert empty == [] and len(lst) == 3 and len(lst[0]) == 2
        assert lst[1][1].shape == want.shape
        assert (lst[1][1].cpu() - want).abs().max().item() <= 1e-3 * scale, fused
        assert torch.allclose(out.cpu(), out_ref, rtol=1e-3, atol=1e-5)
        other = lst[2][0]
        assert (other.cpu() - blocks[2][0]).abs().max().item() <= 1e-3 * float(blocks[2][0].abs().max())
        for xa in cross_attention_modules(lavis):            # the model is handed back untouched
            assert "forward" not in xa.__dict__ and xa.save_attention is False and xa.attention_map is None
    # the whole DropOut loop of DRV:564-722 around it
    g0_o, agg_o, _, _ = O.salience_dropout(
        lambda x: RA.compute_gradcam_ensemble_reference(ref_model, x, caps, tokens)[0][1][1], imgs, 3, 6, argsort_kind="stable")
    A.fused_xattn = True
    g0, agg = R.Inference_BLIP_filteredcaption(A, lavis, tokens.to(dev), imgs, None, ["a", "b"], caps, [["cat", "aeroplane"], ["dog"]], 0)
    assert (g0.cpu() - g0_o).abs().max().item() <= 1e-3 * float(g0_o.abs().max())
    assert (agg.cpu() - agg_o).abs().max().item() <= 1e-3 * float(agg_o.abs().max())


def test_save_img_union_attention_reads_the_reference_layouts(dev, tmp_path):
    """reference_api.save_img_union_attention (DRV:290-521) on a miniature VOC tree: JPEG/PNG inputs and the GPT-4o
    answers are read from disk, the batch (three different ground-truth sizes) runs on the GPU, and the .npy matrices
    land where Calculate_mIoU.py looks for them.  Oracle: oracle.hotpath.batch_confusion on the same inputs."""
    import json
    import types
    from PIL import Image
    from oracle import hotpath as O
    from pnp_ovss_b200 import reference_api as R
    from pnp_ovss_b200.driver import DATASETS
    nms = DATASETS["voc"][0]
    cats = [{"id": i + 1, "name": n} for i, n in enumerate(nms)]
    img_ids = ["2007_000033", "2007_000042", "2007_000061"]
    sizes = [(40, 56), (64, 48), (50, 50)]
    answers = ["[15: 'person', 12: 'dog'], [95%, 80%]", "[1: 'aeroplane', 8: 'cat', 3: 'bird'],\n[90%, 60%, 99%]", "[19: 'train'], [75%]"]
    home = tmp_path / "home"
    for sub in ("VOCdevkit/VOC2012/SegmentationClass", "VOCdevkit/VOC2012/JPEGImages", "GPT4o_classification"):
        (home / sub).mkdir(parents=True)
    gts, guides = [], []
    for k, (img_id, (H, W)) in enumerate(zip(img_ids, sizes)):
        gt = synth.gt_labels(300 + k, H, W, 21).astype(np.uint8)
        gd = synth.guide_image(400 + k, H, W)
        Image.fromarray(gt).save(home / "VOCdevkit/VOC2012/SegmentationClass" / (img_id + ".png"))
        Image.fromarray(gd).save(home / "VOCdevkit/VOC2012/JPEGImages" / (img_id + ".png"))
        (home / "VOCdevkit/VOC2012/JPEGImages" / (img_id + ".png")).rename(home / "VOCdevkit/VOC2012/JPEGImages" / (img_id + ".jpg"))
        g = gt.astype(np.float32)
        g[g == 255] = 0
        gts.append(g)
        guides.append(gd)
    json.dump(dict(zip(img_ids, answers)), open(home / "GPT4o_classification/voc_classification_noboundary.json", "w"))
    tok = synth.SyntheticWordPieceTokenizer()
    class_lists = [["person", "dog"], ["aeroplane", "bird"], ["train"]]
    ids = [[15, 12], [1, 3], [19]]
    caps = ["A picture of " + " ".join(c) for c in class_lists]
    tokens = tok(caps, padding="max_length", max_length=500)
    T = max(len(tok.encode(c)) for c in caps)
    rows = tokens.attention_mask[:, 1:T].float()
    S, P = 96, 6
    imgs = torch.randn(3, 3, S, S, generator=torch.Generator().manual_seed(9)).abs()

    class FakeModel:   # stands where BlipITM stands: .tokenizer and .gradcam(...) (the SynthGradcamFn saliency)
        tokenizer = tok
        layer = [types.SimpleNamespace(crossattention=types.SimpleNamespace(self=types.SimpleNamespace(heads=12)))] * 12

        def __init__(self):
            self.fn = synth.SynthGradcamFn(31, 3, T, P)

        def gradcam(self, x, text_input, tokenized_text, layer, head):
            assert list(text_input) == caps and (layer, head) == (7, 9)
            return self.fn(x.cpu(), rows).to(x.device), None

    wrap = types.SimpleNamespace(module=FakeModel())
    args = types.SimpleNamespace(home_dir=str(home), data_type="voc", save_path=str(tmp_path / "out"), drop_iter=3, img_size=S,
                                 max_att_block_num=8, prune_att_head=9, threshold=0.15, postprocess="blur")
    out = R.save_img_union_attention(wrap, imgs, None, args, [None] * 3, img_ids, 3, None, None, cats, nms, None, 0, 9, max_block_num=8)
    assert out is None
    fn = synth.SynthGradcamFn(31, 3, T, P)
    with np.errstate(all="ignore"):
        h0, hagg, _ = O.batch_confusion(lambda x: fn(x, rows), imgs.clone(), tokens.input_ids, tok.decode, class_lists, ids, gts, guides,
                                        drop_iter=3, patch_num=P, threshold=0.15, data_type="voc", mode="blur", n_class=21,
                                        coco=False, argsort_kind="stable")
    got0 = np.load(tmp_path / "out/hist_withfiltered_caption/img_2007_000033_max_blocknum_8_atthead_9.npy")
    gotagg = np.load(tmp_path / "out/all_drop_hist_with_filtered_caption/img_2007_000033_max_blocknum_8_atthead_9.npy")
    assert got0.dtype == np.float64 and got0.shape == (21, 21)
    assert np.array_equal(got0, h0) and np.array_equal(gotagg, hagg)
    assert got0.sum() == sum(h * w for h, w in sizes)
    assert torch.equal(R.save_img_union_attention.last[1].cpu(), torch.from_numpy(hagg.astype(np.int64)))


def test_postprocess_entry_points_match_oracle(dev):
    from oracle import hotpath as O
    from pnp_ovss_b200 import reference_api as R
    H, W, C = 60, 52, 4
    rng = np.random.default_rng(8)
    x = torch.from_numpy(rng.random((C, H, W)).astype(np.float32))
    x[0] = (x[1:].max(0)[0] < 0.6).float()
    img = synth.guide_image(4, H, W)
    gts = [np.zeros((H, W), np.float32)]

    class A:
        postprocess = "blur"
    got = R.postprocess(A, x.clone(), [img], gts, 0)
    want = O.postprocess("blur", x.clone(), img, (H, W))
    assert got.dtype == want.dtype and np.array_equal(got, want)
    A.postprocess = "blur+crf"
    got = R.postprocess(A, x.clone(), [img], gts, 0)
    want = O.postprocess("blur+crf", x.clone(), img, (H, W))
    assert got.dtype == np.float32 and (got != want).mean() <= 2e-3
    b = R.blurring(x[1], (H, W))
    assert np.allclose(b, O.blurring(x[1], (H, W)), rtol=1e-3, atol=1e-4)
    s = R.Scale_0_1(x.clone().to(dev))
    assert torch.equal(s.cpu(), O.scale_0_1(x.clone()))


def test_config0_full_size_model_single_image(dev):
    """BASELINE.json configs[0]: one 336x336 image, BLIP ITM-large shape (random init), VOC 20 classes, drop_iter 1,
    blur only -- the reference's CPU-runnable case.  CPU: the reference procedure (torch-CPU model pass with hooks and
    full backward, then oracle post-processing).  GPU: the product pipeline with the SAME weights."""
    import copy
    from oracle import hotpath as O
    from oracle import reference_arm as RA
    from pnp_ovss_b200 import pipeline
    from pnp_ovss_b200.blip_itm import BlipITM
    voc = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "table", "dog", "horse",
           "motorbike", "person", "plant", "sheep", "sofa", "train", "television"]
    tok = synth.SyntheticWordPieceTokenizer()
    caps = ["A picture of " + " ".join(voc)]
    tokens = tok(caps, padding="max_length", max_length=500)
    torch.manual_seed(2024)
    model = BlipITM(img_size=336, tokenizer=tok).eval()
    g = torch.Generator().manual_seed(7)
    imgs = torch.randn(1, 3, 336, 336, generator=g)
    gts = [synth.gt_labels(3, 336, 336, 21)]
    guides = [synth.guide_image(3, 336, 336)]
    ids = [list(range(1, 21))]
    ref_model = RA.install_reference_capture(copy.deepcopy(model))
    want = RA.compute_gradcam_ensemble_reference(ref_model, imgs, caps, tokens)[0][7][9]
    gm = model.to(dev).requires_grad_(False)
    got, _ = gm.gradcam(imgs.to(dev), caps, tokens.to(dev), layer=7, head=9)
    scale = float(want.abs().max())
    assert scale > 0
    rel = (got.cpu() - want).abs().max().item() / scale
    assert rel <= 1e-3, "block-8 head-9 GradCAM of the full-size model differs by %g of its max" % rel
    # post-processing of the SAME map on both sides (drop_iter 1 -> only the round-0 pass, Scale_0_1 on, blur only)
    with np.errstate(all="ignore"):
        h_ref, _, _ = O.batch_confusion(lambda x: want, imgs, tokens.input_ids, tok.decode, [voc], ids, gts, guides, drop_iter=1,
                                        patch_num=21, threshold=0.15, data_type="voc", mode="blur", n_class=21)
    import smoke_case
    for lowrank in (False, True):
        pipeline.USE_LOWRANK_BLUR = lowrank
        try:
            h_gpu, h_agg, _ = pipeline.batch_confusion(lambda x: want.to(dev), imgs.to(dev), tokens.input_ids.tolist(), tok.decode, [voc], ids,
                                                       gts, guides, drop_iter=1, patch_num=21, threshold=0.15, data_type="voc", mode="blur",
                                                       n_class=21)
        finally:
            pipeline.USE_LOWRANK_BLUR = True
        assert h_agg is None
        if not lowrank:   # direct kernels (the reference's summation order): identical saliency maps in -> bit-exact confusion matrix
            assert np.array_equal(h_gpu.cpu().numpy(), h_ref)
        else:             # fused low-rank blur: maps equal to 5e-6, so only pixels on a numerical tie between two channels can move
            d = smoke_case.disagreement(h_gpu.cpu().numpy(), h_ref)
            print("low-rank blur: %.3g of the pixels land in another bin" % d)
            assert d <= 1e-4, d
    # and the whole thing end to end from the GPU model's own map: report-level agreement
    h_e2e, _, _ = pipeline.batch_confusion(lambda x: gm.gradcam(x, caps, tokens.to(dev), layer=7, head=9)[0], imgs.to(dev),
                                           tokens.input_ids.tolist(), tok.decode, [voc], ids, gts, guides, drop_iter=1, patch_num=21,
                                           threshold=0.15, data_type="voc", mode="blur", n_class=21)
    assert smoke_case.disagreement(h_e2e.cpu().numpy(), h_ref) <= 0.02


def test_3xtf32_gemm_mode_keeps_fp32_grade_gradcam(dev):
    """--gemm 3xtf32: the ViT linears run as three TF32 tensor-core GEMMs on an exact hi/lo split of both operands.
    The GradCAM must stay far inside the 1e-3 tolerance (plain TF32 does not: profiles/tf32_gradcam_error.py)."""
    from pnp_ovss_b200.blip_itm import BlipITM, _tf32_split
    t = torch.randn(4096, device=dev) * 3
    hi, lo = _tf32_split(t)
    assert torch.equal(hi + lo, t) and int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
    tok = synth.SyntheticWordPieceTokenizer()
    caps = ["A picture of cat aeroplane", "A picture of dog"]
    tokens = tok(caps, padding="max_length", max_length=500).to(dev)
    torch.manual_seed(3)
    m = BlipITM(img_size=96, tokenizer=tok, hidden=128, layers=3, heads=2, inter=256, vit_dim=256, vit_depth=4, vit_heads=4,
                max_pos=64).eval()
    with torch.no_grad():
        for p_ in m.parameters():
            p_.mul_(2.0)
    m = m.to(dev).requires_grad_(False)
    imgs = torch.randn(2, 3, 96, 96, generator=torch.Generator().manual_seed(1)).to(dev)
    ref, _ = m.gradcam(imgs, caps, tokens, layer=1, head=1)
    m.gemm_precision = "3xtf32"
    got, _ = m.gradcam(imgs, caps, tokens, layer=1, head=1)
    assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_tensor_core_gemm_modes_keep_fp32_grade_gradcam_on_the_full_size_model(dev):
    """The shipped default (3xFP16) and 3xTF32 on the full-size random-init BLIP ITM-large: block-8 / head-9 GradCAM against an
    fp64 autograd pass of the same model.  Gates: no further from fp64 than 1.1x torch's own default fp32 path (which lets cuDNN
    run the patch embedding in TF32, as the reference's GPU path does) and far inside the 1e-3 tolerance of the north star.
    Confusion matrices of a 4-round batch (Salience DropOut picks patches by rank, so tiny map differences can move a patch)
    are compared with the STRICT fp32 run's (cuDNN TF32 off, 3e-6 from fp64): each tensor-core mode must agree with it at
    least as well as torch's default fp32 path does, or to 1e-4 of the pixels."""
    import bench
    import smoke_case
    from pnp_ovss_b200 import pipeline
    from pnp_ovss_b200.blip_itm import BlipITM
    cfg = dict(bench.CONFIGS[1], B=4)
    w = bench.make_workload(0, cfg)
    torch.manual_seed(4321)
    model = BlipITM(img_size=336, tokenizer=w["tok"]).to(dev).eval().requires_grad_(False)
    imgs, caps, tok = w["imgs"].to(dev), w["captions"], w["tokens"].to(dev)
    truth = bench.gradcam_fp64(model, imgs, caps, tok, 7, 9, 21)
    sc = truth.abs().max()
    err, hists = {}, {}
    for mode in ("fp32_strict", "fp32", "3xtf32", "3xfp16"):
        model.gemm_precision = "fp32" if mode.startswith("fp32") else mode
        torch.backends.cudnn.allow_tf32 = mode != "fp32_strict"
        try:
            cam = model.gradcam(imgs, caps, tok, layer=7, head=9)[0]
            err[mode] = float(((cam.double() - truth).abs().max() / sc).item())
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            _, hagg, _ = pipeline.batch_confusion(lambda x: model.gradcam(x, caps, tok, layer=7, head=9)[0], imgs.clone(),
                                                  w["tokens"].input_ids.tolist(), w["tok"].decode, w["class_lists"], w["dataset_ids"],
                                                  torch.from_numpy(w["gts"]).to(dev), torch.from_numpy(w["guides"]).to(dev), drop_iter=4,
                                                  patch_num=21, threshold=0.15, data_type="voc", mode="blur+crf", n_class=21, bad_count=bad)
        finally:
            torch.backends.cudnn.allow_tf32 = True
        hists[mode] = hagg.cpu().numpy()
        assert int(bad.item()) == 0
    model.check_fp16_overflow()
    print("GradCAM max error vs fp64 / max: %s" % err)
    dis = {m: smoke_case.disagreement(hists[m], hists["fp32_strict"].astype(np.float64)) for m in ("fp32", "3xtf32", "3xfp16")}
    print("pixels in another bin than the strict-fp32 run: %s" % dis)
    for mode in ("3xtf32", "3xfp16"):
        assert err[mode] <= 1.1 * err["fp32"] and err[mode] <= 1e-4, err
        assert dis[mode] <= max(1e-4, dis["fp32"]), dis
