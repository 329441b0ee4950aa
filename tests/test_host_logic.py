"""CPU: host-side logic of the product (token segment tables, relabel LUT, sharding, metric arithmetic) against the
oracle and the reference-generated golden fixtures; and the 2-rank confusion-matrix all-reduce over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import synth
from make_golden_cases import MERGE_CASES
from oracle import hotpath as O
from pnp_ovss_b200 import host


def _emulate_merge(g, segs):
    """What pnp_token_merge computes, in numpy fp32 (sequential adds, one division)."""
    rows = g[3:]  # pnp_token_merge addresses rows from row_offset = 3; segments never reach the trailing SEP row
    out = np.zeros((len(segs),) + g.shape[1:], np.float32)
    for c, (s, l, d) in enumerate(segs):
        if l == 0:
            continue
        acc = rows[s].copy()
        for i in range(1, l):
            acc = (acc + rows[s + i]).astype(np.float32)
        out[c] = acc / np.float32(d) if d != 1.0 else acc
    return out


def test_token_segments_reproduce_reference_merge(golden):
    tok = synth.SyntheticWordPieceTokenizer()
    for class_lists in MERGE_CASES.values():
        for cl in class_lists:
            tok("A picture of " + " ".join(cl))
    for name, class_lists in MERGE_CASES.items():
        ids = golden["merge_%s_ids" % name]
        g = golden["merge_%s_g" % name]
        for b, cl in enumerate(class_lists):
            toks = host.token_strings(ids[b], tok.decode)
            assert toks == O.token_strings(ids[b], tok.decode)
            segs = host.build_token_segments(toks, len(cl))
            assert max(s + l for s, l, _ in segs) <= g.shape[1] - 4   # stays inside the [3:-1] slice
            assert np.array_equal(_emulate_merge(g[b], segs), golden["merge_%s_out%d" % (name, b)]), (name, b)


def test_token_segments_quirks():
    # split word last: summed, never divided (DRV:844-847)
    assert host.build_token_segments(["dog", "aero", "##plan", "##e"], 2) == [(0, 1, 1.0), (1, 3, 1.0)]
    # split word followed by another word: averaged
    assert host.build_token_segments(["aero", "##plan", "##e", "dog"], 2) == [(0, 3, 3.0), (3, 1, 1.0)]
    # as many pieces as classes: rows taken as they are (DRV:852), even if some are continuations
    assert host.build_token_segments(["pot", "##ted"], 2) == [(0, 1, 1.0), (1, 1, 1.0)]
    with pytest.raises(IndexError):
        host.build_token_segments(["a", "b", "c", "d"], 2)


def test_relabel_lut_matches_sequential_relabel():
    rng = np.random.default_rng(0)
    for _ in range(50):
        C = int(rng.integers(1, 7))
        ids = [int(v) for v in rng.integers(1, 9, C)]
        for with_bg in (True, False):
            n = C + (1 if with_bg else 0)
            m = rng.integers(0, n, (6, 7)).astype(np.float32)
            want = O.relabel_sequential(m.copy(), ids, with_bg)
            lut = host.relabel_lut(ids, with_bg)
            assert np.array_equal(np.asarray(lut, np.float32)[m.astype(int)], want)
    assert host.relabel_lut([2, 1], True) == [0, 2, 2]  # the aliasing quirk (a9)


def test_background_rule_and_sharding():
    assert host.add_background_rule("voc", 5) and host.add_background_rule("coco_object", 9)
    assert host.add_background_rule("ade20k", 2) and not host.add_background_rule("ade20k", 3)
    assert not host.add_background_rule("coco_stuff", 4) and host.add_background_rule("psc", 1)
    for n, w in ((35, 8), (7, 2), (3, 4), (1449, 8)):
        spans = [host.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))            # contiguous, nothing counted twice
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1


def test_metrics_from_hist_matches_reference(golden):
    from pnp_ovss_b200.reference_api import metrics_from_hist
    table, _ = metrics_from_hist(golden["scores_hist"])
    assert table["Mean IoU"] == float(golden["scores_miou"])
    assert table["Pixel Accuracy"] == float(golden["scores_acc"])
    assert table["Frequency Weighted IoU"] == float(golden["scores_fwiou"])


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from pnp_ovss_b200 import pipeline
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 21
    rng = np.random.default_rng(100 + rank)
    s, e = host.shard_range(7, rank, world)
    hist = torch.zeros((n, n), dtype=torch.int64)
    for img in range(s, e):
        gt = synth.gt_labels(img, 40, 36, n)
        pred = np.random.default_rng(img).integers(0, n, gt.shape)
        hist += torch.from_numpy(O.fast_hist(gt.flatten(), pred.flatten(), n))
    pipeline.allreduce_hist(hist)
    if rank == 0:
        np.save(out, hist.numpy())
    dist.destroy_process_group()


def test_allreduce_hist_two_ranks_gloo(tmp_path):
    """N>1 path on CPU: two ranks shard 7 images, all-reduce their int64 matrices, rank 0 holds the global matrix."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "hist.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    n = 21
    want = np.zeros((n, n), np.int64)
    for img in range(7):
        gt = synth.gt_labels(img, 40, 36, n)
        pred = np.random.default_rng(img).integers(0, n, gt.shape)
        want += O.fast_hist(gt.flatten(), pred.flatten(), n)
    assert np.array_equal(np.load(out), want)


def test_save_hist_npy_layout(tmp_path):
    from pnp_ovss_b200 import pipeline
    h = torch.arange(9, dtype=torch.int64).view(3, 3)
    p = pipeline.save_hist_npy(h, str(tmp_path), "all_drop_hist_with_filtered_caption", "2007_000033", 8, 9)
    assert p.endswith("all_drop_hist_with_filtered_caption/img_2007_000033_max_blocknum_8_atthead_9.npy")
    back = np.load(p)
    assert back.dtype == np.float64 and np.array_equal(back, h.numpy())


def test_gpt4o_class_list_parser_matches_reference():
    """host.parse_gpt4o_classes vs the reference's Load_predicted_classes run on its own shipped GPT-4o answers
    (tests/golden/gpt4o_golden.json, made by tests/golden/make_golden_gpt4o.py)."""
    import json
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpt4o_golden.json")))
    n = 0
    for data_type in ("voc", "psc", "ade20k"):
        blob = cases[data_type]
        names = ["name%03d" % i for i in range(blob["n_names"])]
        for c in blob["cases"]:
            if "error" in c:
                with pytest.raises(Exception):
                    host.parse_gpt4o_classes(c["raw"], names)
                continue
            best, cls, cap = host.parse_gpt4o_classes(c["raw"], names)
            assert best == c["best_class_idx"] and cap == c["caption"], (data_type, c["raw"])
            n += 1
    assert n > 300
    assert host.parse_gpt4o_classes("[3: 'bird'], [60%]", ["a", "b", "c"])[0] == [0]       # nothing above 70 -> class 0


def test_calculate_miou_reads_the_reference_layout(tmp_path):
    from pnp_ovss_b200 import calculate_miou, pipeline
    rng = np.random.default_rng(0)
    hs = [rng.integers(0, 50, (5, 5)) for _ in range(3)]
    for i, h in enumerate(hs):  # one file per batch, like the reference writes them (DRV:517-520)
        pipeline.save_hist_npy(torch.from_numpy(h), str(tmp_path), "all_drop_hist_with_filtered_caption", "img%d" % i, 8, 9)
    table, hist = calculate_miou.main(["--save_path", str(tmp_path)])
    assert np.array_equal(hist, sum(hs).astype(np.float64))
    want, _ = O.scores([np.zeros((1, 1))], [np.zeros((1, 1))], 5)  # just for the key set
    assert set(want) == set(table)
    iu = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
    assert table["Mean IoU"] == np.nanmean(iu[hist.sum(1) > 0])


def test_ground_truth_and_guide_image_loaders(tmp_path):
    """Load_GroundTruth / load_OrgImage (DRV:901-955, DRVC:1111-1138) on a miniature copy of the directory layouts."""
    from PIL import Image
    from pnp_ovss_b200 import reference_api as R
    rng = np.random.default_rng(0)
    home = tmp_path

    def put(rel, arr, mode=None):
        p = home / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        Image.fromarray(arr, mode=mode).save(p)

    gt = rng.integers(0, 21, size=(7, 9)).astype(np.uint8)
    gt[0, :3] = 255
    rgb = rng.integers(0, 256, size=(7, 9, 3)).astype(np.uint8)
    put("VOCdevkit/VOC2012/SegmentationClass/2007_000033.png", gt)
    put("VOCdevkit/VOC2012/JPEGImages/2007_000033.png", rgb)          # lossless stand-in, renamed below
    (home / "VOCdevkit/VOC2012/JPEGImages/2007_000033.png").rename(home / "VOCdevkit/VOC2012/JPEGImages/2007_000033.jpg")
    put("mmsegmentation/data/VOCdevkit/VOC2010/SegmentationClassContext/2007_000033.png", gt)
    put("ADEChallengeData2016/annotations/validation/ADE_val_00000012.png", gt)
    put("ADEChallengeData2016/images/validation/ADE_val_00000012.png", rgb)
    (home / "ADEChallengeData2016/images/validation/ADE_val_00000012.png").rename(
        home / "ADEChallengeData2016/images/validation/ADE_val_00000012.jpg")
    put("coco_stuff164k/annotations/val2017/000000000139.png", gt)
    put("coco/images/val2017/000000000139.png", rgb)
    (home / "coco/images/val2017/000000000139.png").rename(home / "coco/images/val2017/000000000139.jpg")

    class A:
        home_dir = str(home)
        data_type = "voc"

    voc = R.Load_GroundTruth(A, ["2007_000033"])[0]
    assert voc.dtype == np.float32 and voc.shape == (7, 9)
    want = gt.astype(np.float32)
    want[want == 255] = 0
    assert np.array_equal(voc, want)
    img = R.load_OrgImage(A, ["2007_000033"])[0]
    assert img.dtype == np.uint8 and np.array_equal(img, rgb)
    A.data_type = "psc"
    assert np.array_equal(R.Load_GroundTruth(A, ["2007_000033"])[0], gt.astype(np.float32))      # 255 kept
    assert np.array_equal(R.load_OrgImage(A, ["2007_000033"])[0], rgb)
    A.data_type = "ade20k"
    assert np.array_equal(R.Load_GroundTruth(A, ["12"])[0], gt.astype(np.float32))
    assert np.array_equal(R.load_OrgImage(A, ["12"])[0], rgb)
    A.data_type = "coco_stuff"
    stuff = R.Load_GroundTruth(A, ["139"])[0]
    ref = gt.astype(np.float32)
    for i in range(ref.shape[0]):               # the reference's loop, DRVC:1117-1122
        for j in range(ref.shape[1]):
            ref[i][j] = 0 if ref[i][j] == 255 else ref[i][j] + 1
    assert stuff.dtype == np.float32 and np.array_equal(stuff, ref)
    assert np.array_equal(R.load_OrgImage(A, ["139"])[0], rgb)
    A.data_type = "coco_object"
    with pytest.raises(ValueError):
        R.Load_GroundTruth(A, ["139"])
    A.data_type = "nope"
    with pytest.raises(ValueError):
        R.load_OrgImage(A, ["1"])


def test_gpt4o_class_list_parser_coco_variant_matches_reference(tmp_path):
    """host.parse_gpt4o_classes_coco and reference_api.Load_predicted_classes vs the COCO driver's own function
    (DRVC:855-963) run on the answers the reference ships (tests/golden/gpt4o_golden.json)."""
    import json
    import types
    from pnp_ovss_b200 import reference_api as R
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpt4o_golden.json")))
    for data_type in ("coco_object", "coco_stuff"):
        block = cases[data_type]
        ids, names = block["cat_ids"], ["name%03d" % i for i in range(block["n_names"])]
        assert len(block["cases"]) > 100
        for c in block["cases"]:
            best, cls, cap = host.parse_gpt4o_classes_coco(c["raw"], ids, names, data_type)
            assert best == c["best_class_idx"] and cap == c["caption"], c["raw"]
    # the file-reading entry points, both signatures, on a miniature GPT4o_classification directory
    d = tmp_path / "GPT4o_classification"
    d.mkdir()
    voc_case, coco_case = cases["voc"]["cases"][0], cases["coco_object"]["cases"][0]
    json.dump({"2007_000033": voc_case["raw"]}, open(d / "voc_classification_noboundary.json", "w"))
    json.dump({"000000000139": coco_case["raw"]}, open(d / "coco_object_classification_noboundary.json", "w"))
    args = types.SimpleNamespace(home_dir=str(tmp_path), data_type="voc")
    b, c, caps = R.Load_predicted_classes(args, ["name%03d" % i for i in range(20)], [], [], [], None, ["2007_000033"], 0, None)
    assert b == [voc_case["best_class_idx"]] and caps == [voc_case["caption"]] and len(c[0]) == len(b[0])
    args.data_type = "coco_object"
    ids = cases["coco_object"]["cat_ids"]
    b, c, caps = R.Load_predicted_classes(args, ["name%03d" % i for i in range(len(ids))], [{"id": i} for i in ids], [], [], [],
                                          [None], [139], 0, None)
    assert b == [coco_case["best_class_idx"]] and caps == [coco_case["caption"]]


def test_token_segments_property_random_wordpiece_shapes():
    """For arbitrary captions (1..9 words of 1..4 pieces each) the host's segment table, applied the way pnp_token_merge
    applies it, reproduces the oracle's restatement of the reference loop bit for bit -- including the cases where the
    number of pieces happens to equal the number of classes (rows taken verbatim) and a split word comes last (summed)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=4), min_size=1, max_size=9), st.integers(min_value=0, max_value=2 ** 31 - 1))
    def check(pieces_per_word, seed):
        toks = []
        for w, n in enumerate(pieces_per_word):
            toks.append("w%d" % w)
            toks.extend("##p%d" % k for k in range(n - 1))
        n_classes = len(pieces_per_word)
        rng = np.random.default_rng(seed)
        g = rng.random((3 + len(toks) + 1, 3, 3)).astype(np.float32)      # rows: "a picture of", the pieces, [SEP]
        want = O.mean_over_filtered_label_tokens(toks, torch.from_numpy(g), n_classes).numpy()
        segs = host.build_token_segments(toks, n_classes)
        assert len(segs) == n_classes
        assert np.array_equal(_emulate_merge(g, segs), want)

    check()


def test_save_img_union_attention_coco_signature_host_glue(tmp_path, monkeypatch):
    """The COCO form of save_img_union_attention (DRVC:338: leading coco_thing, `cats` as a list of dicts with sparse ids,
    183-class matrix, round-0 pass skipped at drop_iter >= 3) with the GPU part stubbed out: what reaches
    pipeline.batch_confusion and what lands on disk."""
    import json
    import types
    from PIL import Image
    from pnp_ovss_b200 import pipeline
    from pnp_ovss_b200 import reference_api as R
    rng = np.random.default_rng(1)
    home = tmp_path / "home"
    (home / "coco_stuff164k/annotations/val2017").mkdir(parents=True)
    (home / "coco/images/val2017").mkdir(parents=True)
    (home / "GPT4o_classification").mkdir()
    gt = rng.integers(0, 171, size=(9, 11)).astype(np.uint8)
    gt[0, 0] = 255
    Image.fromarray(gt).save(home / "coco_stuff164k/annotations/val2017/000000000139.png")
    rgb = rng.integers(0, 256, size=(9, 11, 3)).astype(np.uint8)
    Image.fromarray(rgb).save(home / "coco/images/val2017/000000000139.png")
    (home / "coco/images/val2017/000000000139.png").rename(home / "coco/images/val2017/000000000139.jpg")
    json.dump({"000000000139": "[1: 'person', 18: 'dog', 93: 'branch'], [95%, 40%, 88%]"},
              open(home / "GPT4o_classification/coco_stuff_classification_noboundary.json", "w"))
    cat_ids = [1, 2, 18, 92, 93]
    cats = [{"id": i, "name": "n%d" % i} for i in cat_ids]
    nms = ["person", "bicycle", "dog", "banner", "branch"]
    tok = synth.SyntheticWordPieceTokenizer()

    class Model:
        tokenizer = tok
    seen = {}

    def fake_batch_confusion(gradcam_fn, imgs, token_ids, decode, class_lists, dataset_ids, gts, guides, **kw):
        seen.update(kw, class_lists=class_lists, dataset_ids=dataset_ids, gts=gts, guides=guides, imgs=imgs, token_ids=token_ids)
        return None, torch.full((kw["n_class"], kw["n_class"]), 2, dtype=torch.int64), None

    monkeypatch.setattr(R, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(pipeline, "batch_confusion", fake_batch_confusion)
    args = types.SimpleNamespace(home_dir=str(home), data_type="coco_stuff", save_path=str(tmp_path / "out"), drop_iter=4, img_size=32,
                                 max_att_block_num=8, prune_att_head="9", threshold=0.15, postprocess="blur+crf")
    imgs = torch.zeros(1, 3, 32, 32)
    coco_thing = types.SimpleNamespace(loadImgs=lambda ids: [{"file_name": "000000000139.jpg"}], getImgIds=lambda imgIds: imgIds)
    out = R.save_img_union_attention(coco_thing,
                                     types.SimpleNamespace(module=Model()), imgs, None, args, [None], [139], 4, None, None, cats,
                                     nms, None, 0, "9", 8)
    assert out is None
    assert seen["class_lists"] == [["person", "branch"]] and seen["dataset_ids"] == [[1, 93]]     # category ids, not positions
    assert seen["n_class"] == 183 and seen["coco"] is True and seen["data_type"] == "coco_stuff"
    assert seen["drop_iter"] == 4 and seen["patch_num"] == 2 and seen["mode"] == "blur+crf"
    want_gt = np.where(gt == 255, 0, gt.astype(np.float32) + 1).astype(np.float32)              # DRVC:1117-1122
    assert np.array_equal(seen["gts"][0], want_gt) and np.array_equal(seen["guides"][0], rgb)
    assert len(seen["token_ids"]) == 1 and len(seen["token_ids"][0]) == 500
    saved = np.load(tmp_path / "out/all_drop_hist_with_filtered_caption/img_139_max_blocknum_8_atthead_9.npy")
    assert saved.shape == (183, 183) and saved.dtype == np.float64 and saved.sum() == 2 * 183 * 183
    assert not (tmp_path / "out/hist_withfiltered_caption").exists() or not list((tmp_path / "out/hist_withfiltered_caption").iterdir())
