"""CPU dry run of the host composition (pipeline.batch_confusion) with the C-ABI ops replaced by shape-faithful stubs
(tests/stub_ops.py): call order, bucket handling, the no-host-sync contract and the Python glue are exercised on a box
without a GPU.  The arithmetic is meaningless here -- parity lives in the -m gpu tests."""
import numpy as np
import pytest
import torch

import stub_ops
import synth


@pytest.fixture()
def pipe(monkeypatch):
    from pnp_ovss_b200 import ops, pipeline
    stub = stub_ops.make(ops)
    monkeypatch.setattr(pipeline, "ops", stub)
    monkeypatch.setattr(pipeline, "_SPATIAL", pipeline.SpatialLatticeCache())
    pipeline._SEG_TABLES.clear()
    return pipeline, stub


def _batch(B, C, S=64, G=48, n_class=21, ragged=False):
    tok = synth.SyntheticWordPieceTokenizer()
    names = ["aeroplane", "bicycle", "bird", "boat", "motorbike", "television"][:C]
    class_lists = [names[:(1 + b % C) if ragged else C] for b in range(B)]
    caps = ["A picture of " + " ".join(c) for c in class_lists]
    tokens = tok(caps, padding="max_length", max_length=500)
    T = max(len(tok.encode(c)) for c in caps)
    P = S // 16
    fn = synth.SynthGradcamFn(3, B, T, P)
    rows = tokens.attention_mask[:, 1:T].float()
    imgs = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(0))
    gts = [synth.gt_labels(10 + b, G, G, n_class) for b in range(B)]
    guides = [synth.guide_image(20 + b, G, G) for b in range(B)]
    ids = [[1 + names.index(c) for c in cl] for cl in class_lists]
    return dict(tok=tok, tokens=tokens, class_lists=class_lists, fn=lambda x: fn(x, rows), imgs=imgs, gts=gts, guides=guides, ids=ids, P=P,
                valid=sum(int(((g >= 0) & (g < n_class)).sum()) for g in gts))


@pytest.mark.parametrize("mode,drop_iter,coco", [("blur+crf", 3, False), ("blur", 1, False), ("crf", 4, True), ("", 2, False)])
def test_batch_confusion_composes_on_the_host(pipe, mode, drop_iter, coco):
    pipeline, stub = pipe
    b = _batch(3, 3)
    bad = torch.zeros(1, dtype=torch.int32)
    labels = {}
    h0, hall, chosen = pipeline.batch_confusion(b["fn"], b["imgs"], b["tokens"].input_ids.tolist(), b["tok"].decode, b["class_lists"], b["ids"],
                                                b["gts"], b["guides"], drop_iter=drop_iter, patch_num=b["P"], threshold=0.15,
                                                data_type="coco_object" if coco else "voc", mode=mode, n_class=91 if coco else 21, coco=coco,
                                                overlap=False, labels_out=labels, bad_count=bad)
    n_passes = (0 if (coco and drop_iter >= 3) else 1) + (1 if drop_iter > 1 else 0)
    assert (h0 is None) == (coco and drop_iter >= 3) and (hall is None) == (drop_iter == 1)
    for h in (h0, hall):
        if h is not None:
            assert h.dtype == torch.int64 and int(h.sum()) == b["valid"]     # every valid pixel lands in exactly one bin
    assert sorted(labels) == [k for k, h in (("all_drop", hall), ("round0", h0)) if h is not None]
    assert all(v.shape == (3, 48, 48) for v in labels.values())
    assert stub.calls.count("token_merge") == n_passes == stub.calls.count("confusion_accumulate")
    assert stub.calls.count("crf_inference") == (n_passes if "crf" in mode else 0)
    assert stub.calls.count("lowrank_blur_unary") == (n_passes if "blur" in mode else 0)     # the fused (d) group
    assert stub.calls.count("gaussian_blur") == 0 and stub.calls.count("threshold_upsample") == (0 if "blur" in mode else n_passes)
    assert stub.calls.count("build_lattice") == (2 if "crf" in mode else 0)          # one bilateral (shared by both passes) + one spatial
    assert stub.calls.count("salience_dropout_round") == (drop_iter if drop_iter > 1 else 0)
    assert (chosen is None) == (drop_iter == 1)


def test_ragged_batch_is_bucketed_and_segment_tables_are_cached(pipe):
    pipeline, stub = pipe
    b = _batch(4, 3, ragged=True)
    args = (b["fn"], b["imgs"], b["tokens"].input_ids.tolist(), b["tok"].decode, b["class_lists"], b["ids"], b["gts"], b["guides"])
    kw = dict(drop_iter=2, patch_num=b["P"], threshold=0.15, data_type="voc", mode="blur+crf", n_class=21, overlap=False)
    h0, hall, _ = pipeline.batch_confusion(*args, **kw)
    assert int(h0.sum()) == b["valid"] == int(hall.sum())
    # class counts that pad to the same CRF channel count share a bucket (the fused (d) group takes them per image) ...
    from pnp_ovss_b200 import host
    keys = {((len(c) + int(host.add_background_rule("voc", len(c))) + 3) // 4 * 4, tuple(g.shape)) for c, g in zip(b["class_lists"], b["gts"])}
    assert stub.calls.count("crf_inference") == 2 * len(keys) < 2 * len({len(c) for c in b["class_lists"]})
    # ... and with that switched off every class count is a bucket of its own, with the same matrices
    stub.calls.clear()
    pipeline.PAD_CLASSES_IN_BUCKETS = False
    try:
        h0_exact, hall_exact, _ = pipeline.batch_confusion(*args, **kw)
    finally:
        pipeline.PAD_CLASSES_IN_BUCKETS = True
    assert stub.calls.count("crf_inference") == 2 * len({(len(c), tuple(g.shape)) for c, g in zip(b["class_lists"], b["gts"])})
    assert int(h0_exact.sum()) == b["valid"] == int(hall_exact.sum())
    pipeline.PAD_CLASSES_IN_BUCKETS = False
    try:
        with pytest.raises(pipeline.PnpError):        # label maps of a batch that falls into several buckets have no single shape
            pipeline.batch_confusion(*args, labels_out={}, **kw)
    finally:
        pipeline.PAD_CLASSES_IN_BUCKETS = True
    seg_a = pipeline.segment_tensors(b["tokens"].input_ids.tolist(), b["tok"].decode, b["class_lists"], torch.device("cpu"))
    seg_b = pipeline.segment_tensors(b["tokens"].input_ids.tolist(), b["tok"].decode, b["class_lists"], torch.device("cpu"))
    assert seg_a is seg_b and isinstance(seg_a[3], int)                     # same captions: nothing is rebuilt or uploaded
    assert seg_a[3] == int((seg_a[0] + seg_a[1]).max())


def test_run_on_streams_keeps_order_and_propagates_errors():
    """pipeline._run_on_streams without streams (CPU): results come back in job order, the first exception is re-raised."""
    from pnp_ovss_b200 import pipeline
    assert pipeline._run_on_streams(torch.device("cpu"), [], [lambda i=i: i * i for i in range(7)]) == [i * i for i in range(7)]

    def boom():
        raise ValueError("job failed")

    with pytest.raises(ValueError):
        pipeline._run_on_streams(torch.device("cpu"), [], [lambda: 1, boom, lambda: 3])


def test_ragged_bucket_keys_pad_class_counts_only_in_blurred_modes(pipe):
    """Class counts share a bucket only where the fused low-rank (d) group can take them per image (a blurred mode); without blur the
    exact count stays in the key, and either way every image lands in exactly one bucket (the matrices count every valid pixel once)."""
    pipeline, stub = pipe
    b = _batch(6, 3, ragged=True)
    args = (b["fn"], b["imgs"], b["tokens"].input_ids.tolist(), b["tok"].decode, b["class_lists"], b["ids"], b["gts"], b["guides"])
    counts = {}
    for mode in ("blur", "crf", ""):
        stub.calls.clear()
        h0, _, _ = pipeline.batch_confusion(*args, drop_iter=1, patch_num=b["P"], threshold=0.15, data_type="voc", mode=mode, n_class=21)
        assert int(h0.sum()) == b["valid"]
        counts[mode] = stub.calls.count("confusion_accumulate")
    assert counts["blur"] == 1 and counts["crf"] == counts[""] == 3
