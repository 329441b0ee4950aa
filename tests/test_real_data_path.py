"""The real-data entry of the driver (pnp_ovss_b200/data.py + driver.main_real) on miniature copies of the reference's
directory layouts.  CPU: transform vs torchvision (DS:430-443), tokenizer, id lists, category tables.  GPU: the whole
`--real_data` run (JPEG/PNG/JSON from disk, a real BertTokenizer on a small vocab, a small BlipITM) against the CPU
reference procedure on the same files."""
import copy
import json
import os
import types

import numpy as np
import pytest
import torch

import synth

VOCAB = (["[PAD]"] + ["[unused%d]" % i for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"] +
         ["a", "picture", "of", "cat", "dog", "aero", "##plane", "person", "potted", "##plant", "tv", "##monitor", "bird", "train",
          "bicycle", "boat", "bottle", "bus", "car", "chair", "cow", "table", "horse", "motor", "##bike", "sheep", "sofa"])


def _write_vocab(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(VOCAB) + "\n")
    return str(p)


def _mini_voc(home, sizes, answers):
    from PIL import Image
    root = home / "VOCdevkit/VOC2012"
    for sub in ("SegmentationClass", "JPEGImages"):
        (root / sub).mkdir(parents=True)
    (home / "GPT4o_classification").mkdir()
    ids = ["2007_%06d" % (33 + 9 * k) for k in range(len(sizes))]
    for k, (img_id, (H, W)) in enumerate(zip(ids, sizes)):
        Image.fromarray(synth.gt_labels(300 + k, H, W, 21).astype(np.uint8)).save(root / "SegmentationClass" / (img_id + ".png"))
        Image.fromarray(synth.guide_image(400 + k, H, W)).save(root / "JPEGImages" / (img_id + ".jpg"), quality=95)
    (root / "val.txt").write_text("".join(i + "\n" for i in ids))
    json.dump(dict(zip(ids, answers)), open(home / "GPT4o_classification/voc_classification_noboundary.json", "w"))
    return ids


def test_model_input_transform_matches_torchvision(tmp_path):
    from PIL import Image
    from torchvision import transforms
    from torchvision.transforms.functional import InterpolationMode
    from pnp_ovss_b200 import data
    rgb = synth.guide_image(5, 50, 70)
    path = tmp_path / "x.png"
    Image.fromarray(rgb).save(path)
    x, norm, size = data.load_model_image(str(path), 96)
    ref = transforms.Compose([transforms.Resize((96, 96), interpolation=InterpolationMode.BICUBIC), transforms.ToTensor(),
                              transforms.Normalize(mean=data.CLIP_MEAN, std=data.CLIP_STD)])(Image.open(path).convert("RGB"))
    assert x.dtype == torch.float32 and x.shape == (3, 96, 96) and torch.equal(x, ref)
    assert size == (70, 50) and norm.dtype == np.float32 and norm.shape == (96, 96, 3)
    assert np.array_equal(norm, np.float32(Image.open(path).convert("RGB").resize((96, 96))) / 255)


def test_blip_tokenizer_from_a_local_vocab(tmp_path):
    from pnp_ovss_b200 import data, host
    tok = data.init_tokenizer(_write_vocab(tmp_path))
    n = len(VOCAB)
    assert tok.convert_tokens_to_ids("[DEC]") == n and tok.enc_token_id == n + 1 and len(tok) == n + 2
    assert (tok.cls_token_id, tok.sep_token_id, tok.pad_token_id) == (101, 102, 0)        # DRV:814 hard-codes 102
    enc = tok(["A picture of aeroplane pottedplant", "A picture of dog"], padding="max_length", max_length=500, return_tensors="pt")
    assert enc.input_ids.shape == (2, 500)
    toks = host.token_strings(enc.input_ids[0].tolist(), tok.decode)
    assert toks == ["aero", "##plane", "potted", "##plant"]      # class pieces only: [CLS] a picture of ... [SEP] stripped
    assert host.build_token_segments(toks, 2) == [(0, 2, 2.0), (2, 2, 1.0)]   # last word: summed, not averaged (DRV:844-847)
    longest = tok(["A picture of aeroplane pottedplant", "A picture of dog"], padding="longest", truncation=True, max_length=500,
                  return_tensors="pt")
    assert longest.input_ids.shape == (2, 9) and longest.attention_mask[1].tolist() == [1] * 6 + [0] * 3


def test_id_lists_category_tables_and_shards(tmp_path):
    from pnp_ovss_b200 import data
    ids = _mini_voc(tmp_path, [(8, 9), (7, 7), (9, 8)], ["[]"] * 3)
    args = types.SimpleNamespace(home_dir=str(tmp_path), data_type="voc")
    assert data.image_ids(args) == ids
    cats, nms = data.categories(args)
    assert len(cats) == 20 and cats[16] == "pottedplant" and nms[19] == "tvmonitor"
    assert os.path.isfile(data.image_path(args, ids[0]))
    assert data.batches(ids, 2) == [ids[:2], ids[2:]]
    assert data.batches(ids, 2, rank=0, world_size=2) == [ids[:2]] and data.batches(ids, 2, rank=1, world_size=2) == [ids[2:]]
    args.data_type = "ade20k"
    cats, nms = data.categories(args)
    assert len(cats) == 150 and cats[45] == "chest of drawers" and nms[44] == "chestofdrawers"
    d = tmp_path / "semantic-segmentation-pytorch-master/data"
    d.mkdir(parents=True)
    (d / "validation.odgt").write_text("".join(json.dumps({"fpath_img": "ADEChallengeData2016/images/validation/ADE_val_%08d.jpg" % i,
                                                            "width": 4, "height": 4}) + "\n" for i in (1, 12, 2000)))
    assert data.image_ids(args) == ["1", "12", "2000"]
    assert data.image_path(args, "12").endswith("ADE_val_00000012.jpg")
    args.data_type = "coco_stuff"
    ann = tmp_path / "instances.json"
    json.dump({"categories": [{"id": 3, "name": "car"}, {"id": 1, "name": "person"}]}, open(ann, "w"))
    cats, nms = data.categories(args, str(ann))
    assert [c["id"] for c in cats] == [1, 3] and nms == ["person", "car"]
    (tmp_path / "coco/images/val2017").mkdir(parents=True)
    for n in ("000000000139.jpg", "000000000285.jpg"):
        (tmp_path / "coco/images/val2017" / n).write_bytes(b"")
    assert data.image_ids(args) == ["139", "285"]
    with pytest.raises(ValueError):
        data.categories(args)


@pytest.mark.gpu
def test_real_data_run_matches_the_cpu_reference_procedure(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import hotpath as O
    from oracle import reference_arm as RA
    from pnp_ovss_b200 import data, driver
    from pnp_ovss_b200 import reference_api as R
    from pnp_ovss_b200.blip_itm import BlipITM
    sizes = [(50, 70), (64, 48), (56, 56), (40, 90)]
    answers = ["[15: 'person', 12: 'dog'], [95%, 80%]", "[1: 'aeroplane', 8: 'cat', 3: 'bird'],\n[90%, 60%, 99%]",
               "[16: 'pottedplant', 19: 'train'], [75%, 88%]", "[20: 'tvmonitor'], [99%]"]
    home = tmp_path / "home"
    home.mkdir()
    ids = _mini_voc(home, sizes, answers)
    vocab = _write_vocab(tmp_path)
    S, P = 96, 6
    args = driver.get_args_parser().parse_args([
        "--real_data", "--home_dir", str(home), "--data_type", "voc", "--bert_vocab", vocab, "--img_size", str(S), "--batch_size", "3",
        "--max_att_block_num", "2", "--prune_att_head", "1", "--drop_iter", "3", "--threshold", "0.15", "--postprocess", "blur",
        "--world_size", "1", "--save_path", str(tmp_path / "out")])
    tok = data.init_tokenizer(vocab)
    torch.manual_seed(11)
    model = BlipITM(img_size=S, tokenizer=tok, vocab=len(tok), hidden=128, layers=3, heads=2, inter=256, vit_dim=64, vit_depth=2,
                    vit_heads=2, max_pos=64).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(2.0)
    ref_model = RA.install_reference_capture(copy.deepcopy(model))
    got = driver.main_real(0, 1, args, model=model)

    # CPU: the reference procedure on the same files, batch by batch (batch membership matters: DRV:638 quirk)
    cats, nms = data.categories(args)
    total = np.zeros((21, 21))
    for img_ids in data.batches(ids, 3):
        loaded = [data.load_model_image(data.image_path(args, i), S) for i in img_ids]
        imgs = torch.stack([t for t, _, _ in loaded])
        best, cls, caps = [], [], []
        for k in range(len(img_ids)):
            R.Load_predicted_classes(args, nms, best, cls, caps, None, img_ids, k, None)
        tokens = tok(caps, padding="max_length", max_length=500, return_tensors="pt")
        gts, guides = R.Load_GroundTruth(args, img_ids), R.load_OrgImage(args, img_ids)
        with np.errstate(all="ignore"):
            h0, hagg, _ = O.batch_confusion(lambda x: RA.compute_gradcam_ensemble_reference(ref_model, x, caps, tokens)[0][1][1], imgs,
                                            tokens.input_ids, tok.decode, cls, [[i + 1 for i in b] for b in best], gts, guides,
                                            drop_iter=3, patch_num=P, threshold=0.15, data_type="voc", mode="blur", n_class=21,
                                            coco=False, argsort_kind="stable")
        total += hagg
    assert got.shape == (21, 21) and got.sum() == total.sum() == sum(h * w for h, w in sizes)
    disagree = np.abs(got.astype(np.float64) - total).sum() / 2.0 / total.sum()
    assert disagree <= 0.02, disagree
    first = tmp_path / "out/all_drop_hist_with_filtered_caption" / ("img_%s_max_blocknum_2_atthead_1.npy" % ids[0])
    assert first.is_file() and np.load(first).sum() == sum(h * w for h, w in sizes[:3])
