"""The driver keeps the reference's CLI (DRV:57-106) and shards images without double counting."""
import os

import numpy as np
import pytest
import torch

from pnp_ovss_b200 import driver, host

# (flag, default) exactly as PnP_OVSS_0514_updated_segmentation.py:57-106 declares them
REFERENCE_FLAGS = [
    ("batch_size", 2), ("num_workers", 0), ("gen_multiplecap_withpnpvqa", "label"), ("save_path", "Eval_test_ddp"),
    ("home_dir", "/home/letitiabanana/LAVIS/"), ("master_port", "12355"),
    ("existing_att_path", "./Cbatch_Eval_test_ddp_0521_labelascaption/img_att_forclasses/"),
    ("cam_out_dir", "./Eval_test_ddp_0331/img_att_forclasses/"), ("del_patch_num", None), ("max_att_block_num", 10),
    ("img_size", 768), ("world_size", 4), ("ensemble_blocks", None), ("drop_iter", 10), ("prune_att_head", None),
    ("sort_threshold", None), ("edge_map_for_clip", False), ("final_att_threshold", 0.05), ("search", None), ("layer", None),
    ("cal_token_sim_forall_layerhead", False), ("in_the_wild", False), ("data_type", None), ("postprocess", None),
    ("threshold", None),
]


def test_cli_flags_and_defaults_match_the_reference():
    args = driver.get_args_parser().parse_args([])
    for flag, default in REFERENCE_FLAGS:
        assert hasattr(args, flag), flag
        assert getattr(args, flag) == default, flag
    # the README's recommended command line (README.md:110-121 / Run_seg.sh) parses unchanged
    a = driver.get_args_parser().parse_args(
        "--save_path out --master_port 10990 --home_dir /data --data_type voc --batch_size 35 --max_att_block_num 8 "
        "--img_size 336 --world_size 1 --drop_iter 4 --prune_att_head 9 --del_patch_num sort_thresh005 --sort_threshold 0.05 "
        "--threshold 0.15 --postprocess blur+crf".split())
    assert (a.batch_size, a.img_size, a.drop_iter, a.prune_att_head, a.threshold, a.postprocess) == (35, 336, 4, "9", 0.15, "blur+crf")


def test_synthetic_shards_are_world_size_independent():
    a = driver.get_args_parser().parse_args("--img_size 32 --synthetic_images 7 --data_type voc".split())
    names, n = driver.DATASETS["voc"]
    whole = driver.synthetic_shard(a, names, n, 0, 7)
    parts = []
    for r in range(3):
        s, e = host.shard_range(7, r, 3)
        parts += driver.synthetic_shard(a, names, n, s, e)
    assert [p["img_id"] for p in parts] == [w["img_id"] for w in whole]
    for p, w in zip(parts, whole):
        assert torch.equal(p["img"], w["img"]) and np.array_equal(p["gt"], w["gt"]) and p["classes"] == w["classes"]


@pytest.mark.gpu
@pytest.mark.parametrize("data_type,classes", [("voc", 2), ("ade20k", 4)])
def test_driver_runs_end_to_end_on_one_gpu(tmp_path, data_type, classes):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    a = driver.get_args_parser().parse_args(
        ("--data_type %s --img_size 96 --batch_size 2 --max_att_block_num 8 --prune_att_head 9 --drop_iter 2 --del_patch_num "
         "sort_thresh005 --sort_threshold 0.05 --threshold 0.15 --postprocess blur+crf --world_size 1 --synthetic_images 3 "
         "--synthetic_classes %d --save_path %s" % (data_type, classes, tmp_path)).split())
    hist = driver.main(0, 1, a)
    n = driver.DATASETS[data_type][1]
    names = driver.DATASETS[data_type][0]
    items = driver.synthetic_shard(a, names, n, 0, 3)
    valid = sum(int(((it["gt"] >= 0) & (it["gt"] < n)).sum()) for it in items)
    assert hist.shape == (n, n) and int(hist.sum()) == valid
    saved = np.load(str(tmp_path / "all_drop_hist_with_filtered_caption" / "img_syn_000000_max_blocknum_8_atthead_9.npy"))
    assert saved.dtype == np.float64 and np.array_equal(saved, hist)
    assert np.array_equal(driver.main(0, 1, a), hist)   # deterministic


def _run_driver(world_size, save_path, port):
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "pnp_ovss_b200.driver", "--data_type", "voc", "--img_size", "96", "--batch_size", "2",
           "--max_att_block_num", "8", "--prune_att_head", "9", "--drop_iter", "2", "--del_patch_num", "sort_thresh005",
           "--threshold", "0.15", "--postprocess", "blur+crf", "--world_size", str(world_size), "--synthetic_images", "4",
           "--synthetic_classes", "2", "--save_path", save_path, "--master_port", str(port)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(cmd, check=True, cwd=root, timeout=600)
    return np.load(os.path.join(save_path, "all_drop_hist_with_filtered_caption", "img_syn_000000_max_blocknum_8_atthead_9.npy"))


@pytest.mark.gpu
def test_two_gpu_run_gives_the_same_matrix_as_one_gpu(tmp_path):
    """Data-parallel over images with mp.spawn (DRV:1439) + one NCCL all-reduce: with the same batch membership
    (4 images, batches of 2: [0,1] [2,3] on one GPU, one batch per rank on two) the all-reduced confusion matrix is
    bit-identical.  (Batch membership matters in the reference itself: the DropOut score slices rows [3:-1] of the
    longest caption IN THE BATCH)  Needs two GPUs; skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run_driver(1, str(tmp_path / "w1"), 29611)
    two = _run_driver(2, str(tmp_path / "w2"), 29612)
    assert one.sum() > 0 and np.array_equal(one, two)


def test_driver_gemm_precision_flag_defaults_to_the_shipped_mode():
    """--gemm_precision is the one flag added on top of the reference's parser for the model pass: default 3xfp16 (what bench.py
    times), fp32 / 3xtf32 selectable."""
    import argparse
    from pnp_ovss_b200 import driver
    p = argparse.ArgumentParser(parents=[driver.get_args_parser()])
    assert p.parse_args([]).gemm_precision == "3xfp16"
    assert p.parse_args(["--gemm_precision", "fp32"]).gemm_precision == "fp32"
    with pytest.raises(SystemExit):
        p.parse_args(["--gemm_precision", "fp8"])
