"""Offline scorer with the reference's file layout (Calculate_mIoU.py:204-256): sums every confusion-matrix .npy under
{save_path}/all_drop_hist_with_filtered_caption/ (or another sub-directory) and prints the statistics.  With the
all-reduce of pipeline.allreduce_hist there is one file per run instead of one per batch and rank, but directories
written by the reference itself are read just the same.

    python -m pnp_ovss_b200.calculate_miou --save_path out/ [--subdir hist_withfiltered_caption]
"""
import argparse
import os

import numpy as np

from .reference_api import metrics_from_hist


def sum_hist_dir(path):
    total = None
    n_files = 0
    for fn in sorted(os.listdir(path)):
        if fn.endswith(".npy"):
            h = np.load(os.path.join(path, fn))
            total = h.astype(np.float64) if total is None else total + h
            n_files += 1
    if total is None:
        raise FileNotFoundError("no .npy confusion matrices under %s (drop_iter 1 runs write hist_withfiltered_caption/ only)" % path)
    return total, n_files


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--save_path", required=True)
    ap.add_argument("--subdir", default="all_drop_hist_with_filtered_caption")
    a = ap.parse_args(argv)
    hist, n_files = sum_hist_dir(os.path.join(a.save_path, a.subdir))
    table, _ = metrics_from_hist(hist)
    print("files %d  pixels %d" % (n_files, int(hist.sum())))
    for k in ("Pixel Accuracy", "Mean Accuracy", "Frequency Weighted IoU", "Mean IoU"):
        print("%s: %.6f" % (k, table[k]))
    return table, hist


if __name__ == "__main__":
    main()
