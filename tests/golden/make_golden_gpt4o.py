#!/usr/bin/env python
"""Generate tests/golden/gpt4o_golden.json by running the REFERENCE's own Load_predicted_classes (lifted with ast from
PnP_OVSS_0514_updated_segmentation.py:726-787, unmodified) on the GPT-4o result files it ships
(/root/reference/GPT4o_classification/*.json).  Needs /root/reference; the tests do not.  Only the raw answer strings
(public dataset image ids -> GPT-4o text) and the parsed class indices are stored."""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import DRV, DRVC, REF, base_ns, lift  # noqa: E402

N_PER_SET = 120


def main():
    ns = lift(DRV, ["Load_predicted_classes"], base_ns())
    out = {}
    for data_type, fname, n_names in (("voc", "voc_classification_noboundary.json", 20), ("psc", "psc_classification_noboundary.json", 59),
                                      ("ade20k", "ade20k_classification_noboundary.json", 150)):
        raw = json.load(open(os.path.join(REF, "GPT4o_classification", fname)))
        nms = ["name%03d" % i for i in range(n_names)]
        keys = sorted(raw)
        # every answer with an unusual layout plus an evenly spaced sample of the rest
        odd = [k for k in keys if "\n" in raw[k] or raw[k].count("[") != 2]
        step = max(1, len(keys) // N_PER_SET)
        chosen = sorted(set(odd[:40] + keys[::step]))
        args = types.SimpleNamespace(home_dir=REF, data_type=data_type)
        cases = []
        for k in chosen:
            img_id = k[len("ADE_val_"):].lstrip("0") if data_type == "ade20k" else k
            if data_type == "ade20k" and "ADE_val_" + img_id.rjust(8, "0") != k:
                continue
            try:
                best, cls, caps = ns["Load_predicted_classes"](args, nms, [], [], [], None, [img_id], 0, None)
            except Exception as exc:  # the reference itself fails on this answer: record that
                cases.append({"raw": raw[k], "error": type(exc).__name__})
                continue
            cases.append({"raw": raw[k], "best_class_idx": best[0], "caption": caps[0]})
        out[data_type] = {"n_names": n_names, "cases": cases}
    # the COCO driver's variant (DRVC:855-963): category ids instead of positions, a "no probabilities" branch
    thing_ids = [i for i in range(1, 91) if i not in (12, 26, 29, 30, 45, 66, 68, 69, 71, 83)]
    stuff_ids = list(range(92, 183))
    nsc = lift(DRVC, ["Load_predicted_classes"], base_ns())
    for data_type, fname, ids in (("coco_object", "coco_object_classification_noboundary.json", thing_ids),
                                  ("coco_stuff", "coco_stuff_classification_noboundary.json", thing_ids + stuff_ids)):
        raw = json.load(open(os.path.join(REF, "GPT4o_classification", fname)))
        cats = [{"id": i} for i in ids]
        nms = ["name%03d" % i for i in range(len(ids))]
        keys = sorted(raw)
        odd = [k for k in keys if "\n" in raw[k] or raw[k].count("[") != 2]
        step = max(1, len(keys) // N_PER_SET)
        chosen = sorted(set(odd[:40] + keys[::step]))
        args = types.SimpleNamespace(home_dir=REF, data_type=data_type)
        cases = []
        for k in chosen:
            try:
                best, cls, caps = nsc["Load_predicted_classes"](args, nms, cats, [], [], [], [None], [k], 0, None)
            except Exception as exc:
                cases.append({"raw": raw[k], "error": type(exc).__name__})
                continue
            cases.append({"raw": raw[k], "best_class_idx": best[0], "caption": caps[0]})
        out[data_type] = {"n_names": len(ids), "cat_ids": ids, "cases": cases}
    json.dump(out, open(os.path.join(HERE, "gpt4o_golden.json"), "w"), indent=0)
    print({k: len(v["cases"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
