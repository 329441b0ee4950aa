"""Case tables shared by make_golden.py and the tests (no reference access)."""
VOC_NMS = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "table", "dog",
           "horse", "motorbike", "person", "plant", "sheep", "sofa", "train", "television"]

MERGE_CASES = {
    "plain": [["cat", "dog", "person"], ["bus"]],                       # #pieces == C -> slice path
    "split_mid": [["aeroplane", "dog"], ["cat", "motorbike", "sofa"]],  # split word followed by a word
    "split_last": [["dog", "aeroplane"], ["pottedplant"]],              # split word LAST: summed, not averaged
    "mixed_len": [["cat"], ["television", "cow", "bicycle", "boat"]],   # shorter caption keeps its SEP row
}

# tag: (coco driver?, data_type, class lists, drop_iter)
DRIVER_CASES = {
    "voc_r4": (False, "voc", [["cat", "aeroplane"], ["dog"], ["bus", "car", "person"]], 4),
    "voc_r1": (False, "voc", [["bicycle", "aeroplane"], ["dog"]], 1),
    "voc_alias": (False, "voc", [["bicycle", "aeroplane"], ["bird", "bicycle", "aeroplane"]], 2),
    "ade_r2": (False, "ade20k", [["cat", "dog"], ["bus", "car", "person", "boat"]], 2),
    # the COCO driver (PnP_OVSS_0514_updated_segmentation_coco.py): sparse category ids, Scale_0_1 on both paths,
    # round-0 pass only when drop_iter < 3, n_class 91 / 183
    "coco_obj_r2": (True, "coco_object", [["cat", "aeroplane"], ["dog", "bus", "car"]], 2),
    "coco_stuff_r4": (True, "coco_stuff", [["cat", "dog"], ["bus", "car", "person", "boat"]], 4),
}
