#!/bin/bash
# BASELINE.json configs[2..4] end to end through the driver (model passes + DropOut + both scored maps) on ONE GPU,
# 2 batches of 35 synthetic images each; the second batch's "Time: batch" line is the warm one.
COMMON="--batch_size 35 --synthetic_images 70 --max_att_block_num 8 --prune_att_head 9 --drop_iter 4 --del_patch_num sort_thresh005 --sort_threshold 0.05 --threshold 0.15 --postprocess blur+crf --world_size 1 --save_path /tmp/pnp_out"
echo "== cfg2 ade20k 150 classes @336"
python -m pnp_ovss_b200.driver --data_type ade20k --img_size 336 --synthetic_classes 150 $COMMON 2>&1 | grep -E "Time|images|Error|error" 
echo "== cfg3 coco_stuff 171 classes @336, CRF at 512x512"
python -m pnp_ovss_b200.driver --data_type coco_stuff --img_size 336 --synthetic_classes 171 --synthetic_gt_size 512 $COMMON 2>&1 | grep -E "Time|images|Error|error"
echo "== cfg4 coco_object 80 classes @448"
python -m pnp_ovss_b200.driver --data_type coco_object --img_size 448 --synthetic_classes 80 $COMMON 2>&1 | grep -E "Time|images|Error|error"
