#!/bin/bash
# BASELINE.json configs[2] and configs[4] data-parallel over 2 GPUs through the driver (mp.spawn, one process per GPU,
# one NCCL all-reduce of the confusion matrix at the end); 70 images per rank = 2 batches of 35, the second one warm.
COMMON="--batch_size 35 --synthetic_images 140 --max_att_block_num 8 --prune_att_head 9 --drop_iter 4 --del_patch_num sort_thresh005 --sort_threshold 0.05 --threshold 0.15 --postprocess blur+crf --world_size 2 --save_path /tmp/pnp_out2 --master_port 29611"
echo "== cfg2 ade20k 150 classes @336, 2 GPUs"
python -m pnp_ovss_b200.driver --data_type ade20k --img_size 336 --synthetic_classes 150 $COMMON 2>&1 | grep -E "Time|images|rror"
echo "== cfg4 coco_object 80 classes @448, 2 GPUs"
python -m pnp_ovss_b200.driver --data_type coco_object --img_size 448 --synthetic_classes 80 $COMMON 2>&1 | grep -E "Time|images|rror"
