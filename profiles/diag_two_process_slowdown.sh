COMMON="--data_type ade20k --img_size 336 --synthetic_classes 150 --batch_size 35 --max_att_block_num 8 --prune_att_head 9 --drop_iter 4 --del_patch_num sort_thresh005 --sort_threshold 0.05 --threshold 0.15 --postprocess blur+crf --save_path /tmp/pnp_out2 --master_port 29611"
echo "== (a) world 1"; python -m pnp_ovss_b200.driver $COMMON --world_size 1 --synthetic_images 70 2>&1 | grep -E "Time: batch"
echo "== (b) world 2, OMP_NUM_THREADS=1"; OMP_NUM_THREADS=1 python -m pnp_ovss_b200.driver $COMMON --world_size 2 --synthetic_images 140 2>&1 | grep -E "Time: batch"
echo "== (c) world 2 default + clocks"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 500 > /tmp/smi.log &
SMI=$!
python -m pnp_ovss_b200.driver $COMMON --world_size 2 --synthetic_images 140 2>&1 | grep -E "Time: batch"
kill $SMI
sort /tmp/smi.log | uniq -c | sort -rn | head -8
echo "== (d) two independent single-GPU processes"
CUDA_VISIBLE_DEVICES=0 python -m pnp_ovss_b200.driver $COMMON --world_size 1 --synthetic_images 70 2>&1 | grep -E "Time: batch" | sed 's/^/gpu0 /' &
P0=$!
CUDA_VISIBLE_DEVICES=1 python -m pnp_ovss_b200.driver $COMMON --world_size 1 --synthetic_images 70 --save_path /tmp/pnp_out3 2>&1 | grep -E "Time: batch" | sed 's/^/gpu1 /'
wait $P0
