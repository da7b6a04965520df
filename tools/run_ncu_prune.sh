set -x
# ncu --set full of the two pruner kernels (one launch each, after warm-up launches)
python tools/prune_launches.py > gpurun_out/plain_prune.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"weight_prune_kernel|filter_prune_fused_kernel" -s 4 -c 2 -f -o gpurun_out/prof_prune_r2 python tools/prune_launches.py > gpurun_out/ncu_prune.log 2>&1
tail -2 gpurun_out/ncu_prune.log
