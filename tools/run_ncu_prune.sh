set -x
python tools/profile_prune.py 1 > gpurun_out/plain_prune.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"weight_prune_kernel|filter_sumsq_kernel|filter_mask_fill_kernel|filter_finish_kernel" -s 6 -c 6 -f -o gpurun_out/prof_prune_r1 python tools/profile_prune.py 1 > gpurun_out/ncu_prune.log 2>&1
tail -2 gpurun_out/ncu_prune.log
