set -x
# dominant kernel of the bench workload (shrunk net, batch 64, eager): the 18 tcgen05 conv launches of one forward
python tools/profile_forward.py 64 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 18 -c 18 -f -o gpurun_out/prof_conv_r1 python tools/profile_forward.py 64 > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log
