set -x
# 0. GPU tests; 1. the bench itself (no profiler) -> the numbers; 2. launch list of the same command; 3. ncu --set full
# of the conv GEMM kernels of one eager forward (shrunk net, batch 64): 17 launches (single-CTA + CTA-pair kernels),
# and of the three conv_thin_kernel launches of the same forward
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_short.log 2>&1
python tools/profile_forward.py 64 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 17 -c 17 -f -o gpurun_out/prof_conv_r1 python tools/profile_forward.py 64 > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log
ncu --set full --clock-control none --import-source on -k regex:conv_thin -s 3 -c 3 -f -o gpurun_out/prof_thin_r2 python tools/profile_forward.py 64 > gpurun_out/ncu_thin.log 2>&1
tail -2 gpurun_out/ncu_thin.log
# 4. launch list of the retrain step (BASELINE configs[2])
python tools/profile_train.py > gpurun_out/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/plain_train.log
