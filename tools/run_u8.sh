timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
python tools/e2e_probe.py 2>&1 | tail -8
python bench.py --no-dense --no-retrain --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -2 gpurun_out/bench_e2e.err
