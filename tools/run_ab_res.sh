timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
timeout 150 python tools/bench_layers.py res > gpurun_out/ab_res.jsonl 2>gpurun_out/ab_res.err
timeout 150 python tools/bench_layers.py dense res >> gpurun_out/ab_res.jsonl 2>>gpurun_out/ab_res.err
tail -3 gpurun_out/ab_res.err
