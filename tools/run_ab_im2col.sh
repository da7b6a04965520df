timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_eval.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
MCB200_NO_BRANCH_STREAM=1 timeout 150 python tools/bench_layers.py im2col > gpurun_out/ab_im2col.jsonl 2>gpurun_out/ab_im2col.err
timeout 150 python tools/bench_layers.py im2col+branch >> gpurun_out/ab_im2col.jsonl 2>>gpurun_out/ab_im2col.err
MCB200_NO_BRANCH_STREAM=1 timeout 150 python tools/bench_layers.py dense im2col >> gpurun_out/ab_im2col.jsonl 2>>gpurun_out/ab_im2col.err
timeout 150 python tools/bench_layers.py dense im2col+branch >> gpurun_out/ab_im2col.jsonl 2>>gpurun_out/ab_im2col.err
tail -3 gpurun_out/ab_im2col.err
