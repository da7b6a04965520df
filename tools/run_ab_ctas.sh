set -x
python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
for c in 1 2 3; do MCB200_CONV_CTAS=$c python tools/bench_layers.py ctas$c; done > gpurun_out/ab_ctas.jsonl 2>gpurun_out/ab_ctas.err
python tools/bench_layers.py auto >> gpurun_out/ab_ctas.jsonl 2>>gpurun_out/ab_ctas.err
for c in 1 2 3; do MCB200_CONV_CTAS=$c python tools/bench_layers.py dense ctas$c; done >> gpurun_out/ab_ctas.jsonl 2>>gpurun_out/ab_ctas.err
python tools/bench_layers.py dense auto >> gpurun_out/ab_ctas.jsonl 2>>gpurun_out/ab_ctas.err
tail -3 gpurun_out/ab_ctas.err
