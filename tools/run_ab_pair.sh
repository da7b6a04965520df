for m in 0 1 2; do MCB200_CONV_2CTA=$m timeout 150 python tools/bench_layers.py pair$m; echo "rc=$?" >&2; done > gpurun_out/ab_pair.jsonl 2>gpurun_out/ab_pair.err
tail -5 gpurun_out/ab_pair.err
for m in 0 1; do MCB200_CONV_2CTA=$m timeout 150 python tools/bench_layers.py dense pair$m; echo "rc=$?" >&2; done >> gpurun_out/ab_pair.jsonl 2>>gpurun_out/ab_pair.err
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
