timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
for rep in 1 2; do
MCB200_SHARE_DX=2 timeout 150 python tools/bench_layers.py base
timeout 150 python tools/bench_layers.py share32
done > gpurun_out/ab_share32.jsonl 2>gpurun_out/ab_share32.err
for rep in 1 2; do
MCB200_SHARE_DX=2 timeout 150 python tools/bench_layers.py dense base >> gpurun_out/ab_share32.jsonl 2>>gpurun_out/ab_share32.err
timeout 150 python tools/bench_layers.py dense share32 >> gpurun_out/ab_share32.jsonl 2>>gpurun_out/ab_share32.err
done
tail -2 gpurun_out/ab_share32.err
