timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1; tail -3 gpurun_out/ab_pytest.log
MCB200_CONV_STAGE_OUT=2 timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/ab_pytest2.log 2>&1; tail -3 gpurun_out/ab_pytest2.log
for rep in 1 2; do
timeout 150 python tools/bench_layers.py base
MCB200_CONV_STAGE_OUT=2 timeout 150 python tools/bench_layers.py chunked
done > gpurun_out/ab_stage.jsonl 2>gpurun_out/ab_stage.err
for rep in 1 2; do
timeout 150 python tools/bench_layers.py dense base >> gpurun_out/ab_stage.jsonl 2>>gpurun_out/ab_stage.err
MCB200_CONV_STAGE_OUT=2 timeout 150 python tools/bench_layers.py dense chunked >> gpurun_out/ab_stage.jsonl 2>>gpurun_out/ab_stage.err
done
tail -2 gpurun_out/ab_stage.err
python bench.py --no-dense --no-retrain --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; tail -2 gpurun_out/bench_e2e.err
