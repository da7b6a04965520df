for rep in 1 2; do
MCB200_CONV_STAGE_OUT=0 timeout 150 python tools/bench_layers.py base
timeout 150 python tools/bench_layers.py stage
done > gpurun_out/ab_stage.jsonl 2>gpurun_out/ab_stage.err
for rep in 1 2; do
MCB200_CONV_STAGE_OUT=0 timeout 150 python tools/bench_layers.py dense base >> gpurun_out/ab_stage.jsonl 2>>gpurun_out/ab_stage.err
timeout 150 python tools/bench_layers.py dense stage >> gpurun_out/ab_stage.jsonl 2>>gpurun_out/ab_stage.err
done
tail -2 gpurun_out/ab_stage.err
