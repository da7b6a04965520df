for rep in 1 2; do
MCB200_CONV_ACC=2 MCB200_PITCH=0 timeout 150 python tools/bench_layers.py base
MCB200_PITCH=0 timeout 150 python tools/bench_layers.py acc
timeout 150 python tools/bench_layers.py acc+pitch
MCB200_CONV_ACC=2 timeout 150 python tools/bench_layers.py pitch
done > gpurun_out/ab_acc.jsonl 2>gpurun_out/ab_acc.err
MCB200_CONV_ACC=2 MCB200_PITCH=0 timeout 150 python tools/bench_layers.py dense base >> gpurun_out/ab_acc.jsonl 2>>gpurun_out/ab_acc.err
timeout 150 python tools/bench_layers.py dense acc+pitch >> gpurun_out/ab_acc.jsonl 2>>gpurun_out/ab_acc.err
tail -2 gpurun_out/ab_acc.err
