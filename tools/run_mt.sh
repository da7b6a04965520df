for rep in 1 2; do
timeout 150 python tools/bench_layers.py base
MCB200_CONV_MINTILES=1 timeout 150 python tools/bench_layers.py mt1
done > gpurun_out/ab_mt.jsonl 2>gpurun_out/ab_mt.err
timeout 150 python tools/bench_layers.py dense base >> gpurun_out/ab_mt.jsonl 2>>gpurun_out/ab_mt.err
MCB200_CONV_MINTILES=1 timeout 150 python tools/bench_layers.py dense mt1 >> gpurun_out/ab_mt.jsonl 2>>gpurun_out/ab_mt.err
tail -2 gpurun_out/ab_mt.err
