S="64 64 52 52 64 3  64 48 104 104 32 3  64 64 26 26 96 3  64 40 52 52 80 3"
for r in 0 1; do MCB200_CONV_RESIDENT=$r timeout 120 python tools/bench_single_conv.py $S; done > gpurun_out/ab_res2.jsonl 2> gpurun_out/ab_res2.err
tail -2 gpurun_out/ab_res2.err
timeout 150 python tools/bench_layers.py final > gpurun_out/ab_res.jsonl 2>gpurun_out/ab_res.err
timeout 150 python tools/bench_layers.py dense final >> gpurun_out/ab_res.jsonl 2>>gpurun_out/ab_res.err
