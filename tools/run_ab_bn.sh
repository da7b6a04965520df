for bn in 0 176 192 208 224 240; do MCB200_CONV_BN_BIG=$bn timeout 150 python tools/bench_layers.py bn$bn; done > gpurun_out/ab_bn.jsonl 2>gpurun_out/ab_bn.err
tail -3 gpurun_out/ab_bn.err
