for bn in 0 176 112 208 128; do MCB200_CONV_BN_BIG=$bn python tools/bench_layers.py bn$bn; done > gpurun_out/ab_bn.jsonl 2>gpurun_out/ab_bn.err
tail -3 gpurun_out/ab_bn.err
