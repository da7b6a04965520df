timeout 150 python tools/bench_layers.py hint > gpurun_out/ab_spin.jsonl 2>gpurun_out/ab_spin.err
timeout 150 python tools/bench_layers.py dense hint >> gpurun_out/ab_spin.jsonl 2>>gpurun_out/ab_spin.err
cp modelcompression_b200/libmcb200.so /tmp/orig.so; cp tools/variant/libmcb200_spin.so modelcompression_b200/libmcb200.so
timeout 150 python tools/bench_layers.py spin >> gpurun_out/ab_spin.jsonl 2>>gpurun_out/ab_spin.err
timeout 150 python tools/bench_layers.py dense spin >> gpurun_out/ab_spin.jsonl 2>>gpurun_out/ab_spin.err
cp /tmp/orig.so modelcompression_b200/libmcb200.so
tail -3 gpurun_out/ab_spin.err
