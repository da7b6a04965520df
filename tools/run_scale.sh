S="64 8 104 104 4 1  64 17 104 104 4 1  64 40 104 104 4 1 64 78 52 52 16 1 64 69 26 26 145 3 64 818 13 13 1006 3"
for m in 8 64; do for f in 0 1; do MCB200_TMAP_FULL_LD=$f LD_MULT=$m timeout 200 python tools/bench_single_conv.py $S; done; done > gpurun_out/scale.json 2> gpurun_out/scale.err; tail -2 gpurun_out/scale.err
