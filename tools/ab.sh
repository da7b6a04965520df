#!/bin/bash
# Same-lease A/B of the forward under MCB200_* environment switches (box-to-box variance is ~3 %, so both arms must run
# inside ONE gpurun call).  Usage:  gpurun -- 'bash tools/ab.sh "MCB200_CONV_2CTA=0" "MCB200_CONV_STAGE_OUT=0" ...'
# Every argument is one variant (space-separated VAR=value pairs; "" = defaults); each runs twice on the shrunk and once on
# the dense network.  Result lines: gpurun_out/ab.jsonl (tools/bench_layers.py format).
# Switches: MCB200_CONV_2CTA (0 off, 2 force), MCB200_CONV_CTAS (1..3), MCB200_CONV_MINTILES, MCB200_CONV_BN_BIG=<bn>, MCB200_CONV_BN_MID=<bn>,
# MCB200_CONV_STAGE_OUT (0 off, 2 all wide tiles), MCB200_SHARE_DX (0 off, 2 only 64-wide k-blocks), MCB200_CONV_ACC=<n>,
# MCB200_CONV_RESIDENT=1, MCB200_PITCH=0, MCB200_CONV_TRACE=1.
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
for rep in 1 2; do
  env python tools/bench_layers.py base >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
  for v in "$@"; do env $v python tools/bench_layers.py "$v" >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err; done
done
env python tools/bench_layers.py dense base >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
for v in "$@"; do env $v python tools/bench_layers.py dense "$v" >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err; done
python - <<'PY'
import json
rows = [json.loads(l) for l in open('gpurun_out/ab.jsonl')]
for dense in (False, True):
    rs = [r for r in rows if r['dense'] == dense]
    print('dense' if dense else 'shrunk', [(r['tag'], r['ms_per_step'], r['img_per_s']) for r in rs])
PY
