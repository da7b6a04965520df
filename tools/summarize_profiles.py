"""Turn the raw gpurun_out/ captures of tools/run_final_evidence.sh into the tracked summaries under profiles/:
  r2_launches_bench_final.csv (+ _summary.csv)   per-kernel totals of the launch list of the bench command
  r2_ncu_full_conv_gemm_summary.csv              one row per conv GEMM launch of a forward (ncu --set full)
  r2_ncu_full_conv_gemm.json                     DRAM traffic of those launches (bench.py's roofline.traffic)
Run here (no GPU needed): python tools/summarize_profiles.py"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
GP = os.path.join(ROOT, 'gpurun_out')


def launch_summary():
    src = os.path.join(GP, 'launches_bench_final.csv')
    rows = list(csv.reader(open(src, errors='ignore')))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[h]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[h + 2:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:110]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e6  # ns -> ms
    tot = sum(a[1] for a in agg.values())
    shutil.copy(src, os.path.join(OUT, 'r2_launches_bench_final.csv'))
    with open(os.path.join(OUT, 'r2_launches_bench_final_summary.csv'), 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_ms', 'share_of_captured'])
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, round(ms, 4), round(ms / tot, 4)])
    print('launch list:', len(agg), 'kernels,', round(tot, 3), 'ms captured')


def ncu_summary():
    rep = os.path.join(GP, 'prof_conv_r1.ncu-rep')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ['ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
            'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
            'launch__cluster_size', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(os.path.join(OUT, 'r2_ncu_full_conv_gemm_summary.csv'), 'w') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i].split('(CUtensorMap')[0] if hdr[i] == 'Kernel Name' else r[i] for i in idx])
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')

    def mb(v, unit):
        v = float(v.replace(',', ''))
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[unit]
    rd = sum(mb(r[ir], units[ir]) for r in rows[2:])
    wr = sum(mb(r[iw], units[iw]) for r in rows[2:])
    n = len(rows) - 2
    json.dump({"kernel": "conv_gemm_tcgen05_kernel / conv_gemm_tcgen05_pair_kernel",
               "source": "ncu --set full --clock-control none, tools/run_final_evidence.sh (shrunk bench workload, batch "
                         "64, the %d conv GEMM launches of one forward)" % n,
               "launches": n, "dram_read_mbytes_total": rd, "dram_write_mbytes_total": wr,
               "dram_bytes_per_launch": (rd + wr) * 1e6 / n},
              open(os.path.join(OUT, 'r2_ncu_full_conv_gemm.json'), 'w'), indent=1)
    print('ncu full:', n, 'launches,', round(rd + wr, 1), 'MB DRAM traffic')


def train_summary(steps=7):
    """per-kernel totals of one retrain step from the launch list of tools/profile_train.py (7 steps captured)."""
    src = os.path.join(GP, 'launches_train.csv')
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, errors='ignore')))
    h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[h]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[h + 2:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki].split('(')[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e6
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, 'r2_launches_train_step_summary.csv'), 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches_per_step', 'ms_per_step', 'share'])
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, round(n / steps, 1), round(ms / steps, 4), round(ms / tot, 4)])
        w.writerow(['TOTAL (one retrain step, batch 64, ncu gpu__time_duration, cold cache, serialised)', '',
                    round(tot / steps, 3), 1.0])
    print('train step:', round(tot / steps, 3), 'ms of kernels per step')


if __name__ == '__main__':
    train_summary()
    launch_summary()
    ncu_summary()
    shutil.copy(os.path.join(GP, 'bench_final.json'), os.path.join(OUT, 'r2_bench_final.json'))
