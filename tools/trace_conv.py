"""Per-CTA timeline of the single-CTA conv GEMM kernel (mc_debug_conv_trace) on given shapes: B C H W O k ...
Slots: 0 entry, 1 set-up done, 3 producer finished issuing, 4 first operands landed, 5+i MMAs of tile i issued,
12+2i accumulator of tile i ready (epilogue view), 13+2i epilogue of tile i done, 30 exit.  Times in us from the first CTA's entry."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200 import _lib

dev = 'cuda:0'
args = [int(a) for a in sys.argv[1:]]
lib = _lib.load()
for i in range(0, len(args), 6):
    B, C, H, W, O, k = args[i:i + 6]
    conv = mc.MaskedConv2d(C, O, k, 1, (k - 1) // 2, bias=False).to(dev)
    kb = 32 if C <= 32 else 64
    ld_in = (C + 7) // 8 * 8
    if C > 16:
        ld_in = 32 if C <= 32 else ((C + 63) // 64 * 64 if (C + 63) // 64 * 64 * 4 <= ld_in * 5 else ld_in)
    Kc, Npad, ld_out = (C + kb - 1) // kb * kb, (O + 15) // 16 * 16, (O + 7) // 8 * 8
    rows = B * (H + 1) * (W + 1)
    xin = torch.randn(rows, ld_in, device=dev).to(torch.bfloat16)
    wpack = torch.empty(Npad, k * k * Kc, dtype=torch.bfloat16, device=dev)
    scale, shift = torch.ones(Npad, device=dev), torch.zeros(Npad, device=dev)
    yb = torch.empty(rows, ld_out, dtype=torch.bfloat16, device=dev)
    s = _lib.stream_ptr()
    w = conv.weight.data.float().contiguous()
    _lib.check(lib.mc_pack_conv_weights(w.data_ptr(), None, O, C, k, None, O, None, C, wpack.data_ptr(), Npad, Kc, s), "pack")
    d = _lib.mc_conv_desc()
    d.d_in, d.d_wpack, d.d_scale, d.d_shift, d.d_out = xin.data_ptr(), wpack.data_ptr(), scale.data_ptr(), shift.data_ptr(), yb.data_ptr()
    d.B, d.H, d.W, d.Cin, d.Cin_ld, d.N, d.Npad = B, H, W, C, ld_in, O, Npad
    d.in_cols = ld_in
    d.ksize, d.leaky, d.epi_mode, d.ldc, d.ch_off, d.block_n, d.stages, d.block_k = k, 1, _lib.MC_EPI_PNHWC, ld_out, 0, 0, 0, kb
    need = int(lib.mc_workspace_bytes_conv_fwd(ctypes.byref(d)))
    if need and not os.environ.get('NO_WS'):
        ws = torch.zeros(need, dtype=torch.uint8, device=dev)
        d.d_ws, d.ws_bytes = ws.data_ptr(), need
    for _ in range(3):
        _lib.check(lib.mc_conv_fwd(ctypes.byref(d), s), "conv")
    torch.cuda.synchronize()
    buf = torch.zeros(32 * 600, dtype=torch.int64, device=dev)
    lib.mc_debug_conv_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.mc_conv_fwd(ctypes.byref(d), s), "conv")
    e1.record()
    torch.cuda.synchronize()
    lib.mc_debug_conv_trace(None)
    info = (ctypes.c_int * 8)()
    lib.mc_conv_last_plan(info)
    grid = info[6]
    t = buf.view(600, 32)[:grid].cpu()
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    rel = (t - t0).double() / 1e3
    rel[t == 0] = float('nan')
    print("shape", (B, C, H, W, O, k), "plan [pair,bn,ctas,res,share,stages,grid,kblk]", list(info), "event us %.1f" % (e0.elapsed_time(e1) * 1e3))
    names = {0: 'entry', 1: 'setup', 3: 'prodEnd', 4: 'opnd0', 30: 'exit'}
    for j in range(6):
        names[5 + j] = 'mma%d' % j
    for j in range(8):
        names[12 + 2 * j] = 'acc%d' % j
        names[13 + 2 * j] = 'epi%d' % j
    for cta in sorted(set([0, 1, grid // 2, grid - 1])):
        row = rel[cta]
        print("  cta %3d:" % cta, " ".join("%s=%.1f" % (names[sl], float(row[sl])) for sl in sorted(names) if row[sl] == row[sl]))
    ex = rel[:, 30]
    en = rel[:, 0]
    print("  entry min/max %.1f/%.1f  exit min/median/max %.1f/%.1f/%.1f" % (float(en.min()), float(en.max()), float(ex.min()), float(ex.median()), float(ex.max())))
