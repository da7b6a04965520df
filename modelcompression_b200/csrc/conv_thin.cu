// conv_thin.cu — CUDA-core convolution for the DEGENERATE layers of a filter-pruned network (sm_100a).
//
// Reference op: MaskedConv2d + BatchNorm2d(eval) + LeakyReLU(0.1) [+ MaxPool2d(2,2)] of src/nets.py:779-821 on a layer
// whose surviving channel counts are tiny — after 40 % filter pruning of the seed-0 network conv2 is 4 -> 1 channels,
// conv3 1 -> 17, conv4 17 -> 4 (1x1), conv5 4 -> 11.  Such a layer is 36..400 multiply-adds per output pixel: not a
// contraction a 128 x N x 64 tensor-core tile can be filled with, but a stream over 0.7 M pixels that is bound by
// reading the input once and writing the output once.  On the tcgen05 kernels these four layers took 25-37 us each
// (latency chain TMA -> MMA -> commit -> epilogue per 128-pixel tile) against an HBM floor of 3-9 us.
//
// Formulation: one thread = one output pixel (or one 2x2 pool window).  The thread loads its 3x3 (4x4 for a pool window)
// neighbourhood of PNHWC pixels with vector loads (the zero pad column / pad line of the layout supply the borders, so
// there is no bounds logic beyond "row >= 0"), converts bf16 -> fp32 by a shift, and runs a FULLY UNROLLED multiply-add
// nest whose weights are kernel PARAMETERS: the compiler emits FFMA with a constant-bank operand (c[0x0][imm]), i.e. no
// instruction and no register is spent on a weight.  Scale/shift (folded BatchNorm), leaky-ReLU and the 2x2 max are
// applied in registers; optionally the 1x1 layer BEHIND the 3x3 layer is applied to the bf16-rounded activations in the
// same thread (N2T > 0), so the intermediate tensor (45 MB for conv3 at batch 64) is never stored.
//
// Numerics: activations are the stored bf16 values, weights are bf16-rounded by the caller (the same operands the tensor
// core kernels see), accumulation is fp32 FMA in tap-major order.
#include "common.cuh"

namespace {

constexpr int TH_MAXW = 1536;  // main weights: taps * (2*CW) * NT floats
constexpr int TH_MAXN = 24;
constexpr int TH_MAXN2 = 8;
constexpr int TH_THREADS = 128;  // small blocks: the widest instances hold 128 registers, and a pooled 52x52 layer has only 5.4 K warps

struct ThinParams {
  float w[TH_MAXW];             // [(tap * 2*CW + c) * NT + n], zero padded
  float sc[TH_MAXN], sh[TH_MAXN];
  float w2[TH_MAXN * TH_MAXN2];  // fused 1x1: [n * N2T + o]
  float sc2[TH_MAXN2], sh2[TH_MAXN2];
};

// The one rule that decides which template instances exist (launch table) and which shapes are accepted (geometry).
constexpr bool thin_valid(int ks, int cw, int nt, int pool, int n2t) {
  if (ks == 3) {
    if (!(cw == 1 || cw == 2 || cw == 4)) return false;
    if (pool) return n2t == 0 && nt <= 16 && 72 * cw * nt <= 1800;
    if (n2t) return nt >= 8 && 18 * cw * nt <= 640;
    return 18 * cw * nt <= 640;
  }
  if (ks == 1) return !pool && n2t == 0 && (cw == 4 || cw == 8 || cw == 12 || cw == 16) && nt <= 8;
  return false;
}

__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }

template <int CW>
__device__ __forceinline__ void load_px(const uint32_t* __restrict__ p, bool ok, uint32_t (&w)[CW]) {
  if (!ok) {
#pragma unroll
    for (int i = 0; i < CW; ++i) w[i] = 0u;
    return;
  }
  if constexpr (CW % 4 == 0) {
#pragma unroll
    for (int q = 0; q < CW / 4; ++q) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + q);
      w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
  } else if constexpr (CW == 2) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = v.x; w[1] = v.y;
  } else {
    w[0] = __ldg(p);
  }
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// out row: NV values (NV <= 24), written as ceil(NV/8) 16-byte pieces; channels beyond NV inside the last piece are zero
template <int NV>
__device__ __forceinline__ void store_row(__nv_bfloat16* __restrict__ dst, const float (&v)[NV]) {
  constexpr int G = (NV + 7) / 8;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c0 = g * 8 + 2 * j;
      const float a = c0 < NV ? v[c0 < NV ? c0 : 0] : 0.f;
      const float b = c0 + 1 < NV ? v[c0 + 1 < NV ? c0 + 1 : 0] : 0.f;
      w[j] = pack_bf2(a, b);
    }
    *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ODD: the input has an odd channel count, i.e. the high half of the last word read per pixel is not a real channel
// (zero weights): its multiply-adds are not issued (conv3 of the bench network has ONE input channel).
template <int KS, int CW, int NT, bool POOL, int N2T, bool ODD>
__global__ void __launch_bounds__(TH_THREADS)
conv_thin_kernel(const __grid_constant__ ThinParams P, const uint32_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                 int H, int W, int in_ldw, int ldc, int leaky, int leaky2, unsigned total) {
  constexpr int CT = 2 * CW;
  constexpr int HALO = KS / 2;
  const unsigned i = blockIdx.x * (unsigned)TH_THREADS + threadIdx.x;
  if (!POOL && i >= total) return;  // (pool windows exchange pixels by shuffle: every lane stays until they are done)
  float fin[NT];
  long long out_row;

  if constexpr (!POOL) {
    const unsigned Wp = (unsigned)W + 1u;
    const unsigned line = i / Wp;
    const unsigned x = i - line * Wp;
    const unsigned y = line % ((unsigned)H + 1u);
    if (x == (unsigned)W || y == (unsigned)H) return;  // pad column / pad line: stay zero
    uint32_t px[KS * KS][CW];
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) {
      const int dy = t / KS - HALO, dx = t % KS - HALO;
      const long long r = (long long)i + dy * (int)Wp + dx;
      load_px<CW>(in + r * in_ldw, r >= 0, px[t]);
    }
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.f;
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) {
#pragma unroll
      for (int cw = 0; cw < CW; ++cw) {
        const float a0 = bf_lo(px[t][cw]), a1 = bf_hi(px[t][cw]);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          acc[n] = fmaf(a0, P.w[(t * CT + 2 * cw) * NT + n], acc[n]);
          if (!(ODD && cw == CW - 1)) acc[n] = fmaf(a1, P.w[(t * CT + 2 * cw + 1) * NT + n], acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float v = fmaf(acc[n], P.sc[n], P.sh[n]);
      if (leaky) v = leaky01(v);
      fin[n] = v;
    }
    out_row = (long long)i;
  } else {
    // One thread per pool window: a 4x4 pixel neighbourhood (rows 2yo-1..2yo+2, columns 2xo-1..2xo+2).  Consecutive
    // lanes own consecutive windows of a line, so a thread LOADS only its own two columns (2xo, 2xo+1) of every row and
    // takes column 2xo-1 from the lane below and column 2xo+2 from the lane above by shuffle (lane 0 / lane 31 load
    // them): 8 instead of 16 loads per thread — the per-thread stride of two pixels makes every load instruction touch
    // every 32-byte sector of its span, and the L1 wavefronts were the limiter of the 4 -> 1 layer.  Threads on the pad
    // column / pad line / beyond the end load nothing, hold zeros (what the pad positions contain) and do not store.
    const unsigned Wo = (unsigned)W / 2u, Ho = (unsigned)H / 2u;
    const unsigned Wop = Wo + 1u;
    const unsigned line = i / Wop;
    const unsigned xo = i - line * Wop;
    const unsigned yo = line % (Ho + 1u);
    const unsigned b = line / (Ho + 1u);
    const bool pad = i >= total || xo == Wo || yo == Ho;
    const int Wp = W + 1;
    const long long base = ((long long)b * (H + 1) + 2 * yo) * Wp + 2 * xo;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t px[16][CW];
    // all loads first (8 per thread, 4 more on the two edge lanes), so that they are in flight together ...
#pragma unroll
    for (int row = 0; row < 4; ++row) {
      const long long r0 = base + (row - 1) * Wp;
      const bool ok = !pad && (row > 0 || yo > 0 || b > 0);  // the line above the first image does not exist: zeros
      load_px<CW>(in + r0 * in_ldw, ok, px[row * 4 + 1]);
      load_px<CW>(in + (r0 + 1) * in_ldw, ok, px[row * 4 + 2]);
      // column -1 of the first window of a line is the (zero) pad column of the previous line
      load_px<CW>(in + (r0 - 1) * in_ldw, ok && lane == 0 && xo != 0, px[row * 4 + 0]);
      load_px<CW>(in + (r0 + 2) * in_ldw, ok && lane == 31, px[row * 4 + 3]);  // (column <= W: in bounds)
    }
    // ... then the exchange
#pragma unroll
    for (int row = 0; row < 4; ++row) {
#pragma unroll
      for (int w = 0; w < CW; ++w) {
        const uint32_t l = __shfl_up_sync(0xffffffffu, px[row * 4 + 2][w], 1);
        const uint32_t r = __shfl_down_sync(0xffffffffu, px[row * 4 + 1][w], 1);
        if (lane != 0) px[row * 4 + 0][w] = xo == 0 ? 0u : l;
        if (lane != 31) px[row * 4 + 3][w] = r;
      }
    }
    if (pad) return;
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int py = pos / 2, qx = pos % 2;
      float acc[NT];
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int pi = (py + t / 3) * 4 + (qx + t % 3);
#pragma unroll
        for (int cw = 0; cw < CW; ++cw) {
          const float a0 = bf_lo(px[pi][cw]), a1 = bf_hi(px[pi][cw]);
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            acc[n] = fmaf(a0, P.w[(t * CT + 2 * cw) * NT + n], acc[n]);
            if (!(ODD && cw == CW - 1)) acc[n] = fmaf(a1, P.w[(t * CT + 2 * cw + 1) * NT + n], acc[n]);
          }
        }
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        float v = fmaf(acc[n], P.sc[n], P.sh[n]);
        if (leaky) v = leaky01(v);
        fin[n] = pos == 0 ? v : fmaxf(fin[n], v);
      }
    }
    out_row = (long long)i;  // the pooled flat index IS the destination row
  }

  __nv_bfloat16* dst = out + out_row * ldc;
  if constexpr (N2T == 0) {
    store_row<NT>(dst, fin);
  } else {
    // the 1x1 layer behind this one, on the activations as they would have been stored (bf16)
    float o2[N2T];
#pragma unroll
    for (int o = 0; o < N2T; ++o) o2[o] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float a = __bfloat162float(__float2bfloat16_rn(fin[n]));
#pragma unroll
      for (int o = 0; o < N2T; ++o) o2[o] = fmaf(a, P.w2[n * N2T + o], o2[o]);
    }
#pragma unroll
    for (int o = 0; o < N2T; ++o) {
      float v = fmaf(o2[o], P.sc2[o], P.sh2[o]);
      if (leaky2) v = leaky01(v);
      o2[o] = v;
    }
    store_row<N2T>(dst, o2);
  }
}

struct ThinGeom {
  int cw, nt, n2t, odd;
};

bool thin_pick(int ks, int Cin, int N, int pool, int N2, ThinGeom* g) {
  static const int cw3[] = {1, 2, 4}, cw1[] = {4, 8, 12, 16};
  static const int nts[] = {1, 2, 4, 8, 12, 16, 20, 24}, ntf[] = {8, 12, 16, 20, 24}, n2s[] = {4, 8};
  if (Cin < 1 || N < 1 || N2 < 0) return false;
  int cw = 0, nt = 0, n2t = 0;
  if (ks == 3) { for (int c : cw3) if (!cw && 2 * c >= Cin) cw = c; }
  else if (ks == 1) { for (int c : cw1) if (!cw && 2 * c >= Cin) cw = c; }
  if (!cw) return false;
  if (N2 > 0) {
    for (int n : ntf) if (!nt && n >= N) nt = n;
    for (int n : n2s) if (!n2t && n >= N2) n2t = n;
    if (!n2t) return false;
  } else {
    for (int n : nts) if (!nt && n >= N) nt = n;
  }
  if (!nt || !thin_valid(ks, cw, nt, pool ? 1 : 0, n2t)) return false;
  g->cw = cw; g->nt = nt; g->n2t = n2t; g->odd = (Cin == 2 * cw - 1) ? 1 : 0;
  return true;
}

struct ThinLaunch {
  ThinParams P;
  const uint32_t* in;
  __nv_bfloat16* out;
  int H, W, in_ldw, ldc, leaky, leaky2;
  unsigned total;
  cudaStream_t stream;
  int ks, cw, nt, pool, n2t, odd;
  bool done;
};

template <int KS, int CW, int NT, int POOL, int N2T>
inline void thin_try(ThinLaunch& L) {
  if constexpr (thin_valid(KS, CW, NT, POOL, N2T) && KS * KS * 2 * CW * NT <= TH_MAXW) {
    if (!L.done && L.ks == KS && L.cw == CW && L.nt == NT && L.pool == POOL && L.n2t == N2T) {
      const unsigned grid = (L.total + TH_THREADS - 1u) / TH_THREADS;
      if (L.odd)
        conv_thin_kernel<KS, CW, NT, POOL != 0, N2T, true><<<grid, TH_THREADS, 0, L.stream>>>(
            L.P, L.in, L.out, L.H, L.W, L.in_ldw, L.ldc, L.leaky, L.leaky2, L.total);
      else
        conv_thin_kernel<KS, CW, NT, POOL != 0, N2T, false><<<grid, TH_THREADS, 0, L.stream>>>(
            L.P, L.in, L.out, L.H, L.W, L.in_ldw, L.ldc, L.leaky, L.leaky2, L.total);
      L.done = true;
    }
  }
}

template <int KS, int CW, int NT>
inline void thin_try_nt(ThinLaunch& L) {
  thin_try<KS, CW, NT, 0, 0>(L);
  thin_try<KS, CW, NT, 1, 0>(L);
  thin_try<KS, CW, NT, 0, 4>(L);
  thin_try<KS, CW, NT, 0, 8>(L);
}

template <int KS, int CW>
inline void thin_try_cw(ThinLaunch& L) {
  thin_try_nt<KS, CW, 1>(L);
  thin_try_nt<KS, CW, 2>(L);
  thin_try_nt<KS, CW, 4>(L);
  thin_try_nt<KS, CW, 8>(L);
  thin_try_nt<KS, CW, 12>(L);
  thin_try_nt<KS, CW, 16>(L);
  thin_try_nt<KS, CW, 20>(L);
  thin_try_nt<KS, CW, 24>(L);
}

}  // namespace

extern "C" int mc_conv_thin_geometry(int ksize, int Cin, int N, int pool, int N2, int* ct, int* nt, int* n2t) {
  ThinGeom g;
  if (!thin_pick(ksize, Cin, N, pool, N2, &g)) return 0;
  if (ct) *ct = 2 * g.cw;
  if (nt) *nt = g.nt;
  if (n2t) *n2t = g.n2t;
  return 1;
}

extern "C" int mc_conv_thin_fwd(const void* d_in, const float* h_w, const float* h_scale, const float* h_shift,
                                const float* h_w2, const float* h_scale2, const float* h_shift2, void* d_out, int B,
                                int H, int W, int Cin, int Cin_ld, int N, int ldc, int ksize, int leaky, int pool, int N2,
                                int leaky2, void* stream_) {
  MC_CHECK_ARG(d_in && h_w && h_scale && h_shift && d_out, "mc_conv_thin_fwd: null pointer");
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0, "mc_conv_thin_fwd: bad dims");
  ThinGeom g;
  MC_CHECK_ARG(thin_pick(ksize, Cin, N, pool, N2, &g),
               "mc_conv_thin_fwd: k=%d Cin=%d N=%d pool=%d N2=%d outside the thin path", ksize, Cin, N, pool, N2);
  MC_CHECK_ARG(N2 == 0 || (h_w2 && h_scale2 && h_shift2), "mc_conv_thin_fwd: fused 1x1 layer without weights");
  MC_CHECK_ARG(Cin_ld % 8 == 0 && Cin_ld >= 2 * g.cw, "mc_conv_thin_fwd: input pitch %d < %d channels read per pixel",
               Cin_ld, 2 * g.cw);
  const int nstore = g.n2t ? g.n2t : g.nt;
  MC_CHECK_ARG(ldc % 8 == 0 && ldc >= (nstore + 7) / 8 * 8, "mc_conv_thin_fwd: output pitch %d too small for %d channels",
               ldc, nstore);
  MC_CHECK_ARG(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_out % 16) == 0, "mc_conv_thin_fwd: buffers must be 16-byte aligned");
  if (pool) MC_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "mc_conv_thin_fwd: pooling needs even H,W");
  const long long total = pool ? (long long)B * (H / 2 + 1) * (W / 2 + 1) : (long long)B * (H + 1) * (W + 1);
  MC_CHECK_ARG((long long)B * (H + 1) * (W + 1) < (1ll << 31), "mc_conv_thin_fwd: tensor too large");
  static thread_local ThinLaunch L;
  const int taps = ksize * ksize, ct = 2 * g.cw;
  for (int i = 0; i < taps * ct * g.nt; ++i) L.P.w[i] = h_w[i];
  for (int i = 0; i < g.nt; ++i) { L.P.sc[i] = h_scale[i]; L.P.sh[i] = h_shift[i]; }
  if (g.n2t) {
    for (int i = 0; i < g.nt * g.n2t; ++i) L.P.w2[i] = h_w2[i];
    for (int i = 0; i < g.n2t; ++i) { L.P.sc2[i] = h_scale2[i]; L.P.sh2[i] = h_shift2[i]; }
  }
  L.in = reinterpret_cast<const uint32_t*>(d_in);
  L.out = reinterpret_cast<__nv_bfloat16*>(d_out);
  L.H = H; L.W = W; L.in_ldw = Cin_ld / 2; L.ldc = ldc; L.leaky = leaky; L.leaky2 = leaky2;
  L.total = (unsigned)total;
  L.stream = reinterpret_cast<cudaStream_t>(stream_);
  L.ks = ksize; L.cw = g.cw; L.nt = g.nt; L.pool = pool ? 1 : 0; L.n2t = g.n2t; L.odd = g.odd;
  L.done = false;
  thin_try_cw<3, 1>(L);
  thin_try_cw<3, 2>(L);
  thin_try_cw<3, 4>(L);
  thin_try_cw<1, 4>(L);
  thin_try_cw<1, 8>(L);
  thin_try_cw<1, 12>(L);
  thin_try_cw<1, 16>(L);
  MC_CHECK_ARG(L.done, "mc_conv_thin_fwd: internal: no kernel instance for k=%d cw=%d nt=%d pool=%d n2t=%d", ksize, g.cw,
               g.nt, L.pool, g.n2t);
  MC_LAUNCH_CHECK("conv_thin_kernel");
  return 0;
}
