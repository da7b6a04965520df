// conv_im2col_tc.cu — 3x3 conv (+ folded BN + leaky [+ 2x2/2 max-pool]) for THIN inputs on tcgen05, with the im2col
// A-tile built in shared memory by the CTA's own threads.
//
// Replaces the same reference call sites as conv_tcgen05.cu (MaskedConv2d.forward / BatchNorm2d / LeakyReLU /
// MaxPool2d, src/pruning/weightPruning/layers.py:53-64, src/nets.py:802,809,821) for layers whose input has <= 16
// channels: the 3-channel first layer (fp32 NCHW image) and the first blocks of a filter-pruned network.  The TMA
// formulation of conv_tcgen05.cu spends one 64-wide k-block per tap, i.e. 9 x 64 columns of K for 3..16 real
// channels; here a GEMM row holds the whole receptive field contiguously (K = 9*CL or 16*CL), so one or two
// k-blocks suffice and the op is bound by reading the input once.
//
// Pool fusion ("pool-window GEMM"): a GEMM row is one 2x2 pool window; its K axis is the 4x4xCL input patch the four
// conv outputs of the window depend on, and the N axis is (position-in-window, out-channel) = 4*N columns whose
// weights are the 3x3 filter shifted to each position (zeros elsewhere; built on the host).  The pool is then a
// max over 4 TMEM column groups inside one thread, and each thread stores one pooled pixel.
//
// CTA = 128 threads = 128 GEMM rows (8 x 16 windows/pixels); thread t builds row t, owns TMEM lane t.
#include <cuda.h>
#include "common.cuh"
#include "ptx_sm100.cuh"

int mc_make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows);

namespace {

constexpr int TY = 8, TX = 16;  // tile of GEMM rows

struct Im2colParams {
  const void* in;
  void* out;
  const float* scale;
  const float* shift;
  int B, H, W;          // conv resolution
  int Cin, Cin_ld;      // valid input channels, input pitch (PNHWC) — FIRST: Cin planes of the NCHW image
  int N, npos, nb;      // valid outputs; per-position column stride; UMMA N (pool: 4*npos, else npos)
  int ldc, leaky;
  int nkb, ksteps;      // k-blocks of 64, total UMMA K steps (16 elements each)
  int tiles_x, tiles_y, total_tiles;
  int tmem_cols;
  uint32_t idesc;
};

__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  ptx::tmem_ld_32x32b_x16(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// store `cnt` (<=8) bf16 values starting at channel n0 of dst row
__device__ __forceinline__ void store_group(__nv_bfloat16* dst, int n0, const float* v, int N, bool vec) {
  if (vec && n0 + 8 <= N) {
    __align__(16) __nv_bfloat16 h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2bfloat16_rn(v[j]);
    *reinterpret_cast<uint4*>(dst + n0) = *reinterpret_cast<const uint4*>(h);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < N) dst[n0 + j] = __float2bfloat16_rn(v[j]);
  }
}

template <int CL, bool POOL, bool FIRST>
__global__ void __launch_bounds__(128, 4)
conv_im2col_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const Im2colParams p) {
  constexpr int PR = POOL ? 2 * TY + 2 : TY + 2;   // patch rows / cols (conv pixels incl. halo)
  constexpr int PC = POOL ? 2 * TX + 2 : TX + 2;
  constexpr int NPIX = POOL ? 16 : 9;              // patch pixels per GEMM row
  constexpr int KELEMS = NPIX * CL;
  constexpr int NKB = (KELEMS + 63) / 64;
  constexpr int A_BYTES = NKB * 128 * 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  smem += (1024u - (ptx::smem_u32(smem) & 1023u)) & 1023u;
  uint8_t* a_tile = smem;                                   // [NKB][128 rows][128 B], 128B swizzle
  uint8_t* b_tile = a_tile + A_BYTES;                       // [NKB][nb_pad rows][128 B]
  const int nb_pad = (p.nb + 15) & ~15;
  __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(b_tile + (size_t)NKB * nb_pad * 128);  // [PR][PC][CL]
  uint64_t* b_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + ((PR * PC * CL * 2 + 15) & ~15));
  uint64_t* mma_bar = b_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);

  const int t = threadIdx.x;
  const int warp_idx = t >> 5;
  if (t == 0) {
    ptx::prefetch_tensormap(&tmap_b);
    ptx::mbar_init(b_bar, 1);
    ptx::mbar_init(mma_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 0) {
    ptx::tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (t == 0) {  // expanded weights: loaded once per CTA
    ptx::mbar_arrive_expect_tx(b_bar, (uint32_t)(NKB * nb_pad * 128));
    for (int kb = 0; kb < NKB; ++kb) ptx::tma_load_2d(b_tile + (size_t)kb * nb_pad * 128, &tmap_b, b_bar, kb * 64, 0);
  }

  const int wy = t / TX, wx = t % TX;
  const int Hout = POOL ? p.H / 2 : p.H, Wout = POOL ? p.W / 2 : p.W;
  const bool vec_store = (p.ldc & 7) == 0;
  uint32_t mma_phase = 0;
  bool b_ready = false;

  // Patch staging goes through registers so that the global loads of tile i+1 are in flight while tile i is built,
  // multiplied and stored (one DRAM round trip per tile would otherwise sit on the critical path).
  constexpr int PV = FIRST ? 1 : CL / 8;                       // 16-byte vectors per patch pixel (FIRST: one pixel)
  constexpr int NLOAD = (PR * PC * PV + 127) / 128;
  float pf[FIRST ? NLOAD * 4 : 1];
  uint4 pq[FIRST ? 1 : NLOAD];
  auto load_patch = [&](int tile) {
    const int tx = tile % p.tiles_x;
    const int ty = (tile / p.tiles_x) % p.tiles_y;
    const int b = tile / (p.tiles_x * p.tiles_y);
    const int iy0 = (POOL ? 2 * ty * TY : ty * TY) - 1, ix0 = (POOL ? 2 * tx * TX : tx * TX) - 1;
#pragma unroll
    for (int u = 0; u < NLOAD; ++u) {
      const int i = t + u * 128;
      const int pix = i / PV, v = i - pix * PV;
      const int r = pix / PC, sx = pix - r * PC;
      const int yy = iy0 + r, xx = ix0 + sx;
      const bool ok = (i < PR * PC * PV) && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      if constexpr (FIRST) {
        const float* img = reinterpret_cast<const float*>(p.in);
        const long long o = ((long long)b * p.Cin * p.H + yy) * p.W + xx;
        const long long plane = (long long)p.H * p.W;
        pf[u * 4 + 0] = ok ? img[o] : 0.f;
        pf[u * 4 + 1] = (ok && p.Cin > 1) ? img[o + plane] : 0.f;
        pf[u * 4 + 2] = (ok && p.Cin > 2) ? img[o + 2 * plane] : 0.f;
        pf[u * 4 + 3] = (ok && p.Cin > 3) ? img[o + 3 * plane] : 0.f;
      } else {
        const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.in);
        pq[u] = ok ? *reinterpret_cast<const uint4*>(
                         x + (((long long)b * (p.H + 1) + yy) * (p.W + 1) + xx) * p.Cin_ld + v * 8)
                   : make_uint4(0, 0, 0, 0);
      }
    }
  };
  auto store_patch = [&]() {
#pragma unroll
    for (int u = 0; u < NLOAD; ++u) {
      const int i = t + u * 128;
      if (i < PR * PC * PV) {
        if constexpr (FIRST) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(pf[u * 4 + 0], pf[u * 4 + 1]);
          __nv_bfloat162 hi = __floats2bfloat162_rn(pf[u * 4 + 2], pf[u * 4 + 3]);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(patch + (size_t)i * CL) = pk;  // CL == 4 for FIRST
        } else {
          *reinterpret_cast<uint4*>(patch + (size_t)i * 8) = pq[u];  // pixel-major, PV vectors per pixel
        }
      }
    }
  };

  if ((int)blockIdx.x < p.total_tiles) load_patch(blockIdx.x);
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int tx = tile % p.tiles_x;
    const int ty = (tile / p.tiles_x) % p.tiles_y;
    const int b = tile / (p.tiles_x * p.tiles_y);
    const int oy0 = ty * TY, ox0 = tx * TX;  // first output row/col of the tile

    // ---- 1. patch registers -> shared memory (bf16 [PR][PC][CL]); then prefetch the next tile's patch
    store_patch();
    __syncthreads();
    if (tile + (int)gridDim.x < p.total_tiles) load_patch(tile + gridDim.x);

    // ---- 2. build GEMM row t: NPIX patch pixels x CL channels, pixel-major, zero padded to NKB*64 elements
    {
      const int by = POOL ? 2 * wy : wy, bx = POOL ? 2 * wx : wx;
#pragma unroll
      for (int q = 0; q < NKB * 8; ++q) {  // 16-byte chunk q holds K elements [8q, 8q+8)
        uint4 val = make_uint4(0, 0, 0, 0);
        if (q * 8 < KELEMS) {
          if constexpr (CL == 4) {
            const int pa = 2 * q, pb = 2 * q + 1;  // two pixels per chunk
            uint2 lo = make_uint2(0, 0), hi = make_uint2(0, 0);
            {
              const int r = POOL ? pa / 4 : pa / 3, s = POOL ? pa % 4 : pa % 3;
              lo = *reinterpret_cast<const uint2*>(patch + ((by + r) * PC + bx + s) * CL);
            }
            if (pb < NPIX) {
              const int r = POOL ? pb / 4 : pb / 3, s = POOL ? pb % 4 : pb % 3;
              hi = *reinterpret_cast<const uint2*>(patch + ((by + r) * PC + bx + s) * CL);
            }
            val = make_uint4(lo.x, lo.y, hi.x, hi.y);
          } else {
            constexpr int V = CL / 8;
            const int pix = q / V, v = q % V;
            const int r = POOL ? pix / 4 : pix / 3, s = POOL ? pix % 4 : pix % 3;
            val = *reinterpret_cast<const uint4*>(patch + ((by + r) * PC + bx + s) * CL + v * 8);
          }
        }
        const int kb = q >> 3, j = q & 7;
        *reinterpret_cast<uint4*>(a_tile + (size_t)kb * (128 * 128) + t * 128 + ((j ^ (t & 7)) << 4)) = val;
      }
    }
    ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    ptx::tc_fence_before();
    __syncthreads();

    // ---- 3. MMA (one thread)
    if (t == 0) {
      if (!b_ready) ptx::mbar_wait(b_bar, 0);
      ptx::tc_fence_after();
      const uint32_t a_addr = ptx::smem_u32(a_tile), b_addr = ptx::smem_u32(b_tile);
      for (int ks = 0; ks < p.ksteps; ++ks) {
        const int kb = ks >> 2, k = ks & 3;
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_addr + kb * (128 * 128)) + (uint64_t)(2 * k);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(b_addr + kb * (nb_pad * 128)) + (uint64_t)(2 * k);
        ptx::umma_bf16_ss(tmem_base, adesc, bdesc, p.idesc, ks > 0 ? 1u : 0u);
      }
      ptx::umma_commit(mma_bar);
    }
    b_ready = true;

    // ---- 4. epilogue: TMEM lane t -> scale/shift (-> max over the 4 window positions) -> leaky -> bf16 store
    ptx::mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1u;
    ptx::tc_fence_after();
    const int oy = oy0 + wy, ox = ox0 + wx;
    const bool valid = oy < Hout && ox < Wout;
    __nv_bfloat16* dst =
        reinterpret_cast<__nv_bfloat16*>(p.out) + (((long long)b * (Hout + 1) + oy) * (Wout + 1) + ox) * p.ldc;
    const uint32_t trow = tmem_base + ((uint32_t)(warp_idx * 32) << 16);
    if (POOL) {
      const int np = p.npos;
      if (np % 16 == 0) {
        for (int n0 = 0; n0 < np; n0 += 16) {
          float m[16], v[16];
          tmem_ld_x16(trow + n0, m);
#pragma unroll
          for (int j = 0; j < 16; ++j) m[j] = m[j] * __ldg(p.scale + n0 + j) + __ldg(p.shift + n0 + j);
          for (int pos = 1; pos < 4; ++pos) {
            tmem_ld_x16(trow + pos * np + n0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) m[j] = fmaxf(m[j], v[j] * __ldg(p.scale + n0 + j) + __ldg(p.shift + n0 + j));
          }
          if (p.leaky) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m[j] = leaky01(m[j]);
          }
          if (valid) {
            store_group(dst, n0, m, p.N, vec_store);
            store_group(dst, n0 + 8, m + 8, p.N, vec_store);
          }
        }
      } else {  // np == 4 or 8: the four positions sit in one or two 16-column loads
        float v[32];
        {
          float a[16];
          tmem_ld_x16(trow, a);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = a[j];
          if (np == 8) {
            tmem_ld_x16(trow + 16, a);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[16 + j] = a[j];
          }
        }
        float m[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          m[n] = 0.f;
          if (n < np) {
            const float sc = __ldg(p.scale + n), sh = __ldg(p.shift + n);
            float r = v[n] * sc + sh;
#pragma unroll
            for (int pos = 1; pos < 4; ++pos) r = fmaxf(r, v[pos * np + n] * sc + sh);
            m[n] = p.leaky ? leaky01(r) : r;
          }
        }
        if (valid) store_group(dst, 0, m, p.N, vec_store);
      }
    } else {
      for (int n0 = 0; n0 < p.nb; n0 += 16) {
        float v[16];
        tmem_ld_x16(trow + n0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = v[j] * __ldg(p.scale + n0 + j) + __ldg(p.shift + n0 + j);
          v[j] = p.leaky ? leaky01(a) : a;
        }
        if (valid) {
          store_group(dst, n0, v, p.N, vec_store);
          store_group(dst, n0 + 8, v + 8, p.N, vec_store);
        }
      }
    }
    ptx::tc_fence_before();  // TMEM reads done before the next tile's MMA (ordered by the next __syncthreads)
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

template <int CL, bool POOL, bool FIRST>
int launch_im2col(const CUtensorMap& tm_b, const Im2colParams& p, cudaStream_t stream) {
  constexpr int PR = POOL ? 2 * TY + 2 : TY + 2, PC = POOL ? 2 * TX + 2 : TX + 2;
  constexpr int NPIX = POOL ? 16 : 9;
  constexpr int NKB = (NPIX * CL + 63) / 64;
  const int nb_pad = (p.nb + 15) & ~15;
  const size_t smem = (size_t)NKB * 128 * 128 + (size_t)NKB * nb_pad * 128 + ((PR * PC * CL * 2 + 15) & ~15) + 64 + 1024;
  if (smem > 227 * 1024) return mc_set_error(MC_ERR_SHAPE, "mc_conv_im2col_fwd: %zu B of shared memory", smem);
  auto kern = conv_im2col_tc_kernel<CL, POOL, FIRST>;
  static size_t attr_smem = 0;  // per template instantiation
  if (smem > attr_smem) {
    MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int per_sm = (int)((200 * 1024) / smem);
  const int tmem_lim = 512 / p.tmem_cols;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)mc_num_sms() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<(int)grid, 128, smem, stream>>>(tm_b, p);
  MC_LAUNCH_CHECK("conv_im2col_tc_kernel");
  return 0;
}

}  // namespace

// K elements per GEMM row (before padding to a multiple of 64) and UMMA-N of the expanded weight matrix.
extern "C" int mc_conv_im2col_supported(int Cin, int in_is_nchw_f32, int N, int pool) {
  if (Cin < 1 || N < 1) return 0;
  if (in_is_nchw_f32) {
    if (Cin > 4) return 0;
  } else if (Cin > 16) {
    return 0;
  }
  const int npos = pool ? ((N <= 4) ? 4 : (N <= 8) ? 8 : ((N + 15) & ~15)) : ((N + 15) & ~15);
  const int nb = pool ? 4 * npos : npos;
  return nb <= 256 ? 1 : 0;
}

extern "C" int mc_conv_im2col_geometry(int Cin, int in_is_nchw_f32, int N, int pool, int* cl, int* npos, int* nb,
                                       int* kpad) {
  if (!mc_conv_im2col_supported(Cin, in_is_nchw_f32, N, pool)) return mc_set_error(MC_ERR_SHAPE, "mc_conv_im2col: unsupported shape");
  const int CL = in_is_nchw_f32 ? 4 : (Cin <= 8 ? 8 : 16);
  const int np = pool ? ((N <= 4) ? 4 : (N <= 8) ? 8 : ((N + 15) & ~15)) : ((N + 15) & ~15);
  if (cl) *cl = CL;
  if (npos) *npos = np;
  if (nb) *nb = pool ? 4 * np : np;
  if (kpad) *kpad = (((pool ? 16 : 9) * CL + 63) / 64) * 64;
  return 0;
}

extern "C" int mc_conv_im2col_fwd(const void* d_in, int in_is_nchw_f32, const void* d_wexp, const float* d_scale,
                                  const float* d_shift, void* d_out, int B, int H, int W, int Cin, int Cin_ld, int N,
                                  int ldc, int leaky, int pool, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_wexp && d_scale && d_shift && d_out, "mc_conv_im2col_fwd: null pointer");
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0, "mc_conv_im2col_fwd: bad dims");
  int CL, npos, nb, kpad;
  int rc = mc_conv_im2col_geometry(Cin, in_is_nchw_f32, N, pool, &CL, &npos, &nb, &kpad);
  if (rc) return rc;
  if (!in_is_nchw_f32) MC_CHECK_ARG(Cin_ld == CL, "mc_conv_im2col_fwd: input pitch %d must equal %d", Cin_ld, CL);
  MC_CHECK_ARG(ldc >= N, "mc_conv_im2col_fwd: ldc < N");
  if (pool) MC_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "mc_conv_im2col_fwd: pooling needs even H,W");
  MC_CHECK_ARG(((uintptr_t)d_wexp & 15) == 0 && ((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0,
               "mc_conv_im2col_fwd: pointers must be 16-byte aligned");

  Im2colParams p;
  p.in = d_in;
  p.out = d_out;
  p.scale = d_scale;
  p.shift = d_shift;
  p.B = B; p.H = H; p.W = W;
  p.Cin = Cin; p.Cin_ld = Cin_ld;
  p.N = N; p.npos = npos; p.nb = nb;
  p.ldc = ldc; p.leaky = leaky;
  p.nkb = kpad / 64;
  const int kelems = (pool ? 16 : 9) * CL;
  p.ksteps = (kelems + 15) / 16;
  const int Hout = pool ? H / 2 : H, Wout = pool ? W / 2 : W;
  p.tiles_x = (Wout + TX - 1) / TX;
  p.tiles_y = (Hout + TY - 1) / TY;
  p.total_tiles = B * p.tiles_x * p.tiles_y;
  int tc = 32;
  while (tc < nb) tc <<= 1;
  p.tmem_cols = tc;
  const int nb_pad = (nb + 15) & ~15;
  p.nb = nb_pad;  // UMMA N must be a multiple of 16; padded columns are zero rows of d_wexp
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nb_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  CUtensorMap tm_b;
  rc = mc_make_tmap_2d_bf16(&tm_b, d_wexp, (uint64_t)nb_pad, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)nb_pad);
  if (rc) return rc;

  if (in_is_nchw_f32) return pool ? launch_im2col<4, true, true>(tm_b, p, stream) : launch_im2col<4, false, true>(tm_b, p, stream);
  if (CL == 8) return pool ? launch_im2col<8, true, false>(tm_b, p, stream) : launch_im2col<8, false, false>(tm_b, p, stream);
  return pool ? launch_im2col<16, true, false>(tm_b, p, stream) : launch_im2col<16, false, false>(tm_b, p, stream);
}
