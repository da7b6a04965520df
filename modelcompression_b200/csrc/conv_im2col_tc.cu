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

#include "tmap.cuh"

namespace {

__device__ unsigned int g_im2col_dbg = 0;  // first mbarrier wait that timed out (0 = none)

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
  int nsc;              // entries of scale/shift that may be read
  int nkb, ksteps;      // k-blocks of 64, total UMMA K steps (16 elements each)
  int tiles_x, tiles_y, total_tiles;
  int tmem_cols;
  uint32_t idesc;
};

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  ptx::tmem_ld_32x32b_x16(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// store 8 bf16 values at channels [n0, n0+8) of a destination row.  Channels >= N are the row's zero padding: when the
// 8-group fits inside the pitch it is written as ONE 16-byte store with zeros there (whole-sector, coalesced writes);
// scalar stores only when the pitch is not a multiple of 8.
__device__ __forceinline__ void store_group(__nv_bfloat16* dst, int n0, const float* v, int N, int ldc, bool vec) {
  if (n0 >= N) return;
  if (vec && n0 + 8 <= ldc) {
    __align__(16) __nv_bfloat16 h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2bfloat16_rn(n0 + j < N ? v[j] : 0.f);
    *reinterpret_cast<uint4*>(dst + n0) = *reinterpret_cast<const uint4*>(h);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < N) dst[n0 + j] = __float2bfloat16_rn(v[j]);
  }
}

// ---- epilogue helpers -------------------------------------------------------------------------------------------
// pooled, NP (per-position column stride) = 4 or 8: all four window positions sit in one or two 16-column loads
template <int NP>
__device__ __forceinline__ void epilogue_pool_small(uint32_t trow, const float* s_sc, const float* s_sh, int leaky,
                                                    bool valid, __nv_bfloat16* dst, int N, int ldc, bool vec_store) {
  float v[4 * NP];
  {
    float a[16];
    tmem_ld_x16(trow, a);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = a[j];
    if (NP == 8) {
      tmem_ld_x16(trow + 16, a);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[(16 + j) % (4 * NP)] = a[j];
    }
  }
  float m[8];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    m[n] = 0.f;
    if (n < NP) {
      const float sc = s_sc[n], sh = s_sh[n];
      float r = fmaf(v[n], sc, sh);
#pragma unroll
      for (int pos = 1; pos < 4; ++pos) r = fmaxf(r, fmaf(v[pos * NP + n], sc, sh));
      m[n] = leaky ? fmaxf(r, 0.1f * r) : r;
    }
  }
  if (valid) store_group(dst, 0, m, N, ldc, vec_store);
}

// CTA = 256 threads.  Warps 0-3 ("builders"): thread t builds GEMM row t of the im2col tile from the TMA-staged input
// patch, thread 0 additionally issues the patch TMA and the MMAs.  Warps 4-7: epilogue (TMEM lane quarter = warp%4).
// Everything is double-buffered on the tile parity (patch, A tile, TMEM accumulator), so the build of tile i+1
// overlaps the MMA + epilogue of tile i:
//   patch_full[2]  TMA -> builders          a_free[2]    tcgen05.commit -> builders (MMA has consumed A[s])
//   tmem_full[2]   tcgen05.commit -> epilogue   tmem_free[2] epilogue warps (4 arrivals) -> MMA issuer
template <int CL, bool POOL, int FIRST>  // FIRST: 0 = PNHWC bf16 input, 1 = fp32 NCHW image, 2 = uint8 NCHW image (x/255)
__global__ void __launch_bounds__(256, 4)
conv_im2col_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_in,
                      const Im2colParams p) {
  constexpr int PR = POOL ? 2 * TY + 2 : TY + 2;   // patch rows / cols (conv pixels incl. halo)
  constexpr int PC = POOL ? 2 * TX + 2 : TX + 2;
  constexpr int NPIX = POOL ? 16 : 9;              // patch pixels per GEMM row
  constexpr int KELEMS = NPIX * CL;
  constexpr int NKB = (KELEMS + 63) / 64;
  constexpr int A_BYTES = NKB * 128 * 128;
  // FIRST: fp32 patch [Cin][PR][PCF]; TMA needs a 16-byte aligned innermost start, so the box begins XS = 3 columns
  // left of the halo column (x0-4 instead of x0-1) and PCF = round_up(PC + 3, 4).  Else bf16 patch [PR][PC][CL].
  // uint8 image: same with 16-byte = 16-pixel granularity (XS = 15, PCF = round_up(PC + 15, 16)).
  constexpr int XS = FIRST == 2 ? 15 : 3;
  constexpr int PCF = FIRST == 2 ? ((PC + XS + 15) & ~15) : ((PC + XS + 3) & ~3);
  constexpr int PEL = FIRST == 2 ? 1 : 4;  // bytes per patch element
  // patch element -> float: the uint8 image is scaled like ToTensor / do_detect (img.float().div(255.0),
  // src/nets2_utils.py:346-352) before the bf16 rounding every GEMM operand gets
  // (a 256-entry table in shared memory: one load instead of an IEEE division per element, same values)
  __shared__ float s_u8lut[FIRST == 2 ? 256 : 1];
  if constexpr (FIRST == 2) s_u8lut[threadIdx.x & 255] = __fdiv_rn((float)(threadIdx.x & 255), 255.0f);
  auto pel = [&](const uint8_t* patch, int idx) -> float {
    if constexpr (FIRST == 2) return s_u8lut[patch[idx]];
    else return reinterpret_cast<const float*>(patch)[idx];
  };

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  smem += (1024u - (ptx::smem_u32(smem) & 1023u)) & 1023u;
  uint8_t* a_tile = smem;                                   // [2][NKB][128 rows][128 B], 128B swizzle
  uint8_t* b_tile = a_tile + 2 * A_BYTES;                   // [NKB][nb_pad rows][128 B]
  const int nb_pad = (p.nb + 15) & ~15;
  const uint32_t patch_bytes = FIRST ? (uint32_t)(p.Cin * PR * PCF * PEL) : (uint32_t)(PR * PC * CL * 2);
  const uint32_t patch_stride = (patch_bytes + 127u) & ~127u;
  uint8_t* patch0 = b_tile + (size_t)NKB * nb_pad * 128;  // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(patch0 + 2 * patch_stride);
  uint64_t* b_bar = bars;            // weights landed
  uint64_t* patch_full = bars + 1;   // [2]
  uint64_t* a_free = bars + 3;       // [2]
  uint64_t* tmem_full = bars + 5;    // [2]
  uint64_t* tmem_free = bars + 7;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);
  float* s_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);  // [256] scale, [256] shift
  float* s_sh = s_sc + 256;

  const int t = threadIdx.x;
  const int warp_idx = t >> 5;
  if (t == 0) {
    ptx::prefetch_tensormap(&tmap_b);
    ptx::prefetch_tensormap(&tmap_in);
    ptx::mbar_init(b_bar, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&patch_full[i], 1);
      ptx::mbar_init(&a_free[i], 1);
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_free[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 0) {
    ptx::tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  for (int i = t; i < 256; i += 256) {  // per-channel scale/shift (arrays are padded to >= 16 entries by the host)
    const bool ok = i < p.nsc;
    s_sc[i] = ok ? __ldg(p.scale + i) : 0.f;
    s_sh[i] = ok ? __ldg(p.shift + i) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t acc_stride = (uint32_t)((nb_pad + 31) & ~31);  // TMEM columns between the two accumulator stages
  const unsigned int tiles_xy = (unsigned)(p.tiles_x * p.tiles_y);
  const int Hout = POOL ? p.H / 2 : p.H, Wout = POOL ? p.W / 2 : p.W;

  if (warp_idx < 4) {
    // ============================== builders (+ TMA / MMA issue by thread 0) ==============================
    auto issue_patch = [&](unsigned int tile, int buf) {
      const unsigned int b = tile / tiles_xy, rem = tile - b * tiles_xy;
      const unsigned int ty = rem / (unsigned)p.tiles_x, tx = rem - ty * (unsigned)p.tiles_x;
      const int iy0 = (int)(POOL ? 2 * ty * TY : ty * TY) - 1, ix0 = (int)(POOL ? 2 * tx * TX : tx * TX) - 1;
      ptx::mbar_arrive_expect_tx(&patch_full[buf], patch_bytes);
      if constexpr (FIRST != 0)
        ptx::tma_load_4d(patch0 + buf * patch_stride, &tmap_in, &patch_full[buf], ix0 - XS, iy0, 0, (int)b);
      else
        ptx::tma_load_3d(patch0 + buf * patch_stride, &tmap_in, &patch_full[buf], 0, ix0, (int)b * (p.H + 1) + iy0);
    };
    if (t == 0) {
      ptx::mbar_arrive_expect_tx(b_bar, (uint32_t)(NKB * nb_pad * 128));  // expanded weights: once per CTA
      for (int kb = 0; kb < NKB; ++kb)
        ptx::tma_load_2d(b_tile + (size_t)kb * nb_pad * 128, &tmap_b, b_bar, kb * 64, 0);
      unsigned int tl = blockIdx.x;
      for (int i = 0; i < 2 && tl < (unsigned)p.total_tiles; ++i, tl += gridDim.x) issue_patch(tl, i);
    }
    const int wy = t / TX, wx = t % TX;
    const int by = POOL ? 2 * wy : wy, bx = POOL ? 2 * wx : wx;
    const uint32_t swz = (uint32_t)(t & 7) << 4;
    bool b_ready = false;
    int it = 0;
    for (unsigned int tile = blockIdx.x; tile < (unsigned)p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      ptx::mbar_wait(&a_free[s], ph ^ 1u, &g_im2col_dbg, 0x400u + (unsigned)it);   // MMA of tile it-2 has read A[s]
      ptx::mbar_wait(&patch_full[s], ph, &g_im2col_dbg, 0x100u + (unsigned)it);
      const uint8_t* patch_raw = patch0 + s * patch_stride;
      uint8_t* arow = a_tile + (size_t)s * A_BYTES + t * 128;
      if (FIRST && POOL && p.Cin == 3) {
        // RGB fast path: compile-time offsets from one base pointer.  Row r of the 4x4 window gives chunks 2r
        // (pixels px 0,1) and 2r+1 (pixels px 2,3), each pixel = (c0,c1,c2,0) in bf16.
        // (tried: the second half-warp walking each column pair in the opposite order to avoid the 2-way bank conflict
        //  between the two window rows of a warp — the selects and dynamic offsets cost more than the wavefronts saved:
        //  84 -> 92 us)
        const int base = by * PCF + bx + XS;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float v[3][4];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if constexpr (FIRST == 2) {
              // uint8 image: the window row's 4 bytes come from two aligned 32-bit loads + a funnel shift (two lanes
              // share a word: broadcast) instead of 4 byte loads + 4 table look-ups — the LSU data pipe is this
              // kernel's limiter (ncu), and the byte path made the uint8 kernel slower than the fp32 one.  x/255 is
              // computed exactly as IEEE division: q0 = a*r, q = fma(fma(-255, q0, a), r, q0) with r = RN(1/255)
              // (equal to a/255.0f for all 256 byte values — checked exhaustively).
              const int idx0 = base + (c * PR + r) * PCF;
              const uint32_t* words = reinterpret_cast<const uint32_t*>(patch_raw);
              const uint32_t lo = words[idx0 >> 2], hi = words[(idx0 >> 2) + 1];
              const uint32_t four = __funnelshift_r(lo, hi, (uint32_t)(idx0 & 3) * 8u);
#pragma unroll
              for (int sx = 0; sx < 4; ++sx) {
                const float a = __uint2float_rn((four >> (8 * sx)) & 0xffu);
                const float rcp = 0.00392156886f;  // RN(1/255) = 0x3b808081
                const float q0 = __fmul_rn(a, rcp);
                v[c][sx] = __fmaf_rn(__fmaf_rn(-255.0f, q0, a), rcp, q0);
              }
            } else {
#pragma unroll
              for (int sx = 0; sx < 4; ++sx) v[c][sx] = pel(patch_raw, base + (c * PR + r) * PCF + sx);
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __nv_bfloat162 a0 = __floats2bfloat162_rn(v[0][2 * h], v[1][2 * h]);
            __nv_bfloat162 a1 = __floats2bfloat162_rn(v[2][2 * h], 0.f);
            __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0][2 * h + 1], v[1][2 * h + 1]);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(v[2][2 * h + 1], 0.f);
            uint4 val;
            val.x = *reinterpret_cast<uint32_t*>(&a0);
            val.y = *reinterpret_cast<uint32_t*>(&a1);
            val.z = *reinterpret_cast<uint32_t*>(&b0);
            val.w = *reinterpret_cast<uint32_t*>(&b1);
            *reinterpret_cast<uint4*>(arow + ((uint32_t)((2 * r + h) << 4) ^ swz)) = val;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NKB * 8; ++q) {  // 16-byte chunk q holds K elements [8q, 8q+8)
          uint4 val = make_uint4(0, 0, 0, 0);
          int qs = q;  // chunk position in the A row
          if (q * 8 < KELEMS) {
            if constexpr (FIRST != 0) {
              // two pixels per chunk, 4 bf16 each (c0,c1,c2,c3; absent channels 0); patch is [c][PR][PCF]
              const int pa = 2 * q, pb = 2 * q + 1;
              float va[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < p.Cin) {
                  va[c] = pel(patch_raw, (c * PR + by + (POOL ? pa / 4 : pa / 3)) * PCF + bx + (POOL ? pa % 4 : pa % 3) + XS);
                  if (pb < NPIX)
                    vb[c] = pel(patch_raw, (c * PR + by + (POOL ? pb / 4 : pb / 3)) * PCF + bx + (POOL ? pb % 4 : pb % 3) + XS);
                }
              __nv_bfloat162 a0 = __floats2bfloat162_rn(va[0], va[1]), a1 = __floats2bfloat162_rn(va[2], va[3]);
              __nv_bfloat162 b0 = __floats2bfloat162_rn(vb[0], vb[1]), b1 = __floats2bfloat162_rn(vb[2], vb[3]);
              val.x = *reinterpret_cast<uint32_t*>(&a0);
              val.y = *reinterpret_cast<uint32_t*>(&a1);
              val.z = *reinterpret_cast<uint32_t*>(&b0);
              val.w = *reinterpret_cast<uint32_t*>(&b1);
            } else {
              // Bank conflicts: neighbouring lanes' windows start S = (POOL ? 2 : 1) * CL/8 chunks (16 B) apart, so with
              // every lane reading "its chunk number q" only 8/S of the 8 bank groups are used (2-4x the wavefronts).
              // Lane-dependent XOR on the low bits of the chunk offset inside the window row (and on the matching
              // position in the A row) makes 8 consecutive lanes cover all 8 groups; loads and stores stay a bijection.
              const __nv_bfloat16* patch = reinterpret_cast<const __nv_bfloat16*>(patch_raw);
              constexpr int V = CL / 8;
              constexpr int NPX = POOL ? 4 : 3;
              constexpr int S = (POOL ? 2 : 1) * V;
              const int m = S == 1 ? 0 : (S == 2 ? ((wx >> 2) & 1) : ((wx >> 1) & 3));
              const int pix = q / V, v = q % V;
              const int r = pix / NPX, sx = pix % NPX;
              const int rel = (sx * V + v) ^ m;  // chunk inside the window row, permuted per lane
              val = *reinterpret_cast<const uint4*>(patch + ((by + r) * PC + bx) * CL + rel * 8);
              qs = r * (NPX * V) + rel;
            }
          }
          const int kb = qs >> 3, j = qs & 7;
          *reinterpret_cast<uint4*>(arow + (size_t)kb * (128 * 128) + ((uint32_t)(j << 4) ^ swz)) = val;
        }
      }
      ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      asm volatile("bar.sync 1, 128;" ::: "memory");  // all 128 rows of A[s] written; patch[s] fully consumed

      if (t == 0) {
        // patch[s] is free again: fetch the patch of tile it+2
        const unsigned int nxt = tile + 2u * gridDim.x;
        if (nxt < (unsigned)p.total_tiles) issue_patch(nxt, s);
        if (!b_ready) ptx::mbar_wait(b_bar, 0, &g_im2col_dbg, 0x200u);
        b_ready = true;
        ptx::mbar_wait(&tmem_free[s], ph ^ 1u, &g_im2col_dbg, 0x500u + (unsigned)it);  // epilogue drained TMEM[s]
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(a_tile + (size_t)s * A_BYTES), b_addr = ptx::smem_u32(b_tile);
        const uint32_t tacc = tmem_base + (uint32_t)s * acc_stride;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const int kb = ks >> 2, k = ks & 3;
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_addr + kb * (128 * 128)) + (uint64_t)(2 * k);
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(b_addr + kb * (nb_pad * 128)) + (uint64_t)(2 * k);
          ptx::umma_bf16_ss(tacc, adesc, bdesc, p.idesc, ks > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&a_free[s]);     // A[s] may be rebuilt once these MMAs retire
        ptx::umma_commit(&tmem_full[s]);  // accumulator ready for the epilogue warps
      }
    }
  } else {
    // ============================== epilogue warps 4..7 ==============================
    const int et = t - 128;                 // TMEM lane == GEMM row
    const int quarter = warp_idx & 3;
    const int lane = t & 31;
    const int wy = et / TX, wx = et % TX;
    const bool vec_store = (p.ldc & 7) == 0;
    int it = 0;
    for (unsigned int tile = blockIdx.x; tile < (unsigned)p.total_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      const unsigned int b = tile / tiles_xy, rem = tile - b * tiles_xy;
      const unsigned int ty = rem / (unsigned)p.tiles_x, tx = rem - ty * (unsigned)p.tiles_x;
      const int oy = (int)ty * TY + wy, ox = (int)tx * TX + wx;
      const bool valid = oy < Hout && ox < Wout;
      __nv_bfloat16* dst =
          reinterpret_cast<__nv_bfloat16*>(p.out) + (((long long)b * (Hout + 1) + oy) * (Wout + 1) + ox) * p.ldc;
      ptx::mbar_wait(&tmem_full[s], ph, &g_im2col_dbg, 0x300u + (unsigned)it);
      ptx::tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)s * acc_stride;
      if (POOL) {
        const int np = p.npos;
        if (np == 4) {
          epilogue_pool_small<4>(trow, s_sc, s_sh, p.leaky, valid, dst, p.N, p.ldc, vec_store);
        } else if (np == 8) {
          epilogue_pool_small<8>(trow, s_sc, s_sh, p.leaky, valid, dst, p.N, p.ldc, vec_store);
        } else {
          for (int n0 = 0; n0 < np; n0 += 16) {
            float m[16], v[16];
            tmem_ld_x16(trow + n0, m);
#pragma unroll
            for (int j = 0; j < 16; ++j) m[j] = fmaf(m[j], s_sc[n0 + j], s_sh[n0 + j]);
            for (int pos = 1; pos < 4; ++pos) {
              tmem_ld_x16(trow + pos * np + n0, v);
#pragma unroll
              for (int j = 0; j < 16; ++j) m[j] = fmaxf(m[j], fmaf(v[j], s_sc[n0 + j], s_sh[n0 + j]));
            }
            if (p.leaky) {
#pragma unroll
              for (int j = 0; j < 16; ++j) m[j] = fmaxf(m[j], 0.1f * m[j]);
            }
            if (valid) {
              store_group(dst, n0, m, p.N, p.ldc, vec_store);
              store_group(dst, n0 + 8, m + 8, p.N, p.ldc, vec_store);
            }
          }
        }
      } else {
        for (int n0 = 0; n0 < p.nb; n0 += 16) {
          float v[16];
          tmem_ld_x16(trow + n0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = fmaf(v[j], s_sc[n0 + j], s_sh[n0 + j]);
            v[j] = p.leaky ? fmaxf(a, 0.1f * a) : a;
          }
          if (valid) {
            store_group(dst, n0, v, p.N, p.ldc, vec_store);
            store_group(dst, n0 + 8, v + 8, p.N, p.ldc, vec_store);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_free[s]);  // this warp has drained TMEM[s]
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

template <int CL, bool POOL, int FIRST>
int launch_im2col(const CUtensorMap& tm_b, const Im2colParams& p, cudaStream_t stream) {
  constexpr int PR = POOL ? 2 * TY + 2 : TY + 2, PC = POOL ? 2 * TX + 2 : TX + 2;
  constexpr int PCF = FIRST == 2 ? ((PC + 15 + 15) & ~15) : ((PC + 3 + 3) & ~3);  // see XS in the kernel
  constexpr int PEL = FIRST == 2 ? 1 : 4;
  constexpr int NPIX = POOL ? 16 : 9;
  constexpr int NKB = (NPIX * CL + 63) / 64;
  const int nb_pad = (p.nb + 15) & ~15;
  const size_t patch_bytes = FIRST ? (size_t)p.Cin * PR * PCF * PEL : (size_t)PR * PC * CL * 2;
  const size_t patch_stride = (patch_bytes + 127) & ~(size_t)127;
  const size_t smem = 2 * (size_t)NKB * 128 * 128 + (size_t)NKB * nb_pad * 128 + 2 * patch_stride + 128 + 2048 + 1024;
  if (smem > 227 * 1024) return mc_set_error(MC_ERR_SHAPE, "mc_conv_im2col_fwd: %zu B of shared memory", smem);

  // input patch tensor map
  CUtensorMap tm_in;
  int rc;
  if (FIRST) {  // fp32 / uint8 NCHW [B][Cin][H][W], box [1][Cin][PR][PCF]
    const uint64_t dims[4] = {(uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.Cin, (uint64_t)p.B};
    const uint64_t strides[3] = {(uint64_t)p.W * PEL, (uint64_t)p.W * p.H * PEL, (uint64_t)p.W * p.H * p.Cin * PEL};
    const uint32_t box[4] = {(uint32_t)PCF, (uint32_t)PR, (uint32_t)p.Cin, 1};
    if ((p.W * PEL) % 16 != 0)
      return mc_set_error(MC_ERR_SHAPE, "mc_conv_im2col_fwd: image rows must be a multiple of 16 bytes (width %d)", p.W);
    // short misaligned rows: 128 B promotion halves the L2->SM over-fetch of the default 256 B
    rc = mc_make_tmap(&tm_in, FIRST == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.in, dims,
                      strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  } else {      // bf16 PNHWC viewed as [B*(H+1)][W+1][CL], box [PR][PC][CL]
    const uint64_t dims[3] = {(uint64_t)CL, (uint64_t)(p.W + 1), (uint64_t)p.B * (p.H + 1)};
    const uint64_t strides[2] = {(uint64_t)CL * 2, (uint64_t)(p.W + 1) * CL * 2};
    const uint32_t box[3] = {(uint32_t)CL, (uint32_t)PC, (uint32_t)PR};
    rc = mc_make_tmap(&tm_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  }
  if (rc) return rc;

  auto kern = conv_im2col_tc_kernel<CL, POOL, FIRST>;
  static size_t attr_smem = 0;  // per template instantiation
  if (smem > attr_smem) {
    MC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int per_sm = (int)((200 * 1024) / smem);
  const int tmem_lim = 512 / p.tmem_cols;
  if (per_sm > tmem_lim) per_sm = tmem_lim;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)mc_num_sms() * per_sm;
  if (grid > p.total_tiles) grid = p.total_tiles;
  kern<<<(int)grid, 256, smem, stream>>>(tm_b, tm_in, p);
  MC_LAUNCH_CHECK("conv_im2col_tc_kernel");
  return 0;
}


// shared-memory bytes conv_im2col_tc_kernel needs for a geometry (mirrors the kernel's carve-up)
static size_t im2col_smem_bytes(int CL, int pool, int first, int nb_pad, int Cin) {
  const int PR = pool ? 2 * TY + 2 : TY + 2, PC = pool ? 2 * TX + 2 : TX + 2;
  const int PCF = first == 2 ? ((PC + 15 + 15) & ~15) : ((PC + 3 + 3) & ~3);
  const int NKB = ((pool ? 16 : 9) * CL + 63) / 64;
  const size_t patch_bytes = first ? (size_t)Cin * PR * PCF * (first == 2 ? 1 : 4) : (size_t)PR * PC * CL * 2;
  const size_t patch_stride = (patch_bytes + 127) & ~(size_t)127;
  return 2 * (size_t)NKB * 128 * 128 + (size_t)NKB * nb_pad * 128 + 2 * patch_stride + 128 + 2048 + 1024;
}

}  // namespace

// Debug: code of the first mbarrier wait that timed out in conv_im2col_tc_kernel since the last call (0 = none):
// 0x100+it patch TMA, 0x200 weight TMA, 0x300+it MMA commit.
extern "C" int mc_debug_im2col_timeout(void) {
  unsigned int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, g_im2col_dbg, sizeof(v)) != cudaSuccess) return -1;
  cudaMemcpyToSymbol(g_im2col_dbg, &zero, sizeof(zero));
  return (int)v;
}

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
  if (nb > 256) return 0;
  const int CL = in_is_nchw_f32 ? 4 : (Cin <= 8 ? 8 : 16);
  return im2col_smem_bytes(CL, pool, in_is_nchw_f32, (nb + 15) & ~15, Cin) <= 200 * 1024 ? 1 : 0;
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
  p.nsc = ((pool ? npos : ((N + 15) & ~15)) + 15) & ~15;  // the engine pads scale/shift to max(round_up(npos,16),16)
  p.nkb = kpad / 64;
  const int kelems = (pool ? 16 : 9) * CL;
  p.ksteps = (kelems + 15) / 16;
  const int Hout = pool ? H / 2 : H, Wout = pool ? W / 2 : W;
  p.tiles_x = (Wout + TX - 1) / TX;
  p.tiles_y = (Hout + TY - 1) / TY;
  p.total_tiles = B * p.tiles_x * p.tiles_y;
  const int nb_pad = (nb + 15) & ~15;
  int tc = 32;
  while (tc < 2 * ((nb_pad + 31) & ~31)) tc <<= 1;
  p.tmem_cols = tc;  // two accumulator stages
  p.nb = nb_pad;  // UMMA N must be a multiple of 16; padded columns are zero rows of d_wexp
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nb_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  CUtensorMap tm_b;
  rc = mc_make_tmap_2d_bf16(&tm_b, d_wexp, (uint64_t)nb_pad, (uint64_t)kpad, (uint64_t)kpad, (uint32_t)nb_pad);
  if (rc) return rc;

  if (in_is_nchw_f32 == 2) return pool ? launch_im2col<4, true, 2>(tm_b, p, stream) : launch_im2col<4, false, 2>(tm_b, p, stream);
  if (in_is_nchw_f32) return pool ? launch_im2col<4, true, 1>(tm_b, p, stream) : launch_im2col<4, false, 1>(tm_b, p, stream);
  if (CL == 8) return pool ? launch_im2col<8, true, 0>(tm_b, p, stream) : launch_im2col<8, false, 0>(tm_b, p, stream);
  return pool ? launch_im2col<16, true, 0>(tm_b, p, stream) : launch_im2col<16, false, 0>(tm_b, p, stream);
}
