// conv_tcgen05.cu — Darknet-19 conv + folded-BN + leaky-ReLU as an implicit GEMM on tcgen05/TMEM, fed by TMA.
//
// Replaces: MaskedConv2d.forward -> F.conv2d (src/pruning/weightPruning/layers.py:53-64) followed by
// nn.BatchNorm2d (eval) and nn.LeakyReLU(0.1) (src/nets.py:802,809), the Reorg module (src/nets.py:648-667)
// and the route concat (src/nets.py:735-746), which are folded into the store addressing of the epilogue.
//
// Formulation.  Activations live in "PNHWC" (see include/mcb200.h): a 2-D bf16 matrix [rows, C] in which a
// 3x3 tap (dy,dx) is the constant row offset dy*(W+1)+dx and out-of-image reads hit zero rows/columns (or TMA
// out-of-bounds zero fill at the ends of the buffer).  So
//     Y[p, n] = sum_{tap} sum_{c} X[p + off(tap), c] * Wt[n, tap*Kc + c]
// is a plain GEMM whose A-tile row coordinate is shifted per k-block: no im2col buffer, no bounds logic.
// M tile = 128 rows (UMMA M=128, cta_group::1), N tile = block_n (16..256), K block = 64 bf16 = one
// 128-byte swizzle row (or 32 bf16 / 64-byte swizzle for <= 32 input channels).  Warp roles: warp0 = TMA producer,
// warp1 = TMEM owner + MMA issuer (warp-uniform loop, one elected lane issues), warps 2..5 = epilogue (TMEM -> regs ->
// scale/shift/leaky -> bf16/fp32 stores, staged through shared memory for 32..96-wide PNHWC tiles).
// Two kernels share this file:
//   conv_gemm_tcgen05_kernel<MINB>   persistent, 1..3 CTAs per SM (MINB = register budget), every layer shape;
//   conv_gemm_tcgen05_pair_kernel    cluster of 2 CTAs, tcgen05 cta_group::2: 256-row tiles, each CTA loads half of every
//                                    weight tile — the wide 3x3 layers (block_n >= 192, >= 36 k-blocks).
// mc_conv_fwd picks kernel, tile width, ring depth and CTAs per SM (mc_conv_last_plan reports the choice;
// MCB200_CONV_TRACE=1 prints it; the MCB200_* switches listed in tools/ab.sh override single decisions for A/B runs).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "ptx_sm100.cuh"
#include "tmap.cuh"
#include "decode_math.cuh"

namespace {

constexpr int BLOCK_M = 128;  // k-block: 64 bf16 (128-byte swizzle rows) or 32 bf16 (64-byte swizzle rows), per launch
constexpr int MAX_STAGES = 8;
constexpr int MAX_A_STAGES = 8;   // activation-box ring: 3 beside a weight ring, up to 8 when the weights are resident
constexpr int DEF_A_STAGES = 3;
constexpr int A_BOX_ROWS = BLOCK_M + 8;          // rows m0-1 .. m0+134: the dx = -1, 0, +1 windows of a 128-row tile
constexpr int A_BOX_BYTES = A_BOX_ROWS * 128;   // 17,408 B landed by TMA
constexpr int A_BOX_STRIDE = 18 * 1024;         // ring pitch (1024-byte aligned for the 128-byte swizzle)
constexpr int MAX_ACC = 8;  // accumulator stages in TMEM (2 for wide tiles; up to 8 for narrow one-N-tile layers)
constexpr int NUM_THREADS = 192;

struct ConvKParams {
  int M_rows;      // B*(H+1)*(W+1)
  int N;           // valid output channels
  int Npad;        // length of scale/shift arrays
  int num_kb;      // ntaps * kb_per_tap
  int kb_per_tap;  // Kc / block_k
  int block_k;     // 64 or 32
  int ksize;       // 1 or 3
  int Wp, Hp, W, H;  // Wp = W+1, Hp = H+1
  int block_n, stages, tmem_cols, acc_stride;  // acc_stride: TMEM columns between accumulator stages
  int acc_stages;  // accumulator stages: the MMA issuer runs this many tiles ahead of the epilogue
  int out_pitch;   // > 0: PNHWC epilogue stages each warp's 32 rows in shared memory (row pitch in bytes) and writes them out coalesced
  int out_chunk;   // columns staged at a time (<= 64)
  int tma_bufs;    // > 0: PNHWC epilogue writes whole 64-column chunks with TMA stores from this many 4 KB slabs per warp
  float* stat_sum;    // training forward: per-channel sum / sum of squares of the STORED (bf16-rounded) outputs are
  float* stat_sumsq;  // accumulated from the TMA-store slabs (the batch statistics of BatchNorm); nullptr: off
  unsigned int wp_mul, wp_shr, hp_mul, hp_shr;  // magic-number division by Wp and Hp (row -> x, y, b)
  int last_ksteps;  // UMMA K steps (16 channels) that hold real channels in the LAST channel block of a tap
  int b_resident;   // share_dx only: all weight tiles stay in shared memory for the whole launch (one N tile, small K)
  int share_dx, a_stages;  // 3x3 only: one A box (136 rows) serves the three dx taps of a filter row; separate A / B rings
  int m_tiles, n_tiles;
  uint32_t idesc;
  const float* scale;
  const float* shift;
  void* out;
  int ldc, ch_off, epi_mode, leaky;
  // stream-K tail of the CTA-pair kernel (sk_units > 0): the units of the last, partial wave are cut along K into equal
  // spans, one per cluster; spans that do not end a unit leave their fp32 accumulator in the workspace
  int sk_first, sk_units;        // first stream-K unit (the units before it run one per cluster and round), their number
  int sk_gper, sk_total_g;       // K groups (one activation box = 3 taps) per unit; sk_units * sk_gper
  int sk_clusters;               // clusters that take a stream-K span (<= 3 * sk_units: a unit has at most 4 pieces)
  unsigned int* sk_count;        // [sk_units] arrivals of partial writers (8 epilogue warps each), zeroed per launch
  float* sk_part;                // [sk_units][3][block_n / 4][256 rows][4] partial accumulators
  // MC_EPI_DECODE (region decode fused into the head's epilogue)
  float* dec_boxes;
  float* dec_cls;
  float* dec_head;
  int dec_A, dec_nc, dec_only_obj;
  float dec_thresh;
  McAnchors dec_anc;
  unsigned long long* dbg;  // mc_debug_conv_trace: globaltimer stamps, 32 slots per CTA (nullptr: no tracing)
};

// Timeline stamps of the single-CTA kernel (slot layout in tools/trace_conv.py)
__device__ __forceinline__ void conv_stamp(const ConvKParams& p, int slot) {
  if (p.dbg != nullptr && slot < 32) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[(size_t)blockIdx.x * 32 + slot] = t;
  }
}

// 16 accumulator columns of one row: y = acc*scale + shift (-> leaky) ; rows outside the image become 0.
template <bool LEAKY>
__device__ __forceinline__ void scale_act16(const uint32_t (&r)[16], const float* __restrict__ sc,
                                            const float* __restrict__ sh, bool interior, float (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(sc + 4 * q);  // same address in every lane: smem broadcast
    const float4 b = *reinterpret_cast<const float4*>(sh + 4 * q);
    v[4 * q + 0] = fmaf(__uint_as_float(r[4 * q + 0]), a.x, b.x);
    v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), a.y, b.y);
    v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), a.z, b.z);
    v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), a.w, b.w);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (LEAKY) v[j] = fmaxf(v[j], 0.1f * v[j]);
    v[j] = interior ? v[j] : 0.f;
  }
}

template <int MODE>
__device__ __forceinline__ void store16(const ConvKParams& p, const float (&v)[16], int nbase, long long out_row_base,
                                        bool vec_ok) {
  if (MODE == MC_EPI_NCHW_F32 || MODE == MC_EPI_DECODE) {
    float* out_f = reinterpret_cast<float*>(p.out);
    const long long hw = (long long)p.H * p.W;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (nbase + j < p.N) out_f[out_row_base + (long long)(nbase + j) * hw] = v[j];
  } else {
    __nv_bfloat16* out_bf = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row_base;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int n = nbase + g * 8;
      if (n >= p.N) continue;
      // channels >= N of a PNHWC row are zero padding up to the pitch: a partial 8-group is still written as one
      // 16-byte store (zeros there) when it fits inside the pitch
      const bool full = n + 8 <= p.N;
      if (vec_ok && (full || (MODE == MC_EPI_PNHWC && p.ch_off + n + 8 <= p.ldc))) {
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = (full || n + j < p.N) ? v[g * 8 + j] : 0.f;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(w[0], w[1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(w[2], w[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(w[4], w[5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(w[6], w[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(out_bf + n) = pk;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n + j < p.N) out_bf[n + j] = __float2bfloat16_rn(v[g * 8 + j]);
      }
    }
  }
}

// 16 fp32 -> 16 bf16 -> two 16-byte shared-memory stores
__device__ __forceinline__ void pack16_to_smem(const float (&v)[16], uint8_t* dst) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[g * 8 + 0], v[g * 8 + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[g * 8 + 2], v[g * 8 + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[g * 8 + 4], v[g * 8 + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[g * 8 + 6], v[g * 8 + 7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2);
    pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + g * 16) = pk;
  }
}

// 16 fp32 -> 16 bf16 -> the two 16-byte chunks (index chunk, chunk + 1) of this lane's 128-byte row of a TMA store
// slab.  The slab has the tensor map's 128-byte swizzle: chunk j of row r sits at chunk position j ^ (r & 7).
__device__ __forceinline__ void pack16_to_slab(const float (&v)[16], uint8_t* row, int chunk, int lane) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[g * 8 + 0], v[g * 8 + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[g * 8 + 2], v[g * 8 + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[g * 8 + 4], v[g * 8 + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[g * 8 + 6], v[g * 8 + 7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2);
    pk.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(row + (((chunk + g) ^ (lane & 7)) << 4)) = pk;
  }
}

// Batch statistics from a finished 32-row x 64-column slab: lane L owns columns 2L, 2L+1 (one 4-byte word per row: the
// 32 lanes read 32 different words of a 128-byte row, conflict-free) and adds the 32 rows into this warp's private
// shared-memory accumulators acc[0][col] (sum) / acc[1][col] (sum of squares); pad rows hold zeros.
__device__ __forceinline__ void slab_stats(const uint8_t* sb, float* acc /*[2][256]*/, int col0, int lane) {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const int j = lane >> 2, w = lane & 3;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(sb + r * 128 + ((j ^ (r & 7)) << 4) + w * 4);
    const float a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xffff0000u);
    s0 += a; s1 += b;
    q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1);
  }
  acc[col0 + 2 * lane] += s0;
  acc[col0 + 2 * lane + 1] += s1;
  acc[256 + col0 + 2 * lane] += q0;
  acc[256 + col0 + 2 * lane + 1] += q1;
}
// the warp's accumulated statistics of N tile n0 -> global (one atomic per column, statistic and warp), then zero
__device__ __forceinline__ void flush_stats(const ConvKParams& p, float* acc, int n0, int lane) {
  for (int c = lane; c < p.block_n; c += 32) {
    if (n0 + c < p.N) {
      atomicAdd(p.stat_sum + n0 + c, acc[c]);
      atomicAdd(p.stat_sumsq + n0 + c, acc[256 + c]);
    }
    acc[c] = 0.f;
    acc[256 + c] = 0.f;
  }
  __syncwarp();
}

// PNHWC epilogue of one accumulator tile through TMA stores.  A warp owns 32 rows; every whole 64-column chunk of the
// tile is converted into the warp's 4 KB slab (32 rows x 128 B, swizzled like the tensor map of the output) and leaves
// as ONE bulk tensor store — 128-byte rows, issued by one lane, asynchronous — instead of 16-byte pieces one row pitch
// apart per thread (measured: ~38 cycles per output column and tile, 4..7.6 us for a 208..256-wide tile, which
// throttles short-K layers and is exposed after the last tile of every CTA).  With two slabs per warp the conversion of
// chunk i+1 overlaps the store of chunk i.  The tile's last block_n % 64 columns take direct stores; rows and columns
// outside the output tensor are clipped by TMA.  add(r, col) lets a stream-K finisher add partial accumulators.
template <bool LEAKY, bool PAIR, bool WIDE_LD, typename AddFn>
__device__ __forceinline__ void epilogue_tile_tma(const ConvKParams& p, const CUtensorMap* tm_out, uint32_t taddr_row,
                                                  const float* sc, const float* sh, bool interior, bool in_buf,
                                                  int row0, int n0, long long out_row_base, bool vec_ok, uint8_t* slab,
                                                  int& buf, uint64_t* empty_bar, AddFn add, float* stat_acc = nullptr) {
  const int lane = threadIdx.x & 31;
  int cols_valid = ((p.N - n0 + 7) >> 3) << 3;  // whole 8-groups up to the last one holding a real channel
  if (cols_valid > p.block_n) cols_valid = p.block_n;
  const int full_cols = p.block_n & ~63;
  for (int cc = 0; cc < full_cols && cc < cols_valid; cc += 64) {
    uint8_t* sb = slab + (size_t)buf * 4096;
    if (lane == 0) {  // the store that last read this slab has finished reading it
      if (p.tma_bufs >= 2) ptx::bulk_wait_group_read<1>();
      else ptx::bulk_wait_group_read<0>();
    }
    __syncwarp();
    uint8_t* mine = sb + lane * 128;
    if (WIDE_LD) {
      // all 64 columns of the chunk in flight before the one wait (register budget of the 1-CTA-per-SM variants)
      uint32_t r[4][16];
#pragma unroll
      for (int q = 0; q < 4; ++q) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(cc + 16 * q), r[q]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        add(r[q], cc + 16 * q);
        float v[16];
        scale_act16<LEAKY>(r[q], sc + cc + 16 * q, sh + cc + 16 * q, interior, v);
        pack16_to_slab(v, mine, 2 * q, lane);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[16], r1[16];
        ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(cc + 32 * h), r0);
        ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(cc + 32 * h + 16), r1);
        ptx::tmem_ld_wait();
        add(r0, cc + 32 * h);
        add(r1, cc + 32 * h + 16);
        float v[16];
        scale_act16<LEAKY>(r0, sc + cc + 32 * h, sh + cc + 32 * h, interior, v);
        pack16_to_slab(v, mine, 4 * h, lane);
        scale_act16<LEAKY>(r1, sc + cc + 32 * h + 16, sh + cc + 32 * h + 16, interior, v);
        pack16_to_slab(v, mine, 4 * h + 2, lane);
      }
    }
    ptx::fence_proxy_async_smem();  // this lane's slab writes -> visible to the TMA (async proxy)
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm_out, sb, n0 + cc, row0);
      ptx::bulk_commit_group();
    }
    if (stat_acc != nullptr) slab_stats(sb, stat_acc, cc, lane);
    if (p.tma_bufs >= 2) buf ^= 1;
  }
  for (int c0 = full_cols; c0 < p.block_n; c0 += 32) {
    uint32_t r0[16], r1[16];
    const bool two = c0 + 16 < p.block_n;
    ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)c0, r0);
    if (two) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(c0 + 16), r1);
    ptx::tmem_ld_wait();
    if (!in_buf) continue;
    add(r0, c0);
    if (two) add(r1, c0 + 16);
    float v[16];
    scale_act16<LEAKY>(r0, sc + c0, sh + c0, interior, v);
    store16<MC_EPI_PNHWC>(p, v, n0 + c0, out_row_base, vec_ok);
    if (two) {
      scale_act16<LEAKY>(r1, sc + c0 + 16, sh + c0 + 16, interior, v);
      store16<MC_EPI_PNHWC>(p, v, n0 + c0 + 16, out_row_base, vec_ok);
    }
  }
  // this warp has finished reading the accumulator stage: hand it back to the MMA issuer
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if (PAIR) ptx::mbar_arrive_leader(empty_bar);
    else ptx::mbar_arrive(empty_bar);
  }
}

// Epilogue warps (threads 64..191): per tile, stage the tile's scale/shift in shared memory (double-buffered with
// the accumulator stage, one named barrier per tile), then drain the accumulator 32 columns at a time.
// PAIR (CTA-pair kernel): `tile` counts (m-tile pair, n-tile) units of the cluster, this CTA owns m-tile 2*pair + rank,
// and the accumulator stage is handed back on the LEADER CTA's barrier (the MMA issuer waits for both epilogues).
template <int MODE, bool LEAKY, bool PAIR = false, bool WIDE_LD = false>
__device__ __forceinline__ void epilogue_loop(const ConvKParams& p, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                              uint64_t* tmem_empty_bar, float* s_ss, uint8_t* s_out = nullptr,
                                              const CUtensorMap* tm_out = nullptr,
                                              int tile0 = -1, int tile_step = 0, int total_units = 0, int pair_rank = 0) {
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int et = threadIdx.x - 64;   // 0..127
  const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
  const int total_tiles = PAIR ? total_units : p.m_tiles * p.n_tiles;
  if (!PAIR) { tile0 = blockIdx.x; tile_step = gridDim.x; }
  const bool vec_ok = ((p.ldc | p.ch_off) & 7) == 0 && (MODE != MC_EPI_REORG2 || (p.N & 7) == 0);
  int as = 0;
  uint32_t aphase = 0;
  int slab_buf = 0;  // (TMA-store epilogue) slab of this warp the next chunk goes to
  // (batch statistics) this warp's accumulators behind the slabs; flushed whenever the N tile changes and at the end
  float* stat_acc = nullptr;
  int stat_n0 = -1;
  if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0 && p.stat_sum != nullptr) {
    stat_acc = reinterpret_cast<float*>(s_out + (size_t)4 * p.tma_bufs * 4096) + quarter * 512;
    for (int c = lane; c < 512; c += 32) stat_acc[c] = 0.f;
    __syncwarp();
  }
  // one N tile: every tile of the launch uses the same scale/shift slice, staged once (a narrow layer's epilogue is
  // otherwise a chain of global-load latency + barrier per 128 rows)
  const bool one_n_tile = p.n_tiles == 1;
  if (one_n_tile) {
    for (int i = et; i < p.block_n; i += 128) {
      const bool ok = i < p.Npad;
      s_ss[i] = ok ? __ldg(p.scale + i) : 0.f;
      s_ss[256 + i] = ok ? __ldg(p.shift + i) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  int ti = 0;  // (tracing) tiles done by this CTA
  for (int tile = tile0; tile < total_tiles; tile += tile_step, ++ti) {
    const int n0 = (tile % p.n_tiles) * p.block_n;
    const int m0 = (PAIR ? 2 * (tile / p.n_tiles) + pair_rank : tile / p.n_tiles) * BLOCK_M;
    float* sc = s_ss + (one_n_tile ? 0 : as * 512);
    float* sh = sc + 256;
    if (!one_n_tile) {
      for (int i = et; i < p.block_n; i += 128) {
        const bool ok = (n0 + i) < p.Npad;
        sc[i] = ok ? __ldg(p.scale + n0 + i) : 0.f;
        sh[i] = ok ? __ldg(p.shift + n0 + i) : 0.f;
      }
    }
    const int row = m0 + quarter * 32 + lane;
    const int t = (int)(__umulhi((unsigned int)row, p.wp_mul) >> p.wp_shr);  // row / Wp (magic-number division)
    const int x = row - t * p.Wp;
    const int b = (int)(__umulhi((unsigned int)t, p.hp_mul) >> p.hp_shr);    // t / Hp
    const int y = t - b * p.Hp;
    const bool in_buf = row < p.M_rows;
    const bool interior = in_buf && (x < p.W) && (y < p.H);
    long long out_row_base = 0;
    bool do_store = false;
    if (MODE == MC_EPI_PNHWC) {
      out_row_base = (long long)row * p.ldc + p.ch_off;
      do_store = in_buf;  // pad rows are written as zeros to keep the layout invariant
    } else if (MODE == MC_EPI_REORG2) {
      const int Wo = p.W / 2 + 1, Ho = p.H / 2 + 1;
      const long long orow = ((long long)b * Ho + (y >> 1)) * Wo + (x >> 1);
      out_row_base = orow * p.ldc + p.ch_off + ((y & 1) * 2 + (x & 1)) * p.N;
      do_store = interior;
    } else {  // MC_EPI_NCHW_F32
      out_row_base = ((long long)b * p.N * p.H + y) * p.W + x;  // + n*H*W
      do_store = interior;
    }
    if (!one_n_tile) asm volatile("bar.sync 1, 128;" ::: "memory");  // scale/shift of this tile visible to the 4 epilogue warps

    ptx::mbar_wait(&tmem_full_bar[as], aphase);
    ptx::tc_fence_after();
    if (!PAIR && et == 0 && ti < 8) conv_stamp(p, 12 + 2 * ti);
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.acc_stride);

    if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0) {
      if (stat_acc != nullptr && n0 != stat_n0) {
        if (stat_n0 >= 0) flush_stats(p, stat_acc, stat_n0, lane);
        stat_n0 = n0;
      }
      epilogue_tile_tma<LEAKY, PAIR, WIDE_LD>(p, tm_out, taddr_row, sc, sh, interior, in_buf, m0 + quarter * 32, n0, out_row_base,
                                     vec_ok, s_out + (size_t)quarter * p.tma_bufs * 4096, slab_buf, &tmem_empty_bar[as],
                                     [](uint32_t (&)[16], int) {}, stat_acc);
      if (!PAIR && et == 0 && ti < 8) conv_stamp(p, 13 + 2 * ti);
      if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      continue;
    }

    if (MODE == MC_EPI_PNHWC && s_out != nullptr) {
      // Staged store: a thread owns one output ROW, so direct stores are 16-byte pieces one row pitch apart (32 half-
      // filled sectors per warp instruction: measured ~38 cycles per output column and tile).  Instead the warp parks
      // its 32 rows x block_n bf16 in shared memory (pitch +16 B: conflict-free), hands the TMEM stage back at once,
      // and writes the rows out with consecutive lanes on consecutive 16-byte chunks.
      // Tiles wider than the staging buffer (64 columns) go through it in 64-column chunks.
      const int pitch = p.out_pitch;
      const int chunk = p.out_chunk;
      uint8_t* wbuf = s_out + (size_t)quarter * 32 * pitch;
      uint8_t* mine = wbuf + (size_t)lane * pitch;
      const int row0 = m0 + quarter * 32;
      int cols_left = ((p.N - n0 + 7) >> 3) << 3;  // whole 8-groups up to the last one holding a real channel
      if (cols_left > p.block_n) cols_left = p.block_n;
      for (int cc = 0; cc < p.block_n; cc += chunk) {
        const int cw = (p.block_n - cc < chunk) ? p.block_n - cc : chunk;
        for (int c0 = 0; c0 < cw; c0 += 32) {
          uint32_t r0[16], r1[16];
          const bool two = c0 + 16 < cw;
          ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(cc + c0), r0);
          if (two) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(cc + c0 + 16), r1);
          ptx::tmem_ld_wait();
          float v[16];
          scale_act16<LEAKY>(r0, sc + cc + c0, sh + cc + c0, interior, v);
          pack16_to_smem(v, mine + c0 * 2);
          if (two) {
            scale_act16<LEAKY>(r1, sc + cc + c0 + 16, sh + cc + c0 + 16, interior, v);
            pack16_to_smem(v, mine + (c0 + 16) * 2);
          }
        }
        if (cc + chunk >= p.block_n) {  // last chunk read: hand the TMEM stage back before the copy-out
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
            else ptx::mbar_arrive(&tmem_empty_bar[as]);
          }
        } else {
          __syncwarp();
        }
        int cols_w = cols_left - cc;
        if (cols_w > cw) cols_w = cw;
        if (cols_w > 0) {
          const int cpr = cols_w >> 3;  // 16-byte chunks per row
          int r = lane / cpr, c = lane - r * cpr;
          const int dr = 32 / cpr, dc = 32 - dr * cpr;
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + p.ch_off + n0 + cc;
          while (r < 32) {
            if (row0 + r < p.M_rows)
              *reinterpret_cast<uint4*>(obase + (long long)(row0 + r) * p.ldc + c * 8) =
                  *reinterpret_cast<const uint4*>(wbuf + (size_t)r * pitch + c * 16);
            r += dr; c += dc;
            if (c >= cpr) { c -= cpr; ++r; }
          }
        }
        __syncwarp();  // wbuf is rewritten by the next chunk / tile
      }
      if (!PAIR && et == 0 && ti < 8) conv_stamp(p, 13 + 2 * ti);
      if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      continue;
    }

    if (MODE == MC_EPI_DECODE) {
      // Region decode in the epilogue (get_region_boxes, src/nets2_utils.py:158-205).  The thread owns pixel (b, y, x):
      // its A*(5+nc) logits (= acc*scale + shift, the head's bias) are parked in the warp's shared-memory rows (row
      // pitch block_n+1 floats: lanes hit different banks), the TMEM stage goes back to the MMA issuer at once, and the
      // thread then decodes its A anchors from shared memory with the arithmetic of decode_math.cuh.
      const int pitch_f = p.block_n + 1;
      float* mine = reinterpret_cast<float*>(s_out) + (size_t)(quarter * 32 + lane) * pitch_f;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t r0[16];
        ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)c0, r0);
        ptx::tmem_ld_wait();
        float v[16];
        scale_act16<LEAKY>(r0, sc + c0, sh + c0, true, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) mine[c0 + j] = v[j];
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[as]);
      if (interior) {
        const int A = p.dec_A, nc = p.dec_nc, F = 5 + p.dec_nc;
        const long long P = (long long)p.H * p.W * A;
        const long long slot0 = (long long)(y * p.W + x) * A;
        if (p.dec_head != nullptr) {
          float* hd = p.dec_head + ((long long)b * p.N * p.H + y) * p.W + x;
          const long long hw = (long long)p.H * p.W;
          for (int n = 0; n < p.N; ++n) hd[n * hw] = mine[n];
        }
        for (int a = 0; a < A; ++a) {
          const float* row = mine + a * F;
          const McDecoded d = mc_decode_anchor([row](int f) { return row[f]; }, nc, x, y, p.W, p.H, p.dec_anc.w[a],
                                               p.dec_anc.h[a], p.dec_thresh, p.dec_only_obj);
          float* dst = p.dec_boxes + ((long long)b * P + slot0 + a) * 8;
          reinterpret_cast<float4*>(dst)[0] = make_float4(d.bx, d.by, d.bw, d.bh);
          reinterpret_cast<float4*>(dst)[1] =
              make_float4(d.conf, d.cmax, (float)d.cid, d.cand ? (float)(slot0 + a) : -1.0f);
          if (d.cand && p.dec_cls != nullptr) {
            float* cd = p.dec_cls + ((long long)b * P + slot0 + a) * nc;
            for (int c = 0; c < nc; ++c) cd[c] = __fdiv_rn(expf(__fsub_rn(row[5 + c], d.cls_max_logit)), d.cls_sum);
          }
        }
      }
      __syncwarp();  // the rows are rewritten by the next tile
      if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      continue;
    }

    for (int c0 = 0; c0 < p.block_n; c0 += 32) {
      uint32_t r0[16], r1[16];
      const bool two = c0 + 16 < p.block_n;
      ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)c0, r0);
      if (two) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(c0 + 16), r1);
      ptx::tmem_ld_wait();
      if (!do_store) continue;
      float v[16];
      scale_act16<LEAKY>(r0, sc + c0, sh + c0, interior, v);
      store16<MODE>(p, v, n0 + c0, out_row_base, vec_ok);
      if (two) {
        scale_act16<LEAKY>(r1, sc + c0 + 16, sh + c0 + 16, interior, v);
        store16<MODE>(p, v, n0 + c0 + 16, out_row_base, vec_ok);
      }
    }
    // this warp has finished reading the accumulator stage: hand it back to the MMA issuer
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (PAIR) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
      else ptx::mbar_arrive(&tmem_empty_bar[as]);
    }
    if (!PAIR && et == 0 && ti < 8) conv_stamp(p, 13 + 2 * ti);
    if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
  }
  if (stat_acc != nullptr && stat_n0 >= 0) flush_stats(p, stat_acc, stat_n0, lane);
  if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0 && lane == 0) ptx::bulk_wait_group_read<0>();  // slabs read before the CTA exits
}

// Persistent: grid = min(#tiles, #SMs); CTA c works on tiles c, c+grid, ...  The smem ring runs across tiles, and the
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// MINB (resident CTAs per SM the register budget allows): 1 for the wide compute-bound tiles; 3 for narrow layers
// (few output channels / short K), whose per-tile chain TMA -> MMA -> commit -> epilogue -> TMEM hand-back is latency
// bound inside one CTA — several CTAs per SM (each with its own small smem ring and TMEM slice) overlap those chains.
template <int MINB>
__global__ void __launch_bounds__(NUM_THREADS, MINB)
conv_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_abox, const __grid_constant__ CUtensorMap tmap_out,
                         const ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [stages x (A 16KB | B block_n*128)] | barriers | tmem ptr | scale/shift staging (4 KB)
  const uint32_t a_tile_bytes = (uint32_t)(BLOCK_M * p.block_k * 2);
  const uint32_t b_tile_bytes = (uint32_t)(p.block_n * p.block_k * 2);
  const uint32_t stage_bytes = a_tile_bytes + b_tile_bytes;
  uint8_t* smem = smem_raw;
  {  // dynamic smem base is only guaranteed 16B aligned: align up to 1024 (host adds 1 KB of slack)
    uint32_t a = ptx::smem_u32(smem);
    smem += (1024u - (a & 1023u)) & 1023u;
  }
  uint8_t* tiles = smem;
  // share_dx: [a_stages x A box (18 KB pitch)] [stages x B tile]; else [stages x (A | B)]
  // shared activation box: 136 rows of one k-block row (128 B with 64-wide k-blocks / 128-byte swizzle, 64 B with 32-wide
  // k-blocks / 64-byte swizzle); ring pitch 18 KB / 9 KB keeps every box 1024-byte aligned
  const uint32_t a_row_bytes = (uint32_t)p.block_k * 2u;
  const uint32_t a_box_bytes = (uint32_t)A_BOX_ROWS * a_row_bytes;
  const uint32_t a_box_stride = p.block_k == 64 ? (uint32_t)A_BOX_STRIDE : (uint32_t)A_BOX_STRIDE / 2u;
  uint8_t* b_ring = tiles + (size_t)p.a_stages * a_box_stride;
  uint8_t* aux = p.share_dx ? b_ring + (size_t)(p.b_resident ? p.num_kb : p.stages) * b_tile_bytes
                            : tiles + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;   // [MAX_ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + MAX_ACC; // [MAX_ACC]
  uint64_t* afull_bar = tmem_empty_bar + MAX_ACC;     // [MAX_A_STAGES]
  uint64_t* aempty_bar = afull_bar + MAX_A_STAGES;    // [MAX_A_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aempty_bar + MAX_A_STAGES);
  float* s_ss = reinterpret_cast<float*>(aux + 512);  // [2 acc stages][scale|shift][256]
  uint8_t* s_out = aux + 5120;                        // staged epilogue: [4 warps][32 rows][out_pitch]; TMA-store slabs
                                                      // [4 warps][tma_bufs][4 KB] (1024-byte aligned: swizzle pattern)

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  if (threadIdx.x == 0) conv_stamp(p, 0);

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < p.acc_stages; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
    }
    for (int a = 0; a < MAX_A_STAGES; ++a) {
      ptx::mbar_init(&afull_bar[a], 1);
      ptx::mbar_init(&aempty_bar[a], 1);
    }
    ptx::prefetch_tensormap(&tmap_abox);
    if (p.tma_bufs > 0) ptx::prefetch_tensormap(&tmap_out);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) conv_stamp(p, 1);

  if (warp_idx == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0 && p.share_dx) {
      // A: one 136-row box per (filter row dy, channel block) — the three dx taps read it at row offsets 0, 1, 2, so
      // the activations cross L2 -> smem 3 times per tile instead of 9.  B: one tile per tap, its own ring.
      int s = 0, sa = 0;
      uint32_t phase = 0, pha = 0;
      if (p.b_resident) {  // every weight tile of the layer, once: [num_kb][block_n rows x 128 B]
        ptx::mbar_arrive_expect_tx(&full_bar[0], (uint32_t)p.num_kb * b_tile_bytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          ptx::tma_load_2d(b_ring + (size_t)kb * b_tile_bytes, &tmap_b, &full_bar[0], kb * p.block_k, 0);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * p.block_n;
        const int m0 = (tile / p.n_tiles) * BLOCK_M;
        for (int g = 0; g < 3 * p.kb_per_tap; ++g) {
          const int dy = g / p.kb_per_tap, cb = g - dy * p.kb_per_tap;
          ptx::mbar_wait(&aempty_bar[sa], pha ^ 1u);
          ptx::mbar_arrive_expect_tx(&afull_bar[sa], a_box_bytes);
          ptx::tma_load_2d(tiles + (size_t)sa * a_box_stride, &tmap_abox, &afull_bar[sa], cb * p.block_k,
                           m0 + (dy - 1) * p.Wp - 1);
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
          if (p.b_resident) continue;
          for (int dx = 0; dx < 3; ++dx) {
            const int kb = (dy * 3 + dx) * p.kb_per_tap + cb;
            ptx::mbar_wait(&empty_bar[s], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[s], b_tile_bytes);
            ptx::tma_load_2d(b_ring + (size_t)s * b_tile_bytes, &tmap_b, &full_bar[s], kb * p.block_k, n0);
            if (++s == p.stages) { s = 0; phase ^= 1u; }
          }
        }
      }
      conv_stamp(p, 3);
    } else if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_tiles) * p.block_n;
        const int m0 = (tile / p.n_tiles) * BLOCK_M;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          int row_off = 0;
          if (p.ksize == 3) row_off = (tap / 3 - 1) * p.Wp + (tap % 3 - 1);
          ptx::mbar_wait(&empty_bar[s], phase ^ 1u);
          uint8_t* a_dst = tiles + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + a_tile_bytes;
          ptx::mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          ptx::tma_load_2d(a_dst, &tmap_a, &full_bar[s], cb * p.block_k, m0 + row_off);
          ptx::tma_load_2d(b_dst, &tmap_b, &full_bar[s], kb * p.block_k, n0);
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
      }
      conv_stamp(p, 3);
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp walks the loop (warp-uniform control flow and operands, so descriptors live in uniform
    // registers) and one elected lane issues the tcgen05 instructions.  With the loop inside `if (lane == 0)` the
    // compiler wraps every UTCHMMA in a vote/R2UR.BROADCAST loop (~25 instructions), and that single thread's issue
    // rate — not the tensor pipe — bounds the k-block time.
    if (p.share_dx) {
      int s = 0, sa = 0, as = 0;
      uint32_t phase = 0, pha = 0, aphase = 0;
      bool b_loaded = false;
      int ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * p.acc_stride);
        for (int g = 0; g < 3 * p.kb_per_tap; ++g) {
          const int dy = g / p.kb_per_tap, cb = g - dy * p.kb_per_tap;
          // K steps beyond the layer's real channels multiply TMA zero fill by zero weights: not issued
          const int nk = (cb == p.kb_per_tap - 1) ? p.last_ksteps : (int)(a_row_bytes >> 5);
          ptx::mbar_wait(&afull_bar[sa], pha);
          if (ti == 0 && g == 0 && lane == 0) conv_stamp(p, 4);
          const uint32_t a_addr = ptx::smem_u32(tiles + (size_t)sa * a_box_stride);
          for (int dx = 0; dx < 3; ++dx) {
            uint32_t b_addr;
            if (p.b_resident) {
              if (!b_loaded) { ptx::mbar_wait(&full_bar[0], 0); b_loaded = true; }
              b_addr = ptx::smem_u32(b_ring + (size_t)((dy * 3 + dx) * p.kb_per_tap + cb) * b_tile_bytes);
            } else {
              ptx::mbar_wait(&full_bar[s], phase);
              b_addr = ptx::smem_u32(b_ring + (size_t)s * b_tile_bytes);
            }
            ptx::tc_fence_after();
            // the window of tap dx starts dx rows (dx * 128 B) into the box.  The start is then not aligned to the
            // 1024-byte swizzle pattern; measured on B200: the tensor core applies the 128-byte swizzle to the absolute
            // shared-memory address (as TMA did when it wrote the box), so the descriptor's base-offset field must stay
            // 0 — setting it to dx (or 8-dx) reads garbage (tools/debug_share.py)
            // (32-wide k-blocks: the same holds for the 64-byte swizzle — rows of 64 B, window dx starts dx * 64 B in)
            const bool k64 = p.block_k == 64;
            const uint64_t adesc = k64 ? ptx::make_sw128_kmajor_desc(a_addr + (uint32_t)dx * 128u)
                                       : ptx::make_sw64_kmajor_desc(a_addr + (uint32_t)dx * 64u);
            const uint64_t bdesc = k64 ? ptx::make_sw128_kmajor_desc(b_addr) : ptx::make_sw64_kmajor_desc(b_addr);
            if (ptx::elect_one()) {
              // (compile-time trip count with a per-step predicate: this issue loop is on the critical path, a
              //  run-time trip count costs the narrow-Cin layers ~4 us each)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk)
                  ptx::umma_bf16_ss(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                    (g > 0 || dx > 0 || k > 0) ? 1u : 0u);
              if (!p.b_resident) ptx::umma_commit(&empty_bar[s]);
              if (dx == 2) ptx::umma_commit(&aempty_bar[sa]);  // the box may be refilled once its three taps retire
              if (dx == 2 && g == 3 * p.kb_per_tap - 1) ptx::umma_commit(&tmem_full_bar[as]);
            }
            __syncwarp();
            if (!p.b_resident) {
              if (++s == p.stages) { s = 0; phase ^= 1u; }
            }
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
        }
        if (lane == 0 && ti < 6) conv_stamp(p, 5 + ti);
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      }
    } else {
      int s = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);  // epilogue has drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * p.acc_stride);
        int cb_i = 0;  // channel block inside the tap
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[s], phase);
          ptx::tc_fence_after();
          if (ti == 0 && kb == 0 && lane == 0) conv_stamp(p, 4);
          const uint32_t a_addr = ptx::smem_u32(tiles + (size_t)s * stage_bytes);
          const bool k64 = p.block_k == 64;
          const uint64_t adesc = k64 ? ptx::make_sw128_kmajor_desc(a_addr) : ptx::make_sw64_kmajor_desc(a_addr);
          const uint64_t bdesc = k64 ? ptx::make_sw128_kmajor_desc(a_addr + a_tile_bytes)
                                     : ptx::make_sw64_kmajor_desc(a_addr + a_tile_bytes);
          const int ksteps = (cb_i == p.kb_per_tap - 1) ? p.last_ksteps : p.block_k / 16;
          if (++cb_i == p.kb_per_tap) cb_i = 0;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr>>4) field
              if (k < ksteps)
                ptx::umma_bf16_ss(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
            if (kb == p.num_kb - 1) ptx::umma_commit(&tmem_full_bar[as]);  // accumulator complete
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; phase ^= 1u; }
        }
        if (lane == 0 && ti < 6) conv_stamp(p, 5 + ti);
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..5 =====================
    if (p.epi_mode == MC_EPI_PNHWC) {
      uint8_t* so = (p.out_pitch || p.tma_bufs) ? s_out : nullptr;
      if (p.leaky) epilogue_loop<MC_EPI_PNHWC, true, false, MINB == 1>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, so, &tmap_out);
      else epilogue_loop<MC_EPI_PNHWC, false, false, MINB == 1>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, so, &tmap_out);
    } else if (p.epi_mode == MC_EPI_REORG2) {
      if (p.leaky) epilogue_loop<MC_EPI_REORG2, true>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss);
      else epilogue_loop<MC_EPI_REORG2, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss);
    } else if (p.epi_mode == MC_EPI_DECODE) {
      epilogue_loop<MC_EPI_DECODE, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, s_out);
    } else {
      if (p.leaky) epilogue_loop<MC_EPI_NCHW_F32, true>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss);
      else epilogue_loop<MC_EPI_NCHW_F32, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) conv_stamp(p, 30);
  if (warp_idx == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// Work items of one cluster of the CTA-pair kernel: whole units first (cluster c: units c, c+G, ... < sk_first), then
// its stream-K span.  Stream-K: the sk_units units of the partial last wave hold sk_total_g K groups; cluster c (of the
// first G' = min(G, 3 * sk_units) clusters) owns groups [c*T/G', (c+1)*T/G') of the unit-major sequence, i.e. the END of
// one unit and / or the START of the next (a span is shorter than a unit, so it touches at most two; a unit is cut into
// at most four pieces).  The START piece runs first and leaves its accumulator in the workspace; the END piece is the
// unit's finisher: it adds the (at most three) earlier pieces and runs the epilogue.  The
// earlier pieces were the FIRST thing their clusters did, so the finisher practically never waits.
struct PairWork {
  int unit;    // unit index (m-pair, n-tile)
  int g0, g1;  // K groups [g0, g1) of the unit
  int kind;    // 0 whole unit, 1 partial (dump to slot), 2 finisher (add n partials)
  int aux;     // kind 1: slot; kind 2: number of partial pieces to add
};

struct PairWorkIter {
  int next_full, step, sk_first;
  int sk_state;  // 0: stream-K pieces not started, 1: first piece done, 2: finished
  int c, G, gper, T, units;
  __device__ __forceinline__ void init(const ConvKParams& p, int cluster_id, int num_clusters) {
    next_full = cluster_id;
    step = num_clusters;
    sk_first = p.sk_first;
    c = cluster_id; G = p.sk_clusters; gper = p.sk_gper; T = p.sk_total_g; units = p.sk_units;
    sk_state = (p.sk_units > 0 && c < G) ? 0 : 2;
  }
  __device__ __forceinline__ int span_begin(int cl) const { return (int)(((long long)cl * T) / G); }
  __device__ __forceinline__ void describe(int u, int a, int b, PairWork& w) const {  // groups [a,b) of stream-K unit u
    w.unit = sk_first + u;
    w.g0 = a;
    w.g1 = b;
    // first cluster whose span reaches into unit u
    int cf = c;
    while (cf > 0 && span_begin(cf) > u * gper) --cf;
    if (b == gper) {
      w.kind = (a == 0) ? 0 : 2;
      w.aux = c - cf;
    } else {
      w.kind = 1;
      w.aux = c - cf;
    }
  }
  __device__ __forceinline__ bool next(PairWork& w) {
    if (next_full < sk_first) {
      w.unit = next_full; w.g0 = 0; w.g1 = -1; w.kind = 0; w.aux = 0;  // g1 < 0: the whole K range
      next_full += step;
      return true;
    }
    if (sk_state == 2) return false;
    const int r0 = span_begin(c), r1 = span_begin(c + 1);
    if (r1 <= r0) { sk_state = 2; return false; }
    const int u0 = r0 / gper, u1 = (r1 - 1) / gper;
    if (u0 == u1) {
      sk_state = 2;
      describe(u0, r0 - u0 * gper, r1 - u0 * gper, w);
      return true;
    }
    if (sk_state == 0) {  // the START piece of the later unit first
      sk_state = 1;
      describe(u1, 0, r1 - u1 * gper, w);
      return true;
    }
    sk_state = 2;
    describe(u0, r0 - u0 * gper, gper, w);
    return true;
  }
};

// Epilogue warps of the CTA-pair kernel: one pass per work item.  Whole units and finishers drain the accumulator
// through scale/shift/leaky and store (a finisher first adds the partial accumulators of the unit's earlier K spans);
// partial pieces park their raw fp32 accumulator in the workspace and count themselves in.
template <int MODE, bool LEAKY>
__device__ __forceinline__ void epilogue_loop_pair(const ConvKParams& p, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                   uint64_t* tmem_empty_bar, float* s_ss, int cluster_id, int num_clusters,
                                                   int pair_rank, uint8_t* s_out, const CUtensorMap* tm_out) {
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int et = threadIdx.x - 64;   // 0..127
  const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
  const bool vec_ok = ((p.ldc | p.ch_off) & 7) == 0 && (MODE != MC_EPI_REORG2 || (p.N & 7) == 0);
  int as = 0;
  uint32_t aphase = 0;
  int slab_buf = 0;
  float* stat_acc = nullptr;
  int stat_n0 = -1;
  if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0 && p.stat_sum != nullptr) {
    stat_acc = reinterpret_cast<float*>(s_out + (size_t)4 * p.tma_bufs * 4096) + quarter * 512;
    for (int c = lane; c < 512; c += 32) stat_acc[c] = 0.f;
    __syncwarp();
  }
  const bool one_n_tile = p.n_tiles == 1;
  if (one_n_tile) {
    for (int i = et; i < p.block_n; i += 128) {
      const bool ok = i < p.Npad;
      s_ss[i] = ok ? __ldg(p.scale + i) : 0.f;
      s_ss[256 + i] = ok ? __ldg(p.shift + i) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
  PairWorkIter it;
  it.init(p, cluster_id, num_clusters);
  PairWork w;
  int ti = -1;
  while (it.next(w)) {
    ++ti;
    const int tile = w.unit;
    const int n0 = (tile % p.n_tiles) * p.block_n;
    const int m0 = (2 * (tile / p.n_tiles) + pair_rank) * BLOCK_M;
    float* sc = s_ss + (one_n_tile ? 0 : as * 512);
    float* sh = sc + 256;
    if (!one_n_tile) {
      for (int i = et; i < p.block_n; i += 128) {
        const bool ok = (n0 + i) < p.Npad;
        sc[i] = ok ? __ldg(p.scale + n0 + i) : 0.f;
        sh[i] = ok ? __ldg(p.shift + n0 + i) : 0.f;
      }
    }
    const int row = m0 + quarter * 32 + lane;
    const int t = (int)(__umulhi((unsigned int)row, p.wp_mul) >> p.wp_shr);  // row / Wp (magic-number division)
    const int x = row - t * p.Wp;
    const int b = (int)(__umulhi((unsigned int)t, p.hp_mul) >> p.hp_shr);    // t / Hp
    const int y = t - b * p.Hp;
    const bool in_buf = row < p.M_rows;
    const bool interior = in_buf && (x < p.W) && (y < p.H);
    long long out_row_base = 0;
    bool do_store = false;
    if (MODE == MC_EPI_PNHWC) {
      out_row_base = (long long)row * p.ldc + p.ch_off;
      do_store = in_buf;  // pad rows are written as zeros to keep the layout invariant
    } else if (MODE == MC_EPI_REORG2) {
      const int Wo = p.W / 2 + 1, Ho = p.H / 2 + 1;
      const long long orow = ((long long)b * Ho + (y >> 1)) * Wo + (x >> 1);
      out_row_base = orow * p.ldc + p.ch_off + ((y & 1) * 2 + (x & 1)) * p.N;
      do_store = interior;
    } else {  // MC_EPI_NCHW_F32
      out_row_base = ((long long)b * p.N * p.H + y) * p.W + x;  // + n*H*W
      do_store = interior;
    }
    if (!one_n_tile) asm volatile("bar.sync 1, 128;" ::: "memory");  // scale/shift of this tile visible to the 4 warps

    ptx::mbar_wait(&tmem_full_bar[as], aphase);
    ptx::tc_fence_after();
    if (et == 0 && ti < 8) conv_stamp(p, 12 + 2 * ti);
    const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.acc_stride);
    const int u_sk = tile - p.sk_first;  // (stream-K pieces only)
    // partial tiles are stored [slot][column / 4][256 rows][4 floats]: a warp instruction (32 rows, one float4 each) is one
    // contiguous 512-byte segment, for the writers and for the finisher that reads them back with the same mapping
    const int prow = pair_rank * 128 + quarter * 32 + lane;
    const size_t slot_stride = (size_t)256 * p.block_n;
    float* part_unit = p.sk_part + (size_t)(u_sk < 0 ? 0 : u_sk) * 3 * slot_stride;
    auto part_ptr = [&](int slot, int col) -> float* {  // col multiple of 4
      return part_unit + (size_t)slot * slot_stride + ((size_t)(col >> 2) * 256 + prow) * 4;
    };
    if (w.kind == 1) {
      // partial piece: raw accumulator -> workspace slot, then count this warp in (release)
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r0[16], r1[16];
        const bool two = c0 + 16 < p.block_n;
        ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)c0, r0);
        if (two) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(c0 + 16), r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(part_ptr(w.aux, c0 + 4 * q)) = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
        if (two) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(part_ptr(w.aux, c0 + 16 + 4 * q)) = make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]);
        }
      }
      ptx::tc_fence_before();
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        atomicAdd(p.sk_count + u_sk, 1u);
        ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
      }
    } else {
      const int nparts = w.kind == 2 ? w.aux : 0;
      if (nparts > 0) {
        // the earlier K spans of this unit were the first thing their clusters computed: normally long finished
        if (lane == 0) {
          const unsigned int want = 8u * (unsigned int)nparts;
          unsigned int spins = 0;
          while (*reinterpret_cast<volatile unsigned int*>(p.sk_count + u_sk) < want) {
            if (++spins > (1u << 24)) __trap();
          }
          // every writer has arrived and all 8 finisher warps must see that: the LAST finisher warp to get here puts the
          // counter back to zero (second counter = finisher warps done), so the workspace needs no per-launch memset
          if (atomicAdd(p.sk_count + 256 + u_sk, 1u) == 7u) {
            p.sk_count[256 + u_sk] = 0u;
            p.sk_count[u_sk] = 0u;
          }
        }
        __syncwarp();
        __threadfence();
      }
      if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0) {
        if (stat_acc != nullptr && n0 != stat_n0) {
          if (stat_n0 >= 0) flush_stats(p, stat_acc, stat_n0, lane);
          stat_n0 = n0;
        }
        epilogue_tile_tma<LEAKY, true, true>(p, tm_out, taddr_row, sc, sh, interior, in_buf, m0 + quarter * 32, n0, out_row_base,
                                       vec_ok, s_out + (size_t)quarter * p.tma_bufs * 4096, slab_buf, &tmem_empty_bar[as],
                                       [&](uint32_t (&r)[16], int col) {
                                         for (int s2 = 0; s2 < nparts; ++s2) {
#pragma unroll
                                           for (int q = 0; q < 4; ++q) {
                                             const float4 a = __ldcg(reinterpret_cast<const float4*>(part_ptr(s2, col + 4 * q)));
                                             r[4 * q] = __float_as_uint(__uint_as_float(r[4 * q]) + a.x);
                                             r[4 * q + 1] = __float_as_uint(__uint_as_float(r[4 * q + 1]) + a.y);
                                             r[4 * q + 2] = __float_as_uint(__uint_as_float(r[4 * q + 2]) + a.z);
                                             r[4 * q + 3] = __float_as_uint(__uint_as_float(r[4 * q + 3]) + a.w);
                                           }
                                         }
                                       }, stat_acc);
        if (et == 0 && ti < 8) conv_stamp(p, 13 + 2 * ti);
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
        continue;
      }
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t r0[16], r1[16];
        const bool two = c0 + 16 < p.block_n;
        ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)c0, r0);
        if (two) ptx::tmem_ld_32x32b_x16(taddr_row + (uint32_t)(c0 + 16), r1);
        ptx::tmem_ld_wait();
        if (!do_store) continue;
        for (int s2 = 0; s2 < nparts; ++s2) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(part_ptr(s2, c0 + 4 * q)));
            r0[4 * q] = __float_as_uint(__uint_as_float(r0[4 * q]) + a.x);
            r0[4 * q + 1] = __float_as_uint(__uint_as_float(r0[4 * q + 1]) + a.y);
            r0[4 * q + 2] = __float_as_uint(__uint_as_float(r0[4 * q + 2]) + a.z);
            r0[4 * q + 3] = __float_as_uint(__uint_as_float(r0[4 * q + 3]) + a.w);
          }
          if (two) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = __ldcg(reinterpret_cast<const float4*>(part_ptr(s2, c0 + 16 + 4 * q)));
              r1[4 * q] = __float_as_uint(__uint_as_float(r1[4 * q]) + a.x);
              r1[4 * q + 1] = __float_as_uint(__uint_as_float(r1[4 * q + 1]) + a.y);
              r1[4 * q + 2] = __float_as_uint(__uint_as_float(r1[4 * q + 2]) + a.z);
              r1[4 * q + 3] = __float_as_uint(__uint_as_float(r1[4 * q + 3]) + a.w);
            }
          }
        }
        float v[16];
        scale_act16<LEAKY>(r0, sc + c0, sh + c0, interior, v);
        store16<MODE>(p, v, n0 + c0, out_row_base, vec_ok);
        if (two) {
          scale_act16<LEAKY>(r1, sc + c0 + 16, sh + c0 + 16, interior, v);
          store16<MODE>(p, v, n0 + c0 + 16, out_row_base, vec_ok);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_leader(&tmem_empty_bar[as]);
    }
    if (et == 0 && ti < 8) conv_stamp(p, 13 + 2 * ti);
    if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
  }
  if (stat_acc != nullptr && stat_n0 >= 0) flush_stats(p, stat_acc, stat_n0, lane);
  if (MODE == MC_EPI_PNHWC && p.tma_bufs > 0 && lane == 0) ptx::bulk_wait_group_read<0>();  // slabs read before the CTA exits
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant for the wide 3x3 layers (tcgen05 cta_group::2): two CTAs on the two SMs of a TPC compute a
// 256 x block_n tile together.  Each CTA loads its own 128-row activation boxes but only HALF of every weight tile
// (block_n/2 rows); the tensor cores of both SMs read both halves.  Why: the single-CTA kernel is bound by the
// chip-wide L2 -> shared-memory TMA rate (measured ~50 B/clk/SM; a 128x256 tile needs 5.7 KB of A + 32 KB of B per
// 512 MMA cycles = 74 B/clk) — halving B per SM (5.7 + 16 KB -> 42 B/clk) puts the layer back under the tensor pipe.
// Barriers: TMA of both CTAs counts bytes on the LEADER's full barriers (only the leader issues MMAs); MMA completion
// is multicast to the empty / accumulator-full barriers of both CTAs; the epilogues of both CTAs hand the
// accumulator stage back on the leader's barrier (8 arrivals).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_abox,
                              const __grid_constant__ CUtensorMap tmap_out, const ConvKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t b_half_bytes = (uint32_t)(p.block_n / 2 * 128);
  uint8_t* smem = smem_raw;
  {
    uint32_t a = ptx::smem_u32(smem);
    smem += (1024u - (a & 1023u)) & 1023u;
  }
  uint8_t* tiles = smem;
  uint8_t* b_ring = tiles + (size_t)p.a_stages * A_BOX_STRIDE;
  uint8_t* aux = b_ring + (size_t)p.stages * b_half_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;   // [MAX_ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + MAX_ACC; // [MAX_ACC]
  uint64_t* afull_bar = tmem_empty_bar + MAX_ACC;     // [MAX_A_STAGES]
  uint64_t* aempty_bar = afull_bar + MAX_A_STAGES;    // [MAX_A_STAGES]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aempty_bar + MAX_A_STAGES);
  float* s_ss = reinterpret_cast<float*>(aux + 512);
  uint8_t* s_out = aux + 5120;  // TMA-store slabs [4 warps][tma_bufs][4 KB], 1024-byte aligned

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  if (threadIdx.x == 0) conv_stamp(p, 0);

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_b);
    ptx::prefetch_tensormap(&tmap_abox);
    if (p.tma_bufs > 0) ptx::prefetch_tensormap(&tmap_out);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < p.acc_stages; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], 8);  // 4 epilogue warps of each CTA
    }
    for (int a = 0; a < MAX_A_STAGES; ++a) {
      ptx::mbar_init(&afull_bar[a], 1);
      ptx::mbar_init(&aempty_bar[a], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 1) {
    ptx::tmem_alloc_pair(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // barriers of BOTH CTAs initialised before any remote arrival / complete_tx
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) conv_stamp(p, 1);

  if (warp_idx == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (lane == 0) {
      int s = 0, sa = 0;
      uint32_t phase = 0, pha = 0;
      PairWorkIter it;
      it.init(p, cluster_id, num_clusters);
      PairWork w;
      while (it.next(w)) {
        const int unit = w.unit;
        const int n0 = (unit % p.n_tiles) * p.block_n + rank * (p.block_n / 2);
        const int m0 = (2 * (unit / p.n_tiles) + rank) * BLOCK_M;
        const int gb = w.g0, ge = w.g1 < 0 ? 3 * p.kb_per_tap : w.g1;
        for (int g = gb; g < ge; ++g) {
          const int dy = g / p.kb_per_tap, cb = g - dy * p.kb_per_tap;
          ptx::mbar_wait(&aempty_bar[sa], pha ^ 1u);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&afull_bar[sa], 2 * A_BOX_BYTES);
          ptx::tma_load_2d_pair(tiles + (size_t)sa * A_BOX_STRIDE, &tmap_abox, &afull_bar[sa], cb * 64,
                                m0 + (dy - 1) * p.Wp - 1);
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
          for (int dx = 0; dx < 3; ++dx) {
            const int kb = (dy * 3 + dx) * p.kb_per_tap + cb;
            ptx::mbar_wait(&empty_bar[s], phase ^ 1u);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * b_half_bytes);
            ptx::tma_load_2d_pair(b_ring + (size_t)s * b_half_bytes, &tmap_b, &full_bar[s], kb * 64, n0);
            if (++s == p.stages) { s = 0; phase ^= 1u; }
          }
        }
      }
      conv_stamp(p, 3);
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (warp 1 of the leader CTA; one elected lane issues) =====================
    if (rank == 0) {
      int s = 0, sa = 0, as = 0;
      uint32_t phase = 0, pha = 0, aphase = 0;
      PairWorkIter it;
      it.init(p, cluster_id, num_clusters);
      PairWork w;
      int ti = -1;
      while (it.next(w)) {
        ++ti;
        ptx::mbar_wait(&tmem_empty_bar[as], aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * p.acc_stride);
        const int gb = w.g0, ge = w.g1 < 0 ? 3 * p.kb_per_tap : w.g1;
        for (int g = gb; g < ge; ++g) {
          const int nk = ((g % p.kb_per_tap) == p.kb_per_tap - 1) ? p.last_ksteps : 4;
          ptx::mbar_wait(&afull_bar[sa], pha);
          if (ti == 0 && g == gb && lane == 0) conv_stamp(p, 4);
          const uint32_t a_addr = ptx::smem_u32(tiles + (size_t)sa * A_BOX_STRIDE);
          for (int dx = 0; dx < 3; ++dx) {
            ptx::mbar_wait(&full_bar[s], phase);
            ptx::tc_fence_after();
            const uint64_t adesc = ptx::make_sw128_kmajor_desc(a_addr + (uint32_t)dx * 128u);
            const uint64_t bdesc = ptx::make_sw128_kmajor_desc(ptx::smem_u32(b_ring + (size_t)s * b_half_bytes));
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk)
                  ptx::umma_bf16_ss_pair(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                                         (g > gb || dx > 0 || k > 0) ? 1u : 0u);
              ptx::umma_commit_pair(&empty_bar[s]);
              if (dx == 2) ptx::umma_commit_pair(&aempty_bar[sa]);
              if (dx == 2 && g == ge - 1) ptx::umma_commit_pair(&tmem_full_bar[as]);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; phase ^= 1u; }
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1u; }
        }
        if (lane == 0 && ti < 6) conv_stamp(p, 5 + ti);
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..5 of both CTAs =====================
    if (p.epi_mode == MC_EPI_PNHWC) {
      if (p.leaky) epilogue_loop_pair<MC_EPI_PNHWC, true>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
      else epilogue_loop_pair<MC_EPI_PNHWC, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
    } else if (p.epi_mode == MC_EPI_REORG2) {
      if (p.leaky) epilogue_loop_pair<MC_EPI_REORG2, true>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
      else epilogue_loop_pair<MC_EPI_REORG2, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
    } else {
      if (p.leaky) epilogue_loop_pair<MC_EPI_NCHW_F32, true>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
      else epilogue_loop_pair<MC_EPI_NCHW_F32, false>(p, tmem_base, tmem_full_bar, tmem_empty_bar, s_ss, cluster_id, num_clusters, rank, s_out, &tmap_out);
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // the peer may still read this CTA's smem / signal its barriers until both are done
  if (warp_idx == 1) ptx::tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
}

// Tile-width heuristic.  One 64-wide k-block of a 128 x bn tile costs max(2*bn, 128+bn) SM cycles: 2*bn is the
// tcgen05 issue floor (128*bn*64 MACs at 4096 MAC/clk), 128+bn is the shared-memory read of the A (16 KB) and
// B (bn*128 B) tiles at 128 B/clk.  Cost = waves over the SMs x (k-blocks x that + a fixed prologue/epilogue).
int pick_block_n(int Npad, int m_tiles, int num_kb, int num_sms, bool share_dx) {
  int best = 16;
  double best_cost = 1e30;
  for (int bn = 16; bn <= 256; bn += 16) {  // every legal UMMA N for M=128
    const int n_tiles = (Npad + bn - 1) / bn;
    const long long tiles = (long long)n_tiles * m_tiles;
    const long long rounds = (tiles + num_sms - 1) / num_sms;  // persistent CTAs: tiles per CTA
    // per 64-wide k-block: tcgen05 issue floor 2*bn cycles; smem operand read (16 KB + bn*128 B at 128 B/clk);
    // operand FEED from L2 through TMA, measured on B200 at ~13.3 cycles per KB per SM when all SMs stream
    // (128x256 tiles run at ~80% tensor-active = 640 cycles per k-block for 48 KB)
    double per_kb = 2.0 * bn;
    if (128.0 + bn > per_kb) per_kb = 128.0 + bn;
    // (with the shared activation box a tap's k-block brings 17/3 KB of A instead of 16 KB)
    const double feed = 13.3 * ((share_dx ? 17.0 / 3.0 : 16.0) + bn / 8.0);
    if (feed > per_kb) per_kb = feed;
    // one epilogue (~12 cycles per column) is exposed at the end; the others overlap the next tile's MMAs unless
    // they are longer than the mainloop
    const double main = per_kb * num_kb, epi = 12.0 * bn + 400.0;
    const double cost = (double)rounds * (main > epi ? main : epi) + epi + 2500.0;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

static thread_local int g_last_plan[8] = {0, 0, 0, 0, 0, 0, 0, 0};
static unsigned long long* g_conv_dbg = nullptr;

// Debug: timeline stamps of the single-CTA conv GEMM kernel (32 x uint64 per CTA, globaltimer ns) into d_buf, which
// must hold 32 * 8 * grid bytes; NULL switches tracing off again.  tools/trace_conv.py prints the timelines.
extern "C" int mc_debug_conv_trace(void* d_buf) {
  g_conv_dbg = reinterpret_cast<unsigned long long*>(d_buf);
  return 0;
}

extern "C" int mc_conv_last_plan(int info[8]) {
  MC_CHECK_ARG(info != nullptr, "mc_conv_last_plan: null pointer");
  for (int i = 0; i < 8; ++i) info[i] = g_last_plan[i];
  return 0;
}

// Workspace the stream-K tail of the CTA-pair kernel can use for this layer (0: the layer never takes that path).  Upper
// bound over every tile width the planner may pick: up to 148/2 stream-K units x 3 partial slots x 256 rows x 256 columns
// of fp32, plus 2 KB of counters.  The counters must be ZERO before the first launch; every launch leaves them zero.
extern "C" size_t mc_workspace_bytes_conv_fwd(const mc_conv_desc* d) {
  if (d == nullptr || d->ksize != 3 || d->Cin <= 32 || d->Npad < 192) return 0;
  const int clusters = mc_num_sms() / 2;
  return (size_t)2048 + (size_t)clusters * 3 * 256 * 256 * sizeof(float);
}

extern "C" int mc_conv_fwd(const mc_conv_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d != nullptr, "mc_conv_fwd: null descriptor");
  MC_CHECK_ARG(d->d_in && d->d_wpack && d->d_scale && d->d_shift && (d->d_out || d->epi_mode == MC_EPI_DECODE),
               "mc_conv_fwd: null pointer");
  MC_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "mc_conv_fwd: ksize must be 1 or 3 (got %d)", d->ksize);
  MC_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->N > 0, "mc_conv_fwd: bad dims");
  MC_CHECK_ARG((d->Cin_ld % 8) == 0 && d->Cin <= d->Cin_ld, "mc_conv_fwd: Cin_ld must be a multiple of 8 >= Cin");
  MC_CHECK_ARG((d->Npad % 16) == 0 && d->Npad >= d->N, "mc_conv_fwd: Npad must be a multiple of 16 >= N");
  MC_CHECK_ARG(((uintptr_t)d->d_in & 15) == 0 && ((uintptr_t)d->d_wpack & 15) == 0 && ((uintptr_t)d->d_out & 15) == 0,
               "mc_conv_fwd: pointers must be 16-byte aligned");
  MC_CHECK_ARG(d->epi_mode == MC_EPI_PNHWC || d->epi_mode == MC_EPI_REORG2 || d->epi_mode == MC_EPI_NCHW_F32 ||
                   d->epi_mode == MC_EPI_DECODE,
               "mc_conv_fwd: unsupported epilogue mode %d", d->epi_mode);
  const bool decode = d->epi_mode == MC_EPI_DECODE;
  if (decode) {
    const mc_decode_params* q = d->decode;
    MC_CHECK_ARG(q != nullptr && q->d_boxes != nullptr, "mc_conv_fwd: MC_EPI_DECODE needs decode parameters with d_boxes");
    MC_CHECK_ARG(q->A > 0 && q->A <= MC_MAX_ANCHORS && q->nc > 0 && q->A * (5 + q->nc) == d->N,
                 "mc_conv_fwd: decode needs N == A*(5+nc) (N %d, A %d, nc %d)", d->N, q->A, q->nc);
    MC_CHECK_ARG(d->Npad <= 256 && d->leaky == 0, "mc_conv_fwd: decode needs a linear head of <= 256 channels");
    MC_CHECK_ARG(((uintptr_t)q->d_boxes & 15) == 0, "mc_conv_fwd: d_boxes must be 16-byte aligned");
  }
  if (d->epi_mode == MC_EPI_REORG2) MC_CHECK_ARG((d->H % 2) == 0 && (d->W % 2) == 0, "mc_conv_fwd: reorg needs even H,W");

  const int ntaps = d->ksize * d->ksize;
  const int BLOCK_K = d->block_k == 32 ? 32 : 64;
  MC_CHECK_ARG(d->block_k == 0 || d->block_k == 32 || d->block_k == 64, "mc_conv_fwd: block_k must be 0, 32 or 64");
  const int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
  const int Kc = ((d->Cin + BLOCK_K - 1) / BLOCK_K) * BLOCK_K;
  const long long M_rows = (long long)d->B * (d->H + 1) * (d->W + 1);
  MC_CHECK_ARG(M_rows < (1ll << 31), "mc_conv_fwd: too many rows");
  const int m_tiles = (int)((M_rows + BLOCK_M - 1) / BLOCK_M);

  // 3x3 with 64-wide k-blocks: share one activation box among the three dx taps (MCB200_SHARE_DX=0 disables)
  static int share_env = -1, share_env32 = 1;
  if (share_env < 0) {
    const char* e = mc_tune_env("MCB200_SHARE_DX");
    share_env = (e && e[0] == '0') ? 0 : 1;
    share_env32 = (e && e[0] == '2') ? 0 : 1;
  }
  // (MCB200_SHARE_DX=2 restricts the shared box to 64-wide k-blocks, the state before the 64-byte-swizzle variant)
  // 32-wide k-blocks share the box only on long, feed-bound launches (dense conv2: 21,841 tiles, 306 -> 262 us); short
  // ones lose a little to the two-ring pipeline (shrunk conv13, 365 tiles: 18 -> 21 us)
  const int share_dx = (d->ksize == 3 && share_env && (BLOCK_K == 64 || (share_env32 && m_tiles >= 2048))) ? 1 : 0;
  const int a_box_stride = BLOCK_K == 64 ? A_BOX_STRIDE : A_BOX_STRIDE / 2;
  int block_n = d->block_n;
  // (the operand-feed term keeps the un-shared 16 KB A tile: with the shared-box figure the model picks narrower tiles
  //  that measured 4 % slower end to end on B200)
  {  // experiment switch: MCB200_CONV_BN_BIG=<bn> forces the tile width of the wide 3x3 layers (Npad >= 960)
    static int bn_big = -1;
    if (bn_big < 0) {
      const char* e = mc_tune_env("MCB200_CONV_BN_BIG");
      bn_big = e ? atoi(e) : 0;
    }
    if (block_n <= 0 && bn_big >= 16 && bn_big <= 256 && (bn_big % 16) == 0 && d->Npad >= 960 && d->ksize == 3) block_n = bn_big;
    static int bn_mid = -1;  // MCB200_CONV_BN_MID=<bn>: same for 3x3 layers with 512 <= Npad < 960
    if (bn_mid < 0) {
      const char* e = mc_tune_env("MCB200_CONV_BN_MID");
      bn_mid = e ? atoi(e) : 0;
    }
    if (block_n <= 0 && bn_mid >= 16 && bn_mid <= 256 && (bn_mid % 16) == 0 && d->Npad >= 512 && d->Npad < 960 && d->ksize == 3)
      block_n = bn_mid;
  }
  if (decode) block_n = d->Npad;  // every logit of a cell in ONE accumulator row
  if (block_n <= 0) block_n = pick_block_n(d->Npad, m_tiles, (ntaps * Kc + 63) / 64, mc_num_sms(), false);  // 64-wide k-block units
  const bool want_stats = d->d_stat_sum != nullptr || d->d_stat_sumsq != nullptr;
  if (want_stats) {
    // batch statistics come from the TMA-store slabs: every column of a tile must leave in a whole 64-column chunk
    MC_CHECK_ARG(d->d_stat_sum != nullptr && d->d_stat_sumsq != nullptr, "mc_conv_fwd: d_stat_sum and d_stat_sumsq go together");
    MC_CHECK_ARG(d->epi_mode == MC_EPI_PNHWC && ((d->ldc | d->ch_off) & 7) == 0 && d->N >= 33,
                 "mc_conv_fwd: statistics need the PNHWC epilogue, 8-aligned pitch / offset and more than 32 channels");
    block_n = (block_n + 63) / 64 * 64;
    if (block_n > 256) block_n = 256;
  }
  MC_CHECK_ARG(block_n >= 16 && block_n <= 256 && (block_n % 16) == 0, "mc_conv_fwd: block_n %d invalid", block_n);
  const int n_tiles = (d->Npad + block_n - 1) / block_n;

  const int stage_bytes = A_TILE_BYTES + block_n * BLOCK_K * 2;
  const long long total_tiles = (long long)n_tiles * m_tiles;
  const int acc_stride = block_n == 16 ? 16 : ((block_n + 31) / 32) * 32;
  int tmem_cols = 32;
  while (tmem_cols < 2 * acc_stride) tmem_cols <<= 1;  // at least two accumulator stages (more below for narrow layers)
  const int num_kb = ntaps * (Kc / BLOCK_K);
  constexpr size_t AUX_BYTES = 5120 + 1024;  // barriers (512), scale/shift staging (4096), pad to 1 KB; 1024-byte alignment slack

  // smem ring of one CTA when `ctas` CTAs share an SM (each CTA also costs 1 KB of reserved shared memory)
  // resident weights: one N tile and all taps' weight tiles fit beside the activation ring -> loaded once per CTA
  // (narrow 3x3 layers: the per-tile weight re-load is otherwise as much TMA traffic as the activations)
  const size_t b_res_bytes = (size_t)num_kb * block_n * BLOCK_K * 2;
  // OFF by default (MCB200_CONV_RESIDENT=1 enables): measured on B200 it loses to the weight ring on every narrow
  // 3x3 shape tried (64->64 @52x52: 37 us vs 25 us; 48->32 @104x104: 71 vs 58; 40->80 @52x52: 41 vs 33) because the
  // resident tiles leave room for ONE CTA per SM, and three small CTAs per SM hide the per-tile latency chain better
  // than the saved weight traffic.
  static int res_env = -1;
  if (res_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_RESIDENT");
    res_env = (e && e[0] == '1') ? 1 : 0;
  }
  const int b_resident = (res_env && share_dx && n_tiles == 1 && d->stages <= 0 && b_res_bytes <= 112 * 1024) ? 1 : 0;
  // staged epilogue (PNHWC, tiles up to 128 wide): 128 rows x (block_n bf16 + 16 B) of shared memory
  static int stage_env = -1;  // MCB200_CONV_STAGE_OUT=0 disables (A/B switch)
  if (stage_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_STAGE_OUT");
    stage_env = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1);
  }
  // (measured: tiles 32..96 wide gain — dense conv2 396 -> 300 us, conv4' 60 -> 54, shrunk conv8 33 -> 29; 128-wide tiles
  //  lose smem ring depth to the 34 KB staging buffer (133 -> 154 us) and 16-wide rows are two stores either way)
  // Wider tiles go through the buffer in 64-column chunks where it was measured to pay: 3x3 layers at high resolution
  // (dense conv3/conv5, 128 wide, 5513 tiles: 137 -> 123 us); wide 1x1 layers (conv7: 39 -> 47 us) and the short
  // launches of the 26x26 / 13x13 stages do not.  MCB200_CONV_STAGE_OUT=2 stages every wide tile (A/B switch).
  const bool wide_ok = stage_env == 2 || (d->ksize == 3 && m_tiles >= 1024);
  // TMA-store epilogue (tiles >= 64 wide): whole 64-column chunks leave as bulk tensor stores from 4 KB slabs per
  // epilogue warp — two slabs per warp when the CTA has the SM to itself, one (16 KB per CTA) when several CTAs share
  // it (measured: dense conv4/conv6 at 104x104 run 113 us with 2 CTAs per SM + one slab, 158 us with one CTA + two
  // slabs, 125 us with the staged epilogue).  MCB200_CONV_TMA_OUT=0 disables, =1 / =2 force the slab count.
  static int tma_env = -1;
  if (tma_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_TMA_OUT");
    tma_env = e ? atoi(e) : 3;
    if (tma_env < 0 || tma_env > 3) tma_env = 3;
  }
  const bool tma_out = tma_env && d->epi_mode == MC_EPI_PNHWC && block_n >= 64 && ((d->ldc | d->ch_off) & 7) == 0;
  auto tma_bufs_for = [&](int ctas_) -> int { return !tma_out ? 0 : (tma_env == 3 ? (ctas_ == 1 ? 2 : 1) : tma_env); };
  const bool stage_out = !tma_out && stage_env && d->epi_mode == MC_EPI_PNHWC && block_n >= 32 && (block_n <= 96 || wide_ok) &&
                         ((d->ldc | d->ch_off) & 7) == 0;
  const int out_chunk = block_n <= 96 ? block_n : 64;
  const int out_pitch = stage_out ? out_chunk * 2 + 16 : 0;
  // (decode epilogue: 128 rows x (block_n + 1) floats of logits)
  auto out_stage_for = [&](int ctas_) -> size_t {
    return decode ? (size_t)128 * (block_n + 1) * 4
                  : (tma_out ? (size_t)tma_bufs_for(ctas_) * 4 * 4096 + (want_stats ? 4 * 512 * sizeof(float) : 0)
                             : (size_t)128 * out_pitch);
  };
  auto plan_ring = [&](int ctas, int* st, int* ast, size_t* bytes) -> bool {
    const size_t out_stage_bytes = out_stage_for(ctas);
    const long long cap = (ctas == 1 ? 220 * 1024 - (long long)AUX_BYTES : (227 * 1024) / ctas - 1024 - (long long)AUX_BYTES) - (long long)out_stage_bytes;
    if (b_resident) {
      // the activation ring is the only pipeline left: as deep as the smem beside the weights allows (a tile is 3 boxes)
      long long a = (cap - (long long)b_res_bytes) / a_box_stride;
      if (a > MAX_A_STAGES) a = MAX_A_STAGES;
      if (a < (ctas == 1 ? 3 : 2)) return false;
      *st = 1; *ast = (int)a;
      *bytes = (size_t)a * a_box_stride + b_res_bytes + AUX_BYTES + out_stage_bytes;
      return true;
    }
    if (share_dx) {
      const int b_bytes = block_n * BLOCK_K * 2;
      int a = ctas == 1 ? DEF_A_STAGES : 2;
      long long stg = (cap - (long long)a * a_box_stride) / b_bytes;
      if (stg > MAX_STAGES) stg = MAX_STAGES;
      if (stg < 3) return false;
      *st = (int)stg; *ast = a;
      *bytes = (size_t)a * a_box_stride + (size_t)stg * b_bytes + AUX_BYTES + out_stage_bytes;
    } else {
      long long stg = cap / stage_bytes;
      if (stg > MAX_STAGES) stg = MAX_STAGES;
      if (stg > num_kb * 4) stg = num_kb * 4;  // a ring deeper than four tiles' worth of k-blocks buys nothing
      if (stg < (ctas == 1 ? 1 : 3)) return false;
      *st = (int)stg; *ast = 0;
      *bytes = (size_t)stg * stage_bytes + AUX_BYTES + out_stage_bytes;
    }
    return true;
  };
  // CTAs per SM.  Wide, long-K tiles are tensor-bound: one CTA with the deepest ring.  Narrow layers are bound by the
  // per-tile latency chain, which only more CTAs (or tiles) in flight hide: up to 3 when the rings and the TMEM
  // slices fit and every CTA still gets >= 2 tiles.  MCB200_CONV_CTAS=1|2|3 forces the upper limit.
  static int ctas_env = -1;
  if (ctas_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_CTAS");
    ctas_env = e ? atoi(e) : 0;
    if (ctas_env < 0 || ctas_env > 3) ctas_env = 0;
  }
  static int min_tiles = -1;  // tiles every CTA must get before another CTA per SM is added (MCB200_CONV_MINTILES, A/B)
  if (min_tiles < 0) {
    const char* e = mc_tune_env("MCB200_CONV_MINTILES");
    min_tiles = e ? atoi(e) : 1;  // measured: 1 beats 2 by 0.5 % on the shrunk step (the 365-tile 26x26 layers get 2 CTAs per SM)
    if (min_tiles < 1) min_tiles = 1;
  }
  int ctas = 1;
  {
    const int limit = ctas_env ? ctas_env : 3;
    const double main_cycles = (double)num_kb * (BLOCK_K / 64.0) * (2.0 * block_n > 128.0 + block_n ? 2.0 * block_n : 128.0 + block_n);
    const bool narrow = ctas_env ? true : main_cycles < 6000.0;
    for (int c = limit; c >= 2 && narrow; --c) {
      int st, ast;
      size_t bytes;
      if (c * tmem_cols > 512 || total_tiles < (long long)min_tiles * c * mc_num_sms() || !plan_ring(c, &st, &ast, &bytes)) continue;
      ctas = c;
      break;
    }
  }
  int stages = d->stages;
  int a_stages = 0;
  size_t smem_bytes;
  if (stages <= 0) {
    MC_CHECK_ARG(plan_ring(ctas, &stages, &a_stages, &smem_bytes), "mc_conv_fwd: no shared-memory ring fits block_n %d", block_n);
  } else {
    ctas = 1;
    if (share_dx) {
      a_stages = DEF_A_STAGES;
      MC_CHECK_ARG(stages >= 2 && stages <= MAX_STAGES, "mc_conv_fwd: stages %d invalid", stages);
      smem_bytes = (size_t)a_stages * a_box_stride + (size_t)stages * block_n * BLOCK_K * 2 + AUX_BYTES + out_stage_for(1);
    } else {
      MC_CHECK_ARG(stages >= 1 && stages <= MAX_STAGES, "mc_conv_fwd: stages %d invalid", stages);
      smem_bytes = (size_t)stages * stage_bytes + AUX_BYTES + out_stage_for(1);
    }
  }
  MC_CHECK_ARG(smem_bytes <= 227 * 1024, "mc_conv_fwd: smem %zu too large", smem_bytes);

  // CTA pair (cta_group::2) for the wide 3x3 layers: MCB200_CONV_2CTA=0 off, 1 auto (default), 2 wherever legal
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_2CTA");
    pair_env = e ? atoi(e) : 1;
    if (pair_env < 0 || pair_env > 2) pair_env = 1;
  }
  const bool pair_legal = share_dx && BLOCK_K == 64 && d->stages <= 0 && (block_n % 16) == 0 && block_n >= 64 && m_tiles >= 2;
  const bool use_pair = !decode && pair_legal && (pair_env == 2 || (pair_env == 1 && block_n >= 192 && num_kb >= 36));
  if (use_pair) {
    ctas = 1;
    a_stages = DEF_A_STAGES;
    const int b_half = block_n / 2 * 128;
    const int slab_bytes = tma_bufs_for(1) * 4 * 4096 + (want_stats ? 4 * 512 * (int)sizeof(float) : 0);
    stages = (220 * 1024 - a_stages * A_BOX_STRIDE - slab_bytes) / b_half;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    smem_bytes = (size_t)a_stages * A_BOX_STRIDE + (size_t)stages * b_half + AUX_BYTES + slab_bytes;
  }

  CUtensorMap tm_a, tm_b, tm_abox;
  MC_CHECK_ARG(d->in_cols == 0 || (d->in_cols >= d->Cin && d->in_cols <= d->Cin_ld), "mc_conv_fwd: in_cols %d outside [Cin, Cin_ld]", d->in_cols);
  const uint64_t in_cols = d->in_cols ? (uint64_t)d->in_cols : (uint64_t)d->Cin;
  int rc = mc_make_tmap_2d_bf16_k(&tm_a, d->d_in, (uint64_t)M_rows, in_cols, (uint64_t)d->Cin_ld, BLOCK_M, BLOCK_K);
  if (rc) return rc;
  tm_abox = tm_a;
  if (share_dx) {
    rc = mc_make_tmap_2d_bf16_k(&tm_abox, d->d_in, (uint64_t)M_rows, in_cols, (uint64_t)d->Cin_ld, A_BOX_ROWS, BLOCK_K);
    if (rc) return rc;
  }
  // weights: [n_tiles*block_n >= Npad rows (OOB rows zero-filled), ntaps*Kc]
  rc = mc_make_tmap_2d_bf16_k(&tm_b, d->d_wpack, (uint64_t)d->Npad, (uint64_t)ntaps * Kc, (uint64_t)ntaps * Kc,
                              (uint32_t)(use_pair ? block_n / 2 : block_n), BLOCK_K);
  if (rc) return rc;

  CUtensorMap tm_out = tm_a;
  const int tma_bufs = tma_bufs_for(use_pair ? 1 : ctas);
  if (tma_bufs) {
    // output slice [ch_off, ch_off + N rounded up to 8) of the PNHWC rows: stores are clipped there, so a chunk that
    // reaches past the layer's channels never touches the neighbouring concat slice; channels N .. 8-multiple get zeros
    uint64_t ocols = (uint64_t)((d->N + 7) / 8 * 8);
    if (ocols > (uint64_t)(d->ldc - d->ch_off)) ocols = (uint64_t)(d->ldc - d->ch_off);
    rc = mc_make_tmap_2d_bf16_k(&tm_out, reinterpret_cast<const __nv_bfloat16*>(d->d_out) + d->ch_off, (uint64_t)M_rows, ocols,
                                (uint64_t)d->ldc, 32, 64);
    if (rc) return rc;
  }

  ConvKParams p;
  p.M_rows = (int)M_rows;
  p.N = d->N;
  p.Npad = d->Npad;
  p.kb_per_tap = Kc / BLOCK_K;
  p.block_k = BLOCK_K;
  p.num_kb = ntaps * p.kb_per_tap;
  p.ksize = d->ksize;
  p.W = d->W;
  p.H = d->H;
  p.Wp = d->W + 1;
  p.Hp = d->H + 1;
  p.block_n = block_n;
  p.stages = stages;
  p.share_dx = share_dx;
  p.b_resident = (b_resident && !use_pair) ? 1 : 0;
  {
    const int rem = d->Cin - (Kc / BLOCK_K - 1) * BLOCK_K;  // channels in the last channel block of a tap
    p.last_ksteps = (rem + 15) / 16;
  }
  p.a_stages = a_stages;
  // Accumulator stages: two.  More (a narrow one-N-tile layer's tiles are small, the CTA's TMEM share holds up to 8) was
  // measured on B200 and rejected: no gain on the narrow layers themselves (their epilogue warps, not the hand-back
  // round trip, set the ~1 us per tile and CTA), and the larger TMEM allocations keep the next kernel's CTAs from
  // co-residing with this kernel's tail (dense step 2.25 -> 2.35 ms).  MCB200_CONV_ACC=<n> (3..8) enables it for A/B.
  static int acc_env = -1;
  if (acc_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_ACC");
    acc_env = e ? atoi(e) : 0;
  }
  int acc_stages = 2;
  if (n_tiles == 1 && !use_pair && acc_env > 2) {
    int budget = ctas == 1 ? 512 : (ctas == 2 ? 256 : 128);
    acc_stages = budget / acc_stride;
    if (acc_stages > MAX_ACC) acc_stages = MAX_ACC;
    if (acc_stages > acc_env) acc_stages = acc_env;
    if (acc_stages < 2) acc_stages = 2;
    tmem_cols = 32;
    while (tmem_cols < acc_stages * acc_stride) tmem_cols <<= 1;
  }
  p.acc_stages = acc_stages;
  p.out_pitch = use_pair ? 0 : out_pitch;
  p.out_chunk = out_chunk;
  p.tma_bufs = tma_bufs;
  p.stat_sum = d->d_stat_sum;
  p.stat_sumsq = d->d_stat_sumsq;
  MC_CHECK_ARG(!want_stats || tma_bufs > 0, "mc_conv_fwd: statistics need the TMA-store epilogue (tuning switch disabled it)");
  {
    auto magic = [](unsigned int dv, unsigned int* mul, unsigned int* shr) {
      unsigned int l = 0;
      while ((1u << l) < dv) ++l;
      const unsigned long long pw = 31ull + l;
      *mul = (unsigned int)((((unsigned long long)1 << pw) + dv - 1) / dv);
      *shr = (unsigned int)(pw - 32);
    };
    magic((unsigned int)(d->W + 1), &p.wp_mul, &p.wp_shr);
    magic((unsigned int)(d->H + 1), &p.hp_mul, &p.hp_shr);
  }
  p.acc_stride = acc_stride;
  p.tmem_cols = tmem_cols;
  p.m_tiles = m_tiles;
  p.n_tiles = n_tiles;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  p.scale = d->d_scale;
  p.shift = d->d_shift;
  p.out = d->d_out;
  p.ldc = d->ldc;
  p.ch_off = d->ch_off;
  p.epi_mode = d->epi_mode;
  p.leaky = d->leaky;
  p.sk_first = 0; p.sk_units = 0; p.sk_gper = 0; p.sk_total_g = 0; p.sk_clusters = 0; p.sk_count = nullptr; p.sk_part = nullptr;
  p.dbg = g_conv_dbg;
  p.dec_boxes = p.dec_cls = p.dec_head = nullptr;
  p.dec_A = p.dec_nc = p.dec_only_obj = 0;
  p.dec_thresh = 0.f;
  if (decode) {
    const mc_decode_params* q = d->decode;
    p.dec_boxes = q->d_boxes;
    p.dec_cls = q->d_cls;
    p.dec_head = q->d_head;
    p.dec_A = q->A;
    p.dec_nc = q->nc;
    p.dec_only_obj = q->only_objectness;
    p.dec_thresh = q->conf_thresh;
    for (int a = 0; a < MC_MAX_ANCHORS; ++a) {
      p.dec_anc.w[a] = a < q->A ? q->anchors[2 * a] : 0.f;
      p.dec_anc.h[a] = a < q->A ? q->anchors[2 * a + 1] : 0.f;
    }
  }

  if (use_pair) {
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    p.acc_stages = 2;
    p.acc_stride = ((block_n + 31) / 32) * 32;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.acc_stride) p.tmem_cols <<= 1;
    static int max_clusters = -1;
    if (max_clusters < 0) {
      MC_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (mc_num_sms() / 2));
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = 222 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_gemm_tcgen05_pair_kernel, &cfg) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = mc_num_sms() / 2;
      }
      max_clusters = n;
    }
    const int m_pairs = (m_tiles + 1) / 2;
    const long long units = (long long)m_pairs * n_tiles;
    const int clusters = (int)(units < max_clusters ? units : max_clusters);
    // Stream-K for a SHORT partial last wave (at most half a wave of units left over, e.g. batch 32: 100 units over 74
    // clusters — without it the 26 left-over units cost a whole second round).  Measured on B200 for the bench shapes
    // (196 units = 2.65 waves): no gain, 0.8556 vs 0.8584 ms per step — the 48 clusters of a two-thirds-full last wave
    // run their units faster because the operand feed from L2 is shared by fewer SMs, so whole tiles are kept there.
    // Needs the caller's workspace; MCB200_CONV_STREAMK=0 disables, =2 forces it for any remainder (tuning builds).
    p.sk_first = (int)units;
    p.sk_units = 0;
    p.sk_gper = 3 * p.kb_per_tap;
    p.sk_total_g = 0;
    p.sk_clusters = 0;
    p.sk_count = nullptr;
    p.sk_part = nullptr;
    {
      const long long full = units / clusters, rem = units % clusters;
      const char* e = mc_tune_env("MCB200_CONV_STREAMK");
      const bool on = !(e && e[0] == '0');
      const bool any_rem = e && e[0] == '2';
      const size_t need = mc_workspace_bytes_conv_fwd(d);
      if (on && full >= 1 && rem > 0 && (rem * 2 <= (long long)clusters || any_rem) && d->d_ws != nullptr && need > 0 &&
          d->ws_bytes >= need && ((uintptr_t)d->d_ws & 255) == 0) {
        p.sk_first = (int)(full * clusters);
        p.sk_units = (int)rem;
        p.sk_total_g = (int)rem * p.sk_gper;
        p.sk_clusters = (int)(3 * rem < clusters ? 3 * rem : clusters);
        p.sk_count = reinterpret_cast<unsigned int*>(d->d_ws);  // [256] writer arrivals, [256] finisher warps done
        p.sk_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->d_ws) + 2048);
      }
    }
    g_last_plan[0] = 1; g_last_plan[1] = block_n; g_last_plan[2] = 1; g_last_plan[3] = p.sk_units; g_last_plan[4] = 1;
    g_last_plan[5] = stages; g_last_plan[6] = 2 * clusters; g_last_plan[7] = BLOCK_K;
    conv_gemm_tcgen05_pair_kernel<<<2 * clusters, NUM_THREADS, smem_bytes, stream>>>(tm_b, tm_abox, tm_out, p);
    MC_LAUNCH_CHECK("conv_gemm_tcgen05_pair_kernel");
    return 0;
  }

  static bool attr_set = false;
  if (!attr_set) {
    MC_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MC_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    MC_CUDA(cudaFuncSetAttribute(conv_gemm_tcgen05_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set = true;
  }
  const long long max_ctas = (long long)mc_num_sms() * ctas;
  int grid = (int)(total_tiles < max_ctas ? total_tiles : max_ctas);
  static int trace_env = -1;  // MCB200_CONV_TRACE=1: one line per launch on stderr (tools/bench_layers.py)
  if (trace_env < 0) {
    const char* e = mc_tune_env("MCB200_CONV_TRACE");
    trace_env = (e && e[0] == '1') ? 1 : 0;
  }
  if (trace_env)
    fprintf(stderr, "[mc_conv_fwd] %dx%d k%d Cin %d N %d: m_tiles %d block_n %d n_tiles %d kb %d(x%d) ctas %d stages %d/%d share %d stage_out %d grid %d smem %zu\n",
            d->H, d->W, d->ksize, d->Cin, d->N, m_tiles, block_n, n_tiles, num_kb, BLOCK_K, ctas, stages, a_stages, share_dx,
            p.out_pitch, grid, smem_bytes);
  g_last_plan[0] = 0; g_last_plan[1] = block_n; g_last_plan[2] = ctas; g_last_plan[3] = p.b_resident;
  g_last_plan[4] = share_dx; g_last_plan[5] = stages; g_last_plan[6] = grid; g_last_plan[7] = BLOCK_K;
  if (ctas >= 3)
    conv_gemm_tcgen05_kernel<3><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_a, tm_b, tm_abox, tm_out, p);
  else if (ctas == 2)
    conv_gemm_tcgen05_kernel<2><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_a, tm_b, tm_abox, tm_out, p);
  else
    conv_gemm_tcgen05_kernel<1><<<grid, NUM_THREADS, smem_bytes, stream>>>(tm_a, tm_b, tm_abox, tm_out, p);
  MC_LAUNCH_CHECK("conv_gemm_tcgen05_kernel");
  return 0;
}
