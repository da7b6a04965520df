// prune_filter.cu — global filter pruner ("scaled L2 norm" of each filter) on the GPU.
//
// Replaces quick_filter_prune (src/pruning/weightPruning/methods.py:28-78).  The reference computes, per conv
// weight p[O,C,kh,kw] (float32, NumPy):
//     v = np.square(p).sum(axis=1).sum(axis=1).sum(axis=1) / (C*kh*kw)          methods.py:43-44
//     v = v / np.sqrt(np.square(v).sum());  v /= np.max(v)                      methods.py:46-51
//     thr = np.percentile(concat(v of all layers) as float64, perc)             methods.py:53-55
//     mask[o] = 0 where v[o] < thr                                              methods.py:75
// Bit-exactness needs NumPy's float32 summation ORDER:
//   kh*kw > 1 : s[h,w] = sum over c, sequential in c; then sequential over h; then sequential over w.
//   kh*kw == 1: NumPy pairwise summation over the contiguous C axis (numpy/_core/src/umath/loops_utils.h.src:
//               n<8 sequential; n<=128: 8 strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail;
//               else split at n/2 rounded down to a multiple of 8).  sum(v^2) over O uses the same rule.
// All products/sums use *_rn intrinsics so nvcc cannot contract them into FMAs.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int MAX_LAYERS = MC_MAX_SEGMENTS;

struct LayerTable {
  const float* w[MAX_LAYERS];
  float* mask[MAX_LAYERS];
  int O[MAX_LAYERS], C[MAX_LAYERS], taps[MAX_LAYERS];
  int voff[MAX_LAYERS + 1];        // prefix of O
  long long woff[MAX_LAYERS + 1];  // prefix of O*C*taps
  int toff[MAX_LAYERS + 1];        // prefix of threads used by the generic sumsq path (tiled layers take none)
  int nlayers;
  // tiled path (3x3, C % 32 == 0, O % 32 == 0): work items = groups of 32 filters, layers ordered by descending C
  int torder[MAX_LAYERS];          // layer index of the i-th tiled layer
  int tblk[MAX_LAYERS + 1];        // prefix of work items over torder
  int ntiled;
};

// NumPy pairwise sum of f(a[i]) for i in [0,n), a contiguous.  SQUARE: element is a[i]*a[i] rounded to fp32 first.
template <bool SQUARE>
__device__ float np_pairwise(const float* __restrict__ a, int n) {
  auto el = [&](int i) -> float {
    const float x = a[i];
    return SQUARE ? __fmul_rn(x, x) : x;
  };
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, el(i));
    return res;
  } else if (n <= 128) {
    float r0 = el(0), r1 = el(1), r2 = el(2), r3 = el(3), r4 = el(4), r5 = el(5), r6 = el(6), r7 = el(7);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
      r0 = __fadd_rn(r0, el(i + 0));
      r1 = __fadd_rn(r1, el(i + 1));
      r2 = __fadd_rn(r2, el(i + 2));
      r3 = __fadd_rn(r3, el(i + 3));
      r4 = __fadd_rn(r4, el(i + 4));
      r5 = __fadd_rn(r5, el(i + 5));
      r6 = __fadd_rn(r6, el(i + 6));
      r7 = __fadd_rn(r7, el(i + 7));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)),
                          __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
    for (; i < n; ++i) res = __fadd_rn(res, el(i));
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise<SQUARE>(a, n2), np_pairwise<SQUARE>(a + n2, n - n2));
  }
}

// ---- exact lane-parallel version of NumPy's pairwise sum -----------------------------------------------------
// The recursion splits [0,n) into leaves of <= 128 elements (in order); a leaf is summed with 8 strided accumulators
// (or sequentially when shorter than 8).  Leaves are independent, so lane L of a warp sums leaf L (L+32, ...), and
// lane 0 then combines the leaf sums by replaying the recursion.  Same operations, same order per value => bit-exact.
template <bool SQUARE>
__device__ __forceinline__ float np_leaf_sum(const float* __restrict__ a, int n) {  // n <= 128
  return np_pairwise<SQUARE>(a, n);
}

// walk the recursion; call f(leaf_index, start, len) for every leaf, in order.  Returns the number of leaves.
template <typename F>
__device__ __forceinline__ int np_for_each_leaf(int n, F f) {
  int stack_start[24], stack_len[24];
  int sp = 0, leaf = 0;
  stack_start[0] = 0;
  stack_len[0] = n;
  sp = 1;
  while (sp > 0) {
    --sp;
    const int st = stack_start[sp], ln = stack_len[sp];
    if (ln <= 128) {
      f(leaf, st, ln);
      ++leaf;
    } else {
      int n2 = ln / 2;
      n2 -= n2 % 8;
      // push right first so the left half is visited first (in-order)
      stack_start[sp] = st + n2;
      stack_len[sp] = ln - n2;
      ++sp;
      stack_start[sp] = st;
      stack_len[sp] = n2;
      ++sp;
    }
  }
  return leaf;
}

// combine leaf sums in recursion order (replays the split; leaves are consumed in order)
__device__ float np_combine_leaves(const float* __restrict__ leaf_sums, int n, int* next_leaf) {
  if (n <= 128) return leaf_sums[(*next_leaf)++];
  int n2 = n / 2;
  n2 -= n2 % 8;
  const float l = np_combine_leaves(leaf_sums, n2, next_leaf);
  const float r = np_combine_leaves(leaf_sums, n - n2, next_leaf);
  return __fadd_rn(l, r);
}

constexpr int MAX_LEAVES = 128;  // leaves are >= 57 long once n > 128
constexpr int ROW_MAX = 2048;     // 1x1 rows up to this many channels are staged in shared memory (warp path)

// Raw per-filter sums.  taps>1: one thread per (filter, tap) runs the sequential-in-c chain (loads are issued 32 deep
// so the chain is not exposed to memory latency); the taps of a filter sit in adjacent lanes and are combined in
// NumPy's (h then w) order through shared memory.  taps==1: one WARP per filter runs the lane-parallel pairwise sum.
constexpr int SS_THREADS = 288;  // multiple of 9 and of 32

template <int NT>
__device__ void sumsq_generic_block(const LayerTable& lt, int gblk, float* __restrict__ values, float* s_tap,
                                    float (*s_leaf)[MAX_LEAVES]) {
  const int gt = gblk * NT + threadIdx.x;  // global thread slot (blocks never straddle layers)
  int l = 0;
  while (l + 1 < lt.nlayers && gt >= lt.toff[l + 1]) ++l;
  const int local = gt - lt.toff[l];
  const int taps = lt.taps[l], C = lt.C[l], O = lt.O[l];
  const float* __restrict__ w = lt.w[l];
  if (taps == 1) {
    // one WARP per filter.  The row (C floats, contiguous) is staged in shared memory with ONE round of coalesced
    // loads; the leaves of NumPy's pairwise recursion are then summed four at a time, lane = (leaf, accumulator):
    // a leaf of n >= 8 elements is 8 strided chains r[j] += a[i+j], combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) —
    // the xor-butterfly does exactly these additions (fp add is commutative) — plus a sequential tail.
    const int o = local >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (o >= O) return;  // whole warp
    const float* a = w + (long long)o * C;
    float* ls = s_leaf[wib];                                    // [MAX_LEAVES] leaf sums
    int* linfo = reinterpret_cast<int*>(s_leaf[NT / 32]) + wib * 2 * MAX_LEAVES;  // (start, len) per leaf
    float* srow = reinterpret_cast<float*>(s_leaf[NT / 32]) + (NT / 32) * 2 * MAX_LEAVES + wib * ROW_MAX;
    if (C <= ROW_MAX) {
      if ((C & 3) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        for (int i = lane; i < C / 4; i += 32) reinterpret_cast<float4*>(srow)[i] = ld_stream_f4(a4 + i);
      } else {
        for (int i = lane; i < C; i += 32) srow[i] = a[i];
      }
      int nleaves = 0;
      if (lane == 0) {
        nleaves = np_for_each_leaf(C, [&](int leaf, int st, int ln) {
          linfo[2 * leaf] = st;
          linfo[2 * leaf + 1] = ln;
        });
      }
      nleaves = __shfl_sync(0xffffffffu, nleaves, 0);
      __syncwarp();
      const int lq = lane >> 3, j = lane & 7;
      for (int g0 = 0; g0 < nleaves; g0 += 4) {
        const int leaf = g0 + lq;
        const bool have = leaf < nleaves;
        const int st = have ? linfo[2 * leaf] : 0, ln = have ? linfo[2 * leaf + 1] : 0;
        float res = 0.f;
        if (ln >= 8) {
          float x = srow[st + j];
          float r = __fmul_rn(x, x);
          const int body = ln - (ln % 8);
          for (int i = 8; i < body; i += 8) {
            x = srow[st + i + j];
            r = __fadd_rn(r, __fmul_rn(x, x));
          }
          res = r;
        }
        // all 32 lanes take part in the butterfly (leaves shorter than 8 ignore its result)
        float t = __fadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 1));
        t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
        t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 4));
        if (have && j == 0) {
          float sum;
          int i;
          if (ln >= 8) { sum = t; i = ln - (ln % 8); } else { sum = 0.f; i = 0; }
          for (; i < ln; ++i) {
            const float x = srow[st + i];
            sum = __fadd_rn(sum, __fmul_rn(x, x));
          }
          ls[leaf] = sum;
        }
      }
      __syncwarp();
      if (lane == 0) {
        int next = 0;
        const float sres = np_combine_leaves(ls, C, &next);
        values[lt.voff[l] + o] = __fdiv_rn(sres, (float)C);
      }
    } else if (lane == 0) {
      values[lt.voff[l] + o] = __fdiv_rn(np_pairwise<true>(a, C), (float)C);
    }
    return;
  }
  // blocks hold whole filters: fpb filters x taps threads, the remaining threads of the block idle
  const int fpb = NT / taps;
  const int blk = local / NT, tl = local - blk * NT;
  const int fl = tl / taps, j = tl - fl * taps;
  const int o = (fl < fpb) ? blk * fpb + fl : O;  // O = out of range -> idle
  float acc = 0.f;
  if (o < O) {
    const float* p = w + (long long)o * C * taps + j;
    int c = 0;
    for (; c + 32 <= C; c += 32) {
      float x[32];
#pragma unroll
      for (int u = 0; u < 32; ++u) x[u] = p[(long long)(c + u) * taps];
#pragma unroll
      for (int u = 0; u < 32; ++u) acc = __fadd_rn(acc, __fmul_rn(x[u], x[u]));
    }
    for (; c < C; ++c) {
      const float x = p[(long long)c * taps];
      acc = __fadd_rn(acc, __fmul_rn(x, x));
    }
  }
  s_tap[threadIdx.x] = acc;
  __syncthreads();
  if (o < O && j == 0) {
    // taps = kh*kw with kh == kw (square kernels): s[h][w] at s_tap[base + h*k + w]
    int k = 1;
    while (k * k < taps) ++k;
    const float* s = &s_tap[threadIdx.x];
    float tot = 0.f;
    for (int ww = 0; ww < k; ++ww) {
      float col = 0.f;
      for (int hh = 0; hh < k; ++hh) col = __fadd_rn(col, s[hh * k + ww]);  // .sum(axis=1) over h, sequential
      tot = __fadd_rn(tot, col);                                            // final .sum(axis=1) over w (n<8)
    }
    values[lt.voff[l] + o] = __fdiv_rn(tot, (float)(C * taps));
  }
}

// ---- tiled path for the 3x3 layers (97 % of the weights) ------------------------------------------------------
// A block owns 32 filters and walks their C axis in chunks of 32 channels: the 32 x 288 floats of a chunk are
// contiguous per filter (1152 B) and are copied with 16-byte cp.async into a 3-stage shared-memory ring (two chunks
// = 72 KB in flight per block, no registers involved); thread (f, tap) runs its 32 sequential square-adds on the
// oldest stage meanwhile.  The summation order per (filter, tap) is the same sequential-in-c chain as above.  What
// bounds a block is the LENGTH of that chain (C/32 ring steps), so a step must cost the compute, not a DRAM latency.
constexpr int TL_F = 32;               // filters per block
constexpr int TL_C = 32;               // channels per chunk
constexpr int TL_ROW4 = 74;            // float4 per tile row: 72 data + 2 pad (bank = 8f + tap: 2-way conflicts at most)
constexpr int TL_TILE4 = TL_F * TL_ROW4;
constexpr int TL_STAGES = 3;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// F filters per block (F * 9 threads run the chains and issue the copies; further threads of the block only take part
// in the barriers), STAGES ring slots of F x 32 channels.
template <int F, int STAGES>
__device__ void sumsq_tiled_block(const LayerTable& lt, int item, float* __restrict__ values,
                                  float4* s_tile /*[STAGES][F * TL_ROW4]*/, float* s_tap /*[blockDim.x]*/) {
  constexpr int WORK = F * 9;          // threads with a (filter, tap) chain
  constexpr int TILE4 = F * TL_ROW4;
  int ti = 0;
  while (ti + 1 < lt.ntiled && item >= lt.tblk[ti + 1]) ++ti;
  const int l = lt.torder[ti];
  const int C = lt.C[l];
  const int o0 = (item - lt.tblk[ti]) * F;
  const int t = threadIdx.x;
  const bool worker = t < WORK;
  const int f = worker ? t / 9 : 0, j = worker ? t - f * 9 : 0;
  const int nchunks = C / TL_C;
  // o0*C*9*4 bytes is a multiple of 16 because C % 32 == 0
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(lt.w[l] + (long long)o0 * C * 9);
  // this thread's 8 float4 of a chunk: q = u*288 + t -> (row = q / 72, col4 = q % 72)
  int grow[8], scol[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int q = u * WORK + (worker ? t : 0);
    const int row = q / 72, col4 = q - row * 72;
    grow[u] = row * (C * 9 / 4) + col4;  // float4 index inside the block's filters
    scol[u] = row * TL_ROW4 + col4;
  }
  auto issue = [&](int cc) {
    if (cc < nchunks && worker) {
      float4* tile = s_tile + (cc % STAGES) * TILE4;
      const float4* src = w4 + (long long)cc * (TL_C * 9 / 4);
#pragma unroll
      for (int u = 0; u < 8; ++u) cp_async16(tile + scol[u], src + grow[u]);
    }
    cp_async_commit();  // an empty group keeps the wait_group arithmetic uniform at the tail
  };
#pragma unroll
  for (int p = 0; p < STAGES - 1; ++p) issue(p);
  float acc = 0.f;
  for (int cc = 0; cc < nchunks; ++cc) {
    cp_async_wait<STAGES - 2>();  // this thread's copies of chunk cc have landed
    __syncthreads();              // ... and everyone's; all threads are also done with the stage refilled next
    issue(cc + STAGES - 1);
    const float* row = reinterpret_cast<const float*>(s_tile + (cc % STAGES) * TILE4 + f * TL_ROW4) + j;
#pragma unroll
    for (int c = 0; c < TL_C; ++c) {
      const float x = row[c * 9];
      acc = __fadd_rn(acc, __fmul_rn(x, x));
    }
  }
  cp_async_wait<0>();
  s_tap[t] = acc;
  __syncthreads();
  if (worker && j == 0) {
    const float* s = &s_tap[t];
    float tot = 0.f;
    for (int ww = 0; ww < 3; ++ww) {
      float col = 0.f;
      for (int hh = 0; hh < 3; ++hh) col = __fadd_rn(col, s[hh * 3 + ww]);  // .sum(axis=1) over h, sequential
      tot = __fadd_rn(tot, col);                                            // final .sum(axis=1) over w (n<8)
    }
    values[lt.voff[l] + o0 + f] = __fdiv_rn(tot, (float)(C * 9));
  }
}

// blocks [0, n_tiled): tiled work items; the first `first_wave` items (largest C first) land one per SM, the remaining
// items are taken smallest-first so that an SM's second block complements its first.  Blocks beyond: generic path.
__global__ void __launch_bounds__(SS_THREADS) filter_sumsq_kernel(const LayerTable lt, float* __restrict__ values,
                                                                  int n_tiled, int first_wave) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ float s_tap[SS_THREADS];
  if ((int)blockIdx.x < n_tiled) {
    int item = blockIdx.x;
    if (item >= first_wave) item = first_wave + (n_tiled - 1 - item);
    sumsq_tiled_block<TL_F, TL_STAGES>(lt, item, values, reinterpret_cast<float4*>(dsm), s_tap);
  } else {
    sumsq_generic_block<SS_THREADS>(lt, (int)blockIdx.x - n_tiled, values, s_tap, reinterpret_cast<float(*)[MAX_LEAVES]>(dsm));
  }
}
constexpr size_t SUMSQ_SMEM = TL_STAGES * TL_TILE4 * sizeof(float4);  // 113,664 B (the generic path needs 4.6 KB of it)

// One block per layer: v /= sqrt(pairwise(v^2)); v /= max(v).  v is staged in shared memory so the serial pairwise
// recursion of thread 0 runs at shared-memory latency.
constexpr int NORM_SMEM = 8192;
__global__ void __launch_bounds__(256) filter_norm_kernel(const LayerTable lt, float* __restrict__ values) {
  __shared__ float s_v[NORM_SMEM];
  __shared__ float s_norm;
  __shared__ float s_red[256];
  const int l = blockIdx.x;
  const int O = lt.O[l];
  float* v = values + lt.voff[l];
  const bool staged = O <= NORM_SMEM;
  if (staged)
    for (int o = threadIdx.x; o < O; o += 256) s_v[o] = v[o];
  __syncthreads();
  if (threadIdx.x == 0) s_norm = __fsqrt_rn(np_pairwise<true>(staged ? s_v : v, O));
  __syncthreads();
  const float nrm = s_norm;
  float mx = -INFINITY;
  for (int o = threadIdx.x; o < O; o += 256) {
    const float x = __fdiv_rn(staged ? s_v[o] : v[o], nrm);
    if (staged) s_v[o] = x; else v[o] = x;
    mx = fmaxf(mx, x);
  }
  s_red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] = fmaxf(s_red[threadIdx.x], s_red[threadIdx.x + s]);
    __syncthreads();
  }
  mx = s_red[0];
  for (int o = threadIdx.x; o < O; o += 256) v[o] = __fdiv_rn(staged ? s_v[o] : v[o], mx);
}

// Single-block radix select of TWO ranks at once over n non-negative floats (8 bits per pass, 4 passes), then NumPy's
// float64 _lerp.  Each pass keeps one 256-bin histogram per rank (elements matching that rank's prefix); warp 0 / warp 1
// locate the bins with shuffle prefix sums.
__device__ __forceinline__ void warp_find_bin(const unsigned int* hist, unsigned int rank, unsigned int* out_bin,
                                              unsigned int* out_rem) {
  const int lane = threadIdx.x & 31;
  unsigned int c[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned int excl = incl - tot;
  const bool mine = (rank >= excl) && (rank < incl);
  const unsigned int who = __ballot_sync(0xffffffffu, mine);
  const int src = who ? (__ffs(who) - 1) : 31;  // rank beyond the total: clamp (caller validates)
  if (lane == src) {
    unsigned int cum = excl, b = 0;
    for (; b < 7; ++b) {
      if (cum + c[b] > rank) break;
      cum += c[b];
    }
    *out_bin = lane * 8 + b;
    *out_rem = rank - cum;
  }
}

__global__ void __launch_bounds__(1024) filter_threshold_kernel(const float* __restrict__ v, int n, long long k,
                                                                double gamma, double* __restrict__ thr) {
  __shared__ unsigned int s_hist[2][256];
  __shared__ unsigned int s_bin[2], s_rem[2];
  __shared__ int s_nan;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (v[i] != v[i]) s_nan = 1;  // np.percentile returns nan if any value is nan (an all-zero layer gives 0/0)
  __syncthreads();
  if (s_nan) {
    if (threadIdx.x == 0) *thr = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  const long long k1 = (k + 1 < n) ? k + 1 : (long long)n - 1;
  unsigned int rank[2] = {(unsigned int)k, (unsigned int)k1};
  unsigned int prefix[2] = {0, 0}, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned int key = __float_as_uint(v[i]);
      const unsigned int d = (key >> shift) & 255u;
      if ((key & mask) == prefix[0]) atomicAdd(&s_hist[0][d], 1u);
      if ((key & mask) == prefix[1]) atomicAdd(&s_hist[1][d], 1u);
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5;
    if (wid < 2) warp_find_bin(s_hist[wid], rank[wid], &s_bin[wid], &s_rem[wid]);
    __syncthreads();
    prefix[0] |= s_bin[0] << shift;
    prefix[1] |= s_bin[1] << shift;
    rank[0] = s_rem[0];
    rank[1] = s_rem[1];
    mask |= 255u << shift;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double a = (double)__uint_as_float(prefix[0]);
    const double b = (double)__uint_as_float(prefix[1]);
    const double diff = __dsub_rn(b, a);
    double r = __dadd_rn(a, __dmul_rn(diff, gamma));
    if (gamma >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
    *thr = r;
  }
}

// ---- finish: per-layer normalisation, then (last block to arrive) the float64 percentile and the keep flags -------
// Order statistics by bitwise binary search on the fp32 bit pattern (values are >= 0, so the pattern is monotone):
// 31 rounds of "count keys below prefix|bit", both ranks at once, keys in shared memory — ~4 us for 10,461 values,
// where a shared-memory radix histogram serialises on the few distinct high digits of values that all lie in (0,1].
constexpr int FIN_THREADS = 1024;

__device__ void block_count2(unsigned int c0, unsigned int c1, unsigned int* s_red /*[64]*/, unsigned int* out0,
                             unsigned int* out1) {
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[wid] = c0; s_red[32 + wid] = c1; }
  __syncthreads();
  if (wid == 0) {
    unsigned int a = s_red[lane], b = s_red[32 + lane];  // FIN_THREADS / 32 == 32 warps
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { s_red[0] = a; s_red[32] = b; }
  }
  __syncthreads();
  *out0 = s_red[0];
  *out1 = s_red[32];
  __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS) filter_finish_kernel(const LayerTable lt, float* __restrict__ values,
                                                                    long long k, double gamma, double* __restrict__ thr,
                                                                    uint8_t* __restrict__ keep,
                                                                    unsigned int* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char dsm[];
  float* s_v = reinterpret_cast<float*>(dsm);  // max(O_l, n) floats
  __shared__ float s_norm;
  __shared__ float s_leaf[MAX_LEAVES];
  __shared__ unsigned int s_red[64];
  __shared__ unsigned int s_last;
  const int l = blockIdx.x;
  const int O = lt.O[l];
  float* v = values + lt.voff[l];
  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(ticket) + 1;  // diagnostics: phase stamps of the last block
  auto stamp = [&](int i) {
    if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[i] = t; }
  };
  unsigned long long t_begin = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
  // ---- v /= sqrt(pairwise(v^2)); v /= max(v)   (methods.py:46-51)
  for (int o = threadIdx.x; o < O; o += FIN_THREADS) s_v[o] = v[o];
  __syncthreads();
  if (O <= 128 * MAX_LEAVES) {
    if (threadIdx.x < 32) {  // lane-parallel replay of NumPy's pairwise recursion (leaves are independent)
      const int lane = threadIdx.x;
      np_for_each_leaf(O, [&](int leaf, int st, int ln) {
        if ((leaf & 31) == lane && leaf < MAX_LEAVES) s_leaf[leaf] = np_leaf_sum<true>(s_v + st, ln);
      });
      __syncwarp();
      if (lane == 0) {
        int next = 0;
        s_norm = __fsqrt_rn(np_combine_leaves(s_leaf, O, &next));
      }
    }
  } else if (threadIdx.x == 0) {
    s_norm = __fsqrt_rn(np_pairwise<true>(s_v, O));
  }
  __syncthreads();
  const float nrm = s_norm;
  float mx = -INFINITY;
  for (int o = threadIdx.x; o < O; o += FIN_THREADS) {
    const float x = __fdiv_rn(s_v[o], nrm);
    s_v[o] = x;
    mx = fmaxf(mx, x);
  }
  {
    unsigned int a, b;  // max of non-negative floats == max of their bit patterns; NaN (0/0) handled below
    block_count2(0u, 0u, s_red, &a, &b);  // barrier only
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = __float_as_uint(mx);
    __syncthreads();
    if (threadIdx.x == 0) {
      float m2 = -INFINITY;
      for (int w2 = 0; w2 < FIN_THREADS / 32; ++w2) m2 = fmaxf(m2, __uint_as_float(s_red[w2]));
      s_norm = m2;
    }
    __syncthreads();
  }
  mx = s_norm;
  for (int o = threadIdx.x; o < O; o += FIN_THREADS) v[o] = __fdiv_rn(s_v[o], mx);
  // ---- last block to finish computes the threshold over all layers
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) dbg[0] = t_begin;
  stamp(1);
  const int n = lt.voff[lt.nlayers];
  unsigned int* s_k = reinterpret_cast<unsigned int*>(dsm);
  unsigned int nan = 0;
  for (int i = threadIdx.x; i < n; i += FIN_THREADS) {
    const float x = __ldcg(values + i);
    s_k[i] = __float_as_uint(x);
    nan |= (x != x);
  }
  __syncthreads();
  unsigned int nn, dummy;
  block_count2(nan, 0u, s_red, &nn, &dummy);
  if (nn) {  // np.percentile returns nan if any value is nan (an all-zero layer gives 0/0): nothing is < nan
    if (threadIdx.x == 0) *thr = __longlong_as_double(0x7ff8000000000000LL);
    if (keep)
      for (int i = threadIdx.x; i < n; i += FIN_THREADS) keep[i] = 1;
    return;
  }
  stamp(2);
  const unsigned int r0 = (unsigned int)k, r1 = (unsigned int)((k + 1 < n) ? k + 1 : (long long)n - 1);
  unsigned int p0 = 0, p1 = 0;
  // The rounds run on ONE SM, so what they cost is warp-instructions issued (4 per clock), not latency: only the
  // first SEL_THREADS threads take part (the others leave), keys sit in registers (n <= KPT*SEL_THREADS, else shared
  // memory), the two ranks share one compare while their prefixes coincide, and the cross-warp sum is: redux -> one
  // slot per warp -> ONE named barrier -> every warp adds the slots itself (slots alternate by round parity).
  constexpr int SEL_THREADS = 512, SEL_WARPS = SEL_THREADS / 32, KPT = 24;
  __shared__ unsigned int s_part[2][2][SEL_WARPS];
  __syncthreads();  // s_k complete, NaN decision taken by everyone
  if (threadIdx.x >= SEL_THREADS) return;
  const bool in_regs = n <= KPT * SEL_THREADS;
  unsigned int kreg[KPT];
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int i = u * SEL_THREADS + threadIdx.x;
    kreg[u] = (in_regs && i < n) ? s_k[i] : 0xffffffffu;  // padding never counts as "below"
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int bit = 30, r = 0; bit >= 0; --bit, ++r) {  // keys < 2^31
    const unsigned int c0 = p0 | (1u << bit), c1 = p1 | (1u << bit);
    const bool split = p0 != p1;  // block-uniform
    unsigned int n0 = 0, n1 = 0;
    if (in_regs) {
#pragma unroll
      for (int u = 0; u < KPT; ++u) n0 += kreg[u] < c0;
      if (split) {
#pragma unroll
        for (int u = 0; u < KPT; ++u) n1 += kreg[u] < c1;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += SEL_THREADS) {
        const unsigned int key = s_k[i];
        n0 += key < c0;
        n1 += key < c1;
      }
    }
    n0 = __reduce_add_sync(0xffffffffu, n0);
    if (split || !in_regs) n1 = __reduce_add_sync(0xffffffffu, n1);
    if (lane == 0) {
      s_part[r & 1][0][wid] = n0;
      s_part[r & 1][1][wid] = n1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SEL_THREADS) : "memory");
    const unsigned int t0 = __reduce_add_sync(0xffffffffu, lane < SEL_WARPS ? s_part[r & 1][0][lane] : 0u);
    unsigned int t1 = t0;
    if (split || !in_regs) t1 = __reduce_add_sync(0xffffffffu, lane < SEL_WARPS ? s_part[r & 1][1][lane] : 0u);
    if (t0 <= r0) p0 = c0;  // at most r0 keys below the candidate: the r0-th smallest is >= candidate
    if (t1 <= r1) p1 = c1;
  }
  double t;
  {
    const double a = (double)__uint_as_float(p0);
    const double b = (double)__uint_as_float(p1);
    const double diff = __dsub_rn(b, a);
    t = __dadd_rn(a, __dmul_rn(diff, gamma));
    if (gamma >= 0.5) t = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
  }
  stamp(3);
  if (threadIdx.x == 0) *thr = t;
  if (keep)
    for (int i = threadIdx.x; i < n; i += SEL_THREADS) keep[i] = ((double)__uint_as_float(s_k[i]) < t) ? 0 : 1;
  stamp(4);
}

__global__ void filter_keep_kernel(const float* __restrict__ v, int n, const double* __restrict__ thr,
                                   uint8_t* __restrict__ keep) {
  const double t = *thr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    keep[i] = ((double)v[i] < t) ? 0 : 1;
}

// one block per filter (grid-stride): constant fill of `per` floats, 128-bit stores on the aligned body
__global__ void __launch_bounds__(256) filter_mask_fill_kernel(const LayerTable lt, const float* __restrict__ v,
                                                               const double* __restrict__ thr) {
  const double t = *thr;
  const int nfilters = lt.voff[lt.nlayers];
  for (int f = blockIdx.x; f < nfilters; f += gridDim.x) {
    int l = 0;
    while (l + 1 < lt.nlayers && f >= lt.voff[l + 1]) ++l;
    const int o = f - lt.voff[l];
    const int per = lt.C[l] * lt.taps[l];
    const float val = ((double)v[f] < t) ? 0.f : 1.f;
    float* m = lt.mask[l] + (long long)o * per;
    // scalar head up to 16-byte alignment
    int head = (int)(((16 - (reinterpret_cast<uintptr_t>(m) & 15)) & 15) >> 2);
    if (head > per) head = per;
    if ((int)threadIdx.x < head) m[threadIdx.x] = val;
    const int body4 = (per - head) >> 2;
    float4* m4 = reinterpret_cast<float4*>(m + head);
    const float4 v4 = make_float4(val, val, val, val);
    for (int i = threadIdx.x; i < body4; i += 256) st_stream_f4(m4 + i, v4);
    const int tail0 = head + body4 * 4;
    if ((int)threadIdx.x < per - tail0) m[tail0 + threadIdx.x] = val;
  }
}

// ---- the whole of quick_filter_prune in ONE cooperative launch ------------------------------------------------------
// Persistent blocks (2 per SM, all co-resident) take work items from an atomic queue: groups of FU_F filters of the 3x3
// layers through a 6-stage cp.async ring, and the blocks of the generic path (1x1 layers) slotted in behind the long
// items so that their latency-bound work does not form the tail.  Raw per-filter sums go to a scratch array.  When the
// queue is empty EVERY block waits for the last item, pulls all raw sums (42 KB for Darknet-19) from L2 into shared
// memory and redundantly runs the small serial part itself — per-layer normalisation in NumPy's pairwise order, then a
// two-rank radix select for np.percentile — so nothing is exchanged between blocks after the sums and no block waits
// for another one's result; it then writes its slice of the normalised values / keep flags and fills its share of the
// masks.  Against three launches this removes two full drains + ramps of the grid and the single-block finish kernel
// (~30 us of pure latency between the two bandwidth passes).
constexpr int FU_THREADS = 160;  // 16 filters x 9 taps = 144 chain threads, rounded up to whole warps
constexpr int FU_F = 16;
constexpr int FU_STAGES = 6;
constexpr size_t FU_SMEM = (size_t)FU_STAGES * FU_F * TL_ROW4 * sizeof(float4);  // 113,664 B: two blocks per SM
constexpr int FU_AUX_BYTES = 8192;                                                // leaf table / histograms
constexpr int FU_MAX_N = (int)((FU_SMEM - FU_AUX_BYTES) / sizeof(float));         // 26,368 filters
static_assert((size_t)(FU_THREADS / 32) * MAX_LEAVES * 4 <= (size_t)FU_AUX_BYTES, "leaf sums must fit the aux region");

struct FusedState {
  unsigned int next_item;  // work queue
  unsigned int pad0[31];
  unsigned int items_done;
  unsigned int pad1[31];
  unsigned int flag;  // polled by idle blocks: its own 128-byte line
  unsigned int pad2[31];
  unsigned long long stamp[8];  // diagnostics (globaltimer ns): block 0 start, last block out of work, last block past the
                                // wait, last block with thr, last block done
};
__device__ __forceinline__ void fused_stamp(FusedState* st, int i, bool max_over_blocks) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (max_over_blocks) atomicMax(&st->stamp[i], t);
  else st->stamp[i] = t;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All raw sums -> normalised values in shared memory (methods.py:43-51 for every layer), by one block.
// s_val [n]; aux: leaf table (start, len) + leaf sums.  Returns true (block-uniform) if any value is NaN.
template <int NT>
__device__ bool fused_normalise_all(const LayerTable& lt, const float* __restrict__ raw, float* s_val, unsigned char* aux) {
  const int n = lt.voff[lt.nlayers], nl = lt.nlayers;
  const int tid = threadIdx.x;
  float* s_lsum = reinterpret_cast<float*>(aux);  // [NT / 32][MAX_LEAVES] leaf sums, one row per warp
  for (int i0 = 0; i0 < n; i0 += 16 * NT) {  // 16 independent L2 loads in flight per thread
    float x[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = i0 + u * NT + tid;
      x[u] = i < n ? __ldcg(raw + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = i0 + u * NT + tid;
      if (i < n) s_val[i] = x[u];
    }
  }
  __syncthreads();
  // One warp per layer, start to finish (no block barrier inside).  ||v||: O <= 128 is one leaf of NumPy's pairwise
  // recursion, O = 128 * 2^k splits into 2^k leaves of 128 combined as a balanced tree (the recursion halves exactly);
  // a leaf is 8 strided accumulator chains combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail — lane =
  // (leaf, accumulator), the xor-butterfly performs exactly those additions.  Any other O: lane 0 runs the recursion.
  // Then x = v / ||v||, max(x) (a NaN wins, as np.max propagates it), v = x / max.
  unsigned int nan = 0;
  {
    const int lane = tid & 31, wid = tid >> 5;
    float* ls = s_lsum + wid * MAX_LEAVES;
    for (int l = wid; l < nl; l += NT / 32) {
      const int O = lt.O[l];
      float* vv = s_val + lt.voff[l];
      const int nleaf = O <= 128 ? 1 : O / 128;
      const bool fast = O <= 128 || ((O % 128) == 0 && (nleaf & (nleaf - 1)) == 0 && nleaf <= MAX_LEAVES);
      float nrm = 0.f;
      if (fast) {
        const int lq = lane >> 3, j = lane & 7;
        for (int g0 = 0; g0 < nleaf; g0 += 4) {
          const int leaf = g0 + lq;
          const bool have = leaf < nleaf;
          const int st = leaf * 128, ln = have ? (O <= 128 ? O : 128) : 0;
          float r = 0.f;
          if (ln >= 8) {
            float x = vv[st + j];
            r = __fmul_rn(x, x);
            const int body = ln - (ln % 8);
            for (int i = 8; i < body; i += 8) {
              x = vv[st + i + j];
              r = __fadd_rn(r, __fmul_rn(x, x));
            }
          }
          float t = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
          t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
          t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 4));
          if (have && j == 0) {
            float sum;
            int i;
            if (ln >= 8) { sum = t; i = ln - (ln % 8); } else { sum = 0.f; i = 0; }
            for (; i < ln; ++i) {
              const float x = vv[st + i];
              sum = __fadd_rn(sum, __fmul_rn(x, x));
            }
            ls[leaf] = sum;
          }
        }
        __syncwarp();
        for (int w2 = 1; w2 < nleaf; w2 <<= 1) {  // balanced tree: (L0+L1), (L2+L3), ... then pairs of those, ...
          for (int i = lane * 2 * w2; i + w2 < nleaf; i += 64 * w2) ls[i] = __fadd_rn(ls[i], ls[i + w2]);
          __syncwarp();
        }
        nrm = __fsqrt_rn(ls[0]);
      } else {
        if (lane == 0) ls[0] = __fsqrt_rn(np_pairwise<true>(vv, O));
        __syncwarp();
        nrm = ls[0];
      }
      __syncwarp();
      unsigned int mb = 0u;
      for (int i = lane; i < O; i += 32) {
        const float x = __fdiv_rn(vv[i], nrm);
        vv[i] = x;
        mb = max(mb, __float_as_uint(x) & 0x7fffffffu);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
      const float mx = __uint_as_float(mb);
      for (int i = lane; i < O; i += 32) {
        const float v = __fdiv_rn(vv[i], mx);
        vv[i] = v;
        nan |= (v != v);
      }
    }
  }
  return __syncthreads_or((int)nan) != 0;
}

// np.percentile over the n normalised values in shared memory: the order statistics k and k+1 (both at once), then
// NumPy's float64 _lerp.  The values of a layer crowd into a handful of exponents, so the top byte of the bit patterns
// is resolved by 7 rounds of counting "keys below prefix|bit" (a shared-memory histogram would serialise 32-way on
// those few bins: measured 10 us for that pass alone); the three lower bytes, which are well spread, by 8-bit radix
// passes with shared-memory atomics (one histogram per rank).  Every thread returns the threshold.
template <int NT>
__device__ double fused_select(const float* s_val, int n, long long k, double gamma, unsigned int* s_hist /*[2][256]*/,
                               unsigned int* s_bin /*[4]*/) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int WARPS = NT / 32;
  const long long k1 = (k + 1 < n) ? k + 1 : (long long)n - 1;
  unsigned int rank[2] = {(unsigned int)k, (unsigned int)k1};
  unsigned int prefix[2] = {0, 0};
  unsigned int* s_part = s_hist;  // [2 parities][2 ranks][WARPS] partial counts of the counting rounds
  for (int bit = 30, r = 0; bit >= 24; --bit, ++r) {  // keys < 2^31
    const unsigned int c0 = prefix[0] | (1u << bit), c1 = prefix[1] | (1u << bit);
    unsigned int n0 = 0, n1 = 0;
    for (int i = tid; i < n; i += NT) {
      const unsigned int key = __float_as_uint(s_val[i]);
      n0 += key < c0;
      n1 += key < c1;
    }
    n0 = __reduce_add_sync(0xffffffffu, n0);
    n1 = __reduce_add_sync(0xffffffffu, n1);
    unsigned int* part = s_part + (r & 1) * 2 * WARPS;
    if (lane == 0) {
      part[wid] = n0;
      part[WARPS + wid] = n1;
    }
    __syncthreads();  // (slots alternate by round parity: one barrier per round)
    const unsigned int t0 = __reduce_add_sync(0xffffffffu, lane < WARPS ? part[lane] : 0u);
    const unsigned int t1 = __reduce_add_sync(0xffffffffu, lane < WARPS ? part[WARPS + lane] : 0u);
    if (t0 <= (unsigned int)k) prefix[0] = c0;   // at most k keys below the candidate: the k-th smallest is >= candidate
    if (t1 <= (unsigned int)k1) prefix[1] = c1;
  }
  {  // ranks relative to the keys that share the top byte: subtract the keys below the prefix
    unsigned int n0 = 0, n1 = 0;
    for (int i = tid; i < n; i += NT) {
      const unsigned int key = __float_as_uint(s_val[i]);
      n0 += key < prefix[0];
      n1 += key < prefix[1];
    }
    n0 = __reduce_add_sync(0xffffffffu, n0);
    n1 = __reduce_add_sync(0xffffffffu, n1);
    __syncthreads();
    if (lane == 0) {
      s_part[wid] = n0;
      s_part[WARPS + wid] = n1;
    }
    __syncthreads();
    rank[0] -= __reduce_add_sync(0xffffffffu, lane < WARPS ? s_part[lane] : 0u);
    rank[1] -= __reduce_add_sync(0xffffffffu, lane < WARPS ? s_part[WARPS + lane] : 0u);
    __syncthreads();
  }
  unsigned int mask = 0xff000000u;
  for (int shift = 16; shift >= 0; shift -= 8) {
    for (int i = tid; i < 512; i += NT) s_hist[i] = 0;
    __syncthreads();
    const bool same = prefix[0] == prefix[1];  // block-uniform: one histogram serves both ranks while they agree
    for (int i = tid; i < n; i += NT) {
      const unsigned int key = __float_as_uint(s_val[i]);
      const unsigned int d = (key >> shift) & 255u;
      if ((key & mask) == prefix[0]) atomicAdd(&s_hist[d], 1u);
      else if (!same && (key & mask) == prefix[1]) atomicAdd(&s_hist[256 + d], 1u);
    }
    __syncthreads();
    if (wid == 0) warp_find_bin(s_hist, rank[0], &s_bin[0], &s_bin[2]);
    if (wid == 1) warp_find_bin(same ? s_hist : s_hist + 256, rank[1], &s_bin[1], &s_bin[3]);
    __syncthreads();
    prefix[0] |= s_bin[0] << shift;
    prefix[1] |= s_bin[1] << shift;
    rank[0] = s_bin[2];
    rank[1] = s_bin[3];
    mask |= 255u << shift;
    __syncthreads();
  }
  const double a = (double)__uint_as_float(prefix[0]);
  const double b = (double)__uint_as_float(prefix[1]);
  const double diff = __dsub_rn(b, a);
  double r = __dadd_rn(a, __dmul_rn(diff, gamma));
  if (gamma >= 0.5) r = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
  return r;
}

__global__ void __launch_bounds__(FU_THREADS, 2)
filter_prune_fused_kernel(const LayerTable lt, float* __restrict__ raw, float* __restrict__ values, int n_tiled, int n_big,
                          int n_items, long long k, double gamma, double* __restrict__ thr, uint8_t* __restrict__ keep,
                          FusedState* st, int fill) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ float s_tap[FU_THREADS];
  __shared__ unsigned int s_bin[4];
  __shared__ int s_item;
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  const int n = lt.voff[lt.nlayers];
  if (tid == 0 && blockIdx.x == 0) fused_stamp(st, 0, false);
  if (tid == 0) s_last = 0u;  // (a block that never gets an item is not the finisher)
  __syncthreads();
  // ---- phase 1: raw per-filter sums
  for (;;) {
    if (tid == 0) s_item = (int)atomicAdd(&st->next_item, 1u);
    __syncthreads();
    const int qpos = s_item;
    __syncthreads();
    if (qpos >= n_items) break;
    // queue order: the long tiled items (C >= 1024, one per block: the bandwidth-bound bulk), then the short generic
    // blocks (1x1 layers: latency-bound, hidden behind the bulk instead of forming the tail), then the short tiled items
    const int n_generic = n_items - n_tiled;
    const int item = qpos < n_big ? qpos : (qpos < n_big + n_generic ? n_tiled + (qpos - n_big) : qpos - n_generic);
    if (item < n_tiled)
      sumsq_tiled_block<FU_F, FU_STAGES>(lt, item, raw, reinterpret_cast<float4*>(dsm), s_tap);
    else
      sumsq_generic_block<FU_THREADS>(lt, item - n_tiled, raw, s_tap, reinterpret_cast<float(*)[MAX_LEAVES]>(dsm));
    __threadfence();  // this block's sums are visible device-wide before it counts the item in
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&st->items_done, 1u) == (unsigned int)n_items - 1u) ? 1u : 0u;
    __syncthreads();
    const bool finisher = s_last != 0u;  // block-uniform: this block completed the LAST item
    __syncthreads();
    if (!finisher) continue;
    // ---- phase 2 (the finisher alone; every other block is out of work or about to be): all sums -> normalised values
    //      -> threshold, published for everyone
    __threadfence();
    if (tid == 0) fused_stamp(st, 2, false);
    float* s_val = reinterpret_cast<float*>(dsm);
    unsigned char* aux = dsm + (FU_SMEM - FU_AUX_BYTES);
    const bool any_nan = fused_normalise_all<FU_THREADS>(lt, raw, s_val, aux);
    if (tid == 0) fused_stamp(st, 5, false);
    double t;
    if (any_nan) t = __longlong_as_double(0x7ff8000000000000LL);  // np.percentile returns nan: nothing is < nan
    else t = fused_select<FU_THREADS>(s_val, n, k, gamma, reinterpret_cast<unsigned int*>(aux), s_bin);
    for (int i = tid; i < n; i += FU_THREADS) {
      values[i] = s_val[i];
      if (keep) keep[i] = ((double)s_val[i] < t) ? 0 : 1;
    }
    if (tid == 0) *thr = t;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      fused_stamp(st, 3, false);
      atomicExch(&st->flag, 1u);
    }
    __syncthreads();
  }
  if (tid == 0) fused_stamp(st, 1, true);
  if (!fill) return;
  // ---- phase 3: masks.  A block owns the filters blockIdx.x, blockIdx.x + gridDim.x, ... in both passes below.
  // While the finisher runs the serial part (normalise + percentile: ~38 us during which HBM would sit idle), every other
  // block already fills ITS filters with the PROVISIONAL value — the majority outcome, known from the rank alone: 1 when
  // fewer than half of the filters will be pruned, else 0 — and once the threshold is published only the filters whose
  // flag differs are rewritten (same thread, same addresses: ordered).  The finisher itself writes its filters once.
  const float prov = (2 * k < (long long)n) ? 1.f : 0.f;
  const bool i_finished = (s_last != 0u);  // block-uniform (s_last is only rewritten inside the queue loop)
  auto fill_filters = [&](int pass) {
    const double t = pass ? __ldcg(thr) : 0.0;
    int l = 0;
    for (int f = blockIdx.x; f < n; f += gridDim.x) {
      while (l + 1 < lt.nlayers && f >= lt.voff[l + 1]) ++l;
      float val = prov;
      if (pass) {
        val = ((double)__ldcg(values + f) < t) ? 0.f : 1.f;
        if (!i_finished && val == prov) continue;  // already written by the provisional pass
      }
      const int o = f - lt.voff[l];
      const int per = lt.C[l] * lt.taps[l];
      float* m = lt.mask[l] + (long long)o * per;
      int head = (int)(((16 - (reinterpret_cast<uintptr_t>(m) & 15)) & 15) >> 2);  // scalar head up to 16-byte alignment
      if (head > per) head = per;
      if (tid < head) m[tid] = val;
      const int body4 = (per - head) >> 2;
      float4* m4 = reinterpret_cast<float4*>(m + head);
      const float4 v4 = make_float4(val, val, val, val);
      for (int i = tid; i < body4; i += FU_THREADS) st_stream_f4(m4 + i, v4);
      const int tail0 = head + body4 * 4;
      if (tid < per - tail0) m[tail0 + tid] = val;
    }
  };
  if (!i_finished) fill_filters(0);
  if (tid == 0) {
    unsigned int spins = 0;
    while (ld_acquire_u32(&st->flag) == 0u) {
      __nanosleep(128);
      if (++spins > (1u << 26)) __trap();  // (a broken schedule must not hang the GPU)
    }
  }
  __syncthreads();
  fill_filters(1);
  if (tid == 0) fused_stamp(st, 4, true);
}

int build_layers(LayerTable* lt, const float* const* w, float* const* masks, const int* O, const int* C,
                 const int* taps, int nlayers, const char* who, int nt = SS_THREADS, int tl_f = TL_F) {
  if (nlayers <= 0 || nlayers > MAX_LAYERS) return mc_set_error(MC_ERR_ARG, "%s: nlayers %d out of range", who, nlayers);
  lt->nlayers = nlayers;
  lt->voff[0] = 0;
  lt->woff[0] = 0;
  lt->toff[0] = 0;
  for (int l = 0; l < nlayers; ++l) {
    if (O[l] <= 0 || C[l] <= 0 || taps[l] <= 0) return mc_set_error(MC_ERR_ARG, "%s: layer %d has bad dims", who, l);
    lt->w[l] = w ? w[l] : nullptr;
    lt->mask[l] = masks ? masks[l] : nullptr;
    lt->O[l] = O[l];
    lt->C[l] = C[l];
    lt->taps[l] = taps[l];
    lt->voff[l + 1] = lt->voff[l] + O[l];
    // element offsets are rounded up to 4 so the vectorised mask fill never straddles two layers
    const long long sz = (long long)O[l] * C[l] * taps[l];
    lt->woff[l + 1] = lt->woff[l] + ((sz + 3) / 4) * 4;
    const long long thr = (taps[l] == 1) ? O[l] : (long long)O[l] * taps[l];
    // per-layer thread slots rounded up to whole blocks; for taps>1 a block must hold whole filters
    long long slots;
    if (w && taps[l] == 9 && (C[l] % TL_C) == 0 && (O[l] % tl_f) == 0) slots = 0;  // tiled path
    else if (taps[l] == 1) slots = ((thr * 32 + nt - 1) / nt) * nt;  // one warp per filter
    else {
      const int fpb = nt / taps[l];  // filters per block
      slots = (long long)((O[l] + fpb - 1) / fpb) * nt;
    }
    lt->toff[l + 1] = lt->toff[l] + (int)slots;
  }
  // tiled layers, largest C first (insertion sort; nlayers <= 64)
  lt->ntiled = 0;
  for (int l = 0; l < nlayers; ++l) {
    if (!(w && taps[l] == 9 && (C[l] % TL_C) == 0 && (O[l] % tl_f) == 0)) continue;
    int pos = lt->ntiled++;
    while (pos > 0 && C[lt->torder[pos - 1]] < C[l]) {
      lt->torder[pos] = lt->torder[pos - 1];
      --pos;
    }
    lt->torder[pos] = l;
  }
  lt->tblk[0] = 0;
  for (int i = 0; i < lt->ntiled; ++i) lt->tblk[i + 1] = lt->tblk[i] + O[lt->torder[i]] / tl_f;
  for (int i = lt->ntiled; i < MAX_LAYERS; ++i) {
    lt->torder[i] = 0;
    lt->tblk[i + 1] = lt->tblk[lt->ntiled];
  }
  for (int l = nlayers; l < MAX_LAYERS; ++l) {
    lt->w[l] = nullptr;
    lt->mask[l] = nullptr;
    lt->O[l] = lt->C[l] = lt->taps[l] = 0;
    lt->voff[l + 1] = lt->voff[nlayers];
    lt->woff[l + 1] = lt->woff[nlayers];
    lt->toff[l + 1] = lt->toff[nlayers];
  }
  return 0;
}

}  // namespace

namespace {
int launch_sumsq(const LayerTable& lt, float* d_values, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    MC_CUDA(cudaFuncSetAttribute(filter_sumsq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SUMSQ_SMEM));
    attr_set = true;
  }
  const int n_tiled = lt.tblk[lt.ntiled];
  const int n_generic = lt.toff[lt.nlayers] / SS_THREADS;
  const int first_wave = n_tiled < mc_num_sms() ? n_tiled : mc_num_sms();
  if (n_tiled + n_generic == 0) return 0;
  const char* dbg = mc_tune_env("MCB200_SUMSQ_ONLY");  // timing experiments only (results are incomplete)
  if (dbg && dbg[0] == 't') {
    filter_sumsq_kernel<<<n_tiled, SS_THREADS, SUMSQ_SMEM, stream>>>(lt, d_values, n_tiled, first_wave);
    return 0;
  }
  if (dbg && dbg[0] == 'g') {
    filter_sumsq_kernel<<<n_generic, SS_THREADS, SUMSQ_SMEM, stream>>>(lt, d_values, 0, 0);
    return 0;
  }
  filter_sumsq_kernel<<<n_tiled + n_generic, SS_THREADS, SUMSQ_SMEM, stream>>>(lt, d_values, n_tiled, first_wave);
  MC_LAUNCH_CHECK("filter_sumsq_kernel");
  return 0;
}
}  // namespace

extern "C" int mc_filter_values(const float* const* h_w_ptrs, const int* h_O, const int* h_C, const int* h_kh,
                                const int* h_kw, int nlayers, float* d_values, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(h_w_ptrs && h_O && h_C && h_kh && h_kw && d_values, "mc_filter_values: null pointer");
  MC_CHECK_ARG(nlayers > 0 && nlayers <= MAX_LAYERS, "mc_filter_values: nlayers out of range");
  int taps[MAX_LAYERS];
  for (int l = 0; l < nlayers; ++l) {
    MC_CHECK_ARG(h_kh[l] == h_kw[l] && h_kh[l] >= 1 && h_kh[l] * h_kw[l] <= 49,
                 "mc_filter_values: layer %d kernel %dx%d unsupported (square, <=7x7)", l, h_kh[l], h_kw[l]);
    MC_CHECK_ARG(h_w_ptrs[l] != nullptr, "mc_filter_values: layer %d null weight", l);
    taps[l] = h_kh[l] * h_kw[l];
  }
  LayerTable lt;
  int rc = build_layers(&lt, h_w_ptrs, nullptr, h_O, h_C, taps, nlayers, "mc_filter_values");
  if (rc) return rc;
  rc = launch_sumsq(lt, d_values, stream);
  if (rc) return rc;
  filter_norm_kernel<<<nlayers, 256, 0, stream>>>(lt, d_values);
  MC_LAUNCH_CHECK("filter_norm_kernel");
  return 0;
}

extern "C" size_t mc_workspace_bytes_filter_threshold(int n) {
  (void)n;
  return 0;
}

extern "C" int mc_filter_threshold(const float* d_values, int n, int64_t k, double gamma, double* d_thr, void* d_ws,
                                   size_t ws_bytes, void* stream_) {
  (void)d_ws;
  (void)ws_bytes;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_values && d_thr && n > 0, "mc_filter_threshold: bad argument");
  MC_CHECK_ARG(k >= 0 && k < n, "mc_filter_threshold: rank out of range");
  MC_CHECK_ARG(gamma >= 0.0 && gamma < 1.0, "mc_filter_threshold: gamma must be in [0,1)");
  filter_threshold_kernel<<<1, 1024, 0, stream>>>(d_values, n, (long long)k, gamma, d_thr);
  MC_LAUNCH_CHECK("filter_threshold_kernel");
  return 0;
}

extern "C" int mc_filter_masks(const float* d_values, const double* d_thr, const int* h_O, const int* h_per_filter,
                               int nlayers, float* const* h_mask_ptrs, uint8_t* d_keep, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_values && d_thr && h_O && h_per_filter, "mc_filter_masks: null pointer");
  MC_CHECK_ARG(h_mask_ptrs || d_keep, "mc_filter_masks: nothing to do");
  int ones[MAX_LAYERS];
  for (int l = 0; l < nlayers && l < MAX_LAYERS; ++l) ones[l] = 1;
  LayerTable lt;
  // per-filter element count goes into C, taps = 1
  int rc = build_layers(&lt, nullptr, h_mask_ptrs, h_O, h_per_filter, ones, nlayers, "mc_filter_masks");
  if (rc) return rc;
  const int n = lt.voff[nlayers];
  if (d_keep) {
    filter_keep_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_values, n, d_thr, d_keep);
    MC_LAUNCH_CHECK("filter_keep_kernel");
  }
  if (h_mask_ptrs) {
    for (int l = 0; l < nlayers; ++l) MC_CHECK_ARG(h_mask_ptrs[l] != nullptr, "mc_filter_masks: null mask %d", l);
    int blocks = lt.voff[nlayers];
    const int cap = mc_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    filter_mask_fill_kernel<<<blocks, 256, 0, stream>>>(lt, d_values, d_thr);
    MC_LAUNCH_CHECK("filter_mask_fill_kernel");
  }
  return 0;
}

/* The whole of quick_filter_prune in one call: ONE cooperative launch (filter_prune_fused_kernel); models with more
 * filters than the single-block select holds take the three-launch path.  See include/mcb200.h. */
extern "C" size_t mc_workspace_bytes_filter_prune(void) { return 2048 + (size_t)FU_MAX_N * sizeof(float); }
static_assert(sizeof(FusedState) <= 2048, "FusedState must fit the head of the filter-prune workspace");

extern "C" int mc_filter_prune(const float* const* h_w_ptrs, const int* h_O, const int* h_C, const int* h_kh,
                               const int* h_kw, int nlayers, int64_t k, double gamma, float* d_values, double* d_thr,
                               float* const* h_mask_ptrs, uint8_t* d_keep, void* d_ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(h_w_ptrs && h_O && h_C && h_kh && h_kw && d_values && d_thr && d_ws, "mc_filter_prune: null pointer");
  MC_CHECK_ARG(nlayers > 0 && nlayers <= MAX_LAYERS, "mc_filter_prune: nlayers out of range");
  if (ws_bytes < mc_workspace_bytes_filter_prune()) return mc_set_error(MC_ERR_WS, "mc_filter_prune: workspace too small");
  int taps[MAX_LAYERS];
  int max_o = 0;
  for (int l = 0; l < nlayers; ++l) {
    MC_CHECK_ARG(h_kh[l] == h_kw[l] && h_kh[l] >= 1 && h_kh[l] * h_kw[l] <= 49,
                 "mc_filter_prune: layer %d kernel %dx%d unsupported (square, <=7x7)", l, h_kh[l], h_kw[l]);
    MC_CHECK_ARG(h_w_ptrs[l] != nullptr, "mc_filter_prune: layer %d null weight", l);
    MC_CHECK_ARG(!h_mask_ptrs || h_mask_ptrs[l] != nullptr, "mc_filter_prune: null mask %d", l);
    taps[l] = h_kh[l] * h_kw[l];
    if (h_O[l] > max_o) max_o = h_O[l];
  }
  LayerTable lt;
  int rc = build_layers(&lt, h_w_ptrs, h_mask_ptrs, h_O, h_C, taps, nlayers, "mc_filter_prune");
  if (rc) return rc;
  const int n = lt.voff[nlayers];
  MC_CHECK_ARG(k >= 0 && k < n, "mc_filter_prune: rank out of range");
  MC_CHECK_ARG(gamma >= 0.0 && gamma < 1.0, "mc_filter_prune: gamma must be in [0,1)");
  {
    const char* e = mc_tune_env("MCB200_FILTER_FUSED");  // =0: the three-launch path (A/B)
    const bool fused_on = !(e && e[0] == '0');
    if (fused_on && n <= FU_MAX_N) {
      LayerTable lf;
      rc = build_layers(&lf, h_w_ptrs, h_mask_ptrs, h_O, h_C, taps, nlayers, "mc_filter_prune", FU_THREADS, FU_F);
      if (rc) return rc;
      static int max_blocks = -1;
      if (max_blocks < 0) {
        MC_CUDA(cudaFuncSetAttribute(filter_prune_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FU_SMEM));
        int per_sm = 0;
        MC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter_prune_fused_kernel, FU_THREADS, FU_SMEM));
        if (per_sm < 1) return mc_set_error(MC_ERR_SHAPE, "mc_filter_prune: fused kernel does not fit an SM");
        max_blocks = per_sm * mc_num_sms();
      }
      int n_tiled = lf.tblk[lf.ntiled];
      int n_items = n_tiled + lf.toff[lf.nlayers] / FU_THREADS;
      int n_big = 0;  // tiled items of the layers with >= 1024 input channels (torder is sorted by descending C)
      for (int i2 = 0; i2 < lf.ntiled && h_C[lf.torder[i2]] >= 1024; ++i2) n_big = lf.tblk[i2 + 1];
      int grid = n_items < max_blocks ? n_items : max_blocks;
      if (grid < 1) grid = 1;
      MC_CUDA(cudaMemsetAsync(d_ws, 0, sizeof(FusedState), stream));
      long long kk = (long long)k;
      FusedState* stp = reinterpret_cast<FusedState*>(d_ws);
      float* d_raw = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(d_ws) + 2048);  // raw sums [n]
      int fill = h_mask_ptrs ? 1 : 0;
      void* args[] = {&lf, &d_raw, &d_values, &n_tiled, &n_big, &n_items, &kk, &gamma, &d_thr, &d_keep, &stp, &fill};
      MC_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(filter_prune_fused_kernel), dim3((unsigned)grid),
                                          dim3(FU_THREADS), args, FU_SMEM, stream));
      return 0;
    }
  }
  const size_t fin_smem = (size_t)(n > max_o ? n : max_o) * sizeof(float);
  if (fin_smem > 200 * 1024) return mc_set_error(MC_ERR_SHAPE, "mc_filter_prune: %d filters exceed the single-block select", n);
  static size_t fin_attr = 0;
  if (fin_smem > fin_attr) {
    MC_CUDA(cudaFuncSetAttribute(filter_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    fin_attr = fin_smem;
  }
  MC_CUDA(cudaMemsetAsync(d_ws, 0, 8, stream));
  rc = launch_sumsq(lt, d_values, stream);
  if (rc) return rc;
  filter_finish_kernel<<<nlayers, FIN_THREADS, fin_smem, stream>>>(lt, d_values, (long long)k, gamma, d_thr, d_keep,
                                                                   reinterpret_cast<unsigned int*>(d_ws));
  MC_LAUNCH_CHECK("filter_finish_kernel");
  if (h_mask_ptrs) {
    // the fill kernel addresses layer l's filter o at mask[l] + o*per: per-filter element count in C, taps = 1
    int blocks = n;
    const int cap = mc_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    filter_mask_fill_kernel<<<blocks, 256, 0, stream>>>(lt, d_values, d_thr);
    MC_LAUNCH_CHECK("filter_mask_fill_kernel");
  }
  return 0;
}
