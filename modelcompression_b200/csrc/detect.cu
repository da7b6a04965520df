// detect.cu — region-layer decode, greedy NMS and detection-row compaction, one CTA per image.
//
// Replaces get_region_boxes (src/nets2_utils.py:141-234), nms / bbox_iou / bbox_ious (src/nets2_utils.py:236-259,
// 63-131) and the per-image row emission of the eval loop (src/predict.py:157-173).
//
// Decode contract: channel a*(5+nc)+f holds f = tx,ty,tw,th,to,cls...; candidates are emitted in
// the reference's Python-loop order (cy, cx, anchor); box = [xs/W, ys/H, ws/W, hs/H, conf, cls_max_conf, cls_max_id].
// Transcendentals are accurate expf (no fast-math) — the contract vs the reference is a few-ulp tolerance.  The
// arithmetic lives in decode_math.cuh, shared with the fused decode epilogue of the head convolution.
// NMS contract: bit-exact kept-index lists: every IoU operation is a separately rounded fp32
// op in the reference's order (no FMA), sort key is fl32(1-conf) ascending, ties by ascending candidate index.
//
// Two box-table forms are accepted by the NMS / compaction kernels:
//   compact (d_counts != NULL): rows [0, counts[b]) of image b are its candidates in list order (mc_decode_region);
//   dense   (d_counts == NULL): one row per SLOT (cy*W+cx)*A + a, element 7 = slot index for a candidate and -1 for a
//           non-candidate (the fused decode epilogue, MC_EPI_DECODE).  Slot order == list order, so ties break the
//           same way, and `keep` then holds slot indices.
#include "common.cuh"
#include "decode_math.cuh"

namespace {

constexpr int DEC_THREADS = 1024;

__global__ void __launch_bounds__(DEC_THREADS)
decode_region_kernel(const float* __restrict__ head, int H, int W, int A, int nc, const McAnchors anc, float thresh,
                     int only_objectness, float* __restrict__ boxes, float* __restrict__ cls_out,
                     int* __restrict__ counts) {
  __shared__ int s_warp[DEC_THREADS / 32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int HW = H * W;
  const int P = HW * A;
  const int stride_f = HW;  // channel stride
  const float* hb = head + (long long)b * A * (5 + nc) * HW;
  float* bb = boxes + (long long)b * P * 8;
  float* cb = cls_out ? cls_out + (long long)b * P * nc : nullptr;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();

  for (int p0 = 0; p0 < P; p0 += DEC_THREADS) {
    const int pos = p0 + threadIdx.x;
    McDecoded d;
    d.cand = false;
    const float* src = nullptr;
    if (pos < P) {
      const int a = pos % A;
      const int cell = pos / A;
      const int cy = cell / W, cx = cell - cy * W;
      src = hb + (long long)a * (5 + nc) * stride_f + cell;
      const float* s = src;
      d = mc_decode_anchor([s, stride_f](int f) { return s[f * stride_f]; }, nc, cx, cy, W, H, anc.w[a], anc.h[a],
                           thresh, only_objectness);
    }
    // order-preserving compaction
    const unsigned int bal = __ballot_sync(0xffffffffu, d.cand);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int base = s_base;
    if (d.cand) {
      const int o = base + before + __popc(bal & ((1u << lane) - 1));
      float* dst = bb + (long long)o * 8;
      reinterpret_cast<float4*>(dst)[0] = make_float4(d.bx, d.by, d.bw, d.bh);
      reinterpret_cast<float4*>(dst)[1] = make_float4(d.conf, d.cmax, (float)d.cid, (float)pos);
      if (cb) {
        float* cd = cb + (long long)o * nc;
        for (int c = 0; c < nc; ++c)
          cd[c] = __fdiv_rn(expf(__fsub_rn(src[(5 + c) * stride_f], d.cls_max_logit)), d.cls_sum);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < DEC_THREADS / 32; ++w) tot += s_warp[w];
      s_base = base + tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = s_base;
}

// ------------------------------------------------------------------------------------------------------------
// NMS — chunked greedy, exactly the reference's sequential semantics
//   for i in sorted order: if conf_i > 0: keep i; for j > i: if IoU(i,j) > thr: conf_j = 0
// Thread t owns sorted position t (geometry in registers), warp c owns chunk c (positions 32c..32c+31).
//   (1) every warp builds the 32x32 suppression bit-matrix of its own chunk (all warps in parallel);
//   (2) chunks are resolved in order: the chunk's alive bits (initial conf > 0, minus what earlier kept boxes
//       suppressed) run through a <= 32-step bit loop -> kept bits; every LATER position then tests itself against the
//       kept boxes of the chunk (geometry broadcast from shared memory).  One block barrier per chunk instead of one
//       per kept box; a warp whose 32 positions are all dead skips the tests.
// The IoU arithmetic is the reference's (bbox_iou, centre format), operation by operation.
// ------------------------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 1024;
constexpr int NMS_CHUNKS = NMS_THREADS / 32;

struct NmsBox {
  float l, r, t, b, w, h, area;
};

__device__ __forceinline__ bool nms_suppresses(const NmsBox& a, const NmsBox& c, float thr) {
  const float mx = fminf(a.l, c.l);
  const float Mx = fmaxf(a.r, c.r);
  const float my = fminf(a.t, c.t);
  const float My = fmaxf(a.b, c.b);
  const float uw = __fsub_rn(Mx, mx);
  const float uh = __fsub_rn(My, my);
  const float cw = __fsub_rn(__fadd_rn(a.w, c.w), uw);
  const float ch = __fsub_rn(__fadd_rn(a.h, c.h), uh);
  if (cw <= 0.f || ch <= 0.f) return 0.0f > thr;
  const float carea = __fmul_rn(cw, ch);
  const float uarea = __fsub_rn(__fadd_rn(a.area, c.area), carea);
  return __fdiv_rn(carea, uarea) > thr;
}

// rows a kept box contributes to the detection table (src/predict.py:160-173 / nets2_utils.py:223-228): one, plus in
// validation mode one per other class c with conf * cls[c] > thresh.
__device__ __forceinline__ int detection_rows_of(const float* __restrict__ box, const float* __restrict__ cls, int nc,
                                                 float thresh) {
  if (cls == nullptr) return 1;
  const float conf = box[4];
  const int cid = (int)box[6];
  int n = 1;
  for (int c = 0; c < nc; ++c)
    if (c != cid && __fmul_rn(conf, cls[c]) > thresh) ++n;
  return n;
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(float* __restrict__ boxes, const int* __restrict__ counts, int cap, float thr, int* __restrict__ keep,
           int* __restrict__ keep_counts, const float* __restrict__ cls, int nc, float row_thresh,
           int* __restrict__ row_counts, int* __restrict__ cand_counts) {
  __shared__ unsigned long long s_key[NMS_THREADS];
  __shared__ float4 s_g0[NMS_THREADS];  // l, r, t, b   (sorted order)
  __shared__ float4 s_g1[NMS_THREADS];  // w, h, area, -
  __shared__ unsigned int s_intra[NMS_THREADS];  // row i of its chunk's 32x32 suppression matrix (bits j > i)
  __shared__ unsigned int s_alive[NMS_CHUNKS];
  __shared__ unsigned int s_kept[NMS_CHUNKS];
  __shared__ int s_warp[NMS_CHUNKS];
  __shared__ int s_rows[NMS_CHUNKS];
  __shared__ int s_n;

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* bb = boxes + (long long)img * cap * 8;
  int* kp = keep + (long long)img * cap;
  const bool dense = counts == nullptr;

  // ---- load: thread t reads table row t; key = fl32(1 - conf) ascending, ties by row index (= list order)
  int n_rows = dense ? cap : counts[img];
  if (n_rows > cap) n_rows = cap;
  unsigned long long key = ~0ull;
  bool valid = false;
  if (tid < n_rows) {
    const float4 q = reinterpret_cast<const float4*>(bb + (long long)tid * 8)[1];  // conf, cmax, cid, pos
    valid = !dense || q.w >= 0.f;
    if (valid) {
      // torch.sort ascending on float: order-preserving uint transform of fl32(1-conf) (handles negatives too)
      unsigned int kb = __float_as_uint(__fsub_rn(1.0f, q.x));
      kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
      key = ((unsigned long long)kb << 32) | (unsigned int)tid;
    }
  }
  {
    const unsigned int bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_warp[wid] = __popc(bal);
  }

  // ---- bitonic sort of (key, index) over the 1024 threads — the index in the low word makes the order total, i.e.
  //      stable.  Partners inside a warp exchange through shuffles, partners in other warps through shared memory.
  for (int k = 2; k <= NMS_THREADS; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long other;
      if (j >= 32) {
        s_key[tid] = key;
        __syncthreads();
        other = s_key[tid ^ j];
        __syncthreads();
      } else {
        other = __shfl_xor_sync(0xffffffffu, key, j);
      }
      const bool up = (tid & k) == 0;
      const bool lower = (tid & j) == 0;
      const bool take = lower ? ((key > other) == up) : ((other > key) == up);
      if (take) key = other;
    }
  }
  s_key[tid] = key;
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int w = 0; w < NMS_CHUNKS; ++w) tot += s_warp[w];
    s_n = tot;
  }
  __syncthreads();
  const int n = s_n;

  // ---- geometry of sorted position tid, in registers and in shared memory
  NmsBox me = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int my_row = 0;
  bool alive = false;
  if (tid < n) {
    my_row = (int)(key & 0xffffffffu);
    const float4 g = reinterpret_cast<const float4*>(bb + (long long)my_row * 8)[0];
    const float conf = bb[(long long)my_row * 8 + 4];
    const float hw = __fdiv_rn(g.z, 2.0f), hh = __fdiv_rn(g.w, 2.0f);
    me.l = __fsub_rn(g.x, hw);
    me.r = __fadd_rn(g.x, hw);
    me.t = __fsub_rn(g.y, hh);
    me.b = __fadd_rn(g.y, hh);
    me.w = g.z;
    me.h = g.w;
    me.area = __fmul_rn(g.z, g.w);
    alive = conf > 0.f;
  }
  s_g0[tid] = make_float4(me.l, me.r, me.t, me.b);
  s_g1[tid] = make_float4(me.w, me.h, me.area, 0.f);

  // ---- (1) intra-chunk suppression rows: lane j tests itself against lane i for every i < j
  {
    unsigned int row_bits = 0;  // lane i ends up with row i: bit j set iff j > i and i suppresses j
    for (int i = 0; i < 31; ++i) {
      NmsBox o;
      o.l = __shfl_sync(0xffffffffu, me.l, i);
      o.r = __shfl_sync(0xffffffffu, me.r, i);
      o.t = __shfl_sync(0xffffffffu, me.t, i);
      o.b = __shfl_sync(0xffffffffu, me.b, i);
      o.w = __shfl_sync(0xffffffffu, me.w, i);
      o.h = __shfl_sync(0xffffffffu, me.h, i);
      o.area = __shfl_sync(0xffffffffu, me.area, i);
      const bool sup = lane > i && tid < n && nms_suppresses(o, me, thr);
      const unsigned int bal = __ballot_sync(0xffffffffu, sup);
      if (lane == i) row_bits = bal;
    }
    s_intra[tid] = row_bits;
    const unsigned int ab = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) s_alive[wid] = ab;
  }
  __syncthreads();

  // ---- (2) chunks in order
  const int n_chunks = (n + 31) >> 5;
  for (int c = 0; c < n_chunks; ++c) {
    // resolve chunk c (every warp runs the same uniform bit loop on the same shared words: no broadcast needed)
    unsigned int al = s_alive[c];
    unsigned int kept = 0;
    while (al) {
      const int i = __ffs(al) - 1;
      kept |= 1u << i;
      al &= ~(1u << i);
      al &= ~s_intra[c * 32 + i];
    }
    if (tid == 0) s_kept[c] = kept;
    // later positions test themselves against the kept boxes of the chunk
    if (wid > c && __any_sync(0xffffffffu, alive)) {
      unsigned int kb = kept;
      while (kb) {
        const int i = __ffs(kb) - 1;
        kb &= kb - 1;
        const float4 a0 = s_g0[c * 32 + i], a1 = s_g1[c * 32 + i];
        const NmsBox o = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z};
        if (alive && nms_suppresses(o, me, thr)) alive = false;
      }
      const unsigned int ab = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) s_alive[wid] = ab;
    }
    __syncthreads();  // s_alive of the next chunk is final
  }

  // ---- outputs: kept rows in sorted order, mutated confidences, counts
  const bool is_kept = tid < n && ((s_kept[wid] >> lane) & 1u);
  const unsigned int kbal = __ballot_sync(0xffffffffu, is_kept);
  int rows = 0;
  if (row_counts != nullptr && is_kept)
    rows = detection_rows_of(bb + (long long)my_row * 8, cls ? cls + ((long long)img * cap + my_row) * nc : nullptr, nc,
                             row_thresh);
  for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
  if (lane == 0) {
    s_warp[wid] = __popc(kbal);
    s_rows[wid] = rows;
  }
  __syncthreads();
  int before = 0;
  for (int w = 0; w < wid; ++w) before += s_warp[w];
  if (is_kept) kp[before + __popc(kbal & ((1u << lane) - 1))] = my_row;
  // the reference sets box_j[4] = 0 on every suppressed box (nets2_utils.py:258)
  if (tid < n && !is_kept && bb[(long long)my_row * 8 + 4] > 0.f) bb[(long long)my_row * 8 + 4] = 0.f;
  if (tid == 0) {
    int tot = 0, r = 0;
    for (int w = 0; w < NMS_CHUNKS; ++w) {
      tot += s_warp[w];
      r += s_rows[w];
    }
    keep_counts[img] = tot;
    if (cand_counts != nullptr) cand_counts[img] = n;
    if (row_counts != nullptr) row_counts[img] = r;
  }
}


// More than 1024 boxes per image (e.g. a 608x608 input: 19x19x5 = 1805): one greedy step per kept box, any capacity
// that fits shared memory.  Same semantics, tables and outputs as nms_kernel.
__global__ void __launch_bounds__(NMS_THREADS)
nms_large_kernel(float* __restrict__ boxes, const int* __restrict__ counts, int cap, int cap_pow2, float thr,
                 int* __restrict__ keep, int* __restrict__ keep_counts, const float* __restrict__ cls, int nc,
                 float row_thresh, int* __restrict__ row_counts, int* __restrict__ cand_counts) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(nms_smem);  // [cap_pow2]
  float* s_l = reinterpret_cast<float*>(s_key + cap_pow2);                      // per table row
  float* s_r = s_l + cap;
  float* s_t = s_r + cap;
  float* s_b = s_t + cap;
  float* s_area = s_b + cap;
  float* s_w = s_area + cap;
  float* s_h = s_w + cap;
  float* s_conf = s_h + cap;
  __shared__ int s_nkeep, s_n, s_nrows;

  const int img = blockIdx.x;
  const bool dense = counts == nullptr;
  int n_rows = dense ? cap : counts[img];
  if (n_rows > cap) n_rows = cap;
  float* bb = boxes + (long long)img * cap * 8;
  int* kp = keep + (long long)img * cap;
  if (threadIdx.x == 0) { s_nkeep = 0; s_n = 0; s_nrows = 0; }
  __syncthreads();

  for (int i = threadIdx.x; i < cap_pow2; i += NMS_THREADS) {
    unsigned long long key = ~0ull;
    if (i < n_rows) {
      const float4 g = reinterpret_cast<const float4*>(bb + (long long)i * 8)[0];
      const float4 q = reinterpret_cast<const float4*>(bb + (long long)i * 8)[1];
      const float conf = q.x;
      const float hw = __fdiv_rn(g.z, 2.0f), hh = __fdiv_rn(g.w, 2.0f);
      s_l[i] = __fsub_rn(g.x, hw);
      s_r[i] = __fadd_rn(g.x, hw);
      s_t[i] = __fsub_rn(g.y, hh);
      s_b[i] = __fadd_rn(g.y, hh);
      s_w[i] = g.z;
      s_h[i] = g.w;
      s_area[i] = __fmul_rn(g.z, g.w);
      s_conf[i] = conf;
      if (!dense || q.w >= 0.f) {
        unsigned int kb = __float_as_uint(__fsub_rn(1.0f, conf));
        kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
        key = ((unsigned long long)kb << 32) | (unsigned int)i;
        atomicAdd(&s_n, 1);
      }
    }
    s_key[i] = key;
  }
  __syncthreads();
  const int n = s_n;

  // bitonic sort of (key, index) — index in the low word makes the order total => stable; invalid rows sort last
  for (int k = 2; k <= cap_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < cap_pow2; i += NMS_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s_key[i], c = s_key[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { s_key[i] = c; s_key[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }

  // greedy suppression in sorted order
  for (int i = 0; i < n; ++i) {
    const int bi = (int)(s_key[i] & 0xffffffffu);
    if (!(s_conf[bi] > 0.f)) continue;  // uniform: s_conf only changes behind a barrier
    if (threadIdx.x == 0) {
      kp[s_nkeep] = bi;
      s_nkeep = s_nkeep + 1;
      if (row_counts != nullptr)
        s_nrows += detection_rows_of(bb + (long long)bi * 8, cls ? cls + ((long long)img * cap + bi) * nc : nullptr, nc,
                                     row_thresh);
    }
    const NmsBox a = {s_l[bi], s_r[bi], s_t[bi], s_b[bi], s_w[bi], s_h[bi], s_area[bi]};
    for (int j = i + 1 + threadIdx.x; j < n; j += NMS_THREADS) {
      const int bj = (int)(s_key[j] & 0xffffffffu);
      const NmsBox c = {s_l[bj], s_r[bj], s_t[bj], s_b[bj], s_w[bj], s_h[bj], s_area[bj]};
      if (nms_suppresses(a, c, thr)) s_conf[bj] = 0.f;
    }
    __syncthreads();
  }
  __syncthreads();
  // write back mutated confidences (reference sets box_j[4] = 0) and the counts
  for (int i = threadIdx.x; i < n_rows; i += NMS_THREADS)
    if (s_conf[i] == 0.f && bb[(long long)i * 8 + 4] > 0.f && (!dense || bb[(long long)i * 8 + 7] >= 0.f))
      bb[(long long)i * 8 + 4] = 0.f;
  if (threadIdx.x == 0) {
    keep_counts[img] = s_nkeep;
    if (cand_counts != nullptr) cand_counts[img] = n;
    if (row_counts != nullptr) row_counts[img] = s_nrows;
  }
}

// Detection rows of one image: [img, x, y, w, h, box_conf, cls_conf, cls_id] per kept box (NMS order), in validation
// mode the arg-max class first, then every other class with conf*cls[c] > thresh in ascending class order
// (nets2_utils.py:223-228 as consumed by src/predict.py:167-172).  Rows of image b start at row_offsets[b].
constexpr int CMP_THREADS = 256;

__global__ void __launch_bounds__(CMP_THREADS)
compact_detections_kernel(const float* __restrict__ boxes, const int* __restrict__ keep,
                          const int* __restrict__ keep_counts, const float* __restrict__ cls, int cap, int nc,
                          float thresh, int first_image, const long long* __restrict__ row_offsets,
                          float* __restrict__ out) {
  __shared__ int s_warp[CMP_THREADS / 32];
  __shared__ int s_base;
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* bb = boxes + (long long)img * cap * 8;
  const int* kp = keep + (long long)img * cap;
  const int nk = keep_counts[img];
  float* dst0 = out + row_offsets[img] * 8;
  const float fimg = (float)(first_image + img);
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < nk; k0 += CMP_THREADS) {
    const int k = k0 + tid;
    int rows = 0, row = 0;
    const float* cl = nullptr;
    if (k < nk) {
      row = kp[k];
      cl = cls ? cls + ((long long)img * cap + row) * nc : nullptr;
      rows = detection_rows_of(bb + (long long)row * 8, cl, nc, thresh);
    }
    int incl = rows;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int base = s_base;
    if (k < nk) {
      const float4 g = reinterpret_cast<const float4*>(bb + (long long)row * 8)[0];
      const float4 q = reinterpret_cast<const float4*>(bb + (long long)row * 8)[1];
      float* d = dst0 + (long long)(base + before + incl - rows) * 8;
      reinterpret_cast<float4*>(d)[0] = make_float4(fimg, g.x, g.y, g.z);
      reinterpret_cast<float4*>(d)[1] = make_float4(g.w, q.x, q.y, q.z);
      if (cl != nullptr) {
        const int cid = (int)q.z;
        for (int c = 0; c < nc; ++c) {
          if (c == cid || !(__fmul_rn(q.x, cl[c]) > thresh)) continue;
          d += 8;
          reinterpret_cast<float4*>(d)[0] = make_float4(fimg, g.x, g.y, g.z);
          reinterpret_cast<float4*>(d)[1] = make_float4(g.w, q.x, cl[c], (float)c);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < CMP_THREADS / 32; ++w) tot += s_warp[w];
      s_base = base + tot;
    }
    __syncthreads();
  }
}

// bbox_ious (src/nets2_utils.py:100-131): element-wise IoU of two [4, n] box sets, every operation a separately rounded
// fp32 op in the reference's order (torch element-wise kernels do not contract to FMA), `carea[mask] = 0` included.
__global__ void bbox_ious_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int corners,
                                 float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i], a3 = a[3 * n + i];
    const float b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i], b3 = b[3 * n + i];
    float mx, Mx, my, My, w1, h1, w2, h2;
    if (corners) {
      mx = fminf(a0, b0); Mx = fmaxf(a2, b2); my = fminf(a1, b1); My = fmaxf(a3, b3);
      w1 = __fsub_rn(a2, a0); h1 = __fsub_rn(a3, a1); w2 = __fsub_rn(b2, b0); h2 = __fsub_rn(b3, b1);
    } else {
      mx = fminf(__fsub_rn(a0, __fdiv_rn(a2, 2.0f)), __fsub_rn(b0, __fdiv_rn(b2, 2.0f)));
      Mx = fmaxf(__fadd_rn(a0, __fdiv_rn(a2, 2.0f)), __fadd_rn(b0, __fdiv_rn(b2, 2.0f)));
      my = fminf(__fsub_rn(a1, __fdiv_rn(a3, 2.0f)), __fsub_rn(b1, __fdiv_rn(b3, 2.0f)));
      My = fmaxf(__fadd_rn(a1, __fdiv_rn(a3, 2.0f)), __fadd_rn(b1, __fdiv_rn(b3, 2.0f)));
      w1 = a2; h1 = a3; w2 = b2; h2 = b3;
    }
    const float uw = __fsub_rn(Mx, mx), uh = __fsub_rn(My, my);
    const float cw = __fsub_rn(__fadd_rn(w1, w2), uw), ch = __fsub_rn(__fadd_rn(h1, h2), uh);
    float carea = __fmul_rn(cw, ch);
    if (cw <= 0.f || ch <= 0.f) carea = 0.f;
    const float uarea = __fsub_rn(__fadd_rn(__fmul_rn(w1, h1), __fmul_rn(w2, h2)), carea);
    out[i] = __fdiv_rn(carea, uarea);
  }
}

}  // namespace

extern "C" int mc_bbox_ious(const float* d_boxes1, const float* d_boxes2, int64_t n, int x1y1x2y2, float* d_out,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_boxes1 && d_boxes2 && d_out, "mc_bbox_ious: null pointer");
  MC_CHECK_ARG(n >= 0, "mc_bbox_ious: bad count");
  if (n == 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)mc_num_sms() * 8) blocks = (long long)mc_num_sms() * 8;
  bbox_ious_kernel<<<(int)blocks, 256, 0, stream>>>(d_boxes1, d_boxes2, (long long)n, x1y1x2y2 ? 1 : 0, d_out);
  MC_LAUNCH_CHECK("bbox_ious_kernel");
  return 0;
}

extern "C" int mc_decode_region(const float* d_head, int B, int H, int W, int A, int nc, const float* h_anchors,
                                float conf_thresh, int only_objectness, float* d_boxes, float* d_cls, int* d_counts,
                                void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_head && h_anchors && d_boxes && d_counts, "mc_decode_region: null pointer");
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0 && A > 0 && A <= MC_MAX_ANCHORS && nc > 0, "mc_decode_region: bad dims");
  MC_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0, "mc_decode_region: d_boxes must be 16-byte aligned");
  McAnchors anc;
  for (int a = 0; a < MC_MAX_ANCHORS; ++a) {
    anc.w[a] = a < A ? h_anchors[2 * a] : 0.f;
    anc.h[a] = a < A ? h_anchors[2 * a + 1] : 0.f;
  }
  decode_region_kernel<<<B, DEC_THREADS, 0, stream>>>(d_head, H, W, A, nc, anc, conf_thresh, only_objectness, d_boxes,
                                                      d_cls, d_counts);
  MC_LAUNCH_CHECK("decode_region_kernel");
  return 0;
}

static int nms_launch(float* d_boxes, const int* d_counts, int B, int cap, float nms_thresh, int* d_keep,
                      int* d_keep_counts, const float* d_cls, int nc, float row_thresh, int* d_row_counts,
                      int* d_cand_counts, cudaStream_t stream, const char* who) {
  MC_CHECK_ARG(d_boxes && d_keep && d_keep_counts, "%s: null pointer", who);
  MC_CHECK_ARG(B > 0 && cap > 0, "%s: bad dims", who);
  MC_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0, "%s: d_boxes must be 16-byte aligned", who);
  if (cap > NMS_THREADS) {
    int p2 = 1;
    while (p2 < cap) p2 <<= 1;
    const size_t smem = (size_t)p2 * 8 + (size_t)cap * 8 * sizeof(float);
    if (smem > 220 * 1024) return mc_set_error(MC_ERR_SHAPE, "%s: cap %d needs %zu B of shared memory", who, cap, smem);
    static size_t max_set = 0;
    if (smem > 48 * 1024 && smem > max_set) {
      MC_CUDA(cudaFuncSetAttribute(nms_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      max_set = smem;
    }
    nms_large_kernel<<<B, NMS_THREADS, smem, stream>>>(d_boxes, d_counts, cap, p2, nms_thresh, d_keep, d_keep_counts,
                                                       d_cls, nc, row_thresh, d_row_counts, d_cand_counts);
    MC_LAUNCH_CHECK("nms_large_kernel");
    return 0;
  }
  nms_kernel<<<B, NMS_THREADS, 0, stream>>>(d_boxes, d_counts, cap, nms_thresh, d_keep, d_keep_counts, d_cls, nc,
                                            row_thresh, d_row_counts, d_cand_counts);
  MC_LAUNCH_CHECK("nms_kernel");
  return 0;
}

extern "C" int mc_nms_batched(float* d_boxes, const int* d_counts, int B, int cap, float nms_thresh, int* d_keep,
                              int* d_keep_counts, void* stream_) {
  MC_CHECK_ARG(d_counts != nullptr, "mc_nms_batched: null pointer");
  return nms_launch(d_boxes, d_counts, B, cap, nms_thresh, d_keep, d_keep_counts, nullptr, 0, 0.f, nullptr, nullptr,
                    reinterpret_cast<cudaStream_t>(stream_), "mc_nms_batched");
}

extern "C" int mc_nms_detect(float* d_boxes, const int* d_counts, int B, int cap, float nms_thresh, int* d_keep,
                             int* d_keep_counts, const float* d_cls, int nc, float conf_thresh, int* d_row_counts,
                             int* d_cand_counts, void* stream_) {
  MC_CHECK_ARG(d_cls == nullptr || nc > 0, "mc_nms_detect: class probabilities without a class count");
  return nms_launch(d_boxes, d_counts, B, cap, nms_thresh, d_keep, d_keep_counts, d_cls, nc, conf_thresh, d_row_counts,
                    d_cand_counts, reinterpret_cast<cudaStream_t>(stream_), "mc_nms_detect");
}

extern "C" int mc_compact_detections(const float* d_boxes, const int* d_keep, const int* d_keep_counts,
                                     const float* d_cls, int B, int cap, int nc, float conf_thresh, int first_image,
                                     const int64_t* d_row_offsets, float* d_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_boxes && d_keep && d_keep_counts && d_row_offsets && d_out, "mc_compact_detections: null pointer");
  MC_CHECK_ARG(B > 0 && cap > 0 && (d_cls == nullptr || nc > 0), "mc_compact_detections: bad dims");
  MC_CHECK_ARG((((uintptr_t)d_boxes | (uintptr_t)d_out) & 15) == 0, "mc_compact_detections: tables must be 16-byte aligned");
  compact_detections_kernel<<<B, CMP_THREADS, 0, stream>>>(d_boxes, d_keep, d_keep_counts, d_cls, cap, nc, conf_thresh,
                                                           first_image,
                                                           reinterpret_cast<const long long*>(d_row_offsets), d_out);
  MC_LAUNCH_CHECK("compact_detections_kernel");
  return 0;
}
