// detect.cu — region-layer decode and greedy NMS, one CTA per image.
//
// Replaces get_region_boxes (src/nets2_utils.py:141-234) and nms / bbox_iou (src/nets2_utils.py:236-259, 63-98).
//
// Decode contract: channel a*(5+nc)+f holds f = tx,ty,tw,th,to,cls...; candidates are emitted in
// the reference's Python-loop order (cy, cx, anchor); box = [xs/W, ys/H, ws/W, hs/H, conf, cls_max_conf, cls_max_id].
// Transcendentals are accurate expf (no fast-math) — the contract vs the reference is a few-ulp tolerance.
// NMS contract: bit-exact kept-index lists: every IoU operation is a separately rounded fp32
// op in the reference's order (no FMA), sort key is fl32(1-conf) ascending, ties by ascending candidate index.
#include "common.cuh"

namespace {

constexpr int DEC_THREADS = 1024;
constexpr int MAX_ANCHORS = 16;

struct Anchors {
  float w[MAX_ANCHORS];
  float h[MAX_ANCHORS];
};

__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(DEC_THREADS)
decode_region_kernel(const float* __restrict__ head, int H, int W, int A, int nc, const Anchors anc, float thresh,
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
    bool cand = false;
    float bx = 0, by = 0, bw = 0, bh = 0, conf = 0, cmax = 0;
    int cid = 0;
    const float* src = nullptr;
    float cls_max_logit = 0.f, cls_sum = 1.f;
    if (pos < P) {
      const int a = pos % A;
      const int cell = pos / A;
      const int cy = cell / W, cx = cell - cy * W;
      src = hb + (long long)a * (5 + nc) * stride_f + cell;
      const float tx = src[0], ty = src[stride_f], tw = src[2 * stride_f], th = src[3 * stride_f],
                  to = src[4 * stride_f];
      conf = sigmoid_ref(to);
      // softmax over classes: max, then exp(x-max)/sum
      float mx = -INFINITY;
      for (int c = 0; c < nc; ++c) mx = fmaxf(mx, src[(5 + c) * stride_f]);
      float sum = 0.f;
      for (int c = 0; c < nc; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(src[(5 + c) * stride_f], mx)));
      cls_max_logit = mx;
      cls_sum = sum;
      cmax = -1.f;
      for (int c = 0; c < nc; ++c) {
        const float pc = __fdiv_rn(expf(__fsub_rn(src[(5 + c) * stride_f], mx)), sum);
        if (pc > cmax) { cmax = pc; cid = c; }
      }
      const float score = only_objectness ? conf : __fmul_rn(conf, cmax);
      cand = score > thresh;
      if (cand) {
        bx = __fdiv_rn(__fadd_rn(sigmoid_ref(tx), (float)cx), (float)W);
        by = __fdiv_rn(__fadd_rn(sigmoid_ref(ty), (float)cy), (float)H);
        bw = __fdiv_rn(__fmul_rn(expf(tw), anc.w[a]), (float)W);
        bh = __fdiv_rn(__fmul_rn(expf(th), anc.h[a]), (float)H);
      }
    }
    // order-preserving compaction
    const unsigned int bal = __ballot_sync(0xffffffffu, cand);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < wid; ++w) before += s_warp[w];
    const int base = s_base;
    if (cand) {
      const int o = base + before + __popc(bal & ((1u << lane) - 1));
      float* dst = bb + (long long)o * 8;
      reinterpret_cast<float4*>(dst)[0] = make_float4(bx, by, bw, bh);
      reinterpret_cast<float4*>(dst)[1] = make_float4(conf, cmax, (float)cid, (float)pos);
      if (cb) {
        float* cd = cb + (long long)o * nc;
        for (int c = 0; c < nc; ++c)
          cd[c] = __fdiv_rn(expf(__fsub_rn(src[(5 + c) * stride_f], cls_max_logit)), cls_sum);
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
// NMS
// ------------------------------------------------------------------------------------------------------------
constexpr int NMS_THREADS = 1024;

__global__ void __launch_bounds__(NMS_THREADS)
nms_kernel(float* __restrict__ boxes, const int* __restrict__ counts, int cap, int cap_pow2, float thr,
           int* __restrict__ keep, int* __restrict__ keep_counts) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* s_key = reinterpret_cast<unsigned long long*>(nms_smem);  // [cap_pow2]
  float* s_l = reinterpret_cast<float*>(s_key + cap_pow2);                      // left, right, top, bottom, area, conf
  float* s_r = s_l + cap;
  float* s_t = s_r + cap;
  float* s_b = s_t + cap;
  float* s_area = s_b + cap;
  float* s_w = s_area + cap;
  float* s_h = s_w + cap;
  float* s_conf = s_h + cap;
  __shared__ int s_nkeep;

  const int img = blockIdx.x;
  int n = counts[img];
  if (n > cap) n = cap;
  float* bb = boxes + (long long)img * cap * 8;
  int* kp = keep + (long long)img * cap;

  for (int i = threadIdx.x; i < cap_pow2; i += NMS_THREADS) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const float4 g = reinterpret_cast<const float4*>(bb + (long long)i * 8)[0];
      const float conf = bb[(long long)i * 8 + 4];
      const float hw = __fdiv_rn(g.z, 2.0f), hh = __fdiv_rn(g.w, 2.0f);
      s_l[i] = __fsub_rn(g.x, hw);
      s_r[i] = __fadd_rn(g.x, hw);
      s_t[i] = __fsub_rn(g.y, hh);
      s_b[i] = __fadd_rn(g.y, hh);
      s_w[i] = g.z;
      s_h[i] = g.w;
      s_area[i] = __fmul_rn(g.z, g.w);
      s_conf[i] = conf;
      // torch.sort ascending on float: order-preserving uint transform of fl32(1-conf) (handles negatives too)
      unsigned int kb = __float_as_uint(__fsub_rn(1.0f, conf));
      kb = (kb & 0x80000000u) ? ~kb : (kb | 0x80000000u);
      key = ((unsigned long long)kb << 32) | (unsigned int)i;
    }
    s_key[i] = key;
  }
  if (threadIdx.x == 0) s_nkeep = 0;
  __syncthreads();

  // bitonic sort of (key, index) — index in the low word makes the order total => stable
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
    }
    const float l1 = s_l[bi], r1 = s_r[bi], t1 = s_t[bi], b1 = s_b[bi], w1 = s_w[bi], h1 = s_h[bi], a1 = s_area[bi];
    for (int j = i + 1 + threadIdx.x; j < n; j += NMS_THREADS) {
      const int bj = (int)(s_key[j] & 0xffffffffu);
      const float mx = fminf(l1, s_l[bj]);
      const float Mx = fmaxf(r1, s_r[bj]);
      const float my = fminf(t1, s_t[bj]);
      const float My = fmaxf(b1, s_b[bj]);
      const float uw = __fsub_rn(Mx, mx);
      const float uh = __fsub_rn(My, my);
      const float cw = __fsub_rn(__fadd_rn(w1, s_w[bj]), uw);
      const float ch = __fsub_rn(__fadd_rn(h1, s_h[bj]), uh);
      bool sup;
      if (cw <= 0.f || ch <= 0.f) {
        sup = 0.0f > thr;
      } else {
        const float carea = __fmul_rn(cw, ch);
        const float uarea = __fsub_rn(__fadd_rn(a1, s_area[bj]), carea);
        sup = __fdiv_rn(carea, uarea) > thr;
      }
      if (sup) s_conf[bj] = 0.f;
    }
    __syncthreads();
  }
  __syncthreads();
  // write back mutated confidences (reference sets box_j[4] = 0) and the kept count
  for (int i = threadIdx.x; i < n; i += NMS_THREADS)
    if (s_conf[i] == 0.f) bb[(long long)i * 8 + 4] = 0.f;
  if (threadIdx.x == 0) keep_counts[img] = s_nkeep;
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
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0 && A > 0 && A <= MAX_ANCHORS && nc > 0, "mc_decode_region: bad dims");
  MC_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0, "mc_decode_region: d_boxes must be 16-byte aligned");
  Anchors anc;
  for (int a = 0; a < MAX_ANCHORS; ++a) {
    anc.w[a] = a < A ? h_anchors[2 * a] : 0.f;
    anc.h[a] = a < A ? h_anchors[2 * a + 1] : 0.f;
  }
  decode_region_kernel<<<B, DEC_THREADS, 0, stream>>>(d_head, H, W, A, nc, anc, conf_thresh, only_objectness, d_boxes,
                                                      d_cls, d_counts);
  MC_LAUNCH_CHECK("decode_region_kernel");
  return 0;
}

extern "C" int mc_nms_batched(float* d_boxes, const int* d_counts, int B, int cap, float nms_thresh, int* d_keep,
                              int* d_keep_counts, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_boxes && d_counts && d_keep && d_keep_counts, "mc_nms_batched: null pointer");
  MC_CHECK_ARG(B > 0 && cap > 0, "mc_nms_batched: bad dims");
  MC_CHECK_ARG(((uintptr_t)d_boxes & 15) == 0, "mc_nms_batched: d_boxes must be 16-byte aligned");
  int p2 = 1;
  while (p2 < cap) p2 <<= 1;
  const size_t smem = (size_t)p2 * 8 + (size_t)cap * 8 * sizeof(float);
  if (smem > 220 * 1024) return mc_set_error(MC_ERR_SHAPE, "mc_nms_batched: cap %d needs %zu B of shared memory", cap, smem);
  static size_t max_set = 0;
  if (smem > 48 * 1024 && smem > max_set) {
    MC_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    max_set = smem;
  }
  nms_kernel<<<B, NMS_THREADS, smem, stream>>>(d_boxes, d_counts, cap, p2, nms_thresh, d_keep, d_keep_counts);
  MC_LAUNCH_CHECK("nms_kernel");
  return 0;
}
