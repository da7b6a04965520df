// voc_eval.cu — PASCAL-VOC detection matching on the GPU.
//
// Replaces the file-based scorer of the reference: the per-class '%f' text files written by the eval loop
// (src/predict.py:157-173) and the per-detection Python loop of voc_eval (src/predict.py:305-380).  Two steps:
//   voc_table_kernel   per detection row (image, x, y, w, h, box_conf, cls_conf, cls): the float32 corner / score
//                      arithmetic of :160-171 followed by the '%f' text round trip (6 decimals, read back as float64);
//   voc_match_kernel   per detection, in (class, descending score) order: overlaps with the ground-truth boxes of its
//                      image and class in float64 with the reference's +1 pixel convention, first maximum (np.argmax);
//                      a hit on a non-difficult box competes for it with an atomicMin of the sorted position — the
//                      reference's sequential "first detection to reach a box is the true positive" rule is
//                      order-dependent only per ground-truth box;
//   voc_mark_kernel    tp / fp flags from the winners.
// Sorting and the cumulative sums stay library calls (torch.sort / torch.cumsum) in voc_eval.py.
#include "common.cuh"

namespace {

__device__ __forceinline__ double text_round(double v) { return rint(v * 1e6) / 1e6; }  // round-half-even, like printf

__global__ void voc_table_kernel(const float* __restrict__ dets, long long n, const float* __restrict__ sizes, float def_w,
                                 float def_h, long long* __restrict__ img, long long* __restrict__ cls,
                                 double* __restrict__ conf, double* __restrict__ corners) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* d = dets + i * 8;
  const long long im = (long long)d[0];
  const float width = sizes ? sizes[im * 2] : def_w, height = sizes ? sizes[im * 2 + 1] : def_h;
  const float x = d[1], y = d[2], w = d[3], h = d[4];
  const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);  // w / 2.0, h / 2.0 in float32
  const float x1 = __fmul_rn(__fsub_rn(x, hw), width), y1 = __fmul_rn(__fsub_rn(y, hh), height);
  const float x2 = __fmul_rn(__fadd_rn(x, hw), width), y2 = __fmul_rn(__fadd_rn(y, hh), height);
  img[i] = im;
  cls[i] = (long long)d[7];
  conf[i] = text_round((double)__fmul_rn(d[5], d[6]));
  corners[i * 4 + 0] = text_round((double)x1);
  corners[i * 4 + 1] = text_round((double)y1);
  corners[i * 4 + 2] = text_round((double)x2);
  corners[i * 4 + 3] = text_round((double)y2);
}

__global__ void voc_fill_kernel(long long* __restrict__ first, long long m, long long v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) first[i] = v;
}

// key = class * K + image for detections (in sorted order) and ground-truth boxes (sorted ascending)
__global__ void voc_match_kernel(const long long* __restrict__ key, const double* __restrict__ corners, long long n,
                                 const long long* __restrict__ gkey, const double* __restrict__ gbox,
                                 const unsigned char* __restrict__ gdiff, long long m, double ovthresh,
                                 unsigned long long* __restrict__ first, long long* __restrict__ jmax_out,
                                 unsigned char* __restrict__ state) {
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const long long k = key[d];
  long long lo = 0, hi = m;  // first ground-truth box with gkey >= k
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (gkey[mid] < k) lo = mid + 1; else hi = mid;
  }
  const double bx1 = corners[d * 4], by1 = corners[d * 4 + 1], bx2 = corners[d * 4 + 2], by2 = corners[d * 4 + 3];
  const double barea = __dmul_rn(__dadd_rn(__dsub_rn(bx2, bx1), 1.0), __dadd_rn(__dsub_rn(by2, by1), 1.0));
  double ovmax = -INFINITY;
  long long jmax = -1;
  for (long long j = lo; j < m && gkey[j] == k; ++j) {
    const double gx1 = gbox[j * 4], gy1 = gbox[j * 4 + 1], gx2 = gbox[j * 4 + 2], gy2 = gbox[j * 4 + 3];
    const double ixmin = fmax(gx1, bx1), iymin = fmax(gy1, by1), ixmax = fmin(gx2, bx2), iymax = fmin(gy2, by2);
    const double iw = fmax(__dadd_rn(__dsub_rn(ixmax, ixmin), 1.0), 0.0);
    const double ih = fmax(__dadd_rn(__dsub_rn(iymax, iymin), 1.0), 0.0);
    const double inters = __dmul_rn(iw, ih);
    const double garea = __dmul_rn(__dadd_rn(__dsub_rn(gx2, gx1), 1.0), __dadd_rn(__dsub_rn(gy2, gy1), 1.0));
    const double uni = __dsub_rn(__dadd_rn(barea, garea), inters);
    const double ov = __ddiv_rn(inters, uni);
    if (ov > ovmax) { ovmax = ov; jmax = j; }  // first maximum
  }
  unsigned char st = 0;  // 0: no hit (false positive), 1: hit a difficult box (ignored), 2: candidate for box jmax
  if (jmax >= 0 && ovmax > ovthresh) {
    if (gdiff[jmax]) st = 1;
    else {
      st = 2;
      atomicMin(&first[jmax], (unsigned long long)d);
    }
  }
  jmax_out[d] = jmax;
  state[d] = st;
}

__global__ void voc_mark_kernel(const unsigned long long* __restrict__ first, const long long* __restrict__ jmax,
                                const unsigned char* __restrict__ state, long long n, double* __restrict__ tp,
                                double* __restrict__ fp) {
  const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const unsigned char st = state[d];
  const bool is_tp = st == 2 && first[jmax[d]] == (unsigned long long)d;
  tp[d] = is_tp ? 1.0 : 0.0;
  fp[d] = (st == 0 || (st == 2 && !is_tp)) ? 1.0 : 0.0;
}

}  // namespace

extern "C" int mc_voc_table(const float* d_dets, int64_t n, const float* d_sizes, float def_w, float def_h, int64_t* d_img,
                            int64_t* d_cls, double* d_conf, double* d_corners, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(n >= 0, "mc_voc_table: negative count");
  if (n == 0) return 0;
  MC_CHECK_ARG(d_dets && d_img && d_cls && d_conf && d_corners, "mc_voc_table: null pointer");
  voc_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_dets, (long long)n, d_sizes, def_w, def_h,
                                                                      reinterpret_cast<long long*>(d_img),
                                                                      reinterpret_cast<long long*>(d_cls), d_conf, d_corners);
  MC_LAUNCH_CHECK("voc_table_kernel");
  return 0;
}

extern "C" size_t mc_workspace_bytes_voc_match(int64_t n, int64_t m) {
  return (size_t)(m > 0 ? m : 1) * 8 + (size_t)(n > 0 ? n : 1) * 9 + 64;
}

extern "C" int mc_voc_match(const int64_t* d_key, const double* d_corners, int64_t n, const int64_t* d_gkey,
                            const double* d_gbox, const uint8_t* d_gdiff, int64_t m, double ovthresh, double* d_tp,
                            double* d_fp, void* d_ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(n >= 0 && m >= 0, "mc_voc_match: negative count");
  if (n == 0) return 0;
  MC_CHECK_ARG(d_key && d_corners && d_tp && d_fp && d_ws && (m == 0 || (d_gkey && d_gbox && d_gdiff)),
               "mc_voc_match: null pointer");
  if (ws_bytes < mc_workspace_bytes_voc_match(n, m)) return mc_set_error(MC_ERR_WS, "mc_voc_match: workspace too small");
  MC_CHECK_ARG(((uintptr_t)d_ws & 7) == 0, "mc_voc_match: workspace must be 8-byte aligned");
  unsigned long long* first = reinterpret_cast<unsigned long long*>(d_ws);                 // [m]
  long long* jmax = reinterpret_cast<long long*>(first + (m > 0 ? m : 1));                   // [n]
  unsigned char* state = reinterpret_cast<unsigned char*>(jmax + n);                         // [n]
  if (m > 0) {
    voc_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(reinterpret_cast<long long*>(first), (long long)m,
                                                                      (long long)n);
    MC_LAUNCH_CHECK("voc_fill_kernel");
  }
  voc_match_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(d_key), d_corners,
                                                                      (long long)n, reinterpret_cast<const long long*>(d_gkey),
                                                                      d_gbox, d_gdiff, (long long)m, ovthresh, first, jmax,
                                                                      state);
  MC_LAUNCH_CHECK("voc_match_kernel");
  voc_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(first, jmax, state, (long long)n, d_tp, d_fp);
  MC_LAUNCH_CHECK("voc_mark_kernel");
  return 0;
}
