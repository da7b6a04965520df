// region_loss.cu — YOLOv2 region loss (build_targets + the six loss terms + the gradient w.r.t. the head) on the GPU.
//
// Replaces RegionLoss.forward + build_targets of the reference (src/nets.py:282-440, 468-610), which copy the head to the
// CPU and run three nested Python loops (image x ground-truth box x anchor) per batch, and the autograd backward of the
// loss expression (src/train.py:229-233).  Two launches:
//   region_targets_kernel  one block per image: end of the ground-truth list (first x == 0, :328/:370), best anchor per
//                          box (:395-410; no positive IoU -> Python's [-1] = the LAST anchor), its (anchor, cell), the
//                          "later box overwrites the earlier" rule of the sequential loops (:412-436), the targets
//                          tx/ty/tw/th/tconf/tcls (tw = gw / anchor_w, NOT its logarithm; tconf = IoU with the prediction
//                          at the cell, whose w/h are exp()-ed TWICE, :511-512 and :546-547), nGT / nCorrect;
//   region_loss_kernel     one thread per (image, anchor, cell): sigmoid / exp of its five box logits, best IoU with the
//                          image's boxes -> conf_mask (:340-352), its targets if it is an assigned cell, the six
//                          half-SSE / cross-entropy terms (summed in float64) and d loss / d head.
// All float32 arithmetic uses the round-to-nearest intrinsics in the reference's operation order (no FMA contraction), so
// values agree with the PyTorch float32 result to the last bits; the sums are float64 and therefore differ from PyTorch's
// float32 reduction order by ~1e-7 relative.
#include "common.cuh"

namespace {

constexpr int RL_MAX_BOX = 50;      // dataloader.py:82-96 pads every label list to 50 rows of (cls, x, y, w, h)
constexpr int RL_MAX_ANCHORS = 16;
constexpr int RL_REC = 12;          // floats per box record

struct RlAnchors {
  float w[RL_MAX_ANCHORS], h[RL_MAX_ANCHORS], area[RL_MAX_ANCHORS];
};

// bbox_ious(..., x1y1x2y2=False), src/nets2_utils.py:100-131, float32, same operation order
__device__ __forceinline__ float iou_center(float x1, float y1, float w1, float h1, float x2, float y2, float w2, float h2) {
  const float hw1 = __fmul_rn(w1, 0.5f), hw2 = __fmul_rn(w2, 0.5f), hh1 = __fmul_rn(h1, 0.5f), hh2 = __fmul_rn(h2, 0.5f);
  const float mx = fminf(__fsub_rn(x1, hw1), __fsub_rn(x2, hw2));
  const float Mx = fmaxf(__fadd_rn(x1, hw1), __fadd_rn(x2, hw2));
  const float my = fminf(__fsub_rn(y1, hh1), __fsub_rn(y2, hh2));
  const float My = fmaxf(__fadd_rn(y1, hh1), __fadd_rn(y2, hh2));
  const float uw = __fsub_rn(Mx, mx), uh = __fsub_rn(My, my);
  const float cw = __fsub_rn(__fadd_rn(w1, w2), uw), ch = __fsub_rn(__fadd_rn(h1, h2), uh);
  const bool bad = cw <= 0.f || ch <= 0.f;
  const float area1 = __fmul_rn(w1, h1), area2 = __fmul_rn(w2, h2);
  const float carea = bad ? 0.f : __fmul_rn(cw, ch);
  const float uarea = __fsub_rn(__fadd_rn(area1, area2), carea);
  return __fdiv_rn(carea, uarea);
}
__device__ __forceinline__ float sigmoidf_rn(float v) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v))); }

// rec[b][t] = { anchor*nH*nW + cell (as float bits), win, tx, ty, tw, th, tconf, tcls, gx, gy, gw, gh };  nvalid[b]
__global__ void __launch_bounds__(64) region_targets_kernel(const float* __restrict__ out, const float* __restrict__ target,
                                                            int nA, int nC, int nH, int nW, RlAnchors anc,
                                                            float* __restrict__ rec, int* __restrict__ nvalid,
                                                            int* __restrict__ counts) {
  __shared__ int s_lin[RL_MAX_BOX];
  __shared__ int s_n;
  const int b = blockIdx.x, t = threadIdx.x;
  const float* tg = target + (size_t)b * RL_MAX_BOX * 5;
  if (t == 0) s_n = RL_MAX_BOX;
  __syncthreads();
  float cls = 0.f, bx = 0.f, by = 0.f, bw = 0.f, bh = 0.f;
  if (t < RL_MAX_BOX) {
    cls = tg[t * 5 + 0]; bx = tg[t * 5 + 1]; by = tg[t * 5 + 2]; bw = tg[t * 5 + 3]; bh = tg[t * 5 + 4];
    if (bx == 0.f) atomicMin(&s_n, t);  // the list ends at the first box whose x is 0
  }
  __syncthreads();
  const int n = s_n;
  const bool valid = t < n;
  float gx = 0.f, gy = 0.f, gw = 0.f, gh = 0.f, iou_gt = 0.f;
  int best = 0, gi = 0, gj = 0, lin = -1;
  if (valid) {
    gx = __fmul_rn(bx, (float)nW); gy = __fmul_rn(by, (float)nH);
    gw = __fmul_rn(bw, (float)nW); gh = __fmul_rn(bh, (float)nH);
    gi = (int)gx; gj = (int)gy;  // int(): truncation
    gi = gi < 0 ? 0 : (gi >= nW ? nW - 1 : gi);
    gj = gj < 0 ? 0 : (gj >= nH ? nH - 1 : gj);
    // best anchor: IoU of (0, 0, aw, ah) with (0, 0, gw, gh); first maximum
    float best_iou = 0.f;
    best = -1;
    const float hgw = __fmul_rn(gw, 0.5f), hgh = __fmul_rn(gh, 0.5f), garea = __fmul_rn(gw, gh);
    for (int a = 0; a < nA; ++a) {
      const float haw = __fmul_rn(anc.w[a], 0.5f), hah = __fmul_rn(anc.h[a], 0.5f);
      const float mx = fminf(__fsub_rn(0.f, haw), __fsub_rn(0.f, hgw)), Mx = fmaxf(__fadd_rn(0.f, haw), __fadd_rn(0.f, hgw));
      const float my = fminf(__fsub_rn(0.f, hah), __fsub_rn(0.f, hgh)), My = fmaxf(__fadd_rn(0.f, hah), __fadd_rn(0.f, hgh));
      const float cw = __fsub_rn(__fadd_rn(anc.w[a], gw), __fsub_rn(Mx, mx));
      const float ch = __fsub_rn(__fadd_rn(anc.h[a], gh), __fsub_rn(My, my));
      const float carea = __fmul_rn(cw, ch);
      const float iou = (cw <= 0.f || ch <= 0.f) ? 0.f : __fdiv_rn(carea, __fsub_rn(__fadd_rn(anc.area[a], garea), carea));
      if (iou > best_iou) { best_iou = iou; best = a; }
    }
    if (best < 0) best = nA - 1;  // best_n stays -1 in the reference: Python indexing addresses the last anchor
    lin = (best * nH + gj) * nW + gi;
    // the prediction at that cell; w, h go through exp() twice (once for the loss, once more for the box)
    const size_t hw = (size_t)nH * nW;
    const float* o = out + (((size_t)b * nA + best) * (5 + nC)) * hw + (size_t)gj * nW + gi;
    const float px = __fadd_rn(sigmoidf_rn(o[0]), (float)gi), py = __fadd_rn(sigmoidf_rn(o[hw]), (float)gj);
    const float pw = __fmul_rn(expf(expf(o[2 * hw])), anc.w[best]), ph = __fmul_rn(expf(expf(o[3 * hw])), anc.h[best]);
    iou_gt = iou_center(gx, gy, gw, gh, px, py, pw, ph);
  }
  if (t < RL_MAX_BOX) s_lin[t] = lin;
  __syncthreads();
  if (t < RL_MAX_BOX) {
    bool win = valid;
    for (int u = t + 1; u < n && win; ++u) win = s_lin[u] != lin;  // sequential assignment: the LAST box of a cell wins
    float* r = rec + ((size_t)b * RL_MAX_BOX + t) * RL_REC;
    r[0] = __int_as_float(lin);
    r[1] = win ? 1.f : 0.f;
    r[2] = __fsub_rn(gx, (float)gi);
    r[3] = __fsub_rn(gy, (float)gj);
    r[4] = valid ? __fdiv_rn(gw, anc.w[best]) : 0.f;
    r[5] = valid ? __fdiv_rn(gh, anc.h[best]) : 0.f;
    r[6] = iou_gt;
    r[7] = cls;
    r[8] = gx; r[9] = gy; r[10] = gw; r[11] = gh;
    if (valid && iou_gt > 0.5f) atomicAdd(&counts[1], 1);
  }
  if (t == 0) {
    nvalid[b] = n;
    atomicAdd(&counts[0], n);
  }
}

// sums[0..5] += sum of squares of (x, y, w, h, conf) terms and the cross-entropy sum
__global__ void __launch_bounds__(256) region_loss_kernel(const float* __restrict__ out, int nB, int nA, int nC, int nH, int nW,
                                                          RlAnchors anc, const float* __restrict__ rec,
                                                          const int* __restrict__ nvalid, float coord_scale,
                                                          float noobject_scale, float object_scale, float class_scale,
                                                          float thresh, float* __restrict__ grad, double* __restrict__ sums) {
  __shared__ double s_red[6][8];
  const long long total = (long long)nB * nA * nH * nW;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (idx < total) {
    const int hw = nH * nW;
    const int cell = (int)(idx % hw);
    const long long ba = idx / hw;
    const int a = (int)(ba % nA), b = (int)(ba / nA);
    const int i = cell % nW, j = cell / nW;
    const float* o = out + ((size_t)ba * (5 + nC)) * hw + cell;
    float* g = grad + ((size_t)ba * (5 + nC)) * hw + cell;
    const float x = sigmoidf_rn(o[0]), y = sigmoidf_rn(o[hw]);
    const float w = expf(o[2 * (size_t)hw]), h = expf(o[3 * (size_t)hw]);
    const float conf = sigmoidf_rn(o[4 * (size_t)hw]);
    const float px = __fadd_rn(x, (float)i), py = __fadd_rn(y, (float)j);
    const float pw = __fmul_rn(expf(w), anc.w[a]), ph = __fmul_rn(expf(h), anc.h[a]);
    const int n = nvalid[b];
    const float* rb = rec + (size_t)b * RL_MAX_BOX * RL_REC;
    float cur = 0.f;
    int mine = -1;
    const int lin = a * hw + cell;
    for (int t = 0; t < n; ++t) {
      const float* r = rb + t * RL_REC;
      cur = fmaxf(cur, iou_center(px, py, pw, ph, r[8], r[9], r[10], r[11]));
      if (__float_as_int(r[0]) == lin && r[1] != 0.f) mine = t;
    }
    float conf_mask = cur > thresh ? 0.f : noobject_scale;
    const float inv_nB = 1.f / (float)nB;
    float tconf = 0.f;
    if (mine >= 0) {
      const float* r = rb + mine * RL_REC;
      conf_mask = object_scale;
      tconf = r[6];
      const float dx = __fsub_rn(x, r[2]), dy = __fsub_rn(y, r[3]), dw = __fsub_rn(w, r[4]), dh = __fsub_rn(h, r[5]);
      acc[0] = (double)__fmul_rn(dx, dx); acc[1] = (double)__fmul_rn(dy, dy);
      acc[2] = (double)__fmul_rn(dw, dw); acc[3] = (double)__fmul_rn(dh, dh);
      g[0] = coord_scale * dx * (1.f - x) * x * inv_nB;
      g[hw] = coord_scale * dy * (1.f - y) * y * inv_nB;
      g[2 * (size_t)hw] = coord_scale * dw * w * inv_nB;
      g[3 * (size_t)hw] = coord_scale * dh * h * inv_nB;
      // cross-entropy (sum) on this cell's class logits
      const int tc = (int)r[7];
      float mxl = -INFINITY;
      for (int c = 0; c < nC; ++c) mxl = fmaxf(mxl, o[(size_t)(5 + c) * hw]);
      float se = 0.f;
      for (int c = 0; c < nC; ++c) se += expf(o[(size_t)(5 + c) * hw] - mxl);
      const float lse = mxl + logf(se);
      acc[5] = (double)(lse - o[(size_t)(5 + tc) * hw]);
      for (int c = 0; c < nC; ++c) {
        const float p = expf(o[(size_t)(5 + c) * hw] - lse);
        g[(size_t)(5 + c) * hw] = class_scale * (p - (c == tc ? 1.f : 0.f)) * inv_nB;
      }
    } else {
      g[0] = 0.f; g[hw] = 0.f; g[2 * (size_t)hw] = 0.f; g[3 * (size_t)hw] = 0.f;
      for (int c = 0; c < nC; ++c) g[(size_t)(5 + c) * hw] = 0.f;
    }
    const float cm = sqrtf(conf_mask);
    const float dc = __fsub_rn(__fmul_rn(conf, cm), __fmul_rn(tconf, cm));
    acc[4] = (double)__fmul_rn(dc, dc);
    g[4 * (size_t)hw] = dc * cm * (1.f - conf) * conf * inv_nB;
  }
  // block reduction (float64), one atomic per term and block
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double v = acc[q];
    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    if (lane == 0) s_red[q][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0;
    for (int w2 = 0; w2 < 8; ++w2) v += s_red[threadIdx.x][w2];
    atomicAdd(&sums[threadIdx.x], v);
  }
}

// loss = (cs*sx/2 + cs*sy/2 + cs*sw/2 + cs*sh/2 + sconf/2 + cl*ce) / nB, float32 like the reference's scalar arithmetic
__global__ void region_finalize_kernel(const double* __restrict__ sums, float coord_scale, float class_scale, int nB,
                                       float* __restrict__ loss) {
  const float lx = coord_scale * ((float)sums[0] / 2.0f), ly = coord_scale * ((float)sums[1] / 2.0f);
  const float lw = coord_scale * ((float)sums[2] / 2.0f), lh = coord_scale * ((float)sums[3] / 2.0f);
  const float lc = (float)sums[4] / 2.0f, lcls = class_scale * (float)sums[5];
  *loss = (lx + ly + lw + lh + lc + lcls) / (float)nB;
}

}  // namespace

extern "C" size_t mc_workspace_bytes_region_loss(int nB) {
  if (nB <= 0) return 0;
  return 64 + (size_t)nB * (RL_MAX_BOX * RL_REC * sizeof(float) + sizeof(int));
}

extern "C" int mc_region_loss(const float* d_output, const float* d_target, int nB, int nA, int nC, int nH, int nW,
                              const double* h_anchors, float coord_scale, float noobject_scale, float object_scale,
                              float class_scale, float thresh, float* d_grad, float* d_loss, int* d_counts, void* d_ws,
                              size_t ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_output && d_target && h_anchors && d_grad && d_loss && d_counts && d_ws, "mc_region_loss: null pointer");
  MC_CHECK_ARG(nB > 0 && nA > 0 && nA <= RL_MAX_ANCHORS && nC > 0 && nH > 0 && nW > 0, "mc_region_loss: bad dims");
  MC_CHECK_ARG((long long)nB * nA * nH * nW * (5 + nC) < (1ll << 31), "mc_region_loss: head too large");
  if (ws_bytes < mc_workspace_bytes_region_loss(nB)) return mc_set_error(MC_ERR_WS, "mc_region_loss: workspace too small");
  RlAnchors anc;
  for (int a = 0; a < RL_MAX_ANCHORS; ++a) {
    anc.w[a] = a < nA ? (float)h_anchors[2 * a] : 1.f;
    anc.h[a] = a < nA ? (float)h_anchors[2 * a + 1] : 1.f;
    anc.area[a] = a < nA ? (float)(h_anchors[2 * a] * h_anchors[2 * a + 1]) : 1.f;  // double product, rounded once
  }
  double* sums = reinterpret_cast<double*>(d_ws);  // [6]
  int* nvalid = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(d_ws) + 64);
  float* rec = reinterpret_cast<float*>(nvalid + nB);
  MC_CUDA(cudaMemsetAsync(d_ws, 0, 64, stream));
  MC_CUDA(cudaMemsetAsync(d_counts, 0, 2 * sizeof(int), stream));
  region_targets_kernel<<<nB, 64, 0, stream>>>(d_output, d_target, nA, nC, nH, nW, anc, rec, nvalid, d_counts);
  MC_LAUNCH_CHECK("region_targets_kernel");
  const long long total = (long long)nB * nA * nH * nW;
  region_loss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_output, nB, nA, nC, nH, nW, anc, rec, nvalid,
                                                                          coord_scale, noobject_scale, object_scale,
                                                                          class_scale, thresh, d_grad, sums);
  MC_LAUNCH_CHECK("region_loss_kernel");
  region_finalize_kernel<<<1, 1, 0, stream>>>(sums, coord_scale, class_scale, nB, d_loss);
  MC_LAUNCH_CHECK("region_finalize_kernel");
  return 0;
}
