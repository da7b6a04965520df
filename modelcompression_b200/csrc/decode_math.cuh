// decode_math.cuh — the region-layer arithmetic of get_region_boxes (src/nets2_utils.py:158-205), shared by the
// stand-alone decode kernel (detect.cu) and the fused decode epilogue of the head convolution (conv_tcgen05.cu), so both
// produce bit-identical boxes from identical fp32 logits.  Every operation is a separately rounded fp32 op (no FMA
// contraction); transcendentals are accurate expf.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define MC_MAX_ANCHORS 16

struct McAnchors {
  float w[MC_MAX_ANCHORS];
  float h[MC_MAX_ANCHORS];
};

__device__ __forceinline__ float mc_sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

struct McDecoded {
  float bx, by, bw, bh, conf, cmax;
  int cid;
  bool cand;
  float cls_max_logit, cls_sum;  // softmax pieces: p[c] = exp(l[c] - cls_max_logit) / cls_sum
};

// One anchor of one cell.  L(f) returns logit f (0 tx, 1 ty, 2 tw, 3 th, 4 to, 5.. classes) of this anchor.
template <typename Logit>
__device__ __forceinline__ McDecoded mc_decode_anchor(Logit L, int nc, int cx, int cy, int W, int H, float aw, float ah,
                                                      float thresh, int only_objectness) {
  McDecoded d;
  d.conf = mc_sigmoid_ref(L(4));
  float mx = -INFINITY;
  for (int c = 0; c < nc; ++c) mx = fmaxf(mx, L(5 + c));
  float sum = 0.f;
  for (int c = 0; c < nc; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(L(5 + c), mx)));
  d.cls_max_logit = mx;
  d.cls_sum = sum;
  d.cmax = -1.f;
  d.cid = 0;
  for (int c = 0; c < nc; ++c) {
    const float pc = __fdiv_rn(expf(__fsub_rn(L(5 + c), mx)), sum);
    if (pc > d.cmax) { d.cmax = pc; d.cid = c; }
  }
  const float score = only_objectness ? d.conf : __fmul_rn(d.conf, d.cmax);
  d.cand = score > thresh;
  d.bx = d.by = d.bw = d.bh = 0.f;
  if (d.cand) {
    d.bx = __fdiv_rn(__fadd_rn(mc_sigmoid_ref(L(0)), (float)cx), (float)W);
    d.by = __fdiv_rn(__fadd_rn(mc_sigmoid_ref(L(1)), (float)cy), (float)H);
    d.bw = __fdiv_rn(__fmul_rn(expf(L(2)), aw), (float)W);
    d.bh = __fdiv_rn(__fmul_rn(expf(L(3)), ah), (float)H);
  }
  return d;
}
