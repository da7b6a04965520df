// layout.cu — layout conversion, weight packing, max-pool and the first (3-channel) layer.
//
// Replaces: nn.MaxPool2d(2,2) (src/nets.py:821), the NCHW fp32 tensors the reference passes between layers
// (src/nets.py:731-733), and conv1+bn1+leaky1+maxpool (models.0 / models.1, src/nets.py:789-822).
#include "common.cuh"

namespace {

__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }

// ---- NCHW fp32 -> PNHWC bf16 ---------------------------------------------------------------------------
// one thread per (row, 8-channel group); pad rows / channels >= C are written as zeros.
__global__ void pack_pnhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                                  int W, int C, int ld) {
  const int groups = ld / 8;
  const long long rows = (long long)B * (H + 1) * (W + 1);
  const long long total = rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / groups;
    const int g = (int)(i - row * groups);
    const int x = (int)(row % (W + 1));
    const long long t = row / (W + 1);
    const int y = (int)(t % (H + 1));
    const int b = (int)(t / (H + 1));
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float f = 0.f;
      if (x < W && y < H && c < C) f = in[(((long long)b * C + c) * H + y) * W + x];
      v[j] = __float2bfloat16_rn(f);
    }
    *reinterpret_cast<uint4*>(out + row * ld + g * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// ---- PNHWC bf16 -> NCHW fp32 (channels [ch_off, ch_off+C)) -----------------------------------------------
__global__ void unpack_pnhwc_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B, int H,
                                    int W, int C, int ld, int ch_off) {
  const long long total = (long long)B * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    long long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const long long row = ((long long)b * (H + 1) + y) * (W + 1) + x;
    out[i] = __bfloat162float(in[row * ld + ch_off + c]);
  }
}

// ---- 2x2/2 max-pool, PNHWC -> PNHWC ------------------------------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) pr[j] = __hmax2(pa[j], pb[j]);
  return r;
}

__global__ void maxpool2x2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                                  int W, int C8, int ld_in, int ld_out) {
  const int Ho = H / 2, Wo = W / 2;
  const long long rows = (long long)B * (Ho + 1) * (Wo + 1);
  const long long total = rows * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C8;
    const int g = (int)(i - row * C8);
    const int x = (int)(row % (Wo + 1));
    const long long t = row / (Wo + 1);
    const int y = (int)(t % (Ho + 1));
    const int b = (int)(t / (Ho + 1));
    uint4 r = make_uint4(0, 0, 0, 0);
    if (x < Wo && y < Ho) {
      const long long r00 = ((long long)b * (H + 1) + 2 * y) * (W + 1) + 2 * x;
      const uint4 a = *reinterpret_cast<const uint4*>(in + r00 * ld_in + g * 8);
      const uint4 bq = *reinterpret_cast<const uint4*>(in + (r00 + 1) * ld_in + g * 8);
      const uint4 c = *reinterpret_cast<const uint4*>(in + (r00 + W + 1) * ld_in + g * 8);
      const uint4 d = *reinterpret_cast<const uint4*>(in + (r00 + W + 2) * ld_in + g * 8);
      r = bf16x8_max(bf16x8_max(a, bq), bf16x8_max(c, d));
    }
    *reinterpret_cast<uint4*>(out + row * ld_out + g * 8) = r;
  }
}

// ---- weight packing ---------------------------------------------------------------------------------------
// out[o', tap*Kc + c'] = bf16( w[oidx[o'], cidx[c'], tap] * mask[...] ), zero outside the surviving sets.
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, const float* __restrict__ mask, int O, int C,
                                         int taps, const int* __restrict__ oidx, int n_o,
                                         const int* __restrict__ cidx, int n_c, __nv_bfloat16* __restrict__ out,
                                         int Npad, int Kc) {
  const long long total = (long long)Npad * taps * Kc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % Kc);
    long long t = i / Kc;
    const int tap = (int)(t % taps);
    const int op = (int)(t / taps);
    float v = 0.f;
    if (op < n_o && cp < n_c) {
      const int o = oidx ? oidx[op] : op;  // negative index = all-zero row / column
      const int c = cidx ? cidx[cp] : cp;
      if (o >= 0 && c >= 0) {
        const long long src = ((long long)o * C + c) * taps + tap;
        v = w[src];
        if (mask) v *= mask[src];
      }
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// ---- conv1: fp32 NCHW image, 3 input channels, 3x3, N<=32 filters, fused scale/shift + leaky + 2x2 pool ---------
// One block = 16x16 conv outputs (8x8 pooled); thread t owns conv pixel (2*wy+dy, 2*wx+dx) with
// window = t/4, (dy,dx) = t%4, so a pool window is 4 adjacent lanes and pooling is two shuffles.
constexpr int C1_MAXN = 32;
__global__ void __launch_bounds__(256)
conv1_direct_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ scale,
                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int B, int H, int W, int N,
                    int ldc, int pool) {
  __shared__ float s_in[3][18][19];
  __shared__ float s_w[27][C1_MAXN];
  __shared__ float s_sc[C1_MAXN], s_sh[C1_MAXN];
  const int tiles_x = (W + 15) / 16;
  const int tiles_y = (H + 15) / 16;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int ty = (tile / tiles_x) % tiles_y;
  const int tx = tile % tiles_x;
  const int y0 = ty * 16, x0 = tx * 16;

  for (int i = threadIdx.x; i < 27 * C1_MAXN; i += 256) {
    const int k = i / C1_MAXN, n = i % C1_MAXN;  // k = c*9 + r*3 + s  (weight layout [N,3,3,3])
    s_w[k][n] = (n < N) ? w[n * 27 + k] : 0.f;
  }
  if (threadIdx.x < C1_MAXN) {
    s_sc[threadIdx.x] = threadIdx.x < N ? scale[threadIdx.x] : 0.f;
    s_sh[threadIdx.x] = threadIdx.x < N ? shift[threadIdx.x] : 0.f;
  }
  for (int i = threadIdx.x; i < 3 * 18 * 18; i += 256) {
    const int c = i / 324, r = (i / 18) % 18, s = i % 18;
    const int yy = y0 + r - 1, xx = x0 + s - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = img[(((long long)b * 3 + c) * H + yy) * W + xx];
    // the tensor-core layers consume bf16 activations; round the image the same way for a uniform contract
    s_in[c][r][s] = __bfloat162float(__float2bfloat16_rn(v));
  }
  __syncthreads();

  const int t = threadIdx.x;
  const int win = t >> 2, sub = t & 3;
  const int ly = 2 * (win >> 3) + (sub >> 1);
  const int lx = 2 * (win & 7) + (sub & 1);
  float acc[C1_MAXN];
#pragma unroll
  for (int n = 0; n < C1_MAXN; ++n) acc[n] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float a = s_in[c][ly + r][lx + s];
        const int k = c * 9 + r * 3 + s;
#pragma unroll
        for (int n = 0; n < C1_MAXN; ++n) acc[n] = fmaf(a, s_w[k][n], acc[n]);
      }
  const int gy = y0 + ly, gx = x0 + lx;
#pragma unroll
  for (int n = 0; n < C1_MAXN; ++n) {
    float v = leaky01(acc[n] * s_sc[n] + s_sh[n]);
    if (pool) {
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    }
    acc[n] = v;
  }
  if (pool) {
    if (sub == 0 && gy < H && gx < W) {
      const int Ho = H / 2, Wo = W / 2;
      const long long row = ((long long)b * (Ho + 1) + (gy >> 1)) * (Wo + 1) + (gx >> 1);
      for (int n = 0; n < N; ++n) out[row * ldc + n] = __float2bfloat16_rn(acc[n]);
    }
  } else {
    if (gy < H && gx < W) {
      const long long row = ((long long)b * (H + 1) + gy) * (W + 1) + gx;
      for (int n = 0; n < N; ++n) out[row * ldc + n] = __float2bfloat16_rn(acc[n]);
    }
  }
}

// zero the pad column / pad line of a PNHWC buffer (channels [0, ld)).
__global__ void zero_pads_kernel(__nv_bfloat16* __restrict__ out, int B, int H, int W, int ld8) {
  const long long npad_rows = (long long)B * ((H + 1) + W);  // per image: pad column (H+1 rows) + pad line (W more)
  const long long total = npad_rows * ld8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pr = i / ld8;
    const int g = (int)(i - pr * ld8);
    const int b = (int)(pr / (H + 1 + W));
    const int q = (int)(pr % (H + 1 + W));
    int y, x;
    if (q <= H) { y = q; x = W; } else { y = H; x = q - (H + 1); }
    const long long row = ((long long)b * (H + 1) + y) * (W + 1) + x;
    *reinterpret_cast<uint4*>(out + (row * ld8 + g) * 8) = make_uint4(0, 0, 0, 0);
  }
}

inline int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)mc_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

extern "C" int mc_pack_pnhwc(const float* d_in, void* d_out, int B, int H, int W, int C, int ld_out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_out && B > 0 && H > 0 && W > 0 && C > 0, "mc_pack_pnhwc: bad argument");
  MC_CHECK_ARG(ld_out % 8 == 0 && ld_out >= C, "mc_pack_pnhwc: ld_out must be a multiple of 8 >= C");
  const long long total = (long long)B * (H + 1) * (W + 1) * (ld_out / 8);
  pack_pnhwc_kernel<<<grid_for(total, 256), 256, 0, stream>>>(d_in, (__nv_bfloat16*)d_out, B, H, W, C, ld_out);
  MC_LAUNCH_CHECK("pack_pnhwc_kernel");
  return 0;
}

extern "C" int mc_unpack_pnhwc(const void* d_in, float* d_out, int B, int H, int W, int C, int ld_in, int ch_off,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_out && B > 0 && H > 0 && W > 0 && C > 0 && ch_off >= 0 && ch_off + C <= ld_in,
               "mc_unpack_pnhwc: bad argument");
  const long long total = (long long)B * C * H * W;
  unpack_pnhwc_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __nv_bfloat16*)d_in, d_out, B, H, W, C, ld_in,
                                                                ch_off);
  MC_LAUNCH_CHECK("unpack_pnhwc_kernel");
  return 0;
}

extern "C" int mc_maxpool2x2(const void* d_in, void* d_out, int B, int H, int W, int C, int ld_in, int ld_out,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_out && B > 0 && H > 0 && W > 0 && C > 0, "mc_maxpool2x2: bad argument");
  MC_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "mc_maxpool2x2: H and W must be even");
  MC_CHECK_ARG(ld_in % 8 == 0 && ld_out % 8 == 0 && ld_in >= C && ld_out >= C, "mc_maxpool2x2: bad pitch");
  const int C8 = (C + 7) / 8;
  const long long total = (long long)B * (H / 2 + 1) * (W / 2 + 1) * C8;
  maxpool2x2_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const __nv_bfloat16*)d_in, (__nv_bfloat16*)d_out, B, H,
                                                              W, C8, ld_in, ld_out);
  MC_LAUNCH_CHECK("maxpool2x2_kernel");
  return 0;
}

extern "C" int mc_pack_conv_weights(const float* d_w, const float* d_mask, int O, int C, int ksize, const int* d_oidx,
                                    int n_o, const int* d_cidx, int n_c, void* d_wpack, int Npad, int Kc,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_w && d_wpack && O > 0 && C > 0 && (ksize == 1 || ksize == 3), "mc_pack_conv_weights: bad argument");
  MC_CHECK_ARG(n_o > 0 && n_o <= Npad && n_c > 0 && n_c <= Kc && (Kc % 64) == 0 && (Npad % 16) == 0,
               "mc_pack_conv_weights: bad packed dims (n_o=%d Npad=%d n_c=%d Kc=%d)", n_o, Npad, n_c, Kc);
  MC_CHECK_ARG((d_oidx || n_o <= O) && (d_cidx || n_c <= C), "mc_pack_conv_weights: counts exceed tensor dims");
  const int taps = ksize * ksize;
  const long long total = (long long)Npad * taps * Kc;
  pack_conv_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(d_w, d_mask, O, C, taps, d_oidx, n_o, d_cidx, n_c,
                                                                     (__nv_bfloat16*)d_wpack, Npad, Kc);
  MC_LAUNCH_CHECK("pack_conv_weights_kernel");
  return 0;
}

extern "C" int mc_conv1_fwd(const float* d_img, const float* d_w, const float* d_scale, const float* d_shift,
                            void* d_out, int B, int H, int W, int N, int ldc, int pool, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_img && d_w && d_scale && d_shift && d_out, "mc_conv1_fwd: null pointer");
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0 && N > 0 && N <= C1_MAXN, "mc_conv1_fwd: N must be in 1..%d (got %d)", C1_MAXN, N);
  MC_CHECK_ARG(ldc % 8 == 0 && ldc >= N, "mc_conv1_fwd: ldc must be a multiple of 8 >= N");
  if (pool) MC_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "mc_conv1_fwd: pooling needs even H,W");
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  // pad line/column of the destination (and channels >= N) must be zero: clear the pads, the kernel
  // writes channels [0,N) of interior rows; channels [N,ldc) are never read (TMA extent = Cin).
  {
    const long long total = (long long)B * ((Ho + 1) + Wo) * (ldc / 8);
    zero_pads_kernel<<<grid_for(total, 256), 256, 0, stream>>>((__nv_bfloat16*)d_out, B, Ho, Wo, ldc / 8);
    MC_LAUNCH_CHECK("zero_pads_kernel");
  }
  const int tiles = B * ((H + 15) / 16) * ((W + 15) / 16);
  conv1_direct_kernel<<<tiles, 256, 0, stream>>>(d_img, d_w, d_scale, d_shift, (__nv_bfloat16*)d_out, B, H, W, N, ldc,
                                                 pool);
  MC_LAUNCH_CHECK("conv1_direct_kernel");
  return 0;
}
