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
// one thread = one (output row, input channel): the taps of a filter are contiguous in the source (36 B for 3x3), the
// destination is written tap by tap with consecutive threads on consecutive channels.  Padding rows / columns are
// zeroed by a memset before the launch.
__global__ void pack_conv_weights_kernel(const float* __restrict__ w, const float* __restrict__ mask, int O, int C,
                                         int taps, const int* __restrict__ oidx, int n_o,
                                         const int* __restrict__ cidx, int n_c, __nv_bfloat16* __restrict__ out,
                                         int Npad, int Kc) {
  const unsigned int total = (unsigned int)n_o * (unsigned int)n_c;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned int op = i / (unsigned int)n_c, cp = i - op * (unsigned int)n_c;
    const int o = oidx ? oidx[op] : (int)op;  // negative index = all-zero row / column
    const int c = cidx ? cidx[cp] : (int)cp;
    if (o < 0 || c < 0) continue;
    const size_t src = ((size_t)o * C + c) * taps;
    __nv_bfloat16* dst = out + (size_t)op * taps * Kc + cp;
    for (int t = 0; t < taps; ++t) {
      float v = w[src + t];
      if (mask) v *= mask[src + t];
      dst[(size_t)t * Kc] = __float2bfloat16_rn(v);
    }
  }
}

// ---- small-channel direct conv (CUDA cores) ------------------------------------------------------------------
// For layers whose channel counts are too small to feed a 128 x N x 64 tensor-core tile (the 3-channel first layer,
// and the first blocks of a filter-pruned network: e.g. 3->4, 4->1, 1->17 channels), the op is bound by reading
// the input once, not by math.  One block = 16x16 output pixels; the (16+2)^2 x Cin input patch is staged in shared
// memory as bf16 (the same rounding the tensor-core layers apply to activations); thread t owns pixel
// (2*wy+dy, 2*wx+dx) with window = t/4, (dy,dx) = t%4, so a 2x2 pool window is 4 adjacent lanes (two shuffles).
// FIRST: input is the fp32 NCHW image; else PNHWC bf16.  Output PNHWC bf16 (interior rows only: the destination's
// pad line/column are zero-initialised once by the engine and never written).
constexpr int SD_MAX_CIN = 32;
constexpr int SD_MAX_WELEMS = 9 * 512;  // taps * Cin * NT floats of shared memory (18 KB)

template <int NT, bool FIRST>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const void* __restrict__ in_, const float* __restrict__ w, const float* __restrict__ scale,
                   const float* __restrict__ shift, __nv_bfloat16* __restrict__ out, int B, int H, int W, int Cin,
                   int Cin_ld, int N, int ldc, int ksize, int leaky, int pool) {
  __shared__ __nv_bfloat16 s_in[18 * 18 * SD_MAX_CIN];  // [r][s][c], c fastest
  __shared__ __align__(16) float s_w[SD_MAX_WELEMS];     // [tap][c][n]
  __shared__ float s_sc[NT], s_sh[NT];
  const int taps = ksize * ksize;
  const int halo = ksize / 2;
  const int tiles_x = (W + 15) / 16;
  const int tiles_y = (H + 15) / 16;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_x * tiles_y);
  const int ty = (tile / tiles_x) % tiles_y;
  const int tx = tile % tiles_x;
  const int y0 = ty * 16, x0 = tx * 16;
  const int P = 16 + 2 * halo;

  // weights: global [N][Cin][taps] fp32 -> shared [tap][c][n]
  for (int i = threadIdx.x; i < taps * Cin * NT; i += 256) {
    const int n = i % NT;
    const int c = (i / NT) % Cin;
    const int tp = i / (NT * Cin);
    s_w[i] = (n < N) ? w[((long long)n * Cin + c) * taps + tp] : 0.f;
  }
  if (threadIdx.x < NT) {
    s_sc[threadIdx.x] = threadIdx.x < N ? scale[threadIdx.x] : 0.f;
    s_sh[threadIdx.x] = threadIdx.x < N ? shift[threadIdx.x] : 0.f;
  }
  if (FIRST) {
    const float* img = reinterpret_cast<const float*>(in_);
    for (int i = threadIdx.x; i < Cin * P * P; i += 256) {
      const int c = i / (P * P), r = (i / P) % P, sx = i % P;
      const int yy = y0 + r - halo, xx = x0 + sx - halo;
      float v = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = img[(((long long)b * Cin + c) * H + yy) * W + xx];
      s_in[(r * P + sx) * Cin + c] = __float2bfloat16_rn(v);
    }
  } else {
    const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(in_);
    for (int i = threadIdx.x; i < P * P * Cin; i += 256) {
      const int c = i % Cin, sx = (i / Cin) % P, r = i / (Cin * P);
      const int yy = y0 + r - halo, xx = x0 + sx - halo;
      __nv_bfloat16 v = __float2bfloat16_rn(0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = x[(((long long)b * (H + 1) + yy) * (W + 1) + xx) * Cin_ld + c];
      s_in[(r * P + sx) * Cin + c] = v;
    }
  }
  __syncthreads();

  const int t = threadIdx.x;
  const int win = t >> 2, sub = t & 3;
  const int ly = 2 * (win >> 3) + (sub >> 1);
  const int lx = 2 * (win & 7) + (sub & 1);
  float acc[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n] = 0.f;
  for (int tp = 0; tp < taps; ++tp) {
    const int r = tp / ksize, sx = tp - r * ksize;
    const __nv_bfloat16* ip = &s_in[((ly + r) * P + (lx + sx)) * Cin];
    const float* wp = &s_w[tp * Cin * NT];
    for (int c = 0; c < Cin; ++c) {
      const float a = __bfloat162float(ip[c]);
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[n] = fmaf(a, wp[c * NT + n], acc[n]);
    }
  }
  const int gy = y0 + ly, gx = x0 + lx;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    float v = acc[n] * s_sc[n] + s_sh[n];
    if (leaky) v = leaky01(v);
    if (pool) {
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
      v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    }
    acc[n] = v;
  }
  long long row;
  bool store;
  if (pool) {
    const int Ho = H / 2, Wo = W / 2;
    row = ((long long)b * (Ho + 1) + (gy >> 1)) * (Wo + 1) + (gx >> 1);
    store = (sub == 0) && gy < H && gx < W;
  } else {
    row = ((long long)b * (H + 1) + gy) * (W + 1) + gx;
    store = gy < H && gx < W;
  }
  if (store) {
    __nv_bfloat16* dst = out + row * ldc;
    if ((NT % 8) == 0 && (ldc % 8) == 0) {
#pragma unroll
      for (int g = 0; g < NT / 8; ++g) {
        if (g * 8 + 8 <= N) {
          __align__(16) __nv_bfloat16 v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v8[j] = __float2bfloat16_rn(acc[g * 8 + j]);
          *reinterpret_cast<uint4*>(dst + g * 8) = *reinterpret_cast<const uint4*>(v8);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (g * 8 + j < N) dst[g * 8 + j] = __float2bfloat16_rn(acc[g * 8 + j]);
        }
      }
    } else {
#pragma unroll
      for (int n = 0; n < NT; ++n)
        if (n < N) dst[n] = __float2bfloat16_rn(acc[n]);
    }
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
  MC_CHECK_ARG(n_o > 0 && n_o <= Npad && n_c > 0 && n_c <= Kc && (Kc % 32) == 0 && (Npad % 16) == 0,
               "mc_pack_conv_weights: bad packed dims (n_o=%d Npad=%d n_c=%d Kc=%d)", n_o, Npad, n_c, Kc);
  MC_CHECK_ARG((d_oidx || n_o <= O) && (d_cidx || n_c <= C), "mc_pack_conv_weights: counts exceed tensor dims");
  const int taps = ksize * ksize;
  MC_CHECK_ARG((long long)n_o * n_c < (1ll << 32), "mc_pack_conv_weights: tensor too large");
  MC_CUDA(cudaMemsetAsync(d_wpack, 0, (size_t)Npad * taps * Kc * sizeof(__nv_bfloat16), stream));
  pack_conv_weights_kernel<<<grid_for((long long)n_o * n_c, 256), 256, 0, stream>>>(d_w, d_mask, O, C, taps, d_oidx, n_o,
                                                                                  d_cidx, n_c, (__nv_bfloat16*)d_wpack,
                                                                                  Npad, Kc);
  MC_LAUNCH_CHECK("pack_conv_weights_kernel");
  return 0;
}

template <bool FIRST>
static int launch_direct(const void* d_in, const float* d_w, const float* d_scale, const float* d_shift, void* d_out,
                         int B, int H, int W, int Cin, int Cin_ld, int N, int ldc, int ksize, int leaky, int pool,
                         cudaStream_t stream) {
  const int tiles = B * ((H + 15) / 16) * ((W + 15) / 16);
  __nv_bfloat16* out = (__nv_bfloat16*)d_out;
#define MC_LAUNCH_DIRECT(NT_)                                                                                  \
  conv_direct_kernel<NT_, FIRST><<<tiles, 256, 0, stream>>>(d_in, d_w, d_scale, d_shift, out, B, H, W, Cin, Cin_ld, N, \
                                                           ldc, ksize, leaky, pool)
  if (N <= 4) MC_LAUNCH_DIRECT(4);
  else if (N <= 8) MC_LAUNCH_DIRECT(8);
  else if (N <= 16) MC_LAUNCH_DIRECT(16);
  else MC_LAUNCH_DIRECT(32);
#undef MC_LAUNCH_DIRECT
  MC_LAUNCH_CHECK("conv_direct_kernel");
  return 0;
}

static int direct_nt(int N) { return N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32; }

extern "C" int mc_conv_direct_supported(int Cin, int N, int ksize) {
  if (!(ksize == 1 || ksize == 3) || Cin < 1 || N < 1 || Cin > SD_MAX_CIN || N > 32) return 0;
  return (ksize * ksize * Cin * direct_nt(N) <= SD_MAX_WELEMS) ? 1 : 0;
}

extern "C" int mc_conv_direct_fwd(const void* d_in, int in_is_nchw_f32, const float* d_w, const float* d_scale,
                                  const float* d_shift, void* d_out, int B, int H, int W, int Cin, int Cin_ld, int N,
                                  int ldc, int ksize, int leaky, int pool, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_w && d_scale && d_shift && d_out, "mc_conv_direct_fwd: null pointer");
  MC_CHECK_ARG(B > 0 && H > 0 && W > 0, "mc_conv_direct_fwd: bad dims");
  MC_CHECK_ARG(mc_conv_direct_supported(Cin, N, ksize), "mc_conv_direct_fwd: Cin=%d N=%d k=%d outside the direct path", Cin, N, ksize);
  MC_CHECK_ARG(ldc >= N, "mc_conv_direct_fwd: ldc < N");
  if (!in_is_nchw_f32) MC_CHECK_ARG(Cin_ld >= Cin, "mc_conv_direct_fwd: Cin_ld < Cin");
  if (pool) MC_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "mc_conv_direct_fwd: pooling needs even H,W");
  if (in_is_nchw_f32)
    return launch_direct<true>(d_in, d_w, d_scale, d_shift, d_out, B, H, W, Cin, Cin_ld, N, ldc, ksize, leaky, pool, stream);
  return launch_direct<false>(d_in, d_w, d_scale, d_shift, d_out, B, H, W, Cin, Cin_ld, N, ldc, ksize, leaky, pool, stream);
}


// ------------------------------------------------------------------------------------------------------------
// Stand-alone Reorg (src/nets.py:648-667) on the reference's own layout: fp32 NCHW in, fp32 NCHW out,
//   out[b, (i*s+j)*C + c, y, x] = in[b, c, s*y+i, s*x+j].
// Inside Darknet.forward the same shuffle is the store addressing of the producing conv's epilogue (MC_EPI_REORG2);
// this entry point serves a direct Reorg.forward call.
// ------------------------------------------------------------------------------------------------------------
namespace {
__global__ void reorg_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int H, int W,
                                  int s) {
  const int Ho = H / s, Wo = W / s;
  const long long total = (long long)B * C * H * W;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(o % Wo);
    long long t = o / Wo;
    const int y = (int)(t % Ho);
    t /= Ho;
    const int oc = (int)(t % ((long long)C * s * s));
    const int b = (int)(t / ((long long)C * s * s));
    const int c = oc % C, q = oc / C;
    const int i = q / s, j = q - i * s;
    out[o] = in[(((long long)b * C + c) * H + (s * y + i)) * W + (s * x + j)];
  }
}
}  // namespace

extern "C" int mc_reorg_nchw(const float* d_in, float* d_out, int B, int C, int H, int W, int stride, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MC_CHECK_ARG(d_in && d_out, "mc_reorg_nchw: null pointer");
  MC_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && stride > 0, "mc_reorg_nchw: bad dims");
  MC_CHECK_ARG((H % stride) == 0 && (W % stride) == 0, "mc_reorg_nchw: H, W must be multiples of the stride");
  const long long total = (long long)B * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)mc_num_sms() * 16) blocks = (long long)mc_num_sms() * 16;
  reorg_nchw_kernel<<<(int)blocks, 256, 0, stream>>>(d_in, d_out, B, C, H, W, stride);
  MC_LAUNCH_CHECK("reorg_nchw_kernel");
  return 0;
}
