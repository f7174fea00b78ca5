// ResNet backbone pieces that are not GEMMs (detectron2 build_resnet_backbone, selected by the reference's
// configs/cityscapes/semantic-segmentation/Base-Cityscapes-SemanticSegmentation.yaml:4,8-15; BASELINE.json configs[0]).
// The bottleneck 1x1 convolutions run on gemm_tc (batch norm folded into weights + bias at load time), the 3x3 ones on the
// tcgen05 implicit-GEMM convolution; what is left is
//   stem_conv     maskformer_model.py:255-257 (normalise, zero-pad) + BasicStem 7x7 / stride 2 conv + folded BN + ReLU
//   maxpool       3x3 / stride 2 / pad 1 max-pool of the stem, written as fp32 + split planes
//   bias_act_sub  y = [relu](x + bias) taken at every `stride`-th pixel: the epilogue of the 3x3 convolutions (their stride-2
//                 variants are computed at stride 1 and sub-sampled here), the ReLU after the residual add, and the stride-2
//                 gather in front of the projection shortcuts
// Activations are NHWC fp32 / bf16 split planes like everywhere else in the engine.
#include "kernels.cuh"

namespace rba {

constexpr int ST_K = 147;          // 3 * 7 * 7
constexpr int ST_C = 64;

// One thread = one output pixel x 16 channels (4 threads per pixel, 64 pixels per CTA); the folded filter lives in shared
// memory as [147][64].  FLOPs are negligible (1.2 GMAC at 512 x 1024); the input patch is re-read through L1.
template <typename T>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const T* __restrict__ img, int B, int H, int W, int Hp, int Wp, float m0, float m1, float m2, float s0, float s1,
                 float s2, const float* __restrict__ w /* [64][147], BN folded */, const float* __restrict__ bias,
                 float* __restrict__ out /* (B, Hp/2, Wp/2, 64) */) {
  __shared__ float sW[ST_K * ST_C];
  for (int e = threadIdx.x; e < ST_K * ST_C; e += blockDim.x) {
    const int k = e / ST_C, c = e - k * ST_C;
    sW[e] = w[c * ST_K + k];
  }
  __syncthreads();
  const int Ho = Hp >> 1, Wo = Wp >> 1;
  const int64_t npix = (int64_t)B * Ho * Wo;
  const int cg = (threadIdx.x & 3) * 16;
  for (int64_t px = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); px < npix; px += (int64_t)gridDim.x * 64) {
    const int ox = (int)(px % Wo);
    const int64_t t = px / Wo;
    const int oy = (int)(t % Ho), b = (int)(t / Ho);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = bias[cg + j];
    for (int ch = 0; ch < 3; ++ch) {
      const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
      const float inv = 1.0f / (ch == 0 ? s0 : (ch == 1 ? s1 : s2));
      const T* plane = img + ((int64_t)b * 3 + ch) * H * W;
      for (int ky = 0; ky < 7; ++ky) {
        const int yy = oy * 2 + ky - 3;
        if (yy < 0 || yy >= H) continue;                  // conv padding and the ImageList padding are both zeros of the
        for (int kx = 0; kx < 7; ++kx) {                   // NORMALISED image
          const int xx = ox * 2 + kx - 3;
          if (xx < 0 || xx >= W) continue;
          const float v = ((float)plane[(int64_t)yy * W + xx] - mean) * inv;
          const float4* wr = reinterpret_cast<const float4*>(sW + (ch * 49 + ky * 7 + kx) * ST_C + cg);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = wr[j4];
            acc[4 * j4] = fmaf(w4.x, v, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(w4.y, v, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(w4.z, v, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(w4.w, v, acc[4 * j4 + 3]);
          }
        }
      }
    }
    float4* o = reinterpret_cast<float4*>(out + px * ST_C + cg);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4)
      o[j4] = make_float4(fmaxf(acc[4 * j4], 0.f), fmaxf(acc[4 * j4 + 1], 0.f), fmaxf(acc[4 * j4 + 2], 0.f), fmaxf(acc[4 * j4 + 3], 0.f));
  }
}

int stem_conv(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean, const float* stdv,
              const float* w, const float* bias, float* out, cudaStream_t st) {
  RBA_CHECK(images && w && bias && out, "stem_conv: null pointer");
  RBA_CHECK(Hp % 4 == 0 && Wp % 4 == 0 && H <= Hp && W <= Wp, "stem_conv: bad padded size");
  const int64_t npix = (int64_t)B * (Hp / 2) * (Wp / 2);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(npix, 64), 148 * 8);
  if (img_dtype == RBA_IMG_U8)
    stem_conv_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)images, B, H, W, Hp, Wp, mean[0], mean[1], mean[2], stdv[0],
                                                    stdv[1], stdv[2], w, bias, out);
  else if (img_dtype == RBA_IMG_F32)
    stem_conv_kernel<float><<<grid, 256, 0, st>>>((const float*)images, B, H, W, Hp, Wp, mean[0], mean[1], mean[2], stdv[0],
                                                  stdv[1], stdv[2], w, bias, out);
  else return fail(RBA_ERR_INVALID, "stem_conv: bad image dtype %d", img_dtype);
  RBA_LAUNCHED();
  return RBA_OK;
}

// 3x3 / stride 2 / pad 1 max-pool over NHWC fp32 (inputs are post-ReLU, so padding never wins); 4 channels per thread
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y, uint16_t* __restrict__ y_hi,
                    uint16_t* __restrict__ y_lo) {
  const int Ho = H >> 1, Wo = W >> 1, c4 = C >> 2;
  const int64_t total = (int64_t)B * Ho * Wo * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), b = (int)(t / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = oy * 2 + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = ox * 2 + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + yy) * W + xx) * C + c);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    const int64_t o = (((int64_t)b * Ho + oy) * Wo + ox) * C + c;
    if (y) *reinterpret_cast<float4*>(y + o) = m;
    if (y_hi) store_split4(y_hi, y_lo, o, m.x, m.y, m.z, m.w);
  }
}

int maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, uint16_t* y_hi, uint16_t* y_lo, cudaStream_t st) {
  RBA_CHECK(x && (y || y_hi), "maxpool: null pointer");
  RBA_CHECK(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "maxpool: H, W must be even and C a multiple of 4");
  const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 4);
  maxpool3x3s2_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 148 * 16), 256, 0, st>>>(x, B, H, W, C, y, y_hi, y_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// y[b, oy, ox, :] = act(x[b, oy * s, ox * s, :] + bias)   (x (B,H,W,C) fp32; y (B,H/s,W/s,C) fp32 and / or split planes; y may
// alias x when s == 1)
__global__ void __launch_bounds__(256)
bias_act_sub_kernel(const float* __restrict__ x, const float* __restrict__ bias, int B, int H, int W, int C, int s, int relu,
                    float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo) {
  const int Ho = H / s, Wo = W / s, c4 = C >> 2;
  const int64_t total = (int64_t)B * Ho * Wo * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    int64_t t = i / c4;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), b = (int)(t / Ho);
    float4 v = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + oy * s) * W + ox * s) * C + c);
    if (bias) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + c);
      v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    const int64_t o = (((int64_t)b * Ho + oy) * Wo + ox) * C + c;
    if (y) *reinterpret_cast<float4*>(y + o) = v;
    if (y_hi) store_split4(y_hi, y_lo, o, v.x, v.y, v.z, v.w);
  }
}

int bias_act_sub(const float* x, const float* bias, int B, int H, int W, int C, int stride, int relu, float* y, uint16_t* y_hi,
                 uint16_t* y_lo, cudaStream_t st) {
  RBA_CHECK(x && (y || y_hi), "bias_act_sub: null pointer");
  RBA_CHECK((stride == 1 || stride == 2) && H % stride == 0 && W % stride == 0 && C % 4 == 0, "bias_act_sub: bad shape / stride");
  RBA_CHECK(!(y == x && stride != 1), "bias_act_sub: in-place needs stride 1");
  const int64_t total = (int64_t)B * (H / stride) * (W / stride) * (C / 4);
  if (total == 0) return RBA_OK;
  bias_act_sub_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 148 * 16), 256, 0, st>>>(x, bias, B, H, W, C, stride, relu, y,
                                                                                             y_hi, y_lo);
  RBA_LAUNCHED();
  return RBA_OK;
}

// eval-mode batch norm folded into the preceding bias-free convolution: w'[o,:] = w[o,:] * g[o] / sqrt(var[o] + eps),
// b'[o] = beta[o] - mean[o] * g[o] / sqrt(var[o] + eps)
__global__ void bn_fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int O, int64_t I,
                                    float* __restrict__ w_out, float* __restrict__ b_out) {
  const int64_t total = (int64_t)O * I;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / I);
    const float sc = gamma[o] / sqrtf(var[o] + eps);
    w_out[i] = sc * w[i];
    if (i % I == 0) b_out[o] = beta[o] - mean[o] * sc;
  }
}

int bn_fold_conv(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps, int O,
                 int64_t I, float* w_out, float* b_out, cudaStream_t st) {
  RBA_CHECK(w && gamma && beta && mean && var && w_out && b_out, "bn_fold_conv: null pointer");
  bn_fold_conv_kernel<<<256, 256, 0, st>>>(w, gamma, beta, mean, var, eps, O, I, w_out, b_out);
  RBA_LAUNCHED();
  return RBA_OK;
}

}  // namespace rba

// ---- per-kernel C ABI (parity tests) ----
extern "C" int rba_k_stem_conv(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                               const float* stdv, const float* w, const float* bias, float* out, void* stream) {
  return rba::stem_conv(images, img_dtype, B, H, W, Hp, Wp, mean, stdv, w, bias, out, (cudaStream_t)stream);
}
extern "C" int rba_k_maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  return rba::maxpool3x3s2(x, B, H, W, C, y, y_hi, y_lo, (cudaStream_t)stream);
}
extern "C" int rba_k_bias_act_sub(const float* x, const float* bias, int B, int H, int W, int C, int stride, int relu, float* y,
                                  uint16_t* y_hi, uint16_t* y_lo, void* stream) {
  return rba::bias_act_sub(x, bias, B, H, W, C, stride, relu, y, y_hi, y_lo, (cudaStream_t)stream);
}
