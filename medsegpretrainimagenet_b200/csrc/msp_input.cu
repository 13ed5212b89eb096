// Input-side kernels of the hot path: what the reference does to a batch on the CPU between the
// data loader and the first convolution, moved onto the device so that only the raw bytes cross PCIe.
//
//  * msp_u8_to_f32_nchw: `np.load(fname) / 255` (classification/datasets.py:47) -> float32 cast
//    (transform/transforms.py:63-103, ConvertToType.default_transform) -> `np.repeat(x, repeats, axis=0)`
//    (RepeatChannels, transform/transforms.py:134-142) in one pass over the uint8 image.
//  * msp_color_jitter: torchvision.transforms.ColorJitter as robustness/eval.py:61-66 applies it to the whole image
//    batch (ONE random draw per call: the op order and the four factors are host arguments taken from torchvision's own
//    `get_params`, so the CPU generator is consumed exactly like the reference does).  All four adjustments run per
//    pixel in registers in one pass; only `adjust_contrast` needs a statistic of the whole image (the mean of its
//    gray-scale version at that point of the chain), produced by a first reduction pass that replays the ops before it.
//    The arithmetic follows torchvision/_functional_tensor.py operation by operation with separately rounded fp32
//    multiplies and adds (no FMA contraction), so everything but the contrast mean is bit-identical to the CPU result.
//
// HBM-bound: 4 B read + 4 B written per element (color jitter: + one more read when contrast is active).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

struct JitterParams {
  int order[4];          // op ids in application order: 0 brightness, 1 contrast, 2 saturation, 3 hue, -1 = skipped
  float a[4], b[4];      // blend coefficients per op id: ratio, (1 - ratio) as torch rounds them
  float hue;             // hue shift
};

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
// ratio * img1 + (1 - ratio) * img2, the two products rounded separately like torch's elementwise kernels
__device__ __forceinline__ float blend(float x, float y, float a, float b) {
  return clamp01(__fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y)));
}
__device__ __forceinline__ float gray_of(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
// python-style remainder by 1 for |v| < 2 (torch.remainder / `%` on tensors)
__device__ __forceinline__ float mod1(float v) {
  float r = fmodf(v, 1.0f);
  if (r != 0.f && r < 0.f) r = __fadd_rn(r, 1.0f);
  return r;
}

__device__ __forceinline__ void hue_shift(float& r, float& g, float& b, float hue) {
  // _rgb2hsv
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float s = __fdiv_rn(cr, eqc ? 1.0f : maxc);
  const float div = eqc ? 1.0f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div),
              bc = __fdiv_rn(__fsub_rn(maxc, b), div);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.f;
  const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.0f, rc), bc) : 0.f;
  const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.0f, gc), rc) : 0.f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.0f), 1.0f), 1.0f);
  // shift
  h = mod1(__fadd_rn(h, hue));
  // _hsv2rgb
  const float h6 = __fmul_rn(h, 6.0f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  int i = (int)fi;
  const float v = maxc;
  const float p = clamp01(__fmul_rn(v, __fsub_rn(1.0f, s)));
  const float q = clamp01(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, f))));
  const float t = clamp01(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, __fsub_rn(1.0f, f)))));
  i = ((i % 6) + 6) % 6;
  switch (i) {
    case 0: r = v; g = t; b = p; break;
    case 1: r = q; g = v; b = p; break;
    case 2: r = p; g = v; b = t; break;
    case 3: r = p; g = q; b = v; break;
    case 4: r = t; g = p; b = v; break;
    default: r = v; g = p; b = q; break;
  }
}

// Applies ops order[0..upto) to one RGB pixel; `mean` = gray mean used by the contrast op.
__device__ __forceinline__ void jitter_rgb(float& r, float& g, float& b, const JitterParams& p, int upto, float mean) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= upto) break;
    const int op = p.order[k];
    if (op == 0) {
      r = blend(r, 0.f, p.a[0], p.b[0]); g = blend(g, 0.f, p.a[0], p.b[0]); b = blend(b, 0.f, p.a[0], p.b[0]);
    } else if (op == 1) {
      r = blend(r, mean, p.a[1], p.b[1]); g = blend(g, mean, p.a[1], p.b[1]); b = blend(b, mean, p.a[1], p.b[1]);
    } else if (op == 2) {
      const float y = gray_of(r, g, b);
      r = blend(r, y, p.a[2], p.b[2]); g = blend(g, y, p.a[2], p.b[2]); b = blend(b, y, p.a[2], p.b[2]);
    } else if (op == 3) {
      hue_shift(r, g, b, p.hue);
    }
  }
}
// single-channel images: saturation and hue leave them untouched (torchvision: "match PIL behaviour")
__device__ __forceinline__ float jitter_gray(float v, const JitterParams& p, int upto, float mean) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= upto) break;
    const int op = p.order[k];
    if (op == 0) v = blend(v, 0.f, p.a[0], p.b[0]);
    else if (op == 1) v = blend(v, mean, p.a[1], p.b[1]);
  }
  return v;
}

// pass 1: per-image sum of the gray-scale image after the ops that precede the contrast adjustment
__global__ void __launch_bounds__(256)
jitter_gray_sum_kernel(const float* __restrict__ x, int c, long long hw, JitterParams p, int upto,
                       double* __restrict__ sums) {
  const int n = blockIdx.y;
  const float* xi = x + (long long)n * c * hw;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (c == 3) {
      float r = xi[i], g = xi[hw + i], b = xi[2 * hw + i];
      jitter_rgb(r, g, b, p, upto, 0.f);
      acc += (double)gray_of(r, g, b);
    } else {
      acc += (double)jitter_gray(xi[i], p, upto, 0.f);
    }
  }
  acc = warp_sum_d(acc);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(sums + n, t);
  }
}

// pass 2: the whole chain
__global__ void __launch_bounds__(256)
jitter_apply_kernel(const float* __restrict__ x, int c, long long hw, JitterParams p, const double* __restrict__ sums,
                    float* __restrict__ y) {
  const int n = blockIdx.y;
  const float* xi = x + (long long)n * c * hw;
  float* yi = y + (long long)n * c * hw;
  const float mean = sums ? (float)(sums[n] / (double)hw) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    if (c == 3) {
      float r = xi[i], g = xi[hw + i], b = xi[2 * hw + i];
      jitter_rgb(r, g, b, p, 4, mean);
      yi[i] = r; yi[hw + i] = g; yi[2 * hw + i] = b;
    } else {
      yi[i] = jitter_gray(xi[i], p, 4, mean);
    }
  }
}

__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint8_t* __restrict__ x, long long planes, long long hw, int repeats, double divisor,
                 float* __restrict__ y) {
  // one thread = 4 consecutive pixels of one source plane, written to `repeats` destination planes
  const long long q4 = (hw + 3) >> 2;
  const long long total = planes * q4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long pl = t / q4, i0 = (t - pl * q4) << 2;
    const uint8_t* src = x + pl * hw + i0;
    float v[4];
    const int cnt = (int)((hw - i0) < 4 ? (hw - i0) : 4);
    if (cnt == 4 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(src);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (float)((double)((w >> (8 * j)) & 0xffu) / divisor);
    } else {
      for (int j = 0; j < cnt; ++j) v[j] = (float)((double)src[j] / divisor);
    }
    for (int r = 0; r < repeats; ++r) {
      float* dst = y + (pl * repeats + r) * hw + i0;
      if (cnt == 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int j = 0; j < cnt; ++j) dst[j] = v[j];
      }
    }
  }
}

}  // namespace

extern "C" int msp_u8_to_f32_nchw(const void* x_u8, int n, int c, long long hw, int repeats, double divisor,
                                  float* y, void* stream) {
  MSP_REQUIRE(x_u8 && y, "u8_to_f32: null pointer");
  MSP_REQUIRE(n >= 0 && c > 0 && hw > 0 && repeats >= 1 && divisor != 0.0, "u8_to_f32: bad arguments");
  if (n == 0) return MSP_OK;
  const long long planes = (long long)n * c;
  const long long work = planes * ((hw + 3) >> 2);
  const long long want = (work + 255) / 256;
  const int blocks = (int)(want < (long long)msp_num_sms() * 8 ? want : (long long)msp_num_sms() * 8);
  u8_to_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)x_u8, planes, hw, repeats, divisor, y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_color_jitter(const float* x, int n, int c, long long hw, const int* order_host,
                                float brightness, float contrast, float saturation, float hue,
                                const float* one_minus_host, double* gray_sums, float* y, void* stream) {
  MSP_REQUIRE(x && y && order_host && one_minus_host, "color_jitter: null pointer");
  MSP_REQUIRE(c == 1 || c == 3, "color_jitter: images must have 1 or 3 channels (got %d)", c);
  MSP_REQUIRE(n >= 0 && hw > 0, "color_jitter: bad shape");
  MSP_REQUIRE(hue >= -0.5f && hue <= 0.5f, "color_jitter: hue factor %f is not in [-0.5, 0.5]", (double)hue);
  if (n == 0) return MSP_OK;
  JitterParams p;
  int contrast_at = -1;
  for (int k = 0; k < 4; ++k) {
    const int op = order_host[k];
    MSP_REQUIRE(op >= -1 && op <= 3, "color_jitter: op id %d", op);
    p.order[k] = op;
    if (op == 1 && contrast_at < 0) contrast_at = k;
  }
  p.a[0] = brightness; p.a[1] = contrast; p.a[2] = saturation; p.a[3] = 0.f;
  for (int k = 0; k < 3; ++k) p.b[k] = one_minus_host[k];
  p.b[3] = 0.f;
  p.hue = hue;
  cudaStream_t st = (cudaStream_t)stream;
  const long long want = (hw + 255) / 256;
  const int per_img = (int)(want < 4096 ? want : 4096);
  int gx = (msp_num_sms() * 8 + n - 1) / n;
  if (gx > per_img) gx = per_img;
  if (gx < 1) gx = 1;
  MSP_REQUIRE(n <= 65535, "color_jitter: at most 65535 images per call");
  dim3 grid((unsigned)gx, (unsigned)n);
  if (contrast_at >= 0) {
    MSP_REQUIRE(gray_sums != nullptr, "color_jitter: the contrast adjustment needs the [n] double workspace");
    MSP_CHECK_CUDA(cudaMemsetAsync(gray_sums, 0, sizeof(double) * (size_t)n, st));
    jitter_gray_sum_kernel<<<grid, 256, 0, st>>>(x, c, hw, p, contrast_at, gray_sums);
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
  }
  jitter_apply_kernel<<<grid, 256, 0, st>>>(x, c, hw, p, contrast_at >= 0 ? gray_sums : nullptr, y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
