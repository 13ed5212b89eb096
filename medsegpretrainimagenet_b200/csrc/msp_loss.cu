// Segmentation / classification heads and losses, forward + analytic gradient.
//
// Replaces (reference): final 1x1 conv + sigmoid/softmax (segmentation/models/unet_models.py:442-445,
// 685-686), DiceLoss (segmentation/losses/losses.py:11-58), CrossEntropyLoss / BCELoss
// (classification/losses.py:4-40) and torch.nn.BCELoss (utils/default_dict.py:10).
// All reductions: per-thread fp32 partials over a short strip, warp shuffle, then fp64 atomics so the
// result does not depend on the grid shape beyond fp64 rounding.
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

constexpr int kMaxHeadK = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------
// final 1x1 conv (K <= 8) + activation.  One thread per pixel; weights in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void final_conv_act_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long P,
                                          long long HW, int C, int x_cs, const float* __restrict__ w,
                                          const float* __restrict__ bias, int K, int act,
                                          float* __restrict__ logits, float* __restrict__ prob) {
  extern __shared__ float ws[];  // [K][C] + [K]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) ws[K * C + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P;
       pix += (long long)gridDim.x * blockDim.x) {
    float acc[kMaxHeadK];
#pragma unroll
    for (int k = 0; k < kMaxHeadK; ++k) acc[k] = k < K ? ws[K * C + k] : 0.f;
    const __nv_bfloat16* xp = x + pix * x_cs;
    for (int c = 0; c < C; c += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(xp + c);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
      float xv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16x2(uu[e]);
        xv[2 * e] = f.x;
        xv[2 * e + 1] = f.y;
      }
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k)
        if (k < K) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(xv[j], ws[k * C + c + j], acc[k]);
        }
    }
    const long long n = pix / HW, hw = pix - n * HW;
    float pr[kMaxHeadK];
    if (act == 1) {
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k) pr[k] = sigmoidf_(acc[k]);
    } else if (act == 2) {
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k)
        if (k < K) m = fmaxf(m, acc[k]);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k)
        if (k < K) {
          pr[k] = __expf(acc[k] - m);
          s += pr[k];
        }
      const float inv = 1.f / s;
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k) pr[k] *= inv;
    } else {
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k) pr[k] = acc[k];
    }
#pragma unroll
    for (int k = 0; k < kMaxHeadK; ++k)
      if (k < K) {
        const long long o = (n * K + k) * HW + hw;
        if (logits) logits[o] = acc[k];
        prob[o] = pr[k];
      }
  }
}

// backward: dlogits through the activation, dx (bf16 NHWC), dw / db via per-block smem reduction.
// Block = 256 threads = 256 pixels per iteration; dl (fp32 [K][256]) and the x tile (bf16 [256][C]) are
// staged in smem, then thread t owns (k, c) pairs t, t+256, ... and sums over the 256 pixels.
__global__ void final_conv_act_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long P,
                                          long long HW, int C, int x_cs, const float* __restrict__ w,
                                          int K, int act, const float* __restrict__ prob,
                                          const float* __restrict__ dprob,
                                          __nv_bfloat16* __restrict__ dx, int dx_cs,
                                          float* __restrict__ dw, float* __restrict__ db,
                                          float* __restrict__ rows_ws) {
  extern __shared__ __align__(16) uint8_t sraw[];
  float* ws = reinterpret_cast<float*>(sraw);                    // [K][C]
  float* dls = ws + K * C;                                       // [K][256]
  float* dwacc = dls + K * 256;                                  // [K][C] (+K for db)
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ws + ((2 * K * C + K * 256 + K + 3) & ~3));  // [256][C]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
    ws[i] = w[i];
    dwacc[i] = 0.f;
  }
  for (int i = threadIdx.x; i < K; i += blockDim.x) dwacc[K * C + i] = 0.f;
  __syncthreads();
  const long long tiles = (P + 255) / 256;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long pix = t * 256 + threadIdx.x;
    const bool ok = pix < P;
    float dl[kMaxHeadK];
#pragma unroll
    for (int k = 0; k < kMaxHeadK; ++k) dl[k] = 0.f;
    if (ok) {
      const long long n = pix / HW, hw = pix - n * HW;
      float pr[kMaxHeadK], dp[kMaxHeadK];
#pragma unroll
      for (int k = 0; k < kMaxHeadK; ++k)
        if (k < K) {
          const long long o = (n * K + k) * HW + hw;
          pr[k] = prob[o];
          dp[k] = dprob[o];
        }
      if (act == 1) {
#pragma unroll
        for (int k = 0; k < kMaxHeadK; ++k)
          if (k < K) dl[k] = dp[k] * pr[k] * (1.f - pr[k]);
      } else if (act == 2) {
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxHeadK; ++k)
          if (k < K) dot = fmaf(dp[k], pr[k], dot);
#pragma unroll
        for (int k = 0; k < kMaxHeadK; ++k)
          if (k < K) dl[k] = pr[k] * (dp[k] - dot);
      } else {
#pragma unroll
        for (int k = 0; k < kMaxHeadK; ++k)
          if (k < K) dl[k] = dp[k];
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxHeadK; ++k)
      if (k < K) dls[k * 256 + threadIdx.x] = dl[k];
    // dx for my pixel + stage x
    const __nv_bfloat16* xp = x + pix * x_cs;
    for (int c = 0; c < C; c += 8) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (ok) u = *reinterpret_cast<const uint4*>(xp + c);
      *reinterpret_cast<uint4*>(xs + threadIdx.x * C + c) = u;
      if (ok && dx) {
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxHeadK; ++k)
          if (k < K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaf(dl[k], ws[k * C + c + j], g[j]);
          }
        uint4 o;
        o.x = pack_bf16x2(g[0], g[1]);
        o.y = pack_bf16x2(g[2], g[3]);
        o.z = pack_bf16x2(g[4], g[5]);
        o.w = pack_bf16x2(g[6], g[7]);
        *reinterpret_cast<uint4*>(dx + pix * dx_cs + c) = o;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) {
      float a = 0.f;
      if (i < K * C) {
        const int k = i / C, c = i - k * C;
        const float* dk = dls + k * 256;
#pragma unroll 8
        for (int p = 0; p < 256; ++p) a = fmaf(dk[p], __bfloat162float(xs[p * C + c]), a);
      } else {
        const float* dk = dls + (i - K * C) * 256;
#pragma unroll 8
        for (int p = 0; p < 256; ++p) a += dk[p];
      }
      dwacc[i] += a;
    }
    __syncthreads();
  }
  if (rows_ws) {  // deterministic mode: this block's row [K*C + K], added in fixed order by msp_reduce_rows
    float* row = rows_ws + (long long)blockIdx.x * (K * C + K);
    for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) row[i] = dwacc[i];
    return;
  }
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) atomicAdd(dw + i, dwacc[i]);
  if (db)
    for (int i = threadIdx.x; i < K; i += blockDim.x) atomicAdd(db + i, dwacc[K * C + i]);
}

// ---------------------------------------------------------------------------------------------
// Dice
// ---------------------------------------------------------------------------------------------
constexpr int kMaxDiceC = 8;

// grid: (chunks, N). sums: double [G][Ceff][3] (I, Y, S)
__global__ void dice_sums_kernel(const float* __restrict__ prob, const long long* __restrict__ mask,
                                 int Cp, long long HW, int two_class, int label_offset, int batchwise,
                                 double* __restrict__ sums) {
  const int n = blockIdx.y;
  const int Ceff = two_class ? 2 : Cp;
  float aI[kMaxDiceC], aY[kMaxDiceC], aS[kMaxDiceC];
#pragma unroll
  for (int c = 0; c < kMaxDiceC; ++c) aI[c] = aY[c] = aS[c] = 0.f;
  const float* pn = prob + (long long)n * Cp * HW;
  const long long* mn = mask + (long long)n * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = mn[i] - label_offset;
    if (two_class) {
      const float p1 = pn[i], p0 = 1.f - p1;
      aS[0] = fmaf(p0, p0, aS[0]);
      aS[1] = fmaf(p1, p1, aS[1]);
      if (m == 0) { aI[0] += p0; aY[0] += 1.f; }
      if (m == 1) { aI[1] += p1; aY[1] += 1.f; }
    } else {
#pragma unroll
      for (int c = 0; c < kMaxDiceC; ++c)
        if (c < Cp) {
          const float p = pn[(long long)c * HW + i];
          aS[c] = fmaf(p, p, aS[c]);
          if (m == c) { aI[c] += p; aY[c] += 1.f; }
        }
    }
  }
  __shared__ float red[3 * kMaxDiceC][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < kMaxDiceC; ++c)
    if (c < Ceff) {
      const float i_ = warp_sum(aI[c]), y_ = warp_sum(aY[c]), s_ = warp_sum(aS[c]);
      if (lane == 0) {
        red[3 * c][warp] = i_;
        red[3 * c + 1][warp] = y_;
        red[3 * c + 2][warp] = s_;
      }
    }
  __syncthreads();
  if (threadIdx.x < 3 * Ceff) {
    double a = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) a += (double)red[threadIdx.x][w];
    const int g = batchwise ? 0 : n;
    atomicAdd(sums + (long long)g * Ceff * 3 + threadIdx.x, a);
  }
}

// 4 pixels per thread and iteration (16-byte loads of every class plane, two 16-byte loads of the int64 mask), all
// loads of an iteration issued before the arithmetic; the grid is a few blocks per SM so that a block ends with ONE
// atomic per sum — the scalar kernel above issued 768 x 12 fp64 atomics on 12 addresses and spent most of its 49 us there.
// Requires HW % 4 == 0 and 16-byte aligned planes (checked by the launcher, which otherwise keeps the scalar kernel).
__global__ void __launch_bounds__(256)
dice_sums_vec4_kernel(const float* __restrict__ prob, const long long* __restrict__ mask, int Cp, long long HW4,
                      int two_class, int label_offset, int batchwise, double* __restrict__ sums) {
  const int n = blockIdx.y;
  const int Ceff = two_class ? 2 : Cp;
  float aI[kMaxDiceC], aY[kMaxDiceC], aS[kMaxDiceC];
#pragma unroll
  for (int c = 0; c < kMaxDiceC; ++c) aI[c] = aY[c] = aS[c] = 0.f;
  const float4* pn = reinterpret_cast<const float4*>(prob + (long long)n * Cp * HW4 * 4);
  const longlong2* mn = reinterpret_cast<const longlong2*>(mask + (long long)n * HW4 * 4);
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4; i += (long long)gridDim.x * blockDim.x) {
    const longlong2 ma = mn[2 * i], mb = mn[2 * i + 1];
    float4 pv[kMaxDiceC];
#pragma unroll
    for (int c = 0; c < kMaxDiceC; ++c)
      if (c < Cp) pv[c] = ld_nc_f4(pn + (long long)c * HW4 + i);
    const long long m[4] = {ma.x - label_offset, ma.y - label_offset, mb.x - label_offset, mb.y - label_offset};
    if (two_class) {
      const float p1[4] = {pv[0].x, pv[0].y, pv[0].z, pv[0].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = 1.f - p1[e];
        aS[0] = fmaf(p0, p0, aS[0]);
        aS[1] = fmaf(p1[e], p1[e], aS[1]);
        if (m[e] == 0) { aI[0] += p0; aY[0] += 1.f; }
        if (m[e] == 1) { aI[1] += p1[e]; aY[1] += 1.f; }
      }
    } else {
#pragma unroll
      for (int c = 0; c < kMaxDiceC; ++c)
        if (c < Cp) {
          const float p[4] = {pv[c].x, pv[c].y, pv[c].z, pv[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            aS[c] = fmaf(p[e], p[e], aS[c]);
            if (m[e] == c) { aI[c] += p[e]; aY[c] += 1.f; }
          }
        }
    }
  }
  __shared__ float red[3 * kMaxDiceC][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < kMaxDiceC; ++c)
    if (c < Ceff) {
      const float i_ = warp_sum(aI[c]), y_ = warp_sum(aY[c]), s_ = warp_sum(aS[c]);
      if (lane == 0) {
        red[3 * c][warp] = i_;
        red[3 * c + 1][warp] = y_;
        red[3 * c + 2][warp] = s_;
      }
    }
  __syncthreads();
  if (threadIdx.x < 3 * Ceff) {
    double a = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) a += (double)red[threadIdx.x][w];
    const int g = batchwise ? 0 : n;
    atomicAdd(sums + (long long)g * Ceff * 3 + threadIdx.x, a);
  }
}

__global__ void __launch_bounds__(256)
dice_grad_vec4_kernel(const float* __restrict__ prob, const long long* __restrict__ mask, int Cp, long long HW4,
                      int two_class, int label_offset, int batchwise, const float* __restrict__ coef, float gscale,
                      const float* __restrict__ gdev, float* __restrict__ dprob) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  const int n = blockIdx.y;
  const int Ceff = two_class ? 2 : Cp;
  const int g = batchwise ? 0 : n;
  __shared__ float cf[kMaxDiceC * 2];
  if (threadIdx.x < Ceff * 2) cf[threadIdx.x] = coef[(long long)g * Ceff * 2 + threadIdx.x] * gscale;
  __syncthreads();
  const float4* pn = reinterpret_cast<const float4*>(prob + (long long)n * Cp * HW4 * 4);
  float4* dn = reinterpret_cast<float4*>(dprob + (long long)n * Cp * HW4 * 4);
  const longlong2* mn = reinterpret_cast<const longlong2*>(mask + (long long)n * HW4 * 4);
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4; i += (long long)gridDim.x * blockDim.x) {
    const longlong2 ma = mn[2 * i], mb = mn[2 * i + 1];
    float4 pv[kMaxDiceC];
#pragma unroll
    for (int c = 0; c < kMaxDiceC; ++c)
      if (c < Cp) pv[c] = ld_nc_f4(pn + (long long)c * HW4 + i);
    const long long m[4] = {ma.x - label_offset, ma.y - label_offset, mb.x - label_offset, mb.y - label_offset};
    if (two_class) {
      const float p1[4] = {pv[0].x, pv[0].y, pv[0].z, pv[0].w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = 1.f - p1[e];
        const float d0 = (m[e] == 0 ? cf[0] : 0.f) + cf[1] * p0;
        const float d1 = (m[e] == 1 ? cf[2] : 0.f) + cf[3] * p1[e];
        o[e] = d1 - d0;
      }
      dn[i] = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int c = 0; c < kMaxDiceC; ++c)
        if (c < Cp) {
          const float p[4] = {pv[c].x, pv[c].y, pv[c].z, pv[c].w};
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = (m[e] == c ? cf[2 * c] : 0.f) + cf[2 * c + 1] * p[e];
          dn[(long long)c * HW4 + i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
  }
}

// loss = 1 - mean_{g, c >= c0} (2I+eps)/(Y+S+eps); also the per-(g,c) gradient coefficients
// coef[g][c] = { a = -2/(M*D), b = (2I+eps)*2/(M*D^2) } so that dL/dp = a*y + b*p  (times gscale later)
__global__ void dice_finalize_kernel(const double* __restrict__ sums, int G, int Ceff, int c0,
                                     float eps, float* __restrict__ loss, float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int M = G * (Ceff - c0);
  double acc = 0.0;
  for (int g = 0; g < G; ++g)
    for (int c = 0; c < Ceff; ++c) {
      const double* s = sums + ((long long)g * Ceff + c) * 3;
      // the reference does these in fp32 (losses.py:37-40)
      const float I = (float)s[0], Y = (float)s[1], S = (float)s[2];
      const float num = 2.f * I + eps, den = Y + S + eps;
      float a = 0.f, b = 0.f;
      if (c >= c0) {
        acc += (double)(num / den);
        a = -2.f / ((float)M * den);
        b = 2.f * num / ((float)M * den * den);
      }
      if (coef) {
        coef[((long long)g * Ceff + c) * 2] = a;
        coef[((long long)g * Ceff + c) * 2 + 1] = b;
      }
    }
  if (loss) *loss = (float)(1.0 - acc / (double)M);
}

__global__ void dice_grad_kernel(const float* __restrict__ prob, const long long* __restrict__ mask,
                                 int Cp, long long HW, int two_class, int label_offset, int batchwise,
                                 const float* __restrict__ coef, float gscale,
                                 const float* __restrict__ gdev, float* __restrict__ dprob) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  const int n = blockIdx.y;
  const int Ceff = two_class ? 2 : Cp;
  const int g = batchwise ? 0 : n;
  __shared__ float cf[kMaxDiceC * 2];
  if (threadIdx.x < Ceff * 2) cf[threadIdx.x] = coef[(long long)g * Ceff * 2 + threadIdx.x] * gscale;
  __syncthreads();
  const float* pn = prob + (long long)n * Cp * HW;
  float* dn = dprob + (long long)n * Cp * HW;
  const long long* mn = mask + (long long)n * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = mn[i] - label_offset;
    if (two_class) {
      const float p1 = pn[i], p0 = 1.f - p1;
      const float d0 = (m == 0 ? cf[0] : 0.f) + cf[1] * p0;
      const float d1 = (m == 1 ? cf[2] : 0.f) + cf[3] * p1;
      dn[i] = d1 - d0;
    } else {
#pragma unroll
      for (int c = 0; c < kMaxDiceC; ++c)
        if (c < Cp) {
          const float p = pn[(long long)c * HW + i];
          dn[(long long)c * HW + i] = (m == c ? cf[2 * c] : 0.f) + cf[2 * c + 1] * p;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CE on probabilities, BCE, softmax-CE
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_to(double v, double* out) {
  __shared__ double redd[8];
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) redd[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0 && out != nullptr) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += redd[w];
    atomicAdd(out, a);
  }
}

// loss_sum (double) += sum_pix -sum_c clamp(nan_to_num(log p_c), -100) * t_c ; dprob = gscale * dL/dp
__global__ void ce_prob_kernel(const float* __restrict__ prob, const long long* __restrict__ label,
                               int C, long long HW, float lo, float hi, int smooth_on, float gscale,
                               const float* __restrict__ gdev, double* __restrict__ loss_sum,
                               float* __restrict__ dprob) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  const int n = blockIdx.y;
  const float* pn = prob + (long long)n * C * HW;
  float* dn = dprob ? dprob + (long long)n * C * HW : nullptr;
  const long long* ln = label + (long long)n * HW;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
       i += (long long)gridDim.x * blockDim.x) {
    const long long y = ln[i];
    for (int c = 0; c < C; ++c) {
      const float p = pn[(long long)c * HW + i];
      float t = (y == c) ? 1.f : 0.f;
      if (smooth_on) t = fminf(fmaxf(t, lo), hi);
      float lp = logf(p);
      bool live = true;  // gradient flows only through finite, un-clamped log values
      if (lp != lp) { lp = 0.f; live = false; }
      else if (isinf(lp)) { lp = lp > 0 ? 3.4028234664e38f : -3.4028234664e38f; live = false; }
      if (lp < -100.f) { lp = -100.f; live = false; }
      acc = fmaf(-lp, t, acc);
      if (dn) {
        // autograd of log -> nan_to_num -> clamp: 0 * (1/p) where the chain is cut, i.e. NaN at p == 0 / NaN
        float gr = live ? -t / p * gscale : 0.f;
        if (p == 0.f || p != p) gr = __int_as_float(0x7fc00000);
        dn[(long long)c * HW + i] = gr;
      }
    }
  }
  block_sum_to((double)acc, loss_sum);
}

// mode 0: reference BCELoss (no clamp, autograd gradient); mode 1: torch.nn.BCELoss (log clamped at
// -100, gradient (p - y) / max(p(1-p), 1e-12))
__global__ void bce_kernel(const float* __restrict__ prob, const float* __restrict__ target,
                           long long numel, int clamp_log, float gscale, const float* __restrict__ gdev,
                           double* __restrict__ loss_sum, float* __restrict__ dprob) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x) {
    const float p = prob[i], y = target[i];
    float l1 = logf(p), l0 = logf(1.f - p);
    float g;
    if (clamp_log) {
      l1 = fmaxf(l1, -100.f);
      l0 = fmaxf(l0, -100.f);
      g = (p - y) / fmaxf((1.f - p) * p, 1e-12f);
    } else {
      g = -(y / p - (1.f - y) / (1.f - p));
    }
    acc -= y * l1 + (1.f - y) * l0;
    if (dprob) dprob[i] = g * gscale;
  }
  block_sum_to((double)acc, loss_sum);
}

__global__ void __launch_bounds__(256)
bce_vec4_kernel(const float* __restrict__ prob, const float* __restrict__ target, long long numel4, int clamp_log,
                float gscale, const float* __restrict__ gdev, double* __restrict__ loss_sum, float* __restrict__ dprob) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  float acc = 0.f;
  const float4* p4 = reinterpret_cast<const float4*>(prob);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float4* d4 = reinterpret_cast<float4*>(dprob);
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel4; i += (long long)gridDim.x * blockDim.x) {
    const float4 pq = ld_nc_f4(p4 + i), tq = ld_nc_f4(t4 + i);
    const float pv[4] = {pq.x, pq.y, pq.z, pq.w}, yv[4] = {tq.x, tq.y, tq.z, tq.w};
    float gv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p = pv[e], y = yv[e];
      float l1 = logf(p), l0 = logf(1.f - p);
      if (clamp_log) {
        l1 = fmaxf(l1, -100.f);
        l0 = fmaxf(l0, -100.f);
        gv[e] = (p - y) / fmaxf((1.f - p) * p, 1e-12f);
      } else {
        gv[e] = -(y / p - (1.f - y) / (1.f - p));
      }
      acc -= y * l1 + (1.f - y) * l0;
    }
    if (dprob) d4[i] = make_float4(gv[0] * gscale, gv[1] * gscale, gv[2] * gscale, gv[3] * gscale);
  }
  block_sum_to((double)acc, loss_sum);
}

// one block per row: F.cross_entropy(logits, label, label_smoothing) summed over rows
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ label,
                                  int C, float smooth, float gscale, const float* __restrict__ gdev,
                                  double* __restrict__ loss_sum, float* __restrict__ dlogits) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  const int n = blockIdx.x;
  const float* z = logits + (long long)n * C;
  __shared__ float redf[8];
  __shared__ float bc[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, z[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) redf[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = redf[0];
    for (int w = 1; w < nw; ++w) a = fmaxf(a, redf[w]);
    bc[0] = a;
  }
  __syncthreads();
  m = bc[0];
  float s = 0.f, sz = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s += expf(z[c] - m);
    sz += z[c];
  }
  s = warp_sum(s);
  sz = warp_sum(sz);
  __syncthreads();
  if (lane == 0) redf[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += redf[w];
    bc[1] = a;
  }
  __syncthreads();
  s = bc[1];
  __syncthreads();
  if (lane == 0) redf[warp] = sz;
  __syncthreads();
  const float lse = m + logf(s);
  const long long y = label[n];
  if (threadIdx.x == 0) {
    float tz = 0.f;
    for (int w = 0; w < nw; ++w) tz += redf[w];
    // -(1-s) logp[y] - s/C sum_c logp[c]
    const float nll = (y >= 0 && y < C) ? lse - z[y] : 0.f;
    const float mean_nlp = lse - tz / (float)C;
    if (loss_sum) atomicAdd(loss_sum, (double)((1.f - smooth) * nll + smooth * mean_nlp));
  }
  if (dlogits) {
    float* d = dlogits + (long long)n * C;
    const float off = smooth / (float)C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float p = expf(z[c] - lse);
      const float t = off + ((c == y) ? (1.f - smooth) : 0.f);
      d[c] = (p - t) * gscale;
    }
  }
}

__global__ void scale_double_to_float_kernel(const double* __restrict__ in, double scale,
                                             float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = (float)(*in * scale);
}

// vectorised kernels: plane length a multiple of 4 elements and a 16-byte aligned base
inline bool vec4_ok(const void* p, long long plane) { return plane % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// blocks per row for a kernel whose thread handles one 16-byte group per iteration: ~4 blocks of 256 threads per SM
// over all rows, never more blocks than groups / 256
inline int vec_grid(long long groups, int rows) {
  long long cap = ((long long)msp_num_sms() * 4 + rows - 1) / rows;
  long long need = (groups + 255) / 256;
  long long b = need < cap ? need : cap;
  return (int)(b < 1 ? 1 : b);
}
inline int strip_grid(long long work, int threads, int rows) {
  long long b = (work + (long long)threads * 8 - 1) / ((long long)threads * 8);
  long long cap = ((long long)msp_num_sms() * 8 + rows - 1) / rows;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}


// torch.nn.CrossEntropyLoss with CLASS-PROBABILITY targets (Mixup / CutMix of config/pretraining/resnet50/advanced.yaml:17,48):
// t' = (1 - smooth) t + smooth / C;  loss_n = lse * sum(t') - sum_c t'_c z_c;  dz_c = p_c * sum(t') - t'_c.  One block per row.
__global__ void softmax_ce_soft_kernel(const float* __restrict__ logits, const float* __restrict__ target, int C,
                                       float smooth, float gscale, const float* __restrict__ gdev,
                                       double* __restrict__ loss_sum, float* __restrict__ dlogits) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  const int n = blockIdx.x;
  const float* z = logits + (long long)n * C;
  const float* t = target + (long long)n * C;
  __shared__ float red[3][8];
  __shared__ float bc[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) m = fmaxf(m, z[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = red[0][0];
    for (int w = 1; w < nw; ++w) a = fmaxf(a, red[0][w]);
    bc[0] = a;
  }
  __syncthreads();
  m = bc[0];
  const float off = smooth / (float)C;
  float s = 0.f, st = 0.f, stz = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float tp = fmaf(1.f - smooth, t[c], off);
    s += expf(z[c] - m);
    st += tp;
    stz = fmaf(tp, z[c], stz);
  }
  s = warp_sum(s);
  st = warp_sum(st);
  stz = warp_sum(stz);
  __syncthreads();
  if (lane == 0) { red[0][warp] = s; red[1][warp] = st; red[2][warp] = stz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c2 = 0.f;
    for (int w = 0; w < nw; ++w) { a += red[0][w]; b += red[1][w]; c2 += red[2][w]; }
    bc[1] = a; bc[2] = b; bc[3] = c2;
  }
  __syncthreads();
  const float lse = m + logf(bc[1]), sum_t = bc[2];
  if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, (double)(lse * sum_t - bc[3]));
  if (dlogits) {
    float* d = dlogits + (long long)n * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      d[c] = (expf(z[c] - lse) * sum_t - fmaf(1.f - smooth, t[c], off)) * gscale;
  }
}

// F.cross_entropy on SPATIAL logits (N, C, HW) with class-index targets (N, HW), label smoothing as above: one thread per
// pixel, the class loop strides by HW (coalesced across the pixels of a warp); targets outside [0, C) contribute nothing
// to the NLL term (ignore_index is not implemented and is refused by the host wrapper).
__global__ void __launch_bounds__(256)
softmax_ce_spatial_kernel(const float* __restrict__ logits, const long long* __restrict__ label, int C, long long HW,
                          long long P, float smooth, float gscale, const float* __restrict__ gdev,
                          double* __restrict__ loss_sum, float* __restrict__ dlogits) {
  gscale *= gdev ? __ldg(gdev) : 1.f;
  double acc = 0.0;
  const float off = smooth / (float)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, hw = i - n * HW;
    const float* z = logits + n * C * HW + hw;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, z[(long long)c * HW]);
    float s = 0.f, sz = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = z[(long long)c * HW];
      s += expf(v - m);
      sz += v;
    }
    const float lse = m + logf(s);
    const long long y = label[i];
    const bool ok = y >= 0 && y < C;
    if (loss_sum) {
      const float nll = ok ? lse - z[y * HW] : 0.f;
      acc += (double)((1.f - smooth) * nll + smooth * (lse - sz / (float)C));
    }
    if (dlogits) {
      float* d = dlogits + n * C * HW + hw;
      for (int c = 0; c < C; ++c) {
        const float p = expf(z[(long long)c * HW] - lse);
        const float t = off + ((ok && c == y) ? (1.f - smooth) : 0.f);
        d[(long long)c * HW] = (p - t) * gscale;
      }
    }
  }
  if (loss_sum) {
    acc = warp_sum_d(acc);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += part[w];
      atomicAdd(loss_sum, t);
    }
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int msp_final_conv_act_fwd(const void* x, int N, int H, int W, int C, int x_cs,
                                      const float* w, const float* bias, int K, int act,
                                      float* logits_nchw, float* prob_nchw, void* stream) {
  MSP_REQUIRE(x && w && prob_nchw, "final_conv_act_fwd: null pointer");
  MSP_REQUIRE(K >= 1 && K <= kMaxHeadK, "final_conv_act_fwd: K=%d outside [1,8]", K);
  MSP_REQUIRE(C > 0 && C % 8 == 0 && x_cs % 8 == 0 && x_cs >= C, "final_conv_act_fwd: C/x_cs %% 8");
  MSP_REQUIRE((long long)K * C <= 8192, "final_conv_act_fwd: K*C too large for shared memory");
  MSP_REQUIRE(act >= 0 && act <= 2, "final_conv_act_fwd: bad activation");
  const long long HW = (long long)H * W, P = (long long)N * HW;
  if (P == 0) return MSP_OK;
  const size_t smem = (size_t)(K * C + K) * sizeof(float);
  long long blocks = (P + 255) / 256;
  if (blocks > (long long)msp_num_sms() * 16) blocks = (long long)msp_num_sms() * 16;
  final_conv_act_fwd_kernel<<<(int)blocks, 256, smem, ST>>>((const __nv_bfloat16*)x, P, HW, C, x_cs, w,
                                                            bias, K, act, logits_nchw, prob_nchw);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_final_conv_act_bwd(const void* x, int N, int H, int W, int C, int x_cs,
                                      const float* w, int K, int act, const float* prob_nchw,
                                      const float* dprob_nchw, void* dx, int dx_cs, float* dw,
                                      float* db, float* rows_ws, int ws_rows, void* stream) {
  MSP_REQUIRE(x && w && prob_nchw && dprob_nchw && dw, "final_conv_act_bwd: null pointer");
  MSP_REQUIRE(rows_ws == nullptr || (ws_rows >= 1 && db == dw + (long long)K * C),
              "final_conv_act_bwd: deterministic mode needs db = dw + K*C (one [K*C + K] result vector)");
  MSP_REQUIRE(K >= 1 && K <= kMaxHeadK, "final_conv_act_bwd: K=%d outside [1,8]", K);
  MSP_REQUIRE(C > 0 && C % 8 == 0 && x_cs % 8 == 0 && x_cs >= C, "final_conv_act_bwd: C/x_cs %% 8");
  MSP_REQUIRE(dx == nullptr || (dx_cs % 8 == 0 && dx_cs >= C), "final_conv_act_bwd: dx_cs");
  const long long HW = (long long)H * W, P = (long long)N * HW;
  if (rows_ws == nullptr || P == 0) {
    MSP_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * K * C, ST));
    if (db) MSP_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * K, ST));
  }
  if (P == 0) return MSP_OK;
  const size_t smem = (size_t)((2 * K * C + K * 256 + K + 3) & ~3) * sizeof(float) + (size_t)256 * C * 2;
  MSP_REQUIRE(smem <= 200 * 1024, "final_conv_act_bwd: C=%d too large for the staged tile", C);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(final_conv_act_bwd_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  long long blocks = (P + 255) / 256;
  if (blocks > (long long)msp_num_sms() * 4) blocks = (long long)msp_num_sms() * 4;
  if (rows_ws != nullptr && blocks > ws_rows) blocks = ws_rows;
  final_conv_act_bwd_kernel<<<(int)blocks, 256, smem, ST>>>(
      (const __nv_bfloat16*)x, P, HW, C, x_cs, w, K, act, prob_nchw, dprob_nchw, (__nv_bfloat16*)dx,
      dx_cs, dw, db, rows_ws);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  if (rows_ws != nullptr) return msp_reduce_rows(rows_ws, (int)blocks, K * C + K, dw, 0, stream);
  return MSP_OK;
}

extern "C" int msp_dice_sums(const float* prob, const int64_t* mask, int N, int Cp, long long HW,
                             int two_class, int label_offset, int batchwise, double* sums, void* stream) {
  MSP_REQUIRE(prob && mask && sums, "dice_sums: null pointer");
  const int Ceff = two_class ? 2 : Cp;
  MSP_REQUIRE(Cp >= 1 && Ceff <= kMaxDiceC && (!two_class || Cp == 1), "dice_sums: %d classes unsupported", Cp);
  MSP_REQUIRE(N > 0 && HW > 0, "dice_sums: empty prediction");
  const int G = batchwise ? 1 : N;
  MSP_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * G * Ceff * 3, ST));
  if (vec4_ok(prob, HW) && vec4_ok(mask, HW)) {
    dim3 grid(vec_grid(HW / 4, N), N);
    dice_sums_vec4_kernel<<<grid, 256, 0, ST>>>(prob, (const long long*)mask, Cp, HW / 4, two_class, label_offset,
                                                batchwise, sums);
  } else {
    dim3 grid(strip_grid(HW, 256, N), N);
    dice_sums_kernel<<<grid, 256, 0, ST>>>(prob, (const long long*)mask, Cp, HW, two_class, label_offset,
                                           batchwise, sums);
  }
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_dice_finalize(const double* sums, int G, int Ceff, int class_start, float eps,
                                 float* coef, float* loss, void* stream) {
  MSP_REQUIRE(sums && (coef || loss), "dice_finalize: null pointer");
  MSP_REQUIRE(G >= 1 && Ceff >= 1 && Ceff <= kMaxDiceC && class_start >= 0 && class_start < Ceff,
              "dice_finalize: bad class range");
  dice_finalize_kernel<<<1, 32, 0, ST>>>(sums, G, Ceff, class_start, eps, loss, coef);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_dice_bwd(const float* prob, const int64_t* mask, int N, int Cp, long long HW,
                            int two_class, int label_offset, int batchwise, const float* coef,
                            float gscale, const float* gscale_dev, float* dprob, void* stream) {
  MSP_REQUIRE(prob && mask && coef && dprob, "dice_bwd: null pointer");
  const int Ceff = two_class ? 2 : Cp;
  MSP_REQUIRE(Cp >= 1 && Ceff <= kMaxDiceC, "dice_bwd: %d classes unsupported", Cp);
  if (vec4_ok(prob, HW) && vec4_ok(mask, HW) && vec4_ok(dprob, HW)) {
    dim3 grid(vec_grid(HW / 4, N), N);
    dice_grad_vec4_kernel<<<grid, 256, 0, ST>>>(prob, (const long long*)mask, Cp, HW / 4, two_class, label_offset,
                                                batchwise, coef, gscale, gscale_dev, dprob);
  } else {
    dim3 grid(strip_grid(HW, 256, N), N);
    dice_grad_kernel<<<grid, 256, 0, ST>>>(prob, (const long long*)mask, Cp, HW, two_class, label_offset,
                                           batchwise, coef, gscale, gscale_dev, dprob);
  }
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_ce_prob_fwd_bwd(const float* prob, const int64_t* label, int N, int C, long long HW,
                                   float smooth, float gscale, const float* gscale_dev,
                                   double* loss_sum, float* dprob, void* stream) {
  MSP_REQUIRE(prob && label && (loss_sum || dprob), "ce_prob: null pointer");
  MSP_REQUIRE(N > 0 && C > 0 && HW > 0, "ce_prob: empty prediction");
  if (loss_sum) MSP_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
  dim3 grid(strip_grid(HW, 256, N), N);
  const float lo = smooth / (float)C, hi = 1.f - smooth / (float)C;
  ce_prob_kernel<<<grid, 256, 0, ST>>>(prob, (const long long*)label, C, HW, lo, hi, smooth != 0.f,
                                       gscale, gscale_dev, loss_sum, dprob);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_bce_fwd_bwd(const float* prob, const float* target, long long numel, int clamp_log,
                               float gscale, const float* gscale_dev, double* loss_sum, float* dprob,
                               void* stream) {
  MSP_REQUIRE(prob && target && (loss_sum || dprob) && numel > 0, "bce: bad arguments");
  if (loss_sum) MSP_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
  if (vec4_ok(prob, numel) && vec4_ok(target, numel) && (!dprob || vec4_ok(dprob, numel)))
    bce_vec4_kernel<<<vec_grid(numel / 4, 1), 256, 0, ST>>>(prob, target, numel / 4, clamp_log, gscale, gscale_dev,
                                                            loss_sum, dprob);
  else
    bce_kernel<<<strip_grid(numel, 256, 1), 256, 0, ST>>>(prob, target, numel, clamp_log, gscale, gscale_dev, loss_sum,
                                                          dprob);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_softmax_ce_soft_fwd_bwd(const float* logits, const float* target, int N, int C, float smooth,
                                           float gscale, const float* gscale_dev, double* loss_sum, float* dlogits,
                                           void* stream) {
  MSP_REQUIRE(logits && target && (loss_sum || dlogits) && N > 0 && C > 0, "softmax_ce_soft: bad arguments");
  if (loss_sum) MSP_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
  softmax_ce_soft_kernel<<<N, 256, 0, ST>>>(logits, target, C, smooth, gscale, gscale_dev, loss_sum, dlogits);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_softmax_ce_spatial_fwd_bwd(const float* logits, const int64_t* label, int N, int C, long long HW,
                                              float smooth, float gscale, const float* gscale_dev, double* loss_sum,
                                              float* dlogits, void* stream) {
  MSP_REQUIRE(logits && label && (loss_sum || dlogits) && N > 0 && C > 0 && HW > 0, "softmax_ce_spatial: bad arguments");
  if (loss_sum) MSP_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
  const long long P = (long long)N * HW;
  long long blocks = (P + 255) / 256;
  if (blocks > (long long)msp_num_sms() * 8) blocks = (long long)msp_num_sms() * 8;
  softmax_ce_spatial_kernel<<<(int)blocks, 256, 0, ST>>>(logits, (const long long*)label, C, HW, P, smooth, gscale,
                                                         gscale_dev, loss_sum, dlogits);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_softmax_ce_fwd_bwd(const float* logits, const int64_t* label, int N, int C,
                                      float smooth, float gscale, const float* gscale_dev,
                                      double* loss_sum, float* dlogits, void* stream) {
  MSP_REQUIRE(logits && label && (loss_sum || dlogits) && N > 0 && C > 0, "softmax_ce: bad arguments");
  if (loss_sum) MSP_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), ST));
  softmax_ce_kernel<<<N, 256, 0, ST>>>(logits, (const long long*)label, C, smooth, gscale, gscale_dev, loss_sum,
                                       dlogits);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_scale_to_float(const double* in, double scale, float* out, void* stream) {
  MSP_REQUIRE(in && out, "scale_to_float: null pointer");
  scale_double_to_float_kernel<<<1, 32, 0, ST>>>(in, scale, out);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
