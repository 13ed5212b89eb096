// HBM-bound NHWC bf16 kernels: layout conversion, fused BatchNorm(+act+residual) forward/backward,
// pooling, nearest up-sampling, attention-gate product, channel-slice copies.
// All use 16-byte vector accesses (8 bf16 channels per thread), coalesced along channels.
//
// Replaces (reference): nn.BatchNorm2d + nn.ReLU + zero-fill/stride-2 shortcut + DropPath
// (classification/models.py:203-212,257-290), ConvBlock BN+act (segmentation/models/blocks.py:458-488),
// nn.MaxPool2d (classification/models.py:56; unet_models.py:452), nn.Upsample (blocks.py:532,615),
// skip*p and torch.cat (blocks.py:624-628,635).
#include "msp_common.cuh"
#include <stdlib.h>
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 load_bf16x8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const F8& f) {
  uint4 u;
  u.x = pack_bf16x2(f.v[0], f.v[1]);
  u.y = pack_bf16x2(f.v[2], f.v[3]);
  u.z = pack_bf16x2(f.v[4], f.v[5]);
  u.w = pack_bf16x2(f.v[6], f.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ F8 load_f32x8(const float* p) {
  F8 r;
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

inline int grid_for(long long work_items, int per_block, int max_waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)msp_num_sms() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// persistent grid-stride kernels: exactly as many blocks as are co-resident (a partial second wave idles the GPU)
template <typename K>
inline int resident_grid(K kernel, int threads, size_t smem, long long work_items, int per_block) {
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)msp_num_sms() * occ;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
// threads per block such that blockDim % V == 0 (V = 8-channel vectors per pixel)
inline int threads_for_vecs(int V) {
  if (V >= 256) return 256;
  return V * (256 / V);
}

// ---------------------------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 via a 32x32 smem transpose of (c, hw) per image
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, long long HW, int Cpad,
                                    __nv_bfloat16* __restrict__ y) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long hw0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xn = x + (long long)n * C * HW;
  __nv_bfloat16* yn = y + (long long)n * HW * Cpad;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long hw = hw0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && hw < HW) ? xn[(long long)c * HW + hw] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long hw = hw0 + i;
    const int c = c0 + threadIdx.x;
    if (hw < HW && c < Cpad) yn[hw * Cpad + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
// NCHW fp32 -> W-padded NHWC bf16 [N][H][Wp][cpp] for the row-window convolution (msp_conv.cu): pixel w of
// the image lands at column w + pad_l, everything else (pad columns, channels >= C) is zero.
__device__ __forceinline__ float src_to_f32(float v) { return v; }
__device__ __forceinline__ float src_to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>  // T = float (the reference's batches) or __nv_bfloat16 (half the host -> device bytes)
__global__ void nchw_to_rowwin_kernel(const T* __restrict__ x, int C, int H, int W, int cpp, int pad_l,
                                      int Wp, long long total, __nv_bfloat16* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    const long long t = i / Wp;
    const int h = (int)(t % H);
    const long long n = t / H;
    const int w = wp - pad_l;
    const bool in = w >= 0 && w < W;
    const T* src = x + (n * C * H + h) * (long long)W + w;
    uint32_t pk[8];
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const float a = (in && c < C) ? src_to_f32(src[(long long)c * H * W]) : 0.f;
      const float b = (in && c + 1 < C) ? src_to_f32(src[(long long)(c + 1) * H * W]) : 0.f;
      pk[c >> 1] = pack_bf16x2(a, b);
    }
    uint4* dst = reinterpret_cast<uint4*>(y + i * cpp);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    if (cpp == 16) dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}
// NHWC bf16 (pixel stride cs) -> NCHW fp32
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int C, long long HW, int cs,
                                    float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long hw0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const __nv_bfloat16* xn = x + (long long)n * HW * cs;
  float* yn = y + (long long)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long hw = hw0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (hw < HW && c < C) ? __bfloat162float(xn[hw * cs + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) yn[(long long)c * HW + hw] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm
// ---------------------------------------------------------------------------------------------
// (sum_rows_fixed / kRowGroups: msp_common.cuh — shared with the peer-memory statistic exchange, msp_p2p.cu)
__global__ void __launch_bounds__(32 * kRowGroups) reduce_rows_kernel(float* ws, int rows, int n, float* out, int reset) {
  __shared__ float part[kRowGroups][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
  const float tot = sum_rows_fixed(ws, rows, n, col < n ? col : n - 1, grp, reset && col < n, part);
  if (grp == 0 && col < n) out[col] = tot;
}

// block = 32 channels x 8 row groups; rows == 1: the plain [2][C] accumulators
__global__ void __launch_bounds__(32 * kRowGroups)
bn_finalize_kernel(float* s1, float* s2, int C, double count, float eps, float momentum,
                   float* mean, float* invstd, float* rmean, float* rvar, int reset, int rows) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[kRowGroups][33];
  __shared__ float part2[kRowGroups][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
  const int cc = c < C ? c : C - 1;
  const long long stride = 2ll * C;
  float t1, t2;
  if (s2 == s1 + C) {   // the usual [rows][2][C] layout: both statistics in one pass
    const float2 t = sum_rows_fixed2(s1, rows, stride, cc, C, grp, reset && c < C, part, part2);
    t1 = t.x;
    t2 = t.y;
  } else {
    t1 = sum_rows_fixed(s1, rows, stride, cc, grp, reset && c < C, part);
    t2 = sum_rows_fixed(s2, rows, stride, cc, grp, reset && c < C, part);
  }
  if (grp != 0 || c >= C) return;
  const double m = (double)t1 / count;
  double var = (double)t2 / count - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (rmean != nullptr) {
    const double unb = count > 1 ? var * count / (count - 1) : var;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}
__global__ void bn_eval_prepare_kernel(const float* __restrict__ rvar, int C, float eps,
                                       float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) invstd[c] = 1.f / sqrtf(rvar[c] + eps);
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == MSP_ACT_RELU) return fmaxf(x, 0.f);
  if (act == MSP_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}
// derivative of act expressed through the stored output y
__device__ __forceinline__ float act_bwd(float y, float dy, int act) {
  if (act == MSP_ACT_RELU) return y > 0.f ? dy : 0.f;
  if (act == MSP_ACT_SIGMOID) return dy * y * (1.f - y);
  return dy;
}

__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
  r.v[4] = c.x; r.v[5] = c.y; r.v[6] = d.x; r.v[7] = d.y;
  return r;
}
// pixel index of the (strided, sub-sampled) shortcut tensor that output pixel `pix` of image `n` reads
__device__ __forceinline__ long long shortcut_pixel(const msp_bn_act_desc& d, uint32_t pix, uint32_t n, uint32_t HW) {
  if (d.r_stride <= 1) return pix;
  const uint32_t hw = pix - n * HW;
  const uint32_t h = hw / (uint32_t)d.W, w = hw - h * (uint32_t)d.W;
  return ((long long)n * d.H * d.r_stride + (long long)h * d.r_stride) * ((long long)d.W * d.r_stride) +
         (long long)w * d.r_stride;
}

// The three BatchNorm kernels share one thread mapping: thread -> (8-channel vector cg, pixel lane pl); a block
// sweeps pixels with stride gridDim*ppb and every thread keeps U independent 16-byte loads per tensor in flight
// (read-once tensors: ld.global.nc.L1::no_allocate).  Pixel counts are < 2^31 (checked on the host): the per-image
// index needed by DropPath / strided shortcuts is one 32-bit division, skipped entirely when unused.
template <int U>
__global__ void __launch_bounds__(256, 2)
bn_act_fwd_kernel(const msp_bn_act_desc d, const __nv_bfloat16* __restrict__ x,
                  const float* __restrict__ mean, const float* __restrict__ invstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ sscale, const __nv_bfloat16* __restrict__ res,
                  __nv_bfloat16* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int V = d.C >> 3;
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = cg * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float g = gamma ? gamma[c + j] : 1.f, b = beta ? beta[c + j] : 0.f;
    sc[j] = g * invstd[c + j];
    sh[j] = b - mean[c + j] * sc[j];
  }
  const uint32_t HW = (uint32_t)d.H * (uint32_t)d.W;
  const uint32_t P = (uint32_t)d.N * HW;
  const bool has_res = res != nullptr && c < d.r_C;
  const bool need_n = sscale != nullptr || (has_res && d.r_stride > 1);
  const uint32_t step = gridDim.x * (uint32_t)ppb;
  for (uint32_t p0 = blockIdx.x * (uint32_t)ppb + pl; p0 < P; p0 += step * U) {
    uint4 a[U], r[U];  // raw bf16x8 until the math (half the registers of unpacked floats)
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + u * step;
      s[u] = 1.f;
      if (pix < P) {
        a[u] = ld_nc_v4(x + (long long)pix * d.x_cs + c);
        uint32_t n = 0;
        if (need_n) {
          n = pix / HW;
          if (sscale) s[u] = __ldg(sscale + n);
        }
        if (has_res) r[u] = ld_nc_v4(res + shortcut_pixel(d, pix, n, HW) * d.r_cs + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + u * step;
      if (pix < P) {
        F8 o;
        const F8 av = unpack8(a[u]);
        F8 rv;
        if (has_res) rv = unpack8(r[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(av.v[j], sc[j], sh[j]) * s[u];
          if (has_res) t += rv.v[j];
          o.v[j] = act_fwd(t, d.act);
        }
        store_bf16x8(y + (long long)pix * d.y_cs + c, o);
      }
    }
  }
}

// pass 1 of the backward: per-channel sum(s*g) and sum(s*g*xhat)
template <int U>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_kernel(const msp_bn_act_desc d, const __nv_bfloat16* __restrict__ x,
                         const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ sscale, float* sum_g, float* sum_gx,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float* rows_ws) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];
  const int V = d.C >> 3;
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = cg * 8;
  // y == nullptr (BatchNorm -> ReLU without shortcut / sample scale): the ReLU mask is recomputed from x with the
  // forward's own expression fmaf(x, sc, sh) > 0, and the stored activation is not read at all
  const bool from_x = (y == nullptr);
  float mu[8], is[8], a1[8], a2[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[c + j];
    is[j] = invstd[c + j];
    a1[j] = 0.f;
    a2[j] = 0.f;
    sc[j] = (gamma ? gamma[c + j] : 1.f) * is[j];
    sh[j] = (beta ? beta[c + j] : 0.f) - mu[j] * sc[j];
  }
  const uint32_t HW = (uint32_t)d.H * (uint32_t)d.W;
  const uint32_t P = (uint32_t)d.N * HW;
  const uint32_t step = gridDim.x * (uint32_t)ppb;
  for (uint32_t p0 = blockIdx.x * (uint32_t)ppb + pl; p0 < P; p0 += step * U) {
    uint4 xr[U], yr[U], gr[U];
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + u * step;
      s[u] = 0.f;
      if (pix < P) {
        xr[u] = ld_nc_v4(x + (long long)pix * d.x_cs + c);
        if (!from_x) yr[u] = ld_nc_v4(y + (long long)pix * d.y_cs + c);
        gr[u] = ld_nc_v4(dy + (long long)pix * d.y_cs + c);
        s[u] = sscale ? __ldg(sscale + pix / HW) : 1.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u * step < P) {
        const F8 xv = unpack8(xr[u]), gv = unpack8(gr[u]);
        F8 yv;
        if (from_x) {
#pragma unroll
          for (int j = 0; j < 8; ++j) yv.v[j] = fmaf(xv.v[j], sc[j], sh[j]);  // only its sign is used (ReLU)
        } else {
          yv = unpack8(yr[u]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = act_bwd(yv.v[j], gv.v[j], d.act) * s[u];
          a1[j] += g;
          a2[j] = fmaf(g, (xv.v[j] - mu[j]) * is[j], a2[j]);
        }
      }
    }
  }
  // block reduction over pixel lanes: red[pl][cg*16 + j]
  float* mine = red + ((long long)pl * V + cg) * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = a1[j];
    mine[8 + j] = a2[j];
  }
  __syncthreads();
  // thread t sums column t of the [ppb][V*16] table (coalesced, conflict-free), one atomic per column
  for (int col = threadIdx.x; col < V * 16; col += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < ppb; ++l) acc += red[(long long)l * V * 16 + col];
    const int ch = (col >> 4) * 8 + (col & 7);
    // deterministic mode: this block's own row [2][C] of the workspace (added in fixed order by reduce_rows_kernel)
    if (rows_ws) rows_ws[(long long)blockIdx.x * 2 * d.C + ((col & 8) ? d.C : 0) + ch] = acc;
    else atomicAdd(((col & 8) ? sum_gx : sum_g) + ch, acc);
  }
}

template <int U>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_kernel(const msp_bn_act_desc d, const __nv_bfloat16* __restrict__ x,
                        const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ sscale,
                        const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float inv_count,
                        __nv_bfloat16* __restrict__ dx, __nv_bfloat16* dres, int dres_acc,
                        const float* __restrict__ beta) {
  pdl_trigger();
  pdl_wait();
  const int V = d.C >> 3;
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  const int c = cg * 8;
  const bool from_x = (y == nullptr);  // see bn_act_bwd_reduce_kernel
  float mu[8], is[8], k0[8], k1[8], k2[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mu[j] = mean[c + j];
    is[j] = invstd[c + j];
    const float gi = (gamma ? gamma[c + j] : 1.f) * is[j];
    sh[j] = (beta ? beta[c + j] : 0.f) - mu[j] * gi;
    k0[j] = gi;
    k1[j] = gi * sum_g[c + j] * inv_count;
    k2[j] = gi * sum_gx[c + j] * inv_count;
  }
  const uint32_t HW = (uint32_t)d.H * (uint32_t)d.W;
  const uint32_t P = (uint32_t)d.N * HW;
  const bool has_res = dres != nullptr && c < d.r_C;
  const bool need_n = sscale != nullptr || (has_res && d.r_stride > 1);
  const uint32_t step = gridDim.x * (uint32_t)ppb;
  for (uint32_t p0 = blockIdx.x * (uint32_t)ppb + pl; p0 < P; p0 += step * U) {
    uint4 xr[U], yr[U], gr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + u * step;
      if (pix < P) {
        xr[u] = ld_nc_v4(x + (long long)pix * d.x_cs + c);
        if (!from_x) yr[u] = ld_nc_v4(y + (long long)pix * d.y_cs + c);
        gr[u] = ld_nc_v4(dy + (long long)pix * d.y_cs + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pix = p0 + u * step;
      if (pix < P) {
        uint32_t n = 0;
        float s = 1.f;
        if (need_n) {
          n = pix / HW;
          if (sscale) s = __ldg(sscale + n);
        }
        F8 o, g;
        const F8 xv = unpack8(xr[u]), gv = unpack8(gr[u]);
        F8 yv;
        if (from_x) {
#pragma unroll
          for (int j = 0; j < 8; ++j) yv.v[j] = fmaf(xv.v[j], k0[j], sh[j]);
        } else {
          yv = unpack8(yr[u]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          g.v[j] = act_bwd(yv.v[j], gv.v[j], d.act);
          const float xh = (xv.v[j] - mu[j]) * is[j];
          o.v[j] = k0[j] * s * g.v[j] - k1[j] - xh * k2[j];
        }
        store_bf16x8(dx + (long long)pix * d.x_cs + c, o);
        if (has_res) {
          __nv_bfloat16* rpz = dres + shortcut_pixel(d, pix, n, HW) * d.r_cs + c;
          if (dres_acc) {
            F8 old = load_bf16x8(rpz);
#pragma unroll
            for (int j = 0; j < 8; ++j) g.v[j] += old.v[j];
          }
          store_bf16x8(rpz, g);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pooling / resampling
// ---------------------------------------------------------------------------------------------
// Max pooling, one block-iteration per OUTPUT ROW (n, ho): no 64-bit index arithmetic in the inner loop, 16-byte
// vectors, the in-window arg-max (first maximum in scan order, like ATen; NaN propagates) as one byte per element.
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int x_cs, int k, int s,
                   int pad, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int Ho, int Wo, int y_cs) {
  const uint32_t V = (uint32_t)C >> 3;
  const uint32_t per_row = (uint32_t)Wo * V;
  const int rows = N * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, ho = row - n * Ho;
    const int h_lo = ho * s - pad;
    const __nv_bfloat16* xn = x + (long long)n * H * W * x_cs;
    for (uint32_t t = threadIdx.x; t < per_row; t += blockDim.x) {
      const uint32_t wo = t / V, cg = t - wo * V;
      const int w_lo = (int)wo * s - pad;
      float best[8];
      uint32_t bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
      for (int r = 0; r < k; ++r) {
        const int h = h_lo + r;
        if (h < 0 || h >= H) continue;
        for (int q = 0; q < k; ++q) {
          const int w = w_lo + q;
          if (w < 0 || w >= W) continue;
          const F8 v = load_bf16x8(xn + ((long long)h * W + w) * x_cs + cg * 8);  // windows overlap: keep in L1
          const uint32_t me = (uint32_t)(r * k + q);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (v.v[j] > best[j] || v.v[j] != v.v[j]) { best[j] = v.v[j]; bi[j] = me; }
        }
      }
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = best[j];
      const long long op = (long long)row * Wo + wo;
      store_bf16x8(y + op * y_cs + cg * 8, o);
      if (idx) {
        uint2 u;
        u.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        u.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        *reinterpret_cast<uint2*>(idx + op * C + cg * 8) = u;
      }
    }
  }
}
// Fixed-geometry variants (3x3/2 pad 1 of the ResNet stem, classification/models.py:49; 2x2/2 of the U-Net encoder,
// unet_models.py:150): the window loops are compile-time, so all K*K (forward) / ceil(K/S)^2 (backward) 16-byte loads
// of a thread are in flight before the first compare — the dynamic-bound kernels issue them one dependent load at a
// time and ran at ~1/4 of the HBM roofline.
template <int K, int S, int PAD>
__global__ void __launch_bounds__(256)
maxpool_fwd_fixed_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int x_cs,
                         __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int Ho, int Wo, int y_cs) {
  const uint32_t V = (uint32_t)C >> 3;
  const uint32_t per_row = (uint32_t)Wo * V;
  const int rows = N * Ho;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, ho = row - n * Ho;
    const int h_lo = ho * S - PAD;
    const __nv_bfloat16* xn = x + (long long)n * H * W * x_cs;
    for (uint32_t t = threadIdx.x; t < per_row; t += blockDim.x) {
      const uint32_t wo = t / V, cg = t - wo * V;
      const int w_lo = (int)wo * S - PAD;
      uint4 raw[K * K];
      bool ok[K * K];
#pragma unroll
      for (int r = 0; r < K; ++r) {
#pragma unroll
        for (int q = 0; q < K; ++q) {
          const int h = h_lo + r, w = w_lo + q;
          ok[r * K + q] = h >= 0 && h < H && w >= 0 && w < W;
          raw[r * K + q] = ok[r * K + q]
                               ? *reinterpret_cast<const uint4*>(xn + ((long long)h * W + w) * x_cs + cg * 8)
                               : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      float best[8];
      uint32_t bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
#pragma unroll
      for (int e = 0; e < K * K; ++e) {
        if (!ok[e]) continue;
        const float2 a = unpack_bf16x2(raw[e].x), b = unpack_bf16x2(raw[e].y), c = unpack_bf16x2(raw[e].z),
                     d = unpack_bf16x2(raw[e].w);
        const float v[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j] || v[j] != v[j]) { best[j] = v[j]; bi[j] = (uint32_t)e; }
      }
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = best[j];
      const long long op = (long long)row * Wo + wo;
      store_bf16x8(y + op * y_cs + cg * 8, o);
      if (idx) {
        uint2 u;
        u.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        u.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        *reinterpret_cast<uint2*>(idx + op * C + cg * 8) = u;
      }
    }
  }
}
template <int K, int S, int PAD>
__global__ void __launch_bounds__(256)
maxpool_bwd_fixed_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy, int N, int H, int W,
                         int C, int Ho, int Wo, int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs, int acc) {
  constexpr int D = (K + S - 1) / S;  // windows that can cover one pixel, per dimension
  const uint32_t V = (uint32_t)C >> 3;
  const uint32_t per_row = (uint32_t)W * V;
  const int rows = N * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, h = row - n * H;
    const int ho0 = (h + PAD) / S;  // last window row covering h; window ho0 - d sees it at filter row (h+PAD) - (ho0-d)*S
    for (uint32_t t = threadIdx.x; t < per_row; t += blockDim.x) {
      const uint32_t wu = t / V, cg = t - wu * V;
      const int w = (int)wu;
      const int wo0 = (w + PAD) / S;
      uint2 iu[D * D];
      uint4 gu[D * D];
      uint32_t me[D * D];
      bool ok[D * D];
#pragma unroll
      for (int a = 0; a < D; ++a) {
#pragma unroll
        for (int b = 0; b < D; ++b) {
          const int ho = ho0 - a, wo = wo0 - b;
          const int r = h + PAD - ho * S, q = w + PAD - wo * S;
          const int e = a * D + b;
          ok[e] = ho >= 0 && ho < Ho && wo >= 0 && wo < Wo && r < K && q < K;
          me[e] = (uint32_t)(r * K + q) * 0x01010101u;
          const long long op = ((long long)n * Ho + ho) * Wo + wo;
          iu[e] = ok[e] ? *reinterpret_cast<const uint2*>(idx + op * C + cg * 8) : make_uint2(0u, 0u);
          gu[e] = ok[e] ? *reinterpret_cast<const uint4*>(dy + op * dy_cs + cg * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      __nv_bfloat16* o = dx + ((long long)row * W + w) * dx_cs + cg * 8;
      float s8[8];
      if (acc) {
        const F8 old = load_bf16x8(o);
#pragma unroll
        for (int j = 0; j < 8; ++j) s8[j] = old.v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s8[j] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < D * D; ++e) {
        if (!ok[e]) continue;
        const uint32_t m0 = __vcmpeq4(iu[e].x, me[e]), m1 = __vcmpeq4(iu[e].y, me[e]);  // 0xFF where this pixel won
        const float2 a = unpack_bf16x2(gu[e].x), b = unpack_bf16x2(gu[e].y), c = unpack_bf16x2(gu[e].z),
                     d = unpack_bf16x2(gu[e].w);
        const float g[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (m0 & (1u << (8 * j))) s8[j] += g[j];
          if (m1 & (1u << (8 * j))) s8[4 + j] += g[4 + j];
        }
      }
      F8 r8;
#pragma unroll
      for (int j = 0; j < 8; ++j) r8.v[j] = s8[j];
      store_bf16x8(o, r8);
    }
  }
}
// ---- 3x3 / stride 2 / pad 1 (the ResNet stem's pool, classification/models.py:49), instruction-lean versions ----
// ncu on the fixed-geometry kernels above at 256 x 112 x 112 x 64 (profiles/r02_maxpool.txt): DRAM traffic = algorithmic,
// DRAM 30 % / 17 % of peak, issue slots 62 % / 67 % busy, 720 / 327 instructions per 8-channel item: INSTRUCTION bound
// (fp32 compare + select per channel and tap; four candidate windows per input pixel, 2.25 of them real).
// Forward: compare / select on packed bf16 pairs (mask from __hgt2_mask | NaN, value and tap index merged with the mask:
// 5 instructions per pair and tap instead of ~16); work items flattened over the whole tensor.
__device__ __forceinline__ uint32_t bf2_gt_or_nan_mask(uint32_t v, uint32_t best) {
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v), b = *reinterpret_cast<const __nv_bfloat162*>(&best);
  return __hgt2_mask(a, b) | __hneu2_mask(a, a);   // v > best (ordered: a NaN best stays), or v is NaN (torch: NaN wins)
}
__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int x_cs,
                      __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, int Ho, int Wo, int y_cs) {
  const uint32_t V = (uint32_t)C >> 3;
  const long long total = (long long)N * Ho * Wo * V;
  constexpr uint32_t kNegInf2 = 0xFF80FF80u;   // bf16 -inf pair: never wins, out-of-image taps
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const uint32_t cg = (uint32_t)(it % V);
    const long long pix = it / V;
    const int wo = (int)(pix % Wo);
    const long long rowi = pix / Wo;
    const int ho = (int)(rowi % Ho), n = (int)(rowi / Ho);
    const int h_lo = ho * 2 - 1, w_lo = wo * 2 - 1;
    const __nv_bfloat16* xn = x + ((long long)n * H * W) * x_cs + cg * 8;
    uint4 raw[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int h = h_lo + r, w = w_lo + q;
        const bool ok = (unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W;
        raw[r * 3 + q] = ok ? *reinterpret_cast<const uint4*>(xn + ((long long)h * W + w) * x_cs)
                            : make_uint4(kNegInf2, kNegInf2, kNegInf2, kNegInf2);
      }
    uint32_t best[4] = {kNegInf2, kNegInf2, kNegInf2, kNegInf2}, bi[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int e = 0; e < 9; ++e) {
      const uint32_t v[4] = {raw[e].x, raw[e].y, raw[e].z, raw[e].w};
      const uint32_t ee = (uint32_t)e * 0x00010001u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m = bf2_gt_or_nan_mask(v[j], best[j]);
        best[j] = (v[j] & m) | (best[j] & ~m);
        bi[j] = (ee & m) | (bi[j] & ~m);
      }
    }
    const long long op = pix;
    *reinterpret_cast<uint4*>(y + op * y_cs + cg * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    if (idx) {
      uint2 u;
      u.x = __byte_perm(bi[0], bi[1], 0x6420);   // tap indices of channels 0..3, one byte each
      u.y = __byte_perm(bi[2], bi[3], 0x6420);
      *reinterpret_cast<uint2*>(idx + op * C + cg * 8) = u;
    }
  }
}
// Backward: one thread owns a 2 x 2 block of INPUT pixels (rows 2i, 2i+1; columns 2j, 2j+1) x 8 channels.  Only the
// windows (i, i+1) x (j, j+1) can have their arg-max there: their 4 index / gradient vectors are loaded once and serve
// the block's 9 real (pixel, window) pairs — instead of 4 candidate windows loaded and tested per pixel (16 per block).
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                      int Ho, int Wo, int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs, int acc) {
  const uint32_t V = (uint32_t)C >> 3;
  const int Hb = (H + 1) >> 1, Wb = (W + 1) >> 1;
  const long long total = (long long)N * Hb * Wb * V;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const uint32_t cg = (uint32_t)(it % V);
    const long long blk = it / V;
    const int j = (int)(blk % Wb);
    const long long rowi = blk / Wb;
    const int i = (int)(rowi % Hb), n = (int)(rowi / Hb);
    uint2 iu[4];
    uint4 gu[4];
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        const int ho = i + di, wo = j + dj;
        const bool ok = ho < Ho && wo < Wo;
        const long long op = ((long long)n * Ho + ho) * Wo + wo;
        iu[di * 2 + dj] = ok ? *reinterpret_cast<const uint2*>(idx + op * C + cg * 8) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        gu[di * 2 + dj] = ok ? *reinterpret_cast<const uint4*>(dy + op * dy_cs + cg * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int h = 2 * i + a, w = 2 * j + b;
        if (h >= H || w >= W) continue;
        __nv_bfloat16* o = dx + (((long long)n * H + h) * W + w) * dx_cs + cg * 8;
        float s8[8];
        if (acc) {
          const F8 old = load_bf16x8(o);
#pragma unroll
          for (int c = 0; c < 8; ++c) s8[c] = old.v[c];
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) s8[c] = 0.f;
        }
        // pixel row 2i + a lies in window row i at filter row a + 1, and (a = 1 only) in window row i + 1 at filter row 0
#pragma unroll
        for (int di = 0; di < 2; ++di)
#pragma unroll
          for (int dj = 0; dj < 2; ++dj) {
            const int r = a + 1 - 2 * di, q = b + 1 - 2 * dj;
            if (r < 0 || q < 0) continue;              // compile-time after unrolling
            const int wdw = di * 2 + dj;
            const uint32_t me = (uint32_t)(r * 3 + q) * 0x01010101u;
            const uint32_t m0 = __vcmpeq4(iu[wdw].x, me), m1 = __vcmpeq4(iu[wdw].y, me);   // 0xFF where this pixel won
            const uint32_t g01 = gu[wdw].x & __byte_perm(m0, 0u, 0x1100), g23 = gu[wdw].y & __byte_perm(m0, 0u, 0x3322);
            const uint32_t g45 = gu[wdw].z & __byte_perm(m1, 0u, 0x1100), g67 = gu[wdw].w & __byte_perm(m1, 0u, 0x3322);
            s8[0] += __uint_as_float(g01 << 16); s8[1] += __uint_as_float(g01 & 0xFFFF0000u);
            s8[2] += __uint_as_float(g23 << 16); s8[3] += __uint_as_float(g23 & 0xFFFF0000u);
            s8[4] += __uint_as_float(g45 << 16); s8[5] += __uint_as_float(g45 & 0xFFFF0000u);
            s8[6] += __uint_as_float(g67 << 16); s8[7] += __uint_as_float(g67 & 0xFFFF0000u);
          }
        F8 r8;
#pragma unroll
        for (int c = 0; c < 8; ++c) r8.v[c] = s8[c];
        store_bf16x8(o, r8);
      }
  }
}
// gather formulation (no atomics): every INPUT pixel sums dy of the windows whose arg-max it is; one block-iteration
// per input row (n, h)
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                   int k, int s, int pad, int Ho, int Wo, int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs,
                   int acc) {
  const uint32_t V = (uint32_t)C >> 3;
  const uint32_t per_row = (uint32_t)W * V;
  const int rows = N * H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, h = row - n * H;
    // windows ho with ho*s - pad <= h < ho*s - pad + k
    const int ho_lo = (h + pad - k + 1) <= 0 ? 0 : (h + pad - k + s) / s;
    int ho_hi = (h + pad) / s;
    if (ho_hi > Ho - 1) ho_hi = Ho - 1;
    for (uint32_t t = threadIdx.x; t < per_row; t += blockDim.x) {
      const uint32_t wu = t / V, cg = t - wu * V;
      const int w = (int)wu;
      const int wo_lo = (w + pad - k + 1) <= 0 ? 0 : (w + pad - k + s) / s;
      int wo_hi = (w + pad) / s;
      if (wo_hi > Wo - 1) wo_hi = Wo - 1;
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = 0.f;
      for (int ho = ho_lo; ho <= ho_hi; ++ho) {
        const int r = h + pad - ho * s;
        for (int wo = wo_lo; wo <= wo_hi; ++wo) {
          const uint32_t me = (uint32_t)(r * k + (w + pad - wo * s));
          const long long op = ((long long)n * Ho + ho) * Wo + wo;
          const uint2 u = *reinterpret_cast<const uint2*>(idx + op * C + cg * 8);
          const F8 g = load_bf16x8(dy + op * dy_cs + cg * 8);                  // shared by neighbours: keep in L1
          const uint32_t me4 = me * 0x01010101u;
          const uint32_t m0 = __vcmpeq4(u.x, me4), m1 = __vcmpeq4(u.y, me4);  // 0xFF where this pixel won
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (m0 & (1u << (8 * j))) a[j] += g.v[j];
            if (m1 & (1u << (8 * j))) a[4 + j] += g.v[4 + j];
          }
        }
      }
      __nv_bfloat16* o = dx + ((long long)row * W + w) * dx_cs + cg * 8;
      F8 r8;
      if (acc) {
        const F8 old = load_bf16x8(o);
#pragma unroll
        for (int j = 0; j < 8; ++j) r8.v[j] = old.v[j] + a[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) r8.v[j] = a[j];
      }
      store_bf16x8(o, r8);
    }
  }
}

__global__ void upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                      int x_cs, __nv_bfloat16* __restrict__ y, int y_cs) {
  const int V = C >> 3;
  const int H2 = 2 * H, W2 = 2 * W;
  const long long total = (long long)N * H2 * W2 * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const int w = (int)(p % W2); p /= W2;
    const int h = (int)(p % H2);
    const int n = (int)(p / H2);
    const uint4 v = *reinterpret_cast<const uint4*>(
        x + (((long long)n * H + (h >> 1)) * W + (w >> 1)) * x_cs + cg * 8);
    *reinterpret_cast<uint4*>(y + (((long long)n * H2 + h) * W2 + w) * y_cs + cg * 8) = v;
  }
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (north star: "bilinear/nearest upsampling"; the
// reference's blocks use nearest, blocks.py:532).  torch's source index: max(0, (dst + 0.5) / 2 - 0.5), i.e. per axis
//   dst = 2i   -> 0.25 * x[i-1] + 0.75 * x[i]   (i = 0: x[0]),   dst = 2i+1 -> 0.75 * x[i] + 0.25 * x[min(i+1, H-1)]
// evaluated like ATen: h0 * (w0 * x00 + w1 * x01) + h1 * (w0 * x10 + w1 * x11) in fp32 on the bf16 inputs.
__device__ __forceinline__ void bilin_src(int dst, int size, int* i0, int* i1, float* l0, float* l1) {
  float real = (dst + 0.5f) * 0.5f - 0.5f;
  real = real < 0.f ? 0.f : real;
  const int a = (int)real;
  *i0 = a;
  *i1 = a + 1 < size ? a + 1 : size - 1;
  *l1 = real - (float)a;
  *l0 = 1.f - *l1;
}
__global__ void upsample_bilinear2x_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C,
                                               int x_cs, __nv_bfloat16* __restrict__ y, int y_cs) {
  const int V = C >> 3;
  const int H2 = 2 * H, W2 = 2 * W;
  const long long total = (long long)N * H2 * W2 * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const int w = (int)(p % W2); p /= W2;
    const int h = (int)(p % H2);
    const int n = (int)(p / H2);
    int h0, h1, w0, w1;
    float lh0, lh1, lw0, lw1;
    bilin_src(h, H, &h0, &h1, &lh0, &lh1);
    bilin_src(w, W, &w0, &w1, &lw0, &lw1);
    const __nv_bfloat16* b = x + (long long)n * H * W * x_cs + cg * 8;
    const F8 a00 = load_bf16x8(b + ((long long)h0 * W + w0) * x_cs), a01 = load_bf16x8(b + ((long long)h0 * W + w1) * x_cs),
             a10 = load_bf16x8(b + ((long long)h1 * W + w0) * x_cs), a11 = load_bf16x8(b + ((long long)h1 * W + w1) * x_cs);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o.v[j] = lh0 * (lw0 * a00.v[j] + lw1 * a01.v[j]) + lh1 * (lw0 * a10.v[j] + lw1 * a11.v[j]);
    store_bf16x8(y + (((long long)n * H2 + h) * W2 + w) * y_cs + cg * 8, o);
  }
}
// gather form of the transpose: source pixel (i, j) collects, per axis, from outputs 2i-1 (0.25), 2i (0.75; 1 at i = 0),
// 2i+1 (0.75; 1 at i = H-1) and 2i+2 (0.25) - no atomics, one thread per 8 channels of a source pixel
__device__ __forceinline__ int bilin_taps(int i, int size, int* dst, float* wgt) {
  int n = 0;
  if (i >= 1) { dst[n] = 2 * i - 1; wgt[n++] = 0.25f; }
  dst[n] = 2 * i; wgt[n++] = i == 0 ? 1.0f : 0.75f;
  dst[n] = 2 * i + 1; wgt[n++] = i == size - 1 ? 1.0f : 0.75f;
  if (i + 1 < size) { dst[n] = 2 * i + 2; wgt[n++] = 0.25f; }
  return n;
}
__global__ void upsample_bilinear2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                                               int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs) {
  const int V = C >> 3;
  const int H2 = 2 * H, W2 = 2 * W;
  const long long total = (long long)N * H * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    int hd[4], wd[4];
    float hw_[4], ww[4];
    const int nh = bilin_taps(h, H, hd, hw_), nw = bilin_taps(w, W, wd, ww);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
    const __nv_bfloat16* b = dy + (long long)n * H2 * W2 * dy_cs + cg * 8;
    for (int a = 0; a < nh; ++a)
      for (int c = 0; c < nw; ++c) {
        const F8 g = load_bf16x8(b + ((long long)hd[a] * W2 + wd[c]) * dy_cs);
        const float f = hw_[a] * ww[c];
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = fmaf(f, g.v[j], o.v[j]);
      }
    store_bf16x8(dx + (((long long)n * H + h) * W + w) * dx_cs + cg * 8, o);
  }
}
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                                      int dy_cs, __nv_bfloat16* __restrict__ dx, int dx_cs) {
  const int V = C >> 3;
  const int W2 = 2 * W;
  const long long total = (long long)N * H * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    const __nv_bfloat16* b = dy + (((long long)n * 2 * H + 2 * h) * W2 + 2 * w) * dy_cs + cg * 8;
    F8 a = load_bf16x8(b), b1 = load_bf16x8(b + dy_cs), c0 = load_bf16x8(b + (long long)W2 * dy_cs),
       c1 = load_bf16x8(b + (long long)W2 * dy_cs + dy_cs);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = (a.v[j] + b1.v[j]) + (c0.v[j] + c1.v[j]);
    store_bf16x8(dx + (((long long)n * H + h) * W + w) * dx_cs + cg * 8, o);
  }
}

__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int HW, int C, int x_cs,
                                   __nv_bfloat16* __restrict__ y) {
  const int V = C >> 3;
  const long long total = (long long)N * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    const int n = (int)(i / V);
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      F8 v = load_bf16x8(x + ((long long)n * HW + p) * x_cs + cg * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += v.v[j];
    }
    F8 o;
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = a[j] * inv;
    store_bf16x8(y + (long long)n * C + cg * 8, o);
  }
}
__global__ void avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int N, int HW, int C,
                                   __nv_bfloat16* __restrict__ dx, int dx_cs) {
  const int V = C >> 3;
  const long long total = (long long)N * HW * V;
  const float inv = 1.f / (float)HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    const long long p = i / V;
    const int n = (int)(p / HW);
    F8 g = load_bf16x8(dy + (long long)n * C + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] *= inv;
    store_bf16x8(dx + p * dx_cs + cg * 8, g);
  }
}

// generic 2-input elementwise on channel slices. op: 0 copy(a), 1 relu(a+b), 2 a+b,
// 3 relu_bwd (a = y, b = dy)
template <int OP>
__global__ void ew2_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           long long P, int C, int a_cs, int b_cs, __nv_bfloat16* __restrict__ y,
                           int y_cs) {
  const int V = C >> 3;
  const long long total = P * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    const long long p = i / V;
    if (OP == 0) {
      *reinterpret_cast<uint4*>(y + p * y_cs + cg * 8) =
          *reinterpret_cast<const uint4*>(a + p * a_cs + cg * 8);
    } else {
      F8 av = load_bf16x8(a + p * a_cs + cg * 8), bv = load_bf16x8(b + p * b_cs + cg * 8), o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (OP == 1) o.v[j] = fmaxf(av.v[j] + bv.v[j], 0.f);
        else if (OP == 2) o.v[j] = av.v[j] + bv.v[j];
        else o.v[j] = av.v[j] > 0.f ? bv.v[j] : 0.f;
      }
      store_bf16x8(y + p * y_cs + cg * 8, o);
    }
  }
}

__global__ void gate_mul_fwd_kernel(const __nv_bfloat16* __restrict__ skip,
                                    const __nv_bfloat16* __restrict__ pg, int N, int H, int W, int C,
                                    int skip_cs, int p_cs, __nv_bfloat16* __restrict__ y, int y_cs) {
  const int V = C >> 3;
  const int Hp = H >> 1, Wp = W >> 1;
  const long long total = (long long)N * H * W * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const long long pix = p;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int n = (int)(p / H);
    F8 s = load_bf16x8(skip + pix * skip_cs + cg * 8);
    F8 g = load_bf16x8(pg + (((long long)n * Hp + (h >> 1)) * Wp + (w >> 1)) * p_cs + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) s.v[j] *= g.v[j];
    store_bf16x8(y + pix * y_cs + cg * 8, s);
  }
}
// one thread per low-res pixel x 8 channels: dp = sum_{2x2} dy*skip ; dskip = dy * p
__global__ void gate_mul_bwd_kernel(const __nv_bfloat16* __restrict__ skip,
                                    const __nv_bfloat16* __restrict__ pg,
                                    const __nv_bfloat16* __restrict__ dy, int N, int H, int W, int C,
                                    int skip_cs, int p_cs, int dy_cs, __nv_bfloat16* __restrict__ dskip,
                                    int dskip_cs, int dskip_acc, __nv_bfloat16* __restrict__ dp,
                                    int dp_cs) {
  const int V = C >> 3;
  const int Hp = H >> 1, Wp = W >> 1;
  const long long total = (long long)N * Hp * Wp * V;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % V);
    long long p = i / V;
    const long long ppix = p;
    const int wp = (int)(p % Wp); p /= Wp;
    const int hp = (int)(p % Hp);
    const int n = (int)(p / Hp);
    F8 g = load_bf16x8(pg + ppix * p_cs + cg * 8);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const long long pix = ((long long)n * H + 2 * hp + dh) * W + 2 * wp + dw;
        F8 d = load_bf16x8(dy + pix * dy_cs + cg * 8);
        F8 s = load_bf16x8(skip + pix * skip_cs + cg * 8);
        F8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = fmaf(d.v[j], s.v[j], acc[j]);
          o.v[j] = d.v[j] * g.v[j];
        }
        __nv_bfloat16* dst = dskip + pix * dskip_cs + cg * 8;
        if (dskip_acc) {
          F8 old = load_bf16x8(dst);
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] += old.v[j];
        }
        store_bf16x8(dst, o);
      }
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
    store_bf16x8(dp + ppix * dp_cs + cg * 8, o);
  }
}

// per-channel sum over pixels of an NHWC bf16 tensor -> fp32 (bias gradients)
__global__ void channel_sum_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, int cs,
                                   float* out, float* rows_ws) {
  extern __shared__ float red[];
  const int V = C >> 3;
  const int cg = threadIdx.x % V, pl = threadIdx.x / V, ppb = blockDim.x / V;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  for (long long pix = (long long)blockIdx.x * ppb + pl; pix < P; pix += (long long)gridDim.x * ppb) {
    F8 v = load_bf16x8(x + pix * cs + cg * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += v.v[j];
  }
  float* mine = red + ((long long)pl * V + cg) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = a[j];
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < ppb; ++l) {
      const float* o = red + ((long long)l * V + cg) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += o[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (rows_ws) rows_ws[(long long)blockIdx.x * C + cg * 8 + j] = a[j];   // deterministic mode: fixed-order second stage
      else atomicAdd(out + cg * 8 + j, a[j]);
    }
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)
#define REQ_C8(C, cs, what)                                                                  \
  MSP_REQUIRE((C) > 0 && (C) % 8 == 0 && (cs) % 8 == 0 && (cs) >= (C),                        \
              what ": channels (%d) and pixel stride (%d) must be multiples of 8", (int)(C), \
              (int)(cs))

extern "C" int msp_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int H, int W, int Cpad,
                                         void* y, void* stream) {
  MSP_REQUIRE(x && y && N > 0 && C > 0 && Cpad >= C, "nchw_to_nhwc: bad arguments");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), (unsigned)N);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, ST>>>(x, C, HW, Cpad, (__nv_bfloat16*)y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_nchw_f32_to_rowwin_bf16(const float* x, int N, int C, int H, int W, int cpp, int pad_l,
                                           int Wp, void* y, void* stream) {
  MSP_REQUIRE(x && y && N > 0 && C > 0 && (cpp == 8 || cpp == 16) && C <= cpp && pad_l >= 0 &&
                  Wp >= W + pad_l,
              "nchw_to_rowwin: bad arguments");
  const long long total = (long long)N * H * Wp;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  nchw_to_rowwin_kernel<<<blocks, 256, 0, ST>>>(x, C, H, W, cpp, pad_l, Wp, total, (__nv_bfloat16*)y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_nchw_bf16_to_rowwin_bf16(const void* x, int N, int C, int H, int W, int cpp, int pad_l,
                                           int Wp, void* y, void* stream) {
  MSP_REQUIRE(x && y && N > 0 && C > 0 && (cpp == 8 || cpp == 16) && C <= cpp && pad_l >= 0 &&
                  Wp >= W + pad_l,
              "nchw_to_rowwin: bad arguments");
  const long long total = (long long)N * H * Wp;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  nchw_to_rowwin_kernel<<<blocks, 256, 0, ST>>>((const __nv_bfloat16*)x, C, H, W, cpp, pad_l, Wp, total, (__nv_bfloat16*)y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_nhwc_bf16_to_nchw_f32(const void* x, int N, int C, int H, int W, int x_cs,
                                         float* y, void* stream) {
  MSP_REQUIRE(x && y && N > 0 && C > 0 && x_cs >= C, "nhwc_to_nchw: bad arguments");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, ST>>>((const __nv_bfloat16*)x, C, HW, x_cs, y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_nchw_f32_grad_to_nhwc_bf16(const float* g, int N, int C, int H, int W,
                                              int g_cs_out, void* y, void* stream) {
  return msp_nchw_f32_to_nhwc_bf16(g, N, C, H, W, g_cs_out, y, stream);
}

extern "C" int msp_bn_finalize(float* ch_sum, float* ch_sqsum, int C, double count, float eps,
                               float momentum, float* mean, float* invstd, float* running_mean,
                               float* running_var, int reset_sums, int rows, void* stream) {
  MSP_REQUIRE(ch_sum && ch_sqsum && mean && invstd && C > 0 && count > 0, "bn_finalize: bad arguments");
  MSP_REQUIRE((running_mean == nullptr) == (running_var == nullptr),
              "bn_finalize: need both running buffers or none");
  MSP_REQUIRE(rows >= 1 && (rows == 1 || ch_sqsum == ch_sum + C),
              "bn_finalize: the per-CTA workspace is [rows][2][C] (ch_sqsum = ch_sum + C)");
  MSP_CHECK_CUDA(msp_launch_pdl(bn_finalize_kernel, dim3((C + 31) / 32), dim3(32 * kRowGroups), 0, ST, ch_sum, ch_sqsum, C,
                                count, eps, momentum, mean, invstd, running_mean, running_var, reset_sums, rows));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_reduce_rows(float* ws, int rows, int n, float* out, int reset, void* stream) {
  MSP_REQUIRE(ws && out && rows >= 1 && n >= 1, "reduce_rows: bad arguments");
  reduce_rows_kernel<<<(n + 31) / 32, 32 * kRowGroups, 0, ST>>>(ws, rows, n, out, reset);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_bn_eval_prepare(const float* running_var, int C, float eps, float* invstd,
                                   void* stream) {
  MSP_REQUIRE(running_var && invstd && C > 0, "bn_eval_prepare: bad arguments");
  bn_eval_prepare_kernel<<<(C + 127) / 128, 128, 0, ST>>>(running_var, C, eps, invstd);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

static int check_bn_desc(const msp_bn_act_desc* d) {
  MSP_REQUIRE(d && d->N > 0 && d->H > 0 && d->W > 0, "bn_act: empty tensor");
  REQ_C8(d->C, d->x_cs, "bn_act(x)");
  REQ_C8(d->C, d->y_cs, "bn_act(y)");
  MSP_REQUIRE(d->C <= 2048, "bn_act: C=%d > 2048 unsupported", d->C);
  MSP_REQUIRE(d->act >= 0 && d->act <= 2, "bn_act: bad activation");
  return MSP_OK;
}

extern "C" int msp_bn_act_fwd(const msp_bn_act_desc* d, const void* x, const float* mean,
                              const float* invstd, const float* gamma, const float* beta,
                              const float* sample_scale, const void* residual, void* y, void* stream) {
  int rc = check_bn_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && y && mean && invstd, "bn_act_fwd: null pointer");
  if (residual) {
    MSP_REQUIRE(d->r_C > 0 && d->r_C % 8 == 0 && d->r_cs % 8 == 0 && d->r_stride >= 1,
                "bn_act_fwd: bad residual description");
  }
  const int V = d->C / 8, T = threads_for_vecs(V), ppb = T / V;
  const long long P = (long long)d->N * d->H * d->W;
  MSP_REQUIRE(P < (1ll << 31), "bn_act: too many pixels");
  MSP_CHECK_CUDA(msp_launch_pdl(bn_act_fwd_kernel<4>, dim3(resident_grid(bn_act_fwd_kernel<4>, T, 0, P, ppb * 4)), dim3(T), 0,
                                ST, *d, (const __nv_bfloat16*)x, mean, invstd, gamma, beta, sample_scale,
                                (const __nv_bfloat16*)residual, (__nv_bfloat16*)y));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

static int check_mask_from_x(const msp_bn_act_desc* d, const void* y, const float* sample_scale, const void* dres) {
  if (y != nullptr) return MSP_OK;
  MSP_REQUIRE(d->act == MSP_ACT_RELU && sample_scale == nullptr && dres == nullptr,
              "bn_act_bwd: y may be omitted only for BatchNorm -> ReLU without shortcut / sample scale");
  return MSP_OK;
}

static int bn_bwd_reduce_impl(const msp_bn_act_desc* d, const void* x, const void* y,
                              const void* dy, const float* mean, const float* invstd,
                              const float* gamma, const float* beta, const float* sample_scale,
                              float* sum_g, float* sum_gx, float* rows_ws, int ws_rows, int* rows_used, void* stream) {
  int rc = check_bn_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && dy && mean && invstd && sum_g && sum_gx, "bn_act_bwd_reduce: null pointer");
  rc = check_mask_from_x(d, y, sample_scale, nullptr);
  if (rc) return rc;
  const int V = d->C / 8, T = threads_for_vecs(V), ppb = T / V;
  const long long P = (long long)d->N * d->H * d->W;
  if (rows_ws != nullptr) {
    MSP_REQUIRE(ws_rows >= 1 && sum_gx == sum_g + d->C, "bn_act_bwd_reduce: deterministic mode needs sum_gx = sum_g + C");
  } else if (sum_gx == sum_g + d->C) {  // the usual [2][C] buffer: one memset node
    MSP_CHECK_CUDA(cudaMemsetAsync(sum_g, 0, sizeof(float) * 2 * d->C, ST));
  } else {
    MSP_CHECK_CUDA(cudaMemsetAsync(sum_g, 0, sizeof(float) * d->C, ST));
    MSP_CHECK_CUDA(cudaMemsetAsync(sum_gx, 0, sizeof(float) * d->C, ST));
  }
  const size_t smem = (size_t)T * 16 * sizeof(float);
  MSP_REQUIRE(P < (1ll << 31), "bn_act: too many pixels");
  static int ured = -1;
  if (ured < 0) { const char* e = getenv("MSP_BN_RED_U"); ured = e ? atoi(e) : 4; }  // 4 loads per tensor in flight: 3.61 -> 3.11 ms over the ResNet-50 layers
  int grid = ured == 4 ? resident_grid(bn_act_bwd_reduce_kernel<4>, T, smem, P, ppb * 4)
                       : resident_grid(bn_act_bwd_reduce_kernel<2>, T, smem, P, ppb * 2);
  if (rows_ws != nullptr && grid > ws_rows) grid = ws_rows;
  if (ured == 4)
    bn_act_bwd_reduce_kernel<4><<<grid, T, smem, ST>>>(
        *d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, mean, invstd,
        sample_scale, sum_g, sum_gx, gamma, beta, rows_ws);
  else
    bn_act_bwd_reduce_kernel<2><<<grid, T, smem, ST>>>(
        *d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, mean, invstd,
        sample_scale, sum_g, sum_gx, gamma, beta, rows_ws);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  if (rows_used != nullptr) {  // the caller adds the rows itself (msp_p2p_stats_exchange)
    *rows_used = grid;
  } else if (rows_ws != nullptr) {  // second stage: the blocks' rows in fixed order
    reduce_rows_kernel<<<(2 * d->C + 31) / 32, 32 * kRowGroups, 0, ST>>>(rows_ws, grid, 2 * d->C, sum_g, 0);
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
  }
  return MSP_OK;
}

extern "C" int msp_bn_act_bwd_reduce(const msp_bn_act_desc* d, const void* x, const void* y,
                                     const void* dy, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, const float* sample_scale,
                                     float* sum_g, float* sum_gx, float* rows_ws, int ws_rows, void* stream) {
  return bn_bwd_reduce_impl(d, x, y, dy, mean, invstd, gamma, beta, sample_scale, sum_g, sum_gx, rows_ws, ws_rows, nullptr,
                            stream);
}

extern "C" int msp_bn_act_bwd_reduce_rows(const msp_bn_act_desc* d, const void* x, const void* y,
                                          const void* dy, const float* mean, const float* invstd,
                                          const float* gamma, const float* beta, const float* sample_scale,
                                          float* rows_ws, int ws_rows, int* rows_used, void* stream) {
  MSP_REQUIRE(rows_ws != nullptr && ws_rows >= 1 && rows_used != nullptr, "bn_act_bwd_reduce_rows: needs the row workspace");
  // first stage only: block b's sums land in row b of the [ws_rows][2][C] workspace; *rows_used rows are valid
  return bn_bwd_reduce_impl(d, x, y, dy, mean, invstd, gamma, beta, sample_scale, rows_ws, rows_ws + d->C, rows_ws, ws_rows,
                            rows_used, stream);
}

extern "C" int msp_bn_act_bwd_apply(const msp_bn_act_desc* d, const void* x, const void* y,
                                    const void* dy, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const float* sample_scale,
                                    const float* sum_g, const float* sum_gx, double count, void* dx,
                                    void* dres, int dres_accumulate, void* stream) {
  int rc = check_bn_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && dy && mean && invstd && sum_g && sum_gx && dx && count > 0,
              "bn_act_bwd_apply: null pointer");
  rc = check_mask_from_x(d, y, sample_scale, dres);
  if (rc) return rc;
  const int V = d->C / 8, T = threads_for_vecs(V), ppb = T / V;
  const long long P = (long long)d->N * d->H * d->W;
  MSP_REQUIRE(P < (1ll << 31), "bn_act: too many pixels");
  static int uapp = -1;
  if (uapp < 0) { const char* e = getenv("MSP_BN_APP_U"); uapp = e ? atoi(e) : 4; }  // 4.27 -> 4.09 ms over the ResNet-50 layers
  if (uapp == 4)
    MSP_CHECK_CUDA(msp_launch_pdl(bn_act_bwd_apply_kernel<4>, dim3(resident_grid(bn_act_bwd_apply_kernel<4>, T, 0, P, ppb * 4)),
                                  dim3(T), 0, ST, *d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y,
                                  (const __nv_bfloat16*)dy, mean, invstd, gamma, sample_scale, sum_g, sum_gx,
                                  (float)(1.0 / count), (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, dres_accumulate, beta));
  else
    MSP_CHECK_CUDA(msp_launch_pdl(bn_act_bwd_apply_kernel<2>, dim3(resident_grid(bn_act_bwd_apply_kernel<2>, T, 0, P, ppb * 2)),
                                  dim3(T), 0, ST, *d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y,
                                  (const __nv_bfloat16*)dy, mean, invstd, gamma, sample_scale, sum_g, sum_gx,
                                  (float)(1.0 / count), (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, dres_accumulate, beta));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_maxpool_fwd(const void* x, int N, int H, int W, int C, int x_cs, int k, int stride,
                               int pad, void* y, void* idx, int Ho, int Wo, int y_cs, void* stream) {
  REQ_C8(C, x_cs, "maxpool_fwd(x)");
  REQ_C8(C, y_cs, "maxpool_fwd(y)");
  MSP_REQUIRE(x && y && k >= 1 && k <= 15 && stride >= 1 && pad >= 0, "maxpool_fwd: bad arguments");
  if (k == 3 && stride == 2 && pad == 1)
    maxpool3s2_fwd_kernel<<<resident_grid(maxpool3s2_fwd_kernel, 256, 0, (long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST>>>(
        (const __nv_bfloat16*)x, N, H, W, C, x_cs, (__nv_bfloat16*)y, (uint8_t*)idx, Ho, Wo, y_cs);
  else if (k == 2 && stride == 2 && pad == 0)
    maxpool_fwd_fixed_kernel<2, 2, 0><<<resident_grid(maxpool_fwd_fixed_kernel<2, 2, 0>, 256, 0, (long long)N * Ho, 1), 256, 0, ST>>>(
        (const __nv_bfloat16*)x, N, H, W, C, x_cs, (__nv_bfloat16*)y, (uint8_t*)idx, Ho, Wo, y_cs);
  else
    maxpool_fwd_kernel<<<resident_grid(maxpool_fwd_kernel, 256, 0, (long long)N * Ho, 1), 256, 0, ST>>>(
        (const __nv_bfloat16*)x, N, H, W, C, x_cs, k, stride, pad, (__nv_bfloat16*)y, (uint8_t*)idx, Ho, Wo, y_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_maxpool_bwd(const void* idx, const void* dy, int N, int H, int W, int C, int k,
                               int stride, int pad, int Ho, int Wo, int dy_cs, void* dx, int dx_cs,
                               int accumulate, void* stream) {
  REQ_C8(C, dy_cs, "maxpool_bwd(dy)");
  REQ_C8(C, dx_cs, "maxpool_bwd(dx)");
  MSP_REQUIRE(idx && dy && dx, "maxpool_bwd: null pointer");
  if (k == 3 && stride == 2 && pad == 1 && Ho == (H - 1) / 2 + 1 && Wo == (W - 1) / 2 + 1)
    maxpool3s2_bwd_kernel<<<resident_grid(maxpool3s2_bwd_kernel, 256, 0,
                                          (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8), 256), 256, 0, ST>>>(
        (const uint8_t*)idx, (const __nv_bfloat16*)dy, N, H, W, C, Ho, Wo, dy_cs, (__nv_bfloat16*)dx, dx_cs, accumulate);
  else if (k == 3 && stride == 2 && pad == 1)
    maxpool_bwd_fixed_kernel<3, 2, 1><<<resident_grid(maxpool_bwd_fixed_kernel<3, 2, 1>, 256, 0, (long long)N * H, 1), 256, 0, ST>>>(
        (const uint8_t*)idx, (const __nv_bfloat16*)dy, N, H, W, C, Ho, Wo, dy_cs, (__nv_bfloat16*)dx, dx_cs, accumulate);
  else if (k == 2 && stride == 2 && pad == 0)
    maxpool_bwd_fixed_kernel<2, 2, 0><<<resident_grid(maxpool_bwd_fixed_kernel<2, 2, 0>, 256, 0, (long long)N * H, 1), 256, 0, ST>>>(
        (const uint8_t*)idx, (const __nv_bfloat16*)dy, N, H, W, C, Ho, Wo, dy_cs, (__nv_bfloat16*)dx, dx_cs, accumulate);
  else
    maxpool_bwd_kernel<<<resident_grid(maxpool_bwd_kernel, 256, 0, (long long)N * H, 1), 256, 0, ST>>>(
        (const uint8_t*)idx, (const __nv_bfloat16*)dy, N, H, W, C, k, stride, pad, Ho, Wo, dy_cs, (__nv_bfloat16*)dx,
        dx_cs, accumulate);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_upsample2x_fwd(const void* x, int N, int H, int W, int C, int x_cs, void* y,
                                  int y_cs, void* stream) {
  REQ_C8(C, x_cs, "upsample2x_fwd(x)");
  REQ_C8(C, y_cs, "upsample2x_fwd(y)");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample2x_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)x, N, H, W, C, x_cs,
                                                              (__nv_bfloat16*)y, y_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_upsample2x_bwd(const void* dy, int N, int H, int W, int C, int dy_cs, void* dx,
                                  int dx_cs, void* stream) {
  REQ_C8(C, dy_cs, "upsample2x_bwd(dy)");
  REQ_C8(C, dx_cs, "upsample2x_bwd(dx)");
  const long long total = (long long)N * H * W * (C / 8);
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)dy, N, H, W, C,
                                                              dy_cs, (__nv_bfloat16*)dx, dx_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_upsample_bilinear2x_fwd(const void* x, int N, int H, int W, int C, int x_cs, void* y, int y_cs,
                                           void* stream) {
  REQ_C8(C, x_cs, "upsample_bilinear2x_fwd(x)");
  REQ_C8(C, y_cs, "upsample_bilinear2x_fwd(y)");
  const long long total = (long long)N * 4 * H * W * (C / 8);
  upsample_bilinear2x_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)x, N, H, W, C, x_cs,
                                                                       (__nv_bfloat16*)y, y_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_upsample_bilinear2x_bwd(const void* dy, int N, int H, int W, int C, int dy_cs, void* dx,
                                           int dx_cs, void* stream) {
  REQ_C8(C, dy_cs, "upsample_bilinear2x_bwd(dy)");
  REQ_C8(C, dx_cs, "upsample_bilinear2x_bwd(dx)");
  const long long total = (long long)N * H * W * (C / 8);
  upsample_bilinear2x_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)dy, N, H, W, C,
                                                                       dy_cs, (__nv_bfloat16*)dx, dx_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_avgpool_fwd(const void* x, int N, int HW, int C, int x_cs, void* y, void* stream) {
  REQ_C8(C, x_cs, "avgpool_fwd");
  const long long total = (long long)N * (C / 8);
  avgpool_fwd_kernel<<<grid_for(total, 128), 128, 0, ST>>>((const __nv_bfloat16*)x, N, HW, C, x_cs,
                                                           (__nv_bfloat16*)y);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_avgpool_bwd(const void* dy, int N, int HW, int C, void* dx, int dx_cs, void* stream) {
  REQ_C8(C, dx_cs, "avgpool_bwd");
  const long long total = (long long)N * HW * (C / 8);
  avgpool_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)dy, N, HW, C,
                                                           (__nv_bfloat16*)dx, dx_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

template <int OP>
static int launch_ew2(const void* a, const void* b, long long P, int C, int a_cs, int b_cs, void* y,
                      int y_cs, void* stream) {
  REQ_C8(C, a_cs, "elementwise(a)");
  REQ_C8(C, y_cs, "elementwise(y)");
  if (OP != 0) REQ_C8(C, b_cs, "elementwise(b)");
  MSP_REQUIRE(a && y && (OP == 0 || b) && P > 0, "elementwise: bad arguments");
  ew2_kernel<OP><<<grid_for(P * (C / 8), 256), 256, 0, ST>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, P, C, a_cs, b_cs, (__nv_bfloat16*)y, y_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_copy_channels(const void* x, long long P, int C, int x_cs, void* y, int y_cs,
                                 void* stream) {
  return launch_ew2<0>(x, nullptr, P, C, x_cs, x_cs, y, y_cs, stream);
}
extern "C" int msp_add_relu_fwd(const void* a, const void* b, long long P, int C, int a_cs, int b_cs,
                                void* y, int y_cs, void* stream) {
  return launch_ew2<1>(a, b, P, C, a_cs, b_cs, y, y_cs, stream);
}
extern "C" int msp_add(const void* a, const void* b, long long P, int C, int a_cs, int b_cs, void* y,
                       int y_cs, void* stream) {
  return launch_ew2<2>(a, b, P, C, a_cs, b_cs, y, y_cs, stream);
}
extern "C" int msp_relu_bwd(const void* y, const void* dy, long long P, int C, int y_cs, int dy_cs,
                            void* dx, int dx_cs, void* stream) {
  return launch_ew2<3>(y, dy, P, C, y_cs, dy_cs, dx, dx_cs, stream);
}

extern "C" int msp_gate_mul_fwd(const void* skip, const void* p, int N, int H, int W, int C,
                                int skip_cs, int p_cs, void* y, int y_cs, void* stream) {
  REQ_C8(C, skip_cs, "gate_mul_fwd(skip)");
  REQ_C8(C, p_cs, "gate_mul_fwd(p)");
  REQ_C8(C, y_cs, "gate_mul_fwd(y)");
  MSP_REQUIRE(H % 2 == 0 && W % 2 == 0, "gate_mul_fwd: H, W must be even");
  const long long total = (long long)N * H * W * (C / 8);
  gate_mul_fwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>((const __nv_bfloat16*)skip,
                                                            (const __nv_bfloat16*)p, N, H, W, C,
                                                            skip_cs, p_cs, (__nv_bfloat16*)y, y_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_gate_mul_bwd(const void* skip, const void* p, const void* dy, int N, int H, int W,
                                int C, int skip_cs, int p_cs, int dy_cs, void* dskip, int dskip_cs,
                                int dskip_accumulate, void* dp, int dp_cs, void* stream) {
  REQ_C8(C, skip_cs, "gate_mul_bwd(skip)");
  REQ_C8(C, dy_cs, "gate_mul_bwd(dy)");
  MSP_REQUIRE(H % 2 == 0 && W % 2 == 0 && dskip && dp, "gate_mul_bwd: bad arguments");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  gate_mul_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(
      (const __nv_bfloat16*)skip, (const __nv_bfloat16*)p, (const __nv_bfloat16*)dy, N, H, W, C,
      skip_cs, p_cs, dy_cs, (__nv_bfloat16*)dskip, dskip_cs, dskip_accumulate, (__nv_bfloat16*)dp,
      dp_cs);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
extern "C" int msp_channel_sum(const void* x, long long P, int C, int cs, float* out, float* rows_ws, int ws_rows,
                               void* stream) {
  REQ_C8(C, cs, "channel_sum");
  MSP_REQUIRE(C <= 2048 && out && x, "channel_sum: bad arguments");
  const int V = C / 8, T = threads_for_vecs(V), ppb = T / V;
  int grid = grid_for(P, ppb * 8, 4);
  if (rows_ws != nullptr) {
    MSP_REQUIRE(ws_rows >= 1, "channel_sum: empty row workspace");
    if (grid > ws_rows) grid = ws_rows;
  } else {
    MSP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, ST));
  }
  channel_sum_kernel<<<grid, T, (size_t)T * 8 * sizeof(float), ST>>>((const __nv_bfloat16*)x, P, C, cs, out, rows_ws);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  if (rows_ws != nullptr) {
    reduce_rows_kernel<<<(C + 31) / 32, 32 * kRowGroups, 0, ST>>>(rows_ws, grid, C, out, 0);
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
  }
  return MSP_OK;
}
