// Convolution forward / dgrad / wgrad as implicit GEMMs on the Blackwell tensor cores.
//
// Design:
//  * "tap-GEMM": a convolution is a sum over filter taps of [pixels x C] * [C x K] GEMMs.  For one
//    output tile (a bw x bh x bn box of output pixels, <= 128 rows) and one tap, the A operand is
//    the same box of the NHWC input shifted by the tap offset — a single 4-D TMA tile load with
//    hardware zero fill for the padding halo and `elementStrides` for stride-2 convs.  The B operand
//    is a [BN x 64] slab of the packed weights [K][tap][C] (3-D TMA).  Both land in shared memory
//    in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
//  * warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane,
//    tcgen05.mma cta_group::1, M=128, fp32 accumulators in TMEM), warps 2-5 = epilogue
//    (tcgen05.ld -> bias/ReLU -> bf16 -> global, plus per-channel sum / sum-of-squares for the
//    following BatchNorm).  Two CTAs are co-resident per SM so one CTA's epilogue overlaps the
//    other's main loop.
//  * dgrad = the same kernel on dy with transposed weights; stride-2 dgrad is decomposed into the
//    four output-parity classes, each a small stride-1 tap-GEMM writing a strided sub-grid of dx.
//  * wgrad contracts over pixels: A = dy tile, B = shifted x tile, both consumed MN-major straight
//    from the same TMA boxes; split-K over pixel tiles, fp32 atomics into the packed dW.
//
// Replaces cuDNN conv fwd / bwd-data / bwd-filter behind nn.Conv2d
// (reference classification/models.py:43-46,161-179,234-253; segmentation/models/blocks.py:458,518,590).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

constexpr int kMaxTaps = 49;
constexpr int kBM = 128;  // UMMA M (rows of the output tile)
constexpr int kBK = 64;   // contraction elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kConvThreads = 192;
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB

struct TapGemmParams {
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n;
  int OWs, OHs, N;  // output sub-grid extent
  int sx;           // A coordinate multiplier (conv stride for fprop, 1 for dgrad)
  int C;            // contraction channels per tap
  int ntaps;
  int Kout;         // valid output channels
  int relu;
  int accumulate;  // epilogue adds into the existing output (dgrad on top of a residual grad)
  long long y_off, y_n_stride, y_h_stride, y_w_stride;  // element strides of the output sub-grid
  __nv_bfloat16* y;
  const float* bias;
  float* ch_sum;
  float* ch_sqsum;
  int8_t tap_dh[kMaxTaps];
  int8_t tap_dw[kMaxTaps];
  uint8_t tap_w[kMaxTaps];
};

template <int BN_>
struct TapGemmCfg {
  static constexpr int kBTileBytes = BN_ * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = BN_ >= 256 ? 2 : (BN_ >= 128 ? 3 : 4);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = BN_ < 32 ? 32 : BN_;
  static constexpr int kChunk = BN_ < 32 ? 16 : 32;  // epilogue column chunk
};

template <int BN_>
__global__ void __launch_bounds__(kConvThreads, 2)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TapGemmParams p) {
  using Cfg = TapGemmCfg<BN_>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t full_bar[Cfg::kStages];
  __shared__ uint64_t empty_bar[Cfg::kStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tm = blockIdx.x;
  const int tw = tm % p.tiles_w;
  const int th = (tm / p.tiles_w) % p.tiles_h;
  const int tn = tm / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
  const int co0 = blockIdx.y * BN_;
  const int chunks = (p.C + kBK - 1) / kBK;
  const int kiters = p.ntaps * chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)Cfg::kBTileBytes;
      for (int it = 0; it < kiters; ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (uint32_t)(it / Cfg::kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tap = it / chunks;
        const int ch = it - tap * chunks;
        uint8_t* a_s = smem + s * Cfg::kStageBytes;
        uint8_t* b_s = a_s + kATileBytes;
        mbar_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(a_s, &tmA, &full_bar[s], ch * kBK, w0 * p.sx + p.tap_dw[tap],
                    h0 * p.sx + p.tap_dh[tap], n0);
        tma_load_3d(b_s, &tmB, &full_bar[s], ch * kBK, (int)p.tap_w[tap], co0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN_, 0, 0);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % Cfg::kStages;
        const uint32_t ph = (uint32_t)(it / Cfg::kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const int ch = it % chunks;
        int kvalid = p.C - ch * kBK;
        kvalid = kvalid > kBK ? kBK : kvalid;
        const int nk = (kvalid + 15) >> 4;
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint64_t adesc = umma_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t bdesc = umma_smem_desc_sw128(a_addr + kATileBytes, 16, 1024);
        for (int k = 0; k < nk; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                    (uint32_t)((it | k) != 0));
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int wi = row % p.bw;
    const int t2 = row / p.bw;
    const int hi = t2 % p.bh;
    const int ni = t2 / p.bh;
    const bool valid = row < p.rows && (w0 + wi) < p.OWs && (h0 + hi) < p.OHs && (n0 + ni) < p.N;
    __nv_bfloat16* yrow = p.y + p.y_off + (long long)(n0 + ni) * p.y_n_stride +
                          (long long)(h0 + hi) * p.y_h_stride + (long long)(w0 + wi) * p.y_w_stride;
    float* scratch = reinterpret_cast<float*>(smem) + q * (32 * 33);
    const bool do_stats = p.ch_sum != nullptr;
    constexpr int CW = Cfg::kChunk;
#pragma unroll 1
    for (int c = 0; c < BN_; c += CW) {
      const int cg = co0 + c;
      if (cg >= p.Kout) break;
      uint32_t v[CW];
      if constexpr (CW == 32) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      else tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
      tmem_ld_wait();
      float f[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        float x = __uint_as_float(v[j]);
        if (p.bias != nullptr && cg + j < p.Kout) x += __ldg(p.bias + cg + j);
        if (p.relu) x = fmaxf(x, 0.f);
        f[j] = x;
      }
      uint32_t pk[CW / 2];
#pragma unroll
      for (int j = 0; j < CW / 2; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
      if (valid) {
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
          if (cg + g * 8 < p.Kout) {
            uint4 o = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
            if (p.accumulate) {
              const uint4 old = *reinterpret_cast<const uint4*>(yrow + cg + g * 8);
              const uint32_t ov[4] = {old.x, old.y, old.z, old.w};
              uint32_t nv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 a = unpack_bf16x2(ov[e]);
                nv[e] = pack_bf16x2(a.x + f[g * 8 + 2 * e], a.y + f[g * 8 + 2 * e + 1]);
              }
              o = make_uint4(nv[0], nv[1], nv[2], nv[3]);
            }
            st_v4(yrow + cg + g * 8, o);
          }
        }
      }
      if (do_stats) {
        // per-channel sums over this warp's 32 rows via a padded smem transpose
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) {
          float2 r = unpack_bf16x2(pk[j]);
          scratch[lane * 33 + 2 * j] = valid ? r.x : 0.f;
          scratch[lane * 33 + 2 * j + 1] = valid ? r.y : 0.f;
        }
        __syncwarp();
        if (lane < CW) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const float x = scratch[i * 33 + lane];
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          if (cg + lane < p.Kout) {
            atomicAdd(p.ch_sum + cg + lane, s1);
            atomicAdd(p.ch_sqsum + cg + lane, s2);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[co][tap][ci] = sum_pix dY[pix][co] * X[pix + tap][ci]
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n, tiles_m;
  int sx, pad_t, pad_l, KW;
  int C, Cw;   // stored input channels, packed weight inner dim
  int Kout;    // output channels
  int ntaps, chunks;
  int splits;
  float* dw;
};

constexpr int kWgStages = 2;
constexpr int kWgStageBytes = 3 * kATileBytes;  // dY halves (2 x 16 KB) + X tile (16 KB)
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 1024;

__global__ void __launch_bounds__(kConvThreads, 2)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kWgStages];
  __shared__ uint64_t empty_bar[kWgStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tap = blockIdx.x / p.chunks;
  const int ci0 = (blockIdx.x - tap * p.chunks) * 64;
  const int co0 = blockIdx.y * 128;
  const int r = tap / p.KW, qx = tap - r * p.KW;
  const int split = blockIdx.z;
  const int my_tiles = (p.tiles_m - split + p.splits - 1) / p.splits;

  // rows beyond the TMA box are never written: keep them zero so partial 16-row MMA steps add 0
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kWgStages * kWgStageBytes / 16; i += kConvThreads)
      reinterpret_cast<uint4*>(smem)[i] = z;
    fence_proxy_async_smem();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<64>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmDY);
      tma_prefetch_desc(&tmX);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u * 3u;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kWgStages;
        const uint32_t ph = (uint32_t)(it / kWgStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int tm = split + it * p.splits;
        const int tw = tm % p.tiles_w;
        const int th = (tm / p.tiles_w) % p.tiles_h;
        const int tn = tm / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        uint8_t* a_s = smem + s * kWgStageBytes;
        mbar_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(a_s, &tmDY, &full_bar[s], co0, w0, h0, n0);
        tma_load_4d(a_s + kATileBytes, &tmDY, &full_bar[s], co0 + 64, w0, h0, n0);
        tma_load_4d(a_s + 2 * kATileBytes, &tmX, &full_bar[s], ci0, w0 * p.sx - p.pad_l + qx,
                    h0 * p.sx - p.pad_t + r, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      const int nk = (p.rows + 15) >> 4;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kWgStages;
        const uint32_t ph = (uint32_t)(it / kWgStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kWgStageBytes);
        // MN-major: 8-row (K) groups 1024 B apart; the two 64-channel halves of dY are 16 KB apart
        const uint64_t adesc = umma_smem_desc_sw128(a_addr, kATileBytes, 1024);
        const uint64_t bdesc = umma_smem_desc_sw128(a_addr + 2 * kATileBytes, kATileBytes, 1024);
        for (int k = 0; k < nk; ++k)
          umma_bf16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc,
                    (uint32_t)((it | k) != 0));
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  } else {
    if (my_tiles > 0) {
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
      const int q = warp & 3;
      const int co = co0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
        tmem_ld_wait();
        if (co < p.Kout) {
          float* dst = p.dw + ((long long)co * p.ntaps + tap) * p.Cw + ci0 + c;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + c + j < p.Cw) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_w_fprop_kernel(const float* __restrict__ w, int K, int C, int taps, int Cpad,
                                    __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)K * taps * Cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long t = i / Cpad;
    const int tap = (int)(t % taps);
    const int k = (int)(t / taps);
    const float v = c < C ? w[((long long)k * C + c) * taps + tap] : 0.f;
    out[i] = __float2bfloat16_rn(v);
  }
}
__global__ void pack_w_dgrad_kernel(const float* __restrict__ w, int K, int C, int taps, int Cpad,
                                    int Kpad, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)Cpad * taps * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kpad);
    const long long t = i / Kpad;
    const int tap = (int)(t % taps);
    const int c = (int)(t / taps);
    const float v = (c < C && k < K) ? w[((long long)k * C + c) * taps + tap] : 0.f;
    out[i] = __float2bfloat16_rn(v);
  }
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int K, int C, int taps, int Cpad,
                                    float* __restrict__ out) {
  const long long total = (long long)K * C * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long t = i / taps;
    const int c = (int)(t % C);
    const int k = (int)(t / C);
    out[i] = dwp[((long long)k * taps + tap) * Cpad + c];
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Box {
  int bw, bh, bn;
};

// Largest-utilisation box of <= 128 output pixels: full rows first, then rows, then images.
Box pick_box(int OW, int OH, int N) {
  Box b{1, 1, 1};
  if (OW >= 128) {
    const int t = msp_cdiv(OW, 128);
    b.bw = msp_cdiv(OW, t);
    return b;
  }
  b.bw = OW;
  const int maxh = 128 / OW;
  if (OH > maxh) {
    const int t = msp_cdiv(OH, maxh);
    b.bh = msp_cdiv(OH, t);
    return b;
  }
  b.bh = OH;
  const int maxn = 128 / (OW * OH);
  if (N > maxn) {
    const int t = msp_cdiv(N, maxn);
    b.bn = msp_cdiv(N, t);
  } else {
    b.bn = N;
  }
  return b;
}

template <int BN_>
int launch_tapgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const TapGemmParams& p,
                   cudaStream_t st) {
  using Cfg = TapGemmCfg<BN_>;
  static bool attr_set = false;
  if (!attr_set) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN_>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, msp_cdiv(p.Kout, BN_), 1);
  tapgemm_kernel<BN_><<<grid, kConvThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

int dispatch_tapgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const TapGemmParams& p,
                     cudaStream_t st) {
  if (p.Kout <= 16) return launch_tapgemm<16>(tmA, tmB, p, st);
  if (p.Kout <= 32) return launch_tapgemm<32>(tmA, tmB, p, st);
  if (p.Kout <= 64) return launch_tapgemm<64>(tmA, tmB, p, st);
  return launch_tapgemm<128>(tmA, tmB, p, st);
}

// A-operand tensor map over an NHWC bf16 tensor (C, W, H, N) with a (64, bw*s, bh*s, bn) box.
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, int cs, Box b,
                 int s) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)cs * 2, (uint64_t)W * cs * 2, (uint64_t)H * W * cs * 2};
  uint32_t box[4] = {64, (uint32_t)(b.bw * s), (uint32_t)(b.bh * s), (uint32_t)b.bn};
  uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
  return msp_encode_tmap_bf16(m, base, 4, dims, strides, box, es, 128);
}
// B-operand map over packed weights [rows][taps][inner]: (inner, taps, rows), box (64, 1, box_rows).
int make_w_map(CUtensorMap* m, const void* base, int inner, int taps, int rows, int box_rows) {
  uint64_t dims[3] = {(uint64_t)inner, (uint64_t)taps, (uint64_t)rows};
  uint64_t strides[2] = {(uint64_t)inner * 2, (uint64_t)taps * inner * 2};
  uint32_t box[3] = {64, 1, (uint32_t)box_rows};
  uint32_t es[3] = {1, 1, 1};
  return msp_encode_tmap_bf16(m, base, 3, dims, strides, box, es, 128);
}

int check_desc(const msp_conv_desc* d) {
  MSP_REQUIRE(d != nullptr, "conv: null descriptor");
  MSP_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0, "conv: empty tensor");
  MSP_REQUIRE(d->C > 0 && d->C % 8 == 0 && d->x_cs % 8 == 0 && d->x_cs >= d->C,
              "conv: input channels / pixel stride must be multiples of 8 (C=%d cs=%d)", d->C,
              d->x_cs);
  MSP_REQUIRE(d->K > 0 && d->K % 8 == 0 && d->y_cs % 8 == 0 && d->y_cs >= d->K,
              "conv: output channels / pixel stride must be multiples of 8 (K=%d cs=%d)", d->K,
              d->y_cs);
  MSP_REQUIRE(d->KH >= 1 && d->KW >= 1 && d->KH * d->KW <= kMaxTaps, "conv: filter %dx%d too large",
              d->KH, d->KW);
  MSP_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride %d unsupported", d->stride);
  MSP_REQUIRE(d->pad_t >= 0 && d->pad_l >= 0 && d->pad_t < 64 && d->pad_l < 64, "conv: bad padding");
  return MSP_OK;
}

inline int bn_tile_for(int K) { return K <= 16 ? 16 : (K <= 32 ? 32 : (K <= 64 ? 64 : 128)); }

}  // namespace

extern "C" int msp_pack_weights(const float* w, int K, int C, int KH, int KW, int Cpad, int Kpad,
                                void* w_fprop, void* w_dgrad, void* stream) {
  MSP_REQUIRE(w && (w_fprop || w_dgrad), "pack_weights: null pointer");
  MSP_REQUIRE(Cpad >= C && Cpad % 8 == 0 && Kpad >= K && Kpad % 8 == 0,
              "pack_weights: padded sizes must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = KH * KW;
  if (w_fprop) {
    const long long total = (long long)K * taps * Cpad;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_w_fprop_kernel<<<blocks, 256, 0, st>>>(w, K, C, taps, Cpad, (__nv_bfloat16*)w_fprop);
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
  }
  if (w_dgrad) {
    const long long total = (long long)Cpad * taps * Kpad;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pack_w_dgrad_kernel<<<blocks, 256, 0, st>>>(w, K, C, taps, Cpad, Kpad,
                                                (__nv_bfloat16*)w_dgrad);
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
  }
  return MSP_OK;
}

extern "C" int msp_unpack_wgrad(const float* dwp, int K, int C, int KH, int KW, int Cpad,
                                float* dw_oihw, void* stream) {
  MSP_REQUIRE(dwp && dw_oihw, "unpack_wgrad: null pointer");
  const int taps = KH * KW;
  const long long total = (long long)K * C * taps;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  unpack_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dwp, K, C, taps, Cpad, dw_oihw);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_conv_fprop(const msp_conv_desc* d, const void* x, const void* w_fprop,
                              const float* bias, void* y, float* ch_sum, float* ch_sqsum,
                              void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && w_fprop && y, "conv_fprop: null pointer");
  MSP_REQUIRE((ch_sum == nullptr) == (ch_sqsum == nullptr), "conv_fprop: need both stat buffers");
  const int taps = d->KH * d->KW;
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = (taps == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 &&
                     d->Ho == d->H && d->Wo == d->W);
  CUtensorMap tmA, tmB;
  Box b;
  if (flat) {
    // 1x1 stride-1: pixels form one long row -> perfectly filled 128-row tiles
    const long long P = (long long)d->N * d->H * d->W;
    MSP_REQUIRE(P < (1ll << 31), "conv_fprop: too many pixels");
    b = pick_box((int)P, 1, 1);
    rc = make_act_map(&tmA, x, d->C, (int)P, 1, 1, d->x_cs, b, 1);
    p.OWs = (int)P; p.OHs = 1; p.N = 1;
    p.y_n_stride = 0; p.y_h_stride = 0; p.y_w_stride = d->y_cs;
  } else {
    b = pick_box(d->Wo, d->Ho, d->N);
    rc = make_act_map(&tmA, x, d->C, d->W, d->H, d->N, d->x_cs, b, d->stride);
    p.OWs = d->Wo; p.OHs = d->Ho; p.N = d->N;
    p.y_n_stride = (long long)d->Ho * d->Wo * d->y_cs;
    p.y_h_stride = (long long)d->Wo * d->y_cs;
    p.y_w_stride = d->y_cs;
  }
  if (rc) return rc;
  rc = make_w_map(&tmB, w_fprop, d->C, taps, d->K, bn_tile_for(d->K));
  if (rc) return rc;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
  p.tiles_w = msp_cdiv(p.OWs, b.bw); p.tiles_h = msp_cdiv(p.OHs, b.bh); p.tiles_n = msp_cdiv(p.N, b.bn);
  p.sx = flat ? 1 : d->stride;
  p.C = d->C; p.ntaps = taps; p.Kout = d->K; p.relu = d->relu;
  p.y = (__nv_bfloat16*)y; p.y_off = 0;
  p.bias = bias; p.ch_sum = ch_sum; p.ch_sqsum = ch_sqsum;
  for (int r = 0; r < d->KH; ++r)
    for (int q = 0; q < d->KW; ++q) {
      const int t = r * d->KW + q;
      p.tap_dh[t] = (int8_t)(r - d->pad_t);
      p.tap_dw[t] = (int8_t)(q - d->pad_l);
      p.tap_w[t] = (uint8_t)t;
    }
  return dispatch_tapgemm(tmA, tmB, p, (cudaStream_t)stream);
}

extern "C" int msp_conv_dgrad(const msp_conv_desc* d, const void* dy, const void* w_dgrad, void* dx,
                              int accumulate, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(dy && w_dgrad && dx, "conv_dgrad: null pointer");
  const int taps = d->KH * d->KW;
  const int s = d->stride;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  // weights [Cpad = d->C][taps][Kpad = d->K]: contraction over K (dy channels), outputs = C
  rc = make_w_map(&tmB, w_dgrad, d->K, taps, d->C, bn_tile_for(d->C));
  if (rc) return rc;
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      TapGemmParams p;
      memset(&p, 0, sizeof(p));
      const int OHs = (d->H - ph + s - 1) / s, OWs = (d->W - pw + s - 1) / s;
      if (OHs <= 0 || OWs <= 0) continue;
      int nt = 0;
      for (int r = 0; r < d->KH; ++r) {
        if ((ph + d->pad_t - r) % s != 0) continue;
        for (int q = 0; q < d->KW; ++q) {
          if ((pw + d->pad_l - q) % s != 0) continue;
          const int dh = (ph + d->pad_t - r) / s, dw = (pw + d->pad_l - q) / s;
          p.tap_dh[nt] = (int8_t)dh;
          p.tap_dw[nt] = (int8_t)dw;
          p.tap_w[nt] = (uint8_t)(r * d->KW + q);
          ++nt;
        }
      }
      if (nt == 0) {
        msp_set_error("conv_dgrad: parity class (%d,%d) receives no filter tap (k=%dx%d s=%d)", ph,
                      pw, d->KH, d->KW, s);
        return MSP_ERR_UNSUPPORTED;
      }
      const bool flat = (taps == 1 && s == 1 && d->pad_t == 0 && d->pad_l == 0 && d->Ho == d->H &&
                         d->Wo == d->W);
      Box b;
      if (flat) {
        const long long P = (long long)d->N * d->H * d->W;
        b = pick_box((int)P, 1, 1);
        rc = make_act_map(&tmA, dy, d->K, (int)P, 1, 1, d->y_cs, b, 1);
        p.OWs = (int)P; p.OHs = 1; p.N = 1;
        p.y_w_stride = d->x_cs;
      } else {
        b = pick_box(OWs, OHs, d->N);
        rc = make_act_map(&tmA, dy, d->K, d->Wo, d->Ho, d->N, d->y_cs, b, 1);
        p.OWs = OWs; p.OHs = OHs; p.N = d->N;
        p.y_n_stride = (long long)d->H * d->W * d->x_cs;
        p.y_h_stride = (long long)s * d->W * d->x_cs;
        p.y_w_stride = (long long)s * d->x_cs;
        p.y_off = ((long long)ph * d->W + pw) * d->x_cs;
      }
      if (rc) return rc;
      p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
      p.tiles_w = msp_cdiv(p.OWs, b.bw); p.tiles_h = msp_cdiv(p.OHs, b.bh);
      p.tiles_n = msp_cdiv(p.N, b.bn);
      p.sx = 1; p.C = d->K; p.ntaps = nt; p.Kout = d->C; p.relu = 0; p.accumulate = accumulate;
      p.y = (__nv_bfloat16*)dx;
      rc = dispatch_tapgemm(tmA, tmB, p, st);
      if (rc) return rc;
    }
  return MSP_OK;
}

extern "C" int msp_conv_wgrad(const msp_conv_desc* d, const void* x, const void* dy,
                              float* dw_packed, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && dy && dw_packed, "conv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = d->KH * d->KW;
  static bool attr_set = false;
  if (!attr_set) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kWgSmemBytes));
    attr_set = true;
  }
  MSP_CHECK_CUDA(cudaMemsetAsync(dw_packed, 0, sizeof(float) * (size_t)d->K * taps * d->C, st));
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = (taps == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 &&
                     d->Ho == d->H && d->Wo == d->W);
  CUtensorMap tmDY, tmX;
  Box b;
  int OW, OH, N;
  if (flat) {
    const long long P = (long long)d->N * d->H * d->W;
    MSP_REQUIRE(P < (1ll << 31), "conv_wgrad: too many pixels");
    OW = (int)P; OH = 1; N = 1;
    b = pick_box(OW, OH, N);
    rc = make_act_map(&tmDY, dy, d->K, OW, 1, 1, d->y_cs, b, 1);
    if (rc) return rc;
    rc = make_act_map(&tmX, x, d->C, OW, 1, 1, d->x_cs, b, 1);
    if (rc) return rc;
    p.sx = 1;
  } else {
    OW = d->Wo; OH = d->Ho; N = d->N;
    b = pick_box(OW, OH, N);
    rc = make_act_map(&tmDY, dy, d->K, d->Wo, d->Ho, d->N, d->y_cs, b, 1);
    if (rc) return rc;
    rc = make_act_map(&tmX, x, d->C, d->W, d->H, d->N, d->x_cs, b, d->stride);
    if (rc) return rc;
    p.sx = d->stride;
  }
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
  p.tiles_w = msp_cdiv(OW, b.bw); p.tiles_h = msp_cdiv(OH, b.bh); p.tiles_n = msp_cdiv(N, b.bn);
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.KW = d->KW;
  p.C = d->C; p.Cw = d->C; p.Kout = d->K; p.ntaps = taps; p.chunks = msp_cdiv(d->C, 64);
  p.dw = dw_packed;
  const int gx = taps * p.chunks, gy = msp_cdiv(d->K, 128);
  int splits = msp_cdiv(4 * msp_num_sms(), gx * gy);
  if (splits > p.tiles_m) splits = p.tiles_m;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  p.splits = splits;
  dim3 grid(gx, gy, splits);
  wgrad_kernel<<<grid, kConvThreads, kWgSmemBytes, st>>>(tmDY, tmX, p);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
