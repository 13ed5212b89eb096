// Narrow-channel convolutions: the full-resolution decoder levels of the U-Nets (16 / 32 channels at 256 x 256 and
// 128 x 128: UNet_decoder's last two levels, segmentation/models/unet_models.py:330-371 with decoder_channels
// (..., 32, 16), config/downstream/acdc/resnet50_attention_unet.yaml:43).
//
// These layers are HBM bound (16 -> 16 channels 3x3 on 1.57 M pixels: 7 GFLOP for 100 MB) and a poor fit for the
// tcgen05 path: one tcgen05.mma occupies the tensor pipe for >= 64 cycles whatever its width (profiles/
// r01_mma_issue_rate.txt), so a 128 x 16 x 16 instruction does 1/16 of the work of a 128 x 256 x 16 one in half its
// time, and the weight gradient pads its 16 output channels to the 128-row instruction on top (1/32 useful:
// 246 us for d4_c0, profiles/r02_convbench_unet50_b24.txt).  Here the contraction runs on warp-level
// mma.sync.m16n8k16 (bf16 -> fp32) instead: many warps per SM, every instruction fully used, operands staged once per
// tile in shared memory (cp.async, zero fill = the convolution's padding) and read through ldmatrix.
//
// wgrad:  dW[k][tap][c] = sum over pixels of dy[pixel][k] * x[pixel + tap][c].  A CTA walks 8 x 32-pixel tiles (dy tile
// + x halo tile, double buffered), its warps split the tile's pixels and the (16 k) x (16 c) blocks of the result,
// every warp keeps its block for ALL taps in registers across all the tiles of the CTA; at the end the warps' blocks are
// added in fixed order through shared memory and stored as the CTA's row of the split-K workspace that
// msp_unpack_wgrad_batched sums in fixed order (deterministic, like wgrad_kernel).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

constexpr int kNwTH = 8, kNwTW = 32;            // output pixels per tile
constexpr int kNwPix = kNwTH * kNwTW;           // 256
constexpr int kNwHaloW = kNwTW + 2, kNwHaloH = kNwTH + 2;  // room for 3 x 3 taps
constexpr int kNwThreads = 256;

struct NarrowWgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* part;       // [grid][K][taps][C]
  int N, H, W, Ho, Wo, C, K, KH, KW, pad_t, pad_l, x_cs, y_cs;
  int tiles_w, tiles_h, total_tiles;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled (padding / ragged tile edges)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared-memory tiles are [pixel][CH channels] bf16 with a PADDED pixel pitch of 2*CH + 16 bytes (48 / 80): the 8 rows
// (consecutive pixels) of an ldmatrix 8 x 8 block then fall into 8 different bank groups without an XOR swizzle, so a
// filter tap is a CONSTANT byte offset from the k step's base address (the first version spent 17 instructions per MMA,
// most of them swizzle / index arithmetic: ncu issue-bound at 22 % tensor pipe, profiles/r02_narrow_wgrad.txt).
template <int CH>
struct NwPitch { static constexpr int v = CH * 2 + 16; };

template <int C_, int K_, int KH_>
struct NwCfg {
  static constexpr int kDyBytes = kNwPix * NwPitch<K_>::v;
  static constexpr int kXBytes = (kNwTH + KH_ - 1) * kNwHaloW * NwPitch<C_>::v;
  static constexpr int kStage = kDyBytes + kXBytes;
  static constexpr int kStages = 3 * kStage <= 112 * 1024 ? 3 : 2;     // two CTAs per SM
  static constexpr int kRed = K_ * 9 * C_ * 4;
  static constexpr int kSmem = kStages * kStage > kRed ? kStages * kStage : kRed;
};

// C_ = input channels (16 | 32), K_ = output channels (16 | 32), KH_ x KW_ filter (compile time: the tap loop is
// straight-line code with immediate offsets)
template <int C_, int K_, int KH_, int KW_>
__global__ void __launch_bounds__(kNwThreads, 2) narrow_wgrad_kernel(const NarrowWgradParams p) {
  using Cfg = NwCfg<C_, K_, KH_>;
  constexpr int kCombos = (K_ / 16) * (C_ / 16);        // (16 k) x (16 c) result blocks
  constexpr int kPG = 8 / kCombos;                      // pixel groups (warps per block of the result)
  constexpr int kPixPerWarp = kNwPix / kPG;             // 32, 64 or 128
  constexpr int PK = NwPitch<K_>::v, PC = NwPitch<C_>::v;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTaps = KH_ * KW_;
  constexpr int kHH = kNwTH + KH_ - 1, kHW = kNwTW + KW_ - 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int combo = warp % kCombos, pg = warp / kCombos;
  const int mt = combo / (C_ / 16), nt = combo % (C_ / 16);  // 16-row k block, 16-column c block of this warp

  float acc[kTaps][2][4];
#pragma unroll
  for (int t = 0; t < kTaps; ++t)
#pragma unroll
    for (int h = 0; h < 2; ++h) acc[t][h][0] = acc[t][h][1] = acc[t][h][2] = acc[t][h][3] = 0.f;

  auto load_tile = [&](int tile, int stage) {
    const int tw = tile % p.tiles_w, r1 = tile / p.tiles_w, th = r1 % p.tiles_h, n = r1 / p.tiles_h;
    const int h0 = th * kNwTH, w0 = tw * kNwTW;
    const uint32_t dy_s = s0 + stage * Cfg::kStage, x_s = dy_s + Cfg::kDyBytes;
    // dy tile: 256 pixels x K_/8 chunks; 32-bit offsets relative to the tile's first pixel
    const __nv_bfloat16* dyb = p.dy + ((long long)(n * p.Ho + h0) * p.Wo + w0) * p.y_cs;
    const int hmax = p.Ho - h0, wmax = p.Wo - w0;
#pragma unroll
    for (int it = 0; it < K_ / 8; ++it) {
      const int i = threadIdx.x + it * kNwThreads;
      const int pix = i / (K_ / 8), ch = i % (K_ / 8), py = pix / kNwTW, px = pix % kNwTW;
      const bool ok = py < hmax && px < wmax;
      cp_async16(dy_s + pix * PK + ch * 16, ok ? dyb + (py * p.Wo + px) * p.y_cs + ch * 8 : p.dy, ok);
    }
    // x halo tile: kHH x kHW pixels (pitch kNwHaloW) x C_/8 chunks, origin (h0 - pad_t, w0 - pad_l)
    const int hb = h0 - p.pad_t, wb = w0 - p.pad_l;
    const __nv_bfloat16* xb = p.x + ((long long)(n * p.H + hb) * p.W + wb) * p.x_cs;   // may point before the image: only
    constexpr int kXChunks = kHH * kNwHaloW * (C_ / 8);                               // dereferenced where `ok`
#pragma unroll
    for (int it = 0; it < (kXChunks + kNwThreads - 1) / kNwThreads; ++it) {
      const int i = threadIdx.x + it * kNwThreads;
      const int ch = i % (C_ / 8), q = i / (C_ / 8), hc = q % kNwHaloW, hr = q / kNwHaloW;
      if (i < kXChunks && hc < kHW) {
        const bool ok = (unsigned)(hb + hr) < (unsigned)p.H && (unsigned)(wb + hc) < (unsigned)p.W;
        cp_async16(x_s + q * PC + ch * 16, ok ? xb + (hr * p.W + hc) * p.x_cs + ch * 8 : p.x, ok);
      }
    }
  };

  // per-lane ldmatrix row addresses (matrix j = lane / 8, row lane % 8), relative to the k step's base
  const int j = lane >> 3, rr = lane & 7;
  const uint32_t a_lane = (uint32_t)(((j >> 1) * 8 + rr) * PK + (mt * 2 + (j & 1)) * 16);   // (k block j&1, pixel block j>>1)
  const uint32_t b_lane = (uint32_t)(((j & 1) * 8 + rr) * PC + (nt * 2 + (j >> 1)) * 16);   // (pixel block j&1, c block j>>1)

  // kStages-deep ring: the loads of the next kStages - 1 tiles are in flight while one tile is multiplied
  int stage = 0, load_stage = 0;
  int tile = blockIdx.x, load_tile_i = blockIdx.x;
#pragma unroll 1
  for (int s_ = 0; s_ < kStages - 1; ++s_) {
    if (load_tile_i < p.total_tiles) load_tile(load_tile_i, load_stage);
    cp_async_commit();
    load_tile_i += gridDim.x;
    load_stage = load_stage + 1 == kStages ? 0 : load_stage + 1;
  }
#pragma unroll 1
  for (; tile < p.total_tiles; tile += gridDim.x) {
    // ONE block barrier per tile: it publishes this tile's data AND proves that every warp has finished the previous
    // tile, whose stage the load issued right after it overwrites (kStages >= 2)
    cp_async_wait<kStages - 2>();   // this tile's group has landed (the later ones may still be in flight)
    __syncthreads();
    if (load_tile_i < p.total_tiles) load_tile(load_tile_i, load_stage);
    cp_async_commit();
    load_tile_i += gridDim.x;
    load_stage = load_stage + 1 == kStages ? 0 : load_stage + 1;
    const uint32_t dy_s = s0 + stage * Cfg::kStage, x_s = dy_s + Cfg::kDyBytes;
#pragma unroll 2
    for (int ks = 0; ks < kPixPerWarp / 16; ++ks) {
      const int p0 = pg * kPixPerWarp + ks * 16;          // first pixel of this 16-pixel k step (half a tile row)
      const int th = p0 / kNwTW, tw0 = p0 % kNwTW;
      // A = dy^T (16 k x 16 pixels), stored [pixel][k] -> .trans;  B = x shifted by the tap (16 pixels x 16 c) -> .trans
      uint32_t a[4];
      ldsm_x4_t(dy_s + (uint32_t)(p0 * PK) + a_lane, a);
      const uint32_t xk = x_s + (uint32_t)((th * kNwHaloW + tw0) * PC) + b_lane;
#pragma unroll
      for (int r = 0; r < KH_; ++r)
#pragma unroll
        for (int q = 0; q < KW_; ++q) {
          uint32_t b[4];
          ldsm_x4_t(xk + (uint32_t)((r * kNwHaloW + q) * PC), b);
          mma_bf16_16816(acc[r * KW_ + q][0], a, b[0], b[1]);
          mma_bf16_16816(acc[r * KW_ + q][1], a, b[2], b[3]);
        }
    }
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  __syncthreads();

  // fixed-order reduction over the pixel groups into a [K][taps][C] fp32 block in shared memory (aliases the stages)
  float* red = reinterpret_cast<float*>(smem);
  const int g = lane >> 2, t4 = lane & 3;
  for (int round = 0; round < kPG; ++round) {
    if (pg == round) {
#pragma unroll
      for (int tap = 0; tap < kTaps; ++tap)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = mt * 16 + g + (e >> 1) * 8, c = nt * 16 + h * 8 + t4 * 2 + (e & 1);
            float* dst = red + ((size_t)k * kTaps + tap) * C_ + c;
            *dst = round == 0 ? acc[tap][h][e] : *dst + acc[tap][h][e];
          }
    }
    __syncthreads();
  }
  float* out = p.part + (size_t)blockIdx.x * K_ * kTaps * C_;
  for (int i = threadIdx.x; i < K_ * kTaps * C_ / 4; i += kNwThreads)
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(red)[i];
}

// ------------------------------------------------------------------------------------------------
// fprop / dgrad:  y[pixel][k] = sum over taps and c of x[pixel + tap][c] * w[k][tap][c]  (+ bias, ReLU, BatchNorm
// statistics).  One CTA of 16 warps per SM walks 16 x 32-pixel tiles (x halo tile in a 3-deep cp.async ring, the packed
// weights [K][taps][C] resident in shared memory); warp w owns row w of the tile: two 16-pixel MMA row blocks x all
// output channels, accumulators in registers, epilogue straight from the MMA fragments (4-byte stores: a quad covers a
// 16-byte run of one pixel).  dgrad of a stride-1 convolution is the same kernel on dy with the dgrad operand
// [C][taps][K], the taps walked in reverse and padding KH - 1 - pad.
// ------------------------------------------------------------------------------------------------
constexpr int kNfTH = 16, kNfTW = 32, kNfThreads = 512;

struct NarrowFpropParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;    // [K_][taps][C_] bf16
  __nv_bfloat16* y;
  const float* bias;
  float* ch_sum;
  float* ch_sqsum;
  long long stat_row;        // > 0: deterministic statistics, row blockIdx.x of the [rows][2][K] workspace
  int N, IH, IW, OH, OW, pad_t, pad_l, x_cs, y_cs, relu, flip;
  int tiles_w, tiles_h, total_tiles;
};

template <int C_, int K_, int KH_>
struct NfCfg {
  static constexpr int kXBytes = (kNfTH + KH_ - 1) * kNwHaloW * NwPitch<C_>::v;
  static constexpr int kWBytes = KH_ * KH_ * K_ * NwPitch<C_>::v;
  static constexpr int kStages = 3;
  static constexpr int kSmem = kStages * kXBytes + kWBytes + 2 * K_ * 16 * 4;   // + per-warp statistic rows
};

template <int C_, int K_, int KH_, int KW_>
__global__ void __launch_bounds__(kNfThreads, 1) narrow_fprop_kernel(const NarrowFpropParams p) {
  using Cfg = NfCfg<C_, K_, KH_>;
  constexpr int PC = NwPitch<C_>::v;
  constexpr int kTaps = KH_ * KW_;
  constexpr int kHH = kNfTH + KH_ - 1, kHW = kNfTW + KW_ - 1;
  constexpr int kNT = K_ / 8;            // 8-channel output blocks
  constexpr int kKS = C_ / 16;           // 16-channel contraction steps per tap
  constexpr int kStages = Cfg::kStages;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t s0 = smem_u32(smem);
  const uint32_t w_s = s0 + kStages * Cfg::kXBytes;
  float* stat_s = reinterpret_cast<float*>(smem + kStages * Cfg::kXBytes + Cfg::kWBytes);   // [16 warps][2][K_]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const bool do_stats = p.ch_sum != nullptr;

  // weights -> shared memory, row (tap, k) at pitch PC; `flip`: the dgrad operand's taps in reverse order
  for (int i = threadIdx.x; i < kTaps * K_ * (C_ / 8); i += kNfThreads) {
    const int ch = i % (C_ / 8), q = i / (C_ / 8), k = q % K_, tap = q / K_;
    const int src_tap = p.flip ? kTaps - 1 - tap : tap;
    cp_async16(w_s + (uint32_t)((tap * K_ + k) * PC + ch * 16), p.w + ((long long)k * kTaps + src_tap) * C_ + ch * 8, true);
  }
  // (committed together with the first tile's group below)

  auto load_tile = [&](int tile, int stage) {
    const int tw = tile % p.tiles_w, r1 = tile / p.tiles_w, th = r1 % p.tiles_h, n = r1 / p.tiles_h;
    const int hb = th * kNfTH - p.pad_t, wb = tw * kNfTW - p.pad_l;
    const uint32_t x_s = s0 + stage * Cfg::kXBytes;
    const __nv_bfloat16* xb = p.x + ((long long)(n * p.IH + hb) * p.IW + wb) * p.x_cs;   // only dereferenced where `ok`
    constexpr int kXChunks = kHH * kNwHaloW * (C_ / 8);
#pragma unroll
    for (int it = 0; it < (kXChunks + kNfThreads - 1) / kNfThreads; ++it) {
      const int i = threadIdx.x + it * kNfThreads;
      const int ch = i % (C_ / 8), q = i / (C_ / 8), hc = q % kNwHaloW, hr = q / kNwHaloW;
      if (i < kXChunks && hc < kHW) {
        const bool ok = (unsigned)(hb + hr) < (unsigned)p.IH && (unsigned)(wb + hc) < (unsigned)p.IW;
        cp_async16(x_s + q * PC + ch * 16, ok ? xb + (hr * p.IW + hc) * p.x_cs + ch * 8 : p.x, ok);
      }
    }
  };

  // ldmatrix lane addresses.  A (pixels x channels, stored [pixel][c], no transpose): matrices (pixel block j&1,
  // channel block j>>1).  B (channels x outputs, stored [k_out][c] = [n][k], no transpose): matrices (channel block j&1,
  // output block j>>1) -> registers {b0, b1} of output block 0 and {b0, b1} of output block 1.
  const int j = lane >> 3, rr = lane & 7;
  const uint32_t a_lane = (uint32_t)(((j & 1) * 8 + rr) * PC + (j >> 1) * 16);
  const uint32_t b_lane = (uint32_t)(((j >> 1) * 8 + rr) * PC + (j & 1) * 16);

  float st[kNT][4];   // running (sum col a, sum col b, sumsq a, sumsq b) of this thread's column pair per output block
#pragma unroll
  for (int i = 0; i < kNT; ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
  float bias_r[kNT][2];
#pragma unroll
  for (int i = 0; i < kNT; ++i) {
    bias_r[i][0] = p.bias ? __ldg(p.bias + i * 8 + t4 * 2) : 0.f;
    bias_r[i][1] = p.bias ? __ldg(p.bias + i * 8 + t4 * 2 + 1) : 0.f;
  }

  int stage = 0, load_stage = 0;
  int tile = blockIdx.x, load_tile_i = blockIdx.x;
#pragma unroll 1
  for (int s_ = 0; s_ < kStages - 1; ++s_) {
    if (load_tile_i < p.total_tiles) load_tile(load_tile_i, load_stage);
    cp_async_commit();
    load_tile_i += gridDim.x;
    load_stage = load_stage + 1 == kStages ? 0 : load_stage + 1;
  }
#pragma unroll 1
  for (; tile < p.total_tiles; tile += gridDim.x) {
    cp_async_wait<kStages - 2>();   // one barrier per tile, as in narrow_wgrad_kernel
    __syncthreads();
    if (load_tile_i < p.total_tiles) load_tile(load_tile_i, load_stage);
    cp_async_commit();
    load_tile_i += gridDim.x;
    load_stage = load_stage + 1 == kStages ? 0 : load_stage + 1;
    const uint32_t x_s = s0 + stage * Cfg::kXBytes;
    float acc[2][kNT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < kNT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
    const uint32_t xrow = x_s + (uint32_t)(warp * kNwHaloW * PC) + a_lane;   // tile row `warp`, tap (0, 0)
#pragma unroll
    for (int r = 0; r < KH_; ++r)
#pragma unroll
      for (int q = 0; q < KW_; ++q) {
        const uint32_t xt = xrow + (uint32_t)((r * kNwHaloW + q) * PC);
        const uint32_t wt = w_s + (uint32_t)((r * KW_ + q) * K_ * PC) + b_lane;
#pragma unroll
        for (int ks = 0; ks < kKS; ++ks) {
          uint32_t a0[4], a1[4];
          ldsm_x4(xt + ks * 32, a0);                 // pixels 0-15 of the row
          ldsm_x4(xt + 16 * PC + ks * 32, a1);       // pixels 16-31
#pragma unroll
          for (int np = 0; np < kNT / 2; ++np) {
            uint32_t b[4];
            ldsm_x4(wt + (uint32_t)(np * 16 * PC) + ks * 32, b);
            mma_bf16_16816(acc[0][2 * np], a0, b[0], b[1]);
            mma_bf16_16816(acc[0][2 * np + 1], a0, b[2], b[3]);
            mma_bf16_16816(acc[1][2 * np], a1, b[0], b[1]);
            mma_bf16_16816(acc[1][2 * np + 1], a1, b[2], b[3]);
          }
        }
      }
    // epilogue from the fragments: thread holds pixels (m * 16 + g, m * 16 + g + 8) x channels (n * 8 + 2 t4, + 1)
    {
      const int tw = tile % p.tiles_w, r1 = tile / p.tiles_w, th = r1 % p.tiles_h, n_img = r1 / p.tiles_h;
      const int ho = th * kNfTH + warp, wo0 = tw * kNfTW;
      const bool row_ok = ho < p.OH;
      __nv_bfloat16* yrow = p.y + ((long long)(n_img * p.OH + ho) * p.OW + wo0) * p.y_cs;
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int px = m * 16 + g + hh * 8;
          const bool ok = row_ok && wo0 + px < p.OW;
#pragma unroll
          for (int n = 0; n < kNT; ++n) {
            float v0 = acc[m][n][hh * 2] + bias_r[n][0], v1 = acc[m][n][hh * 2 + 1] + bias_r[n][1];
            if (p.relu) {
              v0 = fmaxf(v0, 0.f);
              v1 = fmaxf(v1, 0.f);
            }
            const uint32_t pk = pack_bf16x2(v0, v1);
            if (ok) {
              *reinterpret_cast<uint32_t*>(yrow + (long long)px * p.y_cs + n * 8 + t4 * 2) = pk;
              if (do_stats) {   // statistics of the bf16-ROUNDED outputs (what BatchNorm normalises)
                const float2 f = unpack_bf16x2(pk);
                st[n][0] += f.x;
                st[n][1] += f.y;
                st[n][2] = fmaf(f.x, f.x, st[n][2]);
                st[n][3] = fmaf(f.y, f.y, st[n][3]);
              }
            }
          }
        }
    }
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  if (do_stats) {
    // lanes with the same t4 hold the same column pair for different pixels: add over g (xor 4, 8, 16), then the 16
    // warps' rows in fixed order, one store (deterministic mode: this CTA's row) or atomic per channel and statistic
#pragma unroll
    for (int n = 0; n < kNT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = st[n][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        st[n][e] = v;
      }
    if (g == 0) {
#pragma unroll
      for (int n = 0; n < kNT; ++n) {
        float* row = stat_s + warp * 2 * K_;
        row[n * 8 + t4 * 2] = st[n][0];
        row[n * 8 + t4 * 2 + 1] = st[n][1];
        row[K_ + n * 8 + t4 * 2] = st[n][2];
        row[K_ + n * 8 + t4 * 2 + 1] = st[n][3];
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * K_) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) tot += stat_s[w * 2 * K_ + threadIdx.x];
      const int which = threadIdx.x / K_, c = threadIdx.x % K_;
      float* dst = (which ? p.ch_sqsum : p.ch_sum) + c;
      if (p.stat_row) dst[(long long)blockIdx.x * p.stat_row] = tot;
      else atomicAdd(dst, tot);
    }
  }
}

bool narrow_wgrad_ok(const msp_conv_desc* d) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MSP_NARROW");
    enabled = e ? atoi(e) : 1;
  }
  if (!enabled) return false;
  return d->win_px == 0 && d->stride == 1 && d->KH == d->KW && (d->KH == 3 || d->KH == 2) && (d->C == 16 || d->C == 32) &&
         (d->K == 16 || d->K == 32) && (long long)d->N * d->Ho * d->Wo >= 65536;
}

int narrow_wgrad_grid(const msp_conv_desc* d) {
  const long long tiles = (long long)d->N * msp_cdiv(d->Ho, kNwTH) * msp_cdiv(d->Wo, kNwTW);
  const long long cap = 2ll * msp_num_sms();   // two CTAs per SM (<= 75 KB of shared memory each)
  return (int)(tiles < cap ? tiles : cap);
}

}  // namespace

// Internal entry points (msp_conv.cu's wgrad dispatch): 0 splits = layer not eligible.
int msp_narrow_wgrad_splits(const msp_conv_desc* d) { return narrow_wgrad_ok(d) ? narrow_wgrad_grid(d) : 0; }

int msp_narrow_wgrad(const msp_conv_desc* d, const void* x, const void* dy, float* partials, void* stream) {
  MSP_REQUIRE(narrow_wgrad_ok(d), "narrow_wgrad: layer not eligible");
  NarrowWgradParams p;
  p.x = (const __nv_bfloat16*)x;
  p.dy = (const __nv_bfloat16*)dy;
  p.part = partials;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.C = d->C; p.K = d->K; p.KH = d->KH; p.KW = d->KW;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.x_cs = d->x_cs; p.y_cs = d->y_cs;
  p.tiles_w = msp_cdiv(d->Wo, kNwTW); p.tiles_h = msp_cdiv(d->Ho, kNwTH);
  p.total_tiles = d->N * p.tiles_w * p.tiles_h;
  const int grid = narrow_wgrad_grid(d);
  cudaStream_t st = (cudaStream_t)stream;
#define MSP_NW_LAUNCH(CC, KK, FF)                                                                          \
  do {                                                                                                     \
    static bool attr = false;                                                                              \
    if (!attr) {                                                                                           \
      MSP_CHECK_CUDA(cudaFuncSetAttribute(narrow_wgrad_kernel<CC, KK, FF, FF>,                             \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, NwCfg<CC, KK, FF>::kSmem)); \
      attr = true;                                                                                         \
    }                                                                                                      \
    narrow_wgrad_kernel<CC, KK, FF, FF><<<grid, kNwThreads, NwCfg<CC, KK, FF>::kSmem, st>>>(p);            \
  } while (0)
#define MSP_NW_BY_FILTER(CC, KK)             \
  do {                                       \
    if (d->KH == 3) MSP_NW_LAUNCH(CC, KK, 3); \
    else MSP_NW_LAUNCH(CC, KK, 2);           \
  } while (0)
  if (d->C == 16 && d->K == 16) MSP_NW_BY_FILTER(16, 16);
  else if (d->C == 16 && d->K == 32) MSP_NW_BY_FILTER(16, 32);
  else if (d->C == 32 && d->K == 16) MSP_NW_BY_FILTER(32, 16);
  else MSP_NW_BY_FILTER(32, 32);
#undef MSP_NW_BY_FILTER
#undef MSP_NW_LAUNCH
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

namespace {
bool narrow_fprop_shape_ok(int C, int K, int KH, int KW, int stride, int win_px, long long pixels) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MSP_NARROW");
    enabled = e ? atoi(e) : 1;
  }
  return (enabled & 2) == 0 && enabled != 0 && win_px == 0 && stride == 1 && KH == KW && (KH == 3 || KH == 2) &&
         (C == 16 || C == 32) && (K == 16 || K == 32) && pixels >= 65536;
}

int narrow_fprop_launch(NarrowFpropParams& p, int C, int K, int KH, int stat_rows, cudaStream_t st) {
  p.tiles_w = msp_cdiv(p.OW, kNfTW);
  p.tiles_h = msp_cdiv(p.OH, kNfTH);
  p.total_tiles = p.N * p.tiles_w * p.tiles_h;
  int grid = msp_num_sms();
  if (grid > p.total_tiles) grid = p.total_tiles;
  if (p.ch_sum != nullptr && stat_rows > 0) {
    MSP_REQUIRE(stat_rows >= grid, "narrow_fprop: statistic workspace has %d rows for %d CTAs", stat_rows, grid);
  }
#define MSP_NF_LAUNCH(CC, KK, FF)                                                                          \
  do {                                                                                                     \
    static bool attr = false;                                                                              \
    if (!attr) {                                                                                           \
      MSP_CHECK_CUDA(cudaFuncSetAttribute(narrow_fprop_kernel<CC, KK, FF, FF>,                             \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, NfCfg<CC, KK, FF>::kSmem)); \
      attr = true;                                                                                         \
    }                                                                                                      \
    narrow_fprop_kernel<CC, KK, FF, FF><<<grid, kNfThreads, NfCfg<CC, KK, FF>::kSmem, st>>>(p);            \
  } while (0)
#define MSP_NF_BY_FILTER(CC, KK)              \
  do {                                        \
    if (KH == 3) MSP_NF_LAUNCH(CC, KK, 3);    \
    else MSP_NF_LAUNCH(CC, KK, 2);            \
  } while (0)
  if (C == 16 && K == 16) MSP_NF_BY_FILTER(16, 16);
  else if (C == 16 && K == 32) MSP_NF_BY_FILTER(16, 32);
  else if (C == 32 && K == 16) MSP_NF_BY_FILTER(32, 16);
  else MSP_NF_BY_FILTER(32, 32);
#undef MSP_NF_BY_FILTER
#undef MSP_NF_LAUNCH
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
}  // namespace

bool msp_narrow_fprop_ok(const msp_conv_desc* d) {
  return narrow_fprop_shape_ok(d->C, d->K, d->KH, d->KW, d->stride, d->win_px, (long long)d->N * d->Ho * d->Wo);
}
bool msp_narrow_dgrad_ok(const msp_conv_desc* d) {
  return narrow_fprop_shape_ok(d->K, d->C, d->KH, d->KW, d->stride, d->win_px, (long long)d->N * d->H * d->W) &&
         d->KH - 1 - d->pad_t >= 0 && d->KW - 1 - d->pad_l >= 0;
}

int msp_narrow_fprop(const msp_conv_desc* d, const void* x, const void* w_fprop, const float* bias, void* y, float* ch_sum,
                     float* ch_sqsum, void* stream) {
  MSP_REQUIRE(msp_narrow_fprop_ok(d), "narrow_fprop: layer not eligible");
  NarrowFpropParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)x; p.w = (const __nv_bfloat16*)w_fprop; p.y = (__nv_bfloat16*)y; p.bias = bias;
  p.ch_sum = ch_sum; p.ch_sqsum = ch_sqsum;
  p.stat_row = (ch_sum != nullptr && d->stat_rows > 0) ? 2ll * d->K : 0;
  p.N = d->N; p.IH = d->H; p.IW = d->W; p.OH = d->Ho; p.OW = d->Wo; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.x_cs = d->x_cs; p.y_cs = d->y_cs; p.relu = d->relu; p.flip = 0;
  return narrow_fprop_launch(p, d->C, d->K, d->KH, d->stat_rows, (cudaStream_t)stream);
}

// dx = conv_transpose(dy, w) of a stride-1 convolution: "fprop" over dy with the dgrad operand [C][taps][K], reversed taps
int msp_narrow_dgrad(const msp_conv_desc* d, const void* dy, const void* w_dgrad, const float* bias, int relu, void* dx,
                     void* stream) {
  MSP_REQUIRE(msp_narrow_dgrad_ok(d), "narrow_dgrad: layer not eligible");
  NarrowFpropParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const __nv_bfloat16*)dy; p.w = (const __nv_bfloat16*)w_dgrad; p.y = (__nv_bfloat16*)dx; p.bias = bias;
  p.N = d->N; p.IH = d->Ho; p.IW = d->Wo; p.OH = d->H; p.OW = d->W;
  p.pad_t = d->KH - 1 - d->pad_t; p.pad_l = d->KW - 1 - d->pad_l;
  p.x_cs = d->y_cs; p.y_cs = d->x_cs; p.relu = relu; p.flip = 1;
  return narrow_fprop_launch(p, d->K, d->C, d->KH, 0, (cudaStream_t)stream);
}
