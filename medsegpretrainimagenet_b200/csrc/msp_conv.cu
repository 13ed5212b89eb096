// Convolution forward / dgrad / wgrad as implicit GEMMs on the Blackwell tensor cores.
//
// Design:
//  * "tap-GEMM": a convolution is a sum over filter taps of [pixels x C] * [C x K] GEMMs.  For one
//    output tile (a bw x bh x bn box of output pixels, <= 128 rows) and one tap, the A operand is
//    the same box of the NHWC input shifted by the tap offset — a single 4-D TMA tile load with
//    hardware zero fill for the padding halo and `elementStrides` for stride-2 convs.  The B operand
//    is a [BN x 64] slab of the packed weights [K][tap][C] (3-D TMA).  Both land in shared memory
//    in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
//  * PERSISTENT, warp-specialised CTA, one per SM: warp 0 = TMA producer running up to kStages
//    k-blocks (and therefore several tiles) ahead, warp 1 = MMA issuer (tcgen05.mma cta_group::1, M=128,
//    N<=256, fp32 accumulators DOUBLE-BUFFERED in TMEM).  Both roles are taken with elect.sync — under
//    `if (lane == 0)` ptxas wraps every UTCHMMA / UTMALDG in a per-lane loop and one MMA costs ~120-170
//    cycles of issue time whatever N is (tools/mma_rate.cu, profiles/r01_mma_issue_rate.txt).  Warps 2-9 =
//    epilogue, DECOUPLED: warp (q, part) owns TMEM lanes [32q, 32q+32) and a fixed set of 32-column steps of
//    every tile: tcgen05.ld (one step ahead) -> bias/ReLU -> bf16 -> private swizzled slab -> coalesced
//    16-byte global stores (dgrad-on-top-of-a-gradient adds the old values here), overlapped with the next
//    tile's main loop; no CTA-wide barrier on the per-tile path.  The BatchNorm batch statistics (per-channel
//    sum / sum of squares of the bf16-rounded outputs) are read back column-wise from the slab and kept in
//    REGISTERS across all the tiles a CTA owns (tiles are walked output-channel-block-major), so a launch
//    issues ~2 x Cout x (#CTAs) atomics instead of 2 x Cout x (#tiles x 4).
//  * dgrad = the same kernel on dy with transposed weights; stride-2 dgrad is decomposed into the
//    four output-parity classes, each a small stride-1 tap-GEMM writing a strided sub-grid of dx.
//  * wgrad contracts over pixels: A = dy tile (128 output channels), B = shifted x tile (up to 256 input
//    channels), both consumed MN-major straight from the same TMA boxes; split-K over pixel tiles, each
//    split writing its own fp32 partial (no atomics, deterministic), summed by msp_unpack_wgrad.
//  * "row-window" mode for the tiny-channel first convolution (7x7/2 stem on 1-3 channels, 3x3 on 3
//    channels): the input is stored zero-padded in W with 8 (or 16) channels per pixel, so the KW taps
//    of one filter row are 64 CONTIGUOUS elements; an overlapping-stride tensor map (dim 1 = output
//    column, stride = conv stride x pixel) turns each filter ROW into one 64-deep k-block: 7 TMA
//    loads / 28 MMAs per tile for the 7x7 stem instead of 49 / 49.
//
// Replaces cuDNN conv fwd / bwd-data / bwd-filter behind nn.Conv2d
// (reference classification/models.py:43-46,161-179,234-253; segmentation/models/blocks.py:458,518,590).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"
#include <stdlib.h>

extern void msp_count_launch(int n);

namespace {

// name of the kernel variant the calling thread launched last (bench.py's per-kernel roofline list)
thread_local const char* g_last_kernel = "";

constexpr int kMaxTaps = 49;
constexpr int kBM = 128;  // UMMA M (rows of the output tile)
constexpr int kBK = 64;   // contraction elements per pipeline stage (= one 128-byte swizzle row)
constexpr int kConvThreads = 192;  // wgrad: producer + MMA + 4 epilogue warps
constexpr int kEpiWarps = 8;       // tap-GEMM epilogue warps (2 per TMEM lane quarter / scheduler; 16 measured slower: barriers, registers)
constexpr int kTapThreads = 64 + 32 * kEpiWarps;  // producer + MMA + epilogue
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kEpiParts = kEpiWarps / 4;           // warps sharing one TMEM lane quarter
constexpr int kEpiSlabs = kEpiThreads / 32;        // epilogue warps (shared-memory budget of their slabs / scratch)
constexpr int kEpiBarrier = 1;  // named barrier id of the epilogue warps

struct TapGemmParams {
  int bw, bh, bn, rows;
  int bw_valid;  // output columns per tile (= bw, except in halo mode where bw is the halo pitch)
  int tiles_w, tiles_h, tiles_n;
  int tiles_m, tiles_co, total_tiles;
  int OWs, OHs, N;  // output sub-grid extent
  int sxw, sxh;     // A coordinate multipliers (conv stride for fprop, 1 for dgrad / row-window W)
  int C;            // contraction channels per tap
  int ntaps;
  int Kout;         // valid output channels
  int relu;
  int accumulate;  // the output tile is ADDED to memory (dgrad on top of a residual gradient)
  int debug;       // MSP_CONV_DEBUG bit mask (profiling experiments only; 0 in production)
  // halo mode (tapgemm_halo_kernel): the A tile is the whole input halo of the output tile, loaded once per
  // 64-channel chunk; tap (r', q') reads it from row r'*bw + q' on (tap_dh / tap_dw are then halo-relative)
  int halo, org_dh, org_dw, halo_box_h, a_stages, b_stages, b_resident;
  int dbg_stages;  // profiling experiments: ring depth override for the single-CTA tap kernel (0 = kStages)
  long long y_off, y_n_stride, y_h_stride, y_w_stride;  // element strides of the output sub-grid
  __nv_bfloat16* y;
  const float* bias;
  float* ch_sum;
  float* ch_sqsum;
  long long stat_row;  // deterministic statistics: elements between the rows of the per-CTA workspace (0: float atomics)
  int8_t tap_dh[kMaxTaps];
  int8_t tap_dw[kMaxTaps];
  uint8_t tap_w[kMaxTaps];
};

template <int BN_>
struct TapGemmCfg {
  static constexpr int kBTileBytes = BN_ * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = BN_ >= 256 ? 4 : (BN_ >= 128 ? 5 : 6);
  static constexpr int kStageBufs = BN_ >= 256 ? 1 : 2;  // 16 KB units of epilogue staging (the warps' slabs, EpiCfg)
  static constexpr int kScratchBytes = kEpiSlabs * 128 * 4 + 4 * 128 * 4;  // + 6 KB (BN = 256 needs part of it)
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStageBufs * kATileBytes + kScratchBytes + 1024;  // + alignment slack
  static constexpr int kTmemCols = 2 * BN_ < 32 ? 32 : 2 * BN_;
};

// Epilogue warps of the tap-GEMM kernels (warps 2..9): drain TMEM accumulators tile by tile.
// Work items are walked as it0, it0 + it_step, ... < it_total; item -> (output-channel block tco = item / m_per_co,
// m index pm = item % m_per_co).  pair_rank < 0: one M tile per item (tm = pm).  pair_rank = 0/1: CTA pair
// (cta_group::2), the item is two M tiles and this CTA owns tm = 2*pm + pair_rank (possibly past the end: masked);
// the accumulator is handed back through the LEADER CTA's tempty barrier.
//
// The warps are DECOUPLED: warp (q, part) owns TMEM lanes [32q, 32q+32) and a fixed set of column steps of every
// tile, with a private staging slab — no CTA-wide barrier on the per-tile path (the barrier-per-chunk version spent
// 2.8x the TMEM-read time per tile, profiles/r01_epilogue_phases.txt; the 1x1 convolutions are epilogue bound).
// Per step (kCW columns): tcgen05.ld -> bias/ReLU -> bf16 -> swizzled slab (lane = row) -> __syncwarp -> the slab is
// read back (a) piece-major for global stores that cover whole 16*P-byte row segments and (b) column-major for the
// BatchNorm statistics, kept in registers per lane (its column pair) and flushed to global memory (2*Cout atomics
// per CTA) only when the CTA leaves an output-channel block.
template <int BN_>
struct EpiCfg {
  static constexpr int kCWraw = BN_ / kEpiParts;
  static constexpr int kCW = kEpiParts == 2 ? (BN_ >= 128 ? 32 : BN_ / 2)
                                            : (kCWraw > 32 ? 32 : (kCWraw < 8 ? 8 : kCWraw));  // columns per step
  static constexpr int kP = kCW / 8;                                          // 16-byte pieces per slab row
  static constexpr int kActiveParts = (BN_ / kCW) < kEpiParts ? (BN_ / kCW) : kEpiParts;
  static constexpr int kSteps = BN_ / (kActiveParts * kCW);                   // steps per (active) warp and tile
  static constexpr int kSlabBytes = 32 * kCW * 2 < 512 * kSteps ? 512 * kSteps : 32 * kCW * 2;
  static constexpr int kBytes = kEpiWarps * kSlabBytes;
  // first column (within the tile) of step s for column part `part`
  __device__ static constexpr int col(int s, int part) { return (part * kSteps + s) * kCW; }
};

template <int BN_>
__device__ __forceinline__ void tap_epilogue(const TapGemmParams& p, uint8_t* staging, float* /*scratch*/,
                                             uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                             int warp, int lane, int it0, int it_step, int it_total,
                                             int m_per_co, int pair_rank) {
  using E = EpiCfg<BN_>;
  constexpr int CW = E::kCW, P = E::kP;
  static_assert(E::kBytes <= TapGemmCfg<BN_>::kStageBufs * kATileBytes + TapGemmCfg<BN_>::kScratchBytes, "epilogue smem");
  const int et = threadIdx.x - 64;  // 0..kEpiThreads-1
  const int ew = et >> 5;           // epilogue warp 0..7
  const int q = warp & 3;           // TMEM lane quarter this warp may access
  const int part = ew >> 2;         // which column part of the tile this warp handles
  const bool active = part < E::kActiveParts;  // very narrow tiles: the other warps only take part in the hand-offs
  const int row = q * 32 + lane;    // this thread's accumulator row
  const bool do_stats = p.ch_sum != nullptr && !(p.debug & 2);
  const bool has_bias = p.bias != nullptr;
  const uint32_t slab_s = smem_u32(staging) + ew * E::kSlabBytes;
  float st_acc[E::kSteps][4];  // running (sum a, sum b, sumsq a, sumsq b) of this lane's column pair, per step
#pragma unroll
  for (int i = 0; i < E::kSteps; ++i) st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
  // Tile-invariant per-thread state: its TMEM row, the row pieces it copies out, their slab and global offsets.
  const int my_wi = row % p.bw, my_hi = (row / p.bw) % p.bh, my_ni = row / (p.bw * p.bh);
  int o_wi[P], o_hi[P], o_ni[P], o_ch[P];
  uint32_t o_lds[P];       // slab offset of (row, piece)
  long long o_rel[P];      // element offset of the piece relative to the tile origin
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int u = j * 32 + lane, rl = u / P, g = u % P, r = q * 32 + rl;
    o_wi[j] = r % p.bw;
    o_hi[j] = (r / p.bw) % p.bh;
    o_ni[j] = r / (p.bw * p.bh);
    o_ch[j] = g * 8;
    o_lds[j] = (uint32_t)((u >> 3) * 128 + (((u & 7) ^ ((u >> 3) & 7)) << 4));
    o_rel[j] = (r < p.rows && o_wi[j] < p.bw_valid) ? (long long)o_ni[j] * p.y_n_stride + (long long)o_hi[j] * p.y_h_stride +
                                                      (long long)o_wi[j] * p.y_w_stride + g * 8
                                                    : -1;
  }
  uint32_t sts_off[P];  // swizzled slab offsets of this thread's (lane = row) 16-byte pieces
#pragma unroll
  for (int g = 0; g < P; ++g) {
    const int u = lane * P + g;
    sts_off[g] = (uint32_t)((u >> 3) * 128 + (((u & 7) ^ ((u >> 3) & 7)) << 4));
  }
  // statistics read: pass j, lane -> 32-bit word w = j*32 + lane of the slab (row w / (4P), column pair w % (4P))
  const int st_cp = lane % (4 * P);
  const uint32_t st_off0 = (uint32_t)((((lane >> 2) & 7) << 4) + (lane & 3) * 4);  // line j: pos16 = lane/4, xor (j & 7)
  const bool flat_tiles = p.tiles_h == 1 && p.tiles_n == 1;  // 1x1 "flat" convolutions: tile = 128 pixels of one row
  uint32_t t = 0;
  int tco = it0 / m_per_co, pm = it0 - tco * m_per_co;
  for (int item = it0; item < it_total; item += it_step, ++t) {
    int tm = pair_rank < 0 ? pm : 2 * pm + pair_rank;
    const bool tile_ok = tm < p.tiles_m;  // odd tile counts: the pair's second CTA idles on a masked duplicate
    if (!tile_ok) tm = p.tiles_m - 1;
    int tw = tm, th = 0, tn = 0;
    if (!flat_tiles) {
      tw = (int)((uint32_t)tm % (uint32_t)p.tiles_w);
      const uint32_t rest = (uint32_t)tm / (uint32_t)p.tiles_w;
      th = (int)(rest % (uint32_t)p.tiles_h);
      tn = (int)(rest / (uint32_t)p.tiles_h);
    }
    const int w0 = tw * p.bw_valid, h0 = th * p.bh, n0 = tn * p.bn;
    const int co0 = tco * BN_;
    const bool valid =
        tile_ok && row < p.rows && my_wi < p.bw_valid && (w0 + my_wi) < p.OWs && (h0 + my_hi) < p.OHs && (n0 + my_ni) < p.N;
    const long long tile_off = p.y_off + (long long)n0 * p.y_n_stride + (long long)h0 * p.y_h_stride +
                               (long long)w0 * p.y_w_stride + co0;
    __nv_bfloat16* o_ptr[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const bool ok = tile_ok && o_rel[j] >= 0 && (w0 + o_wi[j]) < p.OWs && (h0 + o_hi[j]) < p.OHs &&
                      (n0 + o_ni[j]) < p.N;
      o_ptr[j] = ok ? p.y + tile_off + o_rel[j] : nullptr;
    }
    // advance (tco, pm) to this CTA's next item without a division
    int tco_next = tco, pm_next = pm + it_step;
    while (pm_next >= m_per_co) {
      pm_next -= m_per_co;
      ++tco_next;
    }
    const uint32_t as = t & 1u;
    if (lane == 0) mbar_wait(&tfull_bar[as], (t >> 1) & 1u);  // one poller per warp
    __syncwarp();
    tc_fence_after();
    const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN_;
    // TMEM loads run one step ahead of the arithmetic: step s+1 is in flight while step s is packed / stored / summed
    uint32_t v[CW];
    auto tm_issue = [&](int s_) {
      const int c_ = E::col(s_, part);
      if constexpr (CW == 64) {
        tmem_ld_32x32(tmem_row + c_, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld_32x32(tmem_row + c_ + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      } else if constexpr (CW == 32) {
        tmem_ld_32x32(tmem_row + c_, v);
      } else if constexpr (CW == 16) {
        tmem_ld_32x16(tmem_row + c_, v);
      } else {
        tmem_ld_32x8(tmem_row + c_, v);
      }
    };
    if (active) tm_issue(0);
#pragma unroll
    for (int s = 0; s < E::kSteps; ++s) {
      const int cl = E::col(s, part);  // first column of this step within the tile
      const int cb = co0 + cl;         // first output channel
      // dgrad on top of an existing gradient: fetch the old values now, consume them after the transpose
      uint4 old[P];
      if (p.accumulate) {
#pragma unroll
        for (int j = 0; j < P; ++j)
          old[j] = (active && o_ptr[j] != nullptr && cb + o_ch[j] < p.Kout)
                       ? __ldg(reinterpret_cast<const uint4*>(o_ptr[j] + cl))
                       : make_uint4(0u, 0u, 0u, 0u);
      }
      tmem_ld_wait();
      float x[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) x[j] = __uint_as_float(v[j]);
      if (s + 1 < E::kSteps) {
        if (active) tm_issue(s + 1);
      } else {  // accumulator fully read by this warp: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair_rank < 0) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_leader(&tempty_bar[as]);
        }
      }
      if (cb < p.Kout && active) {
        // uniform branches around straight-line blocks (a per-element `if` costs a taken branch each)
        if (has_bias) {
          if (cb + CW <= p.Kout && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + cb);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 bb = __ldg(b4 + j);
              x[4 * j] += bb.x;
              x[4 * j + 1] += bb.y;
              x[4 * j + 2] += bb.z;
              x[4 * j + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (cb + j < p.Kout) x[j] += __ldg(p.bias + cb + j);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < CW; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < CW; ++j) x[j] = 0.f;
        }
        __syncwarp();  // the previous step's slab reads are done
#pragma unroll
        for (int g = 0; g < P; ++g)
          sts_v4(slab_s + sts_off[g], pack_bf16x2(x[8 * g], x[8 * g + 1]), pack_bf16x2(x[8 * g + 2], x[8 * g + 3]),
                 pack_bf16x2(x[8 * g + 4], x[8 * g + 5]), pack_bf16x2(x[8 * g + 6], x[8 * g + 7]));
        __syncwarp();
        // (running the two slab consumers in opposite order on the two column parts measured slower)
        auto copy_out = [&]() {
          if (!(p.debug & 1)) {
            // copy-out: lane -> (row, 16-byte piece); 32 lanes cover 32/P whole row segments of 16*P bytes
            uint4 o[P];
#pragma unroll
            for (int j = 0; j < P; ++j) o[j] = lds_v4(slab_s + o_lds[j]);
            if (p.accumulate) {  // dgrad on top of an existing gradient
#pragma unroll
              for (int j = 0; j < P; ++j) {
                const uint32_t ov[4] = {old[j].x, old[j].y, old[j].z, old[j].w};
                const uint32_t nv[4] = {o[j].x, o[j].y, o[j].z, o[j].w};
                uint32_t rv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 a = unpack_bf16x2(ov[e]), b2 = unpack_bf16x2(nv[e]);
                  rv[e] = pack_bf16x2(a.x + b2.x, a.y + b2.y);
                }
                o[j] = make_uint4(rv[0], rv[1], rv[2], rv[3]);
              }
            }
#pragma unroll
            for (int j = 0; j < P; ++j)
              if (o_ptr[j] != nullptr && cb + o_ch[j] < p.Kout) st_v4(o_ptr[j] + cl, o[j]);
          }
        };
        auto col_stats = [&]() {
          if (do_stats) {
            // column sums of the bf16 slab: lane -> column pair st_cp; 4P conflict-free LDS.32 (pass j = 128-byte line j)
            float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
            for (int j = 0; j < 4 * P; ++j) {
              const float2 f = unpack_bf16x2(lds_u32(slab_s + j * 128 + (st_off0 ^ (uint32_t)((j & 7) << 4))));
              s1a += f.x;
              s2a = fmaf(f.x, f.x, s2a);
              s1b += f.y;
              s2b = fmaf(f.y, f.y, s2b);
            }
            st_acc[s][0] += s1a;
            st_acc[s][1] += s1b;
            st_acc[s][2] += s2a;
            st_acc[s][3] += s2b;
          }
        };
        copy_out();
        col_stats();
      }
    }
    // flush the CTA's statistics when it leaves the output-channel block (or finishes): the warps park their
    // register sums in their (now idle) slabs, then 2*BN threads add the four lane quarters and issue one global
    // atomic per (statistic, channel).  (Shared-memory float atomics compile to a CAS loop: not used.)
    if (do_stats && (item + it_step >= it_total || tco_next != tco)) {
#pragma unroll
      for (int i = 0; i < E::kSteps; ++i) {
#pragma unroll
        for (int o = 16; o >= 4 * P; o >>= 1) {  // lanes that share a column pair (narrow steps)
#pragma unroll
          for (int e = 0; e < 4; ++e) st_acc[i][e] += __shfl_xor_sync(0xffffffffu, st_acc[i][e], o);
        }
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < E::kSteps; ++i) {
        sts_v4(slab_s + (uint32_t)(i * 32 + lane) * 16, __float_as_uint(st_acc[i][0]), __float_as_uint(st_acc[i][1]),
               __float_as_uint(st_acc[i][2]), __float_as_uint(st_acc[i][3]));
        st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
      }
      named_bar_sync(kEpiBarrier, kEpiThreads);
      for (int i = et; i < 2 * BN_; i += kEpiThreads) {
        const int which = i / BN_, cl = i - which * BN_, col = co0 + cl;
        const int gs = cl / CW, prt = gs / E::kSteps, st = gs - prt * E::kSteps, cw = cl - gs * CW;
        if (col < p.Kout) {
          const uint32_t a = smem_u32(staging) + (uint32_t)(prt * 4) * E::kSlabBytes +
                             (uint32_t)(st * 32 + (cw >> 1)) * 16 + (uint32_t)(which * 2 + (cw & 1)) * 4;
          const float tot = lds_f32(a) + lds_f32(a + E::kSlabBytes) + lds_f32(a + 2 * E::kSlabBytes) +
                            lds_f32(a + 3 * E::kSlabBytes);
          float* dst = (which ? p.ch_sqsum : p.ch_sum) + col;
          // deterministic mode: this CTA's own row of the workspace (a CTA leaves every channel block once), added up
          // in fixed order by msp_bn_finalize; otherwise one float atomic per CTA, statistic and channel
          if (p.stat_row) dst[(long long)blockIdx.x * p.stat_row] = tot;
          else atomicAdd(dst, tot);
        }
      }
      named_bar_sync(kEpiBarrier, kEpiThreads);
    }
    tco = tco_next;
    pm = pm_next;
  }
}

template <int BN_>
__global__ void __launch_bounds__(kTapThreads, 1)
tapgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TapGemmParams p) {
  using Cfg = TapGemmCfg<BN_>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  float* scratch = reinterpret_cast<float*>(staging + Cfg::kStageBufs * kATileBytes);
  __shared__ uint64_t full_bar[Cfg::kStages];
  __shared__ uint64_t empty_bar[Cfg::kStages];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int chunks = (p.C + kBK - 1) / kBK;
  const int kiters = p.ntaps * chunks;
  (void)kiters;
  long long dbg_c0 = 0;
  unsigned long long dbg_t0 = 0;
  if ((p.debug & 512) && blockIdx.x == 0 && threadIdx.x == 0) {  // profiling experiment: effective SM clock
    dbg_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiThreads / 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // everything above (barriers, TMEM, descriptors) may overlap the previous kernel's tail

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)Cfg::kBTileBytes;
      // ring slot / phase advance incrementally: `it % nst`, `it / nst` with a run-time ring depth were two integer
      // divisions per k-block in this one thread (~200 cycles against the 256 cycles of four N <= 128 MMAs)
      const uint32_t nst = p.dbg_stages ? (uint32_t)p.dbg_stages : (uint32_t)Cfg::kStages;
      uint32_t ps = 0, pph = 0;
      const bool flat_tiles = p.tiles_h == 1 && p.tiles_n == 1;  // 1x1 "flat" convolutions: no (w, h, n) split
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tco = p.tiles_co == 1 ? 0 : tile / p.tiles_m;
        const int tm = tile - tco * p.tiles_m;
        int tw = tm, th = 0, tn = 0;
        if (!flat_tiles) {
          tw = tm % p.tiles_w;
          th = (tm / p.tiles_w) % p.tiles_h;
          tn = tm / (p.tiles_w * p.tiles_h);
        }
        const int w0 = tw * p.bw * p.sxw, h0 = th * p.bh * p.sxh, n0 = tn * p.bn;
        const int co0 = tco * BN_;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int cw = w0 + p.tap_dw[tap], chh = h0 + p.tap_dh[tap], wt = (int)p.tap_w[tap];
          for (int ch = 0; ch < chunks; ++ch) {
            const uint32_t s = ps, ph = pph;
            if (++ps == nst) {
              ps = 0;
              pph ^= 1u;
            }
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* a_s = smem + s * Cfg::kStageBytes;
            if (p.debug & 128) {  // timing experiment: MMA side alone (no loads, garbage operands)
              mbar_arrive(&full_bar[s]);
              continue;
            }
            mbar_expect_tx(&full_bar[s], tx_bytes);
            tma_load_4d(a_s, &tmA, &full_bar[s], ch * kBK, cw, chh, n0);
            tma_load_3d(a_s + kATileBytes, &tmB, &full_bar[s], ch * kBK, wt, co0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // One thread: keep its instruction stream short (descriptor halves, incremental stage / phase counters).
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN_, 0, 0);
      const uint64_t d0 = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;
      constexpr uint32_t kStageLo = Cfg::kStageBytes >> 4, kBLo = kATileBytes >> 4;
      const int c_tail = p.C - (chunks - 1) * kBK;       // channels of the last chunk
      const int nk_tail = (c_tail + 15) >> 4;
      uint32_t s = 0, ph = 0, t = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
        const uint32_t as = t & 1u;
        mbar_wait(&tempty_bar[as], ((t >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN_;
        uint32_t acc = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + s * kStageLo, b_lo = a_lo + kBLo;
            if (p.debug & 256) {  // timing experiment: load side alone
            } else if (ch + 1 < chunks || nk_tail == 4) {
              umma_bf16_lohi(tmem_d, a_lo, desc_hi, b_lo, desc_hi, idesc, acc);
              umma_bf16_lohi(tmem_d, a_lo + 2, desc_hi, b_lo + 2, desc_hi, idesc, 1u);
              umma_bf16_lohi(tmem_d, a_lo + 4, desc_hi, b_lo + 4, desc_hi, idesc, 1u);
              umma_bf16_lohi(tmem_d, a_lo + 6, desc_hi, b_lo + 6, desc_hi, idesc, 1u);
            } else {
              for (int k = 0; k < nk_tail; ++k)
                umma_bf16_lohi(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, acc | (uint32_t)k);
            }
            acc = 1u;
            umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
            if (++s == (p.dbg_stages ? (uint32_t)p.dbg_stages : (uint32_t)Cfg::kStages)) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    tap_epilogue<BN_>(p, staging, scratch, tmem_base, tfull_bar, tempty_bar, warp, lane, (int)blockIdx.x,
                      (int)gridDim.x, p.total_tiles, p.tiles_m, -1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  if ((p.debug & 512) && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    const long long c1 = clock64();
    printf("msp clock probe: %lld cycles in %llu ns -> %.0f MHz\n", c1 - dbg_c0, t1 - dbg_t0,
           1e3 * (double)(c1 - dbg_c0) / (double)(t1 - dbg_t0));
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two SMs of a TPC form a cluster and share one 256-row x BN tile.
// Each CTA loads ITS 128-row A tile and HALF of the B tile (BN/2 weight rows); one thread of the leader CTA issues
// M=256 MMAs that read both CTAs' shared memory and write each CTA's own TMEM; tcgen05.commit multicasts the
// "slot free" / "accumulator ready" arrivals to both CTAs, both epilogues drain their 128 rows and hand the
// accumulator back through the leader's barrier.  Per SM and k-block that is 16 KB + BN*64 B of operands instead
// of 16 KB + BN*128 B, and twice the MMA issue width (a single-CTA MMA measured ~45 % of the tensor peak here).
// ------------------------------------------------------------------------------------------------
template <int BN_>
struct TapGemm2Cfg {
  static constexpr int kBHalfBytes = (BN_ / 2) * kBK * 2;
  static constexpr int kStageBytes = kATileBytes + kBHalfBytes;
  static constexpr int kStages = BN_ >= 256 ? 5 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + TapGemmCfg<BN_>::kStageBufs * kATileBytes +
                                    TapGemmCfg<BN_>::kScratchBytes + 1024;
};

template <int BN_>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTapThreads, 1)
tapgemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const TapGemmParams p) {
  using Cfg = TapGemmCfg<BN_>;
  using Cfg2 = TapGemm2Cfg<BN_>;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg2::kStages * Cfg2::kStageBytes;
  float* scratch = reinterpret_cast<float*>(staging + Cfg::kStageBufs * kATileBytes);
  __shared__ uint64_t full_bar[Cfg2::kStages];
  __shared__ uint64_t empty_bar[Cfg2::kStages];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int chunks = (p.C + kBK - 1) / kBK;
  const int npm = (p.tiles_m + 1) >> 1;            // M-tile pairs per output-channel block
  const int items = npm * p.tiles_co;
  const int it0 = (int)(blockIdx.x >> 1), it_step = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg2::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * (kEpiThreads / 32));  // the epilogue warps of BOTH CTAs
    }
    mbar_fence_init();
  }
  cluster_sync_all();  // the peer's barriers exist before anything remote touches them
  if (warp == 1) tmem_alloc2<Cfg::kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // everything above (barriers, TMEM, descriptors) may overlap the previous kernel's tail

  if (warp == 0) {
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t tx_bytes = 2u * ((uint32_t)p.rows * 128u + (uint32_t)Cfg2::kBHalfBytes);  // both CTAs
      uint32_t s = 0, ph = 0;
      int tco = it0 / npm, pm = it0 - tco * npm;
      for (int item = it0; item < items; item += it_step) {
        int tm = 2 * pm + rank;
        if (tm >= p.tiles_m) tm = p.tiles_m - 1;
        const int tw = tm % p.tiles_w;
        const int th = (tm / p.tiles_w) % p.tiles_h;
        const int tn = tm / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw * p.sxw, h0 = th * p.bh * p.sxh, n0 = tn * p.bn;
        const int co0 = tco * BN_ + rank * (BN_ / 2);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int cw = w0 + p.tap_dw[tap], chh = h0 + p.tap_dh[tap], wt = (int)p.tap_w[tap];
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* a_s = smem + s * Cfg2::kStageBytes;
            if (p.debug & 128) {  // timing experiment: MMA side alone
              if (rank == 0) mbar_arrive(&full_bar[s]);
            } else {
              if (rank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
              tma_load_4d_2sm(a_s, &tmA, &full_bar[s], ch * kBK, cw, chh, n0);
              tma_load_3d_2sm(a_s + kATileBytes, &tmB, &full_bar[s], ch * kBK, wt, co0);
            }
            if (++s == (uint32_t)Cfg2::kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        pm += it_step;
        while (pm >= npm) {
          pm -= npm;
          ++tco;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN_, 0, 0);
      const uint64_t d0 = umma_smem_desc_sw128(smem_u32(smem), 16, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;
      constexpr uint32_t kStageLo = Cfg2::kStageBytes >> 4, kBLo = kATileBytes >> 4;
      const int c_tail = p.C - (chunks - 1) * kBK;
      const int nk_tail = (c_tail + 15) >> 4;
      uint32_t s = 0, ph = 0, t = 0;
      for (int item = it0; item < items; item += it_step, ++t) {
        const uint32_t as = t & 1u;
        mbar_wait(&tempty_bar[as], ((t >> 1) & 1u) ^ 1u);  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN_;
        uint32_t acc = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int ch = 0; ch < chunks; ++ch) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + s * kStageLo, b_lo = a_lo + kBLo;
            const int nk = (ch + 1 < chunks) ? 4 : nk_tail;
            for (int k = 0; k < nk; ++k)
              umma_bf16_lohi_2sm(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, acc | (uint32_t)k);
            acc = 1u;
            umma_commit_mc(&empty_bar[s], 3u);  // frees the slot in BOTH CTAs once these MMAs have read it
            if (++s == (uint32_t)Cfg2::kStages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
        umma_commit_mc(&tfull_bar[as], 3u);
      }
    }
    __syncwarp();
  } else {
    tap_epilogue<BN_>(p, staging, scratch, tmem_base, tfull_bar, tempty_bar, warp, lane, it0, it_step, items, npm,
                      rank);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA of the pair frees TMEM / exits while the other may still read its shared memory
  if (warp == 1) tmem_dealloc2<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Narrow-channel epilogue of the halo kernel (TN = 16 or 32 output channels per tile; the full-resolution decoder layers
// of the U-Nets: 16 / 32 channels at 256 x 256, HBM bound).  With one 128 x TN accumulator per hand-off the kernel ran at
// 6-20 % of its HBM roofline (profiles/r01_convbench_unet50_b24.txt: 205 us for 15 us of traffic): nine N = 16 MMAs per
// tile take ~600 cycles, the per-tile epilogue chain (barrier wait -> tcgen05.ld -> pack -> slab -> stores -> statistics)
// ~2400 cycles of pure latency whatever the tile's width.  Here G = 64 / TN pixel tiles share one ACCUMULATOR SET of
// 64 TMEM columns (two sets, double buffered): the MMA warp fills the set tile by tile and commits once, the epilogue
// drains all G tiles with ONE pass of the chain — columns [g*TN, (g+1)*TN) of the set belong to tile g, so every 16-byte
// piece of a row carries its own tile origin.  Same slab layout and statistics read-back as tap_epilogue<64>.
// ------------------------------------------------------------------------------------------------
constexpr int kMultiW = 64;  // accumulator set width (TMEM columns)

template <int TN>
__device__ __forceinline__ void halo_multi_epilogue(const TapGemmParams& p, uint8_t* staging, uint32_t tmem_base,
                                                    uint64_t* tfull_bar, uint64_t* tempty_bar, int warp, int lane) {
  constexpr int G = kMultiW / TN;   // pixel tiles per accumulator set
  constexpr int CW = 32, P = 4;     // accumulator columns per warp, 16-byte pieces per slab row
  constexpr int TPW = CW / TN;      // tiles spanned by one warp's columns: 2 (TN = 16) or 1 (TN = 32)
  constexpr int PPT = TN / 8;       // 16-byte pieces per tile and row
  constexpr int kSlabBytes = 32 * CW * 2;
  static_assert(TN == 16 || TN == 32, "narrow tiles only");
  static_assert(kEpiWarps == 8 && kEpiWarps * kSlabBytes <= TapGemmCfg<TN>::kStageBufs * kATileBytes, "slab budget");
  const int et = threadIdx.x - 64;
  const int ew = et >> 5;
  const int q = warp & 3;        // TMEM lane quarter of this warp
  const int part = ew >> 2;      // column half of the set
  const int row = q * 32 + lane;
  const bool do_stats = p.ch_sum != nullptr && !(p.debug & 2);
  const bool has_bias = p.bias != nullptr;
  const uint32_t slab_s = smem_u32(staging) + ew * kSlabBytes;
  const int cl0 = part * CW;     // first accumulator column of this warp within the set
  const int g0 = cl0 / TN;       // first tile (within the set) its columns belong to
  float st_acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int my_wi = row % p.bw, my_hi = row / p.bw;  // halo tiles hold one image: row -> (hi, wi) with the halo pitch
  const bool my_row_ok = row < p.rows && my_wi < p.bw_valid;
  int o_wi[P], o_hi[P], o_tile[P], o_ch[P];
  uint32_t o_lds[P];
  long long o_rel[P];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int u = j * 32 + lane, rl = u / P, gp = u % P, r = q * 32 + rl;
    o_wi[j] = r % p.bw;
    o_hi[j] = r / p.bw;
    o_tile[j] = gp / PPT;
    o_ch[j] = (gp % PPT) * 8;
    o_lds[j] = (uint32_t)((u >> 3) * 128 + (((u & 7) ^ ((u >> 3) & 7)) << 4));
    o_rel[j] = (r < p.rows && o_wi[j] < p.bw_valid)
                   ? (long long)o_hi[j] * p.y_h_stride + (long long)o_wi[j] * p.y_w_stride + o_ch[j]
                   : -1;
  }
  uint32_t sts_off[P];
#pragma unroll
  for (int g = 0; g < P; ++g) {
    const int u = lane * P + g;
    sts_off[g] = (uint32_t)((u >> 3) * 128 + (((u & 7) ^ ((u >> 3) & 7)) << 4));
  }
  const uint32_t st_off0 = (uint32_t)((((lane >> 2) & 7) << 4) + (lane & 3) * 4);
  const int my_tiles = (int)blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nsets = (my_tiles + G - 1) / G;
  for (int u = 0; u < nsets; ++u) {
    // the (up to) TPW tiles this warp's columns belong to
    int w0[TPW], h0[TPW];
    long long tile_off[TPW];
    bool t_ok[TPW], valid[TPW];
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
      const int tl = u * G + g0 + k;  // CTA-local tile index
      t_ok[k] = tl < my_tiles;
      const int tm = (int)blockIdx.x + (t_ok[k] ? tl : 0) * (int)gridDim.x;  // one output-channel block: tile = tm
      const int tw = (int)((uint32_t)tm % (uint32_t)p.tiles_w);
      const uint32_t rest = (uint32_t)tm / (uint32_t)p.tiles_w;
      const int th = (int)(rest % (uint32_t)p.tiles_h), tn = (int)(rest / (uint32_t)p.tiles_h);
      w0[k] = tw * p.bw_valid;
      h0[k] = th * p.bh;
      tile_off[k] = p.y_off + (long long)tn * p.y_n_stride + (long long)h0[k] * p.y_h_stride +
                    (long long)w0[k] * p.y_w_stride;
      valid[k] = t_ok[k] && my_row_ok && (w0[k] + my_wi) < p.OWs && (h0[k] + my_hi) < p.OHs;
    }
    __nv_bfloat16* o_ptr[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = o_tile[j];
      const bool ok = t_ok[k] && o_rel[j] >= 0 && (w0[k] + o_wi[j]) < p.OWs && (h0[k] + o_hi[j]) < p.OHs &&
                      o_ch[j] < p.Kout;
      o_ptr[j] = ok ? p.y + tile_off[k] + o_rel[j] : nullptr;
    }
    const uint32_t as = (uint32_t)u & 1u;
    if (lane == 0) mbar_wait(&tfull_bar[as], ((uint32_t)u >> 1) & 1u);
    __syncwarp();
    tc_fence_after();
    if (p.debug & 4) {  // timing experiment: hand-offs only (the accumulator is not even read)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      continue;
    }
    uint32_t v[CW];
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * kMultiW + cl0, v);
    uint4 old[P];
    if (p.accumulate) {
#pragma unroll
      for (int j = 0; j < P; ++j)
        old[j] = o_ptr[j] != nullptr ? __ldg(reinterpret_cast<const uint4*>(o_ptr[j])) : make_uint4(0u, 0u, 0u, 0u);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[as]);  // the set is in registers: hand it back to the MMA warp
    float x[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) x[j] = __uint_as_float(v[j]);
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if ((j % TN) < p.Kout) x[j] += __ldg(p.bias + (j % TN));
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < CW; ++j) x[j] = fmaxf(x[j], 0.f);
    }
#pragma unroll
    for (int j = 0; j < CW; ++j)
      if (!valid[j / TN]) x[j] = 0.f;
    __syncwarp();  // the previous set's slab reads are done
#pragma unroll
    for (int g = 0; g < P; ++g)
      sts_v4(slab_s + sts_off[g], pack_bf16x2(x[8 * g], x[8 * g + 1]), pack_bf16x2(x[8 * g + 2], x[8 * g + 3]),
             pack_bf16x2(x[8 * g + 4], x[8 * g + 5]), pack_bf16x2(x[8 * g + 6], x[8 * g + 7]));
    __syncwarp();
    uint4 o[P];
#pragma unroll
    for (int j = 0; j < P; ++j) o[j] = lds_v4(slab_s + o_lds[j]);
    if (p.accumulate) {
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const uint32_t ov[4] = {old[j].x, old[j].y, old[j].z, old[j].w};
        const uint32_t nv[4] = {o[j].x, o[j].y, o[j].z, o[j].w};
        uint32_t rv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = unpack_bf16x2(ov[e]), b2 = unpack_bf16x2(nv[e]);
          rv[e] = pack_bf16x2(a.x + b2.x, a.y + b2.y);
        }
        o[j] = make_uint4(rv[0], rv[1], rv[2], rv[3]);
      }
    }
    if (!(p.debug & 1)) {
#pragma unroll
      for (int j = 0; j < P; ++j)
        if (o_ptr[j] != nullptr) st_v4(o_ptr[j], o[j]);
    }
    if (do_stats) {
      float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
      for (int j = 0; j < 4 * P; ++j) {
        const float2 f = unpack_bf16x2(lds_u32(slab_s + j * 128 + (st_off0 ^ (uint32_t)((j & 7) << 4))));
        s1a += f.x;
        s2a = fmaf(f.x, f.x, s2a);
        s1b += f.y;
        s2b = fmaf(f.y, f.y, s2b);
      }
      st_acc[0] += s1a;
      st_acc[1] += s1b;
      st_acc[2] += s2a;
      st_acc[3] += s2b;
    }
  }
  if (do_stats) {
    // one flush per launch: lanes l and l + 16 hold the same column pair (even / odd rows); the warps park their sums
    // in their slabs, 2 * TN threads add the G tiles' columns of the four lane quarters: one atomic per channel and sum
#pragma unroll
    for (int e = 0; e < 4; ++e) st_acc[e] += __shfl_xor_sync(0xffffffffu, st_acc[e], 16);
    __syncwarp();
    sts_v4(slab_s + (uint32_t)lane * 16, __float_as_uint(st_acc[0]), __float_as_uint(st_acc[1]),
           __float_as_uint(st_acc[2]), __float_as_uint(st_acc[3]));
    named_bar_sync(kEpiBarrier, kEpiThreads);
    for (int i = et; i < 2 * TN; i += kEpiThreads) {
      const int which = i / TN, c = i - which * TN;
      if (c < p.Kout) {
        float tot = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int cl = g * TN + c, prt = cl / CW, cw = cl - prt * CW;
          const uint32_t a = smem_u32(staging) + (uint32_t)(prt * 4) * kSlabBytes + (uint32_t)(cw >> 1) * 16 +
                             (uint32_t)(which * 2 + (cw & 1)) * 4;
          tot += lds_f32(a) + lds_f32(a + kSlabBytes) + lds_f32(a + 2 * kSlabBytes) + lds_f32(a + 3 * kSlabBytes);
        }
        float* dst = (which ? p.ch_sqsum : p.ch_sum) + c;
        if (p.stat_row) dst[(long long)blockIdx.x * p.stat_row] = tot;
        else atomicAdd(dst, tot);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Halo variant of the tap-GEMM for stride-1 filters on feature maps up to 126 pixels wide.  The tap-GEMM above
// fetches the A box once per tap (9x for a 3x3 filter) and is L2->SM bandwidth bound on those layers.  Here the
// output tile is `bh` full-width rows stored with the HALO pitch bw = W + KW - 1, so that the input window of tap
// (r', q') is the same shared-memory tile read from row r'*bw + q' on: ONE TMA load per 64-channel chunk, and the
// tap offset goes into the UMMA descriptor's start address.  Output rows that fall on the
// KW-1 pad columns of the pitch are computed and discarded (W/bw of the MMA rows are useful: 56/58, 28/30, 14/16).
// When all the weights of the CTA's output-channel block fit (e.g. 64->64 3x3 = 72 KB) they are loaded once and
// stay resident; otherwise B tiles stream through their own ring.
// ------------------------------------------------------------------------------------------------
constexpr int kHaloABytes = 32768;  // 256 rows x 128 B
constexpr int kHaloMaxB = 12;

template <int BN_, bool MULTI = false>
__global__ void __launch_bounds__(kTapThreads, 1)
tapgemm_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const TapGemmParams p) {
  using Cfg = TapGemmCfg<BN_>;
  constexpr int kTmemCols = MULTI ? 2 * kMultiW : Cfg::kTmemCols;
  constexpr int kG = MULTI ? kMultiW / BN_ : 1;   // pixel tiles per accumulator set (halo_multi_epilogue)
  constexpr int kSetW = MULTI ? kMultiW : BN_;
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + p.a_stages * kHaloABytes;
  uint8_t* staging = b_ring + p.b_stages * Cfg::kBTileBytes;
  float* scratch = reinterpret_cast<float*>(staging + Cfg::kStageBufs * kATileBytes);
  __shared__ uint64_t a_full[3], a_empty[3];
  __shared__ uint64_t b_full[kHaloMaxB], b_empty[kHaloMaxB];
  __shared__ uint64_t bres_bar;
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t tap_lo_s[kMaxTaps];  // descriptor start-address increment of every tap's window

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int chunks = (p.C + kBK - 1) / kBK;

  if (threadIdx.x < p.ntaps) {
    uint32_t rowoff = (uint32_t)(p.tap_dh[threadIdx.x] * p.bw + p.tap_dw[threadIdx.x]);
    if (p.debug & 64) rowoff &= ~7u;  // timing experiment only: 1024-byte aligned windows (wrong results)
    tap_lo_s[threadIdx.x] = rowoff * (128u >> 4);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kHaloMaxB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&bres_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiThreads / 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // everything above (barriers, TMEM, descriptors) may overlap the previous kernel's tail

  if (warp == 0) {
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t a_bytes = (uint32_t)(p.bw * p.halo_box_h) * 128u;
      if (p.b_resident) {  // all weights of the (single) output-channel block: once per CTA
        mbar_expect_tx(&bres_bar, (uint32_t)(p.ntaps * chunks) * (uint32_t)Cfg::kBTileBytes);
        for (int tap = 0; tap < p.ntaps; ++tap)
          for (int ch = 0; ch < chunks; ++ch)
            tma_load_3d(b_ring + (tap * chunks + ch) * Cfg::kBTileBytes, &tmB, &bres_bar, ch * kBK,
                        (int)p.tap_w[tap], 0);
      }
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
      // Tile coordinates advance INCREMENTALLY (tile += gridDim.x = (s_w, s_h, s_n) with carries): five integer divisions
      // per tile in this single thread cost 13 us of a 200 us launch on the 16-channel 256 x 256 layers
      // (profiles/r02_narrow_layer_decomposition.txt).  One output-channel block only; otherwise the plain arithmetic.
      const bool inc = p.tiles_co == 1 && !(p.debug & 8);
      int tw = 0, th = 0, tn = 0, tco = 0;
      int s_w = 0, s_h = 0, s_n = 0;
      if (inc) {
        const int t0 = (int)blockIdx.x, st = (int)gridDim.x;
        tw = t0 % p.tiles_w; th = (t0 / p.tiles_w) % p.tiles_h; tn = t0 / (p.tiles_w * p.tiles_h);
        s_w = st % p.tiles_w; s_h = (st / p.tiles_w) % p.tiles_h; s_n = st / (p.tiles_w * p.tiles_h);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        if (!inc && !(p.debug & 8)) {
          tco = tile / p.tiles_m;
          const int tm = tile - tco * p.tiles_m;
          tw = tm % p.tiles_w;
          th = (tm / p.tiles_w) % p.tiles_h;
          tn = tm / (p.tiles_w * p.tiles_h);
        }
        const int h0 = th * p.bh, w0 = tw * p.bw_valid;
        const int co0 = tco * BN_;
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(&a_empty[sa], pha ^ 1u);
          if (p.debug & 128) {  // timing experiment: MMA / epilogue side alone (no A loads, garbage operands)
            mbar_arrive(&a_full[sa]);
          } else {
            mbar_expect_tx(&a_full[sa], a_bytes);
            tma_load_4d(a_ring + sa * kHaloABytes, &tmA, &a_full[sa], ch * kBK, w0 + p.org_dw, h0 + p.org_dh, tn);
          }
          if (++sa == (uint32_t)p.a_stages) {
            sa = 0;
            pha ^= 1u;
          }
          if (!p.b_resident) {
            for (int tap = 0; tap < p.ntaps; ++tap) {
              mbar_wait(&b_empty[sb], phb ^ 1u);
              mbar_expect_tx(&b_full[sb], (uint32_t)Cfg::kBTileBytes);
              tma_load_3d(b_ring + sb * Cfg::kBTileBytes, &tmB, &b_full[sb], ch * kBK, (int)p.tap_w[tap], co0);
              if (++sb == (uint32_t)p.b_stages) {
                sb = 0;
                phb ^= 1u;
              }
            }
          }
        }
        if (inc) {
          tw += s_w;
          const int c1 = tw >= p.tiles_w;
          tw -= c1 ? p.tiles_w : 0;
          th += s_h + c1;
          const int c2 = th >= p.tiles_h;
          th -= c2 ? p.tiles_h : 0;
          tn += s_n + c2;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN_, 0, 0);
      const uint64_t d0 = umma_smem_desc_sw128(smem_u32(a_ring), 16, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32);
      const uint32_t a_lo0 = (uint32_t)d0;
      const uint32_t b_lo0 = (uint32_t)umma_smem_desc_sw128(smem_u32(b_ring), 16, 1024);
      constexpr uint32_t kALo = kHaloABytes >> 4, kBLo = Cfg::kBTileBytes >> 4;
      const int c_tail = p.C - (chunks - 1) * kBK;
      const int nk_tail = (c_tail + 15) >> 4;
      if (p.b_resident) {
        mbar_wait(&bres_bar, 0);
        tc_fence_after();
      }
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, t = 0;
      // FAST PATH (resident weights, one channel chunk, <= 9 taps: every narrow decoder layer): the per-tap descriptor
      // words live in registers and the tap loop is straight-line code.  The generic loop below spends ~150 cycles of
      // scalar work per tap (shared-memory look-up, index arithmetic, branches) in the ONE issuing thread — invisible
      // next to a 128-cycle N = 256 MMA, but 2.5x the 64-cycle floor of the N = 16 / 32 MMAs: on the 16-channel
      // 256 x 256 layer the loop alone (MMAs removed) took 66 us of the 198 us launch
      // (profiles/r02_narrow_layer_decomposition.txt).
      if (p.b_resident && chunks == 1 && p.ntaps <= 9 && !(p.debug & (16 | 256))) {
        uint32_t ta[9], tb[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          ta[i] = i < p.ntaps ? tap_lo_s[i] : 0u;
          tb[i] = b_lo0 + (uint32_t)i * kBLo;
        }
        const int nk = nk_tail;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
          const uint32_t su = t / (uint32_t)kG, sg = t % (uint32_t)kG;
          const uint32_t as = su & 1u;
          if (sg == 0) {
            mbar_wait(&tempty_bar[as], ((su >> 1) & 1u) ^ 1u);
            tc_fence_after();
          }
          const uint32_t tmem_d = tmem_base + as * kSetW + sg * BN_;
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + sa * kALo;
          uint32_t acc = 0;
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            if (i < p.ntaps) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < nk) {
                  umma_bf16_lohi(tmem_d, a_lo + ta[i] + 2 * k, desc_hi, tb[i] + 2 * k, desc_hi, idesc, acc);
                  acc = 1u;
                }
              }
            }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == (uint32_t)p.a_stages) {
            sa = 0;
            pha ^= 1u;
          }
          if (sg == (uint32_t)(kG - 1) || tile + (int)gridDim.x >= p.total_tiles) umma_commit(&tfull_bar[as]);
        }
      } else
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
        // accumulator set `as` (double buffered); MULTI: kG consecutive tiles of this CTA fill one set side by side
        const uint32_t su = t / (uint32_t)kG, sg = t % (uint32_t)kG;
        const uint32_t as = su & 1u;
        if (sg == 0) {
          mbar_wait(&tempty_bar[as], ((su >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + as * kSetW + sg * BN_;
        uint32_t acc = 0;
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + sa * kALo;
          const int nk = (ch + 1 < chunks) ? 4 : nk_tail;
          for (int tap = 0; tap < ((p.debug & 16) ? 0 : p.ntaps); ++tap) {  // (timing experiment 16: no tap loop at all)
            uint32_t b_lo;
            if (p.b_resident) {
              b_lo = b_lo0 + (uint32_t)(tap * chunks + ch) * kBLo;
            } else {
              mbar_wait(&b_full[sb], phb);
              tc_fence_after();
              b_lo = b_lo0 + sb * kBLo;
            }
            // The 128B swizzle of a UMMA operand is a function of the absolute shared-memory address bits (like
            // the TMA write; measured on B200), so the window of tap (r', q') needs only the start address shifted
            // by r'*bw + q' rows of 128 B — the descriptor's "matrix base offset" field stays 0.
            const uint32_t ta_lo = a_lo + tap_lo_s[tap];
            if (p.debug & 256) {  // timing experiment: load + epilogue side alone (no MMAs)
            } else if (nk == 4) {
              umma_bf16_lohi(tmem_d, ta_lo, desc_hi, b_lo, desc_hi, idesc, acc);
              umma_bf16_lohi(tmem_d, ta_lo + 2, desc_hi, b_lo + 2, desc_hi, idesc, 1u);
              umma_bf16_lohi(tmem_d, ta_lo + 4, desc_hi, b_lo + 4, desc_hi, idesc, 1u);
              umma_bf16_lohi(tmem_d, ta_lo + 6, desc_hi, b_lo + 6, desc_hi, idesc, 1u);
            } else {
              for (int k = 0; k < nk; ++k)
                umma_bf16_lohi(tmem_d, ta_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, acc | (uint32_t)k);
            }
            acc = 1u;
            if (!p.b_resident) {
              umma_commit(&b_empty[sb]);
              if (++sb == (uint32_t)p.b_stages) {
                sb = 0;
                phb ^= 1u;
              }
            }
          }
          umma_commit(&a_empty[sa]);
          if (++sa == (uint32_t)p.a_stages) {
            sa = 0;
            pha ^= 1u;
          }
        }
        if (sg == (uint32_t)(kG - 1) || tile + (int)gridDim.x >= p.total_tiles) umma_commit(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    if constexpr (MULTI)
      halo_multi_epilogue<BN_>(p, staging, tmem_base, tfull_bar, tempty_bar, warp, lane);
    else
      tap_epilogue<BN_>(p, staging, scratch, tmem_base, tfull_bar, tempty_bar, warp, lane, (int)blockIdx.x,
                        (int)gridDim.x, p.total_tiles, p.tiles_m, -1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[co][tap][ci] = sum_pix dY[pix][co] * X[pix + tap][ci]
// ------------------------------------------------------------------------------------------------
struct WgradParams {
  int bw, bh, bn, rows;
  int tiles_w, tiles_h, tiles_n, tiles_m;
  int sxw, sxh, pad_t, pad_l, KW;
  int Cw;      // packed weight inner dim (input channels, or 64 window elements in row-window mode)
  int Kout;    // output channels
  int ntaps, cchunks, nsub;  // nsub = csub * tsub 64-column sub-tiles of the accumulator per CTA (1..4)
  int csub, tsub;            // ... = csub 64-channel chunks x tsub filter taps (narrow layers batch taps)
  int rowwin;
  int splits;
  int dys;                 // element stride of dY's W / H coordinates (2: the folded up-conv reads a parity class of dY)
  long long split_stride;  // elements between the partial dW of consecutive splits
  float* dw;
};

// Pixel tiles of <= 64 rows in a 4-deep ring (192 KB in flight per SM; two 96 KB stages measured the same, 4.08 vs
// 4.09 ms over the ResNet-50 layers: the kernel is bound by the operand re-reads through L2 — every x tile is fetched
// once per output-channel block and tap group — not by the ring depth; profiles/r01_experiments.txt).
constexpr int kWgRows = 64;
constexpr int kWgAtomBytes = kWgRows * 128;  // one [pixels x 64 channels] MN-major atom
constexpr int kWgStages = 4;
constexpr int kWgMaxSub = 4;
constexpr int kWgStageBytes = (2 + kWgMaxSub) * kWgAtomBytes;  // dY halves (2 atoms) + X (<= 4 atoms)
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 1024;
constexpr int kWgTmemCols = 256;

__global__ void __launch_bounds__(kConvThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kWgStages];
  __shared__ uint64_t empty_bar[kWgStages];
  __shared__ uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int tgrp = blockIdx.x / p.cchunks;
  const int tap0 = tgrp * p.tsub;
  const int ci0 = (blockIdx.x - tgrp * p.cchunks) * (64 * p.csub);
  const int co0 = blockIdx.y * 128;
  const int ntap_here = (p.ntaps - tap0) < p.tsub ? (p.ntaps - tap0) : p.tsub;
  const int nsub = ntap_here * p.csub;  // accumulator sub-tiles of THIS block: (tap0 + j / csub, ci0 + 64 * (j % csub))
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
  if (warp == 1) tmem_alloc<kWgTmemCols>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // everything above (barriers, TMEM, descriptors) may overlap the previous kernel's tail

  if (warp == 0) {
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      tma_prefetch_desc(&tmDY);
      tma_prefetch_desc(&tmX);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u * (uint32_t)(2 + nsub);
      int sub_dw[kWgMaxSub], sub_dh[kWgMaxSub], sub_c[kWgMaxSub];  // per sub-tile: tap offset and first channel
#pragma unroll
      for (int j = 0; j < kWgMaxSub; ++j) {
        const int tj = j / p.csub, cj = j - tj * p.csub;
        const int tap = tap0 + tj;
        const int r = p.rowwin ? tap : tap / p.KW;
        const int qx = p.rowwin ? 0 : tap - r * p.KW;
        sub_dw[j] = qx - (p.rowwin ? 0 : p.pad_l);
        sub_dh[j] = r - p.pad_t;
        sub_c[j] = ci0 + 64 * cj;
      }
      // pixel-tile coordinates advance incrementally (tm += splits) instead of three integer divisions per 64-pixel
      // tile in this one thread: ~500 cycles of scalar work against 512 cycles of MMAs made the PRODUCER the bound of
      // the narrow-layer wgrads (profiles/r02_narrow_layer_decomposition.txt)
      int tw = split % p.tiles_w, th = (split / p.tiles_w) % p.tiles_h, tn = split / (p.tiles_w * p.tiles_h);
      const int s_w = p.splits % p.tiles_w, s_h = (p.splits / p.tiles_w) % p.tiles_h,
                s_n = p.splits / (p.tiles_w * p.tiles_h);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kWgStages;
        const uint32_t ph = (uint32_t)(it / kWgStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        {
          tw += s_w;
          const int c1 = tw >= p.tiles_w;
          tw -= c1 ? p.tiles_w : 0;
          th += s_h + c1;
          const int c2 = th >= p.tiles_h;
          th -= c2 ? p.tiles_h : 0;
          tn += s_n + c2;
        }
        uint8_t* a_s = smem + s * kWgStageBytes;
        mbar_expect_tx(&full_bar[s], tx_bytes);
        tma_load_4d(a_s, &tmDY, &full_bar[s], co0, w0 * p.dys, h0 * p.dys, n0);
        tma_load_4d(a_s + kWgAtomBytes, &tmDY, &full_bar[s], co0 + 64, w0 * p.dys, h0 * p.dys, n0);
        const int xw0 = w0 * p.sxw, xh0 = h0 * p.sxh;
#pragma unroll
        for (int j = 0; j < kWgMaxSub; ++j)
          if (j < nsub)
            tma_load_4d(a_s + (2 + j) * kWgAtomBytes, &tmX, &full_bar[s], sub_c[j], xw0 + sub_dw[j], xh0 + sub_dh[j],
                        n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync, not lane==0: ptxas then emits bare UTCHMMA / UTMALDG (no per-lane loop)
      const uint32_t idesc = umma_idesc_bf16(128, 64 * nsub, 1, 1);
      const int nk = (p.rows + 15) >> 4;
      // MN-major: 8-row (K) groups 1024 B apart; consecutive 64-channel atoms are one atom (8 KB) apart
      const uint64_t d0 = umma_smem_desc_sw128(smem_u32(smem), kWgAtomBytes, 1024);
      const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
      constexpr uint32_t kStageLo = kWgStageBytes >> 4, kXLo = (2 * kWgAtomBytes) >> 4;
      uint32_t acc = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t s = (uint32_t)it % kWgStages;
        mbar_wait(&full_bar[s], ((uint32_t)it / kWgStages) & 1u);
        tc_fence_after();
        const uint32_t a_lo = lo0 + s * kStageLo, b_lo = a_lo + kXLo;
        if (nk == 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lohi(tmem_base, a_lo + 128 * k, desc_hi, b_lo + 128 * k, desc_hi, idesc, k ? 1u : acc);
        } else {
          for (int k = 0; k < nk; ++k)
            umma_bf16_lohi(tmem_base, a_lo + 128 * k, desc_hi, b_lo + 128 * k, desc_hi, idesc, acc | (uint32_t)k);
        }
        acc = 1u;
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    float* dbase = p.dw + (long long)split * p.split_stride + (long long)co * p.ntaps * p.Cw;
    const int ncols = 64 * nsub;
    if (my_tiles > 0) {
      if (lane == 0) mbar_wait(&accum_bar, 0);
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c, v);
        tmem_ld_wait();
        const int j = c >> 6, tj = j / p.csub, cj = j - tj * p.csub;
        const int ci = ci0 + 64 * cj + (c & 63);
        float* dst = dbase + (long long)(tap0 + tj) * p.Cw + ci;
        if (co < p.Kout) {
          if (ci + 32 <= p.Cw) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              *reinterpret_cast<float4*>(dst + jj) =
                  make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                              __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (ci + jj < p.Cw) dst[jj] = __uint_as_float(v[jj]);
          }
        }
      }
    } else if (co < p.Kout) {
      for (int c = 0; c < ncols; ++c) {
        const int j = c >> 6, tj = j / p.csub, cj = j - tj * p.csub;
        const int ci = ci0 + 64 * cj + (c & 63);
        if (ci < p.Cw) dbase[(long long)(tap0 + tj) * p.Cw + ci] = 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kWgTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_w_fprop_kernel(const float* __restrict__ w, int K, int C, int taps, int Cpad,
                                    __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
// both packings in one launch (one per convolution per step)
__global__ void pack_w_both_kernel(const float* __restrict__ w, int K, int C, int taps, int Cpad, int Kpad,
                                   __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  pdl_trigger();
  pdl_wait();
  const long long nf = (long long)K * taps * Cpad, nd = (long long)Cpad * taps * Kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nf + nd;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < nf) {
      const int c = (int)(i % Cpad);
      const long long t = i / Cpad;
      const int tap = (int)(t % taps);
      const int k = (int)(t / taps);
      wf[i] = __float2bfloat16_rn(c < C ? w[((long long)k * C + c) * taps + tap] : 0.f);
    } else {
      const long long i2 = i - nf;
      const int k = (int)(i2 % Kpad);
      const long long t = i2 / Kpad;
      const int tap = (int)(t % taps);
      const int c = (int)(t / taps);
      wd[i2] = __float2bfloat16_rn((c < C && k < K) ? w[((long long)k * C + c) * taps + tap] : 0.f);
    }
  }
}
// row-window packing: out[k][r][q * cpp + c] = w[k][c][r][q]  (64 window elements per filter row)
// every convolution of a model in ONE launch: items[i] = {w, wf, wd (0 = none), K, C, taps, Cpad, Kpad, first block,
// blocks} as int64; blocks are dealt out in proportion to the item's element count (a fixed share per item left the
// largest layers of the U-Net with 1/128 of the parallelism they had and cost more than the launches saved).  Same
// element mapping as pack_w_both_kernel.  The per-layer launches cost ~6.7 us each for microseconds of work: 49 of them
// per ResNet-50 step, 79 per U-Net step.
constexpr int kPackItemWords = 10;
constexpr int kPackElemsPerBlock = 256 * 8;
constexpr int kPackTile = 32;       // tiled mode (taps <= 9): one block = 32 output x 32 input channels x all taps
constexpr int kPackTiledMaxTaps = 9;
__global__ void __launch_bounds__(256)
pack_w_batched_kernel(const long long* __restrict__ items, int n_items) {
  int lo = 0, hi = n_items - 1;  // last item whose first block <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (items[(long long)mid * kPackItemWords + 8] <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const long long* it = items + (long long)lo * kPackItemWords;
  const float* __restrict__ w = reinterpret_cast<const float*>(it[0]);
  __nv_bfloat16* __restrict__ wf = reinterpret_cast<__nv_bfloat16*>(it[1]);
  __nv_bfloat16* __restrict__ wd = reinterpret_cast<__nv_bfloat16*>(it[2]);
  const int K = (int)it[3], C = (int)it[4], taps = (int)it[5], Cpad = (int)it[6], Kpad = (int)it[7];
  const long long blk = (long long)blockIdx.x - it[8], nblk = it[9];
  if (taps <= kPackTiledMaxTaps) {
    // Tiled transposition through shared memory: the fp32 OIHW source is read in contiguous runs of 32 x taps floats
    // per output channel, both bf16 destinations are written in 64-byte runs (the element-wise version below read the
    // source with a stride of `taps` floats and spent 450 us per R50 U-Net step on 0.45 GB of traffic).
    __shared__ float tile[kPackTile * (kPackTile * kPackTiledMaxTaps + 1)];
    const int pitch = kPackTile * taps + 1;  // odd pitch: the output-channel-major read below is conflict free
    const int tiles_c = (Cpad + kPackTile - 1) / kPackTile;
    const int tk = (int)(blk / tiles_c), tc = (int)(blk - (long long)tk * tiles_c);
    const int k0 = tk * kPackTile, c0 = tc * kPackTile;
    const int run = kPackTile * taps;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps: no per-element division anywhere below
    for (int kk = ty; kk < kPackTile; kk += 8) {              // warp = one output channel: contiguous source run
      const bool kok = k0 + kk < K;
      const float* src = w + ((long long)(k0 + kk) * C + c0) * taps;
      int cc = tx / taps, tap = tx - cc * taps;               // element r = tx, tx + 32, ... of the run: (cc, tap)
      const int dcc = 32 / taps, dtap = 32 - dcc * taps;
      for (int r = tx; r < run; r += 32) {
        tile[kk * pitch + r] = (kok && c0 + cc < C) ? src[r] : 0.f;
        cc += dcc; tap += dtap;
        if (tap >= taps) { tap -= taps; ++cc; }
      }
    }
    __syncthreads();
    if (c0 + tx < Cpad) {                                     // wf[k][tap][c]: lane = input channel (64-byte runs)
      for (int kk = ty; kk < kPackTile; kk += 8) {
        if (k0 + kk >= K) break;
        __nv_bfloat16* dst = wf + (long long)(k0 + kk) * taps * Cpad + c0 + tx;
        const float* srow = tile + kk * pitch + tx * taps;
        for (int tap = 0; tap < taps; ++tap) dst[(long long)tap * Cpad] = __float2bfloat16_rn(srow[tap]);
      }
    }
    if (wd != nullptr && k0 + tx < Kpad) {                    // wd[c][tap][k]: lane = output channel
      for (int cc = ty; cc < kPackTile; cc += 8) {
        if (c0 + cc >= Cpad) break;
        __nv_bfloat16* dst = wd + (long long)(c0 + cc) * taps * Kpad + k0 + tx;
        const float* scol = tile + tx * pitch + cc * taps;
        for (int tap = 0; tap < taps; ++tap) dst[(long long)tap * Kpad] = __float2bfloat16_rn(scol[tap]);
      }
    }
    return;
  }
  const long long nf = (long long)K * taps * Cpad, nd = wd ? (long long)Cpad * taps * Kpad : 0;
  for (long long i = blk * blockDim.x + threadIdx.x; i < nf + nd; i += nblk * blockDim.x) {
    if (i < nf) {
      const int c = (int)(i % Cpad);
      const long long t = i / Cpad;
      const int tap = (int)(t % taps);
      const int k = (int)(t / taps);
      wf[i] = __float2bfloat16_rn(c < C ? w[((long long)k * C + c) * taps + tap] : 0.f);
    } else {
      const long long i2 = i - nf;
      const int k = (int)(i2 % Kpad);
      const long long t = i2 / Kpad;
      const int tap = (int)(t % taps);
      const int c = (int)(t / taps);
      wd[i2] = __float2bfloat16_rn((c < C && k < K) ? w[((long long)k * C + c) * taps + tap] : 0.f);
    }
  }
}
__global__ void pack_w_rowwin_kernel(const float* __restrict__ w, int K, int C, int KH, int KW, int cpp,
                                     __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)K * KH * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % 64);
    const long long t = i / 64;
    const int r = (int)(t % KH);
    const int k = (int)(t / KH);
    const int q = e / cpp, c = e - q * cpp;
    const float v = (q < KW && c < C) ? w[(((long long)k * C + c) * KH + r) * KW + q] : 0.f;
    out[i] = __float2bfloat16_rn(v);
  }
}
// sum of `splits` packed partials [K][taps][Cpad] -> OIHW fp32
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int splits, long long split_stride,
                                    int K, int C, int taps, int Cpad, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)K * C * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long t = i / taps;
    const int c = (int)(t % C);
    const int k = (int)(t / C);
    const float* src = dwp + ((long long)k * taps + tap) * Cpad + c;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[(long long)s * split_stride];
    out[i] = acc;
  }
}
// row-window partials [K][KH][64] -> OIHW
__global__ void unpack_wgrad_rowwin_kernel(const float* __restrict__ dwp, int splits,
                                           long long split_stride, int K, int C, int KH, int KW,
                                           int cpp, float* __restrict__ out) {
  const long long total = (long long)K * C * KH * KW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % KW);
    long long t = i / KW;
    const int r = (int)(t % KH);
    t /= KH;
    const int c = (int)(t % C);
    const int k = (int)(t / C);
    const float* src = dwp + ((long long)k * KH + r) * 64 + q * cpp + c;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += src[(long long)s * split_stride];
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Folded up-convolution: nn.Upsample(scale_factor=2) (nearest) -> Conv2d(k=2, padding='same') (UpConvBlock,
// segmentation/models/blocks.py:531-535) WITHOUT the x4 tensor.  Output pixel (2i+a, 2j+b) reads the up-sampled rows
// 2i+a+r, r in {0,1}, i.e. the LOW-RES rows i (a = 0: both taps) or i, i+1 (a = 1): four output-parity classes of
// 1x1 / 1x2 / 2x1 / 2x2 convolutions on the low-res input with pre-summed weights — 9 taps instead of 16 per 2x2 output
// block (9/16 of the FLOPs), the A operand 4x smaller, and no up-sample kernels in either direction.
// Folded tap order: class (a,b) = (0,0),(0,1),(1,0),(1,1) at bases 0,1,3,5; inside a class (dr,dq) row-major.
// ------------------------------------------------------------------------------------------------
__constant__ int kFoldBase[4] = {0, 1, 3, 5};
__global__ void fold_upconv_kernel(const float* __restrict__ w, long long KC, float* __restrict__ wf) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < KC; i += (long long)gridDim.x * blockDim.x) {
    const float w00 = w[i * 4], w01 = w[i * 4 + 1], w10 = w[i * 4 + 2], w11 = w[i * 4 + 3];
    float* o = wf + i * 9;
    o[0] = (w00 + w01) + (w10 + w11);                 // (0,0): all four taps land on (i, j)
    o[1] = w00 + w10;  o[2] = w01 + w11;              // (0,1): columns j, j+1
    o[3] = w00 + w01;  o[4] = w10 + w11;              // (1,0): rows i, i+1
    o[5] = w00; o[6] = w01; o[7] = w10; o[8] = w11;   // (1,1)
  }
}
// gradient of the fold: dW[r][q] = sum over classes of the folded tap that (r, q) was summed into
__global__ void unfold_upconv_kernel(const float* __restrict__ g00, const float* __restrict__ g01,
                                     const float* __restrict__ g10, const float* __restrict__ g11, long long KC,
                                     float* __restrict__ dw, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < KC; i += (long long)gridDim.x * blockDim.x) {
    const float a = g00[i];
    const float b0 = g01[i * 2], b1 = g01[i * 2 + 1];
    const float c0 = g10[i * 2], c1 = g10[i * 2 + 1];
    const float d00 = g11[i * 4], d01 = g11[i * 4 + 1], d10 = g11[i * 4 + 2], d11 = g11[i * 4 + 3];
    float r[4] = {((a + b0) + c0) + d00, ((a + b1) + c0) + d01, ((a + b0) + c1) + d10, ((a + b1) + c1) + d11};
#pragma unroll
    for (int t = 0; t < 4; ++t) dw[i * 4 + t] = accumulate ? dw[i * 4 + t] + r[t] : r[t];
  }
}

// Every weight gradient of a backward pass in ONE launch (msp_unpack_wgrad_batched): per item the fixed-order sum of
// the split-K partials, written (or added) in the OIHW layout of nn.Conv2d.weight.grad.  The per-layer launches above
// cost ~8 us each for microseconds of work: 50 of them per ResNet-50 step, 80 per U-Net step (5 % of the cfg3 step).
// The item table travels in the kernel's parameter space (no device table to keep in sync with the allocator).
constexpr int kUnpackMaxItems = 96;
constexpr int kUnpackWideSplits = 48;   // items with at least this many partials: splits shared by 8 thread groups
struct UnpackList {
  msp_unpack_item it[kUnpackMaxItems];
  int first_block[kUnpackMaxItems + 1];
  int n;
};
__global__ void __launch_bounds__(256) unpack_wgrad_batched_kernel(const __grid_constant__ UnpackList L) {
  int lo = 0, hi = L.n - 1;  // last item whose first block <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (L.first_block[mid] <= (int)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const msp_unpack_item& it = L.it[lo];
  const long long blk = (long long)blockIdx.x - L.first_block[lo], nblk = L.first_block[lo + 1] - L.first_block[lo];
  const float* __restrict__ dwp = it.partials;
  float* __restrict__ out = it.dst;
  const int splits = it.splits;
  const long long ss = it.split_stride;
  if (it.rowwin_KH == 0 && splits >= kUnpackWideSplits) {
    // many partials of a small gradient (the narrow-channel wgrad kernel: one partial per CTA, ~300 x 9 KB): a serial
    // loop over the splits is ~300 dependent L2 round trips per thread.  Here 8 thread groups share the splits of 32
    // elements, 4 independent accumulators each; the association is fixed by the code (deterministic).
    __shared__ float part[8][33];
    const int taps = it.taps, C = it.C_true, Cpad = it.Cpad;
    const long long total = (long long)it.K * C * taps;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    for (long long base = blk * 32; base < total; base += nblk * 32) {
      const long long i = base + lane < total ? base + lane : total - 1;
      const int tap = (int)(i % taps);
      const long long t = i / taps;
      const int c = (int)(t % C), k = (int)(t / C);
      const float* src = dwp + ((long long)k * taps + tap) * Cpad + c;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int sp = grp;
      for (; sp + 24 < splits; sp += 32) {
        a0 += src[(long long)sp * ss];
        a1 += src[(long long)(sp + 8) * ss];
        a2 += src[(long long)(sp + 16) * ss];
        a3 += src[(long long)(sp + 24) * ss];
      }
      for (; sp < splits; sp += 8) a0 += src[(long long)sp * ss];
      part[grp][lane] = (a0 + a1) + (a2 + a3);
      __syncthreads();
      if (grp == 0 && base + lane < total) {
        float acc = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) acc += part[g][lane];
        out[i] = it.accumulate ? out[i] + acc : acc;
      }
      __syncthreads();
    }
    return;
  }
  if (it.rowwin_KH > 0) {
    const int KH = it.rowwin_KH, KW = it.rowwin_KW, cpp = it.rowwin_cpp, C = it.C_true;
    const long long total = (long long)it.K * C * KH * KW;
    for (long long i = blk * 256 + threadIdx.x; i < total; i += nblk * 256) {
      const int q = (int)(i % KW);
      long long t = i / KW;
      const int r = (int)(t % KH);
      t /= KH;
      const int c = (int)(t % C);
      const int k = (int)(t / C);
      const float* src = dwp + ((long long)k * KH + r) * 64 + q * cpp + c;
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += src[(long long)sp * ss];
      out[i] = it.accumulate ? out[i] + acc : acc;
    }
  } else {
    const int taps = it.taps, C = it.C_true, Cpad = it.Cpad;
    const long long total = (long long)it.K * C * taps;
    for (long long i = blk * 256 + threadIdx.x; i < total; i += nblk * 256) {
      const int tap = (int)(i % taps);
      const long long t = i / taps;
      const int c = (int)(t % C);
      const int k = (int)(t / C);
      const float* src = dwp + ((long long)k * taps + tap) * Cpad + c;
      float acc = 0.f;
      for (int sp = 0; sp < splits; ++sp) acc += src[(long long)sp * ss];
      out[i] = it.accumulate ? out[i] + acc : acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Box {
  int bw, bh, bn;
};

// Largest-utilisation box of <= max_rows (128) output pixels: full rows first, then rows, then images.
Box pick_box(int OW, int OH, int N, int max_rows = 128) {
  Box b{1, 1, 1};
  if (OW >= max_rows) {
    const int t = msp_cdiv(OW, max_rows);
    b.bw = msp_cdiv(OW, t);
    return b;
  }
  b.bw = OW;
  const int maxh = max_rows / OW;
  if (OH > maxh) {
    const int t = msp_cdiv(OH, maxh);
    b.bh = msp_cdiv(OH, t);
    return b;
  }
  b.bh = OH;
  const int maxn = max_rows / (OW * OH);
  if (N > maxn) {
    const int t = msp_cdiv(N, maxn);
    b.bn = msp_cdiv(N, t);
  } else {
    b.bn = N;
  }
  return b;
}

template <int BN_>
int launch_tapgemm(const CUtensorMap& tmA, const CUtensorMap& tmB, TapGemmParams& p, cudaStream_t st) {
  using Cfg = TapGemmCfg<BN_>;
  static bool attr_set = false;
  if (!attr_set) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_kernel<BN_>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  {
    static int dbg = -1, dst = 0;
    if (dbg < 0) {
      const char* e = getenv("MSP_CONV_DEBUG"); dbg = e ? atoi(e) : 0;
      const char* f = getenv("MSP_CONV_STAGES"); dst = f ? atoi(f) : 0;
      if (dst > TapGemmCfg<BN_>::kStages) dst = TapGemmCfg<BN_>::kStages;
    }
    p.debug = dbg;
    p.dbg_stages = dst;
  }
  p.bw_valid = p.bw;
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_co = msp_cdiv(p.Kout, BN_);
  const long long total = (long long)p.tiles_m * p.tiles_co;
  MSP_REQUIRE(total < (1ll << 31), "conv: too many tiles");
  p.total_tiles = (int)total;
  const int sms = msp_num_sms();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  MSP_CHECK_CUDA(msp_launch_pdl(tapgemm_kernel<BN_>, dim3(grid), dim3(kTapThreads), Cfg::kSmemBytes, st, tmA, tmB, p));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  g_last_kernel = BN_ == 256 ? "tapgemm_kernel<256>" : BN_ == 128 ? "tapgemm_kernel<128>" : BN_ == 64 ? "tapgemm_kernel<64>"
                  : BN_ == 32 ? "tapgemm_kernel<32>" : "tapgemm_kernel<16>";
  return MSP_OK;
}

template <int BN_>
int launch_tapgemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, TapGemmParams& p, cudaStream_t st) {
  using Cfg2 = TapGemm2Cfg<BN_>;
  static bool attr_set = false;
  if (!attr_set) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(tapgemm2_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg2::kSmemBytes));
    attr_set = true;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MSP_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.bw_valid = p.bw;
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_co = msp_cdiv(p.Kout, BN_);
  const long long items = (long long)((p.tiles_m + 1) / 2) * p.tiles_co;
  MSP_REQUIRE(items < (1ll << 30), "conv: too many tiles");
  p.total_tiles = (int)((long long)p.tiles_m * p.tiles_co);
  const int pairs = msp_num_sms() / 2;
  const int grid = 2 * (int)(items < pairs ? items : pairs);
  tapgemm2_kernel<BN_><<<grid, kTapThreads, Cfg2::kSmemBytes, st>>>(tmA, tmB, p);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  g_last_kernel = BN_ == 256 ? "tapgemm2_kernel<256>" : BN_ == 128 ? "tapgemm2_kernel<128>" : "tapgemm2_kernel<64>";
  return MSP_OK;
}

inline int bn_tile_for(int K) {
  return K <= 16 ? 16 : (K <= 32 ? 32 : (K <= 64 ? 64 : (K <= 128 ? 128 : 256)));
}

// Kernel-variant policy (msp_conv_set_policy / environment), measured on ResNet-50 B=256 (profiles/r01_conv_variants.txt):
//  pair (cta_group::2 tap-GEMM): 0 never (default: not faster than the single-CTA kernel on these shapes),
//                                1 for tiles of >= 128 output channels, 2 also for 64
//  halo (A tile loaded once per chunk, taps via descriptor offsets): 0 never, 1 (default) when the weights of the
//                                output-channel block stay resident in shared memory, 2 whenever it applies
int g_pair_policy = -1, g_halo_policy = -1;
int pair_policy() {
  if (g_pair_policy < 0) { const char* e = getenv("MSP_CONV_2CTA"); g_pair_policy = e ? atoi(e) : 0; }
  return g_pair_policy;
}
int halo_policy() {
  if (g_halo_policy < 0) {
    const char* e = getenv("MSP_CONV_NOHALO");
    const char* f = getenv("MSP_CONV_HALO");
    g_halo_policy = (e && atoi(e)) ? 0 : (f ? atoi(f) : 1);
  }
  return g_halo_policy;
}

struct WMapArgs {  // packed weights [rows][taps][inner] (make_w_map)
  const void* base;
  int inner, taps, rows;
};
int make_w_map(CUtensorMap* m, const void* base, int inner, int taps, int rows, int box_rows);

// Output-channel tile width of the tap kernel.  The natural width (bn_tile_for) minimises the A re-reads and the per-tile
// epilogue overhead, and it is kept whenever it yields at least one tile per SM.  A launch with FEWER tiles than SMs leaves
// most of the GPU idle: the deep layers of the R50 U-Net at batch 24 have M = 1536 pixels = 12 row tiles, i.e. 24 CTAs of
// 148 at BN = 256 (profiles/r01_convbench_unet50_b24.txt: 39 us for 5 us of work).  Only then narrower tiles are
// considered, by a makespan estimate per CTA in cycles: main loop = k-blocks x max(MMA issue, L2 -> SM operand feed at
// ~74 B/clk), epilogue ~ 24 cycles per accumulator column.  (Applied to well-filled launches as well, the same estimate
// picked 64-wide tiles for ResNet-50's 1x1 layers at batch 256 and made them up to 2x slower — it underestimates the
// per-tile cost of narrow tiles — profiles/r02_experiments.txt.)  MSP_CONV_BNPOL=0 restores the fixed choice.
int pick_bn(int Kout, long long tiles_m, int kblocks) {
  const int natural = bn_tile_for(Kout);
  static int pol = -1;
  if (pol < 0) { const char* e = getenv("MSP_CONV_BNPOL"); pol = e ? atoi(e) : 1; }
  const int sms = msp_num_sms();
  if (pol == 0 || natural <= 64 || tiles_m * msp_cdiv(Kout, natural) >= sms) return natural;
  int best = natural;
  double best_t = 1e30;
  for (int bn = natural; bn >= 64; bn >>= 1) {
    const long long tiles = tiles_m * msp_cdiv(Kout, bn);
    const double n = (double)((tiles + sms - 1) / sms);
    const double mma = 4.0 * (bn >= 256 ? 128.0 : 64.0), load = (16384.0 + bn * 128.0) / 74.0;
    const double main = kblocks * (mma > load ? mma : load), epi = 24.0 * bn;
    const double t = main + (n - 1.0) * (main > epi ? main : epi) + epi;
    if (t < best_t * 0.97) { best_t = t; best = bn; }   // a narrower tile must win by 3 %
  }
  return best;
}

int dispatch_tapgemm(const CUtensorMap& tmA, const WMapArgs& w, TapGemmParams& p, cudaStream_t st) {
  const int bn = p.halo ? bn_tile_for(p.Kout)
                        : pick_bn(p.Kout, (long long)p.tiles_w * p.tiles_h * p.tiles_n, p.ntaps * msp_cdiv(p.C, kBK));
  const int pol = pair_policy();
  CUtensorMap tmB;
  if (p.tiles_w * p.tiles_h * p.tiles_n >= 2 && !p.halo && ((bn >= 128 && pol >= 1) || (bn == 64 && pol >= 2))) {
    int rc = make_w_map(&tmB, w.base, w.inner, w.taps, w.rows, bn / 2);  // each CTA of the pair loads half of B
    if (rc) return rc;
    if (bn == 256) return launch_tapgemm2<256>(tmA, tmB, p, st);
    if (bn == 128) return launch_tapgemm2<128>(tmA, tmB, p, st);
    return launch_tapgemm2<64>(tmA, tmB, p, st);
  }
  int rc = make_w_map(&tmB, w.base, w.inner, w.taps, w.rows, bn);
  if (rc) return rc;
  switch (bn) {
    case 16: return launch_tapgemm<16>(tmA, tmB, p, st);
    case 32: return launch_tapgemm<32>(tmA, tmB, p, st);
    case 64: return launch_tapgemm<64>(tmA, tmB, p, st);
    case 128: return launch_tapgemm<128>(tmA, tmB, p, st);
    default: return launch_tapgemm<256>(tmA, tmB, p, st);
  }
}

// Decide whether the halo kernel applies (taps already in p.tap_dh/dw as absolute offsets) and re-plan the tiling.
// A tensor = (Ain_W, Ain_H) feature map read by the taps; output sub-grid OW x OH x N, stride 1.
bool plan_halo(TapGemmParams& p, int OW, int OH, int N, int BN, Box* box, int* halo_w) {
  const int policy = halo_policy();
  if (policy == 0 || p.ntaps <= 1) return false;
  int dh0 = 127, dh1 = -127, dw0 = 127, dw1 = -127;
  for (int t = 0; t < p.ntaps; ++t) {
    dh0 = p.tap_dh[t] < dh0 ? p.tap_dh[t] : dh0; dh1 = p.tap_dh[t] > dh1 ? p.tap_dh[t] : dh1;
    dw0 = p.tap_dw[t] < dw0 ? p.tap_dw[t] : dw0; dw1 = p.tap_dw[t] > dw1 ? p.tap_dw[t] : dw1;
  }
  const int ext_h = dh1 - dh0 + 1, ext_w = dw1 - dw0 + 1;
  // tile = bh rows x bwv output columns stored with the halo pitch bw = bwv + ext_w - 1 (<= 128 MMA rows per tile);
  // wide feature maps are cut into column strips.  Pick the pitch with the most useful MMA rows.
  int bw = 0, bwv = 0, bh = 0;
  double best = 0.0;
  const int pmax = OW + ext_w - 1 < 128 ? OW + ext_w - 1 : 128;
  for (int P = ext_w; P <= pmax; ++P) {
    const int v = P - (ext_w - 1);
    int h = 128 / P;
    if (h < 1) continue;
    if (h > OH) h = OH;
    h = msp_cdiv(OH, msp_cdiv(OH, h));  // balance the row tiles of an image (14 rows -> 7 + 7, not 8 + 6)
    if (P * (h + ext_h - 1) > 256) continue;                         // halo tile must fit 32 KB
    if ((ext_h - 1) * P + (ext_w - 1) + 128 > 256) continue;         // the M=128 read window stays inside it
    const double util = (double)OW * OH / ((double)msp_cdiv(OW, v) * msp_cdiv(OH, h) * 128.0);
    if (util > best + 1e-9 || (util > best - 1e-9 && P > bw)) { best = util; bw = P; bwv = v; bh = h; }
  }
  if (bw == 0 || best < 0.55) return false;                          // too few useful MMA rows: tap path
  const int box_h = bh + ext_h - 1;
  const int chunks = msp_cdiv(p.C, kBK);
  const long long bt = (long long)BN * kBK * 2;
  const long long fixed = (BN >= 256 ? 1 : 2) * (long long)kATileBytes + (kEpiSlabs * 128 * 4 + 4 * 128 * 4) + 1024 + 512;
  const long long budget = 232448 - fixed;
  const long long res_bytes = (long long)p.ntaps * chunks * bt;
  p.b_resident = (msp_cdiv(p.Kout, BN) == 1 && res_bytes + 2 * kHaloABytes <= budget) ? 1 : 0;
  if (!p.b_resident && policy < 2) return false;  // streaming-B halo is slower than the tap kernel (MMA-issue bound)
  if (p.b_resident) {
    p.b_stages = p.ntaps * chunks;
    p.a_stages = (res_bytes + 3 * kHaloABytes <= budget) ? 3 : 2;
    { static int cap = -1; if (cap < 0) { const char* e = getenv("MSP_HALO_ASTAGES"); cap = e ? atoi(e) : 3; }
      if (cap >= 1 && cap < p.a_stages) p.a_stages = cap; }   // timing experiments: ring depth
  } else {
    p.a_stages = BN >= 256 ? 2 : 3;
    long long nb = (budget - (long long)p.a_stages * kHaloABytes) / bt;
    if (nb > kHaloMaxB) nb = kHaloMaxB;
    if (nb < 2) return false;
    p.b_stages = (int)nb;
  }
  for (int t = 0; t < p.ntaps; ++t) {
    p.tap_dh[t] = (int8_t)(p.tap_dh[t] - dh0);
    p.tap_dw[t] = (int8_t)(p.tap_dw[t] - dw0);
  }
  p.halo = 1; p.org_dh = dh0; p.org_dw = dw0; p.halo_box_h = box_h;
  p.bw = bw; p.bw_valid = bwv; p.bh = bh; p.bn = 1; p.rows = bw * bh;
  p.tiles_w = msp_cdiv(OW, bwv); p.tiles_h = msp_cdiv(OH, bh); p.tiles_n = N;
  p.OWs = OW; p.OHs = OH; p.N = N;
  box->bw = bw; box->bh = box_h; box->bn = 1;
  *halo_w = bw;
  return true;
}

// narrow tiles (16 / 32 output channels): several pixel tiles per TMEM accumulator set (halo_multi_epilogue).
// MSP_CONV_MULTI=0 keeps one tile per hand-off.
bool multi_policy() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MSP_CONV_MULTI"); on = e ? atoi(e) : 1; }
  return on != 0;
}

template <int BN_, bool MULTI = false>
int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, TapGemmParams& p, cudaStream_t st) {
  using Cfg = TapGemmCfg<BN_>;
  const int smem = p.a_stages * kHaloABytes + p.b_stages * Cfg::kBTileBytes + Cfg::kStageBufs * kATileBytes +
                   Cfg::kScratchBytes + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(tapgemm_halo_kernel<BN_, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        232448 - 1024));
    attr_smem = 232448;
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("MSP_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  p.tiles_co = msp_cdiv(p.Kout, BN_);
  const long long total = (long long)p.tiles_m * p.tiles_co;
  MSP_REQUIRE(total < (1ll << 31), "conv: too many tiles");
  p.total_tiles = (int)total;
  const int sms = msp_num_sms();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  MSP_CHECK_CUDA(msp_launch_pdl(tapgemm_halo_kernel<BN_, MULTI>, dim3(grid), dim3(kTapThreads), (size_t)smem, st, tmA, tmB, p));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  g_last_kernel = BN_ == 256 ? "tapgemm_halo_kernel<256>" : BN_ == 128 ? "tapgemm_halo_kernel<128>"
                  : BN_ == 64 ? "tapgemm_halo_kernel<64>"
                  : BN_ == 32 ? (MULTI ? "tapgemm_halo_kernel<32, multi>" : "tapgemm_halo_kernel<32>")
                              : (MULTI ? "tapgemm_halo_kernel<16, multi>" : "tapgemm_halo_kernel<16>");
  return MSP_OK;
}

int dispatch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, TapGemmParams& p, cudaStream_t st) {
  const bool multi = multi_policy() && p.Kout <= 32 && p.bn == 1;
  switch (bn_tile_for(p.Kout)) {
    case 16: return multi ? launch_halo<16, true>(tmA, tmB, p, st) : launch_halo<16>(tmA, tmB, p, st);
    case 32: return multi ? launch_halo<32, true>(tmA, tmB, p, st) : launch_halo<32>(tmA, tmB, p, st);
    case 64: return launch_halo<64>(tmA, tmB, p, st);
    case 128: return launch_halo<128>(tmA, tmB, p, st);
    default: return launch_halo<256>(tmA, tmB, p, st);
  }
}

// Tensor map over an NHWC bf16 tensor (C, W, H, N) with a (64, bw*s, bh*s, bn) box (s = element stride).
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, int cs, Box b, int s) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)cs * 2, (uint64_t)W * cs * 2, (uint64_t)H * W * cs * 2};
  uint32_t box[4] = {64, (uint32_t)(b.bw * s), (uint32_t)(b.bh * s), (uint32_t)b.bn};
  uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
  return msp_encode_tmap_bf16(m, base, 4, dims, strides, box, es, 128);
}
// Same over a SUB-GRID view: base already points at the first element, (W, H) are the view's extents in elements of the
// parent tensor, row / image strides are the parent's (the folded up-conv reads the parity class (a, b) of dY).
int make_act_map_view(CUtensorMap* m, const void* base, int C, int W, int H, int N, int cs, long long row_stride,
                      long long img_stride, Box b, int s) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)cs * 2, (uint64_t)row_stride * 2, (uint64_t)img_stride * 2};
  uint32_t box[4] = {64, (uint32_t)(b.bw * s), (uint32_t)(b.bh * s), (uint32_t)b.bn};
  uint32_t es[4] = {1, (uint32_t)s, (uint32_t)s, 1};
  return msp_encode_tmap_bf16(m, base, 4, dims, strides, box, es, 128);
}
// Row-window map over the W-padded input [N][H][Wp][cpp]: dim 0 = 64 contiguous elements (the KW taps of
// one filter row), dim 1 = output column (stride = conv stride x pixel, OVERLAPPING windows), dim 2 = input
// row (element stride = conv stride), dim 3 = image.
int make_rowwin_map(CUtensorMap* m, const void* base, int cpp, int Wp, int H, int N, int OW, int s,
                    Box b) {
  uint64_t dims[4] = {64, (uint64_t)OW, (uint64_t)H, (uint64_t)N};
  uint64_t strides[3] = {(uint64_t)s * cpp * 2, (uint64_t)Wp * cpp * 2, (uint64_t)H * Wp * cpp * 2};
  uint32_t box[4] = {64, (uint32_t)b.bw, (uint32_t)(b.bh * s), (uint32_t)b.bn};
  uint32_t es[4] = {1, 1, (uint32_t)s, 1};
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
  if (d->win_px != 0) {
    MSP_REQUIRE(d->win_px == 4 || d->win_px == 8, "conv: row-window width must be 4 or 8 pixels");
    MSP_REQUIRE(d->C * d->win_px == 64 && d->x_cs == d->C,
                "conv: row-window mode needs C*win_px == 64 dense channels (C=%d)", d->C);
    MSP_REQUIRE(d->KW <= d->win_px, "conv: filter width %d exceeds the %d-pixel window", d->KW, d->win_px);
    MSP_REQUIRE(d->Wp >= d->stride * (d->Wo - 1) + d->win_px,
                "conv: padded row pitch Wp=%d too small for the last window", d->Wp);
  }
  return MSP_OK;
}

inline bool is_flat(const msp_conv_desc* d) {
  return d->win_px == 0 && d->KH * d->KW == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 &&
         d->Ho == d->H && d->Wo == d->W;
}

}  // namespace

extern "C" const char* msp_conv_last_kernel(void) { return g_last_kernel; }

extern "C" int msp_conv_set_policy(int pair, int halo) {
  MSP_REQUIRE(pair >= -1 && pair <= 2 && halo >= -1 && halo <= 2, "conv_set_policy: values are -1 (default) .. 2");
  g_pair_policy = pair;
  g_halo_policy = halo;
  return MSP_OK;
}

extern "C" int msp_pack_weights(const float* w, int K, int C, int KH, int KW, int Cpad, int Kpad,
                                void* w_fprop, void* w_dgrad, void* stream) {
  MSP_REQUIRE(w && (w_fprop || w_dgrad), "pack_weights: null pointer");
  MSP_REQUIRE(Cpad >= C && Cpad % 8 == 0 && Kpad >= K && Kpad % 8 == 0,
              "pack_weights: padded sizes must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = KH * KW;
  if (w_fprop && w_dgrad) {
    const long long total = (long long)K * taps * Cpad + (long long)Cpad * taps * Kpad;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    MSP_CHECK_CUDA(msp_launch_pdl(pack_w_both_kernel, dim3(blocks), dim3(256), 0, st, w, K, C, taps, Cpad, Kpad, (__nv_bfloat16*)w_fprop,
                                               (__nv_bfloat16*)w_dgrad));
    MSP_CHECK_LAUNCH();
    msp_count_launch(1);
    return MSP_OK;
  }
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

extern "C" int msp_pack_weights_batched(const long long* items_dev, int n_items, int total_blocks, void* stream) {
  MSP_REQUIRE(items_dev && n_items >= 1 && total_blocks >= n_items, "pack_weights_batched: bad arguments");
  pack_w_batched_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(items_dev, n_items);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_pack_weights_rowwin(const float* w, int K, int C, int KH, int KW, int win_px,
                                       void* w_rowwin, void* stream) {
  MSP_REQUIRE(w && w_rowwin, "pack_weights_rowwin: null pointer");
  MSP_REQUIRE((win_px == 4 || win_px == 8) && KW <= win_px && C <= 64 / win_px,
              "pack_weights_rowwin: C=%d KW=%d do not fit a %d-pixel window", C, KW, win_px);
  const long long total = (long long)K * KH * 64;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_w_rowwin_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, K, C, KH, KW, 64 / win_px,
                                                                 (__nv_bfloat16*)w_rowwin);
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
  if (msp_narrow_fprop_ok(d)) {   // 16 / 32-channel full-resolution layers: warp-level MMA kernel (msp_narrow.cu)
    rc = msp_narrow_fprop(d, x, w_fprop, bias, y, ch_sum, ch_sqsum, stream);
    g_last_kernel = "narrow_fprop_kernel";
    return rc;
  }
  const int taps = d->KH * d->KW;
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = is_flat(d);
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
    if (d->win_px)
      rc = make_rowwin_map(&tmA, x, d->C, d->Wp, d->H, d->N, d->Wo, d->stride, b);
    else
      rc = make_act_map(&tmA, x, d->C, d->W, d->H, d->N, d->x_cs, b, d->stride);
    p.OWs = d->Wo; p.OHs = d->Ho; p.N = d->N;
    p.y_n_stride = (long long)d->Ho * d->Wo * d->y_cs;
    p.y_h_stride = (long long)d->Wo * d->y_cs;
    p.y_w_stride = d->y_cs;
  }
  if (rc) return rc;
  const int bn_tile = bn_tile_for(d->K);
  const WMapArgs wm = d->win_px ? WMapArgs{w_fprop, 64, d->KH, d->K} : WMapArgs{w_fprop, d->C, taps, d->K};
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
  p.tiles_w = msp_cdiv(p.OWs, b.bw); p.tiles_h = msp_cdiv(p.OHs, b.bh); p.tiles_n = msp_cdiv(p.N, b.bn);
  p.sxw = (flat || d->win_px) ? 1 : d->stride;
  p.sxh = flat ? 1 : d->stride;
  p.Kout = d->K; p.relu = d->relu;
  p.y = (__nv_bfloat16*)y; p.y_off = 0;
  p.bias = bias; p.ch_sum = ch_sum; p.ch_sqsum = ch_sqsum;
  if (d->stat_rows > 0 && ch_sum != nullptr) {
    MSP_REQUIRE(d->stat_rows >= msp_num_sms() && ch_sqsum == ch_sum + d->K,
                "conv_fprop: deterministic statistics need a [rows >= %d][2][K] workspace (ch_sqsum = ch_sum + K)", msp_num_sms());
    p.stat_row = 2ll * d->K;
  }
  if (d->win_px) {
    p.C = 64; p.ntaps = d->KH;
    for (int r = 0; r < d->KH; ++r) {
      p.tap_dh[r] = (int8_t)(r - d->pad_t);
      p.tap_dw[r] = 0;
      p.tap_w[r] = (uint8_t)r;
    }
  } else {
    p.C = d->C; p.ntaps = taps;
    for (int r = 0; r < d->KH; ++r)
      for (int q = 0; q < d->KW; ++q) {
        const int t = r * d->KW + q;
        p.tap_dh[t] = (int8_t)(r - d->pad_t);
        p.tap_dw[t] = (int8_t)(q - d->pad_l);
        p.tap_w[t] = (uint8_t)t;
      }
  }
  if (!flat && !d->win_px && d->stride == 1) {
    Box hb;
    int hw = 0;
    if (plan_halo(p, d->Wo, d->Ho, d->N, bn_tile, &hb, &hw)) {
      rc = make_act_map(&tmA, x, d->C, d->W, d->H, d->N, d->x_cs, hb, 1);
      if (rc) return rc;
      rc = make_w_map(&tmB, wm.base, wm.inner, wm.taps, wm.rows, bn_tile);
      if (rc) return rc;
      return dispatch_halo(tmA, tmB, p, (cudaStream_t)stream);
    }
  }
  return dispatch_tapgemm(tmA, wm, p, (cudaStream_t)stream);
}

namespace {
int dgrad_impl(const msp_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, int accumulate,
               const float* bias, int relu, void* stream);
}
extern "C" int msp_conv_dgrad(const msp_conv_desc* d, const void* dy, const void* w_dgrad, void* dx,
                              int accumulate, void* stream) {
  return dgrad_impl(d, dy, w_dgrad, dx, accumulate, nullptr, 0, stream);
}
// nn.ConvTranspose2d forward IS the data gradient of the convolution it transposes: y = dgrad(x) (+ bias, ReLU in the
// same epilogue).  `d` describes that convolution: (N, H, W, C) = the OUTPUT of the transposed conv, (Ho, Wo, K) = its
// input; w_dgrad = msp_pack_weights' [C][tap][K] copy of the (K = in_channels, C = out_channels, kh, kw) weight.
extern "C" int msp_conv_transpose_fprop(const msp_conv_desc* d, const void* x, const void* w_dgrad, const float* bias,
                                        int relu, void* y, void* stream) {
  return dgrad_impl(d, x, w_dgrad, y, 0, bias, relu, stream);
}
namespace {
// Helper streams for the parity classes of a stride-2 data gradient: the 4 classes are independent launches (disjoint output
// pixels); on small feature maps each fills a fraction of the GPU for ~10 us, so they are forked onto three helper streams
// and joined again (events: capturable, the fork / join become edges of the step's CUDA graph) instead of running back
// to back (the U-Net's W_s / stride-2 3x3 layers at batch 24: 44-52 us for four launches of ~11 us).
struct ForkJoin {
  cudaStream_t side[3];
  cudaEvent_t fork, join[3];
  int device;
  bool ok;
};
ForkJoin* fork_join() {
  static ForkJoin fj[16];
  static bool init[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  ForkJoin& f = fj[dev];
  if (!init[dev]) {
    init[dev] = true;
    f.device = dev;
    f.ok = cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3 && f.ok; ++i)
      f.ok = cudaStreamCreateWithFlags(&f.side[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&f.join[i], cudaEventDisableTiming) == cudaSuccess;
  }
  return f.ok ? &f : nullptr;
}
bool dgrad_fork_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MSP_DGRAD_FORK");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
}

int dgrad_impl(const msp_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, int accumulate,
               const float* bias, int relu, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(dy && w_dgrad && dx, "conv_dgrad: null pointer");
  MSP_REQUIRE(d->win_px == 0, "conv_dgrad: not available in row-window mode (the network input needs no gradient)");
  if (!accumulate && msp_narrow_dgrad_ok(d)) {
    rc = msp_narrow_dgrad(d, dy, w_dgrad, bias, relu, dx, stream);
    g_last_kernel = "narrow_fprop_kernel";
    return rc;
  }
  const int taps = d->KH * d->KW;
  const int s = d->stride;
  cudaStream_t st0 = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  // weights [Cpad = d->C][taps][Kpad = d->K]: contraction over K (dy channels), outputs = C
  const WMapArgs wm{w_dgrad, d->K, taps, d->C};
  // stride 2 on a small feature map: the parity classes run side by side (see ForkJoin)
  ForkJoin* fj = nullptr;
  if (s == 2 && dgrad_fork_enabled() &&
      (long long)d->N * ((d->H + 1) / 2) * ((d->W + 1) / 2) <= 128ll * 8 * msp_num_sms())
    fj = fork_join();
  if (fj != nullptr) {
    MSP_CHECK_CUDA(cudaEventRecord(fj->fork, st0));
    for (int i = 0; i < 3; ++i) MSP_CHECK_CUDA(cudaStreamWaitEvent(fj->side[i], fj->fork, 0));
  }
  struct Joiner {   // joins the helper streams on every exit path
    ForkJoin* fj;
    cudaStream_t st0;
    ~Joiner() {
      if (fj == nullptr) return;
      for (int i = 0; i < 3; ++i) {
        cudaEventRecord(fj->join[i], fj->side[i]);
        cudaStreamWaitEvent(st0, fj->join[i], 0);
      }
    }
  } joiner{fj, st0};
  for (int ph = 0; ph < s; ++ph)
    for (int pw = 0; pw < s; ++pw) {
      const int cls = ph * s + pw;
      cudaStream_t st = (fj != nullptr && cls > 0) ? fj->side[cls - 1] : st0;
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
      const bool flat = is_flat(d);
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
      p.sxw = 1; p.sxh = 1; p.C = d->K; p.ntaps = nt; p.Kout = d->C; p.relu = relu; p.accumulate = accumulate;
      p.bias = bias;
      p.y = (__nv_bfloat16*)dx;
      if (!flat && s == 1) {
        Box hb;
        int hw = 0;
        if (plan_halo(p, d->W, d->H, d->N, bn_tile_for(d->C), &hb, &hw)) {
          rc = make_act_map(&tmA, dy, d->K, d->Wo, d->Ho, d->N, d->y_cs, hb, 1);
          if (rc) return rc;
          rc = make_w_map(&tmB, wm.base, wm.inner, wm.taps, wm.rows, bn_tile_for(d->C));
          if (rc) return rc;
          rc = dispatch_halo(tmA, tmB, p, st);
          if (rc) return rc;
          continue;
        }
      }
      rc = dispatch_tapgemm(tmA, wm, p, st);
      if (rc) return rc;
    }
  return MSP_OK;
}
}  // namespace

namespace {
struct WgradPlan {
  Box b;
  int OW, OH, N, tiles_m, nsub, csub, tsub, cchunks, gx, gy, splits, Cw, ntaps;
};
int plan_wgrad(const msp_conv_desc* d, WgradPlan* pl, bool force_2d = false) {
  const bool flat = is_flat(d) && !force_2d;
  if (flat) {
    const long long P = (long long)d->N * d->H * d->W;
    MSP_REQUIRE(P < (1ll << 31), "conv_wgrad: too many pixels");
    pl->OW = (int)P; pl->OH = 1; pl->N = 1;
  } else {
    pl->OW = d->Wo; pl->OH = d->Ho; pl->N = d->N;
  }
  pl->b = pick_box(pl->OW, pl->OH, pl->N, kWgRows);
  pl->tiles_m = msp_cdiv(pl->OW, pl->b.bw) * msp_cdiv(pl->OH, pl->b.bh) * msp_cdiv(pl->N, pl->b.bn);
  pl->Cw = d->win_px ? 64 : d->C;
  pl->ntaps = d->win_px ? d->KH : d->KH * d->KW;
  // accumulator = 128 output channels x (csub 64-channel chunks x tsub taps) <= 256 columns: wide MMA instructions
  // even for narrow layers (a 128x64x16 tcgen05.mma costs about as much as a 128x256x16 one)
  const int c64 = msp_cdiv(pl->Cw, 64);
  pl->csub = c64 < kWgMaxSub ? c64 : kWgMaxSub;
  pl->tsub = kWgMaxSub / pl->csub;
  if (pl->tsub > pl->ntaps) pl->tsub = pl->ntaps;
  pl->nsub = pl->csub * pl->tsub;
  pl->cchunks = msp_cdiv(c64, pl->csub);
  pl->gx = msp_cdiv(pl->ntaps, pl->tsub) * pl->cchunks;
  pl->gy = msp_cdiv(d->K, 128);
  const int base = pl->gx * pl->gy;
  int splits = msp_num_sms() / base;
  if (splits < 1) splits = 1;
  if (splits > pl->tiles_m) splits = pl->tiles_m;
  {
    // optional cap on the bytes of split-K partials per layer (MSP_WGRAD_PART_MB; 0 = none): every partial is written by
    // the wgrad kernel and read again by the unpack, which runs exposed between the backward pass and the optimizer
    static long long cap_bytes = -1;
    if (cap_bytes < 0) {
      const char* e = getenv("MSP_WGRAD_PART_MB");
      cap_bytes = e ? (long long)atoi(e) << 20 : 0;
    }
    if (cap_bytes > 0) {
      const long long wbytes = (long long)d->K * pl->ntaps * pl->Cw * 4;
      long long smax = cap_bytes / (wbytes > 0 ? wbytes : 1);
      if (smax < 1) smax = 1;
      if (splits > smax) splits = (int)smax;
    }
    // ... or relative to the activations the launch reads anyway: partial bytes <= alpha x (x + dy bytes).  At batch 24 the
    // U-Net's mid layers wrote 8-74 partials of their whole weight tensor (1.1 GB per step, written during the backward pass
    // and read again by the unpack); at batch 256 the activations dominate and the rule changes nothing.
    static double alpha = -1.0;
    if (alpha < 0.0) {
      const char* e = getenv("MSP_WGRAD_PART_ALPHA");
      alpha = e ? atof(e) : 1.0;   // measured: cfg3 9.60 -> 9.38 ms (0.5: 9.27), cfg2 unchanged (a 6 MB byte cap: cfg2 +1.7 % slower)
    }
    if (alpha > 0.0) {
      const double wbytes = (double)d->K * pl->ntaps * pl->Cw * 4.0;
      const double abytes = 2.0 * (double)pl->OW * pl->OH * pl->N * ((double)d->C + d->K);
      long long smax = (long long)(alpha * abytes / (wbytes > 0 ? wbytes : 1.0));
      if (smax < 1) smax = 1;
      if (splits > smax) splits = (int)smax;
    }
  }
  if (splits > 65535) splits = 65535;
  pl->splits = splits;
  return MSP_OK;
}
}  // namespace

extern "C" int msp_conv_wgrad_splits(const msp_conv_desc* d) {
  int rc = check_desc(d);
  if (rc) return rc;
  const int narrow = msp_narrow_wgrad_splits(d);   // 16 / 32-channel full-resolution layers: warp-level MMA kernel
  if (narrow > 0) return narrow;
  WgradPlan pl;
  rc = plan_wgrad(d, &pl);
  if (rc) return rc;
  return pl.splits;
}

namespace {
int wgrad_impl(const msp_conv_desc* d, const void* x, const void* dy, float* dw_partials, int dy_es, int dy_oh,
               int dy_ow, void* stream);
}
extern "C" int msp_conv_wgrad(const msp_conv_desc* d, const void* x, const void* dy,
                              float* dw_partials, void* stream) {
  return wgrad_impl(d, x, dy, dw_partials, 1, 0, 0, stream);
}
// One output-parity class of the folded up-convolution: `d` = the class's convolution on the LOW-RES input (KH = 1 + a,
// KW = 1 + b, pad 0, stride 1, Ho = H, Wo = W) whose output gradient is the parity class (a, b) of the full-resolution
// dy (2H x 2W, pixel stride d->y_cs): dY(i, j) = dy[2i + a][2j + b].
extern "C" int msp_upconv2x_wgrad_class(const msp_conv_desc* d, const void* x, const void* dy, int a, int b,
                                        float* dw_partials, void* stream) {
  MSP_REQUIRE(d && (a == 0 || a == 1) && (b == 0 || b == 1) && d->KH == 1 + a && d->KW == 1 + b && d->stride == 1 &&
                  d->pad_t == 0 && d->pad_l == 0 && d->Ho == d->H && d->Wo == d->W && d->win_px == 0,
              "upconv2x_wgrad_class: descriptor does not describe parity class (%d,%d)", a, b);
  return wgrad_impl(d, x, dy, dw_partials, 2, a, b, stream);
}
namespace {
int wgrad_impl(const msp_conv_desc* d, const void* x, const void* dy, float* dw_partials, int dy_es, int dy_oh,
               int dy_ow, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && dy && dw_partials, "conv_wgrad: null pointer");
  if (dy_es == 1 && msp_narrow_wgrad_splits(d) > 0) {
    rc = msp_narrow_wgrad(d, x, dy, dw_partials, stream);
    g_last_kernel = "narrow_wgrad_kernel";
    return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    MSP_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kWgSmemBytes));
    attr_set = true;
  }
  WgradPlan pl;
  rc = plan_wgrad(d, &pl, dy_es != 1);
  if (rc) return rc;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const bool flat = is_flat(d) && dy_es == 1;
  CUtensorMap tmDY, tmX;
  const Box b = pl.b;
  p.dys = dy_es;
  if (dy_es != 1) {
    // dY = parity class (dy_oh, dy_ow) of a (dy_es*Ho x dy_es*Wo) tensor: view from its first element, parent strides
    const int Wf = dy_es * d->Wo, Hf = dy_es * d->Ho;
    const __nv_bfloat16* base = (const __nv_bfloat16*)dy + ((long long)dy_oh * Wf + dy_ow) * d->y_cs;
    rc = make_act_map_view(&tmDY, base, d->K, Wf - dy_ow, Hf - dy_oh, d->N, d->y_cs, (long long)Wf * d->y_cs,
                           (long long)Hf * Wf * d->y_cs, b, dy_es);
    if (rc) return rc;
    rc = make_act_map(&tmX, x, d->C, d->W, d->H, d->N, d->x_cs, b, d->stride);
    if (rc) return rc;
    p.sxw = p.sxh = d->stride;
  } else if (flat) {
    rc = make_act_map(&tmDY, dy, d->K, pl.OW, 1, 1, d->y_cs, b, 1);
    if (rc) return rc;
    rc = make_act_map(&tmX, x, d->C, pl.OW, 1, 1, d->x_cs, b, 1);
    if (rc) return rc;
    p.sxw = p.sxh = 1;
  } else {
    rc = make_act_map(&tmDY, dy, d->K, d->Wo, d->Ho, d->N, d->y_cs, b, 1);
    if (rc) return rc;
    if (d->win_px) {
      rc = make_rowwin_map(&tmX, x, d->C, d->Wp, d->H, d->N, d->Wo, d->stride, b);
      p.sxw = 1;
    } else {
      rc = make_act_map(&tmX, x, d->C, d->W, d->H, d->N, d->x_cs, b, d->stride);
      p.sxw = d->stride;
    }
    if (rc) return rc;
    p.sxh = d->stride;
  }
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
  p.tiles_w = msp_cdiv(pl.OW, b.bw); p.tiles_h = msp_cdiv(pl.OH, b.bh); p.tiles_n = msp_cdiv(pl.N, b.bn);
  p.tiles_m = pl.tiles_m;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.KW = d->KW;
  p.Cw = pl.Cw; p.Kout = d->K; p.ntaps = pl.ntaps; p.cchunks = pl.cchunks; p.nsub = pl.nsub; p.csub = pl.csub; p.tsub = pl.tsub;
  p.rowwin = d->win_px != 0;
  p.splits = pl.splits;
  p.split_stride = (long long)d->K * pl.ntaps * pl.Cw;
  p.dw = dw_partials;
  dim3 grid(pl.gx, pl.gy, pl.splits);
  MSP_CHECK_CUDA(msp_launch_pdl(wgrad_kernel, dim3(grid), dim3(kConvThreads), (size_t)kWgSmemBytes, st, tmDY, tmX, p));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  g_last_kernel = "wgrad_kernel";
  return MSP_OK;
}
}  // namespace

extern "C" int msp_upconv2x_wgrad_splits(const msp_conv_desc* d) {
  int rc = check_desc(d);
  if (rc) return rc;
  WgradPlan pl;
  rc = plan_wgrad(d, &pl, true);
  if (rc) return rc;
  return pl.splits;
}

extern "C" int msp_fold_upconv_weights(const float* w, int K, int C, float* w_folded, void* stream) {
  MSP_REQUIRE(w && w_folded && K > 0 && C > 0, "fold_upconv_weights: bad arguments");
  const long long KC = (long long)K * C;
  const int blocks = (int)((KC + 255) / 256 < 2048 ? (KC + 255) / 256 : 2048);
  fold_upconv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, KC, w_folded);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_unfold_upconv_wgrad(const float* g00, const float* g01, const float* g10, const float* g11, int K,
                                       int C, float* dw, int accumulate, void* stream) {
  MSP_REQUIRE(g00 && g01 && g10 && g11 && dw && K > 0 && C > 0, "unfold_upconv_wgrad: bad arguments");
  const long long KC = (long long)K * C;
  const int blocks = (int)((KC + 255) / 256 < 2048 ? (KC + 255) / 256 : 2048);
  unfold_upconv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g00, g01, g10, g11, KC, dw, accumulate);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

// y = ReLU?(upconv(x) + bias): four stride-1 tap-GEMMs on the low-res input, class (a, b) writing the output sub-grid
// (2i + a, 2j + b).  `d`: (N, H, W, C, x_cs) = the LOW-RES input, (Ho, Wo) = (2H, 2W), K, y_cs, relu.
extern "C" int msp_upconv2x_fprop(const msp_conv_desc* d, const void* x, const void* w_folded_fprop, const float* bias,
                                  void* y, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(x && w_folded_fprop && y, "upconv2x_fprop: null pointer");
  MSP_REQUIRE(d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->win_px == 0, "upconv2x_fprop: output must be 2H x 2W");
  cudaStream_t st = (cudaStream_t)stream;
  const WMapArgs wm{w_folded_fprop, d->C, 9, d->K};
  for (int a = 0; a < 2; ++a)
    for (int b2 = 0; b2 < 2; ++b2) {
      TapGemmParams p;
      memset(&p, 0, sizeof(p));
      const int cls = a * 2 + b2, base = cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? 3 : 5));
      int nt = 0;
      for (int dr = 0; dr <= a; ++dr)
        for (int dq = 0; dq <= b2; ++dq) {
          p.tap_dh[nt] = (int8_t)dr;
          p.tap_dw[nt] = (int8_t)dq;
          p.tap_w[nt] = (uint8_t)(base + nt);
          ++nt;
        }
      Box b = pick_box(d->W, d->H, d->N);
      CUtensorMap tmA, tmB;
      rc = make_act_map(&tmA, x, d->C, d->W, d->H, d->N, d->x_cs, b, 1);
      if (rc) return rc;
      p.OWs = d->W; p.OHs = d->H; p.N = d->N;
      p.y_n_stride = (long long)d->Ho * d->Wo * d->y_cs;
      p.y_h_stride = 2ll * d->Wo * d->y_cs;
      p.y_w_stride = 2ll * d->y_cs;
      p.y_off = ((long long)a * d->Wo + b2) * d->y_cs;
      p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
      p.tiles_w = msp_cdiv(p.OWs, b.bw); p.tiles_h = msp_cdiv(p.OHs, b.bh); p.tiles_n = msp_cdiv(p.N, b.bn);
      p.sxw = 1; p.sxh = 1; p.C = d->C; p.ntaps = nt; p.Kout = d->K; p.relu = d->relu;
      p.y = (__nv_bfloat16*)y;
      p.bias = bias;
      Box hb;
      int hw = 0;
      if (plan_halo(p, d->W, d->H, d->N, bn_tile_for(d->K), &hb, &hw)) {
        rc = make_act_map(&tmA, x, d->C, d->W, d->H, d->N, d->x_cs, hb, 1);
        if (rc) return rc;
        rc = make_w_map(&tmB, wm.base, wm.inner, wm.taps, wm.rows, bn_tile_for(d->K));
        if (rc) return rc;
        rc = dispatch_halo(tmA, tmB, p, st);
      } else {
        rc = dispatch_tapgemm(tmA, wm, p, st);
      }
      if (rc) return rc;
    }
  return MSP_OK;
}

// dx (low-res) = sum over the 9 folded taps of dy[2(i - dr) + a][2(j - dq) + b] * Wfold^T: ONE tap-GEMM whose A operand is
// dy read through a stride-2 tensor map (as a stride-2 forward convolution reads its input).
extern "C" int msp_upconv2x_dgrad(const msp_conv_desc* d, const void* dy, const void* w_folded_dgrad, void* dx,
                                  int accumulate, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(dy && w_folded_dgrad && dx, "upconv2x_dgrad: null pointer");
  MSP_REQUIRE(d->Ho == 2 * d->H && d->Wo == 2 * d->W && d->win_px == 0, "upconv2x_dgrad: dy must be 2H x 2W");
  TapGemmParams p;
  memset(&p, 0, sizeof(p));
  int nt = 0;
  for (int a = 0; a < 2; ++a)
    for (int b2 = 0; b2 < 2; ++b2)
      for (int dr = 0; dr <= a; ++dr)
        for (int dq = 0; dq <= b2; ++dq) {
          p.tap_dh[nt] = (int8_t)(a - 2 * dr);   // dy row 2(i - dr) + a = 2i + (a - 2 dr)
          p.tap_dw[nt] = (int8_t)(b2 - 2 * dq);
          p.tap_w[nt] = (uint8_t)nt;            // folded tap order == this enumeration
          ++nt;
        }
  Box b = pick_box(d->W, d->H, d->N);
  CUtensorMap tmA;
  rc = make_act_map(&tmA, dy, d->K, d->Wo, d->Ho, d->N, d->y_cs, b, 2);
  if (rc) return rc;
  p.OWs = d->W; p.OHs = d->H; p.N = d->N;
  p.y_n_stride = (long long)d->H * d->W * d->x_cs;
  p.y_h_stride = (long long)d->W * d->x_cs;
  p.y_w_stride = d->x_cs;
  p.bw = b.bw; p.bh = b.bh; p.bn = b.bn; p.rows = b.bw * b.bh * b.bn;
  p.tiles_w = msp_cdiv(p.OWs, b.bw); p.tiles_h = msp_cdiv(p.OHs, b.bh); p.tiles_n = msp_cdiv(p.N, b.bn);
  p.sxw = 2; p.sxh = 2; p.C = d->K; p.ntaps = nt; p.Kout = d->C; p.accumulate = accumulate;
  p.y = (__nv_bfloat16*)dx;
  const WMapArgs wm{w_folded_dgrad, d->K, 9, d->C};
  return dispatch_tapgemm(tmA, wm, p, (cudaStream_t)stream);
}

extern "C" int msp_unpack_wgrad(const msp_conv_desc* d, const float* dw_partials, int C_true,
                                float* dw_oihw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  MSP_REQUIRE(dw_partials && dw_oihw && C_true > 0 && C_true <= d->C, "unpack_wgrad: bad arguments");
  WgradPlan pl;
  rc = plan_wgrad(d, &pl);
  if (rc) return rc;
  const int narrow = msp_narrow_wgrad_splits(d);   // the narrow-channel kernel writes one partial per CTA
  if (narrow > 0) {
    msp_unpack_item it;
    memset(&it, 0, sizeof(it));
    it.partials = dw_partials; it.dst = dw_oihw; it.split_stride = (long long)d->K * pl.ntaps * pl.Cw;
    it.splits = narrow; it.K = d->K; it.C_true = C_true; it.taps = pl.ntaps; it.Cpad = pl.Cw;
    return msp_unpack_wgrad_batched(1, &it, stream);
  }
  const long long split_stride = (long long)d->K * pl.ntaps * pl.Cw;
  const long long total = (long long)d->K * C_true * d->KH * d->KW;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (d->win_px)
    unpack_wgrad_rowwin_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        dw_partials, pl.splits, split_stride, d->K, C_true, d->KH, d->KW, d->C, dw_oihw);
  else
    MSP_CHECK_CUDA(msp_launch_pdl(unpack_wgrad_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, dw_partials,
                                  pl.splits, split_stride, d->K, C_true, d->KH * d->KW, d->C, dw_oihw));
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_unpack_wgrad_batched(int n, const msp_unpack_item* items, void* stream) {
  MSP_REQUIRE(items != nullptr && n >= 1 && n <= kUnpackMaxItems, "unpack_wgrad_batched: 1..%d items per call (got %d)",
              kUnpackMaxItems, n);
  UnpackList L;
  memset(&L, 0, sizeof(L));
  int first = 0;
  for (int i = 0; i < n; ++i) {
    const msp_unpack_item& it = items[i];
    MSP_REQUIRE(it.partials && it.dst && it.splits >= 1 && it.K > 0 && it.C_true > 0, "unpack_wgrad_batched: bad item %d", i);
    MSP_REQUIRE(it.rowwin_KH > 0 ? (it.rowwin_KW > 0 && it.rowwin_cpp > 0) : (it.taps > 0 && it.Cpad >= it.C_true),
                "unpack_wgrad_batched: bad geometry in item %d", i);
    L.it[i] = it;
    const long long total = (long long)it.K * it.C_true * (it.rowwin_KH > 0 ? it.rowwin_KH * it.rowwin_KW : it.taps);
    long long nb = (it.rowwin_KH == 0 && it.splits >= kUnpackWideSplits) ? (total + 31) / 32 : (total + 2047) / 2048;
    nb = nb < 1 ? 1 : (nb > 1024 ? 1024 : nb);  // (a 128-block cap left the largest layers of a per-bucket flush on 8 warps per SM)
    L.first_block[i] = first;
    first += (int)nb;
  }
  L.first_block[n] = first;
  L.n = n;
  unpack_wgrad_batched_kernel<<<(unsigned)first, 256, 0, (cudaStream_t)stream>>>(L);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
