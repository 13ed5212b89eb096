// Single-pass metric counters (integer, bit-exact).
//
// Replaces (reference): ConfusionMatrix.calculate_batch (metrics/metrics.py:61-95; eight full-size
// boolean temporaries + six reductions), MultiClassConfusionMatrix.calculate_batch
// (metrics/multiclass_metrics.py:90-107; GPU -> CPU numpy -> sklearn.metrics.confusion_matrix every
// batch) and Top5Accuracy.calculate_batch (metrics/multiclass_metrics.py:424-446).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// counts[0..5] = TP, TN, FP, FN, positives (class count), NaN targets.  grid: (chunks, rows) where a row
// is one (n, c) plane of HW elements; out index = per_channel ? c : 0.
template <typename TT>
__global__ void confusion_binary_kernel(const float* __restrict__ pred, const TT* __restrict__ target,
                                        int C, long long HW, float thr, int per_channel,
                                        unsigned long long* __restrict__ out) {
  const long long row = blockIdx.y;
  const float* p = pred + row * HW;
  const TT* t = target + row * HW;
  unsigned tp = 0, tn = 0, fp = 0, fn = 0, nan = 0;
  // per-thread strips are far below 2^32 elements: (HW / gridDim.x / blockDim.x)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW;
       i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const TT tv = t[i];
    const bool yp = (tv == (TT)1);
    const bool hp = pv >= thr;  // NaN prediction -> negative, like torch
    tp += (yp && hp);
    tn += (!yp && !hp);
    fp += (!yp && hp);
    fn += (yp && !hp);
    if (sizeof(TT) == 4) nan += (tv != tv);
  }
  __shared__ unsigned red[5][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tp = warp_sum_u32(tp); tn = warp_sum_u32(tn); fp = warp_sum_u32(fp); fn = warp_sum_u32(fn);
  nan = warp_sum_u32(nan);
  if (lane == 0) {
    red[0][warp] = tp; red[1][warp] = tn; red[2][warp] = fp; red[3][warp] = fn; red[4][warp] = nan;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    unsigned long long a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[threadIdx.x][w];
    unsigned long long* o = out + (per_channel ? (row % C) * 6 : 0);
    if (a) {
      if (threadIdx.x < 4) atomicAdd(o + threadIdx.x, a);
      else atomicAdd(o + 5, a);
    }
    if (threadIdx.x == 0) {
      // positives = TP + FN
      unsigned long long pos = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) pos += (unsigned long long)red[0][w] + red[3][w];
      if (pos) atomicAdd(o + 4, pos);
    }
  }
}

// 4 elements per thread and iteration (16-byte loads); see confusion_binary_kernel for the semantics.  HW % 4 == 0 and
// 16-byte aligned planes (the launcher checks and otherwise keeps the scalar kernel).
template <typename TT>
__global__ void __launch_bounds__(256)
confusion_binary_vec4_kernel(const float* __restrict__ pred, const TT* __restrict__ target, int C, long long HW4,
                             float thr, int per_channel, unsigned long long* __restrict__ out) {
  const long long row = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(pred + row * HW4 * 4);
  const TT* t = target + row * HW4 * 4;
  unsigned tp = 0, tn = 0, fp = 0, fn = 0, nan = 0;
#pragma unroll 2
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW4; i += (long long)gridDim.x * blockDim.x) {
    const float4 pq = ld_nc_f4(p + i);
    bool yp[4];
    if constexpr (sizeof(TT) == 4) {
      const float4 tq = ld_nc_f4(reinterpret_cast<const float4*>(t) + i);
      const float tv[4] = {tq.x, tq.y, tq.z, tq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        yp[e] = tv[e] == 1.f;
        nan += (tv[e] != tv[e]);
      }
    } else {
      const longlong2 a = reinterpret_cast<const longlong2*>(t)[2 * i], b = reinterpret_cast<const longlong2*>(t)[2 * i + 1];
      yp[0] = a.x == 1; yp[1] = a.y == 1; yp[2] = b.x == 1; yp[3] = b.y == 1;
    }
    const float pv[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool hp = pv[e] >= thr;  // NaN prediction -> negative, like torch
      tp += (yp[e] && hp);
      tn += (!yp[e] && !hp);
      fp += (!yp[e] && hp);
      fn += (yp[e] && !hp);
    }
  }
  __shared__ unsigned red[5][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tp = warp_sum_u32(tp); tn = warp_sum_u32(tn); fp = warp_sum_u32(fp); fn = warp_sum_u32(fn);
  nan = warp_sum_u32(nan);
  if (lane == 0) {
    red[0][warp] = tp; red[1][warp] = tn; red[2][warp] = fp; red[3][warp] = fn; red[4][warp] = nan;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    unsigned long long a = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[threadIdx.x][w];
    unsigned long long* o = out + (per_channel ? (row % C) * 6 : 0);
    if (a) {
      if (threadIdx.x < 4) atomicAdd(o + threadIdx.x, a);
      else atomicAdd(o + 5, a);
    }
    if (threadIdx.x == 0) {
      unsigned long long pos = 0;  // positives = TP + FN
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) pos += (unsigned long long)red[0][w] + red[3][w];
      if (pos) atomicAdd(o + 4, pos);
    }
  }
}

// argmax over the class dimension with torch semantics: first maximal index; a NaN is maximal
// (first NaN wins).
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& bi) {
  const bool best_nan = best != best;
  if (!best_nan && (v > best || v != v)) { best = v; bi = c; }
}

// pixel-parallel variant: thread per (n, hw); loads are coalesced for every class plane.
constexpr int kCmSmemC = 16;
__global__ void confusion_multiclass_pix_kernel(const float* __restrict__ pred,
                                                const void* __restrict__ target, int onehot, int C,
                                                long long HW, long long P,
                                                unsigned long long* __restrict__ cm) {
  __shared__ unsigned hist[kCmSmemC * kCmSmemC];
  const bool use_smem = C <= kCmSmemC;
  if (use_smem) {
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
  }
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long n = pix / HW, hw = pix - n * HW;
    const float* pp = pred + n * C * HW + hw;
    float best = pp[0];
    int bi = 0;
    for (int c = 1; c < C; ++c) argmax_step(pp[(long long)c * HW], c, best, bi);
    long long t;
    if (onehot) {
      const float* tp = reinterpret_cast<const float*>(target) + n * C * HW + hw;
      float tb = tp[0];
      int ti = 0;
      for (int c = 1; c < C; ++c) argmax_step(tp[(long long)c * HW], c, tb, ti);
      t = ti;
    } else {
      t = reinterpret_cast<const long long*>(target)[pix];
    }
    if (t >= 0 && t < C) {  // sklearn ignores labels outside `labels`
      if (use_smem) atomicAdd(&hist[(int)t * C + bi], 1u);
      else atomicAdd(cm + t * C + bi, 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (hist[i]) atomicAdd(cm + i, (unsigned long long)hist[i]);
  }
}

// 4 pixels per thread and iteration for int64 label maps (HW % 4 == 0, 16-byte aligned planes): 16-byte loads of every
// class plane, the shared-memory histogram as above.
__global__ void __launch_bounds__(256)
confusion_multiclass_pix4_kernel(const float* __restrict__ pred, const long long* __restrict__ target, int C,
                                 long long HW4, long long P4, unsigned long long* __restrict__ cm) {
  __shared__ unsigned hist[kCmSmemC * kCmSmemC];
  const bool use_smem = C <= kCmSmemC;
  if (use_smem) {
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
  }
#pragma unroll 2
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < P4; g += (long long)gridDim.x * blockDim.x) {
    const long long n = g / HW4, hw4 = g - n * HW4;
    const float4* pp = reinterpret_cast<const float4*>(pred) + n * C * HW4 + hw4;
    const longlong2 ta = reinterpret_cast<const longlong2*>(target)[2 * g], tb = reinterpret_cast<const longlong2*>(target)[2 * g + 1];
    const float4 p0 = ld_nc_f4(pp);
    float best[4] = {p0.x, p0.y, p0.z, p0.w};
    int bi[4] = {0, 0, 0, 0};
#pragma unroll 4
    for (int c = 1; c < C; ++c) {
      const float4 q = ld_nc_f4(pp + (long long)c * HW4);
      argmax_step(q.x, c, best[0], bi[0]);
      argmax_step(q.y, c, best[1], bi[1]);
      argmax_step(q.z, c, best[2], bi[2]);
      argmax_step(q.w, c, best[3], bi[3]);
    }
    const long long t[4] = {ta.x, ta.y, tb.x, tb.y};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (t[e] >= 0 && t[e] < C) {  // sklearn ignores labels outside `labels`
        if (use_smem) atomicAdd(&hist[(int)t[e] * C + bi[e]], 1u);
        else atomicAdd(cm + t[e] * C + bi[e], 1ull);
      }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
      if (hist[i]) atomicAdd(cm + i, (unsigned long long)hist[i]);
  }
}

// row-parallel variant (HW == 1, many classes: the ImageNet head): one warp per row.
__global__ void confusion_multiclass_row_kernel(const float* __restrict__ pred,
                                                const void* __restrict__ target, int onehot, int C,
                                                long long N, unsigned long long* __restrict__ cm) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  auto row_argmax = [&](const float* z) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    bool has = false;
    for (int c = lane; c < C; c += 32) {
      const float v = z[c];
      if (!has) { best = v; bi = c; has = true; }
      else argmax_step(v, c, best, bi);
    }
    // combine lanes: NaN beats numbers; ties -> lower index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const bool oh = __shfl_xor_sync(0xffffffffu, (int)has, o) != 0;
      if (!oh) continue;
      if (!has) { best = ob; bi = oi; has = true; continue; }
      const bool bn = best != best, on = ob != ob;
      bool take;
      if (bn || on) take = on && (!bn || oi < bi);
      else take = (ob > best) || (ob == best && oi < bi);
      if (take) { best = ob; bi = oi; }
    }
    return bi;
  };
  const int p = row_argmax(pred + row * C);
  long long t;
  if (onehot) t = row_argmax(reinterpret_cast<const float*>(target) + row * C);
  else t = reinterpret_cast<const long long*>(target)[row];
  if (lane == 0 && t >= 0 && t < C) atomicAdd(cm + t * C + p, 1ull);
}

// hit iff fewer than k classes rank before the label: rank = #{s_c > s_L} + #{c < L : s_c == s_L}
__global__ void topk_hits_pix_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                     int C, long long HW, long long P, int k,
                                     unsigned long long* __restrict__ hits) {
  unsigned mine = 0;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < P;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long n = pix / HW, hw = pix - n * HW;
    const long long L = label[pix];
    if (L < 0 || L >= C) continue;
    const float* pp = pred + n * C * HW + hw;
    const float sl = pp[L * HW];
    int rank = 0;
    for (int c = 0; c < C; ++c) {
      const float v = pp[(long long)c * HW];
      rank += (v > sl) || (v == sl && c < L);
    }
    mine += rank < k;
  }
  mine = warp_sum_u32(mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(hits, (unsigned long long)mine);
}
__global__ void topk_hits_row_kernel(const float* __restrict__ pred, const long long* __restrict__ label,
                                     int C, long long N, int k, unsigned long long* __restrict__ hits) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  const long long L = label[row];
  if (L < 0 || L >= C) return;
  const float* z = pred + row * C;
  const float sl = z[L];
  unsigned rank = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = z[c];
    rank += (v > sl) || (v == sl && c < L);
  }
  rank = warp_sum_u32(rank);
  if (lane == 0 && rank < (unsigned)k) atomicAdd(hits, 1ull);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int msp_confusion_binary(const float* pred, const void* target, int target_is_float, int N,
                                    int C, long long HW, float thr, int per_channel, long long* out,
                                    void* stream) {
  MSP_REQUIRE(out, "confusion_binary: null output");
  MSP_REQUIRE(N >= 0 && C >= 1 && HW >= 0, "confusion_binary: bad shape");
  MSP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(long long) * 6 * (per_channel ? C : 1), ST));
  const long long rows = (long long)N * C;
  if (rows == 0 || HW == 0) return MSP_OK;  // empty input: all-zero counts
  MSP_REQUIRE(pred && target, "confusion_binary: null pointer");
  MSP_REQUIRE(rows <= 65535, "confusion_binary: N*C=%lld > 65535", rows);
  long long chunks = (HW + 256 * 16 - 1) / (256 * 16);
  const long long cap = ((long long)msp_num_sms() * 8 + rows - 1) / rows;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)rows);
  const bool vec = HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
  if (vec) {
    // one 16-byte group per thread and iteration, ~4 blocks per SM over all rows
    long long vb = ((long long)msp_num_sms() * 4 + rows - 1) / rows, need = (HW / 4 + 255) / 256;
    if (vb > need) vb = need;
    if (vb < 1) vb = 1;
    dim3 vgrid((unsigned)vb, (unsigned)rows);
    if (target_is_float)
      confusion_binary_vec4_kernel<float><<<vgrid, 256, 0, ST>>>(pred, (const float*)target, C, HW / 4, thr, per_channel,
                                                                 (unsigned long long*)out);
    else
      confusion_binary_vec4_kernel<long long><<<vgrid, 256, 0, ST>>>(pred, (const long long*)target, C, HW / 4, thr,
                                                                     per_channel, (unsigned long long*)out);
  } else if (target_is_float)
    confusion_binary_kernel<float><<<grid, 256, 0, ST>>>(pred, (const float*)target, C, HW, thr,
                                                         per_channel, (unsigned long long*)out);
  else
    confusion_binary_kernel<long long><<<grid, 256, 0, ST>>>(pred, (const long long*)target, C, HW, thr,
                                                             per_channel, (unsigned long long*)out);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_confusion_multiclass(const float* pred, const void* target, int target_is_onehot,
                                        int N, int C, long long HW, long long* cm, void* stream) {
  MSP_REQUIRE(cm && C >= 1, "confusion_multiclass: bad arguments");
  MSP_CHECK_CUDA(cudaMemsetAsync(cm, 0, sizeof(long long) * C * C, ST));
  const long long P = (long long)N * HW;
  if (P <= 0) return MSP_OK;
  MSP_REQUIRE(pred && target, "confusion_multiclass: null pointer");
  if (HW == 1 && C > 32) {
    const long long threads = P * 32;
    confusion_multiclass_row_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ST>>>(
        pred, target, target_is_onehot, C, P, (unsigned long long*)cm);
  } else if (!target_is_onehot && HW % 4 == 0 &&
             ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0) {
    long long blocks = (P / 4 + 255) / 256;
    if (blocks > (long long)msp_num_sms() * 4) blocks = (long long)msp_num_sms() * 4;
    confusion_multiclass_pix4_kernel<<<(unsigned)blocks, 256, 0, ST>>>(pred, (const long long*)target, C, HW / 4, P / 4,
                                                                       (unsigned long long*)cm);
  } else {
    long long blocks = (P + 255) / 256;
    if (blocks > (long long)msp_num_sms() * 8) blocks = (long long)msp_num_sms() * 8;
    confusion_multiclass_pix_kernel<<<(unsigned)blocks, 256, 0, ST>>>(
        pred, target, target_is_onehot, C, HW, P, (unsigned long long*)cm);
  }
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_topk_hits(const float* pred, const int64_t* label, int N, int C, long long HW, int k,
                             long long* hits, void* stream) {
  MSP_REQUIRE(hits && C >= 1 && k >= 1, "topk_hits: bad arguments");
  MSP_CHECK_CUDA(cudaMemsetAsync(hits, 0, sizeof(long long), ST));
  const long long P = (long long)N * HW;
  if (P <= 0) return MSP_OK;
  MSP_REQUIRE(pred && label, "topk_hits: null pointer");
  if (HW == 1 && C > 32) {
    const long long threads = P * 32;
    topk_hits_row_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ST>>>(
        pred, (const long long*)label, C, P, k, (unsigned long long*)hits);
  } else {
    long long blocks = (P + 255) / 256;
    if (blocks > (long long)msp_num_sms() * 8) blocks = (long long)msp_num_sms() * 8;
    topk_hits_pix_kernel<<<(unsigned)blocks, 256, 0, ST>>>(pred, (const long long*)label, C, HW, P, k,
                                                           (unsigned long long*)hits);
  }
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
