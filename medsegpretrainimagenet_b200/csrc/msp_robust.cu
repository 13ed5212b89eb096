// Encoder-robustness distances: cosine, mean-squared and 1-Pearson distance of every row to its
// positive (same index) and negative (fixed permutation) partner, in ONE pass over the data.
//
// Replaces (reference): robustness/distance.py:3-10 (Pearson is a Python loop of torch.corrcoef
// calls) and the k0 construction + hinge of Robustness.__call__ (robustness/eval.py:16-28).
//
// The negative permutation perm = [1, 0, N-1, N-2, ..., 2] is an involution, so rows come in pairs
// {i, j = perm(i)}: loading q_i, q_j, k_i, k_j once yields pos(i), neg(i) = d(q_i, k_j), pos(j) and
// neg(j) = d(q_j, k_i).  Algorithmic traffic: 2 * N * D * 4 bytes for all three distances together.
// Twelve raw moments per pair are accumulated in fp64 (products of fp32 values are exact in fp64, so
// the centred Pearson moments do not suffer the fp32 cancellation of post-ReLU features).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

struct Moments {
  double s[4];   // sum q_i, q_j, k_i, k_j
  double ss[4];  // sum of squares, same order
  double x[4];   // q_i.k_i, q_i.k_j, q_j.k_j, q_j.k_i
};

__device__ __forceinline__ void mom_zero(Moments& m) {
#pragma unroll
  for (int e = 0; e < 4; ++e) m.s[e] = m.ss[e] = m.x[e] = 0.0;
}
__device__ __forceinline__ void mom_add(Moments& m, double qi, double qj, double ki, double kj) {
  m.s[0] += qi; m.s[1] += qj; m.s[2] += ki; m.s[3] += kj;
  m.ss[0] = fma(qi, qi, m.ss[0]); m.ss[1] = fma(qj, qj, m.ss[1]);
  m.ss[2] = fma(ki, ki, m.ss[2]); m.ss[3] = fma(kj, kj, m.ss[3]);
  m.x[0] = fma(qi, ki, m.x[0]); m.x[1] = fma(qi, kj, m.x[1]);
  m.x[2] = fma(qj, kj, m.x[2]); m.x[3] = fma(qj, ki, m.x[3]);
}

__device__ __forceinline__ void write_row(float* out, long long N, long long row, double D, double sq,
                                          double sqq, double skp, double skkp, double xp, double skn,
                                          double skkn, double xn) {
  auto cosd = [](double xy, double xx, double yy) { return (float)(1.0 - xy / sqrt(xx * yy)); };
  auto l2d = [D](double xy, double xx, double yy) { return (float)((xx + yy - 2.0 * xy) / D); };
  auto pear = [D](double xy, double sx, double sy, double xx, double yy) {
    const double cov = xy - sx * sy / D, vx = xx - sx * sx / D, vy = yy - sy * sy / D;
    double r = cov / sqrt(vx) / sqrt(vy);  // torch.corrcoef: c / stddev[:, None] / stddev[None, :]
    r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);  // NaN stays NaN, like torch.clip
    return (float)(1.0 - r);
  };
  out[0 * N + row] = cosd(xp, sqq, skkp);
  out[1 * N + row] = cosd(xn, sqq, skkn);
  out[2 * N + row] = l2d(xp, sqq, skkp);
  out[3 * N + row] = l2d(xn, sqq, skkn);
  out[4 * N + row] = pear(xp, sq, skp, sqq, skkp);
  out[5 * N + row] = pear(xn, sq, skn, sqq, skkn);
}

// G threads cooperate on one pair (G = 32: one warp; G = 256: the whole block).
template <int G>
__global__ void __launch_bounds__(256)
rowpair_kernel(const float* __restrict__ q, const float* __restrict__ k, long long N, long long D,
               int pooled_hw, long long npairs, float* __restrict__ out) {
  constexpr int kGroups = 256 / G;
  const int gid = threadIdx.x / G, tg = threadIdx.x % G;
  const long long pair = (long long)blockIdx.x * kGroups + gid;
  const bool active = pair < npairs;
  long long i = 0, j = 1;
  if (active && pair > 0) { i = pair + 1; j = N + 1 - i; }
  Moments m;
  mom_zero(m);
  if (active) {
    if (pooled_hw <= 1) {
      const float *qi = q + i * D, *qj = q + j * D, *ki = k + i * D, *kj = k + j * D;
      if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) & 15) == 0) {
        const long long D4 = D >> 2;
        for (long long e = tg; e < D4; e += G) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(qi) + e);
          const float4 b = __ldg(reinterpret_cast<const float4*>(qj) + e);
          const float4 c = __ldg(reinterpret_cast<const float4*>(ki) + e);
          const float4 d = __ldg(reinterpret_cast<const float4*>(kj) + e);
          mom_add(m, a.x, b.x, c.x, d.x);
          mom_add(m, a.y, b.y, c.y, d.y);
          mom_add(m, a.z, b.z, c.z, d.z);
          mom_add(m, a.w, b.w, c.w, d.w);
        }
      } else {
        for (long long e = tg; e < D; e += G) mom_add(m, qi[e], qj[e], ki[e], kj[e]);
      }
    } else {
      // pooled: D channels, each the mean over pooled_hw contiguous values.  One warp per channel.
      const int lane = tg & 31, wg = tg >> 5;
      constexpr int kWarps = G / 32;
      const long long rowlen = D * pooled_hw;
      const float hwf = (float)pooled_hw;
      for (long long c = wg; c < D; c += kWarps) {
        const float *a = q + i * rowlen + c * pooled_hw, *b = q + j * rowlen + c * pooled_hw,
                    *cc = k + i * rowlen + c * pooled_hw, *d = k + j * rowlen + c * pooled_hw;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int e = lane; e < pooled_hw; e += 32) { s0 += a[e]; s1 += b[e]; s2 += cc[e]; s3 += d[e]; }
        s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
        // the reference pools in fp32 (torch.mean) before the distance
        if (lane == 0) mom_add(m, s0 / hwf, s1 / hwf, s2 / hwf, s3 / hwf);
      }
    }
  }
  // reduce the 12 moments over the group
  double* v = reinterpret_cast<double*>(&m);
#pragma unroll
  for (int e = 0; e < 12; ++e) v[e] = warp_sum_d(v[e]);
  if (G > 32) {
    __shared__ double red[8][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
      for (int e = 0; e < 12; ++e) red[warp][e] = v[e];
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        double a = 0.0;
        for (int w = 0; w < 8; ++w) a += red[w][e];
        v[e] = a;
      }
    }
  }
  if (active && tg == 0) {
    const double Dd = (double)D;
    write_row(out, N, i, Dd, m.s[0], m.ss[0], m.s[2], m.ss[2], m.x[0], m.s[3], m.ss[3], m.x[1]);
    if (j != i)
      write_row(out, N, j, Dd, m.s[1], m.ss[1], m.s[3], m.ss[3], m.x[2], m.s[2], m.ss[2], m.x[3]);
  }
}

__global__ void triplet_hinge_kernel(const float* __restrict__ dist, long long N,
                                     const float* __restrict__ margins, int M,
                                     float* __restrict__ out) {
  const long long total = (long long)M * 3 * N;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t % N;
    const int d = (int)((t / N) % 3);
    const int mi = (int)(t / (3 * N));
    const float pos = dist[(2 * d) * N + i], neg = dist[(2 * d + 1) * N + i];
    const float s = pos - neg + margins[mi];
    // torch.maximum propagates NaN
    out[t] = (s != s) ? s : fmaxf(0.f, s);
  }
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int msp_rowpair_distances(const float* q, const float* k, int N, long long D, int pooled_hw,
                                     float* out, void* stream) {
  MSP_REQUIRE(q && k && out, "rowpair_distances: null pointer");
  MSP_REQUIRE(N >= 2, "rowpair_distances: need at least 2 rows (the reference's negative shift "
                      "indexes k0[-2]); got %d", N);
  MSP_REQUIRE(D >= 1 && pooled_hw >= 0, "rowpair_distances: bad shape");
  const long long n = N;
  const long long npairs = 1 + (((n + 1) / 2 - 1) > 0 ? ((n + 1) / 2 - 1) : 0);
  const long long work = pooled_hw > 1 ? D * pooled_hw : D;
  if (work <= 4096) {
    rowpair_kernel<32><<<(unsigned)((npairs + 7) / 8), 256, 0, ST>>>(q, k, n, D, pooled_hw, npairs, out);
  } else {
    rowpair_kernel<256><<<(unsigned)npairs, 256, 0, ST>>>(q, k, n, D, pooled_hw, npairs, out);
  }
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_triplet_hinge(const float* dist, int N, const float* margins, int M, float* out,
                                 void* stream) {
  MSP_REQUIRE(dist && margins && out && N >= 1 && M >= 1, "triplet_hinge: bad arguments");
  const long long total = (long long)M * 3 * N;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)msp_num_sms() * 8) blocks = (long long)msp_num_sms() * 8;
  triplet_hinge_kernel<<<(unsigned)blocks, 256, 0, ST>>>(dist, N, margins, M, out);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
