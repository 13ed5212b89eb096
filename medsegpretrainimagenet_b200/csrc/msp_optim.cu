// Multi-tensor optimizer step and gradient norm: the reference calls
// torch.nn.utils.clip_grad_norm_ (train_model.py:93-98, ~340 vector-norm launches per step on the R50 U-Net) and
// optimizer.step() (train_model.py:107 -> optim/optimizer.py:41-48 -> torch.optim.SGD / AdamW) on every batch.
// Here up to 32 parameter tensors travel in the kernel's parameter space per launch (pointer table by value, blocks
// dealt out by tensor size), so a ResNet-50 step is 5 launches per pass whatever the number of layers; fp32 throughout,
// the update formulas are torch's (torch/optim/sgd.py, adamw.py: decoupled weight decay, lerp for the first moment,
// bias corrections from the step count).  HBM-bound: 16 B / parameter for SGD-momentum, 28 B / parameter for AdamW.
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

constexpr int kOptMaxTensors = 32;
constexpr int kOptThreads = 256;
constexpr int kOptElemsPerBlock = kOptThreads * 16;

struct OptList {
  float* p[kOptMaxTensors];
  float* g[kOptMaxTensors];
  float* a[kOptMaxTensors];
  float* b[kOptMaxTensors];
  long long numel[kOptMaxTensors];
  int first_block[kOptMaxTensors + 1];
  int n;
};

struct OptHyper {
  float lr, momentum, dampening, wd, b1, b2, eps, max_norm;
  double b1d, b2d;     // AdamW betas in double: the bias corrections 1 - beta^t lose 1e-5 of their value in fp32
  float omb1, omb2;    // (float)(1 - beta) evaluated in double, as torch passes them (1.f - 0.999f is off by 1.3e-5)
  int nesterov, first;
  const float* step;   // device: AdamW step count (already incremented)
  double* sq;          // device: sum of squares of all gradients
};

// tensor index and element range of this block
__device__ __forceinline__ int opt_locate(const OptList& L, long long* lo, long long* hi) {
  int t = 0;
  while (t + 1 < L.n && L.first_block[t + 1] <= (int)blockIdx.x) ++t;
  const long long base = (long long)((int)blockIdx.x - L.first_block[t]) * kOptElemsPerBlock;
  *lo = base;
  *hi = base + kOptElemsPerBlock < L.numel[t] ? base + kOptElemsPerBlock : L.numel[t];
  return t;
}

__global__ void __launch_bounds__(kOptThreads) optim_sqnorm_kernel(const OptList L, const OptHyper h) {
  long long lo, hi;
  const int t = opt_locate(L, &lo, &hi);
  const float* __restrict__ g = L.g[t];
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  __shared__ double red[kOptThreads / 32];
  double d = warp_sum_d((double)acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
    atomicAdd(h.sq, s);
  }
}

// torch.nn.utils.clip_grad_norm_: grads *= clamp(max_norm / (total_norm + 1e-6), max = 1)
__global__ void __launch_bounds__(kOptThreads) optim_clip_kernel(const OptList L, const OptHyper h) {
  long long lo, hi;
  const int t = opt_locate(L, &lo, &hi);
  const float total = (float)sqrt(*h.sq);
  float coef = h.max_norm / (total + 1e-6f);
  coef = coef < 1.f ? coef : 1.f;
  float* __restrict__ g = L.g[t];
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) g[i] *= coef;
}

// torch.optim.SGD (no maximize): d = g + wd p; buf = first ? d : momentum buf + (1 - dampening) d;
// d = nesterov ? d + momentum buf : buf; p -= lr d
__global__ void __launch_bounds__(kOptThreads) optim_sgd_kernel(const OptList L, const OptHyper h) {
  long long lo, hi;
  const int t = opt_locate(L, &lo, &hi);
  float* __restrict__ p = L.p[t];
  const float* __restrict__ g = L.g[t];
  float* __restrict__ buf = L.a[t];
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) {
    const float pv = p[i];
    float d = g[i];
    if (h.wd != 0.f) d = fmaf(h.wd, pv, d);
    if (h.momentum != 0.f) {
      float bv;
      if (h.first) bv = d;
      else bv = fmaf(h.momentum, buf[i], (1.f - h.dampening) * d);
      buf[i] = bv;
      d = h.nesterov ? fmaf(h.momentum, bv, d) : bv;
    }
    p[i] = fmaf(-h.lr, d, pv);
  }
}

// torch.optim.AdamW (no amsgrad / maximize): p *= 1 - lr wd; m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void __launch_bounds__(kOptThreads) optim_adamw_kernel(const OptList L, const OptHyper h) {
  long long lo, hi;
  const int t = opt_locate(L, &lo, &hi);
  __shared__ float bc[2];
  if (threadIdx.x == 0) {  // as torch does on the host: bias corrections and step size in double, then fp32
    const double step = (double)*h.step;
    const double bias1 = 1.0 - pow(h.b1d, step), bias2 = 1.0 - pow(h.b2d, step);
    bc[0] = (float)((double)h.lr / bias1);
    bc[1] = (float)sqrt(bias2);
  }
  __syncthreads();
  const float step_size = bc[0], sb2 = bc[1], decay = 1.f - h.lr * h.wd;
  float* __restrict__ p = L.p[t];
  const float* __restrict__ g = L.g[t];
  float* __restrict__ m = L.a[t];
  float* __restrict__ v = L.b[t];
  for (long long i = lo + threadIdx.x; i < hi; i += kOptThreads) {
    const float gv = g[i];
    const float pv = p[i] * decay;
    const float mv = fmaf(gv - m[i], h.omb1, m[i]);
    const float vv = fmaf(h.b2, v[i], h.omb2 * gv * gv);
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / sb2 + h.eps;
    p[i] = pv - step_size * (mv / denom);
  }
}

// sqrt of the accumulated sum of squares as the fp32 scalar clip_grad_norm_ returns
__global__ void optim_norm_kernel(const double* __restrict__ sq, float* __restrict__ out) { *out = (float)sqrt(*sq); }

// AdamW step counters (fp32 scalars shared by all parameters of the same age): += value
struct ScalarList {
  float* s[kOptMaxTensors];
  int n;
};
__global__ void optim_add_scalar_kernel(const ScalarList L, float value) {
  if ((int)threadIdx.x < L.n) *L.s[threadIdx.x] += value;
}

int fill_list(OptList* L, int n, void* const* p, void* const* g, void* const* a, void* const* b, const long long* numel,
              int* total_blocks) {
  MSP_REQUIRE(n >= 1 && n <= kOptMaxTensors && g && numel, "optim: 1..%d tensors per call (got %d)", kOptMaxTensors, n);
  memset(L, 0, sizeof(*L));
  int first = 0;
  for (int i = 0; i < n; ++i) {
    MSP_REQUIRE(numel[i] > 0 && g[i], "optim: tensor %d is empty or has no gradient", i);
    L->p[i] = p ? (float*)p[i] : nullptr;
    L->g[i] = (float*)g[i];
    L->a[i] = a ? (float*)a[i] : nullptr;
    L->b[i] = b ? (float*)b[i] : nullptr;
    L->numel[i] = numel[i];
    L->first_block[i] = first;
    const long long nb = (numel[i] + kOptElemsPerBlock - 1) / kOptElemsPerBlock;
    MSP_REQUIRE(first + nb < (1ll << 30), "optim: too many blocks");
    first += (int)nb;
  }
  L->first_block[n] = first;
  L->n = n;
  *total_blocks = first;
  return MSP_OK;
}

}  // namespace

extern "C" int msp_optim_sqnorm(int n, void* const* grads, const long long* numel, double* sq_accum, int zero_first,
                                void* stream) {
  MSP_REQUIRE(sq_accum, "optim_sqnorm: null accumulator");
  OptList L;
  int blocks = 0;
  int rc = fill_list(&L, n, nullptr, grads, nullptr, nullptr, numel, &blocks);
  if (rc) return rc;
  if (zero_first) MSP_CHECK_CUDA(cudaMemsetAsync(sq_accum, 0, sizeof(double), (cudaStream_t)stream));
  OptHyper h;
  memset(&h, 0, sizeof(h));
  h.sq = sq_accum;
  optim_sqnorm_kernel<<<blocks, kOptThreads, 0, (cudaStream_t)stream>>>(L, h);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_optim_norm(const double* sq, float* norm_out, void* stream) {
  MSP_REQUIRE(sq && norm_out, "optim_norm: null pointer");
  optim_norm_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sq, norm_out);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_optim_add_scalar(int n, void* const* scalars, float value, void* stream) {
  MSP_REQUIRE(n >= 1 && n <= kOptMaxTensors && scalars, "optim_add_scalar: 1..%d scalars per call (got %d)", kOptMaxTensors, n);
  ScalarList L;
  memset(&L, 0, sizeof(L));
  for (int i = 0; i < n; ++i) {
    MSP_REQUIRE(scalars[i], "optim_add_scalar: null scalar %d", i);
    L.s[i] = (float*)scalars[i];
  }
  L.n = n;
  optim_add_scalar_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(L, value);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_optim_clip(int n, void* const* grads, const long long* numel, const double* sq, float max_norm,
                              void* stream) {
  MSP_REQUIRE(sq && max_norm >= 0.f, "optim_clip: bad arguments");
  OptList L;
  int blocks = 0;
  int rc = fill_list(&L, n, nullptr, grads, nullptr, nullptr, numel, &blocks);
  if (rc) return rc;
  OptHyper h;
  memset(&h, 0, sizeof(h));
  h.sq = const_cast<double*>(sq);
  h.max_norm = max_norm;
  optim_clip_kernel<<<blocks, kOptThreads, 0, (cudaStream_t)stream>>>(L, h);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_optim_sgd(int n, void* const* params, void* const* grads, void* const* momentum_bufs,
                             const long long* numel, float lr, float momentum, float dampening, float weight_decay,
                             int nesterov, int first_step, void* stream) {
  MSP_REQUIRE(params && (momentum == 0.f || momentum_bufs), "optim_sgd: null pointer");
  OptList L;
  int blocks = 0;
  int rc = fill_list(&L, n, params, grads, momentum_bufs, nullptr, numel, &blocks);
  if (rc) return rc;
  OptHyper h;
  memset(&h, 0, sizeof(h));
  h.lr = lr; h.momentum = momentum; h.dampening = dampening; h.wd = weight_decay;
  h.nesterov = nesterov; h.first = first_step;
  optim_sgd_kernel<<<blocks, kOptThreads, 0, (cudaStream_t)stream>>>(L, h);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_optim_adamw(int n, void* const* params, void* const* grads, void* const* exp_avg,
                               void* const* exp_avg_sq, const long long* numel, float lr, double beta1, double beta2,
                               float eps, float weight_decay, const float* step_dev, void* stream) {
  MSP_REQUIRE(params && exp_avg && exp_avg_sq && step_dev, "optim_adamw: null pointer");
  OptList L;
  int blocks = 0;
  int rc = fill_list(&L, n, params, grads, exp_avg, exp_avg_sq, numel, &blocks);
  if (rc) return rc;
  OptHyper h;
  memset(&h, 0, sizeof(h));
  h.lr = lr; h.b1 = (float)beta1; h.b2 = (float)beta2; h.b1d = beta1; h.b2d = beta2; h.eps = eps; h.wd = weight_decay;
  h.omb1 = (float)(1.0 - beta1); h.omb2 = (float)(1.0 - beta2);
  h.step = step_dev;
  optim_adamw_kernel<<<blocks, kOptThreads, 0, (cudaStream_t)stream>>>(L, h);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
