// Small-vector all-reduce over NVLink peer memory (one process per GPU, one node).
//
// The hot path has ~100 tiny exchanges per training step: the [2C] fp32 BatchNorm sums of every
// BatchNorm layer, forward and backward (SyncBN), and the 3C Dice sums.  A NCCL all-reduce costs 15-25 us each at 8 GPUs
// whatever the size — 2.6 ms of a 26 ms ResNet-50 step.  Here every rank owns a communication buffer that all peers map
// (cudaIpc handles exchanged once through torch.distributed): ONE single-CTA kernel pushes the rank's vector into every
// peer's buffer with plain NVLink stores — 8-byte words that carry the value AND the exchange's sequence number — and
// adds the W vectors in rank order as the peers' words arrive in ITS buffer (bitwise identical result on every rank,
// run to run).  No host round trip, no NCCL; CUDA-graph capturable (the sequence counter lives in device
// memory and advances inside the kernel).  Two slots alternate so that a fast rank's next exchange never overwrites
// what a slow rank is still adding up.
//
// Replaces the dist.all_reduce calls of functional._allreduce_sum / losses._Dice; the reference has no equivalent
// (nn.DataParallel BatchNorm is per replica, train_model.py:194).
#include "msp_common.cuh"
#include "../../include/msp_b200.h"

extern void msp_count_launch(int n);

namespace {

constexpr int kP2PMaxWorld = 16;

struct P2PTable {
  unsigned char* buf[kP2PMaxWorld];  // every rank's communication buffer, as mapped in THIS process
};

// buffer layout: [2 slots][world][max_n] 8-byte words {value bits, sequence number}.  Every word validates itself
// (the "LL" idea of NCCL's low-latency protocol): an 8-byte store is atomic, so the receiver needs no separate flag, no
// fence and no second NVLink round trip — it spins on each word until its sequence half is the current exchange's.
__host__ __device__ inline size_t p2p_total_bytes(int world, int max_n) { return (size_t)2 * world * max_n * 8; }

__device__ __forceinline__ void st_word(uint2* p, float v, unsigned seq) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ uint2 ld_word(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(float* __restrict__ data, int n, int rank, int world, int max_n, P2PTable tbl,
                     unsigned* __restrict__ seq_ptr) {
  __shared__ unsigned seq_s;
  if (threadIdx.x == 0) seq_s = *seq_ptr + 1u;
  __syncthreads();
  const unsigned seq = seq_s;
  const size_t slot_off = (size_t)(seq & 1u) * world * max_n;
  // 1. push this rank's vector, tagged with the exchange number, into every peer's buffer
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    uint2* dst = reinterpret_cast<uint2*>(tbl.buf[p]) + slot_off + (size_t)rank * max_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) st_word(dst + i, data[i], seq);
  }
  // 2. add the W vectors in rank order as their words arrive (bounded spin: a rank that never arrives traps instead
  //    of hanging the GPU)
  const uint2* mine = reinterpret_cast<const uint2*>(tbl.buf[rank]) + slot_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    uint2 w[kP2PMaxWorld];
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)  // all peers' words in flight at once, then re-poll only the late ones
      if (r < world && r != rank) w[r] = ld_word(mine + (size_t)r * max_n + i);
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r) {
      if (r >= world) continue;
      if (r == rank) {
        s += data[i];
        continue;
      }
      unsigned long long spins = 0;
      while (w[r].y != seq) {
        if (++spins > (1ull << 30)) {
          printf("msp p2p all-reduce: rank %d timed out waiting for rank %d (exchange %u, element %d)\n", rank, r, seq, i);
          __trap();
        }
        w[r] = ld_word(mine + (size_t)r * max_n + i);
      }
      s += __uint_as_float(w[r].x);
    }
    data[i] = s;
  }
  if (threadIdx.x == 0) *seq_ptr = seq;
}


// ------------------------------------------------------------------------------------------------
// SyncBN statistic exchange in ONE kernel (msp_p2p_stats_exchange): per-CTA rows -> this rank's sums (fixed order) ->
// push to every peer / add the peers' sums in rank order -> optionally mean / invstd / running statistics.
// Before: reduce_rows (deterministic mode) -> single-CTA all-reduce -> bn_finalize, plus two clone() kernels in the
// backward pass — three to five dependent launches of a few microseconds each per BatchNorm layer and direction, 150
// exchanges per R50 U-Net step on the critical path.  Here a block owns 32 channels of BOTH halves of the [2][C]
// vector (sum and sum of squares / sum g and sum g*xhat) and exchanges just those 64 words, so wide layers use C / 32
// blocks in parallel.  The exchange number is read by every block when it starts and advanced by the LAST block to
// finish (ticket counter), hence identical in all blocks of a launch and on all ranks.
// ------------------------------------------------------------------------------------------------
struct StatsFinalize {
  double count;      // global element count per channel
  float eps, momentum;
  float* mean;       // nullptr: no finalize (backward sums)
  float* invstd;
  float* rmean;      // may be nullptr
  float* rvar;
};

__global__ void __launch_bounds__(32 * kRowGroups)
p2p_stats_kernel(float* __restrict__ ws, int rows, int C, float* __restrict__ local_a, float* __restrict__ local_b,
                 int local_add, float* __restrict__ global_out, int reset, int rank, int world, int max_n, P2PTable tbl, unsigned* __restrict__ seq_ptr,
                 unsigned* __restrict__ ticket, StatsFinalize fin) {
  __shared__ float part[kRowGroups][33];
  __shared__ float part2[kRowGroups][33];
  __shared__ float mine_s[64];
  __shared__ float tot_s[64];
  __shared__ unsigned seq_s;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, c = c0 + lane, cc = c < C ? c : C - 1;
  const int nvalid = C - c0 < 32 ? C - c0 : 32;
  if (threadIdx.x == 0) seq_s = *reinterpret_cast<volatile unsigned*>(seq_ptr) + 1u;
  const long long stride = 2ll * C;
  const float2 t12 = sum_rows_fixed2(ws, rows, stride, cc, C, grp, reset && c < C, part, part2);
  const float t1 = t12.x, t2 = t12.y;
  if (grp == 0) {
    mine_s[lane] = t1;
    mine_s[32 + lane] = t2;
    if (c < C) {  // the rank-local sums: BatchNorm backward's dbeta / dgamma, written or ADDED (param.grad, bucket views)
      if (local_a != nullptr) local_a[c] = local_add ? local_a[c] + t1 : t1;
      if (local_b != nullptr) local_b[c] = local_add ? local_b[c] + t2 : t2;
    }
  }
  __syncthreads();
  const unsigned seq = seq_s;
  const size_t slot_off = (size_t)(seq & 1u) * world * max_n;
  // 1. push the block's 64 words into every peer's buffer (all threads: one word each per pass)
  for (int idx = threadIdx.x; idx < 64 * (world - 1); idx += blockDim.x) {
    const int pi = idx >> 6, j = idx & 63, half = j >> 5, l = j & 31;
    const int p = pi < rank ? pi : pi + 1;
    if (l < nvalid)
      st_word(reinterpret_cast<uint2*>(tbl.buf[p]) + slot_off + (size_t)rank * max_n + (size_t)half * C + c0 + l, mine_s[j], seq);
  }
  // 2. threads 0..63 add the W values of their word in rank order
  if (threadIdx.x < 64 && (threadIdx.x & 31) < nvalid) {
    const int j = threadIdx.x, half = j >> 5, l = j & 31;
    const uint2* mine = reinterpret_cast<const uint2*>(tbl.buf[rank]) + slot_off + (size_t)half * C + c0 + l;
    uint2 w[kP2PMaxWorld];
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r)
      if (r < world && r != rank) w[r] = ld_word(mine + (size_t)r * max_n);
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kP2PMaxWorld; ++r) {
      if (r >= world) continue;
      if (r == rank) {
        s += mine_s[j];
        continue;
      }
      unsigned long long spins = 0;
      while (w[r].y != seq) {
        if (++spins > (1ull << 30)) {
          printf("msp p2p statistic exchange: rank %d timed out waiting for rank %d (exchange %u, block %d)\n", rank, r, seq,
                 (int)blockIdx.x);
          __trap();
        }
        w[r] = ld_word(mine + (size_t)r * max_n);
      }
      s += __uint_as_float(w[r].x);
    }
    tot_s[j] = s;
    if (global_out != nullptr) global_out[(size_t)half * C + c0 + l] = s;
  }
  __syncthreads();
  // 3. BatchNorm forward: mean / invstd / running statistics from the global sums (bn_finalize_kernel's arithmetic)
  if (fin.mean != nullptr && threadIdx.x < 32 && c < C) {
    const double m = (double)tot_s[lane] / fin.count;
    double var = (double)tot_s[32 + lane] / fin.count - m * m;
    if (var < 0) var = 0;
    fin.mean[c] = (float)m;
    fin.invstd[c] = (float)(1.0 / sqrt(var + (double)fin.eps));
    if (fin.rmean != nullptr) {
      const double unb = fin.count > 1 ? var * fin.count / (fin.count - 1) : var;
      fin.rmean[c] = (1.f - fin.momentum) * fin.rmean[c] + fin.momentum * (float)m;
      fin.rvar[c] = (1.f - fin.momentum) * fin.rvar[c] + fin.momentum * (float)unb;
    }
  }
  // 4. the last block to finish advances the exchange number (every block has read the old one by then)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      *ticket = 0u;
      *seq_ptr = seq;
    }
  }
}

}  // namespace

extern "C" long long msp_p2p_buffer_bytes(int world, int max_n) {
  if (world < 1 || world > kP2PMaxWorld || max_n < 1) return -1;
  return (long long)p2p_total_bytes(world, max_n);
}

extern "C" int msp_p2p_alloc(long long bytes, void** ptr, void* handle64) {
  MSP_REQUIRE(bytes > 0 && ptr && handle64, "p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  MSP_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  MSP_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  MSP_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  MSP_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *ptr = p;
  return MSP_OK;
}

extern "C" int msp_p2p_open(const void* handle64, void** ptr) {
  MSP_REQUIRE(handle64 && ptr, "p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  MSP_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return MSP_OK;
}

extern "C" int msp_p2p_close(void* ptr) {
  MSP_REQUIRE(ptr, "p2p_close: null pointer");
  MSP_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return MSP_OK;
}

extern "C" int msp_p2p_free(void* ptr) {
  MSP_REQUIRE(ptr, "p2p_free: null pointer");
  MSP_CHECK_CUDA(cudaFree(ptr));
  return MSP_OK;
}

extern "C" int msp_p2p_allreduce_sum_f32(float* data, int n, int rank, int world, int max_n, void* const* bufs,
                                         unsigned* seq, void* stream) {
  MSP_REQUIRE(data && bufs && seq && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world && n >= 1 &&
                  n <= max_n,
              "p2p_allreduce: bad arguments (n %d, max_n %d, rank %d, world %d)", n, max_n, rank, world);
  P2PTable tbl;
  for (int r = 0; r < kP2PMaxWorld; ++r) tbl.buf[r] = r < world ? (unsigned char*)bufs[r] : nullptr;
  for (int r = 0; r < world; ++r) MSP_REQUIRE(tbl.buf[r] != nullptr, "p2p_allreduce: rank %d buffer not mapped", r);
  p2p_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, rank, world, max_n, tbl, seq);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}

extern "C" int msp_p2p_stats_exchange(float* ws, int rows, int C, float* local_a, float* local_b, int local_add,
                                      float* global_out, int reset, double count, float eps, float momentum, float* mean, float* invstd,
                                      float* running_mean, float* running_var, int rank, int world, int max_n,
                                      void* const* bufs, unsigned* seq, unsigned* ticket, void* stream) {
  MSP_REQUIRE(ws && bufs && seq && ticket && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world && rows >= 1 &&
                  C >= 1 && 2 * C <= max_n,
              "p2p_stats_exchange: bad arguments (rows %d, C %d, max_n %d, rank %d, world %d)", rows, C, max_n, rank, world);
  MSP_REQUIRE((mean == nullptr) == (invstd == nullptr) && (running_mean == nullptr) == (running_var == nullptr),
              "p2p_stats_exchange: mean / invstd and the running buffers come in pairs");
  MSP_REQUIRE(mean == nullptr || count > 0, "p2p_stats_exchange: finalize needs the global element count");
  MSP_REQUIRE(mean != nullptr || global_out != nullptr, "p2p_stats_exchange: nothing to produce");
  P2PTable tbl;
  for (int r = 0; r < kP2PMaxWorld; ++r) tbl.buf[r] = r < world ? (unsigned char*)bufs[r] : nullptr;
  for (int r = 0; r < world; ++r) MSP_REQUIRE(tbl.buf[r] != nullptr, "p2p_stats_exchange: rank %d buffer not mapped", r);
  StatsFinalize fin{count, eps, momentum, mean, invstd, running_mean, running_var};
  p2p_stats_kernel<<<(C + 31) / 32, 32 * kRowGroups, 0, (cudaStream_t)stream>>>(ws, rows, C, local_a, local_b, local_add,
                                                                              global_out, reset, rank, world, max_n, tbl, seq, ticket, fin);
  MSP_CHECK_LAUNCH();
  msp_count_launch(1);
  return MSP_OK;
}
