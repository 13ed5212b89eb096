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
