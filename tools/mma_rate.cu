// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, bf16 -> fp32, M=128) on sm_100a under the
// conv kernels' pipeline structure.  Build + run on the GPU box:
//   nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I medsegpretrainimagenet_b200/csrc \
//        tools/mma_rate.cu -o build/mma_rate && build/mma_rate
// modes: 0 = back-to-back MMAs, one commit at the end
//        1 = + tcgen05.commit after every k-block (4 MMAs)
//        2 = + full/empty mbarrier handshake with a producer thread (no loads)
//        3 = mode 2 with real bulk copies global(L2) -> smem of the operand bytes (A 16 KB + B N*128 B)
//        4 = mode 3, A only reloaded every 9th k-block (halo reuse), B every k-block
#include "msp_common.cuh"
#include <vector>
#include <algorithm>

constexpr int kStages = 4;
constexpr int kStageBytes = 16384 + 32768;

template <int kElect>
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(const uint8_t* __restrict__ src, int N, int mode, int kblocks, long long* out_cycles, int M, int flags,
                int delay) {
  // flags: 1 = alternate two accumulators per MMA, 2 = no tcgen05.fence after the wait, 4 = wait for k-block i+1's
  // barrier before issuing k-block i, 8 = spin `delay` cycles between k-blocks,
  // 16 = roles taken by `lane == 0` instead of elect.sync (ptxas then wraps every UTCHMMA in a per-lane loop)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  (void)lane;
  for (int i = threadIdx.x; i < kStages * kStageBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t b_bytes = (uint32_t)N * 128u;
  if (warp == 0 && mode >= 2 && (kElect ? elect_one() : (lane == 0))) {
    const uint8_t* g = src + (size_t)blockIdx.x * kStageBytes * 8;
    for (int it = 0; it < kblocks; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      if (mode == 2) {
        mbar_arrive(&full_bar[s]);
      } else {
        const bool load_a = (mode == 3) || (it % 9 == 0);
        mbar_expect_tx(&full_bar[s], (load_a ? 16384u : 0u) + b_bytes);
        uint8_t* d = smem + s * kStageBytes;
        const uint8_t* gs = g + (size_t)(it & 7) * kStageBytes;
        if (load_a)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                           "r"(smem_u32(d)), "l"(gs), "r"(16384u), "r"(smem_u32(&full_bar[s])) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                         "r"(smem_u32(d + 16384)), "l"(gs + 16384), "r"(b_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
      }
    }
  } else if (warp == 1 && (kElect ? elect_one() : (lane == 0))) {
    const int mn = (flags & 32) ? 1 : 0;  // 32 = both operands MN-major (the wgrad kernel's form)
    const uint32_t idesc = umma_idesc_bf16(M, N, mn, mn);
    const uint64_t d0 = umma_smem_desc_sw128(smem_u32(smem), mn ? 8192 : 16, 1024);
    const uint32_t kstep = mn ? 128u : 2u;  // descriptor start-address advance per K=16 step (>>4)
    const uint32_t desc_hi = (uint32_t)(d0 >> 32), a_lo0 = (uint32_t)d0;
    uint32_t s = 0, ph = 0;
    const long long c0 = clock64();
    const uint32_t alt = (flags & 1) ? 256u : 0u;
    if ((flags & 4) && mode >= 2) mbar_wait(&full_bar[0], 0);
    for (int it = 0; it < kblocks; ++it) {
      if (mode >= 2 && !(flags & 4)) {
        mbar_wait(&full_bar[s], ph);
        if (!(flags & 2)) tc_fence_after();
      }
      const uint32_t a_lo = a_lo0 + s * (kStageBytes >> 4), b_lo = a_lo + (16384 >> 4);
      const uint32_t tmem_d = tmem_base + (alt ? 0u : ((it / 36) & 1) * (uint32_t)N);
      umma_bf16_lohi(tmem_d, a_lo, desc_hi, b_lo, desc_hi, idesc, it > 0);
      umma_bf16_lohi(tmem_d + alt, a_lo + kstep, desc_hi, b_lo + kstep, desc_hi, idesc, it > 0);
      if ((flags & 4) && mode >= 2 && it + 1 < kblocks) {   // overlap the next barrier wait with queued MMAs
        const uint32_t s1 = (s + 1 == kStages) ? 0 : s + 1, ph1 = (s + 1 == kStages) ? ph ^ 1u : ph;
        mbar_wait(&full_bar[s1], ph1);
        if (!(flags & 2)) tc_fence_after();
      }
      umma_bf16_lohi(tmem_d, a_lo + 2 * kstep, desc_hi, b_lo + 2 * kstep, desc_hi, idesc, 1u);
      umma_bf16_lohi(tmem_d + alt, a_lo + 3 * kstep, desc_hi, b_lo + 3 * kstep, desc_hi, idesc, 1u);
      if (mode >= 1) umma_commit(&empty_bar[s]);
      if (++s == kStages) { s = 0; ph ^= 1u; }
      if (flags & 8) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
    }
    const long long c_issue = clock64();
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long c1 = clock64();
    if (out_cycles) {
      out_cycles[2 * blockIdx.x] = c1 - c0;
      out_cycles[2 * blockIdx.x + 1] = c_issue - c0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src;
  const size_t src_bytes = (size_t)nsm * kStageBytes * 8;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 0x3c, src_bytes);
  long long* out;
  cudaMallocManaged(&out, sizeof(long long) * 2 * nsm);
  const int smem = kStages * kStageBytes + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int kblocks = 2880;
  printf("tcgen05.mma M=128 K=16 bf16, %d SMs, %d k-blocks x 4 MMAs per CTA; floor = N/2 cycles per MMA\n", nsm, kblocks);
  printf("%5s %4s %4s %5s %5s %5s | %12s %12s %10s\n", "grid", "mode", "M", "N", "flags", "delay", "cyc/MMA(med)", "cyc/MMA(max)", "issue/MMA");
  auto run = [&](int grid, int mode, int M, int N, int flags, int delay) {
    for (int rep = 0; rep < 2; ++rep) {
      if (flags & 16) mma_rate_kernel<0><<<grid, 128, smem>>>(src, N, mode, kblocks, out, M, flags, delay);
      else mma_rate_kernel<1><<<grid, 128, smem>>>(src, N, mode, kblocks, out, M, flags, delay);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long mx = 0, iss = 0;
    std::vector<long long> v;
    for (int b = 0; b < grid; ++b) { v.push_back(out[2 * b]); mx = std::max(mx, out[2 * b]); iss += out[2 * b + 1]; }
    std::sort(v.begin(), v.end());
    const double n = 4.0 * kblocks;
    printf("%5d %4d %4d %5d %5d %5d | %12.1f %12.1f %10.1f\n", grid, mode, M, N, flags, delay, v[v.size() / 2] / n, mx / n, iss / n / grid);
  };
  for (int f : {0, 32}) {   // K-major vs MN-major operands
    for (int N : {64, 128, 256}) run(1, 0, 128, N, f, 0);
    for (int N : {64, 128, 256}) run(1, 2, 128, N, f, 0);
    for (int N : {128, 256}) for (int m : {3, 4}) run(nsm, m, 128, N, f, 0);
  }
  return 0;
}
