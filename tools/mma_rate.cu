// Micro-benchmark: cycles per tcgen05.mma (SS mode, M=128, K=16, bf16) as a function of N, operand row pitch
// (64-byte rows / SWIZZLE_64B vs 128-byte rows / SWIZZLE_128B) and the A start row shift between instructions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jatts_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include "tc_common.cuh"
namespace jb { void set_last_error(const std::string&) {} long long g_launch_count = 0; }
using namespace jb;

template <int N, int KROWB>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n_mma, int row_shift, int kcycle, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(128, N, true);
    constexpr uint32_t desc_hi = static_cast<uint32_t>((8 * KROWB) >> 4) | (1u << 14) | (static_cast<uint32_t>(KROWB == 128 ? 2 : 4) << 29);
    const uint32_t a_lo0 = (1u << 16) + (smem_u32(smem) >> 4);               // A slab: 256 rows
    const uint32_t b_lo0 = (1u << 16) + (smem_u32(smem + 48 * 1024) >> 4);   // B tiles
    const uint32_t a_step = (row_shift * KROWB) >> 4;
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kcycle ? (j % (KROWB / 32)) : 0;
        tc_mma_bf16_lohi(tmem, a_lo0 + (j / (KROWB / 32)) * a_step + 2 * k, desc_hi, b_lo0 + j * ((N * KROWB) >> 4) % 2048 + 2 * k, desc_hi, idesc, 1u);
      }
    }
    long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

// CTA pair: M = 256 (128 rows per CTA), N columns with N/2 weight rows in each CTA's shared memory
template <int N>
__global__ void __launch_bounds__(128, 1) rate_pair_kernel(int n_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc(256, N, true);
    constexpr uint32_t desc_hi = static_cast<uint32_t>(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (1u << 16) + (smem_u32(smem) >> 4);
    const uint32_t b_lo0 = (1u << 16) + (smem_u32(smem + 48 * 1024) >> 4);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) tc_mma_bf16_lohi_pair(tmem, a_lo0 + 2 * (j & 3), desc_hi, b_lo0 + 2 * (j & 3), desc_hi, idesc, 1u);
    }
    long long t1 = clock64();
    tc_commit_pair(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);   // the peer's barrier receives the multicast commit too
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int N>
void run_pair() {
  long long* d; cudaMalloc(&d, 16);
  auto kern = rate_pair_kernel<N>;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 4096;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(148); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaLaunchKernelEx(&lc, kern, n, d);
  cudaLaunchKernelEx(&lc, kern, n, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("pair M=256 N=%3d : issue %.1f clk/mma, complete %.1f clk/mma (per-SM math floor %d) %s\n", N, double(h[0]) / n, double(h[1]) / n,
         N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

template <int N, int KROWB>
void run(int row_shift, int kcycle) {
  long long* d; cudaMalloc(&d, 16);
  auto kern = rate_kernel<N, KROWB>;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 4096;
  kern<<<148, 128, smem>>>(n, row_shift, kcycle, d);
  kern<<<148, 128, smem>>>(n, row_shift, kcycle, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d rowbytes=%3d shift=%2d kcycle=%d : issue %.1f clk/mma, complete %.1f clk/mma (floor %d) %s\n", N, KROWB, row_shift, kcycle,
         double(h[0]) / n, double(h[1]) / n, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run_pair<64>(); run_pair<128>(); run_pair<256>();
  for (int kc = 1; kc < 2; ++kc)
    for (int sh : {5}) {
      run<32, 64>(sh, kc); run<32, 128>(sh, kc);
      run<64, 64>(sh, kc); run<64, 128>(sh, kc);
      run<128, 128>(sh, kc); run<256, 128>(sh, kc);
    }
  return 0;
}
