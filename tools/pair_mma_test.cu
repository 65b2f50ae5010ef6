// Validation of the CTA-pair (cta_group::2) building blocks of tc_common.cuh on one tile:
//   D[256 x 128] = A[256 x K] * B[128 x K]^T, K = 128 (two 64-wide chunks), bf16 operands, fp32 result.
// CTA r of the pair loads A rows [128r, 128r+128) and B rows [64r, 64r+64) by TMA (complete_tx on the LEADER's
// mbarrier), the leader issues M = 256 MMAs, both CTAs read their 128 accumulator rows from their own TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jatts_b200/csrc tools/pair_mma_test.cu -o tools/pair_mma_test.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
namespace jb { void set_last_error(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); } long long g_launch_count = 0; }
using namespace jb;

constexpr int N = 128, KC = 2;

__global__ void __launch_bounds__(192, 1) pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                                                      float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                       // [KC][128 x 64] bf16 = 16 KB each
  uint8_t* b_s = smem + KC * 16384;          // [KC][64 x 64] bf16  =  8 KB each (this CTA's half of B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + KC * 8192);
  uint64_t* full = bars;        // leader's: all four loads of both CTAs
  uint64_t* done = bars + 1;    // per CTA: accumulator complete
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && elect_one()) {
    const uint32_t leader_full = mapa_u32(smem_u32(full), 0);
    if (rank == 0) mbar_expect_tx(full, 2 * KC * (16384 + 8192));
    for (int kc = 0; kc < KC; ++kc) {
      tma_load_2d_pair(&tm_a, leader_full, a_s + kc * 16384, kc * 64, static_cast<int>(rank) * 128);
      tma_load_2d_pair(&tm_b, leader_full, b_s + kc * 8192, kc * 64, static_cast<int>(rank) * 64);
    }
  } else if (warp == 1 && rank == 0 && elect_one()) {
    mbar_wait(full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc(256, N, true);
    for (int kc = 0; kc < KC; ++kc)
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_sw128_desc(smem_u32(a_s + kc * 16384)) + 2 * k;
        const uint64_t db = make_sw128_desc(smem_u32(b_s + kc * 8192)) + 2 * k;
        tc_mma_f16_pair(tmem, da, db, idesc, (kc | k) != 0 ? 1u : 0u);
      }
    tc_commit_pair(done);
  } else if (warp >= 2) {
    const int lg = warp & 3;   // warps 2..5 -> lane groups 2,3,0,1
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = lg * 32 + (threadIdx.x & 31);
#pragma unroll 1
    for (int c = 0; c < N; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(lg * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(static_cast<size_t>(rank) * 128 + row) * N + c + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

int main() {
  const int M = 256, K = 64 * KC;
  std::vector<bf16> ha(M * K), hb(N * K);
  std::vector<float> fa(M * K), fb(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = v; }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.0f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  bf16 *da, *db; float* dout;
  cudaMalloc(&da, sizeof(bf16) * M * K); cudaMalloc(&db, sizeof(bf16) * N * K); cudaMalloc(&dout, sizeof(float) * M * N);
  cudaMemcpy(da, ha.data(), sizeof(bf16) * M * K, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), sizeof(bf16) * N * K, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, sizeof(float) * M * N);
  CUtensorMap ta, tb;
  if (make_tmap(&ta, da, M, K, K, 128) || make_tmap(&tb, db, N, K, K, 64)) return 1;
  const int smem = KC * (16384 + 8192) + 4096;
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(2); lc.blockDim = dim3(192); lc.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&lc, pair_kernel, ta, tb, dout);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<float> ho(M * N);
  cudaMemcpy(ho.data(), dout, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += double(fa[m * K + k]) * fb[n * K + k];
      const double err = fabs(ref - ho[m * N + n]);
      if (!(err < 1e-3)) { if (bad < 5) printf("mismatch m=%d n=%d ref=%f got=%f\n", m, n, ref, ho[m * N + n]); ++bad; }
      if (err > maxerr) maxerr = err;
    }
  printf("pair MMA: max err %.3g, %d mismatches of %d\n", maxerr, bad, M * N);
  return bad != 0;
}
