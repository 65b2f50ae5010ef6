// conv_bf16_tma_kernel: the HiFi-GAN convolution kernel (bf16 operands, fp32 TMEM accumulation) with a
// TMA-staged epilogue.  Same implicit-GEMM mainloop as conv_gemm_tc_kernel (TMA producer warp,
// single-thread tcgen05.mma issuer, TMEM double buffer); what changes is how a finished tile leaves:
//
//   warp 2      epilogue loader: TMA-loads the residual / branch-sum tiles of upcoming tiles into a
//               ring of swizzled smem slabs (so no epilogue thread ever waits on a global load)
//   warps 4..7  epilogue: tcgen05.ld the accumulator, read the residual slab from smem, apply
//               bias / LeakyReLU / residual / branch sum / mean, write the bf16 outputs back into the
//               same slab (in place), fence.proxy.async, and one thread issues the TMA stores.
//
// The round-1 baseline epilogue went registers -> global with one row per thread and was latency bound
// (tensor pipe 5 %, DRAM 2-9 %, profiles/r01_summary.md); here every global access of the kernel is a
// bulk asynchronous copy.  Rows that do not belong to an utterance are stored as zeros (the packed
// layout needs its gap rows to stay zero); rows past the end of the tensor are clipped by TMA.
#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace jb {

static constexpr int BLOCK_M2 = 128;
static constexpr int BLOCK_K2 = 64;
static constexpr int UMMA_K2 = 16;
static constexpr int kThreads2 = 256;

struct KernelParams2 {
  int taps, k_chunks, n_pad;
  int tap_off0, tap_stride;
  int m_rows;
  int num_m_tiles, num_n_tiles;
  const uint8_t* frame_mask;
  int rate;
  const float* bias;
  int act;          // ACT_NONE or ACT_LRELU
  float slope;
  float post_scale;
  float out1_slope;
  int has_res, has_acc, has_out0, has_out1;
};

template <int BLOCK_N>
struct Cfg2 {
  static constexpr int SLAB = BLOCK_N >= 64 ? 64 : 32;     // columns per epilogue slab
  static constexpr int N_SLABS = BLOCK_N / SLAB;
  static constexpr int ROWB = SLAB * 2;                     // bytes per slab row = swizzle span
  static constexpr int SW_MASK = ROWB == 128 ? 7 : 3;       // Swizzle<3,4,3> / Swizzle<2,4,3>
  static constexpr int SLAB_BYTES = BLOCK_M2 * ROWB;
  static constexpr int ENTRY_BYTES = 2 * SLAB_BYTES;        // [residual -> out0 | branch sum -> out1]
  static constexpr int EP_ENTRIES = BLOCK_N == 256 ? 2 : (BLOCK_N == 32 ? 4 : 3);
  static constexpr int A_BYTES = BLOCK_M2 * BLOCK_K2 * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K2 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2048;                   // n_pad <= 512
  static constexpr int BAR_BYTES = 512;
  static constexpr int BUDGET = 225 * 1024 - EP_ENTRIES * ENTRY_BYTES - BIAS_BYTES - BAR_BYTES - 1024;
  static constexpr int MAX_STAGES = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EP_ENTRIES * ENTRY_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads2, 1)
conv_bf16_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_acc,
                     const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1,
                     const __grid_constant__ KernelParams2 P) {
  using C = Cfg2<BLOCK_N>;
  constexpr int STAGES = C::STAGES;
  constexpr int E = C::EP_ENTRIES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ep_base = smem + STAGES * C::STAGE_BYTES;                 // 1024-aligned (all sizes are multiples)
  float* bias_s = reinterpret_cast<float*>(ep_base + E * C::ENTRY_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + C::BIAS_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;      // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* epfull_bar = tempty_bar + 2;        // [E]
  uint64_t* epempty_bar = epfull_bar + E;       // [E]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epempty_bar + E);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = P.num_m_tiles * P.num_n_tiles;
  const int k_iters = P.taps * P.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (P.has_res) tma_prefetch_desc(&tm_res);
    if (P.has_acc) tma_prefetch_desc(&tm_acc);
    if (P.has_out0) tma_prefetch_desc(&tm_out0);
    if (P.has_out1) tma_prefetch_desc(&tm_out1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    for (int i = 0; i < E; ++i) {
      mbar_init(&epfull_bar[i], 1);
      mbar_init(&epempty_bar[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(C::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < P.n_pad && i < C::BIAS_BYTES / 4; i += kThreads2) bias_s[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== mainloop TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.num_n_tiles) * BLOCK_M2;
        const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
        for (int tap = 0; tap < P.taps; ++tap) {
          const int arow = m0 + P.tap_off0 + tap * P.tap_stride;
          const int brow = tap * P.n_pad + n0;
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* s = smem + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(&tm_a, &full_bar[stage], s, kc * BLOCK_K2, arow);
            tma_load_2d(&tm_b, &full_bar[stage], s + C::A_BYTES, kc * BLOCK_K2, brow);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M2, BLOCK_N, /*is_bf16=*/true);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t da = make_sw128_desc(sa);
          const uint64_t db = make_sw128_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K2 / UMMA_K2; ++k) {
            const uint64_t koff = static_cast<uint64_t>((k * UMMA_K2 * 2) >> 4);
            tc_mma_bf16(tmem_d, da + koff, db + koff, idesc, (it | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue loader =====================
    if (elect_one()) {
      int e = 0;
      uint32_t ph = 0;
      const uint32_t bytes = (P.has_res ? C::SLAB_BYTES : 0) + (P.has_acc ? C::SLAB_BYTES : 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.num_n_tiles) * BLOCK_M2;
        const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
        for (int s = 0; s < C::N_SLABS; ++s) {
          mbar_wait(&epempty_bar[e], ph ^ 1);
          uint8_t* buf = ep_base + e * C::ENTRY_BYTES;
          if (bytes) {
            mbar_expect_tx(&epfull_bar[e], bytes);
            if (P.has_res) tma_load_2d(&tm_res, &epfull_bar[e], buf, n0 + s * C::SLAB, m0);
            if (P.has_acc) tma_load_2d(&tm_acc, &epfull_bar[e], buf + C::SLAB_BYTES, n0 + s * C::SLAB, m0);
          } else {
            mbar_arrive(&epfull_bar[e]);
          }
          if (++e == E) { e = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..7) =====================
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t row_off = static_cast<uint32_t>(row_in_tile * C::ROWB);
    const uint32_t sw = (row_off >> 7) & C::SW_MASK;   // XOR pattern of this row's 16-byte chunks
    const bool issuer = (warp == 4 && lane == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int e = 0;
    uint32_t ph = 0;
    int groups = 0;  // bulk store groups committed by the issuer
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / P.num_n_tiles) * BLOCK_M2;
      const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
      const int row = m0 + row_in_tile;
      bool valid = row < P.m_rows;
      if (valid && P.frame_mask) valid = P.frame_mask[row / P.rate] != 0;   // issued before the accumulator wait
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < C::N_SLABS; ++s) {
        uint32_t r[C::SLAB];
        {
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          tmem_ld32(lane_addr + static_cast<uint32_t>(acc * BLOCK_N + s * C::SLAB), r0);
          if (C::SLAB == 64) {
            uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[C::SLAB == 64 ? 32 : 0]);
            tmem_ld32(lane_addr + static_cast<uint32_t>(acc * BLOCK_N + s * C::SLAB + 32), r1);
          }
        }
        mbar_wait(&epfull_bar[e], ph);       // residual / branch-sum slabs have landed (or entry is free)
        tmem_ld_wait();
        uint8_t* bufA = ep_base + e * C::ENTRY_BYTES;
        uint8_t* bufB = bufA + C::SLAB_BYTES;
        const float* bs = bias_s + n0 + s * C::SLAB;
#pragma unroll
        for (int c = 0; c < C::SLAB / 8; ++c) {
          const uint32_t off = row_off + ((static_cast<uint32_t>(c) ^ sw) << 4);
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = __uint_as_float(r[c * 8 + i]) + bs[c * 8 + i];
            if (P.act == ACT_LRELU) x = x > 0.f ? x : x * P.slope;
            v[i] = x;
          }
          if (P.has_res) {
            const uint4 t = *reinterpret_cast<const uint4*>(bufA + off);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
          if (P.has_acc) {
            const uint4 t = *reinterpret_cast<const uint4*>(bufB + off);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = valid ? v[i] * P.post_scale : 0.f;
          if (P.has_out0) {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(bufA + off) = o;
          }
          if (P.has_out1) {
            const float sl = P.out1_slope;
            uint4 o;
            o.x = pack_bf16x2(v[0] > 0.f ? v[0] : v[0] * sl, v[1] > 0.f ? v[1] : v[1] * sl);
            o.y = pack_bf16x2(v[2] > 0.f ? v[2] : v[2] * sl, v[3] > 0.f ? v[3] : v[3] * sl);
            o.z = pack_bf16x2(v[4] > 0.f ? v[4] : v[4] * sl, v[5] > 0.f ? v[5] : v[5] * sl);
            o.w = pack_bf16x2(v[6] > 0.f ? v[6] : v[6] * sl, v[7] > 0.f ? v[7] : v[7] * sl);
            *reinterpret_cast<uint4*>(bufB + off) = o;
          }
        }
        if (s == C::N_SLABS - 1) {
          // all TMEM reads of this accumulator are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        fence_proxy_async_smem();            // make the generic-proxy smem writes visible to the TMA engine
        named_bar_sync(1, 128);
        if (issuer) {
          if (P.has_out0) tma_store_2d(&tm_out0, bufA, n0 + s * C::SLAB, m0);
          if (P.has_out1) tma_store_2d(&tm_out1, bufB, n0 + s * C::SLAB, m0);
          tma_store_commit();
          ++groups;
          // the store issued E-1 groups ago has finished reading its slab: release that entry
          tma_store_wait_read<E - 1>();
          if (groups >= E) mbar_arrive(&epempty_bar[(e + 1) % E]);
        }
        if (++e == E) { e = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(C::TMEM_COLS))
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
bool conv_gemm_tc2_eligible(const ConvGemmProblem& p) {
  const ConvGemmEpilogue& e = p.ep;
  if (p.a_lo || p.w_lo || p.up_s > 0) return false;
  if (!(e.act == ACT_NONE || e.act == ACT_LRELU)) return false;
  if (e.res_f32 || e.accum_in || e.out_f32 || e.out_lo) return false;
  if (e.scale != 1.0f) return false;
  if (p.n != p.n_pad || p.n_pad > 512) return false;
  if (p.out_rows != p.m_rows) return false;
  const int slab = p.block_n >= 64 ? 64 : 32;
  auto ok = [&](const void* ptr, int ld) { return ptr == nullptr || (ld % 8 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  if (!ok(e.res_bf16, e.res_ld) || !ok(e.accum_bf16, e.res_ld) || !ok(e.out_hi, e.out_bf_ld) || !ok(e.out_act, e.out_act_ld))
    return false;
  if (!e.out_hi && !e.out_act) return false;
  (void)slab;
  return true;
}

template <int BLOCK_N>
static int launch2(const ConvGemmProblem& p, cudaStream_t stream) {
  using C = Cfg2<BLOCK_N>;
  static_assert(C::STAGES >= 2, "need at least a double buffer");
  const ConvGemmEpilogue& e = p.ep;
  CUtensorMap ta, tb, tres, tacc, to0, to1;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  JB_PROPAGATE(make_tmap(&ta, p.a_hi, p.a_rows, a_cols, p.a_ld, BLOCK_M2));
  JB_PROPAGATE(make_tmap(&tb, p.w_hi, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, BLOCK_N));
  tres = tacc = to0 = to1 = ta;
  if (e.res_bf16) JB_PROPAGATE(make_tmap(&tres, e.res_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  if (e.accum_bf16) JB_PROPAGATE(make_tmap(&tacc, e.accum_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  if (e.out_hi) JB_PROPAGATE(make_tmap(&to0, e.out_hi, p.m_rows, p.n, e.out_bf_ld, BLOCK_M2, C::SLAB));
  if (e.out_act) JB_PROPAGATE(make_tmap(&to1, e.out_act, p.m_rows, p.n, e.out_act_ld, BLOCK_M2, C::SLAB));
  KernelParams2 kp;
  kp.taps = p.taps;
  kp.k_chunks = p.k_pad / BLOCK_K2;
  kp.n_pad = p.n_pad;
  kp.tap_off0 = p.tap_off0;
  kp.tap_stride = p.tap_stride;
  kp.m_rows = p.m_rows;
  kp.num_m_tiles = ceil_div(p.m_rows, BLOCK_M2);
  kp.num_n_tiles = p.n_pad / BLOCK_N;
  kp.frame_mask = p.frame_mask;
  kp.rate = p.rate > 0 ? p.rate : 1;
  kp.bias = e.bias;
  kp.act = e.act;
  kp.slope = e.slope;
  kp.post_scale = e.post_scale;
  kp.out1_slope = e.out_act_slope;
  kp.has_res = e.res_bf16 != nullptr;
  kp.has_acc = e.accum_bf16 != nullptr;
  kp.has_out0 = e.out_hi != nullptr;
  kp.has_out1 = e.out_act != nullptr;
  auto kern = conv_bf16_tma_kernel<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    JB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = kp.num_m_tiles * kp.num_n_tiles;
  if (tiles == 0) return 0;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventCreate(&e0));
    JB_CUDA_OK(cudaEventCreate(&e1));
    JB_CUDA_OK(cudaEventRecord(e0, stream));
  }
  kern<<<grid, kThreads2, C::SMEM_BYTES, stream>>>(ta, tb, tres, tacc, to0, to1, kp);
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, 0});
  }
  return 0;
}

int conv_gemm_tc2(const ConvGemmProblem& p, cudaStream_t stream) {
  switch (p.block_n) {
    case 32: return launch2<32>(p, stream);
    case 64: return launch2<64>(p, stream);
    case 128: return launch2<128>(p, stream);
    case 256: return launch2<256>(p, stream);
  }
  return -2;
}

}  // namespace jb
