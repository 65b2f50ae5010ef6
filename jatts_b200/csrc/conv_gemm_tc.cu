// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel for sm_100a.  See conv_gemm.cuh.
//
// Persistent, warp-specialised CTA (192 threads, 1 CTA per SM):
//   warp 0      TMA producer   (one elected lane; cp.async.bulk.tensor.2d, 128B swizzle)
//   warp 1      MMA issuer     (one elected lane; tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16)
//               + TMEM allocation / deallocation (whole warp)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> registers -> bias/activation/residual -> global)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer full/empty (MMA <-> epilogue),
// static tile scheduler (tile = blockIdx.x + i*gridDim.x, n fastest so CTAs share the A rows in L2).
//
// K loop = taps x (C_in_pad / 64).  Per K block the producer loads a [128 rows, 64 ch] slab of the
// activation matrix at row coordinate m0 + tap_off0 + tap*tap_stride (negative / past-the-end rows
// are zero-filled by TMA) and a [BLOCK_N, 64] slab of the tap's weight matrix.  In SPLIT mode each
// operand is an fp16 pair x ~= hi + lo*2^-11 (common.cuh::split_op16) and every K step issues
// hi*hi into the main TMEM accumulator and lo*hi + hi*lo into a correction accumulator that the
// epilogue rescales by 2^-11 (lo*lo is 2^-24 relative and dropped): fp32-faithful products on the
// 16-bit tensor-core path.  SPLIT accumulation runs in short chains that the epilogue warps add up
// in registers, because tcgen05 truncates (not rounds) when it adds into TMEM.
#include <cstdlib>
#include <mutex>
#include <vector>

#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace jb {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
#ifndef JB_SPLIT_CHUNK
#define JB_SPLIT_CHUNK 8
#endif
static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // bf16 elements = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int kThreads = 192;

struct KernelParams {
  int taps, k_chunks, n_pad;
  int tap_off0, tap_stride;
  int n;  // real output columns (GLU: outputs)
  int m_rows;
  int num_m_tiles, num_n_tiles;
  const uint8_t* frame_mask;
  int rate, out_rows;
  int up_s, up_p, up_cout;
  ConvGemmEpilogue ep;
};

template <int BLOCK_N, bool SPLIT>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int MAX_STAGES = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  // accumulator buffer = [main | correction] in SPLIT mode; two buffers (MMA <-> epilogue ping-pong)
  static constexpr int ACC_COLS = SPLIT ? 2 * BLOCK_N : BLOCK_N;
  static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  // K blocks per tensor-core accumulation chain; 0 = one chain over all of K.  The fp32-faithful
  // (SPLIT) GEMMs keep chains short and add the partial sums on CUDA cores (see the epilogue).
  static constexpr int CHUNK = SPLIT ? JB_SPLIT_CHUNK : 0;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_LRELU) return v > 0.0f ? v : v * slope;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

// Finish one row segment of NV consecutive output columns starting at column `col` of output row
// `orow`; `v` already holds act(acc+bias)*scale (or the GLU product).
template <int NV>
__device__ __forceinline__ void store_row_segment(const ConvGemmEpilogue& ep, float (&v)[NV], long long orow,
                                                  int col, int n_limit) {
  const bool full = (col + NV <= n_limit);
  if (ep.res_f32) {
    const float* r = ep.res_f32 + orow * ep.res_ld + col;
    if (NV % 4 == 0 && full && (ep.res_ld & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(r + i);
        v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) v[i] += r[i];
    }
  }
  if (ep.res_bf16) {
    const bf16* r = ep.res_bf16 + orow * ep.res_ld + col;
    if (NV % 8 == 0 && full && (ep.res_ld & 7) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        uint4 t = *reinterpret_cast<const uint4*>(r + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          v[i + 2 * j] += f.x; v[i + 2 * j + 1] += f.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) v[i] += __bfloat162float(r[i]);
    }
  }
  if (ep.accum_in) {
    const float* r = ep.accum_in + orow * ep.out_f32_ld + col;
    if (NV % 4 == 0 && full && (ep.out_f32_ld & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(r + i);
        v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) v[i] += r[i];
    }
  }
  if (ep.accum_bf16) {
    const bf16* r = ep.accum_bf16 + orow * ep.res_ld + col;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (col + i < n_limit) v[i] += __bfloat162float(r[i]);
  }
  if (ep.post_scale != 1.0f) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= ep.post_scale;
  }
  if (ep.out_f32) {
    float* o = ep.out_f32 + orow * ep.out_f32_ld + col;
    if (NV % 4 == 0 && full && (ep.out_f32_ld & 3) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) o[i] = v[i];
    }
  }
  if (ep.out_hi) {
    // out_lo given: fp16 split operand pair (common.cuh::split_op16); otherwise a single bf16 copy
    bf16* oh = ep.out_hi + orow * ep.out_bf_ld + col;
    bf16* ol = ep.out_lo ? ep.out_lo + orow * ep.out_bf_ld + col : nullptr;
    if (NV % 8 == 0 && full && (ep.out_bf_ld & 7) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = v[i + 2 * j], b = v[i + 2 * j + 1];
          bf16 ah, bh, al, bl;
          if (ol) {
            split_op16(a, ah, al);
            split_op16(b, bh, bl);
            __nv_bfloat162 ll = __halves2bfloat162(al, bl);
            pl[j] = *reinterpret_cast<uint32_t*>(&ll);
          } else {
            ah = __float2bfloat16_rn(a);
            bh = __float2bfloat16_rn(b);
          }
          __nv_bfloat162 hh = __halves2bfloat162(ah, bh);
          ph[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(oh + i) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (ol) *reinterpret_cast<uint4*>(ol + i) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) {
          if (ol) {
            bf16 h, l;
            split_op16(v[i], h, l);
            oh[i] = h;
            ol[i] = l;
          } else {
            oh[i] = __float2bfloat16_rn(v[i]);
          }
        }
    }
  }
  if (ep.out_act) {
    bf16* oa = ep.out_act + orow * ep.out_act_ld + col;
    const float s = ep.out_act_slope;
    if (NV % 8 == 0 && full && (ep.out_act_ld & 7) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        uint32_t pa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = v[i + 2 * j], b = v[i + 2 * j + 1];
          pa[j] = pack_bf16x2(a > 0.f ? a : a * s, b > 0.f ? b : b * s);
        }
        *reinterpret_cast<uint4*>(oa + i) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (col + i < n_limit) {
          float a = v[i];
          oa[i] = __float2bfloat16_rn(a > 0.f ? a : a * s);
        }
    }
  }
}

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                    const __grid_constant__ KernelParams P) {
  using C = Cfg<BLOCK_N, SPLIT>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = P.num_m_tiles * P.num_n_tiles;
  const int k_iters = P.taps * P.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_b_hi);
    if (SPLIT) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_b_lo);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(C::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.num_n_tiles) * BLOCK_M;
        const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
        for (int tap = 0; tap < P.taps; ++tap) {
          const int arow = m0 + P.tap_off0 + tap * P.tap_stride;
          const int brow = tap * P.n_pad + n0;
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* s = smem + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_2d(&tm_a_hi, &full_bar[stage], s, kc * BLOCK_K, arow);
            tma_load_2d(&tm_b_hi, &full_bar[stage], s + C::A_BYTES, kc * BLOCK_K, brow);
            if (SPLIT) {
              tma_load_2d(&tm_a_lo, &full_bar[stage], s + C::A_BYTES + C::B_BYTES, kc * BLOCK_K, arow);
              tma_load_2d(&tm_b_lo, &full_bar[stage], s + 2 * C::A_BYTES + C::B_BYTES, kc * BLOCK_K, brow);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, /*is_bf16=*/!SPLIT);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int it = 0;
        while (it < k_iters) {
          // one accumulation chain = CHUNK K blocks (all of K when CHUNK == 0); see Cfg::CHUNK
          const int it_begin = it;
          const int it_end = (C::CHUNK > 0 && it + C::CHUNK < k_iters) ? it + C::CHUNK : k_iters;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * C::ACC_COLS);
          const uint32_t tmem_c = tmem_d + BLOCK_N;  // SPLIT: lo*hi + hi*lo products (scaled by 2^11)
          for (; it < it_end; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint64_t da_hi = make_sw128_desc(sa);
            const uint64_t db_hi = make_sw128_desc(sa + C::A_BYTES);
            const uint64_t da_lo = make_sw128_desc(sa + C::A_BYTES + C::B_BYTES);
            const uint64_t db_lo = make_sw128_desc(sa + 2 * C::A_BYTES + C::B_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * UMMA_K * 2) >> 4);  // +32 B per K step
              tc_mma_bf16(tmem_d, da_hi + koff, db_hi + koff, idesc, (it != it_begin || k != 0) ? 1u : 0u);
              if (SPLIT) {
                tc_mma_bf16(tmem_c, da_lo + koff, db_hi + koff, idesc, (it != it_begin || k != 0) ? 1u : 0u);
                tc_mma_bf16(tmem_c, da_hi + koff, db_lo + koff, idesc, 1u);
              }
            }
            tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull_bar[acc]);      // chain complete -> epilogue warps
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int lane_group = warp & 3;  // TMEM lanes [32*lane_group, +32) are the ones this warp may read
    const ConvGemmEpilogue& ep = P.ep;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int n_chains = C::CHUNK > 0 ? (k_iters + C::CHUNK - 1) / C::CHUNK : 1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / P.num_n_tiles) * BLOCK_M;
      const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
      const int row = m0 + lane_group * 32 + lane;
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
      // Chained mode: sum the partial accumulators in registers with round-to-nearest fp32 adds.
      // (tcgen05 truncates when it adds into the TMEM accumulator; measured bias 1.7e-5 relative
      // after 288 MMAs -- too much for the 1e-3 mel budget, so chains are kept short.)
      float accr[C::CHUNK > 0 ? BLOCK_N : 1];
      if (C::CHUNK > 0) {
        for (int ch = 0; ch < n_chains; ++ch) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t r[32], rc[32];
            tmem_ld32(lane_addr + static_cast<uint32_t>(acc * C::ACC_COLS + c), r);
            tmem_ld32(lane_addr + static_cast<uint32_t>(acc * C::ACC_COLS + BLOCK_N + c), rc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float x = fmaf(__uint_as_float(rc[i]), 1.0f / kSplitScale, __uint_as_float(r[i]));
              const int idx = C::CHUNK > 0 ? c + i : 0;
              accr[idx] = (ch == 0) ? x : accr[idx] + x;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      }
      const uint32_t taddr = lane_addr + static_cast<uint32_t>(acc * C::ACC_COLS);
      if (ep.act == ACT_GLU) {
        // tile columns [0, BLOCK_N/2) hold the linear half, [BLOCK_N/2, BLOCK_N) the gate half of the
        // same BLOCK_N/2 output channels (weights are interleaved per tile on the host).
        constexpr int HALF = BLOCK_N / 2;
        bool valid = row < P.m_rows;
        if (valid && P.frame_mask) valid = P.frame_mask[row / P.rate] != 0;
#pragma unroll(C::CHUNK > 0 ? HALF / 32 : 1)
        for (int c = 0; c < HALF; c += 32) {
          uint32_t ra[32], rb[32];
          if (C::CHUNK == 0) {
            tmem_ld32(taddr + c, ra);
            tmem_ld32(taddr + HALF + c, rb);
            tmem_ld_wait();
          }
          if (valid) {
            float v[32];
            const int ocol = n0 / 2 + c;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float xa = C::CHUNK > 0 ? accr[(C::CHUNK > 0 ? c + i : 0)] : __uint_as_float(ra[i]);
              const float xg = C::CHUNK > 0 ? accr[(C::CHUNK > 0 ? HALF + c + i : 0)] : __uint_as_float(rb[i]);
              float a = xa + (ep.bias ? __ldg(ep.bias + n0 + c + i) : 0.f);
              float g = xg + (ep.bias ? __ldg(ep.bias + n0 + HALF + c + i) : 0.f);
              v[i] = a * (1.0f / (1.0f + __expf(-g))) * ep.scale;
            }
            store_row_segment<32>(ep, v, row, ocol, P.n);
          }
        }
      } else {
#pragma unroll(C::CHUNK > 0 ? BLOCK_N / 32 : 1)
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t r[32];
          if (C::CHUNK == 0) {
            tmem_ld32(taddr + c, r);
            tmem_ld_wait();
          }
          const int ncol = n0 + c;  // column in GEMM-N space
          long long orow = row;
          int ocol = ncol;
          int n_limit = P.n;
          if (P.up_s > 0) {
            const int q = ncol / P.up_cout;
            orow = static_cast<long long>(row) * P.up_s + q - P.up_p;
            ocol = ncol - q * P.up_cout;
            n_limit = (ncol < P.n) ? P.up_cout : 0;
          }
          bool valid = row < P.m_rows && orow >= 0 && orow < P.out_rows && ocol < n_limit;
          if (valid && P.frame_mask) valid = P.frame_mask[orow / P.rate] != 0;
          if (valid) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = C::CHUNK > 0 ? accr[(C::CHUNK > 0 ? c + i : 0)] : __uint_as_float(r[i]);
              if (ep.bias) x += __ldg(ep.bias + (P.up_s > 0 ? ocol + i : ncol + i < P.n ? ncol + i : 0));
              v[i] = apply_act(x, ep.act, ep.slope) * ep.scale;
            }
            store_row_segment<32>(ep, v, orow, ocol, n_limit);
          }
        }
      }
      if (C::CHUNK == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
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
// host side
// ------------------------------------------------------------------------------------------------
bool g_profile_on = false;
std::vector<ProfileEvent> g_profile_events;

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, bool SPLIT>
static int launch(const ConvGemmProblem& p, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, SPLIT>;
  static_assert(C::STAGES >= 2, "need at least a double buffer");
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  JB_PROPAGATE(make_tmap(&ta_hi, p.a_hi, p.a_rows, a_cols, p.a_ld, BLOCK_M));
  JB_PROPAGATE(make_tmap(&tb_hi, p.w_hi, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, BLOCK_N));
  if (SPLIT) {
    JB_PROPAGATE(make_tmap(&ta_lo, p.a_lo, p.a_rows, a_cols, p.a_ld, BLOCK_M));
    JB_PROPAGATE(make_tmap(&tb_lo, p.w_lo, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, BLOCK_N));
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }
  KernelParams kp;
  kp.taps = p.taps;
  kp.k_chunks = p.k_pad / BLOCK_K;
  kp.n_pad = p.n_pad;
  kp.tap_off0 = p.tap_off0;
  kp.tap_stride = p.tap_stride;
  kp.n = p.n;
  kp.m_rows = p.m_rows;
  kp.num_m_tiles = ceil_div(p.m_rows, BLOCK_M);
  kp.num_n_tiles = p.n_pad / BLOCK_N;
  kp.frame_mask = p.frame_mask;
  kp.rate = p.rate > 0 ? p.rate : 1;
  kp.out_rows = p.out_rows;
  kp.up_s = p.up_s;
  kp.up_p = p.up_p;
  kp.up_cout = p.up_cout > 0 ? p.up_cout : 1;
  kp.ep = p.ep;
  auto kern = conv_gemm_tc_kernel<BLOCK_N, SPLIT>;
  JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
  const int tiles = kp.num_m_tiles * kp.num_n_tiles;
  if (tiles == 0) return 0;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile_on) {  // bench.py's roofline leg: device time of every launch of this kernel
    JB_CUDA_OK(cudaEventCreate(&e0));
    JB_CUDA_OK(cudaEventCreate(&e1));
    JB_CUDA_OK(cudaEventRecord(e0, stream));
  }
  kern<<<grid, kThreads, C::SMEM_BYTES, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, kp);
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, SPLIT ? 1 : 0});
  }
  return 0;
}

static int validate(const ConvGemmProblem& p) {
  JB_REQUIRE(p.a_hi && p.w_hi, -2, "conv_gemm: null operand");
  JB_REQUIRE(p.k_pad > 0 && p.k_pad % BLOCK_K == 0, -2, "conv_gemm: k_pad must be a multiple of 64");
  JB_REQUIRE(p.a_cols >= 0 && p.a_cols <= p.k_pad && p.a_ld >= (p.a_cols > 0 ? p.a_cols : p.k_pad), -2,
             "conv_gemm: need a_cols <= k_pad and a_ld >= a_cols");
  JB_REQUIRE(p.block_n == 32 || p.block_n == 64 || p.block_n == 128 || p.block_n == 256, -2,
             "conv_gemm: block_n must be 32/64/128/256");
  JB_REQUIRE(p.n_pad > 0 && p.n_pad % p.block_n == 0, -2, "conv_gemm: n_pad must be a multiple of block_n");
  JB_REQUIRE(p.taps >= 1, -2, "conv_gemm: taps");
  JB_REQUIRE((p.a_lo == nullptr) == (p.w_lo == nullptr), -2, "conv_gemm: split mode needs both a_lo and w_lo");
  if (p.ep.act == ACT_GLU) {
    JB_REQUIRE(p.block_n >= 64 && p.up_s == 0, -2, "conv_gemm: GLU needs block_n >= 64 and no upsampling");
    JB_REQUIRE(p.n * 2 <= p.n_pad, -2, "conv_gemm: GLU weights must hold 2n columns");
  } else {
    JB_REQUIRE(p.n <= p.n_pad, -2, "conv_gemm: n > n_pad");
  }
  if (p.up_s > 0) JB_REQUIRE(p.up_cout % 32 == 0 && p.n % p.up_cout == 0, -2, "conv_gemm: upsample needs C_out % 32 == 0");
  JB_REQUIRE(p.ep.out_f32 || p.ep.out_hi || p.ep.out_act, -2, "conv_gemm: no output requested");
  if (p.ep.accum_in) JB_REQUIRE(p.ep.out_f32_ld > 0, -2, "conv_gemm: accum_in uses out_f32_ld");
  return 0;
}

int conv_gemm_tc(const ConvGemmProblem& p, cudaStream_t stream) {
  JB_PROPAGATE(validate(p));
  static const bool no_tc2 = getenv("JATTS_B200_NO_TC2") != nullptr;  // A/B switch for profiling
  if (!no_tc2 && conv_gemm_tc2_eligible(p)) return conv_gemm_tc2(p, stream);
  JB_REQUIRE(p.ep.res_inv_slope == 0.f, -1, "res_inv_slope is only implemented by the TMA-epilogue kernel");
  static const bool no_tc3 = getenv("JATTS_B200_NO_TC3") != nullptr;
  if (!no_tc3 && conv_gemm_tc3_eligible(p)) return conv_gemm_tc3(p, stream);
  const bool split = p.a_lo != nullptr;
  switch (p.block_n) {
    case 32: return split ? launch<32, true>(p, stream) : launch<32, false>(p, stream);
    case 64: return split ? launch<64, true>(p, stream) : launch<64, false>(p, stream);
    case 128: return split ? launch<128, true>(p, stream) : launch<128, false>(p, stream);
    case 256:
      JB_REQUIRE(!split, -2, "conv_gemm: split mode supports block_n <= 128");
      return launch<256, false>(p, stream);
  }
  return -2;
}

// ------------------------------------------------------------------------------------------------
// CUDA-core debug twin (test-only; see header)
// ------------------------------------------------------------------------------------------------
__global__ void conv_gemm_simt_kernel(const bf16* __restrict__ a_hi, const bf16* __restrict__ a_lo, int a_rows, int a_ld,
                                      const bf16* __restrict__ w_hi, const bf16* __restrict__ w_lo, int k_pad,
                                      int a_cols, int block_n, KernelParams P) {
  const int ncol = blockIdx.x * blockDim.x + threadIdx.x;  // GEMM-N column
  const int row = blockIdx.y;
  const bool glu = P.ep.act == ACT_GLU;
  const int n_cols = glu ? P.n : P.n;
  if (ncol >= n_cols || row >= P.m_rows) return;
  auto dot = [&](int wcol) {
    float acc = 0.f;
    for (int tap = 0; tap < P.taps; ++tap) {
      const long long ar = static_cast<long long>(row) + P.tap_off0 + tap * P.tap_stride;
      if (ar < 0 || ar >= a_rows) continue;
      const bf16* ah = a_hi + ar * a_ld;
      const bf16* al = a_lo ? a_lo + ar * a_ld : nullptr;
      const long long wr = (static_cast<long long>(tap) * P.n_pad + wcol) * k_pad;
      if (al) {  // fp16 split pairs: x = hi + lo * 2^-11 (common.cuh::split_op16)
        const __half* xh = reinterpret_cast<const __half*>(ah);
        const __half* xl = reinterpret_cast<const __half*>(al);
        const __half* wh = reinterpret_cast<const __half*>(w_hi) + wr;
        const __half* wl = reinterpret_cast<const __half*>(w_lo) + wr;
        for (int c = 0; c < a_cols; ++c) {
          const float x = __half2float(xh[c]) + __half2float(xl[c]) * (1.0f / kSplitScale);
          const float w = __half2float(wh[c]) + __half2float(wl[c]) * (1.0f / kSplitScale);
          acc += x * w;
        }
      } else {
        for (int c = 0; c < a_cols; ++c) acc += __bfloat162float(ah[c]) * __bfloat162float(w_hi[wr + c]);
      }
    }
    return acc;
  };
  const ConvGemmEpilogue& ep = P.ep;
  float v[1];
  long long orow = row;
  int ocol = ncol;
  int n_limit = P.n;
  if (glu) {
    const int half = block_n / 2;
    const int t = ncol / half, i = ncol % half;  // output column ncol lives in weight tile t
    const int ca = t * block_n + i, cb = ca + half;
    float a = dot(ca) + (ep.bias ? ep.bias[ca] : 0.f);
    float g = dot(cb) + (ep.bias ? ep.bias[cb] : 0.f);
    v[0] = a * (1.0f / (1.0f + __expf(-g))) * ep.scale;
  } else {
    if (P.up_s > 0) {
      const int q = ncol / P.up_cout;
      orow = static_cast<long long>(row) * P.up_s + q - P.up_p;
      ocol = ncol - q * P.up_cout;
      n_limit = P.up_cout;
    }
    float x = dot(ncol);
    if (ep.bias) x += ep.bias[P.up_s > 0 ? ocol : ncol];
    v[0] = apply_act(x, ep.act, ep.slope) * ep.scale;
  }
  bool valid = orow >= 0 && orow < P.out_rows;
  if (valid && P.frame_mask) valid = P.frame_mask[orow / P.rate] != 0;
  if (!valid) return;
  store_row_segment<1>(ep, v, orow, ocol, n_limit);
}

int conv_gemm_simt_debug(const ConvGemmProblem& p, cudaStream_t stream) {
  JB_PROPAGATE(validate(p));
  KernelParams kp;
  kp.taps = p.taps;
  kp.k_chunks = p.k_pad / BLOCK_K;
  kp.n_pad = p.n_pad;
  kp.tap_off0 = p.tap_off0;
  kp.tap_stride = p.tap_stride;
  kp.n = p.n;
  kp.m_rows = p.m_rows;
  kp.num_m_tiles = 0;
  kp.num_n_tiles = 0;
  kp.frame_mask = p.frame_mask;
  kp.rate = p.rate > 0 ? p.rate : 1;
  kp.out_rows = p.out_rows;
  kp.up_s = p.up_s;
  kp.up_p = p.up_p;
  kp.up_cout = p.up_cout > 0 ? p.up_cout : 1;
  kp.ep = p.ep;
  if (p.m_rows == 0) return 0;
  dim3 grid(ceil_div(p.n, 128), p.m_rows);
  conv_gemm_simt_kernel<<<grid, 128, 0, stream>>>(p.a_hi, p.a_lo, p.a_rows, p.a_ld, p.w_hi, p.w_lo, p.k_pad,
                                                  p.a_cols > 0 ? p.a_cols : p.k_pad, p.block_n, kp);
  JB_KERNEL_OK();
  return 0;
}

}  // namespace jb
