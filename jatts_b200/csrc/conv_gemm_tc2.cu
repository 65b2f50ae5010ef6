// conv_bf16_tma_kernel: the HiFi-GAN convolution kernel (bf16 operands, fp32 TMEM accumulation) with a
// TMA-staged epilogue and three mainloop modes:
//   STREAM    one [128 x KCH] activation tile + one [N x KCH] weight tile per (tap, K chunk) stage
//   HALO      one activation SLAB per K chunk = tile rows + the taps' halo; each tap is an MMA whose A
//             descriptor starts tap*dilation rows into the slab (128B/64B swizzle is a function of the
//             absolute smem address, so row-shifted starts need no base_offset -- measured); weights
//             stream through their own ring
//   RESIDENT  HALO + all taps' weights loaded into smem once per CTA: no per-tap barrier traffic at
//             all (the single-thread MMA issuer was the bottleneck for C <= 64: 88 clk per 16-clk MMA)
// Epilogue (all modes):
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
#include <cstdlib>

#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace jb {

static constexpr int BLOCK_M2 = 128;
static constexpr int UMMA_K2 = 16;
static constexpr int kThreads2 = 256;
enum : int { MODE_STREAM = 0, MODE_HALO = 1, MODE_RESIDENT = 2 };

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
  int halo_rows;    // HALO/RESIDENT: rows of the activation box (128 + (taps-1)*tap_stride, padded to 8)
};

template <int BLOCK_N, int KCH, int MODE>
struct Cfg2 {
  static constexpr int SLAB = BLOCK_N >= 64 ? 64 : 32;     // columns per epilogue slab
  static constexpr int N_SLABS = BLOCK_N / SLAB;
  static constexpr int ROWB = SLAB * 2;                     // bytes per slab row = swizzle span
  static constexpr int SW_MASK = ROWB == 128 ? 7 : 3;       // Swizzle<3,4,3> / Swizzle<2,4,3>
  static constexpr int SLAB_BYTES = BLOCK_M2 * ROWB;
  static constexpr int ENTRY_BYTES = 2 * SLAB_BYTES;        // [residual -> out0 | branch sum -> out1]
  static constexpr int EP_ENTRIES = MODE == MODE_RESIDENT ? (BLOCK_N == 32 ? 4 : 2)
                                                          : (BLOCK_N == 256 ? 2 : (BLOCK_N == 32 ? 4 : 3));
  static constexpr int KROWB = KCH * 2;                     // operand row bytes (128 or 64) = swizzle span
  static constexpr int A_BYTES = BLOCK_M2 * KROWB;
  static constexpr int B_BYTES = BLOCK_N * KROWB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2048;                   // n_pad <= 512
  static constexpr int BAR_BYTES = 1024;                    // keeps the resident weight area 1024-aligned
  static constexpr int FIXED_TAIL = EP_ENTRIES * ENTRY_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
  static constexpr int BUDGET = 225 * 1024 - FIXED_TAIL;
  // STREAM
  static constexpr int MAX_STAGES = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  // HALO / RESIDENT: activation slabs [HALO_MAX_ROWS x KCH]
  static constexpr int HALO_MAX_ROWS = 192;                 // 128 + (11-1)*5 = 178, padded
  static constexpr int A_SLAB_BYTES = HALO_MAX_ROWS * KROWB;  // multiple of 1024
  static constexpr int A_STAGES = MODE == MODE_RESIDENT ? (BLOCK_N == 32 ? 4 : 2) : (BLOCK_N == 256 ? 2 : 3);
  static constexpr int B_MAX = (BUDGET - A_STAGES * A_SLAB_BYTES) / B_BYTES;
  static constexpr int B_STAGES = B_MAX > 8 ? 8 : B_MAX;
  static constexpr int MAIN_BYTES = MODE == MODE_STREAM ? STAGES * STAGE_BYTES
                                  : MODE == MODE_HALO   ? A_STAGES * A_SLAB_BYTES + B_STAGES * B_BYTES
                                                        : A_STAGES * A_SLAB_BYTES;
  static constexpr int SMEM_FIXED = MAIN_BYTES + FIXED_TAIL;   // + resident weight bytes in RESIDENT mode
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert(MODE != MODE_STREAM || STAGES >= 2, "stream pipeline depth");
  static_assert(MODE != MODE_HALO || B_STAGES >= 2, "weight ring depth");
  static_assert(A_STAGES <= 4 && STAGES <= 8 && B_STAGES <= 8, "barrier arrays");
};

// K-major operand tile with rows of exactly one swizzle span (128 B: KCH = 64, 64 B: KCH = 32)
template <int KCH>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((8 * KCH * 2) >> 4) << 32;    // SBO: 8 rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(KCH == 64 ? 2 : 4) << 61;     // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

template <int BLOCK_N, int KCH, int MODE>
__global__ void __launch_bounds__(kThreads2, 1)
conv_bf16_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_acc,
                     const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1,
                     const __grid_constant__ KernelParams2 P) {
  using C = Cfg2<BLOCK_N, KCH, MODE>;
  constexpr int STAGES = C::STAGES;
  constexpr int E = C::EP_ENTRIES;
  constexpr int KSTEPS = KCH / UMMA_K2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_base = smem + C::A_STAGES * C::A_SLAB_BYTES;             // HALO: weight ring after the A slabs
  uint8_t* ep_base = smem + C::MAIN_BYTES;                            // 1024-aligned (all sizes are multiples)
  float* bias_s = reinterpret_cast<float*>(ep_base + E * C::ENTRY_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + C::BIAS_BYTES);
  uint8_t* w_base = reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES;  // RESIDENT: all weight tiles (1024-aligned)
  constexpr int NB = 8;
  uint64_t* full_bar = bars;                    // [NB]   STREAM: A+B stage | HALO: weight ring
  uint64_t* empty_bar = full_bar + NB;          // [NB]
  uint64_t* afull_bar = empty_bar + NB;         // [4]    HALO/RESIDENT: activation slabs
  uint64_t* aempty_bar = afull_bar + 4;         // [4]
  uint64_t* tfull_bar = aempty_bar + 4;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* epfull_bar = tempty_bar + 2;        // [E]
  uint64_t* epempty_bar = epfull_bar + E;       // [E]
  uint64_t* ready_bar = epempty_bar + E;        // [E]    epilogue -> store warp: slab written
  uint64_t* wfull_bar = ready_bar + E;          // [1]    RESIDENT: weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = P.num_m_tiles * P.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (P.has_res) tma_prefetch_desc(&tm_res);
    if (P.has_acc) tma_prefetch_desc(&tm_acc);
    if (P.has_out0) tma_prefetch_desc(&tm_out0);
    if (P.has_out1) tma_prefetch_desc(&tm_out1);
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    for (int i = 0; i < E; ++i) {
      mbar_init(&epfull_bar[i], 1);
      mbar_init(&epempty_bar[i], 1);
      mbar_init(&ready_bar[i], 128);
    }
    mbar_init(wfull_bar, 1);
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
      if (MODE == MODE_STREAM) {
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
              tma_load_2d(&tm_a, &full_bar[stage], s, kc * KCH, arow);
              tma_load_2d(&tm_b, &full_bar[stage], s + C::A_BYTES, kc * KCH, brow);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else {
        if (MODE == MODE_RESIDENT) {
          // every weight tile of this convolution, once per CTA (num_n_tiles == 1 in this mode)
          mbar_expect_tx(wfull_bar, static_cast<uint32_t>(P.taps * P.k_chunks) * C::B_BYTES);
          for (int kc = 0; kc < P.k_chunks; ++kc)
            for (int tap = 0; tap < P.taps; ++tap)
              tma_load_2d(&tm_b, wfull_bar, w_base + (kc * P.taps + tap) * C::B_BYTES, kc * KCH, tap * P.n_pad);
        }
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        const uint32_t a_bytes = static_cast<uint32_t>(P.halo_rows) * C::KROWB;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int m0 = (tile / P.num_n_tiles) * BLOCK_M2;
          const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(&aempty_bar[as], aph ^ 1);
            mbar_expect_tx(&afull_bar[as], a_bytes);
            tma_load_2d(&tm_a, &afull_bar[as], smem + as * C::A_SLAB_BYTES, kc * KCH, m0 + P.tap_off0);
            if (++as == C::A_STAGES) { as = 0; aph ^= 1; }
            if (MODE == MODE_HALO) {
              for (int tap = 0; tap < P.taps; ++tap) {
                mbar_wait(&empty_bar[bs], bph ^ 1);
                mbar_expect_tx(&full_bar[bs], C::B_BYTES);
                tma_load_2d(&tm_b, &full_bar[bs], b_base + bs * C::B_BYTES, kc * KCH, tap * P.n_pad + n0);
                if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M2, BLOCK_N, /*is_bf16=*/true);
      int acc = 0;
      uint32_t acc_phase = 0;
      if (MODE == MODE_STREAM) {
        const int k_iters = P.taps * P.k_chunks;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
          for (int it = 0; it < k_iters; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint64_t da = make_kmajor_desc<KCH>(sa);
            const uint64_t db = make_kmajor_desc<KCH>(sa + C::A_BYTES);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * UMMA_K2 * 2) >> 4);
              tc_mma_bf16(tmem_d, da + koff, db + koff, idesc, (it | k) != 0 ? 1u : 0u);
            }
            tc_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      } else {
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        if (MODE == MODE_RESIDENT) {
          mbar_wait(wfull_bar, 0);
          tc_fence_after();
        }
        const uint32_t w_addr = smem_u32(w_base);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(&afull_bar[as], aph);
            tc_fence_after();
            const uint32_t slab = smem_u32(smem + as * C::A_SLAB_BYTES);
            for (int tap = 0; tap < P.taps; ++tap) {
              uint32_t b_addr;
              if (MODE == MODE_HALO) {
                mbar_wait(&full_bar[bs], bph);
                tc_fence_after();
                b_addr = smem_u32(b_base + bs * C::B_BYTES);
              } else {
                b_addr = w_addr + static_cast<uint32_t>(kc * P.taps + tap) * C::B_BYTES;
              }
              // tap t reads activation rows [t*dilation, t*dilation + 128) of the slab: a row-shifted
              // start address.  Measured on B200: the swizzle is a function of the absolute smem address
              // bits, so a start that is not 1024 B aligned needs NO base_offset in the descriptor.
              const uint32_t a_addr = slab + static_cast<uint32_t>(tap * P.tap_stride) * C::KROWB;
              const uint64_t da = make_kmajor_desc<KCH>(a_addr);
              const uint64_t db = make_kmajor_desc<KCH>(b_addr);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t koff = static_cast<uint64_t>((k * UMMA_K2 * 2) >> 4);
                tc_mma_bf16(tmem_d, da + koff, db + koff, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
              if (MODE == MODE_HALO) {
                tc_commit(&empty_bar[bs]);
                if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
              }
            }
            tc_commit(&aempty_bar[as]);
            if (++as == C::A_STAGES) { as = 0; aph ^= 1; }
          }
          tc_commit(&tfull_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
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
  } else if (warp == 3) {
    // ===================== store warp: slab -> global by TMA =====================
    if (elect_one()) {
      int e = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / P.num_n_tiles) * BLOCK_M2;
        const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
        for (int s = 0; s < C::N_SLABS; ++s) {
          mbar_wait(&ready_bar[e], ph);      // all 128 epilogue threads wrote + fenced this slab
          uint8_t* bufA = ep_base + e * C::ENTRY_BYTES;
          if (P.has_out0) tma_store_2d(&tm_out0, bufA, n0 + s * C::SLAB, m0);
          if (P.has_out1) tma_store_2d(&tm_out1, bufA + C::SLAB_BYTES, n0 + s * C::SLAB, m0);
          tma_store_commit();
          tma_store_wait_read<0>();          // smem has been read: the entry can be refilled
          mbar_arrive(&epempty_bar[e]);
          if (++e == E) { e = 0; ph ^= 1; }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..7) =====================
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t row_off = static_cast<uint32_t>(row_in_tile * C::ROWB);
    const uint32_t sw = (row_off >> 7) & C::SW_MASK;   // XOR pattern of this row's 16-byte chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    int e = 0;
    uint32_t ph = 0;
    // row validity of the NEXT tile is fetched while the current one is processed (one global byte per
    // thread per tile; its latency used to sit on the per-tile critical path)
    auto row_valid = [&](int tile) -> bool {
      if (tile >= num_tiles) return false;
      const int row = (tile / P.num_n_tiles) * BLOCK_M2 + row_in_tile;
      if (row >= P.m_rows) return false;
      return P.frame_mask ? P.frame_mask[row / P.rate] != 0 : true;
    };
    bool valid_next = row_valid(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile % P.num_n_tiles) * BLOCK_N;
      const bool valid = valid_next;
      valid_next = row_valid(tile + gridDim.x);
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
        mbar_arrive(&ready_bar[e]);          // hand the slab to the store warp; nobody waits here
        if (++e == E) { e = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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


template <int BLOCK_N, int KCH, int MODE>
static int launch2(const ConvGemmProblem& p, cudaStream_t stream) {
  using C = Cfg2<BLOCK_N, KCH, MODE>;
  const ConvGemmEpilogue& e = p.ep;
  CUtensorMap ta, tb, tres, tacc, to0, to1;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  const int k_chunks = ceil_div(a_cols, KCH);   // channel padding beyond a_cols is all-zero: skip it
  const int halo_rows = round_up(BLOCK_M2 + (p.taps - 1) * p.tap_stride, 8);
  JB_PROPAGATE(make_tmap(&ta, p.a_hi, p.a_rows, a_cols, p.a_ld, MODE == MODE_STREAM ? BLOCK_M2 : halo_rows, KCH));
  JB_PROPAGATE(make_tmap(&tb, p.w_hi, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, BLOCK_N, KCH));
  tres = tacc = to0 = to1 = ta;
  if (e.res_bf16) JB_PROPAGATE(make_tmap(&tres, e.res_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  if (e.accum_bf16) JB_PROPAGATE(make_tmap(&tacc, e.accum_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  if (e.out_hi) JB_PROPAGATE(make_tmap(&to0, e.out_hi, p.m_rows, p.n, e.out_bf_ld, BLOCK_M2, C::SLAB));
  if (e.out_act) JB_PROPAGATE(make_tmap(&to1, e.out_act, p.m_rows, p.n, e.out_act_ld, BLOCK_M2, C::SLAB));
  KernelParams2 kp;
  kp.taps = p.taps;
  kp.k_chunks = k_chunks;
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
  kp.halo_rows = halo_rows;
  const int w_bytes = MODE == MODE_RESIDENT ? p.taps * k_chunks * C::B_BYTES : 0;
  const int smem_bytes = C::SMEM_FIXED + w_bytes;
  JB_REQUIRE(smem_bytes <= 227 * 1024, -2, "conv_gemm_tc2: shared memory budget exceeded");
  auto kern = conv_bf16_tma_kernel<BLOCK_N, KCH, MODE>;
  static int attr_bytes = 0;
  if (smem_bytes > attr_bytes) {
    JB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_bytes = smem_bytes;
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
  kern<<<grid, kThreads2, smem_bytes, stream>>>(ta, tb, tres, tacc, to0, to1, kp);
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, 0});
  }
  return 0;
}

int conv_gemm_tc2(const ConvGemmProblem& p, cudaStream_t stream) {
  static const int max_mode = getenv("JATTS_B200_TC2_MODE") ? atoi(getenv("JATTS_B200_TC2_MODE")) : 2;  // A/B switch
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  const bool halo_ok = max_mode >= MODE_HALO && p.taps > 1 && p.tap_stride > 0 &&
                       round_up(BLOCK_M2 + (p.taps - 1) * p.tap_stride, 8) <= 192;
  // RESIDENT when every weight tile of the convolution fits next to the pipeline buffers
  auto resident_ok = [&](int fixed, int b_bytes, int kch) {
    return halo_ok && max_mode >= MODE_RESIDENT && p.n_pad == p.block_n &&
           fixed + p.taps * ceil_div(a_cols, kch) * b_bytes <= 227 * 1024;
  };
  switch (p.block_n) {
    case 32:
      if (a_cols <= 32 && resident_ok(Cfg2<32, 32, MODE_RESIDENT>::SMEM_FIXED, Cfg2<32, 32, MODE_RESIDENT>::B_BYTES, 32))
        return launch2<32, 32, MODE_RESIDENT>(p, stream);
      if (resident_ok(Cfg2<32, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<32, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<32, 64, MODE_RESIDENT>(p, stream);
      return launch2<32, 64, MODE_STREAM>(p, stream);
    case 64:
      if (resident_ok(Cfg2<64, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<64, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<64, 64, MODE_RESIDENT>(p, stream);
      return launch2<64, 64, MODE_STREAM>(p, stream);
    case 128:
      if (resident_ok(Cfg2<128, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<128, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<128, 64, MODE_RESIDENT>(p, stream);
      return halo_ok ? launch2<128, 64, MODE_HALO>(p, stream) : launch2<128, 64, MODE_STREAM>(p, stream);
    case 256:
      return halo_ok ? launch2<256, 64, MODE_HALO>(p, stream) : launch2<256, 64, MODE_STREAM>(p, stream);
  }
  return -2;
}

}  // namespace jb
