// gemm_split_tma_kernel: the FastSpeech2 GEMM / Conv1d kernel.  fp32-faithful products from fp16 split
// operand pairs (x ~= hi + lo*2^-11, common.cuh::split_op16) on tcgen05 tensor cores, with the TMA-staged
// epilogue of conv_gemm_tc2.cu.
//
//   warp 0   TMA producer.  Activations: per 64-wide K chunk ONE slab pair A_hi, A_lo of [128 + (taps-1)*stride
//            rows x 64] that serves every tap (tap t is an MMA whose A descriptor starts t*stride rows into
//            the slab), weights: per (K chunk, tap) one B_hi, B_lo [128 x 64] pair.  The kernel is bound by
//            the bytes an SM can take in (~40 B/clk measured across three kernels): loading the
//            activations once per K chunk instead of once per tap takes a k = 3 convolution from 85 to 56
//            bytes per MMA clock.
//   warp 1   MMA issuer: per K step hi*hi -> main accumulator, lo*hi + hi*lo -> correction accumulator
//            (both in TMEM, fp32); accumulation runs in CHAINS of 8 K blocks because tcgen05 truncates when
//            it adds into TMEM (measured bias 1.7e-5 after 288 MMAs) -- chains ping-pong between
//            two TMEM buffers
//   warp 2   epilogue loader: TMA-loads the fp32 residual slabs of upcoming tiles into the epilogue ring
//   warp 3   store warp: TMA stores of finished slabs (fp32 master and/or fp16 split pair)
//   warps 4-7 epilogue: drain every chain (tcgen05.ld) and add the partial sums in REGISTERS with
//            round-to-nearest fp32 (main + corr * 2^-11); after the last chain: bias, ReLU / tanh / GLU,
//            scale, fp32 residual (from the smem slab), row mask, outputs written in place into the
//            swizzled smem slabs, fence.proxy.async, mbarrier hand-off to the store warp.
//
// The first version of this path (conv_gemm_tc_kernel<128, true>) had the same mainloop but finished
// tiles with per-thread global loads/stores: measured 52 k clk to store one 128x128 tile against 10-14 k
// clk of MMA work (tools/gpu_trace1.py).
#include <cstdlib>

#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace jb {

extern long long* g_trace_ptr;

namespace {

constexpr int BM = 128;       // rows per tile (UMMA M)
constexpr int BN = 128;       // columns per tile (UMMA N)
constexpr int BK = 64;        // fp16 elements per K block = one 128-byte swizzle row
constexpr int UK = 16;
constexpr int kThreads3 = 256;
constexpr int CHUNK = 8;      // K blocks per accumulation chain
constexpr int A_STAGES = 2;
constexpr int MAX_B_STAGES = 6;                // 3 full stages, or 6 half stages when a CTA pair shares the weights
constexpr int OP_BYTES = BM * BK * 2;          // 16 KB: one weight tile (hi or lo)
constexpr int B_STAGE_BYTES = 2 * OP_BYTES;    // B_hi | B_lo
constexpr int A_SLAB_ROWS = 144;               // 128 + halo of the taps (<= 16 rows)
constexpr int A_OP_BYTES = A_SLAB_ROWS * BK * 2;   // 18 KB (multiple of 1024)
constexpr int A_STAGE_BYTES = 2 * A_OP_BYTES;  // A_hi | A_lo
constexpr int SLAB = 32;                        // columns per epilogue slab
constexpr int F32_SLAB_BYTES = BM * SLAB * 4;   // 16 KB, 128-byte rows (SWIZZLE_128B)
constexpr int H16_SLAB_BYTES = BM * SLAB * 2;   // 8 KB, 64-byte rows (SWIZZLE_64B)
constexpr int MAX_ENTRIES = 6;
constexpr int BIAS_BYTES = 8192;                // n_pad <= 2048
constexpr int BAR_BYTES = 1024;
constexpr int SMEM_MISC = BIAS_BYTES + BAR_BYTES + 1024;
constexpr int smem_fixed(int b_stages) { return A_STAGES * A_STAGE_BYTES + b_stages * B_STAGE_BYTES + SMEM_MISC; }
constexpr int TMEM_COLS = 512;                  // 2 buffers x (main 128 | corr 128)

struct Params3 {
  int taps, k_chunks, n_pad;
  int tap_off0, tap_stride;
  int n;            // real output columns (GLU: outputs = half of the weight columns used)
  int m_rows;
  int num_m_tiles, num_n_tiles;
  const uint8_t* frame_mask;
  const float* bias;
  int act;          // ACT_NONE / ACT_RELU / ACT_TANH / ACT_GLU
  float scale;
  int has_res, has_f32, has_split;
  int entries, entry_bytes;
  int b_stages;     // weight ring depth (2..3)
  int num_groups;   // pair mode: ceil(num_m_tiles / 2) * num_n_tiles (a pair walks 2 M tiles x one N tile at a time)
  int halo_rows;    // rows of the activation slab: 128 + (taps-1)*tap_stride rounded up to 8
  long long* trace;
};

#define JB_TRACE3(role, ev, idx)                                                                         \
  do {                                                                                                   \
    if (P.trace && blockIdx.x == 0 && (idx) < 64) P.trace[((role) * 8 + (ev)) * 64 + (idx)] = clock64(); \
  } while (0)

__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {  // K-major, 128-byte swizzle
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// number of 32-column output slabs of the tile that starts at GEMM column n0
__device__ __forceinline__ int tile_slabs(const Params3& P, int n0) {
  if (P.act == ACT_GLU) {
    const int o0 = n0 / 2;
    const int left = P.n - o0;
    return left >= 64 ? 2 : (left <= 0 ? 0 : (left + SLAB - 1) / SLAB);
  }
  const int left = P.n - n0;
  return left >= BN ? 4 : (left <= 0 ? 0 : (left + SLAB - 1) / SLAB);
}

// PAIR: the two CTAs of a cluster issue M = 256 cta_group::2 MMAs; each keeps its own 128 activation rows and HALF of
// every weight tile pair (a template parameter: kernels with cta_group::2 code need an even cluster size to launch)
template <bool PAIR>
__global__ void __launch_bounds__(kThreads3, 1)
gemm_split_tma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_f32,
                      const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                      const __grid_constant__ Params3 P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem + A_STAGES * A_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(b_ring + P.b_stages * B_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + BIAS_BYTES);
  uint8_t* ep_base = reinterpret_cast<uint8_t*>(bars) + BAR_BYTES;   // 1024-aligned
  uint64_t* full_bar = bars;                       // [MAX_B_STAGES]  weight ring
  uint64_t* empty_bar = full_bar + MAX_B_STAGES;   // [MAX_B_STAGES]
  uint64_t* afull_bar = empty_bar + MAX_B_STAGES;  // [A_STAGES]      activation slabs
  uint64_t* aempty_bar = afull_bar + A_STAGES;     // [A_STAGES]
  uint64_t* tfull_bar = aempty_bar + A_STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* epfull_bar = tempty_bar + 2;           // [MAX_ENTRIES]
  uint64_t* epempty_bar = epfull_bar + MAX_ENTRIES;
  uint64_t* ready_bar = epempty_bar + MAX_ENTRIES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready_bar + MAX_ENTRIES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile schedule: work unit u -> (M tile, N tile).  Single CTA: u = tile index.  Pair: u = group index, CTA `rank`
  // owns M tile 2 * (u / num_n_tiles) + rank (past the end: loads zero-fill, stores clip, rows are masked)
  constexpr int CL = PAIR ? 2 : 1;
  const int rank = PAIR ? static_cast<int>(blockIdx.x) & 1 : 0;
  const int u0 = static_cast<int>(blockIdx.x) / CL;
  const int ustep = static_cast<int>(gridDim.x) / CL;
  const int num_units = PAIR ? P.num_groups : P.num_m_tiles * P.num_n_tiles;
  constexpr int B_OP = PAIR ? OP_BYTES / 2 : OP_BYTES;          // one weight tile (hi or lo) as this CTA stores it
  constexpr int B_STG = 2 * B_OP;                                // B_hi | B_lo
  const int b_depth = PAIR ? 2 * P.b_stages : P.b_stages;
  const int k_iters = P.taps * P.k_chunks;
  const int n_chains = (k_iters + CHUNK - 1) / CHUNK;
  const int E = P.entries;
  const int f32_off = 0;
  const int hi_off = (P.has_res || P.has_f32) ? F32_SLAB_BYTES : 0;
  const int lo_off = hi_off + H16_SLAB_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    for (int i = 0; i < MAX_B_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 8 : 4); }
    for (int i = 0; i < MAX_ENTRIES; ++i) {
      mbar_init(&epfull_bar[i], 1);
      mbar_init(&epempty_bar[i], 1);
      mbar_init(&ready_bar[i], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(static_cast<uint32_t>(TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(static_cast<uint32_t>(TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  {
    const int n_bias = P.act == ACT_GLU ? 2 * P.n : P.n;   // entries the bias tensor really has
    for (int i = threadIdx.x; i < P.n_pad && i < BIAS_BYTES / 4; i += kThreads3)
      bias_s[i] = (P.bias && i < n_bias) ? P.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's mbarriers exist before any remote arrive / pair load
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlapped the previous kernel's tail; its outputs are visible from here

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      const uint32_t a_bytes = static_cast<uint32_t>(P.halo_rows) * BK * 2;
      for (int u = u0; u < num_units; u += ustep) {
        const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
        const int n0 = (u % P.num_n_tiles) * BN;
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(&aempty_bar[as], aphase ^ 1);
          uint8_t* a = smem + as * A_STAGE_BYTES;
          if constexpr (PAIR) {
            // both CTAs' slabs complete on the LEADER's barrier (only its MMA thread waits on it)
            const uint32_t lb = mapa_u32(smem_u32(&afull_bar[as]), 0);
            if (rank == 0) mbar_expect_tx(&afull_bar[as], 4 * a_bytes);
            tma_load_2d_pair(&tm_a_hi, lb, a, kc * BK, m0 + P.tap_off0);
            tma_load_2d_pair(&tm_a_lo, lb, a + A_OP_BYTES, kc * BK, m0 + P.tap_off0);
          } else {
            mbar_expect_tx(&afull_bar[as], 2 * a_bytes);
            tma_load_2d(&tm_a_hi, &afull_bar[as], a, kc * BK, m0 + P.tap_off0);
            tma_load_2d(&tm_a_lo, &afull_bar[as], a + A_OP_BYTES, kc * BK, m0 + P.tap_off0);
          }
          if (++as == A_STAGES) { as = 0; aphase ^= 1; }
          for (int tap = 0; tap < P.taps; ++tap) {
            const int brow = tap * P.n_pad + n0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* b = b_ring + stage * B_STG;
            if constexpr (PAIR) {
              // this CTA's half (rows rank * 64 ...) of the hi and lo weight tiles stays in its own shared memory
              const uint32_t lb = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_expect_tx(&full_bar[stage], B_STAGE_BYTES);
              tma_load_2d_pair(&tm_b_hi, lb, b, kc * BK, brow + rank * (BN / 2));
              tma_load_2d_pair(&tm_b_lo, lb, b + B_OP, kc * BK, brow + rank * (BN / 2));
            } else {
              mbar_expect_tx(&full_bar[stage], B_STAGE_BYTES);
              tma_load_2d(&tm_b_hi, &full_bar[stage], b, kc * BK, brow);
              tma_load_2d(&tm_b_lo, &full_bar[stage], b + B_OP, kc * BK, brow);
            }
            if (++stage == b_depth) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if ((!PAIR || rank == 0) && elect_one()) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 2 * BM : BM, BN, /*is_bf16=*/false);   // fp16 operands
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
        if constexpr (PAIR) tc_mma_f16_pair(d, da, db, idesc, acc);
        else tc_mma_bf16(d, da, db, idesc, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) tc_commit_pair(bar);
        else tc_commit(bar);
      };
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      int buf = 0;
      uint32_t buf_phase = 0;
      for (int u = u0, seq = 0; u < num_units; u += ustep, ++seq) {
        JB_TRACE3(1, 0, seq);
        int it = 0;          // (K chunk, tap) steps issued for this tile
        int in_chain = 0;    // steps accumulated into the current chain
        uint32_t tmem_d = 0, tmem_c = 0;
        for (int kc = 0; kc < P.k_chunks; ++kc) {
          mbar_wait(&afull_bar[as], aphase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + as * A_STAGE_BYTES);
          for (int tap = 0; tap < P.taps; ++tap, ++it) {
            if (in_chain == 0) {
              mbar_wait(&tempty_bar[buf], buf_phase ^ 1);
              tc_fence_after();
              tmem_d = tmem_base + static_cast<uint32_t>(buf * 2 * BN);   // main
              tmem_c = tmem_d + BN;                                        // correction
            }
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sb = smem_u32(b_ring + stage * B_STG);
            // tap t reads activation rows [t*stride, t*stride + 128) of the slab (row-shifted descriptor start)
            const uint32_t a_tap = sa + static_cast<uint32_t>(tap * P.tap_stride) * (BK * 2);
            const uint64_t da_hi = desc128(a_tap), db_hi = desc128(sb);
            const uint64_t da_lo = desc128(a_tap + A_OP_BYTES), db_lo = desc128(sb + B_OP);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t koff = static_cast<uint64_t>((k * UK * 2) >> 4);
              const uint32_t first = (in_chain != 0 || k != 0) ? 1u : 0u;
              mma(tmem_d, da_hi + koff, db_hi + koff, first);
              mma(tmem_c, da_lo + koff, db_hi + koff, first);
              mma(tmem_c, da_hi + koff, db_lo + koff, 1u);
            }
            commit(&empty_bar[stage]);
            if (++stage == b_depth) { stage = 0; phase ^= 1; }
            if (++in_chain == CHUNK || it + 1 == k_iters) {
              commit(&tfull_bar[buf]);
              if (++buf == 2) { buf = 0; buf_phase ^= 1; }
              in_chain = 0;
            }
          }
          commit(&aempty_bar[as]);
          if (++as == A_STAGES) { as = 0; aphase ^= 1; }
        }
        JB_TRACE3(1, 2, seq);
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue loader =====================
    if (elect_one()) {
      int e = 0;
      uint32_t ph = 0;
      for (int u = u0; u < num_units; u += ustep) {
        const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
        const int n0 = (u % P.num_n_tiles) * BN;
        const int ns = tile_slabs(P, n0);
        const int o0 = P.act == ACT_GLU ? n0 / 2 : n0;
        for (int s = 0; s < ns; ++s) {
          mbar_wait(&epempty_bar[e], ph ^ 1);
          if (P.has_res) {
            mbar_expect_tx(&epfull_bar[e], F32_SLAB_BYTES);
            tma_load_2d(&tm_res, &epfull_bar[e], ep_base + e * P.entry_bytes + f32_off, o0 + s * SLAB, m0);
          } else {
            mbar_arrive(&epfull_bar[e]);
          }
          if (++e == E) { e = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== store warp =====================
    if (elect_one()) {
      int e = 0;
      uint32_t ph = 0;
      for (int u = u0; u < num_units; u += ustep) {
        const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
        const int n0 = (u % P.num_n_tiles) * BN;
        const int ns = tile_slabs(P, n0);
        const int o0 = P.act == ACT_GLU ? n0 / 2 : n0;
        for (int s = 0; s < ns; ++s) {
          mbar_wait(&ready_bar[e], ph);
          uint8_t* buf = ep_base + e * P.entry_bytes;
          if (P.has_f32) tma_store_2d(&tm_f32, buf + f32_off, o0 + s * SLAB, m0);
          if (P.has_split) {
            tma_store_2d(&tm_hi, buf + hi_off, o0 + s * SLAB, m0);
            tma_store_2d(&tm_lo, buf + lo_off, o0 + s * SLAB, m0);
          }
          tma_store_commit();
          tma_store_wait_read<0>();          // smem has been read: the entry can be refilled
          mbar_arrive(&epempty_bar[e]);
          if (++e == E) { e = 0; ph ^= 1; }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..7) =====================
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t f_row = static_cast<uint32_t>(row_in_tile) * 128u;        // fp32 slab row (128 B)
    const uint32_t f_sw = static_cast<uint32_t>(row_in_tile) & 7u;
    const uint32_t h_row = static_cast<uint32_t>(row_in_tile) * 64u;         // fp16 slab row (64 B)
    const uint32_t h_sw = (h_row >> 7) & 3u;
    int buf = 0;
    uint32_t buf_phase = 0;
    int e = 0;
    uint32_t ph = 0;
    for (int u = u0, seq = 0; u < num_units; u += ustep, ++seq) {
      const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
      const int n0 = (u % P.num_n_tiles) * BN;
      const int row = m0 + row_in_tile;
      unsigned mask_byte = 1u;   // loaded now, compared after the drains (off the critical path)
      if (row < P.m_rows && P.frame_mask) mask_byte = __ldg(P.frame_mask + row);
      // ---- drain the chains: round-to-nearest fp32 sum of (main + corr * 2^-11) partials
      float accr[BN];
      if (warp == 4 && lane == 0) JB_TRACE3(4, 3, seq);
      for (int ch = 0; ch < n_chains; ++ch) {
        mbar_wait(&tfull_bar[buf], buf_phase);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += 16) {   // 16 columns at a time keeps the live register set small
          uint32_t r[16], rc[16];
          tmem_ld16(lane_addr + static_cast<uint32_t>(buf * 2 * BN + c), r);
          tmem_ld16(lane_addr + static_cast<uint32_t>(buf * 2 * BN + BN + c), rc);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = fmaf(__uint_as_float(rc[i]), 1.0f / kSplitScale, __uint_as_float(r[i]));
            accr[c + i] = (ch == 0) ? x : accr[c + i] + x;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0));   // the leader's MMA thread waits
          else mbar_arrive(&tempty_bar[buf]);
        }
        if (++buf == 2) { buf = 0; buf_phase ^= 1; }
      }
      if (warp == 4 && lane == 0) JB_TRACE3(4, 0, seq);
      const bool valid = row < P.m_rows && mask_byte != 0u;
      const int ns = tile_slabs(P, n0);
      // ---- finish: one 32-column slab at a time through the epilogue ring
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s < ns) {
          mbar_wait(&epfull_bar[e], ph);
          uint8_t* ebuf = ep_base + e * P.entry_bytes;
          float v[32];
          const uint32_t bias_addr = smem_u32(bias_s);
          auto bias32 = [&](int col0, float (&b)[32]) {   // 32 consecutive bias values by shared-space 128-bit loads
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 t = lds128(bias_addr + static_cast<uint32_t>(col0 + 4 * q) * 4u);
              b[4 * q] = __uint_as_float(t.x); b[4 * q + 1] = __uint_as_float(t.y);
              b[4 * q + 2] = __uint_as_float(t.z); b[4 * q + 3] = __uint_as_float(t.w);
            }
          };
          if (P.act == ACT_GLU) {
            // tile columns [0,64) linear half, [64,128) gate half of the same 64 output channels
            float ba[32], bg[32];
            bias32(n0 + (s & 1) * 32, ba);
            bias32(n0 + 64 + (s & 1) * 32, bg);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int ca = (s & 1) * 32 + i;   // s < 2 in GLU mode
              const float a = accr[ca] + ba[i];
              const float g = accr[64 + ca] + bg[i];
              v[i] = a * (1.0f / (1.0f + __expf(-g))) * P.scale;
            }
          } else {
            float bb[32];
            bias32(n0 + s * 32, bb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x = accr[s * 32 + i] + bb[i];
              if (P.act == ACT_RELU) x = fmaxf(x, 0.f);
              else if (P.act == ACT_TANH) x = tanhf(x);
              v[i] = x * P.scale;
            }
          }
          if (P.has_res || P.has_f32) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {   // 8 chunks of 4 floats
              const uint32_t off = f_row + ((static_cast<uint32_t>(c) ^ f_sw) << 4);
              float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
              if (P.has_res) {
                const uint4 ru = lds128(smem_u32(ebuf) + f32_off + off);
                const float4 r4 = make_float4(__uint_as_float(ru.x), __uint_as_float(ru.y), __uint_as_float(ru.z), __uint_as_float(ru.w));
                o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                v[4 * c] = o.x; v[4 * c + 1] = o.y; v[4 * c + 2] = o.z; v[4 * c + 3] = o.w;
              }
              if (!valid) o = make_float4(0.f, 0.f, 0.f, 0.f);
              if (P.has_f32)
                sts128(smem_u32(ebuf) + f32_off + off, make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
            }
          }
          if (P.has_split) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {   // 4 chunks of 8 halves
              const uint32_t off = h_row + ((static_cast<uint32_t>(c) ^ h_sw) << 4);
              uint32_t ph4[4], pl4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                bf16 ah, al, bh, bl;
                split_op16(valid ? v[8 * c + 2 * j] : 0.f, ah, al);
                split_op16(valid ? v[8 * c + 2 * j + 1] : 0.f, bh, bl);
                __nv_bfloat162 hh = __halves2bfloat162(ah, bh), ll = __halves2bfloat162(al, bl);
                ph4[j] = *reinterpret_cast<uint32_t*>(&hh);
                pl4[j] = *reinterpret_cast<uint32_t*>(&ll);
              }
              sts128(smem_u32(ebuf) + hi_off + off, make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]));
              sts128(smem_u32(ebuf) + lo_off + off, make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]));
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&ready_bar[e]);
          if (++e == E) { e = 0; ph ^= 1; }
        }
      }
      if (warp == 4 && lane == 0) JB_TRACE3(4, 2, seq);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // no CTA exits while its peer may still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(TMEM_COLS))
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(TMEM_COLS))
                   : "memory");
  }
}

// 2-D fp32 row-major [rows, ld] matrix, box = [128 rows, 32 cols] (128-byte rows), 128B swizzle
int make_tmap_f32(CUtensorMap* map, const float* base, long long rows, int cols, int ld) {
  EncodeTiledFn fn = get_encode_fn();
  JB_REQUIRE(fn != nullptr, -3, "cuTensorMapEncodeTiled entry point not available");
  JB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 4) % 16 == 0, -2, "fp32 TMA alignment");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {SLAB, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JB_REQUIRE(r == CUDA_SUCCESS, -3, "cuTensorMapEncodeTiled(fp32) failed (code " + std::to_string(static_cast<int>(r)) + ")");
  return 0;
}

}  // namespace

bool conv_gemm_tc3_eligible(const ConvGemmProblem& p) {
  const ConvGemmEpilogue& e = p.ep;
  if (!p.a_lo || !p.w_lo || p.up_s > 0 || p.block_n != BN || p.w_tap_stride != 0) return false;
  if (!(e.act == ACT_NONE || e.act == ACT_RELU || e.act == ACT_TANH || e.act == ACT_GLU)) return false;
  if (e.res_bf16 || e.accum_in || e.accum_bf16 || e.out_act || e.res_inv_slope != 0.f) return false;
  if ((e.out_hi != nullptr) != (e.out_lo != nullptr)) return false;
  if (e.post_scale != 1.0f) return false;
  if (p.n_pad > BIAS_BYTES / 4 || p.out_rows != p.m_rows) return false;
  if (p.rate > 1) return false;
  if (p.tap_stride < 0 || (p.taps - 1) * p.tap_stride > A_SLAB_ROWS - BM) return false;
  if (e.act == ACT_GLU && (e.out_hi || e.res_f32)) return false;
  auto f32_ok = [&](const float* ptr, int ld) { return ptr == nullptr || (ld % 4 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  auto h_ok = [&](const bf16* ptr, int ld) { return ptr == nullptr || (ld % 8 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  if (!f32_ok(e.res_f32, e.res_ld) || !f32_ok(e.out_f32, e.out_f32_ld) || !h_ok(e.out_hi, e.out_bf_ld) || !h_ok(e.out_lo, e.out_bf_ld))
    return false;
  return e.out_f32 || e.out_hi;
}

int conv_gemm_tc3(const ConvGemmProblem& p, cudaStream_t stream) {
  const ConvGemmEpilogue& e = p.ep;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tres, tf32, thi, tlo;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  const int halo_rows = round_up(BM + (p.taps - 1) * p.tap_stride, 8);
  JB_PROPAGATE(make_tmap(&ta_hi, p.a_hi, p.a_rows, a_cols, p.a_ld, halo_rows));
  JB_PROPAGATE(make_tmap(&ta_lo, p.a_lo, p.a_rows, a_cols, p.a_ld, halo_rows));
  // CTA pairs (cta_group::2): each SM is sent, stores and reads half of every weight tile pair, and the weight ring
  // is twice as deep for the same bytes; needs >= 4 M tiles to be worth the lockstep
  static const int env_pair = getenv("JATTS_B200_TC3_PAIR") ? atoi(getenv("JATTS_B200_TC3_PAIR")) : 1;
  const bool pair = env_pair != 0 && ceil_div(p.m_rows, BM) >= 4;
  JB_PROPAGATE(make_tmap(&tb_hi, p.w_hi, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, pair ? BN / 2 : BN));
  JB_PROPAGATE(make_tmap(&tb_lo, p.w_lo, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, pair ? BN / 2 : BN));
  tres = tf32 = thi = tlo = ta_hi;
  if (e.res_f32) JB_PROPAGATE(make_tmap_f32(&tres, e.res_f32, p.m_rows, p.n, e.res_ld));
  if (e.out_f32) JB_PROPAGATE(make_tmap_f32(&tf32, e.out_f32, p.m_rows, p.n, e.out_f32_ld));
  if (e.out_hi) {
    JB_PROPAGATE(make_tmap(&thi, e.out_hi, p.m_rows, p.n, e.out_bf_ld, BM, SLAB));
    JB_PROPAGATE(make_tmap(&tlo, e.out_lo, p.m_rows, p.n, e.out_bf_ld, BM, SLAB));
  }
  Params3 kp;
  kp.taps = p.taps;
  kp.k_chunks = ceil_div(a_cols, BK);
  kp.n_pad = p.n_pad;
  kp.tap_off0 = p.tap_off0;
  kp.tap_stride = p.tap_stride;
  kp.n = p.n;
  kp.m_rows = p.m_rows;
  kp.num_m_tiles = ceil_div(p.m_rows, BM);
  kp.num_n_tiles = p.n_pad / BN;
  kp.num_groups = ceil_div(kp.num_m_tiles, 2) * kp.num_n_tiles;
  kp.frame_mask = p.frame_mask;
  kp.bias = e.bias;
  kp.act = e.act;
  kp.scale = e.scale;
  kp.has_res = e.res_f32 != nullptr;
  kp.has_f32 = e.out_f32 != nullptr;
  kp.has_split = e.out_hi != nullptr;
  kp.entry_bytes = ((kp.has_res || kp.has_f32) ? F32_SLAB_BYTES : 0) + (kp.has_split ? 2 * H16_SLAB_BYTES : 0);
  kp.halo_rows = halo_rows;
  // three weight stages when the epilogue ring still gets three entries, else two
  kp.b_stages = (227 * 1024 - smem_fixed(3)) / kp.entry_bytes >= 3 ? 3 : 2;
  const int SMEM_FIXED = smem_fixed(kp.b_stages);
  int entries = (227 * 1024 - SMEM_FIXED) / kp.entry_bytes;
  if (entries > MAX_ENTRIES) entries = MAX_ENTRIES;
  JB_REQUIRE(entries >= 2, -2, "conv_gemm_tc3: shared memory budget exceeded");
  kp.entries = entries;
  kp.trace = g_trace_ptr;
  const int smem_bytes = SMEM_FIXED + entries * kp.entry_bytes;
  JB_PROPAGATE(ensure_dynamic_smem(pair ? reinterpret_cast<const void*>(gemm_split_tma_kernel<true>)
                                        : reinterpret_cast<const void*>(gemm_split_tma_kernel<false>), smem_bytes));
  const int units = pair ? kp.num_groups : kp.num_m_tiles * kp.num_n_tiles;
  if (units == 0) return 0;
  const int max_units = pair ? num_sms() / 2 : num_sms();
  const int grid = (units < max_units ? units : max_units) * (pair ? 2 : 1);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventCreate(&e0));
    JB_CUDA_OK(cudaEventCreate(&e1));
    JB_CUDA_OK(cudaEventRecord(e0, stream));
  }
  if (!pair)
    JB_CUDA_OK(launch_tc(gemm_split_tma_kernel<false>, grid, kThreads3, smem_bytes, stream, 1, ta_hi, ta_lo, tb_hi, tb_lo, tres, tf32, thi, tlo, kp));
  else
    JB_CUDA_OK(launch_tc(gemm_split_tma_kernel<true>, grid, kThreads3, smem_bytes, stream, 2, ta_hi, ta_lo, tb_hi, tb_lo, tres, tf32, thi, tlo, kp));
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, 1});
  }
  return 0;
}

}  // namespace jb
