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
//   warps 4-7 epilogue: drain every chain (tcgen05.ld) and add the partial sums in REGISTERS with
//            round-to-nearest fp32 (main + corr * 2^-11); after the last chain, 64 columns at a time: bias,
//            ReLU / tanh / GLU, scale in the thread-per-row domain, then through a per-warp staging tile into
//            the row-contiguous domain (a ROLLED loop: 2 rows x 16 lanes x 16 B per instruction) where the fp32
//            residual is added (coalesced loads, prefetched four steps ahead), the row mask applied and the fp32
//            master and / or fp16 split pair stored with coalesced global stores.
//
// History of the epilogue (tools/gpu_trace_gemm.py): per-thread scattered global stores took 52 k clk per 128x128
// tile; TMA stores from a ring of swizzled slabs took 11-13 k clk (one store in flight at a time, ~2.9 k clk of
// store latency per 32-column slab) against 4.6-14 k clk of MMA work per tile, so every GEMM of the model was
// epilogue bound; the staged coalesced stores need no ring, no store latency and free 48 KB for the weight ring.
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
constexpr int MAX_B_STAGES = 8;                // 4 full stages, or 8 half stages when a CTA pair shares the weights
constexpr int OP_BYTES = BM * BK * 2;          // 16 KB: one weight tile (hi or lo)
constexpr int B_STAGE_BYTES = 2 * OP_BYTES;    // B_hi | B_lo
constexpr int A_SLAB_ROWS = 144;               // 128 + halo of the taps (<= 16 rows)
constexpr int A_OP_BYTES = A_SLAB_ROWS * BK * 2;   // 18 KB (multiple of 1024)
constexpr int A_STAGE_BYTES = 2 * A_OP_BYTES;  // A_hi | A_lo
constexpr int HALF = 64;                        // output columns per epilogue pass
constexpr int STG_PITCH = HALF + 4;             // staging row pitch in words: 64 data + 4 pad (conflict-free 128-bit rows)
constexpr int STG_BYTES = 32 * STG_PITCH * 4;   // one epilogue warp's [32 rows x 64 words] transpose tile
constexpr int STG1_PITCH = 32 + 4;              // single-chain kernels: 8 epilogue warps, 32-column passes
constexpr int STG1_BYTES = 32 * STG1_PITCH * 4;
constexpr int STG_TOTAL = 8 * STG1_BYTES > 4 * STG_BYTES ? 8 * STG1_BYTES : 4 * STG_BYTES;
constexpr int BIAS_BYTES = 8192;                // n_pad <= 2048
constexpr int BAR_BYTES = 1024;
constexpr int SMEM_MISC = BIAS_BYTES + BAR_BYTES + STG_TOTAL + 1024;
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
  int act;          // ACT_NONE / ACT_RELU / ACT_TANH / ACT_GLU / ACT_SNAKE
  const float* sn_a; const float* sn_ib;   // ACT_SNAKE: per-column exp(alpha), 1 / (exp(beta) + 1e-9) (n_pad floats each)
  float scale;
  const float* res; int res_ld;          // fp32 residual added to the output (may alias out_f32)
  float* out_f32; int out_f32_ld;        // optional fp32 output
  bf16* out_hi; bf16* out_lo; int out_h_ld;   // optional fp16 (hi, lo * 2^11) split pair
  int b_stages;     // weight ring depth (2..3)
  int num_groups;   // pair mode: ceil(num_m_tiles / 2) * num_n_tiles (a pair walks 2 M tiles x one N tile at a time)
  int halo_rows;    // rows of the activation slab: 128 + (taps-1)*tap_stride rounded up to 8
  int dbg;          // debug (JATTS_B200_TC3_DEBUG, results are WRONG): 1 = no global stores, 2 = no row-contiguous pass at all,
                    // 3 = also no staging stores; isolates what bounds the epilogue in the per-role timelines
  long long* trace;
};

#ifdef JB_ENABLE_TRACE
#define JB_TRACE3(role, ev, idx)                                                                         \
  do {                                                                                                   \
    if (P.trace && blockIdx.x == 0 && (idx) < 64) P.trace[((role) * 8 + (ev)) * 64 + (idx)] = clock64(); \
  } while (0)
#else
#define JB_TRACE3(role, ev, idx) do { } while (0)
#endif

__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {  // K-major, 128-byte swizzle
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Row-contiguous half of an epilogue pass: the warp's staging tile holds [32 rows x COLS] finished fp32 values
// (pitch words per row); COLS/4 lanes x 16 B cover one row, so a step stores 128 / COLS rows.  The fp32 residual is added
// here (coalesced loads, fetched one block of 4 steps ahead), masked rows become zeros, and the fp32 master and / or the
// fp16 (hi, lo * 2^11) pair are written with coalesced global stores.
template <int COLS>
__device__ __forceinline__ void store_pass(const Params3& P, uint32_t stg, int pitch, int row0, int ocol0, uint32_t inbits,
                                           uint32_t keepbits, int lane) {
  constexpr int LPR = COLS / 4;        // lanes per row
  constexpr int RPS = 32 / LPR;        // rows per step
  constexpr int STEPS = 32 / RPS;
  const int c4 = (lane % LPR) * 4;
  const int rsub = lane / LPR;
  const int col = ocol0 + c4;
  const bool col_ok = col < P.n;
  auto fetch_res = [&](float4 (&d)[4], int blk) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = (blk * 4 + i) * RPS + rsub;
      d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (P.res != nullptr && col_ok && ((inbits >> rl) & 1u))
        d[i] = *reinterpret_cast<const float4*>(P.res + static_cast<long long>(row0 + rl) * P.res_ld + col);
    }
  };
  if (P.dbg >= 2) return;
  float4 rcur[4];
  fetch_res(rcur, 0);
#pragma unroll 1
  for (int blk = 0; blk < STEPS / 4; ++blk) {
    float4 rnxt[4];
    if (blk + 1 < STEPS / 4) fetch_res(rnxt, blk + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = (blk * 4 + i) * RPS + rsub;
      const uint4 t = lds128(stg + static_cast<uint32_t>((rl * pitch + c4) * 4));
      float4 o = make_float4(__uint_as_float(t.x) + rcur[i].x, __uint_as_float(t.y) + rcur[i].y,
                             __uint_as_float(t.z) + rcur[i].z, __uint_as_float(t.w) + rcur[i].w);
      if (!((keepbits >> rl) & 1u)) o = make_float4(0.f, 0.f, 0.f, 0.f);   // masked rows are stored as zeros
      if (col_ok && ((inbits >> rl) & 1u) && P.dbg == 0) {
        const long long row = row0 + rl;
        if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + row * P.out_f32_ld + col) = o;
        if (P.out_hi) {
          uint32_t ha, la, hb, lb;
          split_pair16_sat(o.x, o.y, ha, la);
          split_pair16_sat(o.z, o.w, hb, lb);
          *reinterpret_cast<uint2*>(P.out_hi + row * P.out_h_ld + col) = make_uint2(ha, hb);
          *reinterpret_cast<uint2*>(P.out_lo + row * P.out_h_ld + col) = make_uint2(la, lb);
        }
      }
    }
    if (blk + 1 < STEPS / 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rcur[i] = rnxt[i];
    }
  }
}

// PAIR: the two CTAs of a cluster issue M = 256 cta_group::2 MMAs; each keeps its own 128 activation rows and HALF of
// every weight tile pair (a template parameter: kernels with cta_group::2 code need an even cluster size to launch)
// ACTK: 0 = none / ReLU (a floor of -inf / 0: one code path), 1 = tanh, 2 = GLU, 3 = SnakeBeta (single-chain kernels
// only) -- a template parameter because the
// unrolled epilogue with a run-time activation switch compiled to 10 k instructions per kernel (instruction-cache bound)
// SINGLE: every tile is one accumulation chain (K <= 8 blocks: the projections and pointwise convolutions, 4.6 k clk of
// MMAs per tile).  Four epilogue warps issue one dependent instruction stream per scheduler and need 11-12 k clk per
// tile (tools/gpu_trace_gemm.py), so these launches run TWO epilogue groups (12 warps), one per TMEM buffer, that
// read the accumulators straight into the staging tiles (no register copy of the tile, so 384 threads fit).
template <bool PAIR, int ACTK, bool SINGLE>
__global__ void __launch_bounds__(SINGLE ? 384 : kThreads3, 1)
gemm_split_tma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                      const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                      const __grid_constant__ Params3 P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_ring = smem + A_STAGES * A_STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(b_ring + P.b_stages * B_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + BIAS_BYTES);
  uint8_t* stg_base = reinterpret_cast<uint8_t*>(bars) + BAR_BYTES;   // [4 epilogue warps][STG_BYTES]
  uint64_t* full_bar = bars;                       // [MAX_B_STAGES]  weight ring
  uint64_t* empty_bar = full_bar + MAX_B_STAGES;   // [MAX_B_STAGES]
  uint64_t* afull_bar = empty_bar + MAX_B_STAGES;  // [A_STAGES]      activation slabs
  uint64_t* aempty_bar = afull_bar + A_STAGES;     // [A_STAGES]
  uint64_t* tfull_bar = aempty_bar + A_STAGES;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    for (int i = 0; i < MAX_B_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&afull_bar[i], 1); mbar_init(&aempty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 8 : 4); }
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
    for (int i = threadIdx.x; i < P.n_pad && i < BIAS_BYTES / 4; i += blockDim.x)
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
  } else if (warp >= 4 && SINGLE) {
    // ===================== epilogue, single-chain tiles: group e = (warp - 4) / 4 owns TMEM buffer e =====================
    const int lane_group = warp & 3;
    const int e = (warp - 4) >> 2;
    const uint32_t tb = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + static_cast<uint32_t>(e * 2 * BN);
    const uint32_t stg = smem_u32(stg_base + (warp - 4) * STG1_BYTES);
    const uint32_t stg_w = stg + static_cast<uint32_t>(lane * STG1_PITCH * 4);
    const float act_floor = P.act == ACT_RELU ? 0.f : -INFINITY;
    const uint32_t bias_addr = smem_u32(bias_s);
    constexpr int NPASS = ACTK == 2 ? 2 : 4;
    uint32_t ph = 0;
    for (int u = u0 + e * ustep, seq = e; u < num_units; u += 2 * ustep, seq += 2) {
      const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
      const int n0 = (u % P.num_n_tiles) * BN;
      const int o0 = ACTK == 2 ? n0 / 2 : n0;
      const int row0 = m0 + lane_group * 32;
      const bool in_me = row0 + lane < P.m_rows;
      const bool keep_me = in_me && (P.frame_mask == nullptr || __ldg(P.frame_mask + row0 + lane) != 0);
      const uint32_t inbits = __ballot_sync(0xffffffffu, in_me), keepbits = __ballot_sync(0xffffffffu, keep_me);
      if (warp == 4 && lane == 0) JB_TRACE3(4, 3, seq);
      mbar_wait(&tfull_bar[e], ph);
      ph ^= 1;
      tc_fence_after();
      if (warp == 4 && lane == 0) JB_TRACE3(4, 0, seq);
#pragma unroll 1
      for (int ps = 0; ps < NPASS; ++ps) {
        const int ocol0 = o0 + ps * 32;
        const bool last = ps == NPASS - 1 || ocol0 + 32 >= P.n;
        // ---- thread = row: 32 columns of the accumulator pair -> bias, activation, scale -> my staging row
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int tc = ps * 32 + hf * 16;   // tile column of this group (GLU: linear half; the gate is 64 further)
          uint32_t m[16], c[16], mg[16], cg[16];
          tmem_ld16(tb + static_cast<uint32_t>(tc), m);
          tmem_ld16(tb + static_cast<uint32_t>(BN + tc), c);
          if constexpr (ACTK == 2) {
            tmem_ld16(tb + static_cast<uint32_t>(64 + tc), mg);
            tmem_ld16(tb + static_cast<uint32_t>(BN + 64 + tc), cg);
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 bb = lds128(bias_addr + static_cast<uint32_t>(n0 + tc + 4 * q) * 4u);
            const float bv[4] = {__uint_as_float(bb.x), __uint_as_float(bb.y), __uint_as_float(bb.z), __uint_as_float(bb.w)};
            float v[4];
            if constexpr (ACTK == 2) {
              const uint4 bg = lds128(bias_addr + static_cast<uint32_t>(n0 + 64 + tc + 4 * q) * 4u);
              const float gv[4] = {__uint_as_float(bg.x), __uint_as_float(bg.y), __uint_as_float(bg.z), __uint_as_float(bg.w)};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float a = fmaf(__uint_as_float(c[4 * q + i]), 1.0f / kSplitScale, __uint_as_float(m[4 * q + i])) + bv[i];
                const float gt = fmaf(__uint_as_float(cg[4 * q + i]), 1.0f / kSplitScale, __uint_as_float(mg[4 * q + i])) + gv[i];
                v[i] = a * (1.0f / (1.0f + __expf(-gt))) * P.scale;
              }
            } else {
              float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
              if constexpr (ACTK == 3) {   // warp-uniform addresses: one broadcast transaction each, L1 resident
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(P.sn_a + n0 + tc + 4 * q));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.sn_ib + n0 + tc + 4 * q));
                sa[0] = a4.x; sa[1] = a4.y; sa[2] = a4.z; sa[3] = a4.w;
                sb[0] = b4.x; sb[1] = b4.y; sb[2] = b4.z; sb[3] = b4.w;
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x = fmaf(__uint_as_float(c[4 * q + i]), 1.0f / kSplitScale, __uint_as_float(m[4 * q + i])) + bv[i];
                if constexpr (ACTK == 1) x = tanhf(x);
                else if constexpr (ACTK == 3) {
                  // sin^2 has period pi: reduce the argument to [-pi/2, pi/2] with a two-constant Cody-Waite step (exact to
                  // ~1e-7 for |arg| up to a few hundred), then sin.approx, whose error is 2^-21 on that interval -- six
                  // instructions instead of the ~40 of sinf, which made this epilogue the bound of the GEMM; sin.approx on the
                  // raw argument lost |arg| * 2^-24 in the hardware's own reduction (8e-5 on outputs of magnitude 10)
                  const float arg = x * sa[i];
                  const float kq = rintf(arg * 0.318309886f);
                  const float red = fmaf(kq, 8.74227766e-8f, fmaf(kq, -3.14159274f, arg));
                  const float t = __sinf(red);
                  x = fmaf(sb[i], t * t, x);
                } else x = fmaxf(x, act_floor);
                v[i] = x * P.scale;
              }
            }
            if (P.dbg < 3) sts128(stg_w + static_cast<uint32_t>((hf * 16 + 4 * q) * 4), make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]),
                                                                                     __float_as_uint(v[2]), __float_as_uint(v[3])));
          }
        }
        if (last) {   // the accumulators of this tile have been read: hand the buffer back to the MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[e]), 0));
            else mbar_arrive(&tempty_bar[e]);
          }
        }
        __syncwarp();
        store_pass<32>(P, stg, STG1_PITCH, row0, ocol0, inbits, keepbits, lane);
        __syncwarp();
        if (last) break;
      }
      if (warp == 4 && lane == 0) JB_TRACE3(4, 2, seq);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..7) =====================
    const int lane_group = warp & 3;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t stg = smem_u32(stg_base + (warp - 4) * STG_BYTES);
    const uint32_t stg_w = stg + static_cast<uint32_t>(lane * STG_PITCH * 4);   // my row of the staging tile
    int buf = 0;
    uint32_t buf_phase = 0;
    for (int u = u0, seq = 0; u < num_units; u += ustep, ++seq) {
      const int m0 = ((u / P.num_n_tiles) * CL + rank) * BM;
      const int n0 = (u % P.num_n_tiles) * BN;
      if (warp == 4 && lane == 0) JB_TRACE3(4, 3, seq);
      // ---- drain the chains: round-to-nearest fp32 sum of (main + corr * 2^-11) partials
      float accr[BN];
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
      // ---- finish: 64 output columns per pass (GLU: the tile's 64 outputs in one pass)
      const int o0 = ACTK == 2 ? n0 / 2 : n0;
      const int row0 = m0 + lane_group * 32;
      // validity of the warp's 32 rows as bit masks: inside the matrix / unmasked (lane = row here)
      const bool in_me = row0 + lane < P.m_rows;
      const bool keep_me = in_me && (P.frame_mask == nullptr || __ldg(P.frame_mask + row0 + lane) != 0);
      const uint32_t inbits = __ballot_sync(0xffffffffu, in_me), keepbits = __ballot_sync(0xffffffffu, keep_me);
      const float act_floor = P.act == ACT_RELU ? 0.f : -INFINITY;
      const uint32_t bias_addr = smem_u32(bias_s);
#pragma unroll
      for (int h = 0; h < (ACTK == 2 ? 1 : 2); ++h) {
        const int ocol0 = o0 + h * HALF;
        if (ocol0 >= P.n) break;
        // ---- thread = row: bias, activation, scale; 16 x 128-bit stores into my staging row
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float v[4];
          if constexpr (ACTK == 2) {
            // tile columns [0,64) linear half, [64,128) gate half of the same 64 output channels
            const uint4 ba = lds128(bias_addr + static_cast<uint32_t>(n0 + 4 * q) * 4u);
            const uint4 bg = lds128(bias_addr + static_cast<uint32_t>(n0 + 64 + 4 * q) * 4u);
            const float bav[4] = {__uint_as_float(ba.x), __uint_as_float(ba.y), __uint_as_float(ba.z), __uint_as_float(ba.w)};
            const float bgv[4] = {__uint_as_float(bg.x), __uint_as_float(bg.y), __uint_as_float(bg.z), __uint_as_float(bg.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a = accr[4 * q + i] + bav[i];
              const float gt = accr[64 + 4 * q + i] + bgv[i];
              v[i] = a * (1.0f / (1.0f + __expf(-gt))) * P.scale;
            }
          } else {
            const uint4 bb = lds128(bias_addr + static_cast<uint32_t>(n0 + h * HALF + 4 * q) * 4u);
            const float bv[4] = {__uint_as_float(bb.x), __uint_as_float(bb.y), __uint_as_float(bb.z), __uint_as_float(bb.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x = accr[h * HALF + 4 * q + i] + bv[i];
              if constexpr (ACTK == 1) x = tanhf(x);
              else x = fmaxf(x, act_floor);
              v[i] = x * P.scale;
            }
          }
          sts128(stg_w + static_cast<uint32_t>(q * 16), make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]),
                                                                    __float_as_uint(v[2]), __float_as_uint(v[3])));
        }
        __syncwarp();
        store_pass<HALF>(P, stg, STG_PITCH, row0, ocol0, inbits, keepbits, lane);
        __syncwarp();
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

}  // namespace

bool conv_gemm_tc3_eligible(const ConvGemmProblem& p) {
  const ConvGemmEpilogue& e = p.ep;
  if (!p.a_lo || !p.w_lo || p.up_s > 0 || p.block_n != BN || p.w_tap_stride != 0) return false;
  if (!(e.act == ACT_NONE || e.act == ACT_RELU || e.act == ACT_TANH || e.act == ACT_GLU || e.act == ACT_SNAKE)) return false;
  if (e.act == ACT_SNAKE && (!e.snake_a || !e.snake_ib || p.taps * ceil_div(p.a_cols > 0 ? p.a_cols : p.k_pad, BK) > CHUNK ||
                             p.n != p.n_pad))
    return false;   // SnakeBeta lives in the single-chain epilogue only
  if (e.res_bf16 || e.accum_in || e.accum_bf16 || e.out_act || e.res_inv_slope != 0.f) return false;
  if ((e.out_hi != nullptr) != (e.out_lo != nullptr)) return false;
  if (e.post_scale != 1.0f) return false;
  if (p.n_pad > BIAS_BYTES / 4 || p.out_rows != p.m_rows) return false;
  if (p.rate > 1) return false;
  if (p.tap_stride < 0 || (p.taps - 1) * p.tap_stride > A_SLAB_ROWS - BM) return false;
  if (e.act == ACT_GLU && (e.out_hi || e.res_f32)) return false;
  if (p.n % 4 != 0) return false;   // the epilogue stores 4 columns per lane
  auto f32_ok = [&](const float* ptr, int ld) { return ptr == nullptr || (ld % 4 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  auto h_ok = [&](const bf16* ptr, int ld) { return ptr == nullptr || (ld % 8 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  if (!f32_ok(e.res_f32, e.res_ld) || !f32_ok(e.out_f32, e.out_f32_ld) || !h_ok(e.out_hi, e.out_bf_ld) || !h_ok(e.out_lo, e.out_bf_ld))
    return false;
  return e.out_f32 || e.out_hi;
}

int conv_gemm_tc3(const ConvGemmProblem& p, cudaStream_t stream) {
  const ConvGemmEpilogue& e = p.ep;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
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
  kp.sn_a = e.snake_a; kp.sn_ib = e.snake_ib;
  kp.scale = e.scale;
  kp.res = e.res_f32; kp.res_ld = e.res_ld;
  kp.out_f32 = e.out_f32; kp.out_f32_ld = e.out_f32_ld;
  kp.out_hi = e.out_hi; kp.out_lo = e.out_lo; kp.out_h_ld = e.out_bf_ld;
  kp.halo_rows = halo_rows;
  // the weight ring takes what the activation slabs, bias table and staging tiles leave: 3 stages (6 half stages per
  // CTA in pair mode)
  kp.b_stages = (227 * 1024 - smem_fixed(0)) / B_STAGE_BYTES;
  if (kp.b_stages > MAX_B_STAGES / 2) kp.b_stages = MAX_B_STAGES / 2;
  JB_REQUIRE(kp.b_stages >= 2, -2, "conv_gemm_tc3: shared memory budget exceeded");
  kp.trace = g_trace_ptr;
  static const int env_dbg = getenv("JATTS_B200_TC3_DEBUG") ? atoi(getenv("JATTS_B200_TC3_DEBUG")) : 0;
  kp.dbg = env_dbg;
  const int smem_bytes = smem_fixed(kp.b_stages);
  const int actk = e.act == ACT_GLU ? 2 : (e.act == ACT_TANH ? 1 : (e.act == ACT_SNAKE ? 3 : 0));
  // single-chain launches (K <= 8 blocks) run the 12-warp kernel with two epilogue groups
  const int single = kp.taps * kp.k_chunks <= CHUNK ? 1 : 0;
  using KernFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Params3);
  static const KernFn kerns[2][2][4] = {
      {{gemm_split_tma_kernel<false, 0, false>, gemm_split_tma_kernel<false, 1, false>, gemm_split_tma_kernel<false, 2, false>, nullptr},
       {gemm_split_tma_kernel<true, 0, false>, gemm_split_tma_kernel<true, 1, false>, gemm_split_tma_kernel<true, 2, false>, nullptr}},
      {{gemm_split_tma_kernel<false, 0, true>, gemm_split_tma_kernel<false, 1, true>, gemm_split_tma_kernel<false, 2, true>,
        gemm_split_tma_kernel<false, 3, true>},
       {gemm_split_tma_kernel<true, 0, true>, gemm_split_tma_kernel<true, 1, true>, gemm_split_tma_kernel<true, 2, true>,
        gemm_split_tma_kernel<true, 3, true>}}};
  const KernFn kern = kerns[single][pair ? 1 : 0][actk];
  JB_REQUIRE(kern != nullptr, -1, "conv_gemm_tc3: SnakeBeta needs a single accumulation chain (K <= 512)");
  JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_bytes));
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
  JB_CUDA_OK(launch_tc(kern, grid, single ? 384 : kThreads3, smem_bytes, stream, pair ? 2 : 1, ta_hi, ta_lo, tb_hi, tb_lo, kp));
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, 1});
  }
  return 0;
}

}  // namespace jb
