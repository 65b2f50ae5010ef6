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
static constexpr int kThreads2 = 384;  // producer, MMA issuer, loader, store, 2 x 4 epilogue warps
enum : int { MODE_STREAM = 0, MODE_HALO = 1, MODE_RESIDENT = 2 };

struct KernelParams2 {
  int taps, k_chunks, n_pad;
  int tap_off0, tap_stride;
  int m_rows;
  int num_m_tiles, num_n_tiles;
  int cl;           // CTAs per cluster sharing the streamed weight tiles by TMA multicast (HALO mode), else 1
  int pair;         // 1: the cluster is a CTA pair issuing M = 256 cta_group::2 MMAs (cl == 2, HALO): each CTA keeps its own
                    // 128 activation rows and HALF of every weight tile in shared memory
  int num_groups;   // ceil(num_m_tiles / cl) * num_n_tiles: a cluster walks groups of cl M tiles x one N tile
  const uint8_t* frame_mask;
  int rate;
  const float* bias;
  int act;          // ACT_NONE or ACT_LRELU
  float slope;
  float post_scale;
  float out1_slope;
  float res_inv_slope;   // != 0: the residual tensor stores LeakyReLU(x); recover x before adding
  int has_res, has_acc, has_out0, has_out1;
  int halo_rows;    // HALO/RESIDENT: rows of the activation box (128 + (taps-1)*tap_stride, padded to 8)
  int ep_entries;   // epilogue ring depth (2..8)
  int ep_bufs;      // slabs per entry: 1 (one input-or-output tensor) or 2
  int w_bytes;      // RESIDENT: bytes of the weight area
  int store_depth;  // TMA store groups kept in flight by the store warp (0..2)
  int w_row0, w_tap_stride;       // weight rows of tap t: w_row0 + t*w_tap_stride (+ n0)
  int store_row_off;              // output row coordinate = m0 + store_row_off
  int mask_mul, mask_add, out_rows;  // validity: frame_mask[(row*mask_mul + mask_add) / rate], 0 <= . < out_rows
  int n_slabs;                       // epilogue slabs per tile that hold real output columns (< N_SLABS for a partly filled N tile)
  int n_phases, phase_pp;            // > 0: N tile q is phase q of a polyphase transposed convolution (ConvGemmProblem::phases):
                                     // tap_off0, mask_add and the output tensor map depend on q
  long long* trace; // debug: per-role clock64 timeline of CTA 0 (jatts_debug_set_trace), else null
};

long long* g_trace_ptr = nullptr;
#ifdef JB_ENABLE_TRACE
#define JB_TRACE(role, ev, idx)                                                                \
  do {                                                                                         \
    if (P.trace && blockIdx.x == 0 && (idx) < 64) P.trace[((role) * 8 + (ev)) * 64 + (idx)] = clock64(); \
  } while (0)
#else
#define JB_TRACE(role, ev, idx) do { } while (0)   // stamps compiled out: the kernels are instruction-cache sensitive
#endif

template <int BLOCK_N, int KCH, int MODE>
struct Cfg2 {
  // epilogue slabs: [128 rows x 32 columns] bf16 = 64-byte rows (SWIZZLE_64B), 8 KB.  An epilogue ring
  // entry holds 1 or 2 slabs ([residual -> out0 | branch sum -> out1]); the ring depth is whatever
  // shared memory is left (launch2), because the per-tile critical path is latency x concurrency:
  // residual prefetch distance and the number of TMA stores in flight both scale with it.
  static constexpr int SLAB = 32;
  static constexpr int N_SLABS = BLOCK_N / SLAB;
  static constexpr int ROWB = SLAB * 2;
  static constexpr int SW_MASK = 3;                         // Swizzle<2,4,3>
  static constexpr int SLAB_BYTES = BLOCK_M2 * ROWB;        // 8192
  static constexpr int MAX_ENTRIES = 8;
  static constexpr int KROWB = KCH * 2;                     // operand row bytes (128 or 64) = swizzle span
  static constexpr int A_BYTES = BLOCK_M2 * KROWB;
  static constexpr int B_BYTES = BLOCK_N * KROWB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2048;                   // n_pad <= 512
  static constexpr int BAR_BYTES = 1024;                    // keeps the resident weight area 1024-aligned
  static constexpr int FIXED_TAIL = BIAS_BYTES + BAR_BYTES + 1024;
  static constexpr int MIN_EP_BYTES = 4 * 2 * SLAB_BYTES;   // the mainloop leaves at least 4 two-slab entries
  static constexpr int BUDGET = 227 * 1024 - FIXED_TAIL - MIN_EP_BYTES;
  // STREAM
  static constexpr int MAX_STAGES = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES >= 6 ? 6 : (MAX_STAGES >= 4 ? 4 : 2);   // even: half per pipeline
  // HALO / RESIDENT: activation slabs [HALO_MAX_ROWS x KCH]
  static constexpr int HALO_MAX_ROWS = 192;                 // 128 + (11-1)*5 = 178, padded
  static constexpr int A_SLAB_BYTES = HALO_MAX_ROWS * KROWB;  // multiple of 1024
  static constexpr int A_STAGES = (MODE == MODE_RESIDENT && BLOCK_N == 32) ? 4 : 2;   // even: half per pipeline when two MMA warps issue
  static constexpr int B_MAX = (BUDGET - A_STAGES * A_SLAB_BYTES) / B_BYTES;
  static constexpr int B_STAGES = B_MAX >= 4 ? 4 : 2;
  static constexpr int MAIN_BYTES = MODE == MODE_STREAM ? STAGES * STAGE_BYTES
                                  : MODE == MODE_HALO   ? A_STAGES * A_SLAB_BYTES + B_STAGES * B_BYTES
                                                        : A_STAGES * A_SLAB_BYTES;
  static constexpr int SMEM_FIXED = MAIN_BYTES + FIXED_TAIL;   // + epilogue ring + resident weights (launch2)
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert(MODE != MODE_STREAM || MAX_STAGES >= 2, "stream pipeline depth");
  static_assert(MODE != MODE_HALO || B_MAX >= 2, "weight ring depth");
  static_assert(A_STAGES <= 4 && STAGES <= 8 && B_STAGES <= 8, "barrier arrays");
};

// Ring cursor: slot index + mbarrier phase parity, advanced incrementally (no integer divisions on the
// per-tile critical path).  Every mbarrier has exactly one producer and one consumer walking it in
// lockstep -- an mbarrier parity wait cannot tell "two phases ahead" from "done", so the two epilogue
// groups (even / odd tiles) own disjoint halves of the epilogue ring: slots [base, base + depth).
struct Ring {
  int idx, base, depth;
  uint32_t phase;
  __device__ __forceinline__ Ring(int base_, int depth_) : idx(base_), base(base_), depth(depth_), phase(0) {}
  __device__ __forceinline__ void next() {
    if (++idx == base + depth) { idx = base; phase ^= 1u; }
  }
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

// PAIR is a template parameter, not a runtime flag: a kernel that contains cta_group::2 instructions can only be
// launched with an even cluster size ("cluster misconfiguration" otherwise, even if the path is never taken)
template <int BLOCK_N, int KCH, int MODE, bool PAIR>
__global__ void __launch_bounds__(kThreads2, 1)
conv_bf16_tma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_bm,
                     const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_acc,
                     const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1,
                     const __grid_constant__ CUtensorMap tm_ph4, const __grid_constant__ KernelParams2 P) {
  using C = Cfg2<BLOCK_N, KCH, MODE>;
  constexpr int STAGES = C::STAGES;
  constexpr int KSTEPS = KCH / UMMA_K2;
  // weight ring: a CTA of a pair keeps half of each tile, so the same bytes give twice the depth (Little's law:
  // bytes in flight = consumption rate x L2 latency; measured 55-60 % of the MMA rate with 4 x 16 KB per CTA)
  constexpr int B_STAGE_BYTES = PAIR ? C::B_BYTES / 2 : C::B_BYTES;
  constexpr int B_DEPTH = PAIR ? 2 * C::B_STAGES : C::B_STAGES;
  static_assert(B_DEPTH <= 8, "barrier arrays");
  const int E = P.ep_entries;
  const int entry_bytes = P.ep_bufs * C::SLAB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* b_base = smem + C::A_STAGES * C::A_SLAB_BYTES;             // HALO: weight ring after the A slabs
  float* bias_s = reinterpret_cast<float*>(smem + C::MAIN_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + C::BIAS_BYTES);
  uint8_t* w_base = reinterpret_cast<uint8_t*>(bars) + C::BAR_BYTES;  // RESIDENT: all weight tiles (1024-aligned)
  uint8_t* ep_base = w_base + P.w_bytes;                              // epilogue ring (1024-aligned)
  constexpr int NB = 8;
  uint64_t* full_bar = bars;                    // [NB]   STREAM: A+B stage | HALO: weight ring
  uint64_t* empty_bar = full_bar + NB;          // [NB]
  uint64_t* afull_bar = empty_bar + NB;         // [4]    HALO/RESIDENT: activation slabs
  uint64_t* aempty_bar = afull_bar + 4;         // [4]
  uint64_t* tfull_bar = aempty_bar + 4;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  constexpr int ME = C::MAX_ENTRIES;
  uint64_t* epfull_bar = tempty_bar + 2;        // [ME]
  uint64_t* epempty_bar = epfull_bar + ME;      // [ME]
  uint64_t* ready_bar = epempty_bar + ME;       // [ME]   epilogue -> store warp: slab written
  uint64_t* wfull_bar = ready_bar + ME;         // [1]    RESIDENT: weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // tile schedule: cluster `cid` takes groups cid, cid + ncl, ...; CTA `rank` of the cluster owns M tile
  // (grp / num_n_tiles) * cl + rank of the group (past the end: loads zero-fill, stores clip)
  const int cl = P.cl;
  const int rank = static_cast<int>(blockIdx.x) % cl;
  const int cid = static_cast<int>(blockIdx.x) / cl;
  const int ncl = static_cast<int>(gridDim.x) / cl;
  const uint16_t cl_mask = static_cast<uint16_t>((1u << cl) - 1u);
  // phase mode: per-phase tap offset / first output row / output view (the five tensor-map slots hold phases 0..4)
  const int n_ph = P.n_phases;
  auto tap_off_of = [&](int q) { return n_ph ? (q >= P.phase_pp ? -1 : 0) : P.tap_off0; };
  auto mask_add_of = [&](int q) { return n_ph ? (q >= P.phase_pp ? q - P.phase_pp : q - P.phase_pp + n_ph) : P.mask_add; };
  const CUtensorMap* const ph_maps[5] = {&tm_out1, &tm_out0, &tm_res, &tm_acc, &tm_ph4};

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (P.has_res) tma_prefetch_desc(&tm_res);
    if (P.has_acc) tma_prefetch_desc(&tm_acc);
    if (P.has_out0) tma_prefetch_desc(&tm_out0);
    if (P.has_out1) tma_prefetch_desc(&tm_out1);
    for (int q = 1; q < n_ph; ++q) tma_prefetch_desc(ph_maps[q]);
    for (int i = 0; i < NB; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], (MODE == MODE_HALO && !PAIR) ? cl : 1);   // multicast weight stage: freed by every CTA of the cluster
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 8 : 4);   // pair: the leader's barrier also collects the peer's epilogue warps
    }
    for (int i = 0; i < ME; ++i) {
      mbar_init(&epfull_bar[i], 1);
      mbar_init(&epempty_bar[i], 1);
      mbar_init(&ready_bar[i], 128);
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(static_cast<uint32_t>(C::TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(static_cast<uint32_t>(C::TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  for (int i = threadIdx.x; i < P.n_pad && i < C::BIAS_BYTES / 4; i += kThreads2) bias_s[i] = P.bias ? P.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // peers' mbarriers are initialised before any multicast load / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel's prologue may start on SMs this grid has left
  pdl_wait();                // everything above ran under the previous kernel's tail; its outputs are visible from here

  if (warp == 0) {
    // ===================== mainloop TMA producer =====================
    if (elect_one()) {
      if (MODE == MODE_RESIDENT) {
        // every weight tile of this convolution, once per CTA (num_n_tiles == 1 in this mode)
        mbar_expect_tx(wfull_bar, static_cast<uint32_t>(P.taps * P.k_chunks) * C::B_BYTES);
        for (int kc = 0; kc < P.k_chunks; ++kc)
          for (int tap = 0; tap < P.taps; ++tap)
            tma_load_2d(&tm_b, wfull_bar, w_base + (kc * P.taps + tap) * C::B_BYTES, kc * KCH, P.w_row0 + tap * P.w_tap_stride);
      }
      const uint32_t a_bytes = static_cast<uint32_t>(P.halo_rows) * C::KROWB;
      Ring rs(0, STAGES), ra(0, C::A_STAGES), rb(0, B_DEPTH);
      for (int grp = cid, seq = 0; grp < P.num_groups; grp += ncl, ++seq) {
        const int m0 = ((grp / P.num_n_tiles) * cl + rank) * BLOCK_M2;
        const int n0 = (grp % P.num_n_tiles) * BLOCK_N;
        if (MODE == MODE_STREAM) {
          for (int tap = 0; tap < P.taps; ++tap) {
            const int arow = m0 + tap_off_of(grp % P.num_n_tiles) + tap * P.tap_stride;
            const int brow = P.w_row0 + tap * P.w_tap_stride + n0;
            for (int kc = 0; kc < P.k_chunks; ++kc) {
              mbar_wait(&empty_bar[rs.idx], rs.phase ^ 1);
              uint8_t* st = smem + rs.idx * C::STAGE_BYTES;
              mbar_expect_tx(&full_bar[rs.idx], C::STAGE_BYTES);
              tma_load_2d(&tm_a, &full_bar[rs.idx], st, kc * KCH, arow);
              tma_load_2d(&tm_b, &full_bar[rs.idx], st + C::A_BYTES, kc * KCH, brow);
              rs.next();
            }
          }
        } else {
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            mbar_wait(&aempty_bar[ra.idx], ra.phase ^ 1);
            if (kc == 0) JB_TRACE(0, 0, seq);
            if constexpr (!PAIR) {
              mbar_expect_tx(&afull_bar[ra.idx], a_bytes);
              tma_load_2d(&tm_a, &afull_bar[ra.idx], smem + ra.idx * C::A_SLAB_BYTES, kc * KCH, m0 + tap_off_of(grp % P.num_n_tiles));
            } else {
              // both CTAs' slabs complete on the LEADER's barrier (only the leader's MMA thread waits on it)
              if (rank == 0) mbar_expect_tx(&afull_bar[ra.idx], 2 * a_bytes);
              tma_load_2d_pair(&tm_a, mapa_u32(smem_u32(&afull_bar[ra.idx]), 0), smem + ra.idx * C::A_SLAB_BYTES, kc * KCH,
                               m0 + tap_off_of(grp % P.num_n_tiles));
            }
            ra.next();
            if (MODE == MODE_HALO) {
              for (int tap = 0; tap < P.taps; ++tap) {
                mbar_wait(&empty_bar[rb.idx], rb.phase ^ 1);
                if constexpr (PAIR) {
                  // this CTA's half of the weight tile (rows rank * N/2 ...) stays in its own shared memory
                  if (rank == 0) mbar_expect_tx(&full_bar[rb.idx], C::B_BYTES);
                  tma_load_2d_pair(&tm_bm, mapa_u32(smem_u32(&full_bar[rb.idx]), 0), b_base + rb.idx * B_STAGE_BYTES, kc * KCH,
                                   P.w_row0 + tap * P.w_tap_stride + n0 + rank * (BLOCK_N / 2));
                  rb.next();
                  continue;
                }
                mbar_expect_tx(&full_bar[rb.idx], C::B_BYTES);
                if (cl == 1) {
                  tma_load_2d(&tm_b, &full_bar[rb.idx], b_base + rb.idx * B_STAGE_BYTES, kc * KCH, P.w_row0 + tap * P.w_tap_stride + n0);
                } else {
                  // this CTA fetches rows [rank, rank + 1) * BLOCK_N / cl of the weight tile for the whole cluster
                  const int slice = BLOCK_N / cl;
                  tma_load_2d_mc(&tm_bm, &full_bar[rb.idx], b_base + rb.idx * B_STAGE_BYTES + rank * slice * C::KROWB, kc * KCH,
                                 P.w_row0 + tap * P.w_tap_stride + n0 + rank * slice, cl_mask);
                }
                rb.next();
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Measured (tools/gpu_trace.py): for N <= 64 this single issuing thread is the bottleneck -- ~46 clk to
    // issue one tcgen05.mma that executes in 16-32 clk, plus several hundred clk of mbarrier handling per
    // tile.  A second issuing warp made it worse (84 clk per MMA: the issue port is shared).  Tiles alternate
    // between the two TMEM accumulators / epilogue groups.
    if ((!PAIR || rank == 0) && elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BLOCK_M2, BLOCK_N, /*is_bf16=*/true);
      constexpr uint32_t idesc2 = make_idesc(2 * BLOCK_M2, BLOCK_N, /*is_bf16=*/true);
      constexpr uint32_t idesc = PAIR ? idesc2 : idesc1;
      const int pipe = 0;
      Ring rs(0, STAGES), ra(0, C::A_STAGES), rb(0, B_DEPTH);
      const uint32_t w_addr = smem_u32(w_base);
      // descriptor words: lo = (addr >> 4) | LBO(1) << 16 ; hi = SBO (8 rows) | version 1 | swizzle mode
      constexpr uint32_t desc_lo0 = 1u << 16;
      constexpr uint32_t desc_hi = static_cast<uint32_t>((8 * KCH * 2) >> 4) | (1u << 14) | (static_cast<uint32_t>(KCH == 64 ? 2 : 4) << 29);
      bool weights_ready = MODE != MODE_RESIDENT;
      for (int grp = cid, seq = 0; grp < P.num_groups; grp += ncl, ++seq) {
        const int acc = seq & 1;   // accumulator (and epilogue group) of this tile
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        mbar_wait(&tempty_bar[acc], ((seq >> 1) & 1) ^ 1);
        if (pipe == 0) JB_TRACE(1, 0, seq);
        tc_fence_after();
        if (!weights_ready) {
          mbar_wait(wfull_bar, 0);
          tc_fence_after();
          weights_ready = true;
        }
        if (MODE == MODE_STREAM) {
          const int k_iters = P.taps * P.k_chunks;
          for (int it = 0; it < k_iters; ++it) {
            const int stage = rs.idx;
            mbar_wait(&full_bar[stage], rs.phase);
            rs.next();
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
          }
        } else {
          for (int kc = 0; kc < P.k_chunks; ++kc) {
            const int as = ra.idx;
            mbar_wait(&afull_bar[as], ra.phase);
            ra.next();
            if (pipe == 0 && kc == 0) JB_TRACE(1, 1, seq);
            tc_fence_after();
            const uint32_t slab = smem_u32(smem + as * C::A_SLAB_BYTES);
            for (int tap = 0; tap < P.taps; ++tap) {
              uint32_t b_addr;
              int bs = 0;
              if (MODE == MODE_HALO) {
                bs = rb.idx;
                mbar_wait(&full_bar[bs], rb.phase);
                rb.next();
                tc_fence_after();
                b_addr = smem_u32(b_base + bs * B_STAGE_BYTES);
              } else {
                b_addr = w_addr + static_cast<uint32_t>(kc * P.taps + tap) * C::B_BYTES;
              }
              // tap t reads activation rows [t*dilation, t*dilation + 128) of the slab: a row-shifted
              // start address.  Measured on B200: the swizzle is a function of the absolute smem address
              // bits, so a start that is not 1024 B aligned needs NO base_offset in the descriptor.
              const uint32_t a_lo = desc_lo0 + ((slab + static_cast<uint32_t>(tap * P.tap_stride) * C::KROWB) >> 4);
              const uint32_t b_lo = desc_lo0 + (b_addr >> 4);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if constexpr (MODE == MODE_HALO && PAIR)
                  tc_mma_bf16_lohi_pair(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (kc | tap | k) != 0 ? 1u : 0u);
                else
                  tc_mma_bf16_lohi(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
              if (MODE == MODE_HALO) {
                if constexpr (PAIR) tc_commit_pair(&empty_bar[bs]);
                else if (cl == 1) tc_commit(&empty_bar[bs]);
                else tc_commit_mc(&empty_bar[bs], cl_mask);
              }
            }
            if constexpr (MODE == MODE_HALO && PAIR) tc_commit_pair(&aempty_bar[as]);
            else tc_commit(&aempty_bar[as]);
          }
        }
        if constexpr (MODE == MODE_HALO && PAIR) tc_commit_pair(&tfull_bar[acc]);
        else tc_commit(&tfull_bar[acc]);
        if (pipe == 0) JB_TRACE(1, 2, seq);
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue loader =====================
    if (elect_one()) {
      const uint32_t bytes = (P.has_res ? C::SLAB_BYTES : 0) + (P.has_acc ? C::SLAB_BYTES : 0);
      Ring rg0(0, E / 2), rg1(E / 2, E / 2);   // the two epilogue groups' halves of the ring
      for (int grp = cid, seq = 0; grp < P.num_groups; grp += ncl, ++seq) {
        Ring& re = (seq & 1) ? rg1 : rg0;
        const int m0 = ((grp / P.num_n_tiles) * cl + rank) * BLOCK_M2;
        const int n0 = (grp % P.num_n_tiles) * BLOCK_N;
        for (int s = 0; s < P.n_slabs; ++s) {
          const int e = re.idx;
          mbar_wait(&epempty_bar[e], re.phase ^ 1);
          re.next();
          if (s == 0) JB_TRACE(2, 0, seq);
          uint8_t* buf = ep_base + e * entry_bytes;
          if (bytes) {
            mbar_expect_tx(&epfull_bar[e], bytes);
            if (P.has_res) tma_load_2d(&tm_res, &epfull_bar[e], buf, n0 + s * C::SLAB, m0);
            if (P.has_acc) tma_load_2d(&tm_acc, &epfull_bar[e], buf + (P.ep_bufs - 1) * C::SLAB_BYTES, n0 + s * C::SLAB, m0);
          } else {
            mbar_arrive(&epfull_bar[e]);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== store warp: slab -> global by TMA =====================
    if (elect_one()) {
      int groups = 0;
      int hist0 = 0, hist1 = 0;   // entries of the two most recently committed store groups
      const int depth = P.store_depth;   // older groups kept in flight behind the one just committed (0..2)
      Ring rg0(0, E / 2), rg1(E / 2, E / 2);
      for (int grp = cid, seq = 0; grp < P.num_groups; grp += ncl, ++seq) {
        Ring& re = (seq & 1) ? rg1 : rg0;
        const int m0 = ((grp / P.num_n_tiles) * cl + rank) * BLOCK_M2;
        const int n0 = (grp % P.num_n_tiles) * BLOCK_N;
        for (int s = 0; s < P.n_slabs; ++s) {
          const int e = re.idx;
          mbar_wait(&ready_bar[e], re.phase);  // the 128 threads of the tile's epilogue group wrote + fenced this slab
          re.next();
          if (s == 0) JB_TRACE(3, 0, seq);
          uint8_t* bufA = ep_base + e * entry_bytes;
          if (n_ph) {   // phase q's strided output view; its columns are the phase's own C_out channels
            tma_store_2d(ph_maps[grp % P.num_n_tiles], bufA + (P.ep_bufs - 1) * C::SLAB_BYTES, s * C::SLAB, m0);
          } else {
            if (P.has_out0) tma_store_2d(&tm_out0, bufA, n0 + s * C::SLAB, m0 + P.store_row_off);
            if (P.has_out1) tma_store_2d(&tm_out1, bufA + (P.ep_bufs - 1) * C::SLAB_BYTES, n0 + s * C::SLAB, m0 + P.store_row_off);
          }
          tma_store_commit();
          ++groups;
          // `depth` older store groups stay in flight; the entry of the group that has certainly finished
          // reading its slab goes back to the loader
          int release;
          if (depth == 0) { tma_store_wait_read<0>(); release = e; }
          else if (depth == 1) { tma_store_wait_read<1>(); release = hist0; }
          else { tma_store_wait_read<2>(); release = hist1; }
          if (groups > depth) mbar_arrive(&epempty_bar[release]);
          hist1 = hist0;
          hist0 = e;
          if (s == 0) JB_TRACE(3, 1, seq);
        }
      }
      // (entries still held at the end are never needed again)
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue (warps 4..7: even tiles, warps 8..11: odd tiles) =====================
    constexpr int CW = C::SLAB;               // one warp handles its 32 rows x the whole 32-column slab
    const int lane_group = warp & 3;
    const int pipe = (warp - 4) >> 2;
    const int half = 0;
    const float act_slope = P.act == ACT_LRELU ? P.slope : 1.0f;
    const int row_in_tile = lane_group * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t row_off = static_cast<uint32_t>(row_in_tile * C::ROWB);
    const uint32_t sw = (row_off >> 7) & C::SW_MASK;   // XOR pattern of this row's 16-byte chunks
    const int acc = pipe;
    uint32_t n_done = 0;
    Ring re(pipe * (E / 2), E / 2);
    // row validity of this group's NEXT tile is fetched while the current one is processed; the raw byte
    // stays in a register and is only compared one tile later, so the load is off the critical path
    auto row_valid = [&](int grp) -> unsigned {
      if (grp >= P.num_groups) return 0u;
      const int row = ((grp / P.num_n_tiles) * cl + rank) * BLOCK_M2 + row_in_tile;
      if (row >= P.m_rows) return 0u;
      const long long orow = static_cast<long long>(row) * P.mask_mul + mask_add_of(grp % P.num_n_tiles);
      if (orow < 0 || orow >= P.out_rows) return 0u;
      return P.frame_mask ? static_cast<unsigned>(__ldg(P.frame_mask + orow / P.rate)) : 1u;
    };
    unsigned valid_next = row_valid(cid + pipe * ncl);
    for (int grp = cid, seq = 0; grp < P.num_groups; grp += ncl, ++seq) {
      if ((seq & 1) != pipe) continue;
      const int n0 = (grp % P.num_n_tiles) * BLOCK_N;
      const float row_scale = valid_next != 0u ? P.post_scale : 0.0f;
      valid_next = row_valid(grp + 2 * ncl);
      if (warp == 4 && lane == 0) JB_TRACE(4, 3, seq);
      mbar_wait(&tfull_bar[acc], n_done & 1);
      ++n_done;
      if (warp == 4 && lane == 0) JB_TRACE(4, 0, seq);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < P.n_slabs; ++s) {
        // each of the two warps of a TMEM lane group owns half of the slab's columns
        uint32_t r[CW];
        {
          const uint32_t ta = lane_addr + static_cast<uint32_t>(acc * BLOCK_N + s * C::SLAB + half * CW);
          if (CW == 32) {
            uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
            tmem_ld32(ta, r0);
          } else {
            uint32_t(&r0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
            tmem_ld16(ta, r0);
          }
        }
        const int e = re.idx;
        mbar_wait(&epfull_bar[e], re.phase);   // residual / branch-sum slabs have landed (or entry is free)
        re.next();
        if (s == 0 && warp == 4 && lane == 0) JB_TRACE(4, 1, seq);
        tmem_ld_wait();
        if (s == P.n_slabs - 1) {
          // every TMEM read of this accumulator has completed: hand it back to the MMA warp right away
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // the leader's MMA thread waits
            else mbar_arrive(&tempty_bar[acc]);
          }
        }
        uint8_t* bufA = ep_base + e * entry_bytes;
        uint8_t* bufB = bufA + (P.ep_bufs - 1) * C::SLAB_BYTES;   // == bufA when an entry is a single slab
        const uint32_t bs_addr = smem_u32(bias_s) + static_cast<uint32_t>((n_ph ? 0 : n0) + s * C::SLAB + half * CW) * 4u;
#pragma unroll
        for (int c = 0; c < CW / 8; ++c) {
          const uint32_t off = row_off + ((static_cast<uint32_t>(half * (CW / 8) + c) ^ sw) << 4);
          float v[8];
          {
            const uint4 b0 = lds128(bs_addr + c * 32), b1 = lds128(bs_addr + c * 32 + 16);   // 8 bias values, shared-space loads
            const uint32_t bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = __uint_as_float(r[c * 8 + i]) + __uint_as_float(bv[i]);
              v[i] = fmaxf(x, x * act_slope);   // LeakyReLU (slope in (0,1)); act_slope == 1 -> identity
            }
          }
          if (P.has_res) {
            const uint4 t = lds128(smem_u32(bufA) + off);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 f = __bfloat1622float2(h[j]);
              if (P.res_inv_slope != 0.f) {
                f.x = f.x >= 0.f ? f.x : f.x * P.res_inv_slope;
                f.y = f.y >= 0.f ? f.y : f.y * P.res_inv_slope;
              }
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
          if (P.has_acc) {
            const uint4 t = lds128(smem_u32(bufB) + off);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              v[2 * j] += f.x; v[2 * j + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] *= row_scale;   // post_scale, or 0 for rows outside every utterance
          if (P.has_out0) {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            sts128(smem_u32(bufA) + off, o);
          }
          if (P.has_out1) {
            const float sl = P.out1_slope;
            uint4 o;
            o.x = pack_bf16x2(fmaxf(v[0], v[0] * sl), fmaxf(v[1], v[1] * sl));
            o.y = pack_bf16x2(fmaxf(v[2], v[2] * sl), fmaxf(v[3], v[3] * sl));
            o.z = pack_bf16x2(fmaxf(v[4], v[4] * sl), fmaxf(v[5], v[5] * sl));
            o.w = pack_bf16x2(fmaxf(v[6], v[6] * sl), fmaxf(v[7], v[7] * sl));
            sts128(smem_u32(bufB) + off, o);
          }
        }
        fence_proxy_async_smem();            // make the generic-proxy smem writes visible to the TMA engine
        mbar_arrive(&ready_bar[e]);          // hand the slab to the store warp; nobody waits here
        if (s == 0 && warp == 4 && lane == 0) JB_TRACE(4, 2, seq);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cl > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"(static_cast<uint32_t>(C::TMEM_COLS))
                   : "memory");
    else
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
  if (e.res_inv_slope != 0.f && !e.res_bf16) return false;
  const bool phase = p.w_tap_stride != 0;
  // n < n_pad (one partly filled N tile): the output tensor map has n columns, so the TMA stores of the padding slabs
  // are clipped; the weight rows and bias entries beyond n are zero (the narrow transposed convolutions, N = 96 of 128)
  const bool partial_tile = !phase && p.n < p.n_pad && p.n_pad == p.block_n && p.n % 32 == 0 && !e.res_bf16 && !e.accum_bf16;
  if ((!phase && p.n != p.n_pad && !partial_tile) || p.n > 512 || (p.n % p.block_n != 0 && !partial_tile)) return false;
  if (!phase && p.out_rows != p.m_rows) return false;
  const int slab = p.block_n >= 64 ? 64 : 32;
  auto ok = [&](const void* ptr, int ld) { return ptr == nullptr || (ld % 8 == 0 && ld >= p.n && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0); };
  if (!ok(e.res_bf16, e.res_ld) || !ok(e.accum_bf16, e.res_ld) || !ok(e.out_hi, e.out_bf_ld) || !ok(e.out_act, e.out_act_ld))
    return false;
  if (!e.out_hi && !e.out_act) return false;
  if (p.phases != 0) {   // all phases of a transposed convolution in one launch
    if (p.phases < 2 || p.phases > 5 || p.phase_pp < 0 || p.phase_pp >= p.phases || !phase || p.n != p.block_n || p.taps != 2 ||
        p.tap_stride != 1 || e.res_bf16 || e.accum_bf16 || e.out_hi || !e.out_act || p.out_pitch_mul != p.phases ||
        p.mask_mul != p.phases)
      return false;
  }
  (void)slab;
  return true;
}


template <int BLOCK_N, int KCH, int MODE, bool PAIR = false>
static int launch2(const ConvGemmProblem& p, cudaStream_t stream) {
  using C = Cfg2<BLOCK_N, KCH, MODE>;
  const ConvGemmEpilogue& e = p.ep;
  CUtensorMap ta, tb, tbm, tres, tacc, to0, to1, tph4;
  const int a_cols = p.a_cols > 0 ? p.a_cols : p.k_pad;
  const int k_chunks = ceil_div(a_cols, KCH);   // channel padding beyond a_cols is all-zero: skip it
  const int halo_rows = round_up(BLOCK_M2 + (p.taps - 1) * p.tap_stride, 8);
  JB_PROPAGATE(make_tmap(&ta, p.a_hi, p.a_rows, a_cols, p.a_ld, MODE == MODE_STREAM ? BLOCK_M2 : halo_rows, KCH));
  const long long w_rows = p.w_tap_stride != 0 ? static_cast<long long>(p.w_rows_total) : static_cast<long long>(p.taps) * p.n_pad;
  JB_PROPAGATE(make_tmap(&tb, p.w_hi, w_rows, p.k_pad, p.k_pad, BLOCK_N, KCH));
  // HALO mode streams every weight tile once per M tile from L2 (measured: ~7.5 TB/s aggregate, the bound of the
  // C = 128 / 256 stages): CTAs of a cluster work on neighbouring M tiles in lockstep and share each weight tile
  // through TMA multicast, every CTA fetching 1/cl of it.
  static const int env_cl = getenv("JATTS_B200_TC2_CL") ? atoi(getenv("JATTS_B200_TC2_CL")) : 1;
  // CTA pairs (cta_group::2, M = 256): each SM reads and is sent half of every weight tile, which takes the
  // N = 128 convolutions off the shared-memory port ceiling
  int cl = 1;
  const int pair = PAIR ? 1 : 0;
  if (PAIR) cl = 2;
  else if (MODE == MODE_HALO && (env_cl == 2 || env_cl == 4) && ceil_div(p.m_rows, BLOCK_M2) >= 2 * env_cl) cl = env_cl;
  tbm = tb;
  if (cl > 1) JB_PROPAGATE(make_tmap(&tbm, p.w_hi, w_rows, p.k_pad, p.k_pad, BLOCK_N / cl, KCH));
  tres = tacc = to0 = to1 = tph4 = ta;
  if (e.res_bf16) JB_PROPAGATE(make_tmap(&tres, e.res_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  if (e.accum_bf16) JB_PROPAGATE(make_tmap(&tacc, e.accum_bf16, p.m_rows, p.n, e.res_ld, BLOCK_M2, C::SLAB));
  const int pitch_mul = p.out_pitch_mul > 0 ? p.out_pitch_mul : 1;
  const long long view_rows = p.out_pitch_mul > 0 ? p.out_view_rows : p.m_rows;
  if (e.out_hi) JB_PROPAGATE(make_tmap(&to0, e.out_hi, view_rows, p.n, e.out_bf_ld * pitch_mul, BLOCK_M2, C::SLAB));
  if (e.out_act && p.phases == 0)
    JB_PROPAGATE(make_tmap(&to1, e.out_act, view_rows, p.n, e.out_act_ld * pitch_mul, BLOCK_M2, C::SLAB));
  if (p.phases > 0) {
    // phase q stores to every s-th row starting at first(q): one strided view per phase, in the slots the kernel's
    // ph_maps[] names (out1, out0, res, acc, ph4)
    CUtensorMap* slots[5] = {&to1, &to0, &tres, &tacc, &tph4};
    for (int q = 0; q < p.phases; ++q) {
      const long long first = q >= p.phase_pp ? q - p.phase_pp : q - p.phase_pp + p.phases;
      const long long rows_q = (static_cast<long long>(p.out_rows) - first + p.phases - 1) / p.phases;
      JB_PROPAGATE(make_tmap(slots[q], e.out_act + first * e.out_act_ld, rows_q > 0 ? rows_q : 1, p.n, e.out_act_ld * p.phases,
                             BLOCK_M2, C::SLAB));
    }
  }
  KernelParams2 kp;
  kp.taps = p.taps;
  kp.k_chunks = k_chunks;
  kp.n_pad = p.w_tap_stride != 0 ? p.n : p.n_pad;   // bias entries / columns per tile row
  kp.tap_off0 = p.tap_off0;
  kp.tap_stride = p.tap_stride;
  kp.m_rows = p.m_rows;
  kp.num_m_tiles = ceil_div(p.m_rows, BLOCK_M2);
  kp.num_n_tiles = p.phases > 0 ? p.phases : (p.w_tap_stride != 0 ? p.n : p.n_pad) / BLOCK_N;
  // a partly filled N tile (n < n_pad, one tile): the slabs beyond n hold no output column and are not processed at all
  kp.n_slabs = (p.phases == 0 && p.n < p.n_pad) ? ceil_div(p.n, C::SLAB) : C::N_SLABS;
  kp.n_phases = p.phases;
  kp.phase_pp = p.phase_pp;
  kp.cl = cl;
  kp.pair = pair;
  kp.num_groups = ceil_div(kp.num_m_tiles, cl) * kp.num_n_tiles;
  kp.frame_mask = p.frame_mask;
  kp.rate = p.rate > 0 ? p.rate : 1;
  kp.bias = e.bias;
  kp.act = e.act;
  kp.slope = e.slope;
  kp.post_scale = e.post_scale;
  kp.out1_slope = e.out_act_slope;
  kp.res_inv_slope = e.res_inv_slope;
  kp.has_res = e.res_bf16 != nullptr;
  kp.has_acc = e.accum_bf16 != nullptr;
  kp.has_out0 = e.out_hi != nullptr;
  kp.has_out1 = e.out_act != nullptr;
  JB_REQUIRE(p.phases == 0 || (!kp.has_res && !kp.has_acc && !kp.has_out0 && kp.has_out1 && MODE != MODE_RESIDENT), -2,
             "conv_gemm_tc2: phase mode takes a plain convolution with one activated output and streamed weights");
  kp.halo_rows = halo_rows;
  kp.w_row0 = p.w_row0;
  kp.w_tap_stride = p.w_tap_stride != 0 ? p.w_tap_stride : p.n_pad;
  kp.store_row_off = p.store_row_off;
  kp.mask_mul = p.mask_mul > 0 ? p.mask_mul : 1;
  kp.mask_add = p.mask_add;
  kp.out_rows = p.out_rows;
  const int w_bytes = MODE == MODE_RESIDENT ? round_up(p.taps * k_chunks * C::B_BYTES, 1024) : 0;
  // epilogue ring: every byte the mainloop leaves, up to 8 entries
  kp.ep_bufs = (kp.has_res + kp.has_acc > 1 || kp.has_out0 + kp.has_out1 > 1) ? 2 : 1;
  const int entry = kp.ep_bufs * C::SLAB_BYTES;
  int entries = (227 * 1024 - C::SMEM_FIXED - w_bytes) / entry;
  if (entries > C::MAX_ENTRIES) entries = C::MAX_ENTRIES;
  entries &= ~1;   // half of the ring per pipeline
  JB_REQUIRE(entries >= 2, -2, "conv_gemm_tc2: shared memory budget exceeded");
  kp.ep_entries = entries;
  kp.w_bytes = w_bytes;
  kp.trace = g_trace_ptr;
  kp.store_depth = entries >= 6 ? 2 : (entries >= 4 ? 1 : 0);
  static const char* sd_env = getenv("JATTS_B200_STORE_DEPTH");
  if (sd_env) kp.store_depth = atoi(sd_env) < entries - 1 ? atoi(sd_env) : entries - 2;
  const int smem_bytes = C::SMEM_FIXED + w_bytes + entries * entry;
  auto kern = conv_bf16_tma_kernel<BLOCK_N, KCH, MODE, PAIR>;
  JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_bytes));
  if (kp.num_groups == 0) return 0;
  int max_clusters = num_sms() / cl;
  if (cl > 2) {
    // clusters of 4 must fit inside a GPC: ask how many can be resident with this much shared memory
    static int cached[8] = {0};
    if (cached[cl] == 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(num_sms() / cl * cl);
      qc.blockDim = dim3(kThreads2);
      qc.dynamicSmemBytes = smem_bytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = cl; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int n = 0;
      JB_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &qc));
      cached[cl] = n > 0 ? n : 1;
    }
    if (cached[cl] < max_clusters) max_clusters = cached[cl];
  }
  const int grid = (kp.num_groups < max_clusters ? kp.num_groups : max_clusters) * cl;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventCreate(&e0));
    JB_CUDA_OK(cudaEventCreate(&e1));
    JB_CUDA_OK(cudaEventRecord(e0, stream));
  }
  JB_CUDA_OK(launch_tc(kern, grid, kThreads2, smem_bytes, stream, cl, ta, tb, tbm, tres, tacc, to0, to1, tph4, kp));
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
    return p.phases == 0 && halo_ok && max_mode >= MODE_RESIDENT && (p.w_tap_stride != 0 ? p.n : p.n_pad) == p.block_n &&
           fixed + round_up(p.taps * ceil_div(a_cols, kch) * b_bytes, 1024) + 4 * 2 * 8192 <= 227 * 1024;
  };
  // CTA pairs (cta_group::2, M = 256): each SM reads and is sent half of every weight tile, which takes the
  // N >= 128 convolutions off the shared-memory port ceiling
  static const int env_pair = getenv("JATTS_B200_TC2_PAIR") ? atoi(getenv("JATTS_B200_TC2_PAIR")) : 1;
  const bool pair_ok = env_pair != 0 && ceil_div(p.m_rows, BLOCK_M2) >= 4;
  switch (p.block_n) {
    case 32:
      if (a_cols <= 32 && resident_ok(Cfg2<32, 32, MODE_RESIDENT>::SMEM_FIXED, Cfg2<32, 32, MODE_RESIDENT>::B_BYTES, 32))
        return launch2<32, 32, MODE_RESIDENT>(p, stream);
      if (resident_ok(Cfg2<32, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<32, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<32, 64, MODE_RESIDENT>(p, stream);
      if (p.phases > 0 && halo_ok) return launch2<32, 64, MODE_HALO>(p, stream);
      return launch2<32, 64, MODE_STREAM>(p, stream);
    case 64:
      if (resident_ok(Cfg2<64, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<64, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<64, 64, MODE_RESIDENT>(p, stream);
      if (p.phases > 0 && halo_ok) return launch2<64, 64, MODE_HALO>(p, stream);
      return launch2<64, 64, MODE_STREAM>(p, stream);
    case 128:
      if (resident_ok(Cfg2<128, 64, MODE_RESIDENT>::SMEM_FIXED, Cfg2<128, 64, MODE_RESIDENT>::B_BYTES, 64))
        return launch2<128, 64, MODE_RESIDENT>(p, stream);
      if (halo_ok && pair_ok) return launch2<128, 64, MODE_HALO, true>(p, stream);
      return halo_ok ? launch2<128, 64, MODE_HALO>(p, stream) : launch2<128, 64, MODE_STREAM>(p, stream);
    case 256:
      if (halo_ok && pair_ok) return launch2<256, 64, MODE_HALO, true>(p, stream);
      return halo_ok ? launch2<256, 64, MODE_HALO>(p, stream) : launch2<256, 64, MODE_STREAM>(p, stream);
  }
  return -2;
}

}  // namespace jb
