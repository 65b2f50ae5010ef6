// mrf_pair_kernel: one HiFi-GAN residual unit (ResBlock1 inner step, oracle/hifigan.py `resblock`) in ONE kernel:
//
//     t   = LeakyReLU( Conv1d(xa, W1, dilation d) + b1 )              xa = LeakyReLU(x) as stored in HBM
//     v   = Conv1d(t, W2, dilation 1) + b2 + x  [+ branch sum]        x recovered from xa: xa >= 0 ? xa : xa / slope
//     out = bf16( LeakyReLU(v * post_scale, out_slope) )              out_slope = 1 -> plain store (branch sum)
//
// for the narrow stages (C = 32 / 64 channels) where the unfused convolutions are HBM bound: per pair the
// unfused path moves 6 activation tensors through HBM (read xa, write t, read t, read x, write x, write xa),
// this kernel moves 2 (read xa with its halo, write out).  The intermediate t never leaves shared memory.
//
// Tile = 128 rows of t = OUT_M = 128 - (k-1) output rows.  Per tile, per CTA (persistent, one CTA per SM):
//
//   warp 0      TMA producer: activation slab [128 + (k-1)*d (+ alignment pad) rows x C] -> smem (SWIZZLE = row)
//   warp 1      MMA issuer (one thread): conv1 = k taps, A descriptor starts tap*d rows into the slab (the
//               swizzle is a function of the absolute smem address, so row-shifted starts need no base offset);
//               conv2 = k taps over the t slab the E1 warps wrote.  Issue order C1(n+1), C2(n), C1(n+2), ... so
//               the tensor pipe runs conv1 of the next tile while E1 converts the current one.
//   warp 2      TMA store: the output tile is written IN PLACE over the residual rows of the activation slab
//               (row o + halo, 1024-byte aligned thanks to the pad) and stored from there; the slab goes back
//               to the producer when the store has read it.
//   warp 3      (W2 streamed) second TMA producer for the conv2 weight ring, when both weight sets do not fit
//   warps 4-11  E1 (two groups of 4 warps: even / odd tiles): tcgen05.ld T -> +b1 -> LeakyReLU -> row mask -> bf16 ->
//               swizzled K-major t slab (UMMA A operand)
//   warps 12-19 E2 (two groups, even / odd tiles): tcgen05.ld U -> +b2 + x (from the slab) [+ branch sum,
//               prefetched to registers] -> out.  Two groups per role because one tile's handoff chain
//               (commit -> mbarrier wake-up -> ld -> math -> fence -> arrive) is latency, not issue, bound.
//
// TMEM: T[2], U[2] accumulators of C columns each.  Weights: W1 resident; W2 resident or streamed.
// Rows outside every utterance are stored as zeros (the packed layout's gap rows stay zero); t rows
// outside an utterance are zeroed too, which IS conv2's zero padding at the utterance ends.
#include <cstdlib>

#include "../../include/jatts_b200.h"
#include "conv_gemm.cuh"
#include "mrf_pair.cuh"
#include "tc_common.cuh"

namespace jb {

extern long long* g_trace_ptr;   // conv_gemm_tc2.cu

namespace {

constexpr int kThreadsP = 640;   // 4 control warps (producer, conv1 issuer, store, conv2 issuer) + 2 x 4 E1 + 2 x 4 E2 warps
constexpr int kTileM = 128;
constexpr int kTRows = 144;   // t slab rows: 128 written + 16 zero rows read by the shifted conv2 taps (k <= 17)
constexpr int kMaxSA = 8, kMaxRB = 8;

struct PairParams {
  int taps, dil;
  int pad;            // alignment rows in front of the slab so that the residual rows start 1024-B aligned
  int halo;           // (k-1)/2 * (dil + 1): slab row of output row 0 is halo + pad
  int h2;             // (k-1)/2
  int out_m;          // 128 - (k-1)
  int slab_rows;      // rows loaded per tile
  int m_rows, num_tiles;
  int sa, nt, rb;     // activation slabs, t slabs, W2 ring depth (0 = W2 resident)
  int dbg;            // debug (JATTS_B200_PAIR_DEBUG, results are WRONG): 1 = epilogues skip their shared-memory traffic,
                      // 2 = also skip the TMEM loads; isolates the MMA/TMA pipeline from the epilogue work in timelines
  int dual;           // 1: conv1 phases are issued by warp 1, conv2 phases by warp 3 (two issuing threads; W2 resident only)
  int la;             // conv1 look-ahead: conv1 of tile n+la is issued before conv2 of tile n (1..3); la+1 T accumulators
  int store_lag;      // 1: keep one TMA store in flight behind the newest (its slab is released one tile later)
  int leader_poll;    // 1: one warp per epilogue group polls the mbarriers, the others sleep in bar.sync
  int slab_bytes;     // bytes between activation slabs (slab_rows * row bytes rounded up to 1024)
  int w_rows_per_tap; // weight rows between taps (n_pad)
  const uint8_t* frame_mask;
  int rate;
  float slope, inv_slope, post_scale, out_slope;
  const bf16* acc;
  int acc_ld;
  long long* trace;
  // biases live in the kernel parameters (constant bank): the unrolled epilogues read them as immediate-offset
  // constant operands, with no shared-memory traffic -- the MMA operand reads already saturate the smem port
  float b1[64], b2[64];   // debug: clock64 timeline of CTA 0 (jatts_debug_set_trace), else null
};

#ifdef JB_ENABLE_TRACE
#define PT(role, ev, idx)                                                                                      \
  do {                                                                                                         \
    if (P.trace && blockIdx.x == 0 && (idx) < 64) P.trace[((role) * 8 + (ev)) * 64 + (idx)] = clock64();       \
  } while (0)
#else
#define PT(role, ev, idx) do { } while (0)   // the ~40 stamp sites cost ~300 instructions of a cache-sensitive kernel
#endif

struct Ring {
  int idx, depth;
  uint32_t phase;
  __device__ __forceinline__ explicit Ring(int depth_) : idx(0), depth(depth_), phase(0) {}
  __device__ __forceinline__ void next() {
    if (++idx == depth) { idx = 0; phase ^= 1u; }
  }
};

template <int C>
struct CfgP {
  static constexpr int KROWB = C * 2;                    // operand row bytes = swizzle span (64 or 128)
  static constexpr int KSTEPS = C / 16;
  static constexpr int ROW_ALIGN = 1024 / KROWB;         // rows per 1024 B
  static constexpr int SLAB_CAP = C == 32 ? 208 : 192;   // 128 + 10*5 + pad (<= ROW_ALIGN-1), bytes % 1024 == 0
  static constexpr int A_SLAB_BYTES = SLAB_CAP * KROWB;
  static constexpr int T_BYTES = kTRows * KROWB;
  static constexpr int B_BYTES = C * KROWB;              // one tap's [C x C] weight tile
  static constexpr int TAIL_BYTES = 2048;                // biases (2*C floats) + barriers
  static constexpr int TMEM_COLS = 8 * C;                // T[<= 4] | U[2] accumulators of C columns (6 C used): 256 / 512
  static_assert(A_SLAB_BYTES % 1024 == 0 && T_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "swizzle alignment");
};

template <int C>
__device__ __forceinline__ uint32_t swz(int row) {   // XOR pattern of the 16-byte chunks of this smem row
  return C == 64 ? static_cast<uint32_t>(row & 7) : static_cast<uint32_t>((row >> 1) & 3);
}

template <int C, int TAPS>
__global__ void __launch_bounds__(kThreadsP, 1)
mrf_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out,
                const __grid_constant__ PairParams P) {
  using K = CfgP<C>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab_base = smem;                                   // [sa][A_SLAB_BYTES]
  uint8_t* t_base = slab_base + P.sa * P.slab_bytes;        // [nt][T_BYTES]
  float* bias_s = reinterpret_cast<float*>(t_base + P.nt * K::T_BYTES);   // [2][C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + 1024);
  uint8_t* w1_base = reinterpret_cast<uint8_t*>(bias_s) + K::TAIL_BYTES;  // [taps][B_BYTES]
  uint8_t* w2_base = w1_base + P.taps * K::B_BYTES;                       // [taps] resident or [rb] ring
  uint64_t* xa_full = bars;                 // [4]
  uint64_t* xa_empty = xa_full + kMaxSA;    // [4]
  uint64_t* out_ready = xa_empty + kMaxSA;  // [4]
  uint64_t* T_full = out_ready + kMaxSA;    // [4]
  uint64_t* T_empty = T_full + 4;           // [4]
  uint64_t* U_full = T_empty + 4;           // [2]
  uint64_t* U_empty = U_full + 2;           // [2]
  uint64_t* t_full = U_empty + 2;           // [2]
  uint64_t* t_empty = t_full + 2;           // [2]
  uint64_t* w2_full = t_empty + 2;          // [8]
  uint64_t* w2_empty = w2_full + kMaxRB;    // [8]
  uint64_t* w_full = w2_empty + kMaxRB;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int my_tiles = (P.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const bool stream_w2 = P.rb > 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_out);
    for (int i = 0; i < kMaxSA; ++i) {
      mbar_init(&xa_full[i], 1);
      mbar_init(&xa_empty[i], 1);
      mbar_init(&out_ready[i], 4);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&T_full[i], 1);
      mbar_init(&T_empty[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&U_full[i], 1);
      mbar_init(&U_empty[i], 4);
      mbar_init(&t_full[i], 4);
      mbar_init(&t_empty[i], 1);
    }
    for (int i = 0; i < kMaxRB; ++i) {
      mbar_init(&w2_full[i], 1);
      mbar_init(&w2_empty[i], 1);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(K::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the t slabs' 16 tail rows are read by the shifted conv2 taps but never written: zero everything once
  for (int i = threadIdx.x; i < P.nt * K::T_BYTES / 16; i += kThreadsP)
    sts128(smem_u32(t_base) + static_cast<uint32_t>(i) * 16u, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // prologue above overlapped the previous kernel's tail; its outputs are visible from here

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one activation slab per tile =====================
    if (elect_one()) {
      mbar_expect_tx(w_full, static_cast<uint32_t>(P.taps * (stream_w2 ? 1 : 2)) * K::B_BYTES);
      for (int tap = 0; tap < P.taps; ++tap) tma_load_2d(&tm_w1, w_full, w1_base + tap * K::B_BYTES, 0, tap * P.w_rows_per_tap);
      if (!stream_w2)
        for (int tap = 0; tap < P.taps; ++tap) tma_load_2d(&tm_w2, w_full, w2_base + tap * K::B_BYTES, 0, tap * P.w_rows_per_tap);
      const uint32_t a_bytes = static_cast<uint32_t>(P.slab_rows) * K::KROWB;
      Ring ra(P.sa);
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        mbar_wait(&xa_empty[ra.idx], ra.phase ^ 1);
        PT(0, 0, i);
        mbar_expect_tx(&xa_full[ra.idx], a_bytes);
        tma_load_2d(&tm_a, &xa_full[ra.idx], slab_base + ra.idx * P.slab_bytes, 0, tile * P.out_m - P.halo - P.pad);
        ra.next();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(kTileM, C, /*is_bf16=*/true);
      constexpr uint32_t desc_lo0 = 1u << 16;
      constexpr uint32_t desc_hi = static_cast<uint32_t>((8 * K::KROWB) >> 4) | (1u << 14) | (static_cast<uint32_t>(C == 64 ? 2 : 4) << 29);
      const uint32_t w1_addr = smem_u32(w1_base), w2_addr = smem_u32(w2_base);
      const uint32_t slab_addr = smem_u32(slab_base), t_addr = smem_u32(t_base);
      Ring ra(P.sa), rt(P.nt), rb(stream_w2 ? P.rb : 1), rT(P.la + 1);
      mbar_wait(w_full, 0);
      tc_fence_after();
      // Issue order C1(0..la-1), then C1(n+la), C2(n) alternately: E1 of tile n (accumulator hand-off, conversion,
      // fence, hand-back: ~1.7-2.7 k clk) has la conv1 phases plus a conv2 phase of tensor-pipe work to hide behind.
      // Measured with the epilogue work disabled (JATTS_B200_PAIR_DEBUG=1): the MMAs run at the SS rate (44 clk at
      // N = 32) but every phase switch (commit, two mbarrier polls, fence, descriptor set-up: ~190 clk, 4 per tile)
      // is idle tensor-pipe time because the issue queue is only 1-2 instructions deep: 2700 clk per tile at k = 11
      // against 1971 clk of MMAs.  Probing the next phase's barriers before the last tap (mbarrier.test_wait) made
      // every shape 3-15 % slower and was dropped.
      for (int i = 0; i < my_tiles + P.la; ++i) {
        if (i < my_tiles) {
          // ---- conv1 of tile i -> T[i % (la+1)]
          const int b = rT.idx;
          mbar_wait(&xa_full[ra.idx], ra.phase);
          PT(1, 4, i);
          mbar_wait(&T_empty[b], rT.phase ^ 1);
          PT(1, 0, i);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(b * C);
          const uint32_t a0 = slab_addr + static_cast<uint32_t>(ra.idx * P.slab_bytes + P.pad * K::KROWB);
          {
            const uint32_t a_lo0 = desc_lo0 + (a0 >> 4);
            const uint32_t b_lo0 = desc_lo0 + (w1_addr >> 4);
            const uint32_t a_step = (static_cast<uint32_t>(P.dil) * K::KROWB) >> 4;
#pragma unroll 1
            for (int tap = 0; tap < TAPS; ++tap) {
#pragma unroll
              for (int k = 0; k < K::KSTEPS; ++k)
                tc_mma_bf16_lohi(tmem_d, a_lo0 + tap * a_step + 2 * k, desc_hi, b_lo0 + tap * (K::B_BYTES >> 4) + 2 * k, desc_hi,
                                 idesc, (tap | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&T_full[b]);
          PT(1, 1, i);
          ra.next();
          rT.next();
        }
        if (i >= P.la && !P.dual) {
          // ---- conv2 of tile i-la -> U[(i-la) & 1]
          const int m = i - P.la, b = m & 1;
          mbar_wait(&t_full[rt.idx], rt.phase);
          PT(1, 5, m);
          mbar_wait(&U_empty[b], ((m >> 1) & 1) ^ 1);
          PT(1, 2, m);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(4 * C + b * C);
          const uint32_t a0 = t_addr + static_cast<uint32_t>(rt.idx * K::T_BYTES);
#pragma unroll 1
          for (int tap = 0; tap < TAPS; ++tap) {
            uint32_t b_addr;
            if (stream_w2) {
              mbar_wait(&w2_full[rb.idx], rb.phase);
              tc_fence_after();
              b_addr = w2_addr + static_cast<uint32_t>(rb.idx) * K::B_BYTES;
            } else {
              b_addr = w2_addr + static_cast<uint32_t>(tap) * K::B_BYTES;
            }
            const uint32_t a_lo = desc_lo0 + ((a0 + static_cast<uint32_t>(tap) * K::KROWB) >> 4);
            const uint32_t b_lo = desc_lo0 + (b_addr >> 4);
#pragma unroll
            for (int k = 0; k < K::KSTEPS; ++k)
              tc_mma_bf16_lohi(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (tap | k) != 0 ? 1u : 0u);
            if (stream_w2) {
              tc_commit(&w2_empty[rb.idx]);
              rb.next();
            }
          }
          tc_commit(&U_full[b]);
          tc_commit(&t_empty[rt.idx]);
          PT(1, 3, m);
          rt.next();
        }
      }
    }
  } else if (warp == 2) {
    // ===================== store thread: output rows of the slab -> global by TMA =====================
    if (elect_one()) {
      Ring ra(P.sa);
      int prev = 0;
      const int src_off = (P.halo + P.pad) * K::KROWB;   // multiple of 1024
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        mbar_wait(&out_ready[ra.idx], ra.phase);
        PT(4, 0, i);
        tma_store_2d(&tm_out, slab_base + ra.idx * P.slab_bytes + src_off, 0, tile * P.out_m);
        tma_store_commit();
        if (P.store_lag) {
          // keep one store in flight behind the one just committed; the slab of the older one goes back to the producer
          tma_store_wait_read<1>();
          if (i > 0) mbar_arrive(&xa_empty[prev]);
        } else {
          tma_store_wait_read<0>();
          mbar_arrive(&xa_empty[ra.idx]);
        }
        PT(4, 1, i);
        prev = ra.idx;
        ra.next();
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp == 3) {
    if (stream_w2) {
      // ===================== conv2 weight ring (W2 not resident: warp 1 issues both convolutions) =====================
      if (elect_one()) {
        Ring rb(P.rb);
        for (int i = 0; i < my_tiles; ++i)
          for (int tap = 0; tap < P.taps; ++tap) {
            mbar_wait(&w2_empty[rb.idx], rb.phase ^ 1);
            mbar_expect_tx(&w2_full[rb.idx], K::B_BYTES);
            tma_load_2d(&tm_w2, &w2_full[rb.idx], w2_base + rb.idx * K::B_BYTES, 0, tap * P.w_rows_per_tap);
            rb.next();
          }
      }
    } else if (P.dual && elect_one()) {
      // ===================== second MMA issuer: every conv2 phase (W2 resident) =====================
      // While one issuing thread is between phases (commit, mbarrier polls, fence, descriptor set-up: ~190 clk with a
      // 1-2 deep issue queue) the other one's MMAs keep the tensor pipe busy; the two only meet on the pipe itself.
      // Measured per launch (us): C=32 k=3/7/11 285/453/738 -> 249/350/390, C=64 k=3/7 218/379 -> 153/268.
      constexpr uint32_t idesc = make_idesc(kTileM, C, /*is_bf16=*/true);
      constexpr uint32_t desc_lo0 = 1u << 16;
      constexpr uint32_t desc_hi = static_cast<uint32_t>((8 * K::KROWB) >> 4) | (1u << 14) | (static_cast<uint32_t>(C == 64 ? 2 : 4) << 29);
      const uint32_t w2_addr = smem_u32(w2_base), t_addr = smem_u32(t_base);
      Ring rt(P.nt), rb(stream_w2 ? P.rb : 1);
      mbar_wait(w_full, 0);
      tc_fence_after();
      for (int m = 0; m < my_tiles; ++m) {
        const int b = m & 1;
        mbar_wait(&t_full[rt.idx], rt.phase);
        PT(1, 5, m);
        mbar_wait(&U_empty[b], ((m >> 1) & 1) ^ 1);
        PT(1, 2, m);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(4 * C + b * C);
        const uint32_t a0 = t_addr + static_cast<uint32_t>(rt.idx * K::T_BYTES);
#pragma unroll 1
        for (int tap = 0; tap < TAPS; ++tap) {
          uint32_t b_addr;
          if (stream_w2) {
            mbar_wait(&w2_full[rb.idx], rb.phase);
            tc_fence_after();
            b_addr = w2_addr + static_cast<uint32_t>(rb.idx) * K::B_BYTES;
          } else {
            b_addr = w2_addr + static_cast<uint32_t>(tap) * K::B_BYTES;
          }
          const uint32_t a_lo = desc_lo0 + ((a0 + static_cast<uint32_t>(tap) * K::KROWB) >> 4);
          const uint32_t b_lo = desc_lo0 + (b_addr >> 4);
#pragma unroll
          for (int k = 0; k < K::KSTEPS; ++k)
            tc_mma_bf16_lohi(tmem_d, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (tap | k) != 0 ? 1u : 0u);
          if (stream_w2) {
            tc_commit(&w2_empty[rb.idx]);
            rb.next();
          }
        }
        tc_commit(&U_full[b]);
        tc_commit(&t_empty[rt.idx]);
        PT(1, 3, m);
        rt.next();
      }
    }
  } else if (warp < 12) {
    // ===================== E1 (group = tile parity): T -> t slab (bf16, swizzled K-major) =====================
    const int grp = (warp - 4) >> 2;
    const int lane_group = warp & 3;
    const int row = lane_group * 32 + lane;                   // t row of this thread
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16);
    const uint32_t row_off = static_cast<uint32_t>(row) * K::KROWB;
    const uint32_t sw = swz<C>(row);
    const int tb = P.nt == 2 ? grp : 0;                       // t slab (and its barriers) of this group
    uint32_t n_done = 0;                                      // tiles this group has processed
    Ring rT(P.la + 1);                                        // T accumulator of tile i: i % (la+1)
    if (grp) rT.next();
    for (int i = grp; i < my_tiles; i += 2, ++n_done) {
      const int tbuf = rT.idx;
      const uint32_t tphase = rT.phase;
      rT.next();
      rT.next();
      const uint32_t lane_addr = lane_base + static_cast<uint32_t>(tbuf * C);
      const int tile = blockIdx.x + i * gridDim.x;
      const int g = tile * P.out_m - P.h2 + row;
      float keep = 0.f;
      if (g >= 0 && g < P.m_rows) keep = (P.frame_mask == nullptr || __ldg(P.frame_mask + g / P.rate)) ? 1.f : 0.f;
      // one warp of the group polls the mbarriers, the other three sleep in bar.sync (a polling warp takes
      // issue slots from the warps doing the math: 20 warps share 4 schedulers)
      if (!P.leader_poll || lane_group == grp) {
        mbar_wait(&T_full[tbuf], tphase);
        // t slab free: conv2 of the tile that used it last has been committed (nt == 1: the previous tile, i-1)
        mbar_wait(&t_empty[tb], ((P.nt == 2 ? n_done : static_cast<uint32_t>(i)) & 1) ^ 1);
      }
      if (P.leader_poll) named_bar_sync(1 + grp, 128);
      if (lane_group == 0 && lane == 0) PT(2, 0, i);
      if (lane_group == 0 && lane == 0) PT(2, 1, i);
      tc_fence_after();
      const uint32_t trow = smem_u32(t_base + tb * K::T_BYTES) + row_off;
#pragma unroll 1
      for (int s = 0; s < C / 32; ++s) {
        uint32_t r[32];
        if (P.dbg < 2) {
          tmem_ld32(lane_addr + static_cast<uint32_t>(s * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        if (s == 0 && lane_group == 0 && lane == 0) PT(2, 3, i);
        if (s == C / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&T_empty[tbuf]);
        }
        if (P.dbg) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = __uint_as_float(r[c * 8 + j]) + P.b1[s * 32 + c * 8 + j];
            v[j] = fmaxf(x, x * P.slope) * keep;
          }
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
          sts128(trow + ((static_cast<uint32_t>(s * 4 + c) ^ sw) << 4), o);
        }
      }
      if (lane_group == 0 && lane == 0) PT(2, 4, i);
      fence_proxy_async_smem();
      if (lane_group == 0 && lane == 0) PT(2, 5, i);
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_full[tb]);
      if (lane_group == 0 && lane == 0) PT(2, 2, i);
    }
  } else {
    // ===================== E2 (group = tile parity): U + b2 + x [+ branch sum] -> output tile, in place =====================
    const int grp = (warp - 12) >> 2;
    const int lane_group = warp & 3;
    const int row = lane_group * 32 + lane;                   // output row of this thread within the tile
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(lane_group * 32) << 16) + static_cast<uint32_t>(4 * C + grp * C);
    const int srow = row + P.halo + P.pad;                    // slab row holding x[row] (and receiving out[row])
    const uint32_t row_off = static_cast<uint32_t>(srow) * K::KROWB;
    const uint32_t sw = swz<C>(srow);
    const bool active = row < P.out_m;
    Ring ra(P.sa);
    if (grp) ra.next();
    uint32_t n_done = 0;
    for (int i = grp; i < my_tiles; i += 2, ++n_done) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int g = tile * P.out_m + row;
      float keep = 0.f;
      if (active && g < P.m_rows) keep = (P.frame_mask == nullptr || __ldg(P.frame_mask + g / P.rate)) ? P.post_scale : 0.f;
      uint4 accv[C / 8];
      if (P.acc != nullptr) {
        // branch sum row, fetched before the accumulator wait so the latency hides behind conv2
        if (keep != 0.f) {
          const uint4* ap = reinterpret_cast<const uint4*>(P.acc + static_cast<long long>(g) * P.acc_ld);
#pragma unroll
          for (int c = 0; c < C / 8; ++c) accv[c] = __ldg(ap + c);
        } else {
#pragma unroll
          for (int c = 0; c < C / 8; ++c) accv[c] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (!P.leader_poll || lane_group == 2 + grp) {
        mbar_wait(&U_full[grp], n_done & 1);
        mbar_wait(&xa_full[ra.idx], ra.phase);   // completed long ago; acquires the TMA-written slab for the group
      }
      if (P.leader_poll) named_bar_sync(3 + grp, 128);
      if (lane_group == 0 && lane == 0) PT(3, 0, i);
      tc_fence_after();
      const uint32_t srow_p = smem_u32(slab_base + ra.idx * P.slab_bytes) + row_off;
#pragma unroll 1
      for (int s = 0; s < C / 32; ++s) {
        uint32_t r[32];
        if (P.dbg < 2) {
          tmem_ld32(lane_addr + static_cast<uint32_t>(s * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        if (s == 0 && lane_group == 0 && lane == 0) PT(3, 2, i);
        if (s == C / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&U_empty[grp]);
        }
        if (active && !P.dbg) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t p = srow_p + ((static_cast<uint32_t>(s * 4 + c) ^ sw) << 4);
            const uint4 xr = lds128(p);
            const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr);
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(xh[j]);
              // x from LeakyReLU(x): negative values were scaled by `slope`
              v[2 * j] = __uint_as_float(r[c * 8 + 2 * j]) + P.b2[s * 32 + c * 8 + 2 * j] + (f.x >= 0.f ? f.x : f.x * P.inv_slope);
              v[2 * j + 1] = __uint_as_float(r[c * 8 + 2 * j + 1]) + P.b2[s * 32 + c * 8 + 2 * j + 1] + (f.y >= 0.f ? f.y : f.y * P.inv_slope);
            }
            if (P.acc != nullptr) {
              const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&accv[s * 4 + c]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(ah[j]);
                v[2 * j] += f.x; v[2 * j + 1] += f.y;
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float y = v[j] * keep;
              v[j] = fmaxf(y, y * P.out_slope);
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            sts128(p, o);
          }
        }
      }
      if (lane_group == 0 && lane == 0) PT(3, 3, i);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_ready[ra.idx]);
      if (lane_group == 0 && lane == 0) PT(3, 1, i);
      ra.next();
      ra.next();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(K::TMEM_COLS))
                 : "memory");
  }
}

template <int C, int TAPS>
int launch_pair(const MrfPairProblem& p, cudaStream_t stream) {
  using K = CfgP<C>;
  PairParams kp{};
  kp.taps = p.taps;
  kp.dil = p.dilation;
  kp.h2 = (p.taps - 1) / 2;
  kp.halo = kp.h2 * (p.dilation + 1);
  kp.pad = (K::ROW_ALIGN - kp.halo % K::ROW_ALIGN) % K::ROW_ALIGN;
  kp.out_m = kTileM - (p.taps - 1);
  kp.slab_rows = kTileM + (p.taps - 1) * p.dilation + kp.pad;
  kp.m_rows = p.rows;
  kp.num_tiles = ceil_div(p.rows, kp.out_m);
  kp.w_rows_per_tap = p.n_pad;
  kp.frame_mask = p.frame_mask;
  kp.rate = p.rate > 0 ? p.rate : 1;
  for (int i = 0; i < C; ++i) {
    kp.b1[i] = p.h_b1[i];
    kp.b2[i] = p.h_b2[i];
  }
  kp.slope = p.slope;
  kp.inv_slope = 1.0f / p.slope;
  kp.post_scale = p.post_scale;
  kp.out_slope = p.out_slope;
  kp.acc = p.accum;
  kp.acc_ld = p.accum_ld;
  kp.trace = g_trace_ptr;
  static const int env_dbg = getenv("JATTS_B200_PAIR_DEBUG") ? atoi(getenv("JATTS_B200_PAIR_DEBUG")) : 0;
  kp.dbg = env_dbg;
  JB_REQUIRE(kp.slab_rows <= K::SLAB_CAP && kp.slab_rows <= 256, JATTS_E_UNSUPPORTED, "mrf_pair: receptive field too wide");
  JB_REQUIRE(kp.h2 + kTileM <= kTRows, JATTS_E_UNSUPPORTED, "mrf_pair: kernel size too large");
  // shared memory plan: everything resident with as many activation slabs as fit (a slab lives from its TMA load
  // to the TMA store of the tile, ~4-6 tile periods); if fewer than 4 slabs fit, stream W2 through a ring
  // (and, if still short, a single t slab: conv1 of the next tile covers E1)
  const int limit = 227 * 1024 - 1024;   // alignment slack
  const int w_one = p.taps * K::B_BYTES;
  const int slab_bytes = round_up(kp.slab_rows * K::KROWB, 1024);
  auto fixed = [&](int nt, int w2_bytes) { return nt * K::T_BYTES + K::TAIL_BYTES + w_one + w2_bytes; };
  static const int force_stream = getenv("JATTS_B200_PAIR_STREAM") ? atoi(getenv("JATTS_B200_PAIR_STREAM")) : 0;
  static const int force_sa = getenv("JATTS_B200_PAIR_SA") ? atoi(getenv("JATTS_B200_PAIR_SA")) : 0;
  static const int env_lag = getenv("JATTS_B200_PAIR_LAG") ? atoi(getenv("JATTS_B200_PAIR_LAG")) : -1;
  static const int env_poll = getenv("JATTS_B200_PAIR_POLL") ? atoi(getenv("JATTS_B200_PAIR_POLL")) : -1;
  int nt = 2, rb = 0;
  int sa = (limit - fixed(2, w_one)) / slab_bytes;
  if (force_stream || sa < 4) {
    rb = force_stream > 1 ? force_stream : 4;
    sa = (limit - fixed(nt, rb * K::B_BYTES)) / slab_bytes;
    if (sa < 3) {
      nt = 1;
      sa = (limit - fixed(nt, rb * K::B_BYTES)) / slab_bytes;
    }
    JB_REQUIRE(sa >= 2, JATTS_E_UNSUPPORTED, "mrf_pair: shared memory budget exceeded");
    if (sa > kMaxSA) sa = kMaxSA;
    // leftover bytes deepen the weight ring
    const int spare = (limit - fixed(nt, rb * K::B_BYTES) - sa * slab_bytes) / K::B_BYTES;
    rb = rb + spare > kMaxRB ? kMaxRB : rb + spare;
  }
  if (sa > kMaxSA) sa = kMaxSA;
  if (force_sa > 1 && force_sa < sa) sa = force_sa;
  kp.sa = sa;
  static const int env_dual = getenv("JATTS_B200_PAIR_DUAL") ? atoi(getenv("JATTS_B200_PAIR_DUAL")) : 1;
  kp.dual = (env_dual && rb == 0) ? 1 : 0;   // with a streamed W2 warp 3 is the weight producer: single issuer
  static const int env_la = getenv("JATTS_B200_PAIR_LA") ? atoi(getenv("JATTS_B200_PAIR_LA")) : 0;
  // a slab is held from its load until the tile's store, so la + 1 tiles are in use when conv1 of the next one
  // needs its slab: la <= sa - 2 (more would deadlock the producer against the store)
  // measured: la = 2 pays at C = 64 (k = 3: 213 -> 190 us, k = 7: 421 -> 370 us per launch), nothing at C = 32
  kp.la = C == 64 ? 2 : 1;
  if (env_la >= 1) kp.la = env_la;
  if (kp.la > sa - 2) kp.la = sa - 2;
  if (kp.la > 3) kp.la = 3;
  if (kp.la < 1) kp.la = 1;
  kp.store_lag = env_lag >= 0 ? env_lag : (sa >= 6 ? 1 : 0);
  kp.leader_poll = env_poll >= 0 ? env_poll : 1;
  kp.slab_bytes = slab_bytes;
  kp.nt = nt;
  kp.rb = rb;
  const int smem_bytes = sa * slab_bytes + fixed(nt, rb > 0 ? rb * K::B_BYTES : w_one) + 1024;
  CUtensorMap ta, tw1, tw2, tout;
  JB_PROPAGATE(make_tmap(&ta, p.xa, p.rows, C, p.ld, kp.slab_rows, C));
  JB_PROPAGATE(make_tmap(&tw1, p.w1, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, C, C));
  JB_PROPAGATE(make_tmap(&tw2, p.w2, static_cast<long long>(p.taps) * p.n_pad, p.k_pad, p.k_pad, C, C));
  JB_PROPAGATE(make_tmap(&tout, p.out, p.rows, C, p.out_ld, kp.out_m, C));
  auto kern = mrf_pair_kernel<C, TAPS>;
  JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem_bytes));
  if (kp.num_tiles == 0) return 0;
  const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventCreate(&e0));
    JB_CUDA_OK(cudaEventCreate(&e1));
    JB_CUDA_OK(cudaEventRecord(e0, stream));
  }
  JB_CUDA_OK(launch_tc(kern, grid, kThreadsP, smem_bytes, stream, 1, ta, tw1, tw2, tout, kp));
  JB_KERNEL_OK();
  if (g_profile_on) {
    JB_CUDA_OK(cudaEventRecord(e1, stream));
    g_profile_events.push_back({e0, e1, 0});
  }
  return 0;
}

}  // namespace

bool mrf_pair_eligible(const MrfPairProblem& p) {
  if (!(p.c == 32 || p.c == 64)) return false;
  if (p.n_pad != p.c || p.k_pad < p.c) return false;
  if ((p.taps & 1) == 0 || p.taps < 3 || p.taps > 11 || p.dilation < 1) return false;
  if (kTileM + (p.taps - 1) * p.dilation + (p.c == 32 ? 15 : 7) > (p.c == 32 ? 208 : 192)) return false;
  if (!(p.slope > 0.f && p.slope <= 1.f) || !(p.out_slope > 0.f && p.out_slope <= 1.f)) return false;
  auto ok = [&](const void* ptr, int ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && ld >= p.c; };
  if (!p.xa || !p.out || !p.w1 || !p.w2 || !p.h_b1 || !p.h_b2) return false;
  if (!ok(p.xa, p.ld) || !ok(p.out, p.out_ld)) return false;
  if (p.accum && !ok(p.accum, p.accum_ld)) return false;
  return true;
}

int mrf_pair(const MrfPairProblem& p, cudaStream_t stream) {
  JB_REQUIRE(mrf_pair_eligible(p), JATTS_E_UNSUPPORTED, "mrf_pair: problem not eligible for the fused kernel");
  switch (p.taps) {
    case 3: return p.c == 32 ? launch_pair<32, 3>(p, stream) : launch_pair<64, 3>(p, stream);
    case 5: return p.c == 32 ? launch_pair<32, 5>(p, stream) : launch_pair<64, 5>(p, stream);
    case 7: return p.c == 32 ? launch_pair<32, 7>(p, stream) : launch_pair<64, 7>(p, stream);
    case 9: return p.c == 32 ? launch_pair<32, 9>(p, stream) : launch_pair<64, 9>(p, stream);
    case 11: return p.c == 32 ? launch_pair<32, 11>(p, stream) : launch_pair<64, 11>(p, stream);
  }
  return JATTS_E_UNSUPPORTED;
}

}  // namespace jb
