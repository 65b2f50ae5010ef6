// relpos_attention_tc_kernel: legacy relative-position self attention (jatts/modules/transformer/attention.py:164-206,
// rel_shift :142-162) on tcgen05 tensor cores, TMEM accumulators and TMA operand loads, for any utterance length.
//
// Inputs are what the fused projection GEMM writes: one [rows, 4*D] matrix of fp16 (hi, lo*2^11) operand pairs per row
//     [ q + bias_u | q + bias_v | k | v ]              (the two biased copies of q cost one extra N block of that GEMM)
// and the per-layer table p = linear_pos(pe) [max_len, D] as the same kind of pair.  Per (utterance, head, tile of 127
// query rows) one persistent CTA runs, with a = query row, b = key, n = position row:
//
//   phase 1   AC[a][b]  = (q_a + u) . k_b          128 x 128 score tiles, 3 MMAs per K step (hi.hi -> main accumulator,
//   phase 2   BD[a][n]  = (q_a + v) . p_n          hi.lo' + lo'.hi -> a second one that the epilogue scales by 2^-11)
//             both are written UNSHIFTED to a per-CTA fp32 scratch (L2 resident: 2 x 128 x T floats per CTA)
//   softmax   the legacy rel-shift is a pure index map on the read side:
//                 s[a][b] = AC[a][b] + ( b <= a ? BD[a][T-1-a+b] : b == a+1 ? 0 : BD[a+1][b-a-2] )
//             one warp per row, lanes along b (coalesced); online max / sum in fp32; needs BD row a+1, hence 127 owned
//             rows per 128-row tile
//   phase 3   ctx = softmax(s / sqrt(d_k)) . V     probabilities are split into (hi, lo') pairs and written as the
//             swizzled K-major A operand, 64 keys at a time; V tiles arrive by TMA exactly as they lie in memory
//             ([key][d], i.e. the MN-major B operand); accumulators main / correction in TMEM
//   epilogue  ctx -> (hi, lo') pair of the output projection's operand buffer
//
// No (B, H, T, T) tensor exists in HBM beyond the per-CTA scratch, no length limit below the positional table, and no
// other attention kernel behind it: an unsupported head size is an error at engine creation.
//
// The same kernel runs plain scaled-dot-product attention (the Matcha decoder's transformer blocks, diffusers
// `Attention` as configured by jatts/modules/matchatts/transformer.py:216-232): nph = 1 skips phase 2 and the
// positional term of the index map; the projection matrix is then [q | k | v].
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-9 = workers
// (score epilogue: the four whose warp-id % 4 covers the TMEM lane quadrants; softmax and the P operand: all eight).
#include "../../include/jatts_b200.h"
#include "conv_gemm.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace jb {
extern long long* g_trace_ptr;   // conv_gemm_tc2.cu
namespace {

constexpr int kAttThreads = 320;
constexpr int kOwnRows = 127;                 // rows a tile stores (row 127 only supplies BD[a+1])
constexpr int kTile16 = 128 * 128;            // one [128 rows x 64 cols] 16-bit tile, 128-byte swizzled: 16 KB
constexpr int kStageB = 2 * kTile16;          // hi + lo
constexpr int kMaxRing = 3;
constexpr int kMaxNC = 4;                     // d_k <= 256
// Operand region: 6 stages of 32 KB.  Phases 1 / 2: Q (one stage per 64-wide d_k chunk) followed by the K / position
// ring (3 stages for d_k <= 192, 2 for d_k = 256).  Phase 3 aliases it: the two P chunk buffers, then the two V stages.
constexpr int kRegion = 6 * kStageB;
constexpr int kVBlock = 64 * 128;             // [64 keys x 64 dims] 8 KB
constexpr int kStgPitch = 20;                // staging row pitch in words: 16 data + 4 pad (conflict-free 128-bit rows)
constexpr int kStgBytes = 32 * kStgPitch * 4;   // one worker warp's [32 rows x 16 words] transpose buffer
constexpr int kSmemBytes = kRegion + 1024 /*barriers*/ + 8 * kStgBytes + 1024 /*row statistics*/ + 1024 /*alignment*/;
static_assert(2 * kStageB + 2 * 2 * kMaxNC * kVBlock <= kRegion, "phase-3 buffers alias the phase-1/2 region");

struct AttParams {
  const int* seg_start;
  const int* seg_len;
  int n_head, dk, nc;
  int nph;           // 2: content + positional scores (legacy rel-pos attention); 1: content scores only (plain attention)
  int q_col[2];      // first column of (q + bias_u), (q + bias_v) in the projection matrix; k_col / v_col likewise
  int k_col, v_col;
  int y_off;         // byte offset of the K / position ring inside the operand region (= size of the Q stages)
  int ring;          // ring stages (2 or 3)
  int v_off;         // byte offset of the two V stages (behind the two P chunk buffers)
  int v_stage;       // bytes per V stage: 2 * nc * kVBlock
  int tiles_per_utt, total_tiles;
  int tp;            // scratch row pitch in floats (multiple of 128)
  float* scratch;    // [gridDim.x][2][128][tp]
  bf16* out_hi;
  bf16* out_lo;
  int out_ld;
  float scale;       // 1 / sqrt(d_k)
  long long* trace;  // debug: clock64 stamps of CTA 0 (jatts_debug_set_trace), else null
};

#ifdef JB_ENABLE_TRACE
#define ATR(slot, idx)                                                                                       \
  do {                                                                                                       \
    if (P.trace && blockIdx.x == 0 && (idx) < 8) P.trace[(slot) * 8 + (idx)] = clock64();                    \
  } while (0)
#else
#define ATR(slot, idx) do { } while (0)
#endif

struct TileInfo {
  int h, a0, T, seg0;
  bool valid;
};
__device__ __forceinline__ TileInfo decode_tile(const AttParams& P, int tile) {
  TileInfo t;
  const int rt = tile % P.tiles_per_utt;
  const int bh = tile / P.tiles_per_utt;
  const int b = bh / P.n_head;
  t.h = bh - b * P.n_head;
  t.a0 = rt * kOwnRows;
  t.T = __ldg(P.seg_len + b);
  t.seg0 = __ldg(P.seg_start + b);
  t.valid = t.a0 < t.T;
  return t;
}

// fp16 operands (format 0), fp32 accumulate; B either K-major or MN-major (bit 16)
__host__ __device__ constexpr uint32_t att_idesc(int n, bool b_mn_major) {
  return (1u << 4) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {   // 2^x, 2 ulp; 2^-inf = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two values -> packed fp16 (hi, hi) and (lo', lo') words of the split operand pair (common.cuh::split_op16 without
// the saturation: probabilities and convex combinations of fp16-representable values cannot overflow)
__device__ __forceinline__ void split_pair16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a - hf.x) * kSplitScale, (b - hf.y) * kSplitScale);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kAttThreads, 1)
relpos_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                           const __grid_constant__ CUtensorMap tm_v_hi, const __grid_constant__ CUtensorMap tm_v_lo,
                           const __grid_constant__ CUtensorMap tm_p_hi, const __grid_constant__ CUtensorMap tm_p_lo,
                           const __grid_constant__ AttParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* X = smem;
  uint8_t* Y = smem + P.y_off;
  uint8_t* V = smem + P.v_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRegion);
  uint64_t* q_full = bars + 0;
  uint64_t* q_free = bars + 1;
  uint64_t* s2_done = bars + 2;
  uint64_t* ctx_full = bars + 3;
  uint64_t* tile_free = bars + 4;
  uint64_t* b_full = bars + 5;     // [3]
  uint64_t* b_empty = bars + 8;    // [3]
  uint64_t* s_full = bars + 11;    // [2]
  uint64_t* s_empty = bars + 13;   // [2]
  uint64_t* p_full = bars + 15;    // [2]
  uint64_t* p_empty = bars + 17;   // [2]
  uint64_t* v_full = bars + 19;    // [2]
  uint64_t* v_empty = bars + 21;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  uint8_t* stg_base = smem + kRegion + 1024;   // [8 worker warps][kStgBytes]
  float* stat_m = reinterpret_cast<float*>(stg_base + 8 * kStgBytes);   // [128] row maximum (log2 domain)
  float* stat_l = stat_m + 128;                                         // [128] 1 / row sum (0 = row not owned)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo);
    tma_prefetch_desc(&tm_v_hi); tma_prefetch_desc(&tm_v_lo);
    tma_prefetch_desc(&tm_p_hi); tma_prefetch_desc(&tm_p_lo);
    mbar_init(q_full, 1); mbar_init(q_free, 1); mbar_init(s2_done, 1); mbar_init(ctx_full, 1);
    mbar_init(tile_free, 8);
    for (int i = 0; i < kMaxRing; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
      mbar_init(&p_full[i], 8); mbar_init(&p_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();   // the projection GEMM's outputs are visible from here

  const uint32_t ring = static_cast<uint32_t>(P.ring);
  if (warp == 0) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      uint32_t nb = 0, nv = 0, it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileInfo ti = decode_tile(P, tile);
        if (!ti.valid) continue;
        const int nkt = (ti.T + 127) >> 7, nkc = (ti.T + 63) >> 6;
        const int hcol = ti.h * P.dk;
        if (it > 0) mbar_wait(ctx_full, (it - 1) & 1);   // every MMA of the previous tile has read its operands
        ATR(0, it);
        auto load_q = [&](int which) {
          mbar_expect_tx(q_full, static_cast<uint32_t>(P.nc) * kStageB);
          for (int c = 0; c < P.nc; ++c) {
            tma_load_2d(&tm_x_hi, q_full, X + c * kStageB, P.q_col[which] + hcol + c * 64, ti.seg0 + ti.a0);
            tma_load_2d(&tm_x_lo, q_full, X + c * kStageB + kTile16, P.q_col[which] + hcol + c * 64, ti.seg0 + ti.a0);
          }
        };
        auto load_b = [&](const CUtensorMap* mh, const CUtensorMap* ml, int col0, int row) {
          for (int c = 0; c < P.nc; ++c) {
            const uint32_t st = nb % ring;
            mbar_wait(&b_empty[st], ((nb / ring) & 1) ^ 1);
            mbar_expect_tx(&b_full[st], kStageB);
            tma_load_2d(mh, &b_full[st], Y + st * kStageB, col0 + c * 64, row);
            tma_load_2d(ml, &b_full[st], Y + st * kStageB + kTile16, col0 + c * 64, row);
            ++nb;
          }
        };
        load_q(0);                                                                         // q + bias_u
        for (int j = 0; j < nkt; ++j) load_b(&tm_x_hi, &tm_x_lo, P.k_col + hcol, ti.seg0 + j * 128);   // keys
        ATR(1, it);
        if (P.nph == 2) {
          mbar_wait(q_free, it & 1);                                                       // phase-1 MMAs done with Q
          ATR(2, it);
          load_q(1);                                                                       // q + bias_v
          for (int j = 0; j < nkt; ++j) load_b(&tm_p_hi, &tm_p_lo, hcol, j * 128);         // positions 0 .. T-1
        }
        mbar_wait(s2_done, it & 1);                                                        // ring and Q regions are free
        ATR(3, it);
        for (int kc = 0; kc < nkc; ++kc) {
          const uint32_t st = nv & 1;
          mbar_wait(&v_empty[st], ((nv >> 1) & 1) ^ 1);
          mbar_expect_tx(&v_full[st], static_cast<uint32_t>(P.v_stage));
          for (int c = 0; c < P.nc; ++c) {
            tma_load_2d(&tm_v_hi, &v_full[st], V + st * P.v_stage + c * kVBlock, P.v_col + hcol + c * 64, ti.seg0 + kc * 64);
            tma_load_2d(&tm_v_lo, &v_full[st], V + st * P.v_stage + (P.nc + c) * kVBlock, P.v_col + hcol + c * 64, ti.seg0 + kc * 64);
          }
          ++nv;
        }
        ATR(4, it);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc_c = att_idesc(P.dk, true);
      constexpr uint32_t desc_hi = static_cast<uint32_t>(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, v1, SWIZZLE_128B
      constexpr uint32_t k_lo0 = 1u << 16;                                      // K-major: LBO unused
      constexpr uint32_t v_lo0 = static_cast<uint32_t>(kVBlock >> 4) << 16;     // MN-major: LBO = next 64-wide block of d
      const uint32_t x_addr = smem_u32(X), y_addr = smem_u32(Y), v_addr = smem_u32(V);
      uint32_t nb = 0, nv = 0, sb = 0, pc = 0, qf = 0, it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileInfo ti = decode_tile(P, tile);
        if (!ti.valid) continue;
        const int nkt = (ti.T + 127) >> 7, nkc = (ti.T + 63) >> 6;
        if (it > 0) mbar_wait(tile_free, (it - 1) & 1);   // the context accumulators of the previous tile were read
        tc_fence_after();
        ATR(8, it);
        for (int ph = 0; ph < P.nph; ++ph) {
          mbar_wait(q_full, qf & 1);
          ++qf;
          tc_fence_after();
          ATR(9 + 2 * ph, it);
          for (int j = 0; j < nkt; ++j) {
            const uint32_t buf = sb & 1;
            mbar_wait(&s_empty[buf], ((sb >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_main = tmem_base + buf * 256u, d_corr = d_main + 128u;
            // the last key / position tile only computes the 16-column groups that hold real keys
            const uint32_t idesc_j = att_idesc(min(128, round_up(ti.T - j * 128, 16)), false);
            for (int c = 0; c < P.nc; ++c) {
              const uint32_t st = nb % ring;
              mbar_wait(&b_full[st], (nb / ring) & 1);
              tc_fence_after();
              const uint32_t a0 = x_addr + c * kStageB, b0 = y_addr + st * kStageB;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t ah = k_lo0 + ((a0 + k * 32) >> 4), al = k_lo0 + ((a0 + kTile16 + k * 32) >> 4);
                const uint32_t bh = k_lo0 + ((b0 + k * 32) >> 4), bl = k_lo0 + ((b0 + kTile16 + k * 32) >> 4);
                const uint32_t acc = (c | k) != 0 ? 1u : 0u;
                tc_mma_bf16_lohi(d_main, ah, desc_hi, bh, desc_hi, idesc_j, acc);
                tc_mma_bf16_lohi(d_corr, ah, desc_hi, bl, desc_hi, idesc_j, acc);
                tc_mma_bf16_lohi(d_corr, al, desc_hi, bh, desc_hi, idesc_j, 1u);
              }
              tc_commit(&b_empty[st]);
              ++nb;
            }
            tc_commit(&s_full[buf]);
            ++sb;
          }
          tc_commit(ph == P.nph - 1 ? s2_done : q_free);
          ATR(10 + 2 * ph, it);
        }
        for (int kc = 0; kc < nkc; ++kc) {
          const uint32_t pb = pc & 1, vs = nv & 1;
          mbar_wait(&p_full[pb], (pc >> 1) & 1);
          mbar_wait(&v_full[vs], (nv >> 1) & 1);
          tc_fence_after();
          if (kc == 0) ATR(13, it);
          const uint32_t a0 = x_addr + pb * kStageB, b0 = v_addr + vs * static_cast<uint32_t>(P.v_stage);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 16 keys per MMA: 32 B along the A rows, two 8-key groups (2 KB) of V
            const uint32_t ah = k_lo0 + ((a0 + k * 32) >> 4), al = k_lo0 + ((a0 + kTile16 + k * 32) >> 4);
            const uint32_t bh = v_lo0 + ((b0 + k * 2048) >> 4), bl = v_lo0 + ((b0 + P.nc * kVBlock + k * 2048) >> 4);
            const uint32_t acc = (kc | k) != 0 ? 1u : 0u;
            tc_mma_bf16_lohi(tmem_base, ah, desc_hi, bh, desc_hi, idesc_c, acc);
            tc_mma_bf16_lohi(tmem_base + 256u, ah, desc_hi, bl, desc_hi, idesc_c, acc);
            tc_mma_bf16_lohi(tmem_base + 256u, al, desc_hi, bh, desc_hi, idesc_c, 1u);
          }
          tc_commit(&p_empty[pb]);
          tc_commit(&v_empty[vs]);
          ++pc;
          ++nv;
        }
        tc_commit(ctx_full);
        ATR(14, it);
        ++it;
      }
    }
  } else {
    // ============================ workers ============================
    const int w = warp - 2;            // 0..7: rows w*16 .. w*16+15 in the softmax / P phases
    const int quad = warp & 3;         // TMEM lane quadrant this warp may read (warps 2..5 and 6..9 both cover 2,3,0,1)
    const int hsel = w >> 2;           // which half of an accumulator's columns this warp moves in the epilogues
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int tp = P.tp;
    float* scr = P.scratch + static_cast<size_t>(blockIdx.x) * 2 * 128 * tp;
    float* AC = scr;
    const float* BD = scr + static_cast<size_t>(128) * tp;
    const uint32_t x_addr = smem_u32(X);
    constexpr float kInvSplit = 1.0f / kSplitScale;
    // staging tile of this warp: a thread owns one accumulator row (32 rows x 16 words + pad); the rows leave it
    // 8 at a time, 4 lanes x 16 B per row, so that global stores are row-contiguous instead of 32 scattered pieces
    const uint32_t stg = smem_u32(stg_base + w * kStgBytes);
    const uint32_t stg_w = stg + static_cast<uint32_t>(lane * kStgPitch * 4);
    const uint32_t stg_r = stg + static_cast<uint32_t>(((lane >> 2) * kStgPitch + (lane & 3) * 4) * 4);
    const float sc2 = P.scale * 1.4426950408889634f;   // scores are kept in the log2 domain: p = 2^(s - max)
    const bool pos = P.nph == 2;
    uint32_t se = 0, pc = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileInfo ti = decode_tile(P, tile);
      if (!ti.valid) continue;
      const int T = ti.T, a0 = ti.a0;
      const int nkt = (T + 127) >> 7, nkc = (T + 63) >> 6;
      const bool tr0 = warp == 2 && lane == 0, tr4 = warp == 6 && lane == 0;
      if (tr0) ATR(16, it);
      // ---- score epilogue: TMEM (main + corr * 2^-11) -> scratch; every warp moves 4 of the 8 16-column groups
      for (int ph = 0; ph < P.nph; ++ph) {
        float* dst = scr + (static_cast<size_t>(ph) * 128 + quad * 32 + (lane >> 2)) * tp + hsel * 64 + (lane & 3) * 4;
        for (int j = 0; j < nkt; ++j) {
          const uint32_t buf = se & 1;
          mbar_wait(&s_full[buf], (se >> 1) & 1);
          tc_fence_after();
          if (tr0 && ph == 0 && j == 0) ATR(17, it);
          const int ng = min(4, ((min(128, T - j * 128) + 15) >> 4) - hsel * 4);   // 16-column groups with real keys
          if (ng <= 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[buf]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g >= ng) break;
            uint32_t m[16], c[16];
            tmem_ld16(lane_base + buf * 256u + static_cast<uint32_t>(hsel * 64 + g * 16), m);
            tmem_ld16(lane_base + buf * 256u + 128u + static_cast<uint32_t>(hsel * 64 + g * 16), c);
            tmem_ld_wait();
            if (g == ng - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_empty[buf]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 v;
              v.x = __float_as_uint(__uint_as_float(m[4 * i]) + __uint_as_float(c[4 * i]) * kInvSplit);
              v.y = __float_as_uint(__uint_as_float(m[4 * i + 1]) + __uint_as_float(c[4 * i + 1]) * kInvSplit);
              v.z = __float_as_uint(__uint_as_float(m[4 * i + 2]) + __uint_as_float(c[4 * i + 2]) * kInvSplit);
              v.w = __float_as_uint(__uint_as_float(m[4 * i + 3]) + __uint_as_float(c[4 * i + 3]) * kInvSplit);
              sts128(stg_w + static_cast<uint32_t>(i * 16), v);
            }
            __syncwarp();
            float* o = dst + j * 128 + g * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // rows 8i .. 8i+7 of the warp's 32
              const uint4 v = lds128(stg_r + static_cast<uint32_t>(i * 8 * kStgPitch * 4));
              *reinterpret_cast<uint4*>(o + static_cast<size_t>(i * 8) * tp) = v;
            }
            __syncwarp();
          }
          ++se;
        }
      }
      if (tr0) ATR(18, it);
      if (tr4) ATR(24, it);
      __threadfence_block();
      named_bar_sync(1, 256);
      if (tr0) ATR(19, it);
      if (tr4) ATR(25, it);
      // ---- softmax statistics: 16 rows per warp, two rows x 8 column blocks of loads in flight per lane; the
      //      combined score (log2 domain) replaces AC in place; max / sum are online across 256-column chunks only
      // (the row-pair loop is ROLLED and its results go through shared memory: unrolled 8x it was 4 k instructions, and
      //  the warp-specialised kernels are instruction-cache sensitive)
#pragma unroll 1
      for (int rp = 0; rp < 8; ++rp) {
        float M[2], L[2];
        float* pac[2];
        const float* plo[2];
        const float* phi[2];
        int aa[2];
        bool rok[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int r = w * 16 + rp * 2 + q, a = a0 + r;
          M[q] = -INFINITY;
          L[q] = 0.f;
          aa[q] = a;
          rok[q] = r < kOwnRows && a < T;
          pac[q] = AC + static_cast<size_t>(r) * tp;
          plo[q] = BD + static_cast<size_t>(r) * tp + (T - 1 - a);        // b <= a   : BD[a][T-1-a+b]
          phi[q] = BD + static_cast<size_t>(r + 1) * tp - (a + 2);        // b >= a+2 : BD[a+1][b-a-2]
          if (!rok[q]) {   // the loads below are unconditional: a row outside the tile reads (and discards) row 0
            pac[q] = AC;
            plo[q] = BD;
            phi[q] = BD;
          }
        }
        if (rok[0]) {   // warp uniform (row 2q+1 may still be outside: its loads are predicated off)
          for (int c0 = 0; c0 < T; c0 += 256) {
            // every load is unconditional and in bounds of the scratch (clamped column; the rel-shift pointers of a
            // row outside the utterance still point into the scratch), so all 32 are in flight before the first use
            float xv[2][8], yv[2][8], sv[2][8];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int b = min(c0 + 32 * i + lane, T - 1);
                const float* pb = b <= aa[q] ? plo[q] : phi[q];
                xv[q][i] = pac[q][b];
                yv[q][i] = pos ? pb[b] : 0.f;
              }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int b = c0 + 32 * i + lane;
                const float y = b != aa[q] + 1 ? yv[q][i] : 0.f;
                sv[q][i] = (rok[q] && b < T) ? (xv[q][i] + y) * sc2 : -INFINITY;
              }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float cm = -INFINITY;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int b = c0 + 32 * i + lane;
                if (sv[q][i] > -INFINITY) pac[q][b] = sv[q][i];
                cm = fmaxf(cm, sv[q][i]);
              }
              const float mn = fmaxf(M[q], cm);
              if (mn > -INFINITY) {
                float acc = L[q] * ex2_approx(M[q] - mn);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc += ex2_approx(sv[q][i] - mn);
                L[q] = acc;
                M[q] = mn;
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float Mw = warp_max(M[q]);
          const float Lw = warp_sum(M[q] > -INFINITY ? L[q] * ex2_approx(M[q] - Mw) : 0.f);
          if (lane == 0) {
            stat_m[w * 16 + rp * 2 + q] = Lw > 0.f ? Mw : 0.f;          // rows this tile does not own have no scores:
            stat_l[w * 16 + rp * 2 + q] = Lw > 0.f ? 1.0f / Lw : 0.f;   // probability 0 (2^(-inf - 0) * 0)
          }
        }
      }
      __syncwarp();   // the in-place scores and the statistics of a row are read below by other lanes of this warp
      float mrow[16], lrow[16];
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        mrow[rr] = stat_m[w * 16 + rr];
        lrow[rr] = stat_l[w * 16 + rr];
      }
      if (tr0) ATR(20, it);
      if (tr4) ATR(26, it);
      // ---- phase 3: probabilities of 64 keys -> (hi, lo') A-operand chunk (lane = 2 adjacent keys of 16 rows);
      //      the scores of the next chunk are fetched before this one is converted
      {
        float2 cur[16];
        auto fetch = [&](float2 (&d)[16], int kc) {
          const int key = kc * 64 + 2 * lane;   // < tp: unconditional loads, masked afterwards
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) d[rr] = *reinterpret_cast<const float2*>(AC + static_cast<size_t>(w * 16 + rr) * tp + key);
#pragma unroll
          for (int rr = 0; rr < 16; ++rr)
            if (!(lrow[rr] > 0.f && key < T)) d[rr] = make_float2(-INFINITY, -INFINITY);
        };
        fetch(cur, 0);
        for (int kc = 0; kc < nkc; ++kc) {
          float2 nxt[16];
          if (kc + 1 < nkc) fetch(nxt, kc + 1);
          const uint32_t pb = pc & 1;
          mbar_wait(&p_empty[pb], ((pc >> 1) & 1) ^ 1);
          const bool second = kc * 64 + 2 * lane + 1 < T;
          const uint32_t dst0 = x_addr + pb * kStageB + static_cast<uint32_t>((lane & 3) * 4);
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = w * 16 + rr;
            const float p0 = ex2_approx(cur[rr].x - mrow[rr]) * lrow[rr];                      // 2^-inf = 0 outside
            const float p1 = second ? ex2_approx(cur[rr].y - mrow[rr]) * lrow[rr] : 0.f;
            uint32_t hw, lw;
            split_pair16(p0, p1, hw, lw);
            const uint32_t a = dst0 + static_cast<uint32_t>(r * 128) + (static_cast<uint32_t>((lane >> 2) ^ (r & 7)) << 4);
            sts32(a, hw);
            sts32(a + kTile16, lw);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[pb]);
          ++pc;
          if (kc + 1 < nkc) {
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) cur[rr] = nxt[rr];
          }
        }
      }
      if (tr0) ATR(21, it);
      if (tr4) ATR(27, it);
      // ---- context epilogue: TMEM -> (hi, lo') rows of the output projection's operand; staging row of a thread =
      //      [16 hi values (32 B) | 16 lo values (32 B)], every warp moves half of the 16-column groups
      {
        mbar_wait(ctx_full, it & 1);
        tc_fence_after();
        if (tr0) ATR(22, it);
        const int n16 = P.dk >> 4, g0 = hsel * (n16 >> 1), g1 = hsel ? n16 : (n16 >> 1);
        bf16* const obase = (lane & 2) ? P.out_lo : P.out_hi;
        for (int g = g0; g < g1; ++g) {
          uint32_t m[16], c[16];
          tmem_ld16(lane_base + static_cast<uint32_t>(g * 16), m);
          tmem_ld16(lane_base + 256u + static_cast<uint32_t>(g * 16), c);
          tmem_ld_wait();
          if (g == g1 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tile_free);
          }
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            split_pair16(__uint_as_float(m[2 * i]) + __uint_as_float(c[2 * i]) * kInvSplit,
                         __uint_as_float(m[2 * i + 1]) + __uint_as_float(c[2 * i + 1]) * kInvSplit, hw[i], lw[i]);
          sts128(stg_w, make_uint4(hw[0], hw[1], hw[2], hw[3]));
          sts128(stg_w + 16, make_uint4(hw[4], hw[5], hw[6], hw[7]));
          sts128(stg_w + 32, make_uint4(lw[0], lw[1], lw[2], lw[3]));
          sts128(stg_w + 48, make_uint4(lw[4], lw[5], lw[6], lw[7]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {   // rows 8i .. 8i+7; lanes 0,1 of a row carry hi, lanes 2,3 lo
            const int r = quad * 32 + i * 8 + (lane >> 2), a = a0 + r;
            const uint4 v = lds128(stg_r + static_cast<uint32_t>(i * 8 * kStgPitch * 4));
            if (r < kOwnRows && a < T)
              *reinterpret_cast<uint4*>(obase + static_cast<size_t>(ti.seg0 + a) * P.out_ld + ti.h * P.dk + g * 16 + (lane & 1) * 8) = v;
          }
          __syncwarp();
        }
      }
      if (tr0) ATR(23, it);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

bool relpos_attention_supported(int n_head, int d_model) {
  if (n_head <= 0 || d_model % n_head != 0) return false;
  const int dk = d_model / n_head;
  return dk % 64 == 0 && dk <= 64 * kMaxNC;
}

size_t relpos_attention_scratch_bytes(int max_len, int nseg, int n_head) {
  if (max_len <= 0 || nseg <= 0) return 0;
  const long long tiles = static_cast<long long>(nseg) * n_head * ceil_div(max_len, kOwnRows);
  const long long grid = tiles < num_sms() ? tiles : num_sms();
  return static_cast<size_t>(grid) * 2 * 128 * round_up(max_len, 128) * sizeof(float);
}

// x: [x_rows, x_ld] projection matrix (operand pairs); q_col0 / q_col1 / k_col / v_col: first column of each block.
// pos_hi == null: plain softmax(q k^T / sqrt(d_k)) v (q_col1 unused).
static int attention_launch(const bf16* x_hi, const bf16* x_lo, long long x_rows, int x_ld, int q_col0, int q_col1, int k_col,
                            int v_col, const bf16* pos_hi, const bf16* pos_lo, int pos_rows, int n_head, int d_model,
                            RowLayout L, int max_len, float* scratch, size_t scratch_bytes, bf16* out_hi, bf16* out_lo,
                            int out_ld, cudaStream_t s) {
  JB_REQUIRE(relpos_attention_supported(n_head, d_model), JATTS_E_UNSUPPORTED,
             "attention: d_k must be 64, 128, 192 or 256 (tcgen05 kernel; there is no other attention path)");
  JB_REQUIRE(x_hi && x_lo && out_hi && out_lo && (pos_hi != nullptr) == (pos_lo != nullptr), JATTS_E_INVALID, "attention: null operand");
  const bool pos = pos_hi != nullptr;
  JB_REQUIRE(!pos || max_len <= pos_rows, JATTS_E_UNSUPPORTED, "attention: utterance longer than the positional table");
  JB_REQUIRE(out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0,
             JATTS_E_INVALID, "attention: output rows must be 16-byte aligned");
  JB_REQUIRE(x_ld % 8 == 0 && q_col0 % 8 == 0 && q_col1 % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0, JATTS_E_INVALID,
             "attention: operand blocks must start on 16-byte boundaries");
  if (L.nseg == 0 || max_len == 0) return 0;
  JB_REQUIRE(scratch != nullptr && scratch_bytes >= relpos_attention_scratch_bytes(max_len, L.nseg, n_head), JATTS_E_INVALID,
             "attention: scratch buffer too small");
  AttParams P{};
  P.seg_start = L.seg_start;
  P.seg_len = L.seg_len;
  P.n_head = n_head;
  P.dk = d_model / n_head;
  P.nc = P.dk / 64;
  P.nph = pos ? 2 : 1;
  P.q_col[0] = q_col0; P.q_col[1] = q_col1; P.k_col = k_col; P.v_col = v_col;
  P.y_off = (P.nc <= 3 ? 3 : 4) * kStageB;
  P.ring = (kRegion - P.y_off) / kStageB;
  P.v_stage = 2 * P.nc * kVBlock;
  P.v_off = P.nc <= 3 ? P.y_off : 2 * kStageB;   // d_k = 256: the V stages start right behind the two P buffers
  JB_REQUIRE(P.ring >= 2 && P.ring <= kMaxRing && P.v_off >= 2 * kStageB && P.v_off + 2 * P.v_stage <= kRegion, JATTS_E_UNSUPPORTED,
             "attention: shared-memory plan");
  P.tiles_per_utt = ceil_div(max_len, kOwnRows);
  P.total_tiles = L.nseg * n_head * P.tiles_per_utt;
  P.tp = round_up(max_len, 128);
  P.scratch = scratch;
  P.out_hi = out_hi;
  P.out_lo = out_lo;
  P.out_ld = out_ld;
  P.scale = 1.0f / sqrtf(static_cast<float>(P.dk));
  P.trace = g_trace_ptr;
  CUtensorMap mxh, mxl, mvh, mvl, mph, mpl;
  JB_PROPAGATE(make_tmap(&mxh, x_hi, x_rows, x_ld, x_ld, 128));
  JB_PROPAGATE(make_tmap(&mxl, x_lo, x_rows, x_ld, x_ld, 128));
  JB_PROPAGATE(make_tmap(&mvh, x_hi, x_rows, x_ld, x_ld, 64));
  JB_PROPAGATE(make_tmap(&mvl, x_lo, x_rows, x_ld, x_ld, 64));
  if (pos) {
    JB_PROPAGATE(make_tmap(&mph, pos_hi, pos_rows, d_model, d_model, 128));
    JB_PROPAGATE(make_tmap(&mpl, pos_lo, pos_rows, d_model, d_model, 128));
  } else {   // never dereferenced (nph == 1); a valid descriptor keeps the prefetch harmless
    mph = mxh;
    mpl = mxl;
  }
  JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(relpos_attention_tc_kernel), kSmemBytes));
  const int grid = P.total_tiles < num_sms() ? P.total_tiles : num_sms();
  ProfileScope prof(s, PROF_ATTENTION);
  JB_CUDA_OK(launch_tc(relpos_attention_tc_kernel, grid, kAttThreads, kSmemBytes, s, 1, mxh, mxl, mvh, mvl, mph, mpl, P));
  JB_KERNEL_OK();
  return 0;
}

int relpos_attention(const bf16* x_hi, const bf16* x_lo, long long x_rows, const bf16* pos_hi, const bf16* pos_lo,
                     int pos_rows, int n_head, int d_model, RowLayout L, int max_len, float* scratch,
                     size_t scratch_bytes, bf16* out_hi, bf16* out_lo, int out_ld, cudaStream_t s) {
  JB_REQUIRE(pos_hi && pos_lo, JATTS_E_INVALID, "attention: null positional table");
  return attention_launch(x_hi, x_lo, x_rows, 4 * d_model, 0, d_model, 2 * d_model, 3 * d_model, pos_hi, pos_lo, pos_rows,
                          n_head, d_model, L, max_len, scratch, scratch_bytes, out_hi, out_lo, out_ld, s);
}

int plain_attention(const bf16* x_hi, const bf16* x_lo, long long x_rows, int n_head, int d_model, RowLayout L, int max_len,
                    float* scratch, size_t scratch_bytes, bf16* out_hi, bf16* out_lo, int out_ld, cudaStream_t s) {
  return attention_launch(x_hi, x_lo, x_rows, 3 * d_model, 0, 0, d_model, 2 * d_model, nullptr, nullptr, 0, n_head, d_model, L,
                          max_len, scratch, scratch_bytes, out_hi, out_lo, out_ld, s);
}

}  // namespace jb
