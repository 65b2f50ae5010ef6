// Bandwidth-bound / small kernels of the synthesis path (CUDA cores, fp32 arithmetic).
// Every kernel works on the packed-with-gaps row layout (common.cuh): gap rows are never written.
#include <cstdlib>
#include <type_traits>

#include "conv_gemm.cuh"
#include "kernels.cuh"

namespace jb {

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
__global__ void fill_layout_kernel(const int* __restrict__ seg_start, const int* __restrict__ seg_len, int nseg,
                                   int n_rows, uint8_t* __restrict__ mask, int* __restrict__ seg) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int lo = 0, hi = nseg - 1, b = -1;
  while (lo <= hi) {  // last segment with start <= r
    const int mid = (lo + hi) >> 1;
    if (seg_start[mid] <= r) { b = mid; lo = mid + 1; } else { hi = mid - 1; }
  }
  const bool v = b >= 0 && r < seg_start[b] + seg_len[b];
  mask[r] = v ? 1 : 0;
  seg[r] = v ? b : -1;
}
int fill_layout(const int* seg_start, const int* seg_len, int nseg, int n_rows, uint8_t* frame_mask, int* frame_seg,
                cudaStream_t s) {
  if (n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(fill_layout_kernel, dim3(ceil_div(n_rows, 256)), dim3(256), 0, s, seg_start, seg_len, nseg, n_rows, frame_mask, frame_seg));
  JB_KERNEL_OK();
  return 0;
}

// grid = (gap rows per utterance, utterances, buffers): only the rows that must be zero are touched; up to 16 buffers in
// one launch (the engines zero ~14 operand buffers per layout)
struct GapBufs {
  uint4* p[16];
  int row_vec[16];
};
__global__ void zero_gap_rows_multi_kernel(GapBufs B, const int* __restrict__ seg_start, const int* __restrict__ seg_len) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, k = blockIdx.z;
  const long long r = static_cast<long long>(seg_start[b]) + seg_len[b] + blockIdx.x;
  const int rv = B.row_vec[k];
  uint4* p = B.p[k] + r * rv;
  for (int i = threadIdx.x; i < rv; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}
int zero_gap_rows_multi(void* const* bufs, const int* row_bytes, int n, RowLayout L, cudaStream_t s) {
  JB_REQUIRE(n >= 0 && n <= 16, -2, "zero_gap_rows_multi: at most 16 buffers");
  if (L.nseg == 0 || n == 0) return 0;
  GapBufs B{};
  for (int i = 0; i < n; ++i) {
    JB_REQUIRE(row_bytes[i] % 16 == 0, -2, "zero_gap_rows_multi: row_bytes % 16");
    B.p[i] = static_cast<uint4*>(bufs[i]);
    B.row_vec[i] = row_bytes[i] / 16;
  }
  JB_CUDA_OK(launch_pdl(zero_gap_rows_multi_kernel, dim3(dim3(kGapRows, L.nseg, n)), dim3(64), 0, s, B, L.seg_start, L.seg_len));
  JB_KERNEL_OK();
  return 0;
}
// ------------------------------------------------------------------------------------------------
// embedding
// ------------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const long long* __restrict__ tokens, const int* __restrict__ tok_off,
                                    const float* __restrict__ emb, int vocab, int d, float scale, RowLayout L,
                                    float* __restrict__ x) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  long long tok = tokens[tok_off[b] + (r - L.seg_start[b])];
  if (tok < 0 || tok >= vocab) tok = 0;  // host validates ids; never index out of the table
  const float4* e = reinterpret_cast<const float4*>(emb + tok * d);
  float4* o = reinterpret_cast<float4*>(x + static_cast<long long>(r) * d);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
    float4 v = e[i];
    o[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}
int embed_tokens(const long long* tokens, const int* tok_off, const float* emb, int vocab, int d, float scale,
                 RowLayout L, float* x, cudaStream_t s) {
  JB_REQUIRE(d % 4 == 0, -2, "embed: d % 4");
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(embed_tokens_kernel, dim3(L.n_rows), dim3(96), 0, s, tokens, tok_off, emb, vocab, d, scale, L, x));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (one warp per row, two-pass in registers)
// ------------------------------------------------------------------------------------------------
static constexpr int LN_MAXV = 4;  // float4 per lane -> C <= 512

template <bool DOT>
__global__ void layernorm_kernel(const float* __restrict__ x, int c, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, RowLayout L, float* __restrict__ y,
                                 bf16* __restrict__ hi, bf16* __restrict__ lo, int bf_ld,
                                 const float* __restrict__ w, float wb, float* __restrict__ dot_out) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.n_rows || !L.frame_mask[warp]) return;
  const long long r = warp;
  const int nv = c >> 2;  // float4 count
  float4 v[LN_MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      v[i] = reinterpret_cast<const float4*>(x + r * c)[j];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / c + eps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[j];
      const float4 bb = reinterpret_cast<const float4*>(beta)[j];
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bb.x;
      o.y = (v[i].y - mean) * rstd * g.y + bb.y;
      o.z = (v[i].z - mean) * rstd * g.z + bb.z;
      o.w = (v[i].w - mean) * rstd * g.w + bb.w;
      if (DOT) {
        const float4 ww = reinterpret_cast<const float4*>(w)[j];
        dot += (o.x * ww.x + o.y * ww.y) + (o.z * ww.z + o.w * ww.w);
      } else {
        if (y) reinterpret_cast<float4*>(y + r * c)[j] = o;
        if (hi) {
          bf16 h0, h1, h2, h3, l0, l1, l2, l3;
          split_op16(o.x, h0, l0); split_op16(o.y, h1, l1); split_op16(o.z, h2, l2); split_op16(o.w, h3, l3);
          __nv_bfloat162 p0 = __halves2bfloat162(h0, h1), p1 = __halves2bfloat162(h2, h3);
          uint2 ph = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
          reinterpret_cast<uint2*>(hi + r * bf_ld)[j] = ph;
          if (lo) {
            __nv_bfloat162 q0 = __halves2bfloat162(l0, l1), q1 = __halves2bfloat162(l2, l3);
            uint2 pl = make_uint2(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1));
            reinterpret_cast<uint2*>(lo + r * bf_ld)[j] = pl;
          }
        }
      }
    }
  }
  if (DOT) {
    dot = warp_sum(dot);
    if (lane == 0) dot_out[r] = dot + wb;
  }
}
int layernorm_rows(const float* x, int c, const float* gamma, const float* beta, float eps, RowLayout L, float* y,
                   bf16* hi, bf16* lo, int bf_ld, cudaStream_t s) {
  ProfileScope prof(s, PROF_LAYERNORM);
  JB_REQUIRE(c % 4 == 0 && c <= 128 * LN_MAXV, -2, "layernorm: C must be a multiple of 4 and <= 512");
  JB_REQUIRE(bf_ld % 4 == 0, -2, "layernorm: bf_ld % 4");
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(layernorm_kernel<false>, dim3(ceil_div(L.n_rows, 8)), dim3(256), 0, s, x, c, gamma, beta, eps, L, y, hi, lo, bf_ld, nullptr,
                                                                0.f, nullptr));
  JB_KERNEL_OK();
  return 0;
}
int ln_dot_rows(const float* x, int c, const float* gamma, const float* beta, float eps, const float* w, float b,
                RowLayout L, float* out, cudaStream_t s) {
  ProfileScope prof(s, PROF_LAYERNORM);
  JB_REQUIRE(c % 4 == 0 && c <= 128 * LN_MAXV, -2, "ln_dot: C must be a multiple of 4 and <= 512");
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(layernorm_kernel<true>, dim3(ceil_div(L.n_rows, 8)), dim3(256), 0, s, x, c, gamma, beta, eps, L, nullptr, nullptr, nullptr, 4,
                                                               w, b, out));
  JB_KERNEL_OK();
  return 0;
}

__global__ void split_rows_kernel(const float* __restrict__ x, int c, RowLayout L, bf16* __restrict__ hi,
                                  bf16* __restrict__ lo, int bf_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  if (!L.frame_mask[r]) return;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    bf16 h, l;
    split_op16(x[static_cast<long long>(r) * c + i], h, l);
    hi[static_cast<long long>(r) * bf_ld + i] = h;
    lo[static_cast<long long>(r) * bf_ld + i] = l;
  }
}
int split_rows(const float* x, int c, RowLayout L, bf16* hi, bf16* lo, int bf_ld, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(split_rows_kernel, dim3(L.n_rows), dim3(128), 0, s, x, c, L, hi, lo, bf_ld));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// depthwise conv + folded BatchNorm + Swish
// ------------------------------------------------------------------------------------------------
// One CTA = one strip of K*M consecutive frames of one utterance, all channels.  The strip plus its K-1 halo rows is
// staged in shared memory with coalesced 16-byte loads (zeros outside the utterance = the convolution's own padding);
// then thread c walks down the strip for channel c with the K-row window and the K taps in registers (the fully
// unrolled (p, j) loops make every window index static): one conflict-free shared-memory load per output instead of
// K global loads through L1 (31 at the decoder's kernel size), and every input row leaves L2 once per strip.
template <int K, int M>
__global__ void __launch_bounds__(256)
dwconv_swish_strip_kernel(const float* __restrict__ g, int c, const float* __restrict__ wT, const float* __restrict__ bias,
                          RowLayout L, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PAD = (K - 1) / 2, TR = K * M, ROWS = TR + K - 1;
  extern __shared__ __align__(16) float dw_tile[];   // [ROWS][cb]: this CTA's block of channels
  const int b = blockIdx.y;
  const int cb = blockDim.x;                          // channels per CTA (multiple of 32, divides c)
  const int c0 = blockIdx.z * cb;
  const int T = L.seg_len[b];
  const int r0 = blockIdx.x * TR;                     // first output frame of the strip (utterance-relative)
  if (r0 >= T) return;
  const long long base = L.seg_start[b];
  const int c4n = cb >> 2;
  const int n4 = ROWS * c4n;
  // the whole tile is put in flight at once with cp.async (16 bytes per request, zero fill outside the utterance: src-size 0);
  // with register-staged loads a CTA had 12 KB of its 70 KB tile in flight and spent most of its life waiting for it
  const uint32_t tile_s = static_cast<uint32_t>(__cvta_generic_to_shared(dw_tile));
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const int rr = i / c4n, q = i - rr * c4n;
    const int t = r0 - PAD + rr;
    const bool in = t >= 0 && t < T;
    const float* src = in ? g + (base + t) * c + c0 + q * 4 : g;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + static_cast<uint32_t>(i) * 16u), "l"(src),
                 "r"(in ? 16 : 0)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int ch = c0 + threadIdx.x;
  float w[K], win[K];
#pragma unroll
  for (int j = 0; j < K; ++j) w[j] = wT[j * c + ch];
  const float bv = bias[ch];
  const float* col = dw_tile + threadIdx.x;
#pragma unroll
  for (int j = 0; j < K - 1; ++j) win[j] = col[j * cb];
#pragma unroll 1
  for (int m = 0; m < M; ++m) {
#pragma unroll
    for (int p = 0; p < K; ++p) {
      const int o = m * K + p;                        // output row of the strip; its newest input is tile row o + K - 1
      win[(K - 1 + p) % K] = col[(o + K - 1) * cb];
      float acc = bv;
#pragma unroll
      for (int j = 0; j < K; ++j) acc = fmaf(win[(p + j) % K], w[j], acc);
      const float y = acc * (1.0f / (1.0f + __expf(-acc)));   // swish.py:13-18 (2 ulp exp: far inside the 1e-3 budget)
      const float yn = __shfl_down_sync(0xffffffffu, y, 1);  // even lanes store a channel pair (c is even)
      const int t = r0 + o;
      if (t < T && (ch & 1) == 0) {
        uint32_t hw, lw;
        split_pair16_sat(y, yn, hw, lw);
        *reinterpret_cast<uint32_t*>(out_hi + (base + t) * out_ld + ch) = hw;
        if (out_lo) *reinterpret_cast<uint32_t*>(out_lo + (base + t) * out_ld + ch) = lw;
      }
    }
  }
}

// generic tap count (no shipped configuration): one CTA per row, taps re-read through L1
__global__ void dwconv_swish_kernel(const float* __restrict__ g, int c, const float* __restrict__ wT,
                                    const float* __restrict__ bias, int k, RowLayout L, bf16* __restrict__ out_hi,
                                    bf16* __restrict__ out_lo, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  const int s0 = L.seg_start[b], s1 = s0 + L.seg_len[b];
  const int pad = (k - 1) / 2;
  for (int c4 = threadIdx.x; c4 < c / 4; c4 += blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(bias)[c4];
    for (int j = 0; j < k; ++j) {
      const int rr = r + j - pad;
      if (rr < s0 || rr >= s1) continue;
      const float4 x = reinterpret_cast<const float4*>(g + static_cast<long long>(rr) * c)[c4];
      const float4 w = reinterpret_cast<const float4*>(wT + j * c)[c4];
      acc.x += x.x * w.x; acc.y += x.y * w.y; acc.z += x.z * w.z; acc.w += x.w * w.w;
    }
    float o[4] = {acc.x, acc.y, acc.z, acc.w};
    bf16 hh[4], ll[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y = o[i] * (1.0f / (1.0f + expf(-o[i])));  // swish.py:13-18
      split_op16(y, hh[i], ll[i]);
    }
    __nv_bfloat162 p0 = __halves2bfloat162(hh[0], hh[1]), p1 = __halves2bfloat162(hh[2], hh[3]);
    reinterpret_cast<uint2*>(out_hi + static_cast<long long>(r) * out_ld)[c4] =
        make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
    if (out_lo) {
      __nv_bfloat162 q0 = __halves2bfloat162(ll[0], ll[1]), q1 = __halves2bfloat162(ll[2], ll[3]);
      reinterpret_cast<uint2*>(out_lo + static_cast<long long>(r) * out_ld)[c4] =
          make_uint2(*reinterpret_cast<uint32_t*>(&q0), *reinterpret_cast<uint32_t*>(&q1));
    }
  }
}
int dwconv_swish(const float* g, int c, const float* wT, const float* bias, int k, RowLayout L, int max_len, bf16* out_hi,
                 bf16* out_lo, int out_ld, cudaStream_t s) {
  ProfileScope prof(s, PROF_DWCONV);
  JB_REQUIRE(c % 4 == 0 && out_ld % 4 == 0 && (k & 1) == 1, -2, "dwconv: C % 4, odd k");
  if (L.n_rows == 0 || L.nseg == 0) return 0;
  static const bool generic_only = getenv("JATTS_B200_DWCONV_GENERIC") != nullptr;   // A/B switch for profiling
  if (!generic_only && c % 32 == 0 && max_len > 0 && (k == 31 || k == 7)) {
    // channels per CTA: the largest multiple of 32 <= 192 that divides c (3 CTAs of 70 KB per SM at c = 384, k = 31:
    // one CTA's tile load runs under the others' arithmetic)
    int cb = 0;
    for (int t = 192; t >= 32; t -= 32)
      if (c % t == 0) { cb = t; break; }
    JB_REQUIRE(cb > 0, -2, "dwconv: C must be a multiple of 32");
    if (k == 31) {
      constexpr int M = 2;
      const int smem = (31 * M + 30) * cb * 4;
      JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(dwconv_swish_strip_kernel<31, M>), smem));
      dim3 grid(ceil_div(max_len, 31 * M), L.nseg, c / cb);
      JB_CUDA_OK(launch_pdl(dwconv_swish_strip_kernel<31, M>, dim3(grid), dim3(cb), smem, s, g, c, wT, bias, L, out_hi, out_lo, out_ld));
    } else {
      constexpr int M = 9;
      const int smem = (7 * M + 6) * cb * 4;
      JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(dwconv_swish_strip_kernel<7, M>), smem));
      dim3 grid(ceil_div(max_len, 7 * M), L.nseg, c / cb);
      JB_CUDA_OK(launch_pdl(dwconv_swish_strip_kernel<7, M>, dim3(grid), dim3(cb), smem, s, g, c, wT, bias, L, out_hi, out_lo, out_ld));
    }
    JB_KERNEL_OK();
    return 0;
  }
  JB_CUDA_OK(launch_pdl(dwconv_swish_kernel, dim3(L.n_rows), dim3(96), 0, s, g, c, wT, bias, k, L, out_hi, out_lo, out_ld));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// speaker embedding integration ("add")
// ------------------------------------------------------------------------------------------------
__global__ void add_speaker_kernel(const float* __restrict__ spembs, int spk_dim, const float* __restrict__ w,
                                   const float* __restrict__ bvec, int d, RowLayout L, float* __restrict__ hs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sh[];  // [spk_dim] normalised embedding, then [d] projected
  float* e = sh;
  float* proj = sh + spk_dim;
  __shared__ float red[32];
  const int b = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < spk_dim; i += blockDim.x) {
    const float v = spembs[b * spk_dim + i];
    e[i] = v;
    ss += v * v;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize eps (fastspeech2.py:752)
  for (int o = threadIdx.x; o < d; o += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < spk_dim; ++i) acc += (e[i] * inv) * w[o * spk_dim + i];
    proj[o] = acc + bvec[o];
  }
  __syncthreads();
  const int s0 = L.seg_start[b], n = L.seg_len[b];
  for (long long i = threadIdx.x; i < static_cast<long long>(n) * d; i += blockDim.x)
    hs[static_cast<long long>(s0) * d + i] += proj[i % d];
}
int add_speaker(const float* spembs, int spk_dim, const float* w, const float* b, int d, RowLayout L, float* hs,
                cudaStream_t s) {
  if (L.nseg == 0) return 0;
  JB_CUDA_OK(launch_pdl(add_speaker_kernel, dim3(L.nseg), dim3(256), sizeof(float) * (spk_dim + d), s, spembs, spk_dim, w, b, d, L, hs));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// durations + per-utterance inclusive scan (one warp per utterance)
// ------------------------------------------------------------------------------------------------
// A huge or infinite log-duration (bad checkpoint) must not wrap the int32 frame count into [0, max_len]: per-token
// counts are clamped to 2^20 and totals saturate at 2^30 (32 x 2^20 fits an int32 warp sum), so such an utterance
// always reports more frames than any max_len and jatts_fs2_plan fails loudly.
static constexpr int kFrameSat = 1 << 30;
__device__ __forceinline__ int regulator_frames(float dv, float alpha) {
  const float f = (alpha != 1.0f) ? rintf(dv * alpha) : dv;
  return static_cast<int>(fminf(f, 1048576.0f));
}
__global__ void durations_kernel(const float* __restrict__ logd, float alpha, RowLayout L,
                                 const int* __restrict__ tok_off, long long* __restrict__ dur_out,
                                 int* __restrict__ cum, int* __restrict__ n_frames) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const int s0 = L.seg_start[b], T = L.seg_len[b];
  // pass 1: predicted durations (duration_predictor.py:86-90), the regulator's alpha-scaled copy
  // (length_regulator.py:81-83) and its exact integer total
  int tot = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int dl = 0;
    if (t < T) {
      const float dv = fmaxf(rintf(expf(logd[s0 + t]) - 1.0f), 0.0f);
      dur_out[tok_off[b] + t] = static_cast<long long>(fminf(dv, 9.0e18f));
      dl = regulator_frames(dv, alpha);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
    tot = min(tot + dl, kFrameSat);
  }
  const bool all_zero = (tot == 0);  // length_regulator.py:86-94 with B == 1: every token becomes 1
  int carry = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int dl = 0;
    if (t < T) {
      if (all_zero) {
        dl = 1;
        if (alpha == 1.0f) dur_out[tok_off[b] + t] = 1;  // in-place mutation visible to the caller (quirk 6)
      } else {
        const float dv = fmaxf(rintf(expf(logd[s0 + t]) - 1.0f), 0.0f);
        dl = regulator_frames(dv, alpha);
      }
    }
    int incl = dl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (t < T) cum[s0 + t] = min(carry + incl, kFrameSat);
    carry = min(carry + __shfl_sync(0xffffffffu, incl, 31), kFrameSat);
  }
  if (lane == 0) n_frames[b] = carry;
}
int durations_and_scan(const float* logd, float alpha, RowLayout L, const int* tok_off, long long* dur_out, int* cum,
                       int* n_frames, cudaStream_t s) {
  if (L.nseg == 0) return 0;
  JB_CUDA_OK(launch_pdl(durations_kernel, dim3(L.nseg), dim3(32), 0, s, logd, alpha, L, tok_off, dur_out, cum, n_frames));
  JB_KERNEL_OK();
  return 0;
}

__global__ void gather_scalar_kernel(const float* __restrict__ in, RowLayout L, const int* __restrict__ tok_off,
                                     float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= L.n_rows) return;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  out[tok_off[b] + (r - L.seg_start[b])] = in[r];
}
int gather_scalar(const float* in, RowLayout L, const int* tok_off, float* out, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(gather_scalar_kernel, dim3(ceil_div(L.n_rows, 256)), dim3(256), 0, s, in, L, tok_off, out));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LengthRegulator: prefix-sum gather (one warp per output frame row)
// ------------------------------------------------------------------------------------------------
__global__ void length_regulate_kernel(const float* __restrict__ hs, const float* __restrict__ pitch,
                                       const float* __restrict__ energy, const float* __restrict__ wp,
                                       const float* __restrict__ bp, const float* __restrict__ we,
                                       const float* __restrict__ be, int d, float scale, RowLayout Lt,
                                       const int* __restrict__ cum, RowLayout Lf, const int* __restrict__ frame_off,
                                       float* __restrict__ x_out, int* __restrict__ lr_index) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= Lf.n_rows) return;
  const int b = Lf.frame_seg[warp];
  if (b < 0) return;
  const int f = warp - Lf.seg_start[b];
  const int t0 = Lt.seg_start[b], T = Lt.seg_len[b];
  int lo = 0, hi = T - 1;  // first token whose inclusive cumulative duration exceeds f
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[t0 + mid] > f) hi = mid; else lo = mid + 1;
  }
  const int trow = t0 + lo;
  if (lane == 0) lr_index[frame_off[b] + f] = lo;
  const float p = pitch[trow], e = energy[trow];
  const float4* src = reinterpret_cast<const float4*>(hs + static_cast<long long>(trow) * d);
  float4* dst = reinterpret_cast<float4*>(x_out + static_cast<long long>(warp) * d);
  for (int i = lane; i < d / 4; i += 32) {
    const float4 h = src[i];
    const float4 a = reinterpret_cast<const float4*>(wp)[i], ab = reinterpret_cast<const float4*>(bp)[i];
    const float4 c = reinterpret_cast<const float4*>(we)[i], cb = reinterpret_cast<const float4*>(be)[i];
    float4 o;
    // fastspeech2.py:616: hs + e_embs + p_embs  (same association order)
    o.x = ((h.x + (e * c.x + cb.x)) + (p * a.x + ab.x)) * scale;
    o.y = ((h.y + (e * c.y + cb.y)) + (p * a.y + ab.y)) * scale;
    o.z = ((h.z + (e * c.z + cb.z)) + (p * a.z + ab.z)) * scale;
    o.w = ((h.w + (e * c.w + cb.w)) + (p * a.w + ab.w)) * scale;
    dst[i] = o;
  }
}
int length_regulate(const float* hs, const float* pitch, const float* energy, const float* wp, const float* bp,
                    const float* we, const float* be, int d, float scale, RowLayout Ltext, const int* cum,
                    RowLayout Lframe, const int* frame_off, float* x_out, int* lr_index, cudaStream_t s) {
  ProfileScope prof(s, PROF_LENGTH_REGULATE);
  JB_REQUIRE(d % 4 == 0, -2, "length_regulate: d % 4");
  if (Lframe.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(length_regulate_kernel, dim3(ceil_div(Lframe.n_rows, 8)), dim3(256), 0, s, hs, pitch, energy, wp, bp, we, be, d, scale, Ltext,
                                                                    cum, Lframe, frame_off, x_out, lr_index));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void unpack_rows_kernel(const float* __restrict__ in, int in_ld, int c, RowLayout L,
                                   const int* __restrict__ off, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  const long long orow = off[b] + (r - L.seg_start[b]);
  for (int i = threadIdx.x; i < c; i += blockDim.x) out[orow * c + i] = in[static_cast<long long>(r) * in_ld + i];
}
int unpack_rows(const float* in, int in_ld, int c, RowLayout L, const int* off, float* out, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(unpack_rows_kernel, dim3(L.n_rows), dim3(96), 0, s, in, in_ld, c, L, off, out));
  JB_KERNEL_OK();
  return 0;
}
__global__ void pack_mel_affine_kernel(const float* __restrict__ mel, int c, const float* __restrict__ a,
                                       const float* __restrict__ bb, RowLayout L, const int* __restrict__ off,
                                       bf16* __restrict__ out, int out_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  if (b < 0) {  // gap row: the operand buffer must read as zero padding here
    for (int i = threadIdx.x; i < out_ld; i += blockDim.x) out[static_cast<long long>(r) * out_ld + i] = __float2bfloat16_rn(0.f);
    return;
  }
  const long long irow = off[b] + (r - L.seg_start[b]);
  for (int i = threadIdx.x; i < out_ld; i += blockDim.x)
    out[static_cast<long long>(r) * out_ld + i] =
        i < c ? __float2bfloat16_rn(mel[irow * c + i] * a[i] + bb[i]) : __float2bfloat16_rn(0.f);
}
int pack_mel_affine(const float* mel, int c, const float* a, const float* b, RowLayout L, const int* off, bf16* out,
                    int out_ld, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(pack_mel_affine_kernel, dim3(L.n_rows), dim3(96), 0, s, mel, c, a, b, L, off, out, out_ld));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// HiFi-GAN output_conv (C -> 1, k taps) + tanh.  One thread = OC_PER consecutive output samples: an
// 8-channel slice of the k + OC_PER - 1 input rows is held in registers and every weight (one shared-memory
// load) feeds OC_PER FMAs.
// ------------------------------------------------------------------------------------------------
static constexpr int OC_PER = 4;
static constexpr int OC_MAXK = 7;

__global__ void output_conv_tanh_kernel(const bf16* __restrict__ x, int ld, int c, const float* __restrict__ w,
                                        float bias, int k, RowLayout L, int rate, const int* __restrict__ frame_off,
                                        float* __restrict__ wave, short* __restrict__ pcm, long long total_rows) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float ws[];  // [k][c]
  for (int i = threadIdx.x; i < k * c; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const long long row0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * OC_PER;
  if (row0 >= total_rows) return;
  const int pad = (k - 1) / 2;
  float acc[OC_PER];
#pragma unroll
  for (int o = 0; o < OC_PER; ++o) acc[o] = bias;
  for (int c8 = 0; c8 < c / 8; ++c8) {
    float xin[OC_MAXK + OC_PER - 1][8];
#pragma unroll
    for (int i = 0; i < OC_MAXK + OC_PER - 1; ++i) {
      const long long rr = row0 + i - pad;   // gap rows are zero: no per-tap utterance test is needed
      uint4 u = make_uint4(0, 0, 0, 0);
      if (i < k + OC_PER - 1 && rr >= 0 && rr < total_rows) u = reinterpret_cast<const uint4*>(x + rr * ld)[c8];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        xin[i][2 * q] = f.x;
        xin[i][2 * q + 1] = f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < OC_MAXK; ++j) {
      if (j < k) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const float wv = ws[j * c + c8 * 8 + ch];
#pragma unroll
          for (int o = 0; o < OC_PER; ++o) acc[o] += xin[o + j][ch] * wv;
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OC_PER; ++o) {
    const long long row = row0 + o;
    if (row >= total_rows) break;
    const int b = L.frame_seg[static_cast<int>(row / rate)];
    if (b < 0) continue;
    const long long t = row - static_cast<long long>(L.seg_start[b]) * rate;
    const float y = tanhf(acc[o]);
    const long long idx = static_cast<long long>(frame_off[b]) * rate + t;
    if (wave) wave[idx] = y;
    // PCM_16 as libsndfile writes float data (sf.write(..., "PCM_16"), tts_decode.py:250-255): lrintf(y * 0x7FFF)
    if (pcm) pcm[idx] = static_cast<short>(__float2int_rn(y * 32767.0f));
  }
}
// C = 32 fast path: a CTA stages OCT_TILE + k - 1 input rows in shared memory with coalesced 16-byte loads (the
// per-thread version above reads 16 bytes per 256-byte stride: 4x the sectors), row pitch 80 B so that the 128-bit
// reads of 8 consecutive rows hit 8 different bank groups.  Thread t computes rows t, t+256, ... of the tile at once
// (OCT_TILE / 256 outputs per thread), so every weight vector (a broadcast shared-memory load) feeds several outputs.
static constexpr int OCT_THREADS = 256, OCT_PITCH = 80;

template <int OCT_TILE>
__global__ void __launch_bounds__(OCT_THREADS)
output_conv32_tiled_kernel(const bf16* __restrict__ x, const float* __restrict__ w, float bias, int k, RowLayout L, int rate,
                           const int* __restrict__ frame_off, float* __restrict__ wave, short* __restrict__ pcm,
                           long long total_rows) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t oct_sm[];
  float* ws = reinterpret_cast<float*>(oct_sm);   // [k][32]
  uint8_t* xs = oct_sm + 1024;                     // [OCT_TILE + k - 1][80 B]
  const int tid = threadIdx.x;
  const long long tile0 = static_cast<long long>(blockIdx.x) * OCT_TILE;
  const int pad = (k - 1) / 2;
  for (int i = tid; i < k * 32; i += OCT_THREADS) ws[i] = w[i];
  const int n_chunks = (OCT_TILE + k - 1) * 4;
  for (int i = tid; i < n_chunks; i += OCT_THREADS) {
    const int r = i >> 2, ch = i & 3;
    const long long rr = tile0 + r - pad;   // gap rows are zero: no per-tap utterance test is needed
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (rr >= 0 && rr < total_rows) u = __ldg(reinterpret_cast<const uint4*>(x + rr * 32) + ch);
    *reinterpret_cast<uint4*>(xs + r * OCT_PITCH + ch * 16) = u;
  }
  __syncthreads();
  constexpr int OPT = OCT_TILE / OCT_THREADS;   // outputs per thread: rows tid, tid + 256, ...
  float acc[OPT];
#pragma unroll
  for (int o = 0; o < OPT; ++o) acc[o] = bias;
  for (int j = 0; j < k; ++j) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + j * 32 + ch * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + j * 32 + ch * 8 + 4);
#pragma unroll
      for (int o = 0; o < OPT; ++o) {
        const uint4 u = *reinterpret_cast<const uint4*>(xs + (tid + o * OCT_THREADS + j) * OCT_PITCH + ch * 16);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
        const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
        float a = acc[o];
        a = fmaf(f0.x, w0.x, a); a = fmaf(f0.y, w0.y, a); a = fmaf(f1.x, w0.z, a); a = fmaf(f1.y, w0.w, a);
        a = fmaf(f2.x, w1.x, a); a = fmaf(f2.y, w1.y, a); a = fmaf(f3.x, w1.z, a); a = fmaf(f3.y, w1.w, a);
        acc[o] = a;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const long long row = tile0 + tid + o * OCT_THREADS;
    if (row >= total_rows) break;
    const int b = L.frame_seg[static_cast<int>(row / rate)];
    if (b < 0) continue;
    const long long t = row - static_cast<long long>(L.seg_start[b]) * rate;
    const float y = tanhf(acc[o]);
    const long long idx = static_cast<long long>(frame_off[b]) * rate + t;
    if (wave) wave[idx] = y;
    if (pcm) pcm[idx] = static_cast<short>(__float2int_rn(y * 32767.0f));   // lrintf(y * 0x7FFF), see above
  }
}

int output_conv_tanh(const bf16* x, int ld, int c, const float* w, float bias, int k, RowLayout L, int rate,
                     const int* frame_off, float* wave, short* pcm, cudaStream_t s) {
  ProfileScope prof(s, PROF_OUTPUT_CONV);
  JB_REQUIRE(c % 8 == 0 && ld % 8 == 0 && k <= OC_MAXK, -2, "output_conv: C % 8, k <= 7");
  const long long total = static_cast<long long>(L.n_rows) * rate;
  if (total == 0) return 0;
  if (c == 32 && ld == 32 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    // rows per CTA: 512 (41 KB of shared memory, 5 CTAs per SM whose load and compute phases overlap) measured 0.217 ms per
    // 5.9 M samples against 0.279 ms with 1024 rows (2 CTAs per SM); JATTS_B200_OCT_TILE is the A/B switch
    static const int tile = getenv("JATTS_B200_OCT_TILE") ? atoi(getenv("JATTS_B200_OCT_TILE")) : 512;
    auto go = [&](auto tag) -> int {
      constexpr int T = decltype(tag)::value;
      const int smem = 1024 + (T + k - 1) * OCT_PITCH;
      JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(output_conv32_tiled_kernel<T>), 1024 + (T + OC_MAXK - 1) * OCT_PITCH));
      JB_CUDA_OK(launch_pdl(output_conv32_tiled_kernel<T>, dim3(static_cast<unsigned>((total + T - 1) / T)), dim3(OCT_THREADS), smem, s,
                            x, w, bias, k, L, rate, frame_off, wave, pcm, total));
      return 0;
    };
    if (tile == 1024) JB_PROPAGATE(go(std::integral_constant<int, 1024>{}));
    else JB_PROPAGATE(go(std::integral_constant<int, 512>{}));   // 256 rows per CTA measured the same as 512
    JB_KERNEL_OK();
    return 0;
  }
  const long long threads = (total + OC_PER - 1) / OC_PER;
  JB_CUDA_OK(launch_pdl(output_conv_tanh_kernel, dim3(static_cast<unsigned>((threads + 127) / 128)), dim3(128), sizeof(float) * k * c, s, 
      x, ld, c, w, bias, k, L, rate, frame_off, wave, pcm, total));
  JB_KERNEL_OK();
  return 0;
}

}  // namespace jb
