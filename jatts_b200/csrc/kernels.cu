// Bandwidth-bound / small kernels of the synthesis path (CUDA cores, fp32 arithmetic).
// Every kernel works on the packed-with-gaps row layout (common.cuh): gap rows are never written.
#include "kernels.cuh"

namespace jb {

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
__global__ void fill_layout_kernel(const int* __restrict__ seg_start, const int* __restrict__ seg_len, int nseg,
                                   int n_rows, uint8_t* __restrict__ mask, int* __restrict__ seg) {
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
  fill_layout_kernel<<<ceil_div(n_rows, 256), 256, 0, s>>>(seg_start, seg_len, nseg, n_rows, frame_mask, frame_seg);
  JB_KERNEL_OK();
  return 0;
}

// grid = (gap rows per utterance at this rate, utterances): only the rows that must be zero are touched
__global__ void zero_gap_rows_kernel(uint4* __restrict__ buf, int row_vec, const int* __restrict__ seg_start,
                                     const int* __restrict__ seg_len, int rate) {
  const int b = blockIdx.y;
  const long long r = (static_cast<long long>(seg_start[b]) + seg_len[b]) * rate + blockIdx.x;
  uint4* p = buf + r * row_vec;
  for (int i = threadIdx.x; i < row_vec; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}
int zero_gap_rows(void* buf, int row_bytes, RowLayout L, int rate, cudaStream_t s) {
  JB_REQUIRE(row_bytes % 16 == 0, -2, "zero_gap_rows: row_bytes % 16");
  if (L.nseg == 0) return 0;
  dim3 grid(kGapRows * rate, L.nseg);
  zero_gap_rows_kernel<<<grid, 64, 0, s>>>(static_cast<uint4*>(buf), row_bytes / 16, L.seg_start, L.seg_len, rate);
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// embedding
// ------------------------------------------------------------------------------------------------
__global__ void embed_tokens_kernel(const long long* __restrict__ tokens, const int* __restrict__ tok_off,
                                    const float* __restrict__ emb, int vocab, int d, float scale, RowLayout L,
                                    float* __restrict__ x) {
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
  embed_tokens_kernel<<<L.n_rows, 96, 0, s>>>(tokens, tok_off, emb, vocab, d, scale, L, x);
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
  JB_REQUIRE(c % 4 == 0 && c <= 128 * LN_MAXV, -2, "layernorm: C must be a multiple of 4 and <= 512");
  JB_REQUIRE(bf_ld % 4 == 0, -2, "layernorm: bf_ld % 4");
  if (L.n_rows == 0) return 0;
  layernorm_kernel<false><<<ceil_div(L.n_rows, 8), 256, 0, s>>>(x, c, gamma, beta, eps, L, y, hi, lo, bf_ld, nullptr,
                                                                0.f, nullptr);
  JB_KERNEL_OK();
  return 0;
}
int ln_dot_rows(const float* x, int c, const float* gamma, const float* beta, float eps, const float* w, float b,
                RowLayout L, float* out, cudaStream_t s) {
  JB_REQUIRE(c % 4 == 0 && c <= 128 * LN_MAXV, -2, "ln_dot: C must be a multiple of 4 and <= 512");
  if (L.n_rows == 0) return 0;
  layernorm_kernel<true><<<ceil_div(L.n_rows, 8), 256, 0, s>>>(x, c, gamma, beta, eps, L, nullptr, nullptr, nullptr, 4,
                                                               w, b, out);
  JB_KERNEL_OK();
  return 0;
}

__global__ void split_rows_kernel(const float* __restrict__ x, int c, RowLayout L, bf16* __restrict__ hi,
                                  bf16* __restrict__ lo, int bf_ld) {
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
  split_rows_kernel<<<L.n_rows, 128, 0, s>>>(x, c, L, hi, lo, bf_ld);
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// legacy relative-position attention (attention.py:164-206), fp32 on CUDA cores.
//
// One CTA = (query tile of R rows, utterance, head).  With BD[a,n] = (q_a + v) . p_n the reference's
// rel_shift (attention.py:142-162) has the closed form
//     shifted[a,b] = BD[a, T-1-a+b]   (b <= a),   0 (b == a+1),   BD[a+1, b-a-2]   (b > a+1)
// so score row a needs q_a against keys and against p, and q_{a+1} against p.
// ------------------------------------------------------------------------------------------------
static constexpr int ATT_TK = 64;       // keys / positions per smem tile
static constexpr int ATT_THREADS = 256;

template <int R>
__global__ void __launch_bounds__(ATT_THREADS)
relpos_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ pos, const float* __restrict__ bias_u,
                        const float* __restrict__ bias_v, int d_model, int dk, RowLayout L, int t_pad,
                        bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int out_ld) {
  constexpr int RPT = R / 4;  // query rows per thread
  const int b = blockIdx.y, h = blockIdx.z;
  const int T = L.seg_len[b];
  const int a0 = blockIdx.x * R;
  if (a0 >= T) return;
  const long long base = L.seg_start[b];
  const int ks = dk + 4;  // padded smem row stride (floats): conflict-free LDS.128
  extern __shared__ float sm[];
  float* qu = sm;                    // [R][dk]
  float* qv = qu + R * dk;           // [R+1][dk]
  float* S = qv + (R + 1) * dk;      // [R][t_pad]
  float* tile = S + R * t_pad;       // [ATT_TK][ks]
  const int tid = threadIdx.x;
  const int tx = tid & 63, ty = tid >> 6;
  const int ld3 = 3 * d_model;
  const int dk4 = dk >> 2;

  for (int i = tid; i < (R + 1) * dk; i += ATT_THREADS) {
    const int r = i / dk, d = i - r * dk;
    const int a = a0 + r;
    float q = 0.f;
    if (a < T) q = qkv[(base + a) * ld3 + h * dk + d];
    if (r < R) qu[r * dk + d] = (a < T) ? q + bias_u[h * dk + d] : 0.f;
    qv[r * dk + d] = (a < T) ? q + bias_v[h * dk + d] : 0.f;
  }
  __syncthreads();

  // ---- phase A: S[r][key] = (q_a + u) . k_key
  for (int k0 = 0; k0 < T; k0 += ATT_TK) {
    for (int i = tid; i < ATT_TK * dk4; i += ATT_THREADS) {
      const int kk = i / dk4, d4 = i - kk * dk4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < T) v = *reinterpret_cast<const float4*>(qkv + (base + k0 + kk) * ld3 + d_model + h * dk + d4 * 4);
      *reinterpret_cast<float4*>(tile + kk * ks + d4 * 4) = v;
    }
    __syncthreads();
    float acc[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
    const float* kr = tile + tx * ks;
    for (int d = 0; d < dk; d += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(qu + (ty * RPT + i) * dk + d);
        acc[i] += (q.x * kv.x + q.y * kv.y) + (q.z * kv.z + q.w * kv.w);
      }
    }
    if (k0 + tx < T) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) S[(ty * RPT + i) * t_pad + k0 + tx] = acc[i];
    }
    __syncthreads();
  }

  // ---- phase B: add the shifted positional term
  for (int n0 = 0; n0 < T; n0 += ATT_TK) {
    for (int i = tid; i < ATT_TK * dk4; i += ATT_THREADS) {
      const int kk = i / dk4, d4 = i - kk * dk4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + kk < T) v = *reinterpret_cast<const float4*>(pos + static_cast<long long>(n0 + kk) * d_model + h * dk + d4 * 4);
      *reinterpret_cast<float4*>(tile + kk * ks + d4 * 4) = v;
    }
    __syncthreads();
    float acc[RPT + 1];
#pragma unroll
    for (int i = 0; i <= RPT; ++i) acc[i] = 0.f;
    const float* pr = tile + tx * ks;
    for (int d = 0; d < dk; d += 4) {
      const float4 pv = *reinterpret_cast<const float4*>(pr + d);
#pragma unroll
      for (int i = 0; i <= RPT; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(qv + (ty * RPT + i) * dk + d);
        acc[i] += (q.x * pv.x + q.y * pv.y) + (q.z * pv.z + q.w * pv.w);
      }
    }
    const int n = n0 + tx;
    if (n < T) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = ty * RPT + i;
        const int a = a0 + r;
        if (a < T) {
          const int b1 = n - (T - 1 - a);
          if (b1 >= 0 && b1 <= a) S[r * t_pad + b1] += acc[i];
          const int b2 = n + a + 2;
          if (b2 < T) S[r * t_pad + b2] += acc[i + 1];
        }
      }
    }
    __syncthreads();
  }

  // ---- phase C: softmax over keys (scores / sqrt(dk))
  {
    const float scale = rsqrtf(static_cast<float>(dk));
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += ATT_THREADS / 32) {
      if (a0 + r >= T) continue;
      float* row = S + r * t_pad;
      float m = -INFINITY;
      for (int j = lane; j < T; j += 32) m = fmaxf(m, row[j]);
      m = warp_max(m) * scale;
      float sum = 0.f;
      for (int j = lane; j < T; j += 32) {
        const float e = expf(row[j] * scale - m);
        row[j] = e;
        sum += e;
      }
      const float inv = 1.0f / warp_sum(sum);
      for (int j = lane; j < T; j += 32) row[j] *= inv;
    }
  }
  __syncthreads();

  // ---- phase D: ctx = P . V
  float ctx[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ctx[i][j] = 0.f;
  for (int k0 = 0; k0 < T; k0 += ATT_TK) {
    for (int i = tid; i < ATT_TK * dk4; i += ATT_THREADS) {
      const int kk = i / dk4, d4 = i - kk * dk4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < T) v = *reinterpret_cast<const float4*>(qkv + (base + k0 + kk) * ld3 + 2 * d_model + h * dk + d4 * 4);
      *reinterpret_cast<float4*>(tile + kk * ks + d4 * 4) = v;
    }
    __syncthreads();
    const int kn = min(ATT_TK, T - k0);
    for (int kk = 0; kk < kn; ++kk) {
      float p[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) p[i] = S[(ty * RPT + i) * t_pad + k0 + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = tx + 64 * j;
        if (d < dk) {
          const float v = tile[kk * ks + d];
#pragma unroll
          for (int i = 0; i < RPT; ++i) ctx[i][j] += p[i] * v;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int a = a0 + ty * RPT + i;
    if (a >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = tx + 64 * j;
      if (d < dk) {
        bf16 hh, ll;
        split_op16(ctx[i][j], hh, ll);
        out_hi[(base + a) * out_ld + h * dk + d] = hh;
        if (out_lo) out_lo[(base + a) * out_ld + h * dk + d] = ll;
      }
    }
  }
}

template <int R>
static int launch_attention(const float* qkv, const float* pos, const float* bias_u, const float* bias_v, int n_head,
                            int d_model, RowLayout L, int max_len, bf16* out_hi, bf16* out_lo, int out_ld,
                            cudaStream_t s) {
  const int dk = d_model / n_head;
  const int t_pad = round_up(max_len, 4) + 4;
  const size_t smem = sizeof(float) * (static_cast<size_t>(R) * dk + (R + 1) * dk + static_cast<size_t>(R) * t_pad +
                                       ATT_TK * (dk + 4));
  JB_REQUIRE(smem <= 227 * 1024, -2, "attention: utterance too long for shared memory");
  auto kern = relpos_attention_kernel<R>;
  JB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(ceil_div(max_len, R), L.nseg, n_head);
  kern<<<grid, ATT_THREADS, smem, s>>>(qkv, pos, bias_u, bias_v, d_model, dk, L, t_pad, out_hi, out_lo, out_ld);
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Register-tiled variant for utterances up to ~440 positions (every shipped workload): one CTA =
// 64 query rows x (utterance, head), 256 threads, 4 rows x 4 keys per thread in the two score passes
// (8 FMA per shared-memory load instead of 3.2), float4-along-keys in the P.V pass.  The bias terms are
// split off the dot products: (q+u).k = q.k + u.k and (q+v).p = q.p + v.p, so one copy of the 65 query
// rows serves both passes; u.k / v.p are one extra 192-long dot per key / position of each tile.
// ------------------------------------------------------------------------------------------------
static constexpr int ATT64_R = 64;

__global__ void __launch_bounds__(ATT_THREADS, 1)
relpos_attention64_kernel(const float* __restrict__ qkv, const float* __restrict__ pos,
                          const float* __restrict__ bias_u, const float* __restrict__ bias_v, int d_model, int dk,
                          RowLayout L, int t_pad, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int out_ld) {
  constexpr int R = ATT64_R;
  const int b = blockIdx.y, h = blockIdx.z;
  const int T = L.seg_len[b];
  const int a0 = blockIdx.x * R;
  if (a0 >= T) return;
  const long long base = L.seg_start[b];
  const int ks = dk + 4;
  extern __shared__ float sm[];
  float* q = sm;                         // [R+1][ks]   raw queries (row R = first row of the next tile)
  float* S = q + (R + 1) * ks;           // [R][t_pad]
  float* tile = S + R * t_pad;           // [ATT_TK][ks]
  float* cvec = tile + ATT_TK * ks;      // [ATT_TK]    u.k (pass A) / v.p (pass B) of the tile
  float* ub = cvec + ATT_TK;             // [dk] bias_u of this head, then [dk] bias_v
  const int tid = threadIdx.x;
  const int ld3 = 3 * d_model;
  const int dk4 = dk >> 2;
  const int tx = tid & 15, ty = tid >> 4;   // 16 key lanes (4 keys each: tx, tx+16, tx+32, tx+48), 16 row groups of 4

  for (int i = tid; i < (R + 1) * dk4; i += ATT_THREADS) {
    const int r = i / dk4, d4 = i - r * dk4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a0 + r < T) v = *reinterpret_cast<const float4*>(qkv + (base + a0 + r) * ld3 + h * dk + d4 * 4);
    *reinterpret_cast<float4*>(q + r * ks + d4 * 4) = v;
  }
  for (int i = tid; i < 2 * dk; i += ATT_THREADS) ub[i] = i < dk ? bias_u[h * dk + i] : bias_v[h * dk + i - dk];
  __syncthreads();

  // load one 64-row tile of k / p / v (zero rows past T) and, for the score passes, bias . row
  auto load_tile = [&](const float* src, long long row_stride, int r0, const float* bias_vec) {
    for (int i = tid; i < ATT_TK * dk4; i += ATT_THREADS) {
      const int kk = i / dk4, d4 = i - kk * dk4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + kk < T) v = *reinterpret_cast<const float4*>(src + static_cast<long long>(r0 + kk) * row_stride + d4 * 4);
      *reinterpret_cast<float4*>(tile + kk * ks + d4 * 4) = v;
    }
    __syncthreads();
    if (bias_vec) {
      const int kk = tid >> 2, part = tid & 3;   // 4 threads per row, dk/4 elements each
      float acc = 0.f;
      const int per = dk >> 2;
      for (int d = part * per; d < (part + 1) * per; ++d) acc += tile[kk * ks + d] * bias_vec[d];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) cvec[kk] = acc;
      __syncthreads();
    }
  };

  // ---- pass A: S[r][key] = q_a . k_key + u . k_key
  for (int k0 = 0; k0 < T; k0 += ATT_TK) {
    load_tile(qkv + base * ld3 + d_model + h * dk, ld3, k0, ub);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d = 0; d < dk; d += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(q + (ty * 4 + i) * ks + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(tile + (tx + 16 * j) * ks + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] += (qv[i].x * kv[j].x + qv[i].y * kv[j].y) + (qv[i].z * kv[j].z + qv[i].w * kv[j].w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int key = k0 + tx + 16 * j;
      if (key < T) {
        const float cu = cvec[tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) S[(ty * 4 + i) * t_pad + key] = acc[i][j] + cu;
      }
    }
    __syncthreads();
  }

  // ---- pass B: BD[a][n] = q_a . p_n + v . p_n, scattered through the closed-form legacy rel-shift
  for (int n0 = 0; n0 < T; n0 += ATT_TK) {
    load_tile(pos + h * dk, d_model, n0, ub + dk);
    float acc[5][4];   // rows ty*4 .. ty*4+4 (the extra row feeds the "a+1" branch of the shift)
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int d = 0; d < dk; d += 4) {
      float4 qv[5], pv[4];
#pragma unroll
      for (int i = 0; i < 5; ++i) qv[i] = *reinterpret_cast<const float4*>(q + (ty * 4 + i) * ks + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) pv[j] = *reinterpret_cast<const float4*>(tile + (tx + 16 * j) * ks + d);
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[i][j] += (qv[i].x * pv[j].x + qv[i].y * pv[j].y) + (qv[i].z * pv[j].z + qv[i].w * pv[j].w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < T) {
        const float cv = cvec[tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = ty * 4 + i;
          const int a = a0 + r;
          if (a < T) {
            const int b1 = n - (T - 1 - a);
            if (b1 >= 0 && b1 <= a) S[r * t_pad + b1] += acc[i][j] + cv;
            const int b2 = n + a + 2;
            if (b2 < T) S[r * t_pad + b2] += acc[i + 1][j] + cv;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- softmax over keys (scores / sqrt(dk))
  {
    const float scale = rsqrtf(static_cast<float>(dk));
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += ATT_THREADS / 32) {
      if (a0 + r >= T) continue;
      float* row = S + r * t_pad;
      float m = -INFINITY;
      for (int j = lane; j < T; j += 32) m = fmaxf(m, row[j]);
      m = warp_max(m) * scale;
      float sum = 0.f;
      for (int j = lane; j < T; j += 32) {
        const float e = expf(row[j] * scale - m);
        row[j] = e;
        sum += e;
      }
      const float inv = 1.0f / warp_sum(sum);
      for (int j = lane; j < T; j += 32) row[j] *= inv;
      for (int j = T + lane; j < t_pad; j += 32) row[j] = 0.f;   // the P.V pass reads keys in groups of 4
    }
  }
  __syncthreads();

  // ---- ctx = P . V : thread = (16-row group, dim lane), 3 dims (d, d+64, d+128), keys 4 at a time
  const int dx = tid & 63, rg = tid >> 6;
  float ctx[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ctx[i][j] = 0.f;
  for (int k0 = 0; k0 < T; k0 += ATT_TK) {
    load_tile(qkv + base * ld3 + 2 * d_model + h * dk, ld3, k0, nullptr);
    const int kn = min(ATT_TK, T - k0);
    for (int kk = 0; kk < kn; kk += 4) {
      float vv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = dx + 64 * j;
          vv[u][j] = d < dk ? tile[(kk + u) * ks + d] : 0.f;   // rows past T are zero-filled
        }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 p = *reinterpret_cast<const float4*>(S + (rg * 16 + i) * t_pad + k0 + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          ctx[i][j] += (p.x * vv[0][j] + p.y * vv[1][j]) + (p.z * vv[2][j] + p.w * vv[3][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int a = a0 + rg * 16 + i;
    if (a >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = dx + 64 * j;
      if (d < dk) {
        bf16 hh, ll;
        split_op16(ctx[i][j], hh, ll);
        out_hi[(base + a) * out_ld + h * dk + d] = hh;
        if (out_lo) out_lo[(base + a) * out_ld + h * dk + d] = ll;
      }
    }
  }
}

static int launch_attention64(const float* qkv, const float* pos, const float* bias_u, const float* bias_v, int n_head,
                              int d_model, RowLayout L, int max_len, bf16* out_hi, bf16* out_lo, int out_ld,
                              cudaStream_t s, bool* launched) {
  const int dk = d_model / n_head;
  const int t_pad = round_up(max_len, 4) + 4;
  const size_t smem = sizeof(float) * (static_cast<size_t>(ATT64_R + 1) * (dk + 4) + static_cast<size_t>(ATT64_R) * t_pad +
                                       ATT_TK * (dk + 4) + ATT_TK + 2 * dk);
  *launched = false;
  if (smem > 227 * 1024 || (dk & 15) != 0) return 0;   // too long for this variant: caller falls back
  static size_t attr = 0;
  if (smem > attr) {
    JB_CUDA_OK(cudaFuncSetAttribute(relpos_attention64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  dim3 grid(ceil_div(max_len, ATT64_R), L.nseg, n_head);
  relpos_attention64_kernel<<<grid, ATT_THREADS, smem, s>>>(qkv, pos, bias_u, bias_v, d_model, dk, L, t_pad, out_hi, out_lo, out_ld);
  JB_KERNEL_OK();
  *launched = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variant (legacy warp-level MMA, m16n8k8 TF32) with the 3xTF32 split: x = hi + lo with
// hi = x truncated to TF32 (exact remainder lo = x - hi, rounded to TF32); a.b ~= hi.hi + hi.lo + lo.hi
// accumulated in fp32, which keeps scores and context FP32-faithful (the dropped lo.lo term is ~2^-21
// relative).  tcgen05 is not used here: the operands are per-(utterance, head) 64 x d_k tiles that need a
// register-level split and a scatter epilogue (the rel-shift), which the warp-level fragment layout gives
// directly.
//
// One CTA = 64 query rows a0..a0+63 of one (utterance, head), of which it OWNS the first 63: the legacy
// rel-shift sends BD[a'][n] = (q_a' + v) . p_n to row a' (keys <= a') and to row a'-1 (keys >= a'+1), so
// row r is complete once BD rows r and r+1 are known; tiles advance by 63 rows and the 64th row only
// contributes its BD.  8 warps = 4 row blocks of 16 x 2 halves of the 64-key tile (score passes) or of the
// head dimension (P.V pass).  The biases are added while the A fragments are built ((q+u).k, (q+v).p as
// the reference computes them).  The next k / p / v tile is prefetched into registers while the current
// one is in the MMAs.  Shared-memory strides: d_k+4 (== 4 mod 32) for operands read as [g][t], d_k+8
// (== 8 mod 32) for the V tile read as [t][g], t_pad with t_pad/4 odd: fragment loads are conflict-free.
// ------------------------------------------------------------------------------------------------
static constexpr int ATT_MMA_OWN = 63;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;   // the MMA ignores the low 13 bits: +half ulp = round
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], float b0,
                                           float b1) {
  uint32_t b0h, b0l, b1h, b1l;
  split_tf32(b0, b0h, b0l);
  split_tf32(b1, b1h, b1l);
  mma_tf32(c, alo, b0h, b1h);
  mma_tf32(c, ahi, b0l, b1l);
  mma_tf32(c, ahi, b0h, b1h);
}

template <int DK>
__global__ void __launch_bounds__(ATT_THREADS, 1)
relpos_attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ pos,
                            const float* __restrict__ bias_u, const float* __restrict__ bias_v, int d_model,
                            RowLayout L, int t_pad, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int out_ld) {
  constexpr int R = 64;
  constexpr int KS = DK + 4;    // q / k / p row stride (words)
  constexpr int KV = DK + 8;    // v row stride
  constexpr int DK4 = DK / 4;
  constexpr int NT_C = DK / 16; // 8-wide n-tiles per warp in the P.V pass
  constexpr int NPRE = ATT_TK * DK4 / ATT_THREADS;   // float4 per thread per tile
  static_assert(ATT_TK * DK4 % ATT_THREADS == 0, "tile must divide evenly over the CTA");
  const int b = blockIdx.y, h = blockIdx.z;
  const int T = L.seg_len[b];
  const int a0 = blockIdx.x * ATT_MMA_OWN;
  if (a0 >= T) return;
  const long long base = L.seg_start[b];
  extern __shared__ float sm[];
  float* q = sm;                         // [R][KS]
  float* S = q + R * KS;                 // [R][t_pad]
  float* tile = S + R * t_pad;           // [ATT_TK][KV]
  float* ub = tile + ATT_TK * KV;        // [2*DK] bias_u, bias_v of this head
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int rb = warp & 3, half = warp >> 2;
  const int ld3 = 3 * d_model;
  const int n_tiles = (T + ATT_TK - 1) / ATT_TK;

  // tile jobs: n_tiles of k, n_tiles of p, n_tiles of v; job j+1 is fetched while job j is computed
  float4 pre[NPRE];
  auto fetch = [&](int job) {
    const int pass = job / n_tiles, r0 = (job - pass * n_tiles) * ATT_TK;
    const float* src = pass == 1 ? pos + h * DK : qkv + base * ld3 + (pass == 0 ? d_model : 2 * d_model) + h * DK;
    const long long row_stride = pass == 1 ? d_model : ld3;
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
      const int idx = tid + i * ATT_THREADS;
      const int kk = idx / DK4, d4 = idx - kk * DK4;
      pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + kk < T) pre[i] = *reinterpret_cast<const float4*>(src + static_cast<long long>(r0 + kk) * row_stride + d4 * 4);
    }
  };
  auto commit = [&](int stride) {
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
      const int idx = tid + i * ATT_THREADS;
      const int kk = idx / DK4, d4 = idx - kk * DK4;
      *reinterpret_cast<float4*>(tile + kk * stride + d4 * 4) = pre[i];
    }
  };

  fetch(0);
  for (int i = tid; i < R * DK4; i += ATT_THREADS) {
    const int r = i / DK4, d4 = i - r * DK4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a0 + r < T) v = *reinterpret_cast<const float4*>(qkv + (base + a0 + r) * ld3 + h * DK + d4 * 4);
    *reinterpret_cast<float4*>(q + r * KS + d4 * 4) = v;
  }
  for (int i = tid; i < 2 * DK; i += ATT_THREADS) ub[i] = i < DK ? bias_u[h * DK + i] : bias_v[h * DK + i - DK];

  // 16 rows (rb) x 32 tile rows (half) x DK:  acc[nt] = (q + bias) . tile^T
  auto score_tile = [&](float (&acc)[4][4], const float* bias) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* qa = q + (rb * 16 + g) * KS + t;
    const float* kb = tile + (half * 32 + g) * KS + t;
    const float* bv = bias + t;
#pragma unroll 4
    for (int k = 0; k < DK; k += 8) {
      uint32_t ahi[4], alo[4];
      const float u0 = bv[k], u1 = bv[k + 4];
      split_tf32(qa[k] + u0, ahi[0], alo[0]);
      split_tf32(qa[8 * KS + k] + u0, ahi[1], alo[1]);
      split_tf32(qa[k + 4] + u1, ahi[2], alo[2]);
      split_tf32(qa[8 * KS + k + 4] + u1, ahi[3], alo[3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_3xtf32(acc[nt], ahi, alo, kb[nt * 8 * KS + k], kb[nt * 8 * KS + k + 4]);
    }
  };

  int job = 0;
  // ---- pass A: S[r][key] = (q_a + u) . k_key
  for (int k0 = 0; k0 < T; k0 += ATT_TK, ++job) {
    commit(KS);
    __syncthreads();
    fetch(job + 1);
    float acc[4][4];
    score_tile(acc, ub);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = rb * 16 + g + (e >> 1) * 8;
        const int key = k0 + half * 32 + nt * 8 + 2 * t + (e & 1);
        if (key < T) S[r * t_pad + key] = acc[nt][e];
      }
    __syncthreads();
  }

  // ---- pass B: BD[a'][n] = (q_a' + v) . p_n lands in row a' (keys <= a') and row a'-1 (keys >= a'+1)
  for (int n0 = 0; n0 < T; n0 += ATT_TK, ++job) {
    commit(KS);
    __syncthreads();
    fetch(job + 1);
    float acc[4][4];
    score_tile(acc, ub + DK);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = rb * 16 + g + (e >> 1) * 8;
        const int n = n0 + half * 32 + nt * 8 + 2 * t + (e & 1);
        const int a = a0 + r;
        if (n < T && a < T) {
          const int b1 = n - (T - 1 - a);
          if (b1 >= 0 && b1 <= a) S[r * t_pad + b1] += acc[nt][e];
          const int b2 = n + a + 1;
          if (r > 0 && b2 < T) S[(r - 1) * t_pad + b2] += acc[nt][e];
        }
      }
    __syncthreads();
  }

  // ---- softmax over keys (scores / sqrt(dk)); the first v tile is already in flight
  {
    const float scale = rsqrtf(static_cast<float>(DK));
    for (int r = warp; r < R; r += ATT_THREADS / 32) {
      float* row = S + r * t_pad;
      if (r >= ATT_MMA_OWN || a0 + r >= T) {   // rows that are not stored still enter the MMA: keep them finite
        for (int j = lane; j < t_pad; j += 32) row[j] = 0.f;
        continue;
      }
      float m = -INFINITY;
      for (int j = lane; j < T; j += 32) m = fmaxf(m, row[j]);
      m = warp_max(m) * scale;
      float sum = 0.f;
      for (int j = lane; j < T; j += 32) {
        const float e = expf(row[j] * scale - m);
        row[j] = e;
        sum += e;
      }
      const float inv = 1.0f / warp_sum(sum);
      for (int j = lane; j < T; j += 32) row[j] *= inv;
      for (int j = T + lane; j < t_pad; j += 32) row[j] = 0.f;
    }
  }

  // ---- ctx = P . V : warp = 16 rows x DK/2 dims
  float ctx[NT_C][4];
#pragma unroll
  for (int i = 0; i < NT_C; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ctx[i][j] = 0.f;
  for (int k0 = 0; k0 < T; k0 += ATT_TK, ++job) {
    commit(KV);
    __syncthreads();
    if (job + 1 < 3 * n_tiles) fetch(job + 1);
    const int kn = min(ATT_TK, T - k0);
    const float* pa = S + (rb * 16 + g) * t_pad + k0 + t;
    const float* vb = tile + t * KV + half * (DK / 2) + g;
    for (int kk = 0; kk < kn; kk += 8) {
      uint32_t ahi[4], alo[4];
      split_tf32(pa[kk], ahi[0], alo[0]);
      split_tf32(pa[8 * t_pad + kk], ahi[1], alo[1]);
      split_tf32(pa[kk + 4], ahi[2], alo[2]);
      split_tf32(pa[8 * t_pad + kk + 4], ahi[3], alo[3]);
#pragma unroll
      for (int nt = 0; nt < NT_C; ++nt) mma_3xtf32(ctx[nt], ahi, alo, vb[kk * KV + nt * 8], vb[(kk + 4) * KV + nt * 8]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < NT_C; ++nt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int r = rb * 16 + g + hr * 8;
      const int a = a0 + r;
      if (r >= ATT_MMA_OWN || a >= T) continue;
      const int d = half * (DK / 2) + nt * 8 + 2 * t;
      bf16 h0, l0, h1, l1;
      split_op16(ctx[nt][hr * 2], h0, l0);
      split_op16(ctx[nt][hr * 2 + 1], h1, l1);
      const long long o = (base + a) * out_ld + h * DK + d;
      *reinterpret_cast<uint32_t*>(out_hi + o) =
          static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      if (out_lo)
        *reinterpret_cast<uint32_t*>(out_lo + o) =
            static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
}

template <int DK>
static int launch_attention_mma(const float* qkv, const float* pos, const float* bias_u, const float* bias_v, int n_head,
                                int d_model, RowLayout L, int max_len, bf16* out_hi, bf16* out_lo, int out_ld,
                                cudaStream_t s, bool* launched) {
  const int t_pad = round_up(max_len, 8) + 4;   // t_pad / 4 odd
  const size_t smem = sizeof(float) * (static_cast<size_t>(64) * (DK + 4) + static_cast<size_t>(64) * t_pad +
                                       ATT_TK * (DK + 8) + 2 * DK);
  *launched = false;
  if (smem > 227 * 1024 || (out_ld & 1)) return 0;
  static size_t attr = 0;
  if (smem > attr) {
    JB_CUDA_OK(cudaFuncSetAttribute(relpos_attention_mma_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  dim3 grid(ceil_div(max_len, ATT_MMA_OWN), L.nseg, n_head);
  relpos_attention_mma_kernel<DK><<<grid, ATT_THREADS, smem, s>>>(qkv, pos, bias_u, bias_v, d_model, L, t_pad, out_hi, out_lo, out_ld);
  JB_KERNEL_OK();
  *launched = true;
  return 0;
}

int relpos_attention(const float* qkv, const float* pos, const float* bias_u, const float* bias_v, int n_head,
                     int d_model, RowLayout L, int max_len, bf16* out_hi, bf16* out_lo, int out_ld, cudaStream_t s) {
  const int dk = d_model / n_head;
  JB_REQUIRE(dk * n_head == d_model && dk % 4 == 0 && dk <= 256, -2, "attention: d_k must be a multiple of 4, <= 256");
  if (L.nseg == 0 || max_len == 0) return 0;
  if (!getenv("JATTS_B200_NO_ATT_MMA")) {
    bool launched = false;
    switch (dk) {
      case 32: JB_PROPAGATE(launch_attention_mma<32>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s, &launched)); break;
      case 64: JB_PROPAGATE(launch_attention_mma<64>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s, &launched)); break;
      case 128: JB_PROPAGATE(launch_attention_mma<128>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s, &launched)); break;
      case 192: JB_PROPAGATE(launch_attention_mma<192>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s, &launched)); break;
      default: break;
    }
    if (launched) return 0;
  }
  {
    bool launched = false;
    JB_PROPAGATE(launch_attention64(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s, &launched));
    if (launched) return 0;
  }
  if (max_len <= 1800) return launch_attention<16>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s);
  if (max_len <= 4000) return launch_attention<8>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s);
  return launch_attention<4>(qkv, pos, bias_u, bias_v, n_head, d_model, L, max_len, out_hi, out_lo, out_ld, s);
}

// ------------------------------------------------------------------------------------------------
// depthwise conv + folded BatchNorm + Swish
// ------------------------------------------------------------------------------------------------
__global__ void dwconv_swish_kernel(const float* __restrict__ g, int c, const float* __restrict__ wT,
                                    const float* __restrict__ bias, int k, RowLayout L, bf16* __restrict__ out_hi,
                                    bf16* __restrict__ out_lo, int out_ld) {
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
int dwconv_swish(const float* g, int c, const float* wT, const float* bias, int k, RowLayout L, bf16* out_hi,
                 bf16* out_lo, int out_ld, cudaStream_t s) {
  JB_REQUIRE(c % 4 == 0 && out_ld % 4 == 0 && (k & 1) == 1, -2, "dwconv: C % 4, odd k");
  if (L.n_rows == 0) return 0;
  dwconv_swish_kernel<<<L.n_rows, 96, 0, s>>>(g, c, wT, bias, k, L, out_hi, out_lo, out_ld);
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// speaker embedding integration ("add")
// ------------------------------------------------------------------------------------------------
__global__ void add_speaker_kernel(const float* __restrict__ spembs, int spk_dim, const float* __restrict__ w,
                                   const float* __restrict__ bvec, int d, RowLayout L, float* __restrict__ hs) {
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
  add_speaker_kernel<<<L.nseg, 256, sizeof(float) * (spk_dim + d), s>>>(spembs, spk_dim, w, b, d, L, hs);
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
  durations_kernel<<<L.nseg, 32, 0, s>>>(logd, alpha, L, tok_off, dur_out, cum, n_frames);
  JB_KERNEL_OK();
  return 0;
}

__global__ void gather_scalar_kernel(const float* __restrict__ in, RowLayout L, const int* __restrict__ tok_off,
                                     float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= L.n_rows) return;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  out[tok_off[b] + (r - L.seg_start[b])] = in[r];
}
int gather_scalar(const float* in, RowLayout L, const int* tok_off, float* out, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  gather_scalar_kernel<<<ceil_div(L.n_rows, 256), 256, 0, s>>>(in, L, tok_off, out);
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
  JB_REQUIRE(d % 4 == 0, -2, "length_regulate: d % 4");
  if (Lframe.n_rows == 0) return 0;
  length_regulate_kernel<<<ceil_div(Lframe.n_rows, 8), 256, 0, s>>>(hs, pitch, energy, wp, bp, we, be, d, scale, Ltext,
                                                                    cum, Lframe, frame_off, x_out, lr_index);
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pack / unpack
// ------------------------------------------------------------------------------------------------
__global__ void unpack_rows_kernel(const float* __restrict__ in, int in_ld, int c, RowLayout L,
                                   const int* __restrict__ off, float* __restrict__ out) {
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  if (b < 0) return;
  const long long orow = off[b] + (r - L.seg_start[b]);
  for (int i = threadIdx.x; i < c; i += blockDim.x) out[orow * c + i] = in[static_cast<long long>(r) * in_ld + i];
}
int unpack_rows(const float* in, int in_ld, int c, RowLayout L, const int* off, float* out, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  unpack_rows_kernel<<<L.n_rows, 96, 0, s>>>(in, in_ld, c, L, off, out);
  JB_KERNEL_OK();
  return 0;
}
__global__ void pack_mel_affine_kernel(const float* __restrict__ mel, int c, const float* __restrict__ a,
                                       const float* __restrict__ bb, RowLayout L, const int* __restrict__ off,
                                       bf16* __restrict__ out, int out_ld) {
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
  pack_mel_affine_kernel<<<L.n_rows, 96, 0, s>>>(mel, c, a, b, L, off, out, out_ld);
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
// reads of 8 consecutive rows hit 8 different bank groups.  Thread t computes rows t, t+256, t+512, t+768 of the
// tile at once, so every weight vector (a broadcast shared-memory load) feeds 4 outputs.
static constexpr int OCT_TILE = 1024, OCT_THREADS = 256, OCT_PITCH = 80;

__global__ void __launch_bounds__(OCT_THREADS)
output_conv32_tiled_kernel(const bf16* __restrict__ x, const float* __restrict__ w, float bias, int k, RowLayout L, int rate,
                           const int* __restrict__ frame_off, float* __restrict__ wave, short* __restrict__ pcm,
                           long long total_rows) {
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
  float acc[4] = {bias, bias, bias, bias};
  for (int j = 0; j < k; ++j) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + j * 32 + ch * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + j * 32 + ch * 8 + 4);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
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
  for (int o = 0; o < 4; ++o) {
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
  JB_REQUIRE(c % 8 == 0 && ld % 8 == 0 && k <= OC_MAXK, -2, "output_conv: C % 8, k <= 7");
  const long long total = static_cast<long long>(L.n_rows) * rate;
  if (total == 0) return 0;
  if (c == 32 && ld == 32 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int smem = 1024 + (OCT_TILE + k - 1) * OCT_PITCH;
    JB_PROPAGATE(ensure_dynamic_smem(reinterpret_cast<const void*>(output_conv32_tiled_kernel),
                                     1024 + (OCT_TILE + OC_MAXK - 1) * OCT_PITCH));
    output_conv32_tiled_kernel<<<static_cast<unsigned>((total + OCT_TILE - 1) / OCT_TILE), OCT_THREADS, smem, s>>>(
        x, w, bias, k, L, rate, frame_off, wave, pcm, total);
    JB_KERNEL_OK();
    return 0;
  }
  const long long threads = (total + OC_PER - 1) / OC_PER;
  output_conv_tanh_kernel<<<static_cast<unsigned>((threads + 127) / 128), 128, sizeof(float) * k * c, s>>>(
      x, ld, c, w, bias, k, L, rate, frame_off, wave, pcm, total);
  JB_KERNEL_OK();
  return 0;
}

}  // namespace jb
