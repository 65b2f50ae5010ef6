// Bandwidth-bound kernels of the Matcha-TTS flow-matching decoder (jatts/modules/matchatts/decoder.py, transformer.py):
// GroupNorm + Mish (Block1D) and the packing of the initial noise (SnakeBeta is an epilogue of the split GEMM).  fp32 arithmetic, packed-with-gaps
// channels-last rows (common.cuh); the dense contractions around them run on the split-operand tcgen05 GEMM.
#include "conv_gemm.cuh"
#include "kernels.cuh"

namespace jb {

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics (decoder.py:67-71 Block1D: torch.nn.GroupNorm(groups, dim_out), eps 1e-5) of one utterance:
// mean / biased variance over (T frames) x (C / groups channels).  One CTA per (group, utterance); sums in double.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ x, int c, int groups, RowLayout L, float eps, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.x, b = blockIdx.y;
  const int cg = c / groups, cg4 = cg >> 2;
  const int T = L.seg_len[b];
  const float* base = x + static_cast<long long>(L.seg_start[b]) * c + g * cg;
  double s = 0.0, ss = 0.0;
  const int n4 = T * cg4;
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {   // four independent 16-byte loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) {
        const int r = i / cg4, q = i - r * cg4;
        v[u] = *reinterpret_cast<const float4*>(base + static_cast<long long>(r) * c + q * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float ps = (v[u].x + v[u].y) + (v[u].z + v[u].w);
      const float pq = (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      s += ps;
      ss += pq;
    }
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; q += red[1][w]; }
    const double n = static_cast<double>(T) * cg;
    const double mean = n > 0 ? a / n : 0.0;
    double var = n > 0 ? q / n - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    stats[b * groups + g] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
  }
}

// torch.nn.functional.mish: x * tanh(softplus(x)) with softplus(x) = x above the threshold 20.  With n = e^x,
// tanh(log(1 + n)) = (n^2 + 2n) / (n^2 + 2n + 2): one exponential and one division, no cancellation (every term is positive);
// same fp32 accuracy as the composed form (max abs difference to an fp64 evaluation 1.2e-6 on [-30, 30], both)
__device__ __forceinline__ float mish_f(float x) {
  const float n = expf(x);
  const float w = fmaf(n, n, 2.f * n);
  return x > 20.f ? x : x * (w / (w + 2.f));
}

// y = Mish(GroupNorm(x)) [+ add[c]] for the rows of an utterance, zeros elsewhere (the result is the operand of a k = 3
// convolution: its gap rows are that convolution's zero padding).  One warp per row.
__global__ void __launch_bounds__(256)
groupnorm_mish_kernel(const float* __restrict__ x, int c, int groups, const float2* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ add,
                      RowLayout L, float* __restrict__ y, bf16* __restrict__ hi, bf16* __restrict__ lo, int bf_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= L.n_rows) return;
  const int b = L.frame_seg[r];
  const int cg = c / groups;
  constexpr int MAXV = 4;   // float4 per lane: C <= 512 (checked by the launcher)
  float4 v[MAXV];
  // the row's loads are issued before any arithmetic (the exponential and the division of Mish would otherwise sit
  // between them and leave one 16-byte load in flight per lane)
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int i = lane * 4 + u * 128;
    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b >= 0 && i < c) v[u] = *reinterpret_cast<const float4*>(x + static_cast<long long>(r) * c + i);
  }
#pragma unroll
  for (int u = 0; u < MAXV; ++u) {
    const int i = lane * 4 + u * 128;
    if (i >= c) break;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b >= 0) {
      const float2 st = stats[b * groups + i / cg];   // cg % 4 == 0: the four channels share a group
      const float4 ga = *reinterpret_cast<const float4*>(gamma + i), be = *reinterpret_cast<const float4*>(beta + i);
      o.x = mish_f((v[u].x - st.x) * st.y * ga.x + be.x);
      o.y = mish_f((v[u].y - st.x) * st.y * ga.y + be.y);
      o.z = mish_f((v[u].z - st.x) * st.y * ga.z + be.z);
      o.w = mish_f((v[u].w - st.x) * st.y * ga.w + be.w);
      if (add != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(add + i);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
    }
    if (y != nullptr) *reinterpret_cast<float4*>(y + static_cast<long long>(r) * c + i) = o;
    if (hi != nullptr) {
      uint32_t h0, l0, h1, l1;
      split_pair16_sat(o.x, o.y, h0, l0);
      split_pair16_sat(o.z, o.w, h1, l1);
      *reinterpret_cast<uint2*>(hi + static_cast<long long>(r) * bf_ld + i) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(lo + static_cast<long long>(r) * bf_ld + i) = make_uint2(l0, l1);
    }
  }
}

int groupnorm_mish_rows(const float* x, int c, int groups, const float* gamma, const float* beta, float eps, const float* add,
                        RowLayout L, float2* stats, float* y, bf16* hi, bf16* lo, int bf_ld, cudaStream_t s) {
  ProfileScope prof(s, PROF_LAYERNORM);
  JB_REQUIRE(groups > 0 && c % groups == 0 && (c / groups) % 4 == 0 && c <= 512, -2,
             "groupnorm: channels per group must be a multiple of 4, C <= 512");
  JB_REQUIRE((hi != nullptr) == (lo != nullptr) && (y != nullptr || hi != nullptr) && bf_ld % 4 == 0, -2, "groupnorm: outputs");
  if (L.n_rows == 0 || L.nseg == 0) return 0;
  JB_CUDA_OK(launch_pdl(groupnorm_stats_kernel, dim3(dim3(groups, L.nseg)), dim3(256), 0, s, x, c, groups, L, eps, stats));
  JB_KERNEL_OK();
  JB_CUDA_OK(launch_pdl(groupnorm_mish_kernel, dim3(ceil_div(L.n_rows, 8)), dim3(256), 0, s, x, c, groups, stats, gamma, beta, add, L, y, hi, lo, bf_ld));
  JB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// utterance-contiguous fp32 [sum T, c] -> packed rows: y = in * scale (fp32 master) and its operand pair; gap rows
// of the pair are written as zeros (flow_matching.py:64: z = randn_like(mu) * temperature, then cat([x, mu]))
// ------------------------------------------------------------------------------------------------
__global__ void pack_rows_split_kernel(const float* __restrict__ in, int c, float scale, RowLayout L, const int* __restrict__ off,
                                       float* __restrict__ y, int y_ld, bf16* __restrict__ hi, bf16* __restrict__ lo, int bf_ld) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x;
  const int b = L.frame_seg[r];
  const long long irow = b >= 0 ? off[b] + (r - L.seg_start[b]) : 0;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const float v = b >= 0 ? in[irow * c + i] * scale : 0.f;
    bf16 h, l;
    split_op16(v, h, l);
    if (b >= 0) y[static_cast<long long>(r) * y_ld + i] = v;
    hi[static_cast<long long>(r) * bf_ld + i] = h;
    lo[static_cast<long long>(r) * bf_ld + i] = l;
  }
}
int pack_rows_split(const float* in, int c, float scale, RowLayout L, const int* off, float* y, int y_ld, bf16* hi, bf16* lo,
                    int bf_ld, cudaStream_t s) {
  if (L.n_rows == 0) return 0;
  JB_CUDA_OK(launch_pdl(pack_rows_split_kernel, dim3(L.n_rows), dim3(96), 0, s, in, c, scale, L, off, y, y_ld, hi, lo, bf_ld));
  JB_KERNEL_OK();
  return 0;
}

}  // namespace jb
