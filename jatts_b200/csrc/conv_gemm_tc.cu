// Dispatcher of the implicit-GEMM convolution (conv_gemm.cuh) and its CUDA-core test twin.
//
// Every product launch goes to one of two tcgen05 kernels: conv_bf16_tma_kernel (conv_gemm_tc2.cu, bf16 HiFi-GAN
// convolutions) or gemm_split_tma_kernel (conv_gemm_tc3.cu, fp16-split FastSpeech2 GEMMs).  A problem neither of them
// implements is an error (JATTS_E_UNSUPPORTED): there is no third kernel behind them.  (The first tcgen05 kernel of
// this path, with per-thread global stores in its epilogue, lived here; no shipped configuration launched it any
// more and it was removed.)
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/jatts_b200.h"
#include "conv_gemm.cuh"
#include "tc_common.cuh"

namespace jb {

static constexpr int BLOCK_K = 64;  // 16-bit elements = one 128-byte swizzle row

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
  if (conv_gemm_tc2_eligible(p)) return conv_gemm_tc2(p, stream);
  if (conv_gemm_tc3_eligible(p)) return conv_gemm_tc3(p, stream);
  JB_REQUIRE(false, JATTS_E_UNSUPPORTED,
             "conv_gemm: this combination of operands / epilogue options is implemented by neither tensor-core kernel");
  return JATTS_E_UNSUPPORTED;
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
