// Shared device/host helpers for the jatts_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

namespace jb {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns 0 or a negative JATTS_E_* code; the message of the last
// failure on this thread is kept for jatts_last_error().
// ----------------------------------------------------------------------------------------------
extern long long g_launch_count;
void set_last_error(const std::string& msg);
const char* get_last_error();
// opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device (cached per device and kernel)
int ensure_dynamic_smem(const void* kernel, int bytes);

#define JB_CUDA_OK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      jb::set_last_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +     \
                         __FILE__ + ":" + std::to_string(__LINE__));                              \
      return -3; /* JATTS_E_CUDA */                                                               \
    }                                                                                             \
  } while (0)

// after every kernel launch: count it (jatts_launch_count) and surface launch errors
#define JB_KERNEL_OK()                                                                            \
  do {                                                                                            \
    ++jb::g_launch_count;                                                                         \
    JB_CUDA_OK(cudaGetLastError());                                                               \
  } while (0)

#define JB_REQUIRE(cond, code, msg)                                                               \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      jb::set_last_error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +                 \
                         std::to_string(__LINE__));                                               \
      return (code);                                                                              \
    }                                                                                             \
  } while (0)

#define JB_PROPAGATE(expr)                                                                        \
  do {                                                                                            \
    int _rc = (expr);                                                                             \
    if (_rc != 0) return _rc;                                                                     \
  } while (0)

// ----------------------------------------------------------------------------------------------
// packed-with-gaps row layout shared by every kernel
//
// A batch of utterances is stored as ONE [rows, C] channels-last matrix.  Utterance b owns rows
// [seg_start[b]*rate, (seg_start[b]+seg_len[b])*rate); between utterances (and before the first /
// after the last) there are >= GAP zero rows at rate 1, i.e. GAP*rate rows at an upsampled rate.
// Convolutions therefore see each utterance's own zero padding with no per-tap boundary test:
// gap rows are zero at allocation and no kernel ever writes them (stores are masked by
// frame_mask[row / rate]).
// ----------------------------------------------------------------------------------------------
struct RowLayout {
  const uint8_t* frame_mask;  // [n_frames_rows] 1 = row belongs to an utterance (rate-1 granularity)
  const int* frame_seg;       // [n_frames_rows] utterance index or -1
  const int* seg_start;       // [nseg] first row (rate-1 units)
  const int* seg_len;         // [nseg] length  (rate-1 units)
  int nseg;
  int n_rows;                 // rows at rate 1 (including gaps)
};

static constexpr int kGapRows = 8;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

#ifdef __CUDACC__
// Split operand pair of the fp32-faithful GEMMs: x ~= hi + lo * 2^-11 with hi = fp16(x) and
// lo = fp16((x - hi) * 2^11), i.e. ~22 significand bits (|err| <= 2^-23 |x|), versus 2^-18 for a
// bf16 hi/lo pair -- measured: the bf16 pair leaves 3-8e-4 max-abs mel error on the JSUT model, too
// close to the 1e-3 budget.  Scaling lo keeps it out of the fp16 subnormal range; the GEMM keeps the
// lo*hi / hi*lo products in their own TMEM accumulator and the epilogue rescales it by 2^-11.
// The 16-bit patterns are carried in bf16-typed storage (the buffers are opaque operand memory).
static constexpr float kSplitScale = 2048.0f;
__device__ __forceinline__ void split_op16(float x, bf16& hi, bf16& lo) {
  x = x > 65504.f ? 65504.f : (x < -65504.f ? -65504.f : x);  // saturate; NaN propagates
  const __half h = __float2half_rn(x);
  const __half l = __float2half_rn((x - __half2float(h)) * kSplitScale);
  hi = *reinterpret_cast<const bf16*>(&h);
  lo = *reinterpret_cast<const bf16*>(&l);
}
// Two values at once with the packed conversions (one F2FP per pair instead of three scalar F2F per value, which run
// at a quarter of the ALU rate): returns the packed (hi, hi) and (lo', lo') fp16 words; saturates like split_op16.
__device__ __forceinline__ void split_pair16_sat(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = a > 65504.f ? 65504.f : (a < -65504.f ? -65504.f : a);   // comparisons, not fmin / fmax: NaN propagates
  b = b > 65504.f ? 65504.f : (b < -65504.f ? -65504.f : b);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a - hf.x) * kSplitScale, (b - hf.y) * kSplitScale);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// programmatic dependent launch (every kernel of the library is launched with the attribute, launch_pdl / launch_tc):
// launch_dependents lets the NEXT kernel of the stream be scheduled as soon as every CTA of this one has started, so its
// launch latency and prologue run under this kernel's tail; wait() returns once the PREVIOUS kernel has completed and its
// writes are visible -- nothing a kernel reads or writes may be touched before it.  No-ops without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

// launch with programmatic stream serialization allowed (JATTS_B200_PDL=0 switches it off for A/B runs)
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem_bytes, cudaStream_t stream, Args... args) {
  static const int pdl = getenv("JATTS_B200_PDL") ? atoi(getenv("JATTS_B200_PDL")) : 1;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem_bytes;
  lc.stream = stream;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = la;
  lc.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kern, args...);
}

}  // namespace jb
