// Implicit-GEMM Conv1d / ConvTranspose1d / Linear on tcgen05 tensor cores (sm_100a).
//
//   out[row, n] = epilogue( sum_{tap j} sum_{c} A[row + tap_off0 + j*tap_stride, c] * W[j][n][c] )
//
// A is a channels-last activation matrix [rows, C_in_pad] in bf16 (optionally a hi/lo pair: "split"
// mode, 3 MMAs per K step = fp32-faithful products, SURVEY.md 8(a) precision budget), W is
// [taps][N_pad][C_in_pad] bf16 (hi/lo in split mode).  Row shifts between taps are TMA coordinates;
// per-utterance zero padding comes from the packed-with-gaps layout (common.cuh) and from TMA
// out-of-bounds zero fill at the ends of the buffer.
#pragma once
#include <vector>

#include "common.cuh"

namespace jb {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3, ACT_GLU = 4, ACT_SNAKE = 5 };

struct ConvGemmEpilogue {
  const float* bias;        // [N] (GLU: [N_pad] permuted like the weights) or null
  int act;                  // ACT_*
  float slope;              // for ACT_LRELU
  float scale;              // v = act(acc + bias) * scale
  const float* res_f32;     // v += res_f32[row*res_ld + n]
  const bf16* res_bf16;     // v += res_bf16[row*res_ld + n]
  float res_inv_slope;      // != 0: res_bf16 holds LeakyReLU(x, s) and x is recovered as r >= 0 ? r : r * res_inv_slope (= 1/s)
  int res_ld;
  const float* accum_in;    // v += accum_in[row*out_f32_ld + n]   (fp32 running sum)
  const bf16* accum_bf16;   // v += accum_bf16[row*res_ld + n]     (MRF branch sum kept in bf16)
  float post_scale;         // v *= post_scale
  float* out_f32;           // optional fp32 output
  int out_f32_ld;
  bf16* out_hi;             // optional bf16(v)
  bf16* out_lo;             // optional bf16(v - hi)
  int out_bf_ld;
  bf16* out_act;            // optional bf16(leaky_relu(v, out_act_slope))  (next conv's operand)
  float out_act_slope;
  int out_act_ld;
  // ACT_SNAKE (SnakeBeta, jatts/modules/matchatts/transformer.py:28-102): x + sin(x * snake_a[n])^2 * snake_ib[n], tables of
  // n_pad floats (a = exp(alpha), ib = 1 / (exp(beta) + 1e-9)); split GEMM with a single accumulation chain only
  const float* snake_a;
  const float* snake_ib;
};

struct ConvGemmProblem {
  // operands
  const bf16* a_hi;  const bf16* a_lo;   // [a_rows, a_ld]; a_lo null => plain bf16
  int a_rows, a_ld;
  int a_cols;                            // real channels of A (<= k_pad; TMA zero-fills the rest); 0 => k_pad
  const bf16* w_hi;  const bf16* w_lo;   // [taps*n_pad, k_pad]
  int taps, n_pad, k_pad;                // k_pad = C_in padded to 64
  int tap_off0, tap_stride;
  int n;                                 // real output columns (GLU: real outputs = n, weights hold 2n)
  int m_rows;                            // rows to compute (output row space before remap)
  int block_n;                           // 32 / 64 / 128 / 256
  // row validity (output side): rows are valid iff frame_mask[out_row / rate] != 0; null => all valid
  const uint8_t* frame_mask;
  int rate;                              // rows per mask entry on the OUTPUT side
  int out_rows;                          // bound on out_row
  // transposed-conv (polyphase) row remap: 0 => plain.  out_row = row*up_s + (n / up_cout) - up_p,
  // out_col = n % up_cout
  int up_s, up_p, up_cout;
  // One PHASE of a polyphase transposed convolution expressed as a plain 2-tap convolution for the
  // TMA-epilogue kernel (all zero = not used): weight rows of tap t start at w_row0 + t*w_tap_stride,
  // output row of GEMM row r is (r + store_row_off) in an output view whose row pitch is out_pitch_mul
  // times the tensor's (the caller offsets the out pointers to the view's first row and gives its row
  // count in out_view_rows), and row validity is frame_mask[(r*mask_mul + mask_add) / rate].
  int w_row0, w_tap_stride, w_rows_total;
  int store_row_off, out_pitch_mul, out_view_rows;
  int mask_mul, mask_add;
  // ALL phases of the polyphase transposed convolution in ONE launch (conv_gemm_tc2 only): phases > 0 = up-sampling factor s,
  // phase_pp = its padding p.  Phase q is the N tile q (block_n == n == C_out): weight rows w_row0 + t*w_tap_stride + q*n,
  // tap_off0 = q >= p ? -1 : 0, output rows first(q) + r*s with first(q) = q >= p ? q - p : q - p + s in a view of pitch s rows
  // (out_act = the tensor's base; out_pitch_mul = mask_mul = s; out_view_rows / mask_add / tap_off0 / store_row_off unused).
  // Units are walked M-tile major, so the s phases of an M tile run at the same time and its activation slab leaves HBM once.
  int phases, phase_pp;
  ConvGemmEpilogue ep;
};

// Launch the tcgen05 kernel (dispatches to the TMA-epilogue variant of conv_gemm_tc2.cu when the
// problem is a plain bf16 convolution with bf16 inputs/outputs).  Returns 0 / negative JATTS_E_*.
int conv_gemm_tc(const ConvGemmProblem& p, cudaStream_t stream);
bool conv_gemm_tc2_eligible(const ConvGemmProblem& p);
int conv_gemm_tc2(const ConvGemmProblem& p, cudaStream_t stream);
// fp16-split GEMMs (FastSpeech2) with the TMA-staged epilogue (conv_gemm_tc3.cu)
bool conv_gemm_tc3_eligible(const ConvGemmProblem& p);
int conv_gemm_tc3(const ConvGemmProblem& p, cudaStream_t stream);
// Plain CUDA-core kernel with identical semantics; used ONLY by the test entry point to bisect
// tensor-core kernel bugs from host-side packing bugs.  Never on the product path.
int conv_gemm_simt_debug(const ConvGemmProblem& p, cudaStream_t stream);

int num_sms();

// optional per-launch device timing (jatts_profile_begin / jatts_profile_end*): CUDA events recorded around every
// launch of a class of kernels on the launching stream; bench.py turns them into achieved TFLOP/s / GB/s
enum : int { PROF_BF16_CONV = 0, PROF_SPLIT_GEMM = 1, PROF_ATTENTION = 2, PROF_LAYERNORM = 3, PROF_DWCONV = 4,
             PROF_LENGTH_REGULATE = 5, PROF_OUTPUT_CONV = 6, PROF_CLASSES = 7 };
struct ProfileEvent { cudaEvent_t e0, e1; int split; };   // split = PROF_* class
extern bool g_profile_on;
extern std::vector<ProfileEvent> g_profile_events;
// RAII: e0 before the launch, e1 after it (no-op unless profiling is armed)
struct ProfileScope {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t s;
  int cls;
  ProfileScope(cudaStream_t stream, int c) : s(stream), cls(c) {
    if (!g_profile_on) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
    cudaEventRecord(e0, s);
  }
  ~ProfileScope() {
    if (!e0) return;
    cudaEventRecord(e1, s);
    g_profile_events.push_back({e0, e1, cls});
  }
};

}  // namespace jb
