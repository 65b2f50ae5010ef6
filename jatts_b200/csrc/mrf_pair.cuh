// Fused HiFi-GAN residual unit (conv k,d -> LeakyReLU -> conv k,1 -> + x) for the narrow MRF stages; see mrf_pair.cu.
#pragma once
#include "common.cuh"

namespace jb {

struct MrfPairProblem {
  const bf16* xa;            // [rows, ld] bf16, LeakyReLU(x, slope) of the unit's input x
  int rows, ld;
  int c;                     // channels (32 or 64)
  const bf16* w1;            // [taps][n_pad][k_pad] bf16 (conv with dilation `dilation`)
  const bf16* w2;            // [taps][n_pad][k_pad] bf16 (conv with dilation 1)
  int taps, n_pad, k_pad, dilation;
  const float* h_b1;         // HOST copies of the two biases [c] (they travel in the kernel parameters)
  const float* h_b2;
  float slope;               // LeakyReLU slope between the convolutions, also the slope xa was stored with
  const uint8_t* frame_mask; // row validity: frame_mask[row / rate] (null = all rows valid)
  int rate;
  const bf16* accum;         // optional [rows, accum_ld]: v += accum (MRF branch sum)
  int accum_ld;
  float post_scale;          // v *= post_scale
  float out_slope;           // out = LeakyReLU(v, out_slope); 1 = identity
  bf16* out;                 // [rows, out_ld]; must not alias xa (other tiles read xa's halo rows)
  int out_ld;
};

bool mrf_pair_eligible(const MrfPairProblem& p);
int mrf_pair(const MrfPairProblem& p, cudaStream_t stream);

}  // namespace jb
