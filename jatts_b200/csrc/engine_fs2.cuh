// Internal interface of the FastSpeech2 engine (engine_fs2.cu) for the engines that reuse its text side: the Matcha-TTS
// engine (engine_matcha.cu) runs the same Conformer encoder / duration predictor / LengthRegulator bookkeeping.
#pragma once
#include <cmath>

#include "engine_common.cuh"

namespace jb {

struct ConformerLayerW {
  const float *ln_ffm_g, *ln_ffm_b, *ln_mha_g, *ln_mha_b, *ln_conv_g, *ln_conv_b, *ln_ff_g, *ln_ff_b, *ln_fin_g,
      *ln_fin_b;
  ConvW ffm_w1, ffm_w2, ff_w1, ff_w2, qkv, out, pw1, pw2;
  const bf16 *pos_hi, *pos_lo;         // [max_len, D] fp16 (hi, lo*2^11) pair of linear_pos(pe); the biases u / v
                                       // are folded into the projection GEMM (its two biased copies of q)
  const float *dw_wT, *dw_b;           // BatchNorm folded
  int dw_k;
};
struct ConformerW {
  std::vector<ConformerLayerW> layers;
  const float *after_g, *after_b;
};
struct PredictorW {
  std::vector<ConvW> conv;
  std::vector<const float*> ln_g, ln_b;
  const float* lin_w;
  float lin_b;
  int chans;
};

}  // namespace jb

struct jatts_fs2 {
  jatts_fs2_config cfg;
  bool text_only = false;   // encoder + duration predictor only (Matcha-TTS): no decoder, pitch / energy, postnet weights
  int device = 0;
  jb::WeightTable wt;
  jb::ConformerW enc, dec;
  jb::PredictorW dur, pitch, energy;
  const float *emb, *pitch_w, *pitch_b, *energy_w, *energy_b, *spk_w, *spk_b;
  jb::ConvW feat_out;
  std::vector<jb::ConvW> postnet;

  jb::Arena arena;
  int cap_rows = 0, cap_utt = 0;
  // workspace views (valid after ensure_workspace)
  float *x, *hs, *g, *pf, *before, *after, *s_dur, *s_pitch, *s_energy, *o_pitch, *o_energy;
  jb::bf16 *q4_hi, *q4_lo;   // [rows, 4*D] projection output: q + u | q + v | k | v as operand pairs
  float* att_scratch = nullptr;   // per-CTA score rows of the attention kernel (own allocation, grown on demand)
  size_t att_scratch_bytes = 0;
  jb::bf16 *h_hi, *h_lo, *t_hi, *t_lo, *c_hi, *c_lo, *p_hi, *p_lo, *b_hi, *b_lo, *pa_hi, *pa_lo, *pb_hi, *pb_lo;
  long long* o_dur;
  int *cum, *lr_index, *d_nframes;
  uint8_t* mask;
  int *seg, *d_small;  // d_small: [seg_start | seg_len | off] x 2 phases
  int* h_small = nullptr;  // pinned staging, same shape
  int* h_nframes = nullptr;

  // state between plan and run
  bool planned = false;
  int n_utt = 0;
  float alpha = 1.0f;
  jb::HostLayout text, frame;
  jb::RowLayout Lt, Lf;
  const int* d_text_off = nullptr;
  const int* d_frame_off = nullptr;
};


namespace jb {
// jatts_fs2_create with the option above
int fs2_create_impl(const jatts_fs2_config* cfg, const jatts_tensor* weights, int32_t n_weights, bool text_only, jatts_fs2** out);
// device view of the text-level (phase 0) or frame-level (phase 1) layout uploaded by the last plan / run
RowLayout fs2_device_layout(const jatts_fs2* h, const HostLayout& hl, int phase);
// out = epilogue(conv(A)) with "same" padding over the packed layout, on the split-operand tcgen05 GEMM
int split_conv(const ConvW& w, const bf16* a_hi, const bf16* a_lo, int a_ld, const RowLayout& L, ConvGemmEpilogue ep,
               cudaStream_t s, int dilation = 1);
}  // namespace jb
