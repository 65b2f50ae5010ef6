/* jatts_b200 -- C ABI of the B200-native JATTS batched-synthesis path.
 *
 * This is the drop-in boundary for the reference's two hot-path calls:
 *
 *   jatts/models/fastspeech2.py:655-735   FastSpeech2.inference(text, spembs=..., alpha=...)
 *   jatts/vocoder/vocoder.py:56-67        Vocoder.decode(c)  (-> parallel_wavegan HiFiGANGenerator.inference)
 *
 * as they are called from jatts/bin/tts_decode.py:230,249 and jatts/trainers/fastspeech2.py:183,205.
 * The reference is pure Python; the host side that mirrors its classes lives in the Python package jatts_b200 and binds
 * these symbols with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (PyTorch); h_* is a HOST pointer.
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered.  The only host
 *     synchronisation is inside jatts_fs2_plan (the per-utterance frame counts must reach the host so
 *     the caller can size its outputs) -- the same point where the reference synchronises
 *     (length_regulator.py:86, repeat_interleave).
 *   - return value: 0 on success, a negative JATTS_E_* code otherwise; jatts_last_error() returns the
 *     message of the last failure on the calling thread.  There is no CPU fallback of any kind.
 *   - a handle is bound to the CUDA device that was current at create time and is not re-entrant.
 */
#ifndef JATTS_B200_H_
#define JATTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JATTS_B200_ABI_VERSION 1
#if defined(__GNUC__)
#define JATTS_API __attribute__((visibility("default")))
#else
#define JATTS_API
#endif

enum {
  JATTS_OK = 0,
  JATTS_E_UNSUPPORTED = -1, /* configuration outside what the kernels implement */
  JATTS_E_INVALID = -2,     /* bad argument / shape */
  JATTS_E_CUDA = -3,        /* CUDA runtime or driver failure (message has the code) */
  JATTS_E_STATE = -4        /* call order violated (run before plan, ...) */
};

enum { JATTS_F32 = 0, JATTS_BF16 = 1, JATTS_I64 = 2, JATTS_I32 = 3, JATTS_F16 = 4 };

/* a named, already-repacked weight tensor in device memory (the Python host side does the repacking:
 * BatchNorm folding, tap-major layout, fp16 (hi, lo*2^11) split pairs, channel padding) */
typedef struct {
  const char* name;
  const void* d_ptr;
  int64_t numel;
  int32_t dtype;
} jatts_tensor;

typedef struct jatts_fs2 jatts_fs2;         /* FastSpeech2 text2mel engine */
typedef struct jatts_hifigan jatts_hifigan; /* HiFi-GAN V1 generator engine */

JATTS_API int jatts_abi_version(void);
JATTS_API const char* jatts_last_error(void);
/* number of kernels this library has launched since load (all handles); bench.py reports the delta */
JATTS_API int64_t jatts_launch_count(void);

/* Per-launch device timing of the tcgen05 convolution kernel (CUDA events on the launching stream),
 * for bench.py's roofline leg.  begin() arms it; end() waits for the recorded launches and returns the
 * summed duration / launch count of the bf16 (HiFi-GAN) and split-fp16 (FastSpeech2) instantiations. */
JATTS_API int jatts_profile_begin(void);
JATTS_API int jatts_profile_end(double* ms_bf16, int64_t* n_bf16, double* ms_split, int64_t* n_split);

/* Same, per kernel class: ms[c] / n[c] for c = 0 bf16 convolutions, 1 split GEMMs, 2 attention, 3 LayerNorm (incl. the
 * predictors' LayerNorm + Linear tails), 4 depthwise conv + Swish, 5 length regulator, 6 output conv + tanh.
 * n_classes >= 7. */
JATTS_API int jatts_profile_end_classes(double* ms, int64_t* n, int32_t n_classes);

/* debug: per-role clock64 timeline of CTA 0 of the TMA-epilogue convolution kernel (d_buf: 5*8*64 int64, or NULL) */
JATTS_API int jatts_debug_set_trace(void* d_buf);

/* ---- FastSpeech2 (replaces jatts/models/fastspeech2.py:566-735 on the inference path) ------------ */
typedef struct {
  int32_t idim, odim, adim, aheads;
  int32_t elayers, eunits, dlayers, dunits;
  int32_t ffn_kernel;                       /* positionwise_conv_kernel_size */
  int32_t enc_cnn_kernel, dec_cnn_kernel;   /* conformer_{enc,dec}_kernel_size */
  int32_t dur_layers, dur_chans, dur_kernel;
  int32_t pitch_layers, pitch_chans, pitch_kernel;
  int32_t energy_layers, energy_chans, energy_kernel;
  int32_t postnet_layers, postnet_chans, postnet_filts;
  int32_t spk_embed_dim;                    /* 0 = no speaker conditioning ("add" integration otherwise) */
  int32_t max_len;                          /* rows of the precomputed linear_pos(pe) tables (<= 5000) */
} jatts_fs2_config;

JATTS_API int jatts_fs2_create(const jatts_fs2_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                     jatts_fs2** out);
JATTS_API void jatts_fs2_destroy(jatts_fs2* h);

/* Phase 1: encoder + variance adaptor + durations.  d_tokens: int64 [sum(h_text_lens)], utterances
 * back to back.  d_spembs: fp32 [n_utt, spk_embed_dim] or NULL.  Writes the number of mel frames of
 * every utterance to h_n_frames[n_utt] (host) and returns after the stream has drained. */
JATTS_API int jatts_fs2_plan(jatts_fs2* h, const int64_t* d_tokens, const int32_t* h_text_lens, int32_t n_utt,
                   const float* d_spembs, float alpha, int32_t* h_n_frames, void* stream);
/* Phase 2: LengthRegulator + decoder + postnet for the batch planned last.  Outputs, utterances back
 * to back: d_mel fp32 [sum frames, odim]; d_durations int64 [sum text]; d_pitch, d_energy fp32
 * [sum text]; d_lr_index int32 [sum frames] (token index each frame was expanded from; may be NULL). */
JATTS_API int jatts_fs2_run(jatts_fs2* h, float* d_mel, int64_t* d_durations, float* d_pitch, float* d_energy,
                  int32_t* d_lr_index, void* stream);

/* ---- Matcha-TTS (replaces jatts/models/matchatts.py:390-560 on the inference path; BASELINE config 5) ---------
 * Text side = the FastSpeech2 encoder / duration predictor / LengthRegulator (`text`: idim, odim, adim, aheads, elayers,
 * eunits, ffn_kernel, enc_cnn_kernel, dur_*, spk_embed_dim, max_len are used, the remaining fields only have to be valid);
 * decoder = the conditional-flow-matching U-Net of jatts/modules/matchatts/{flow_matching,decoder,transformer}.py. */
typedef struct jatts_matcha jatts_matcha;
typedef struct {
  jatts_fs2_config text;
  int32_t n_channels;                       /* len(decoder_channels): 2 */
  int32_t channels[4];                      /* decoder_channels */
  int32_t n_blocks, n_mid_blocks;           /* decoder_n_blocks, decoder_num_mid_blocks */
  int32_t n_heads, head_dim;                /* decoder_num_heads, decoder_attention_head_dim */
} jatts_matcha_config;
JATTS_API int jatts_matcha_create(const jatts_matcha_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                        jatts_matcha** out);
JATTS_API void jatts_matcha_destroy(jatts_matcha* h);
/* ResnetBlock1D count of the decoder (2 down + mid + 2 up): the second dimension of the time-embedding table below */
JATTS_API int32_t jatts_matcha_n_resnets(const jatts_matcha* h);
/* Phase 1: encoder + duration predictor (matchatts.py:404-427).  h_n_frames[n_utt] (host) receives the number of mel
 * frames of every utterance AFTER the truncation to an even length of matchatts.py:453-455. */
JATTS_API int jatts_matcha_plan(jatts_matcha* h, const int64_t* d_tokens, const int32_t* h_text_lens, int32_t n_utt,
                      const float* d_spembs, int32_t* h_n_frames, void* stream);
/* Phase 2: LengthRegulator + encoder_proj + n_steps fixed Euler steps of the flow-matching decoder
 * (flow_matching.py:48-95).  d_noise: fp32 [sum frames, odim] standard-normal z (the reference draws it with
 * torch.randn_like inside CFM.inference; the caller draws it so that a run is reproducible), scaled by `temperature`.
 * d_temb: fp32 [n_steps, n_resnets, C]: mlp_r(mish(time_mlp(sinusoid(t_step)))) of every ResnetBlock1D (decoder.py:91,
 * :421-422): depends on the step only.  h_dt: host float[n_steps] step widths.  Outputs: d_mel fp32 [sum frames, odim],
 * d_durations int64 [sum text]. */
JATTS_API int jatts_matcha_run(jatts_matcha* h, const float* d_noise, float temperature, const float* d_temb,
                     const float* h_dt, int32_t n_steps, float* d_mel, int64_t* d_durations, void* stream);

/* ---- HiFi-GAN generator (replaces parallel_wavegan HiFiGANGenerator.inference behind vocoder.py:64) */
typedef struct {
  int32_t in_channels, out_channels, channels, kernel_size;
  int32_t n_upsamples;
  int32_t upsample_scales[8];
  int32_t n_resblocks;                      /* residual blocks per stage (3) */
  int32_t resblock_kernels[8];
  int32_t n_dilations;                      /* dilations per residual block (3) */
  int32_t resblock_dilations[8][8];
  float lrelu_slope;                        /* 0.1 */
} jatts_hifigan_config;

JATTS_API int jatts_hifigan_create(const jatts_hifigan_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                         jatts_hifigan** out);
JATTS_API void jatts_hifigan_destroy(jatts_hifigan* h);
/* d_mel: fp32 [sum(h_mel_lens), in_channels] utterances back to back, in the caller's normalisation;
 * the per-bin affine  c*scale + shift  of vocoder.py:57-61 is applied on load (weights "mel_scale",
 * "mel_shift").  d_wave: fp32 [sum(h_mel_lens) * hop]. */
JATTS_API int jatts_hifigan_run(jatts_hifigan* h, const float* d_mel, const int32_t* h_mel_lens, int32_t n_utt,
                      float* d_wave, void* stream);

/* Same generator with the output written as PCM_16 samples, int16 = lrintf(wave * 32767) -- what libsndfile stores
 * when jatts/bin/tts_decode.py:250-255 calls sf.write(path, y, sr, "PCM_16"); the conversion is fused into the
 * output convolution so the batched front-end (jatts_b200/decode.py) copies 2 bytes per sample to the host. */
JATTS_API int jatts_hifigan_run_pcm16(jatts_hifigan* h, const float* d_mel, const int32_t* h_mel_lens, int32_t n_utt,
                            int16_t* d_pcm, void* stream);

/* ---- op-level entry points used by the parity tests (tests/test_ops_gpu.py) ------------------------ */
typedef struct {
  const void* d_a_hi; const void* d_a_lo;     /* [a_rows, a_ld]; bf16 if a_lo NULL, else fp16 (hi, lo*2^11) pair */
  int32_t a_rows, a_ld, a_cols;                 /* a_cols: real channels (<= k_pad), 0 = k_pad */
  const void* d_w_hi; const void* d_w_lo;     /* bf16 [taps*n_pad, k_pad] */
  int32_t taps, n_pad, k_pad, tap_off0, tap_stride;
  int32_t n, m_rows, block_n;
  const uint8_t* d_frame_mask; int32_t rate, out_rows;
  int32_t up_s, up_p, up_cout;
  const float* d_bias; int32_t act; float slope, scale;
  const float* d_res_f32; const void* d_res_bf16; int32_t res_ld;
  const float* d_accum_in; const void* d_accum_bf16; float post_scale;
  float* d_out_f32; int32_t out_f32_ld;
  void* d_out_hi; void* d_out_lo; int32_t out_bf_ld;
  void* d_out_act; float out_act_slope; int32_t out_act_ld;
  /* act == 5 (SnakeBeta, jatts/modules/matchatts/transformer.py:28-102, after its own Linear): x + sin(x * a[n])^2 * ib[n];
   * n_pad floats each, a = exp(alpha), ib = 1 / (exp(beta) + 1e-9).  Split GEMMs with K <= 512 only. */
  const float* d_snake_a; const float* d_snake_ib;
} jatts_conv_gemm_args;
/* impl 0 = tcgen05 kernel (the product kernel), 1 = CUDA-core twin (test-only cross-check) */
JATTS_API int jatts_op_conv_gemm(const jatts_conv_gemm_args* a, int32_t impl, void* stream);

/* Fused HiFi-GAN residual unit of the narrow MRF stages (C = 32 / 64), the product kernel behind
 * jatts_hifigan_run for those stages (jatts_b200/csrc/mrf_pair.cu):
 *   t = lrelu(conv(xa, w1, dilation) + b1);  v = conv(t, w2, 1) + b2 + x [+ accum];  out = lrelu(v * post_scale, out_slope)
 * with xa = lrelu(x, slope) as stored and x recovered from it. */
typedef struct {
  const void* d_xa; int32_t rows, ld, c;
  const void* d_w1; const void* d_w2;           /* bf16 [taps][n_pad][k_pad] */
  int32_t taps, n_pad, k_pad, dilation;
  const float* d_b1; const float* d_b2;
  float slope;
  const uint8_t* d_frame_mask; int32_t rate;
  const void* d_accum; int32_t accum_ld;
  float post_scale, out_slope;
  void* d_out; int32_t out_ld;
} jatts_mrf_pair_args;
JATTS_API int jatts_op_mrf_pair(const jatts_mrf_pair_args* a, void* stream);

/* Legacy relative-position self-attention core (jatts/modules/transformer/attention.py:164-206 between the input
 * projections and linear_out), the product kernel behind jatts_fs2_plan / jatts_fs2_run (jatts_b200/csrc/attention_tc.cu).
 * x: [x_rows, 4*d_model] fp16 (hi, lo*2^11) pairs, per row [q + pos_bias_u | q + pos_bias_v | k | v]; utterance i owns
 * rows seg_start[i] .. seg_start[i] + seg_len[i] - 1.  pos: [pos_rows, d_model] pair of linear_pos(pe).
 * out: (hi, lo*2^11) pair of softmax(scores / sqrt(d_k)) . v, [x_rows, out_ld].  max_len >= every seg_len. */
typedef struct {
  const void* d_x_hi; const void* d_x_lo; int64_t x_rows;
  const void* d_pos_hi; const void* d_pos_lo; int32_t pos_rows;
  int32_t n_head, d_model;
  const int32_t* d_seg_start; const int32_t* d_seg_len; int32_t nseg, max_len;
  void* d_out_hi; void* d_out_lo; int32_t out_ld;
} jatts_relpos_attention_args;
JATTS_API int jatts_op_relpos_attention(const jatts_relpos_attention_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JATTS_B200_H_ */
