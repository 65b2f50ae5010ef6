// Bandwidth-bound / small kernels of the synthesis path (CUDA cores, fp32).  See kernels.cu.
#pragma once
#include "common.cuh"

namespace jb {

// layout: fill frame_mask / frame_seg for n_rows rows from seg_start/seg_len (device arrays)
int fill_layout(const int* seg_start, const int* seg_len, int nseg, int n_rows, uint8_t* frame_mask, int* frame_seg,
                cudaStream_t s);
// zero the kGapRows rows that follow every utterance of up to 16 [n_rows, row_bytes] buffers (row_bytes % 16 == 0), one launch
int zero_gap_rows_multi(void* const* bufs, const int* row_bytes, int n, RowLayout L, cudaStream_t s);

// x[row,:] = emb[token,:] * scale for valid rows (fastspeech2.py:270-272 + positional_encoding.py:233)
int embed_tokens(const long long* tokens, const int* tok_off, const float* emb, int vocab, int d, float scale,
                 RowLayout L, float* x, cudaStream_t s);

// LayerNorm over the channel dim (layer_norm.py:12-42, eps 1e-12), valid rows only.
// outputs (each optional): y fp32, hi/lo bf16 operand pair.
int layernorm_rows(const float* x, int c, const float* gamma, const float* beta, float eps, RowLayout L,
                   float* y, bf16* hi, bf16* lo, int bf_ld, cudaStream_t s);

// fp32 -> bf16 hi/lo operand pair for valid rows
int split_rows(const float* x, int c, RowLayout L, bf16* hi, bf16* lo, int bf_ld, cudaStream_t s);

// predictor tail: LayerNorm(C) -> Linear(C -> 1) (duration_predictor.py:84, variance_predictor.py:81)
int ln_dot_rows(const float* x, int c, const float* gamma, const float* beta, float eps, const float* w, float b,
                RowLayout L, float* out, cudaStream_t s);

// legacy relative-position self attention core (attention.py:164-206) for every utterance / head, on tcgen05 tensor
// cores (attention_tc.cu).  x: [x_rows, 4*D] fp16 (hi, lo*2^11) pairs per row = [q + bias_u | q + bias_v | k | v] as
// the fused projection GEMM writes them; pos: [pos_rows, D] pair = linear_pos(pe); out: (hi, lo) pair [rows, out_ld].
// scratch: relpos_attention_scratch_bytes(max_len, nseg, n_head) bytes of device memory (per-CTA score rows).
bool relpos_attention_supported(int n_head, int d_model);
size_t relpos_attention_scratch_bytes(int max_len, int nseg, int n_head);
int relpos_attention(const bf16* x_hi, const bf16* x_lo, long long x_rows, const bf16* pos_hi, const bf16* pos_lo,
                     int pos_rows, int n_head, int d_model, RowLayout L, int max_len, float* scratch,
                     size_t scratch_bytes, bf16* out_hi, bf16* out_lo, int out_ld, cudaStream_t s);

// plain scaled-dot-product self attention on the same kernel: x = [q | k | v] ([x_rows, 3*D] operand pairs, no bias
// folding), softmax(q k^T / sqrt(d_k)) v per utterance and head (the Matcha decoder's transformer blocks)
int plain_attention(const bf16* x_hi, const bf16* x_lo, long long x_rows, int n_head, int d_model, RowLayout L, int max_len,
                    float* scratch, size_t scratch_bytes, bf16* out_hi, bf16* out_lo, int out_ld, cudaStream_t s);

// depthwise Conv1d (BatchNorm folded into wT/bias on the host) -> Swish (convolution.py:74-75)
// g: [rows, C] fp32 (GLU output), wT: [k][C], out: hi/lo bf16
// max_len >= every utterance length (sizes the strip grid)
int dwconv_swish(const float* g, int c, const float* wT, const float* bias, int k, RowLayout L, int max_len, bf16* out_hi,
                 bf16* out_lo, int out_ld, cudaStream_t s);

// hs[row,:] += proj(normalize(spemb[b])) (fastspeech2.py:737-761, "add")
int add_speaker(const float* spembs, int spk_dim, const float* w, const float* b, int d, RowLayout L, float* hs,
                cudaStream_t s);

// duration head (duration_predictor.py:86-95) + LengthRegulator bookkeeping (length_regulator.py:81-94).
//  logd: [rows] fp32 at text level.  dur_out: int64 [sum T_text] (utterance-contiguous, what inference() returns)
//  cum:  int32 [rows] inclusive prefix sum of the durations the regulator uses, per utterance
//  n_frames: int32 [nseg]
int durations_and_scan(const float* logd, float alpha, RowLayout L, const int* tok_off, long long* dur_out, int* cum,
                       int* n_frames, cudaStream_t s);

// scalar per-row copy into utterance-contiguous order: out[tok_off[b]+t] = in[row]
int gather_scalar(const float* in, RowLayout L, const int* tok_off, float* out, cudaStream_t s);

// LengthRegulator expansion as a prefix-sum gather, fused with "+ pitch_embed + energy_embed" and the
// decoder's x*sqrt(D) (fastspeech2.py:614-617, encoder.py:138-141):
//  x_out[frow,:] = (hs[trow,:] + p[trow]*wp + bp + e[trow]*we + be) * scale ; lr_index[frame_off[b]+f] = token
int length_regulate(const float* hs, const float* pitch, const float* energy, const float* wp, const float* bp,
                    const float* we, const float* be, int d, float scale, RowLayout Ltext, const int* cum,
                    RowLayout Lframe, const int* frame_off, float* x_out, int* lr_index, cudaStream_t s);

// rows -> utterance-contiguous fp32 [sum T, c] (and the reverse with an affine, to bf16, for the vocoder)
int unpack_rows(const float* in, int in_ld, int c, RowLayout L, const int* off, float* out, cudaStream_t s);
int pack_mel_affine(const float* mel, int c, const float* a, const float* b, RowLayout L, const int* off, bf16* out,
                    int out_ld, cudaStream_t s);

// HiFi-GAN output_conv: Conv1d(C -> 1, k) + tanh on an (already LeakyReLU'd) bf16 [rows*rate, ld] matrix.
// wave: fp32, utterance-contiguous [sum T*rate] and / or pcm: int16 = lrintf(wave * 32767) (either may be null)
int output_conv_tanh(const bf16* x, int ld, int c, const float* w /*[k][c]*/, float bias, int k, RowLayout L,
                     int rate, const int* frame_off, float* wave, short* pcm, cudaStream_t s);

// ---- Matcha-TTS flow-matching decoder (kernels_matcha.cu) ----
// y = Mish(GroupNorm_groups(x)) [+ add[c]] per utterance (decoder.py:65-96 Block1D / ResnetBlock1D); stats: scratch of
// nseg * groups float2; outputs (each optional): y fp32, hi/lo operand pair whose gap rows are written as zeros
int groupnorm_mish_rows(const float* x, int c, int groups, const float* gamma, const float* beta, float eps, const float* add,
                        RowLayout L, float2* stats, float* y, bf16* hi, bf16* lo, int bf_ld, cudaStream_t s);
// utterance-contiguous fp32 [sum T, c] * scale -> packed fp32 rows + operand pair (gap rows of the pair zeroed)
int pack_rows_split(const float* in, int c, float scale, RowLayout L, const int* off, float* y, int y_ld, bf16* hi, bf16* lo,
                    int bf_ld, cudaStream_t s);

}  // namespace jb
