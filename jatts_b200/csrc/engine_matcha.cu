// Matcha-TTS text2mel engine (BASELINE config 5): the batched, B200-native restatement of
// jatts/models/matchatts.py:390-480 `_forward(is_inference=True)` with per-utterance semantics (row i of the batch ==
// reference inference(x_i) given the same noise).
//
//   text side     = the FastSpeech2 engine's (engine_fs2.cu, text_only): Conformer encoder, duration predictor,
//                   LengthRegulator bookkeeping
//   mu            = encoder_proj(LengthRegulator(hs)) truncated to an even number of frames (matchatts.py:446-457)
//   CFM decoder   = jatts/modules/matchatts/flow_matching.py:48-95 fixed-step Euler over the 1-D U-Net of decoder.py:243-487:
//                   ResnetBlock1D (Conv1d k3 -> GroupNorm(8) -> Mish, + time embedding, + 1x1 residual conv),
//                   BasicTransformerBlock (LayerNorm, self attention, SnakeBeta feed-forward), stride-2 down-sampling,
//                   ConvTranspose1d(4, 2, 1) up-sampling, skip connections.
//
// Every dense contraction is the split-operand tcgen05 GEMM of conv_gemm_tc3.cu (fp32-faithful: the 10 Euler steps
// feed their own output back, bf16 operands do not hold 1e-3); attention is the plain mode of attention_tc.cu.
// Layout tricks on the packed-with-gaps rows (utterances start on even rows and have even lengths):
//   * the stride-2 convolution reads the activations as [rows / 2, 2C] (two frames per row): a 2-tap GEMM with K = 2C
//   * ConvTranspose1d(4, 2, 1) is a 3-tap GEMM with N = 2C whose output rows [rows / 2, 2C] ARE the up-sampled [rows, C]
//   * torch.cat([x, skip]) is a [rows, 2C] operand buffer whose halves are written by the producing GEMMs' epilogues
//   * the Euler update x += dt * v is the epilogue of the final 1x1 projection, which also emits the next step's operand
// The time embedding (sinusoid -> MLP -> per-block Linear of Mish) depends on the step index only; the host passes the
// [steps][resnet blocks][C] table (jatts_b200/matchatts.py computes it once per step count).
#include "engine_fs2.cuh"

namespace jb {

struct ResW {
  ConvW c1, c2, res;
  const float *g1, *b1, *g2, *b2;
};
struct TrW {
  const float *ln1_g, *ln1_b, *ln3_g, *ln3_b, *sn_a, *sn_ib;
  ConvW qkv, out, ff1, ff2;
};

}  // namespace jb

using namespace jb;

struct jatts_matcha {
  jatts_matcha_config cfg;
  int device = 0;
  jatts_fs2* core = nullptr;
  WeightTable wt;
  int C = 0, inner = 0, n_res = 0, in_ld = 0;
  ConvW enc_proj, down0, down1, up0, up1, fin_c, proj;
  const float *fin_g = nullptr, *fin_b = nullptr;
  std::vector<ResW> res;                 // down0, down1, mid..., up0, up1
  std::vector<std::vector<TrW>> tr;      // [resnet block][n_blocks]

  Arena arena;
  int cap_rows = 0, cap_utt = 0;
  float *xt, *lr_x, *A, *B, *X, *zeros;
  bf16 *in_hi, *in_lo, *p0_hi, *p0_lo, *p1_hi, *p1_lo, *p2_hi, *p2_lo, *qkv_hi, *qkv_lo, *ctx_hi, *ctx_lo, *ff_hi, *ff_lo,
      *cat0_hi, *cat0_lo, *cat1_hi, *cat1_lo;
  float2* stats;
  int* lr_index;
  uint8_t *mask, *mask_h;
  int *seg, *seg_h;
  int* d_small = nullptr;   // [seg_start | seg_len | off] full rate, [seg_start | seg_len] half rate
  int* h_small = nullptr;
  float* att_scratch = nullptr;
  size_t att_scratch_bytes = 0;

  bool planned = false;
  std::vector<int> frames;  // even-truncated frame counts of the planned batch
};

namespace jb {

static int load_ln2(const WeightTable& wt, const std::string& n, int c, const float** g, const float** b) {
  JB_PROPAGATE(wt.f32(n + ".g", c, g));
  JB_PROPAGATE(wt.f32(n + ".b", c, b));
  return 0;
}

static int ensure_workspace(jatts_matcha* h, int rows, int n_utt) {
  if (n_utt > h->cap_utt) {
    const int cu = round_up(n_utt, 64);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->d_small) cudaFree(h->d_small);
    h->h_small = nullptr; h->d_small = nullptr;
    JB_CUDA_OK(cudaMallocHost(&h->h_small, sizeof(int) * 5 * cu));
    JB_CUDA_OK(cudaMalloc(&h->d_small, sizeof(int) * 5 * cu));
    h->cap_utt = cu;
    h->cap_rows = 0;   // stats is sized by utterances
  }
  if (rows <= h->cap_rows) return 0;
  const int R = round_up(rows + 256, 1024), Rh = R / 2;
  const int C = h->C, inner = h->inner, d = h->core->cfg.adim, od = h->cfg.text.odim;
  const int pw = std::max(C, round_up(d, 64));
  size_t bytes = 0;
  auto f32 = [&](size_t rws, size_t cols) { bytes += Arena::padded(sizeof(float) * rws * cols); };
  auto b16 = [&](size_t rws, size_t cols) { bytes += 2 * Arena::padded(sizeof(bf16) * rws * cols); };
  f32(R, od); f32(R, d); f32(R, C); f32(R, C); f32(R, C); f32(std::max(R, d), 1);
  b16(R, h->in_ld); b16(R, pw); b16(R, C); b16(R, C); b16(R, 3 * inner); b16(R, inner); b16(R, 4 * C); b16(R, 2 * C); b16(Rh, 2 * C);
  bytes += Arena::padded(sizeof(float2) * h->cap_utt * 8) + Arena::padded(sizeof(int) * R);
  bytes += Arena::padded(R) + Arena::padded(Rh) + Arena::padded(sizeof(int) * R) + Arena::padded(sizeof(int) * Rh);
  JB_PROPAGATE(h->arena.reserve(bytes));   // zero-filled: `zeros` and the padding columns of `in` are never written
  Arena& a = h->arena;
  a.reset();
  h->xt = a.take<float>(size_t(R) * od);
  h->lr_x = a.take<float>(size_t(R) * d);
  h->A = a.take<float>(size_t(R) * C);
  h->B = a.take<float>(size_t(R) * C);
  h->X = a.take<float>(size_t(R) * C);
  h->zeros = a.take<float>(std::max(R, d));
  h->in_hi = a.take<bf16>(size_t(R) * h->in_ld); h->in_lo = a.take<bf16>(size_t(R) * h->in_ld);
  h->p0_hi = a.take<bf16>(size_t(R) * pw); h->p0_lo = a.take<bf16>(size_t(R) * pw);
  h->p1_hi = a.take<bf16>(size_t(R) * C); h->p1_lo = a.take<bf16>(size_t(R) * C);
  h->p2_hi = a.take<bf16>(size_t(R) * C); h->p2_lo = a.take<bf16>(size_t(R) * C);
  h->qkv_hi = a.take<bf16>(size_t(R) * 3 * inner); h->qkv_lo = a.take<bf16>(size_t(R) * 3 * inner);
  h->ctx_hi = a.take<bf16>(size_t(R) * inner); h->ctx_lo = a.take<bf16>(size_t(R) * inner);
  h->ff_hi = a.take<bf16>(size_t(R) * 4 * C); h->ff_lo = a.take<bf16>(size_t(R) * 4 * C);
  h->cat0_hi = a.take<bf16>(size_t(R) * 2 * C); h->cat0_lo = a.take<bf16>(size_t(R) * 2 * C);
  h->cat1_hi = a.take<bf16>(size_t(Rh) * 2 * C); h->cat1_lo = a.take<bf16>(size_t(Rh) * 2 * C);
  h->stats = a.take<float2>(size_t(h->cap_utt) * 8);
  h->lr_index = a.take<int>(R);
  h->mask = a.take<uint8_t>(R);
  h->mask_h = a.take<uint8_t>(Rh);
  h->seg = a.take<int>(R);
  h->seg_h = a.take<int>(Rh);
  h->cap_rows = R;
  return 0;
}

// [rows, C] pair -> one half of a [rows, 2C] concatenation buffer (gap rows included: they are zeros)
static int copy_into_cat(const bf16* src_hi, const bf16* src_lo, int C, int rows, bf16* dst_hi, bf16* dst_lo, int half, cudaStream_t s) {
  const size_t w = sizeof(bf16) * C;
  JB_CUDA_OK(cudaMemcpy2DAsync(dst_hi + half * C, 2 * w, src_hi, w, w, rows, cudaMemcpyDeviceToDevice, s));
  JB_CUDA_OK(cudaMemcpy2DAsync(dst_lo + half * C, 2 * w, src_lo, w, w, rows, cudaMemcpyDeviceToDevice, s));
  return 0;
}

// a convolution whose tap offsets are not centred (the paired-row views)
static int conv_taps(const ConvW& w, const bf16* a_hi, const bf16* a_lo, int a_ld, int tap_off0, const RowLayout& L,
                     ConvGemmEpilogue ep, cudaStream_t s) {
  ConvGemmProblem p{};
  p.a_hi = a_hi; p.a_lo = a_lo; p.a_rows = L.n_rows; p.a_ld = a_ld;
  p.w_hi = w.hi; p.w_lo = w.lo; p.taps = w.taps; p.n_pad = w.n_pad; p.k_pad = w.k_pad;
  p.tap_off0 = tap_off0; p.tap_stride = 1;
  p.n = w.n; p.m_rows = L.n_rows; p.block_n = w.block_n;
  p.frame_mask = L.frame_mask; p.rate = 1; p.out_rows = L.n_rows;
  ep.bias = w.bias;
  if (ep.scale == 0.f) ep.scale = 1.f;
  if (ep.post_scale == 0.f) ep.post_scale = 1.f;
  p.ep = ep;
  return conv_gemm_tc(p, s);
}

// decoder.py:79-96 ResnetBlock1D; input operand pair [rows, in_ld]; result: fp32 h->X
static int resnet_block(jatts_matcha* h, const ResW& W, const bf16* in_hi, const bf16* in_lo, int in_ld, const RowLayout& L,
                        const float* temb, cudaStream_t s) {
  const int C = h->C;
  const float eps = 1e-5f;   // torch.nn.GroupNorm default
  ConvGemmEpilogue e1{};
  e1.out_f32 = h->A; e1.out_f32_ld = C;
  JB_PROPAGATE(split_conv(W.c1, in_hi, in_lo, in_ld, L, e1, s));
  JB_PROPAGATE(groupnorm_mish_rows(h->A, C, 8, W.g1, W.b1, eps, temb, L, h->stats, nullptr, h->p1_hi, h->p1_lo, C, s));
  ConvGemmEpilogue e2{};
  e2.out_f32 = h->A; e2.out_f32_ld = C;
  JB_PROPAGATE(split_conv(W.c2, h->p1_hi, h->p1_lo, C, L, e2, s));
  JB_PROPAGATE(groupnorm_mish_rows(h->A, C, 8, W.g2, W.b2, eps, nullptr, L, h->stats, h->B, nullptr, nullptr, C, s));
  ConvGemmEpilogue e3{};
  e3.res_f32 = h->B; e3.res_ld = C; e3.out_f32 = h->X; e3.out_f32_ld = C;
  JB_PROPAGATE(split_conv(W.res, in_hi, in_lo, in_ld, L, e3, s));
  return 0;
}

// transformer.py:160-364 BasicTransformerBlock as configured by decoder.py:354-362 (pre-LN self attention + SnakeBeta
// feed-forward, the activation fused into the first Linear's epilogue) on h->X; the last GEMM also emits the operand pair of whatever follows (out_hi may be null)
static int transformer_block(jatts_matcha* h, const TrW& W, const RowLayout& L, int max_len, bf16* out_hi, bf16* out_lo,
                             int out_ld, cudaStream_t s) {
  const int C = h->C, inner = h->inner;
  const float eps = 1e-5f;   // torch.nn.LayerNorm default
  JB_PROPAGATE(layernorm_rows(h->X, C, W.ln1_g, W.ln1_b, eps, L, nullptr, h->p1_hi, h->p1_lo, C, s));
  ConvGemmEpilogue eq{};
  eq.out_hi = h->qkv_hi; eq.out_lo = h->qkv_lo; eq.out_bf_ld = 3 * inner;
  JB_PROPAGATE(split_conv(W.qkv, h->p1_hi, h->p1_lo, C, L, eq, s));
  JB_PROPAGATE(plain_attention(h->qkv_hi, h->qkv_lo, L.n_rows, h->cfg.n_heads, inner, L, max_len, h->att_scratch,
                               h->att_scratch_bytes, h->ctx_hi, h->ctx_lo, inner, s));
  ConvGemmEpilogue eo{};
  eo.res_f32 = h->X; eo.res_ld = C; eo.out_f32 = h->X; eo.out_f32_ld = C;
  JB_PROPAGATE(split_conv(W.out, h->ctx_hi, h->ctx_lo, inner, L, eo, s));
  JB_PROPAGATE(layernorm_rows(h->X, C, W.ln3_g, W.ln3_b, eps, L, nullptr, h->p1_hi, h->p1_lo, C, s));
  // SnakeBeta in the epilogue of its own Linear: the feed-forward's hidden layer never exists in fp32 (as a separate
  // pass it cost 3.4 ms per batch of 64; the fused epilogue with sin.approx adds 0.6 ms to the GEMMs)
  ConvGemmEpilogue ef{};
  ef.act = ACT_SNAKE; ef.snake_a = W.sn_a; ef.snake_ib = W.sn_ib;
  ef.out_hi = h->ff_hi; ef.out_lo = h->ff_lo; ef.out_bf_ld = 4 * C;
  JB_PROPAGATE(split_conv(W.ff1, h->p1_hi, h->p1_lo, C, L, ef, s));
  ConvGemmEpilogue e2{};
  e2.res_f32 = h->X; e2.res_ld = C; e2.out_f32 = h->X; e2.out_f32_ld = C;
  e2.out_hi = out_hi; e2.out_lo = out_lo; e2.out_bf_ld = out_ld;
  JB_PROPAGATE(split_conv(W.ff2, h->ff_hi, h->ff_lo, 4 * C, L, e2, s));
  return 0;
}

static int transformers(jatts_matcha* h, int r, const RowLayout& L, int max_len, bf16* out_hi, bf16* out_lo, int out_ld,
                        cudaStream_t s) {
  const int nb = static_cast<int>(h->tr[r].size());
  for (int j = 0; j < nb; ++j) {
    const bool last = j == nb - 1;
    JB_PROPAGATE(transformer_block(h, h->tr[r][j], L, max_len, last ? out_hi : nullptr, last ? out_lo : nullptr, out_ld, s));
  }
  return 0;
}

// decoder.py:413-487 Decoder.forward for one Euler step; x += dt * estimator(x, mu, t) lands in h->xt and in the
// first odim columns of the `in` operand
static int estimator_step(jatts_matcha* h, const RowLayout& L, const RowLayout& Lh, int max_len, const float* temb, float dt,
                          cudaStream_t s) {
  const int C = h->C, od = h->cfg.text.odim;
  const int nm = h->cfg.n_mid_blocks;
  int r = 0;
  // ---- down 0 (full rate): resnet(cat[x, mu]) -> transformers -> skip 0 + stride-2 convolution
  JB_PROPAGATE(resnet_block(h, h->res[r], h->in_hi, h->in_lo, h->in_ld, L, temb + r * C, s));
  JB_PROPAGATE(transformers(h, r, L, max_len, h->p0_hi, h->p0_lo, C, s));
  JB_PROPAGATE(copy_into_cat(h->p0_hi, h->p0_lo, C, L.n_rows, h->cat0_hi, h->cat0_lo, 1, s));
  {
    // Downsample1D Conv1d(k3, stride 2, padding 1): out[j] = W0 x[2j-1] + W1 x[2j] + W2 x[2j+1]; on rows viewed as
    // pairs [x[2j] | x[2j+1]] that is tap(-1) = [0 | W0], tap(0) = [W1 | W2]
    ConvGemmEpilogue e{};
    e.out_hi = h->p2_hi; e.out_lo = h->p2_lo; e.out_bf_ld = C;
    JB_PROPAGATE(conv_taps(h->down0, h->p0_hi, h->p0_lo, 2 * C, -1, Lh, e, s));
  }
  ++r;
  // ---- down 1 (half rate): resnet -> transformers -> skip 1 + Conv1d(k3)
  JB_PROPAGATE(resnet_block(h, h->res[r], h->p2_hi, h->p2_lo, C, Lh, temb + r * C, s));
  // skip 1 is written straight into the second half of cat1 (row pitch 2C) and the convolution reads it from there
  JB_PROPAGATE(transformers(h, r, Lh, max_len / 2, h->cat1_hi + C, h->cat1_lo + C, 2 * C, s));
  {
    ConvGemmEpilogue e{};
    e.out_hi = h->p2_hi; e.out_lo = h->p2_lo; e.out_bf_ld = C;
    JB_PROPAGATE(split_conv(h->down1, h->cat1_hi + C, h->cat1_lo + C, 2 * C, Lh, e, s));
  }
  ++r;
  // ---- mid blocks (half rate); the last one writes the first half of cat1
  for (int i = 0; i < nm; ++i, ++r) {
    JB_PROPAGATE(resnet_block(h, h->res[r], h->p2_hi, h->p2_lo, C, Lh, temb + r * C, s));
    const bool last = i == nm - 1;
    JB_PROPAGATE(transformers(h, r, Lh, max_len / 2, last ? h->cat1_hi : h->p2_hi, last ? h->cat1_lo : h->p2_lo, last ? 2 * C : C, s));
  }
  // ---- up 0 (half rate): resnet(cat[x, skip 1]) -> transformers -> ConvTranspose1d(4, 2, 1)
  JB_PROPAGATE(resnet_block(h, h->res[r], h->cat1_hi, h->cat1_lo, 2 * C, Lh, temb + r * C, s));
  JB_PROPAGATE(transformers(h, r, Lh, max_len / 2, h->p0_hi, h->p0_lo, C, s));
  {
    // out[2j] = W1^T x[j] + W3^T x[j-1], out[2j+1] = W2^T x[j] + W0^T x[j+1]: a 3-tap GEMM with N = 2C whose output row j
    // [out[2j] | out[2j+1]] is rows 2j, 2j+1 of the full-rate [rows, C] buffer
    ConvGemmEpilogue e{};
    e.out_hi = h->p2_hi; e.out_lo = h->p2_lo; e.out_bf_ld = 2 * C;
    JB_PROPAGATE(split_conv(h->up0, h->p0_hi, h->p0_lo, C, Lh, e, s));
    JB_PROPAGATE(copy_into_cat(h->p2_hi, h->p2_lo, C, L.n_rows, h->cat0_hi, h->cat0_lo, 0, s));
  }
  ++r;
  // ---- up 1 (full rate): resnet(cat[x, skip 0]) -> transformers -> Conv1d(k3)
  JB_PROPAGATE(resnet_block(h, h->res[r], h->cat0_hi, h->cat0_lo, 2 * C, L, temb + r * C, s));
  JB_PROPAGATE(transformers(h, r, L, max_len, h->p0_hi, h->p0_lo, C, s));
  {
    ConvGemmEpilogue e{};
    e.out_hi = h->p2_hi; e.out_lo = h->p2_lo; e.out_bf_ld = C;
    JB_PROPAGATE(split_conv(h->up1, h->p0_hi, h->p0_lo, C, L, e, s));
  }
  // ---- final Block1D + 1x1 projection; epilogue = the Euler update and the next step's operand
  {
    ConvGemmEpilogue e{};
    e.out_f32 = h->A; e.out_f32_ld = C;
    JB_PROPAGATE(split_conv(h->fin_c, h->p2_hi, h->p2_lo, C, L, e, s));
    JB_PROPAGATE(groupnorm_mish_rows(h->A, C, 8, h->fin_g, h->fin_b, 1e-5f, nullptr, L, h->stats, nullptr, h->p1_hi, h->p1_lo, C, s));
    ConvGemmEpilogue ep{};
    ep.scale = dt; ep.res_f32 = h->xt; ep.res_ld = od; ep.out_f32 = h->xt; ep.out_f32_ld = od;
    ep.out_hi = h->in_hi; ep.out_lo = h->in_lo; ep.out_bf_ld = h->in_ld;
    JB_PROPAGATE(split_conv(h->proj, h->p1_hi, h->p1_lo, C, L, ep, s));
  }
  return 0;
}

}  // namespace jb

extern "C" int jatts_matcha_create(const jatts_matcha_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                                   jatts_matcha** out) {
  JB_REQUIRE(cfg && weights && out, JATTS_E_INVALID, "matcha_create: null argument");
  JB_REQUIRE(cfg->n_channels == 2 && cfg->channels[0] == cfg->channels[1], JATTS_E_UNSUPPORTED,
             "decoder_channels must be two equal widths (one down-sampling stage, as every shipped recipe)");
  const int C = cfg->channels[0];
  JB_REQUIRE(C % 64 == 0 && C <= 512 && (C / 8) % 4 == 0, JATTS_E_UNSUPPORTED, "decoder width must be a multiple of 64, <= 512");
  JB_REQUIRE(cfg->n_blocks >= 1 && cfg->n_blocks <= 8 && cfg->n_mid_blocks >= 1 && cfg->n_mid_blocks <= 8, JATTS_E_UNSUPPORTED,
             "decoder_n_blocks / decoder_num_mid_blocks must be in 1..8");
  const int inner = cfg->n_heads * cfg->head_dim;
  JB_REQUIRE(cfg->n_heads >= 1 && relpos_attention_supported(cfg->n_heads, inner), JATTS_E_UNSUPPORTED,
             "decoder_attention_head_dim must be 64, 128, 192 or 256 (head size of the tcgen05 attention kernel)");
  JB_REQUIRE(cfg->text.odim % 8 == 0 && cfg->text.odim > 0, JATTS_E_UNSUPPORTED, "odim must be a multiple of 8");
  jatts_matcha* h = new jatts_matcha();
  h->cfg = *cfg;
  h->C = C;
  h->inner = inner;
  h->in_ld = round_up(2 * cfg->text.odim, 64);
  auto fail = [&](int rc) {
    if (h->core) jatts_fs2_destroy(h->core);
    delete h;
    return rc;
  };
  if (cudaGetDevice(&h->device) != cudaSuccess) { set_last_error("cudaGetDevice failed (no CUDA device?)"); return fail(JATTS_E_CUDA); }
  int rc = fs2_create_impl(&cfg->text, weights, n_weights, true, &h->core);
  if (rc) return fail(rc);
  if ((rc = h->wt.init(weights, n_weights))) return fail(rc);
  const WeightTable& wt = h->wt;
  const int d = cfg->text.adim, od = cfg->text.odim;
  if ((rc = load_conv(wt, "enc_proj", 1, od, d, true, true, od, &h->enc_proj))) return fail(rc);
  h->n_res = 4 + cfg->n_mid_blocks;
  h->res.resize(h->n_res);
  h->tr.resize(h->n_res);
  for (int r = 0; r < h->n_res; ++r) {
    const std::string p = "dec.res" + std::to_string(r);
    const bool up = r >= 2 + cfg->n_mid_blocks;
    const int cin = r == 0 ? 2 * od : (up ? 2 * C : C);
    ResW& W = h->res[r];
    if ((rc = load_conv(wt, p + ".c1", 3, C, cin, true, true, C, &W.c1))) return fail(rc);
    if ((rc = load_conv(wt, p + ".c2", 3, C, C, true, true, C, &W.c2))) return fail(rc);
    if ((rc = load_conv(wt, p + ".res", 1, C, cin, true, true, C, &W.res))) return fail(rc);
    if ((rc = load_ln2(wt, p + ".gn1", C, &W.g1, &W.b1))) return fail(rc);
    if ((rc = load_ln2(wt, p + ".gn2", C, &W.g2, &W.b2))) return fail(rc);
    h->tr[r].resize(cfg->n_blocks);
    for (int j = 0; j < cfg->n_blocks; ++j) {
      const std::string q = "dec.tr" + std::to_string(r) + "_" + std::to_string(j);
      TrW& T = h->tr[r][j];
      if ((rc = load_ln2(wt, q + ".ln1", C, &T.ln1_g, &T.ln1_b))) return fail(rc);
      if ((rc = load_ln2(wt, q + ".ln3", C, &T.ln3_g, &T.ln3_b))) return fail(rc);
      if ((rc = load_conv(wt, q + ".qkv", 1, 3 * inner, C, true, false, 3 * inner, &T.qkv))) return fail(rc);
      if ((rc = load_conv(wt, q + ".out", 1, C, inner, true, true, C, &T.out))) return fail(rc);
      if ((rc = load_conv(wt, q + ".ff1", 1, 4 * C, C, true, true, 4 * C, &T.ff1))) return fail(rc);
      if ((rc = load_conv(wt, q + ".ff2", 1, C, 4 * C, true, true, C, &T.ff2))) return fail(rc);
      if ((rc = wt.f32(q + ".snake.a", 4 * C, &T.sn_a))) return fail(rc);
      if ((rc = wt.f32(q + ".snake.ib", 4 * C, &T.sn_ib))) return fail(rc);
    }
  }
  if ((rc = load_conv(wt, "dec.down0", 2, C, 2 * C, true, true, C, &h->down0))) return fail(rc);
  if ((rc = load_conv(wt, "dec.down1", 3, C, C, true, true, C, &h->down1))) return fail(rc);
  if ((rc = load_conv(wt, "dec.up0", 3, 2 * C, C, true, true, 2 * C, &h->up0))) return fail(rc);
  if ((rc = load_conv(wt, "dec.up1", 3, C, C, true, true, C, &h->up1))) return fail(rc);
  if ((rc = load_conv(wt, "dec.final", 3, C, C, true, true, C, &h->fin_c))) return fail(rc);
  if ((rc = load_ln2(wt, "dec.final.gn", C, &h->fin_g, &h->fin_b))) return fail(rc);
  if ((rc = load_conv(wt, "dec.proj", 1, od, C, true, true, od, &h->proj))) return fail(rc);
  *out = h;
  return 0;
}

extern "C" void jatts_matcha_destroy(jatts_matcha* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->core) jatts_fs2_destroy(h->core);
  h->arena.release();
  if (h->att_scratch) cudaFree(h->att_scratch);
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->d_small) cudaFree(h->d_small);
  delete h;
}

extern "C" int32_t jatts_matcha_n_resnets(const jatts_matcha* h) { return h ? h->n_res : 0; }

extern "C" int jatts_matcha_plan(jatts_matcha* h, const int64_t* d_tokens, const int32_t* h_text_lens, int32_t n_utt,
                                 const float* d_spembs, int32_t* h_n_frames, void* stream) {
  JB_REQUIRE(h && h_n_frames, JATTS_E_INVALID, "matcha_plan: bad argument");
  h->planned = false;
  JB_PROPAGATE(jatts_fs2_plan(h->core, d_tokens, h_text_lens, n_utt, d_spembs, 1.0f, h_n_frames, stream));
  h->frames.resize(n_utt);
  for (int i = 0; i < n_utt; ++i) {
    // "since there is 2x upsampling in the decoder, truncate length to multiply of 2" (matchatts.py:453-455)
    h_n_frames[i] -= h_n_frames[i] % 2;
    h->frames[i] = h_n_frames[i];
  }
  h->planned = true;
  return 0;
}

extern "C" int jatts_matcha_run(jatts_matcha* h, const float* d_noise, float temperature, const float* d_temb,
                                const float* h_dt, int32_t n_steps, float* d_mel, int64_t* d_durations, void* stream) {
  JB_REQUIRE(h && d_noise && d_temb && h_dt && d_mel && d_durations && n_steps >= 1, JATTS_E_INVALID, "matcha_run: bad argument");
  JB_REQUIRE(h->planned && h->core->planned, JATTS_E_STATE, "matcha_run called without a successful matcha_plan");
  JB_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  jatts_fs2* core = h->core;
  h->planned = false;
  core->planned = false;
  const int n_utt = core->n_utt, d = core->cfg.adim, od = h->cfg.text.odim, C = h->C;
  JB_CUDA_OK(cudaMemcpyAsync(d_durations, core->o_dur, sizeof(long long) * core->text.total, cudaMemcpyDeviceToDevice, s));
  HostLayout fl;
  fl.build(h->frames.data(), n_utt);
  if (fl.total == 0) return 0;
  for (int i = 0; i < n_utt; ++i)
    JB_REQUIRE(h->frames[i] >= 2, JATTS_E_UNSUPPORTED,
               "an utterance expands to fewer than 2 frames (the reference's decoder output is empty for it)");
  JB_PROPAGATE(ensure_workspace(h, std::max(fl.n_rows, core->text.n_rows), n_utt));   // `zeros` is indexed by text rows too
  // ---- layouts: full rate and half rate (even starts and lengths: halves are exact)
  const int cu = h->cap_utt;
  for (int i = 0; i < n_utt; ++i) {
    h->h_small[i] = fl.seg_start[i];
    h->h_small[cu + i] = fl.seg_len[i];
    h->h_small[2 * cu + i] = fl.off[i];
    h->h_small[3 * cu + i] = fl.seg_start[i] / 2;
    h->h_small[4 * cu + i] = fl.seg_len[i] / 2;
  }
  JB_CUDA_OK(cudaMemcpyAsync(h->d_small, h->h_small, sizeof(int) * 5 * cu, cudaMemcpyHostToDevice, s));
  RowLayout L{}, Lh{};
  L.seg_start = h->d_small; L.seg_len = h->d_small + cu; L.frame_mask = h->mask; L.frame_seg = h->seg;
  L.nseg = n_utt; L.n_rows = fl.n_rows;
  Lh.seg_start = h->d_small + 3 * cu; Lh.seg_len = h->d_small + 4 * cu; Lh.frame_mask = h->mask_h; Lh.frame_seg = h->seg_h;
  Lh.nseg = n_utt; Lh.n_rows = fl.n_rows / 2;
  const int* d_off = h->d_small + 2 * cu;
  JB_PROPAGATE(fill_layout(L.seg_start, L.seg_len, L.nseg, L.n_rows, h->mask, h->seg, s));
  JB_PROPAGATE(fill_layout(Lh.seg_start, Lh.seg_len, Lh.nseg, Lh.n_rows, h->mask_h, h->seg_h, s));
  const size_t need = relpos_attention_scratch_bytes(fl.max_len, n_utt, h->cfg.n_heads);
  if (need > h->att_scratch_bytes) {
    if (h->att_scratch) JB_CUDA_OK(cudaFree(h->att_scratch));
    h->att_scratch = nullptr;
    h->att_scratch_bytes = 0;
    JB_CUDA_OK(cudaMalloc(&h->att_scratch, need));
    h->att_scratch_bytes = need;
  }
  // ---- mu = encoder_proj(LengthRegulator(hs)) (matchatts.py:426, 451): written as columns odim .. 2*odim of `in`
  RowLayout Lt = fs2_device_layout(core, core->text, 0);
  JB_PROPAGATE(length_regulate(core->hs, h->zeros, h->zeros, h->zeros, h->zeros, h->zeros, h->zeros, d, 1.0f, Lt, core->cum, L,
                               d_off, h->lr_x, h->lr_index, s));
  JB_PROPAGATE(split_rows(h->lr_x, d, L, h->p0_hi, h->p0_lo, round_up(d, 64), s));
  {
    ConvGemmEpilogue e{};
    e.out_hi = h->in_hi + od; e.out_lo = h->in_lo + od; e.out_bf_ld = h->in_ld;
    JB_PROPAGATE(split_conv(h->enc_proj, h->p0_hi, h->p0_lo, round_up(d, 64), L, e, s));
  }
  // ---- x_0 = z * temperature (flow_matching.py:64)
  JB_PROPAGATE(pack_rows_split(d_noise, od, temperature, L, d_off, h->xt, od, h->in_hi, h->in_lo, h->in_ld, s));
  // ---- fixed-step Euler (flow_matching.py:70-95)
  for (int step = 0; step < n_steps; ++step)
    JB_PROPAGATE(estimator_step(h, L, Lh, fl.max_len, d_temb + static_cast<size_t>(step) * h->n_res * C, h_dt[step], s));
  JB_PROPAGATE(unpack_rows(h->xt, od, od, L, d_off, d_mel, s));
  return 0;
}
