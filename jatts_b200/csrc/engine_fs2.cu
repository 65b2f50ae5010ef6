// FastSpeech2 text2mel engine: the batched, B200-native restatement of
// jatts/models/fastspeech2.py:566-653 `_forward(is_inference=True)` with per-utterance semantics
// (row i of the batch == reference inference(x_i); SURVEY.md finding 6).
//
// Precision: every dense contraction runs on tcgen05 tensor cores with bf16 hi/lo split operands
// (3 MMAs per K step, fp32 TMEM accumulation); activations between kernels are fp32 masters plus the
// hi/lo operand copies the next GEMM needs.  Attention core, LayerNorm, depthwise conv, predictors'
// tails and the length regulator are fp32 CUDA-core kernels.
#include "engine_fs2.cuh"

using namespace jb;

namespace jb {

static int load_ln(const WeightTable& wt, const std::string& n, int c, const float** g, const float** b) {
  JB_PROPAGATE(wt.f32(n + ".g", c, g));
  JB_PROPAGATE(wt.f32(n + ".b", c, b));
  return 0;
}

static int load_conformer(const WeightTable& wt, const std::string& pre, int n_layers, int d, int units, int ffn_k,
                          int dw_k, int max_len, int heads, ConformerW* out) {
  out->layers.resize(n_layers);
  for (int i = 0; i < n_layers; ++i) {
    ConformerLayerW& L = out->layers[i];
    const std::string p = pre + "." + std::to_string(i) + ".";
    JB_PROPAGATE(load_ln(wt, p + "ln_ffm", d, &L.ln_ffm_g, &L.ln_ffm_b));
    JB_PROPAGATE(load_ln(wt, p + "ln_mha", d, &L.ln_mha_g, &L.ln_mha_b));
    JB_PROPAGATE(load_ln(wt, p + "ln_conv", d, &L.ln_conv_g, &L.ln_conv_b));
    JB_PROPAGATE(load_ln(wt, p + "ln_ff", d, &L.ln_ff_g, &L.ln_ff_b));
    JB_PROPAGATE(load_ln(wt, p + "ln_final", d, &L.ln_fin_g, &L.ln_fin_b));
    JB_PROPAGATE(load_conv(wt, p + "ffm_w1", ffn_k, units, d, true, true, units, &L.ffm_w1));
    JB_PROPAGATE(load_conv(wt, p + "ffm_w2", ffn_k, d, units, true, true, d, &L.ffm_w2));
    JB_PROPAGATE(load_conv(wt, p + "ff_w1", ffn_k, units, d, true, true, units, &L.ff_w1));
    JB_PROPAGATE(load_conv(wt, p + "ff_w2", ffn_k, d, units, true, true, d, &L.ff_w2));
    JB_PROPAGATE(load_conv(wt, p + "qkv", 1, 4 * d, d, true, true, 4 * d, &L.qkv));   // [q+u | q+v | k | v]
    JB_PROPAGATE(load_conv(wt, p + "out", 1, d, d, true, true, d, &L.out));
    JB_PROPAGATE(load_conv(wt, p + "pw1", 1, 2 * d, d, true, true, d, &L.pw1));  // GLU: d outputs from 2d columns
    JB_PROPAGATE(load_conv(wt, p + "pw2", 1, d, d, true, true, d, &L.pw2));
    JB_PROPAGATE(wt.get(p + "pos.hi", JATTS_F16, static_cast<long long>(max_len) * d, reinterpret_cast<const void**>(&L.pos_hi)));
    JB_PROPAGATE(wt.get(p + "pos.lo", JATTS_F16, static_cast<long long>(max_len) * d, reinterpret_cast<const void**>(&L.pos_lo)));
    JB_PROPAGATE(wt.f32(p + "dw.wT", static_cast<long long>(dw_k) * d, &L.dw_wT));
    JB_PROPAGATE(wt.f32(p + "dw.b", d, &L.dw_b));
    L.dw_k = dw_k;
  }
  JB_PROPAGATE(load_ln(wt, pre + ".after_norm", d, &out->after_g, &out->after_b));
  (void)heads;
  return 0;
}

static int load_predictor(const WeightTable& wt, const std::string& pre, int n_layers, int chans, int k, int d,
                          PredictorW* out) {
  out->conv.resize(n_layers);
  out->ln_g.resize(n_layers);
  out->ln_b.resize(n_layers);
  out->chans = chans;
  for (int i = 0; i < n_layers; ++i) {
    const std::string p = pre + ".conv" + std::to_string(i);
    JB_PROPAGATE(load_conv(wt, p, k, chans, i == 0 ? d : chans, true, true, chans, &out->conv[i]));
    JB_PROPAGATE(load_ln(wt, pre + ".ln" + std::to_string(i), chans, &out->ln_g[i], &out->ln_b[i]));
  }
  JB_PROPAGATE(wt.f32(pre + ".lin_w", chans, &out->lin_w));
  const float* lb;
  JB_PROPAGATE(wt.f32(pre + ".lin_b", 1, &lb));
  JB_CUDA_OK(cudaMemcpy(&out->lin_b, lb, sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

RowLayout fs2_device_layout(const jatts_fs2* h, const HostLayout& hl, int phase) {
  RowLayout L;
  const int cu = h->cap_utt;
  const int* base = h->d_small + phase * 3 * cu;
  L.seg_start = base;
  L.seg_len = base + cu;
  L.frame_mask = h->mask;
  L.frame_seg = h->seg;
  L.nseg = static_cast<int>(hl.seg_len.size());
  L.n_rows = hl.n_rows;
  return L;
}

static int upload_layout(jatts_fs2* h, const HostLayout& hl, int phase, cudaStream_t s, RowLayout* L,
                         const int** d_off) {
  const int cu = h->cap_utt;
  const int n = static_cast<int>(hl.seg_len.size());
  int* hb = h->h_small + phase * 3 * cu;
  for (int i = 0; i < n; ++i) {
    hb[i] = hl.seg_start[i];
    hb[cu + i] = hl.seg_len[i];
    hb[2 * cu + i] = hl.off[i];
  }
  int* db = h->d_small + phase * 3 * cu;
  JB_CUDA_OK(cudaMemcpyAsync(db, hb, sizeof(int) * 3 * cu, cudaMemcpyHostToDevice, s));
  *L = fs2_device_layout(h, hl, phase);
  *d_off = db + 2 * cu;
  JB_PROPAGATE(fill_layout(L->seg_start, L->seg_len, L->nseg, L->n_rows, h->mask, h->seg, s));
  return 0;
}

static int ensure_workspace(jatts_fs2* h, int rows, int n_utt, int n_text) {
  const jatts_fs2_config& c = h->cfg;
  if (n_utt > h->cap_utt) {
    // the small-int tables live outside the arena so the arena can grow between plan and run
    const int cu = round_up(n_utt, 64);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->h_nframes) cudaFreeHost(h->h_nframes);
    if (h->d_small) cudaFree(h->d_small);
    if (h->d_nframes) cudaFree(h->d_nframes);
    h->h_small = nullptr; h->h_nframes = nullptr; h->d_small = nullptr; h->d_nframes = nullptr;
    JB_CUDA_OK(cudaMallocHost(&h->h_small, sizeof(int) * 6 * cu));
    JB_CUDA_OK(cudaMallocHost(&h->h_nframes, sizeof(int) * cu));
    JB_CUDA_OK(cudaMalloc(&h->d_small, sizeof(int) * 6 * cu));
    JB_CUDA_OK(cudaMalloc(&h->d_nframes, sizeof(int) * cu));
    h->cap_utt = cu;
  }
  (void)n_text;
  if (rows <= h->cap_rows) return 0;
  const int R = round_up(rows + 256, 1024);
  const int d = c.adim, u = std::max(c.eunits, c.dunits);
  const int pc = std::max(std::max(c.dur_chans, c.pitch_chans), c.energy_chans);
  const int od_pad = round_up(c.odim, 64), pn = round_up(c.postnet_chans, 64);
  size_t bytes = 0;
  auto f32 = [&](size_t cols) { bytes += Arena::padded(sizeof(float) * R * cols); };
  auto b16 = [&](size_t cols) { bytes += 2 * Arena::padded(sizeof(bf16) * R * cols); };
  f32(d); f32(d); b16(4 * d); f32(d); f32(pc); f32(c.odim); f32(c.odim);      // x hs q4 g pf before after
  for (int i = 0; i < 5; ++i) f32(1);                                          // s_dur s_pitch s_energy o_pitch o_energy
  b16(d); b16(u); b16(d); b16(pc); b16(od_pad); b16(pn); b16(pn);              // h t c p b pa pb
  bytes += Arena::padded(sizeof(long long) * R) + 2 * Arena::padded(sizeof(int) * R);  // o_dur cum lr_index
  bytes += Arena::padded(R) + Arena::padded(sizeof(int) * R);                  // mask seg
  JB_PROPAGATE(h->arena.reserve(bytes));
  Arena& a = h->arena;
  a.reset();
  h->x = a.take<float>(size_t(R) * d);
  h->hs = a.take<float>(size_t(R) * d);
  h->q4_hi = a.take<bf16>(size_t(R) * 4 * d);
  h->q4_lo = a.take<bf16>(size_t(R) * 4 * d);
  h->g = a.take<float>(size_t(R) * d);
  h->pf = a.take<float>(size_t(R) * pc);
  h->before = a.take<float>(size_t(R) * c.odim);
  h->after = a.take<float>(size_t(R) * c.odim);
  h->s_dur = a.take<float>(R); h->s_pitch = a.take<float>(R); h->s_energy = a.take<float>(R);
  h->o_pitch = a.take<float>(R); h->o_energy = a.take<float>(R);
  h->h_hi = a.take<bf16>(size_t(R) * d); h->h_lo = a.take<bf16>(size_t(R) * d);
  h->t_hi = a.take<bf16>(size_t(R) * u); h->t_lo = a.take<bf16>(size_t(R) * u);
  h->c_hi = a.take<bf16>(size_t(R) * d); h->c_lo = a.take<bf16>(size_t(R) * d);
  h->p_hi = a.take<bf16>(size_t(R) * pc); h->p_lo = a.take<bf16>(size_t(R) * pc);
  h->b_hi = a.take<bf16>(size_t(R) * od_pad); h->b_lo = a.take<bf16>(size_t(R) * od_pad);
  h->pa_hi = a.take<bf16>(size_t(R) * pn); h->pa_lo = a.take<bf16>(size_t(R) * pn);
  h->pb_hi = a.take<bf16>(size_t(R) * pn); h->pb_lo = a.take<bf16>(size_t(R) * pn);
  h->o_dur = a.take<long long>(R);
  h->cum = a.take<int>(R);
  h->lr_index = a.take<int>(R);
  h->mask = a.take<uint8_t>(R);
  h->seg = a.take<int>(R);
  h->cap_rows = R;
  return 0;
}

// zero the gap rows of every bf16 operand buffer for the current layout
static int zero_operand_gaps(jatts_fs2* h, const RowLayout& L, cudaStream_t s) {
  const jatts_fs2_config& c = h->cfg;
  const int d = c.adim, u = std::max(c.eunits, c.dunits);
  const int pc = std::max(std::max(c.dur_chans, c.pitch_chans), c.energy_chans);
  const int od_pad = round_up(c.odim, 64), pn = round_up(c.postnet_chans, 64);
  struct B { bf16* p; int cols; } bufs[] = {{h->h_hi, d}, {h->h_lo, d}, {h->t_hi, u}, {h->t_lo, u}, {h->c_hi, d},
                                            {h->c_lo, d}, {h->p_hi, pc}, {h->p_lo, pc}, {h->b_hi, od_pad},
                                            {h->b_lo, od_pad}, {h->pa_hi, pn}, {h->pa_lo, pn}, {h->pb_hi, pn},
                                            {h->pb_lo, pn}};
  void* ptrs[16];
  int bytes[16];
  int n = 0;
  for (auto& b : bufs) {
    ptrs[n] = b.p;
    bytes[n++] = b.cols * 2;
  }
  return zero_gap_rows_multi(ptrs, bytes, n, L, s);   // one launch instead of 14
}

// out = epilogue(conv(A)) with "same" padding over the packed layout
int split_conv(const ConvW& w, const bf16* a_hi, const bf16* a_lo, int a_ld, const RowLayout& L, ConvGemmEpilogue ep,
               cudaStream_t s, int dilation) {
  ConvGemmProblem p{};
  p.a_hi = a_hi; p.a_lo = a_lo; p.a_rows = L.n_rows; p.a_ld = a_ld;
  p.w_hi = w.hi; p.w_lo = w.lo; p.taps = w.taps; p.n_pad = w.n_pad; p.k_pad = w.k_pad;
  p.tap_off0 = -((w.taps - 1) / 2) * dilation; p.tap_stride = dilation;
  p.n = w.n; p.m_rows = L.n_rows; p.block_n = w.block_n;
  p.frame_mask = L.frame_mask; p.rate = 1; p.out_rows = L.n_rows;
  ep.bias = w.bias;
  if (ep.scale == 0.f) ep.scale = 1.f;
  if (ep.post_scale == 0.f) ep.post_scale = 1.f;
  p.ep = ep;
  return conv_gemm_tc(p, s);
}

static int conv_ffn(jatts_fs2* h, const ConvW& w1, const ConvW& w2, const RowLayout& L, cudaStream_t s) {
  const int d = h->cfg.adim, u = w1.n;
  ConvGemmEpilogue e1{};  // Conv1d -> ReLU (multi_layer_conv.py:62)
  e1.act = ACT_RELU; e1.out_hi = h->t_hi; e1.out_lo = h->t_lo; e1.out_bf_ld = u;
  JB_PROPAGATE(split_conv(w1, h->h_hi, h->h_lo, d, L, e1, s));
  ConvGemmEpilogue e2{};  // x = residual + 0.5 * ffn (encoder_layer.py:114-120, ff_scale)
  e2.scale = 0.5f; e2.res_f32 = h->x; e2.res_ld = d; e2.out_f32 = h->x; e2.out_f32_ld = d;
  JB_PROPAGATE(split_conv(w2, h->t_hi, h->t_lo, u, L, e2, s));
  return 0;
}

static int conformer_stack(jatts_fs2* h, const ConformerW& W, const RowLayout& L, int max_len, cudaStream_t s) {
  const jatts_fs2_config& c = h->cfg;
  const int d = c.adim;
  const size_t need = relpos_attention_scratch_bytes(max_len, L.nseg, c.aheads);
  if (need > h->att_scratch_bytes) {
    if (h->att_scratch) JB_CUDA_OK(cudaFree(h->att_scratch));   // cudaFree waits for the kernels still using it
    h->att_scratch = nullptr;
    h->att_scratch_bytes = 0;
    JB_CUDA_OK(cudaMalloc(&h->att_scratch, need));
    h->att_scratch_bytes = need;
  }
  const float eps = 1e-12f;  // layer_norm.py:23
  for (const ConformerLayerW& Lw : W.layers) {
    // macaron FFN (encoder_layer.py:112-122)
    JB_PROPAGATE(layernorm_rows(h->x, d, Lw.ln_ffm_g, Lw.ln_ffm_b, eps, L, nullptr, h->h_hi, h->h_lo, d, s));
    JB_PROPAGATE(conv_ffn(h, Lw.ffm_w1, Lw.ffm_w2, L, s));
    // self attention (encoder_layer.py:124-147, attention.py:164-206)
    JB_PROPAGATE(layernorm_rows(h->x, d, Lw.ln_mha_g, Lw.ln_mha_b, eps, L, nullptr, h->h_hi, h->h_lo, d, s));
    ConvGemmEpilogue eq{};
    eq.out_hi = h->q4_hi; eq.out_lo = h->q4_lo; eq.out_bf_ld = 4 * d;
    JB_PROPAGATE(split_conv(Lw.qkv, h->h_hi, h->h_lo, d, L, eq, s));
    JB_PROPAGATE(relpos_attention(h->q4_hi, h->q4_lo, h->cap_rows, Lw.pos_hi, Lw.pos_lo, c.max_len, c.aheads, d, L, max_len,
                                  h->att_scratch, h->att_scratch_bytes, h->c_hi, h->c_lo, d, s));
    ConvGemmEpilogue eo{};
    eo.res_f32 = h->x; eo.res_ld = d; eo.out_f32 = h->x; eo.out_f32_ld = d;
    JB_PROPAGATE(split_conv(Lw.out, h->c_hi, h->c_lo, d, L, eo, s));
    // convolution module (encoder_layer.py:149-156, convolution.py:56-79)
    JB_PROPAGATE(layernorm_rows(h->x, d, Lw.ln_conv_g, Lw.ln_conv_b, eps, L, nullptr, h->h_hi, h->h_lo, d, s));
    ConvGemmEpilogue eg{};
    eg.act = ACT_GLU; eg.out_f32 = h->g; eg.out_f32_ld = d;
    JB_PROPAGATE(split_conv(Lw.pw1, h->h_hi, h->h_lo, d, L, eg, s));
    JB_PROPAGATE(dwconv_swish(h->g, d, Lw.dw_wT, Lw.dw_b, Lw.dw_k, L, max_len, h->c_hi, h->c_lo, d, s));
    ConvGemmEpilogue e2{};
    e2.res_f32 = h->x; e2.res_ld = d; e2.out_f32 = h->x; e2.out_f32_ld = d;
    JB_PROPAGATE(split_conv(Lw.pw2, h->c_hi, h->c_lo, d, L, e2, s));
    // FFN (encoder_layer.py:158-168) and the block's final norm (:170-171)
    JB_PROPAGATE(layernorm_rows(h->x, d, Lw.ln_ff_g, Lw.ln_ff_b, eps, L, nullptr, h->h_hi, h->h_lo, d, s));
    JB_PROPAGATE(conv_ffn(h, Lw.ff_w1, Lw.ff_w2, L, s));
    JB_PROPAGATE(layernorm_rows(h->x, d, Lw.ln_fin_g, Lw.ln_fin_b, eps, L, h->x, nullptr, nullptr, d, s));
  }
  return 0;
}

// n x [Conv1d -> ReLU -> LayerNorm(channels)] -> Linear(-> 1); input operand = hs hi/lo in h_hi/h_lo
static int predictor(jatts_fs2* h, const PredictorW& W, const RowLayout& L, float* out, cudaStream_t s) {
  const int d = h->cfg.adim;
  const float eps = 1e-12f;
  const int n = static_cast<int>(W.conv.size());
  for (int i = 0; i < n; ++i) {
    ConvGemmEpilogue e{};
    e.act = ACT_RELU; e.out_f32 = h->pf; e.out_f32_ld = W.chans;
    if (i == 0) JB_PROPAGATE(split_conv(W.conv[i], h->h_hi, h->h_lo, d, L, e, s));
    else JB_PROPAGATE(split_conv(W.conv[i], h->p_hi, h->p_lo, W.chans, L, e, s));
    if (i + 1 < n)
      JB_PROPAGATE(layernorm_rows(h->pf, W.chans, W.ln_g[i], W.ln_b[i], eps, L, nullptr, h->p_hi, h->p_lo, W.chans, s));
    else
      JB_PROPAGATE(ln_dot_rows(h->pf, W.chans, W.ln_g[i], W.ln_b[i], eps, W.lin_w, W.lin_b, L, out, s));
  }
  return 0;
}

}  // namespace jb

namespace jb {
int fs2_create_impl(const jatts_fs2_config* cfg, const jatts_tensor* weights, int32_t n_weights, bool text_only,
                    jatts_fs2** out) {
  JB_REQUIRE(cfg && weights && out, JATTS_E_INVALID, "fs2_create: null argument");
  JB_REQUIRE(cfg->adim % 64 == 0 && cfg->adim <= 512, JATTS_E_UNSUPPORTED, "adim must be a multiple of 64, <= 512");
  JB_REQUIRE(cfg->adim % cfg->aheads == 0, JATTS_E_INVALID, "adim % aheads");
  JB_REQUIRE(relpos_attention_supported(cfg->aheads, cfg->adim), JATTS_E_UNSUPPORTED,
             "adim / aheads must be 64, 128, 192 or 256 (head size of the tcgen05 attention kernel)");
  JB_REQUIRE(cfg->eunits % 64 == 0 && cfg->dunits % 64 == 0, JATTS_E_UNSUPPORTED, "ffn units must be a multiple of 64");
  JB_REQUIRE(cfg->dur_chans % 64 == 0 && cfg->pitch_chans % 64 == 0 && cfg->energy_chans % 64 == 0 &&
                 cfg->dur_chans <= 512 && cfg->pitch_chans <= 512 && cfg->energy_chans <= 512,
             JATTS_E_UNSUPPORTED, "predictor channels must be multiples of 64, <= 512");
  JB_REQUIRE(text_only || (cfg->postnet_chans % 64 == 0 && cfg->postnet_layers >= 1), JATTS_E_UNSUPPORTED, "postnet channels % 64");
  JB_REQUIRE(cfg->max_len > 0 && cfg->max_len <= 5000, JATTS_E_INVALID, "max_len must be in (0, 5000]");
  JB_REQUIRE(cfg->dur_layers >= 1 && (text_only || (cfg->pitch_layers >= 1 && cfg->energy_layers >= 1)), JATTS_E_UNSUPPORTED,
             "predictors need >= 1 layer");
  // "same" convolutions over the packed layout rely on the kGapRows zero rows between utterances: an even kernel
  // would be centred differently from the reference's padding=(k-1)//2 and a half width above the gap would read
  // the neighbouring utterance (the depthwise convolution bounds its taps per utterance and has no such limit)
  for (int k : {cfg->ffn_kernel, cfg->dur_kernel, cfg->pitch_kernel, cfg->energy_kernel, cfg->postnet_filts})
    JB_REQUIRE(k >= 1 && (k & 1) == 1 && (k - 1) / 2 <= kGapRows, JATTS_E_UNSUPPORTED,
               "convolution kernel sizes must be odd with (k-1)/2 <= 8 (gap rows of the packed layout)");
  JB_REQUIRE((cfg->enc_cnn_kernel & 1) == 1 && (cfg->dec_cnn_kernel & 1) == 1, JATTS_E_UNSUPPORTED,
             "conformer depthwise kernel sizes must be odd");
  jatts_fs2* h = new jatts_fs2();
  h->cfg = *cfg;
  h->text_only = text_only;
  h->d_small = nullptr; h->d_nframes = nullptr;
  auto fail = [&](int rc) { delete h; return rc; };
  if (cudaGetDevice(&h->device) != cudaSuccess) { set_last_error("cudaGetDevice failed (no CUDA device?)"); return fail(JATTS_E_CUDA); }
  int rc = h->wt.init(weights, n_weights);
  if (rc) return fail(rc);
  const int d = cfg->adim;
  if ((rc = load_conformer(h->wt, "enc", cfg->elayers, d, cfg->eunits, cfg->ffn_kernel, cfg->enc_cnn_kernel, cfg->max_len, cfg->aheads, &h->enc))) return fail(rc);
  if (!text_only)
  if ((rc = load_conformer(h->wt, "dec", cfg->dlayers, d, cfg->dunits, cfg->ffn_kernel, cfg->dec_cnn_kernel, cfg->max_len, cfg->aheads, &h->dec))) return fail(rc);
  if ((rc = load_predictor(h->wt, "dur", cfg->dur_layers, cfg->dur_chans, cfg->dur_kernel, d, &h->dur))) return fail(rc);
  if (!text_only)
  if ((rc = load_predictor(h->wt, "pitch", cfg->pitch_layers, cfg->pitch_chans, cfg->pitch_kernel, d, &h->pitch))) return fail(rc);
  if (!text_only)
  if ((rc = load_predictor(h->wt, "energy", cfg->energy_layers, cfg->energy_chans, cfg->energy_kernel, d, &h->energy))) return fail(rc);
  if ((rc = h->wt.f32("emb", static_cast<long long>(cfg->idim) * d, &h->emb))) return fail(rc);
  h->pitch_w = h->pitch_b = h->energy_w = h->energy_b = nullptr;
  if (!text_only) {
    if ((rc = h->wt.f32("pitch_embed.w", d, &h->pitch_w))) return fail(rc);
    if ((rc = h->wt.f32("pitch_embed.b", d, &h->pitch_b))) return fail(rc);
    if ((rc = h->wt.f32("energy_embed.w", d, &h->energy_w))) return fail(rc);
    if ((rc = h->wt.f32("energy_embed.b", d, &h->energy_b))) return fail(rc);
  }
  h->spk_w = h->spk_b = nullptr;
  if (cfg->spk_embed_dim > 0) {
    if ((rc = h->wt.f32("spk.w", static_cast<long long>(d) * cfg->spk_embed_dim, &h->spk_w))) return fail(rc);
    if ((rc = h->wt.f32("spk.b", d, &h->spk_b))) return fail(rc);
  }
  if (text_only) {
    *out = h;
    return 0;
  }
  if ((rc = load_conv(h->wt, "feat_out", 1, cfg->odim, d, true, true, cfg->odim, &h->feat_out))) return fail(rc);
  h->postnet.resize(cfg->postnet_layers);
  for (int i = 0; i < cfg->postnet_layers; ++i) {
    const int ci = i == 0 ? cfg->odim : cfg->postnet_chans;
    const int co = i == cfg->postnet_layers - 1 ? cfg->odim : cfg->postnet_chans;
    if ((rc = load_conv(h->wt, "postnet" + std::to_string(i), cfg->postnet_filts, co, ci, true, true, co, &h->postnet[i]))) return fail(rc);
  }
  *out = h;
  return 0;
}
}  // namespace jb

extern "C" int jatts_fs2_create(const jatts_fs2_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                                jatts_fs2** out) {
  return fs2_create_impl(cfg, weights, n_weights, false, out);
}

extern "C" void jatts_fs2_destroy(jatts_fs2* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->arena.release();
  if (h->att_scratch) cudaFree(h->att_scratch);
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->h_nframes) cudaFreeHost(h->h_nframes);
  if (h->d_small) cudaFree(h->d_small);
  if (h->d_nframes) cudaFree(h->d_nframes);
  delete h;
}

extern "C" int jatts_fs2_plan(jatts_fs2* h, const int64_t* d_tokens, const int32_t* h_text_lens, int32_t n_utt,
                              const float* d_spembs, float alpha, int32_t* h_n_frames, void* stream) {
  JB_REQUIRE(h && d_tokens && h_text_lens && h_n_frames && n_utt > 0, JATTS_E_INVALID, "fs2_plan: bad argument");
  JB_REQUIRE(alpha > 0.f, JATTS_E_INVALID, "alpha must be > 0 (length_regulator.py:82)");
  JB_REQUIRE((h->cfg.spk_embed_dim > 0) == (d_spembs != nullptr), JATTS_E_INVALID,
             "spembs must be given iff the model has spk_embed_dim");
  JB_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  h->planned = false;
  for (int i = 0; i < n_utt; ++i)
    JB_REQUIRE(h_text_lens[i] > 0 && h_text_lens[i] <= h->cfg.max_len, JATTS_E_INVALID,
               "every utterance needs 1..max_len tokens");
  h->text.build(h_text_lens, n_utt);
  JB_PROPAGATE(ensure_workspace(h, std::max(h->text.n_rows, h->cap_rows), n_utt, h->text.total));
  h->n_utt = n_utt;
  h->alpha = alpha;
  const jatts_fs2_config& c = h->cfg;
  const int d = c.adim;
  JB_PROPAGATE(upload_layout(h, h->text, 0, s, &h->Lt, &h->d_text_off));
  const RowLayout& L = h->Lt;
  JB_PROPAGATE(zero_operand_gaps(h, L, s));
  // embedding * sqrt(D) (fastspeech2.py:270-272, positional_encoding.py:233)
  JB_PROPAGATE(embed_tokens(reinterpret_cast<const long long*>(d_tokens), h->d_text_off, h->emb, c.idim, d,
                            std::sqrt(static_cast<float>(d)), L, h->x, s));
  JB_PROPAGATE(conformer_stack(h, h->enc, L, h->text.max_len, s));
  // after_norm -> hs ; + speaker projection ; operand copies for the predictors
  JB_PROPAGATE(layernorm_rows(h->x, d, h->enc.after_g, h->enc.after_b, 1e-12f, L, h->hs,
                              d_spembs ? nullptr : h->h_hi, d_spembs ? nullptr : h->h_lo, d, s));
  if (d_spembs) {
    JB_PROPAGATE(add_speaker(d_spembs, c.spk_embed_dim, h->spk_w, h->spk_b, d, L, h->hs, s));
    JB_PROPAGATE(split_rows(h->hs, d, L, h->h_hi, h->h_lo, d, s));
  }
  if (!h->text_only) {
    JB_PROPAGATE(predictor(h, h->pitch, L, h->s_pitch, s));
    JB_PROPAGATE(predictor(h, h->energy, L, h->s_energy, s));
  }
  JB_PROPAGATE(predictor(h, h->dur, L, h->s_dur, s));
  JB_PROPAGATE(durations_and_scan(h->s_dur, alpha, L, h->d_text_off, h->o_dur, h->cum, h->d_nframes, s));
  if (!h->text_only) {
    JB_PROPAGATE(gather_scalar(h->s_pitch, L, h->d_text_off, h->o_pitch, s));
    JB_PROPAGATE(gather_scalar(h->s_energy, L, h->d_text_off, h->o_energy, s));
  }
  JB_CUDA_OK(cudaMemcpyAsync(h->h_nframes, h->d_nframes, sizeof(int) * n_utt, cudaMemcpyDeviceToHost, s));
  JB_CUDA_OK(cudaStreamSynchronize(s));
  for (int i = 0; i < n_utt; ++i) {
    h_n_frames[i] = h->h_nframes[i];
    JB_REQUIRE(h->h_nframes[i] >= 0 && h->h_nframes[i] <= c.max_len, JATTS_E_UNSUPPORTED,
               "an utterance expands to more frames than max_len");
  }
  h->frame.build(h->h_nframes, n_utt);
  h->planned = true;
  return 0;
}

extern "C" int jatts_fs2_run(jatts_fs2* h, float* d_mel, int64_t* d_durations, float* d_pitch, float* d_energy,
                             int32_t* d_lr_index, void* stream) {
  JB_REQUIRE(h && d_mel && d_durations && d_pitch && d_energy, JATTS_E_INVALID, "fs2_run: null argument");
  JB_REQUIRE(!h->text_only, JATTS_E_STATE, "fs2_run on a text-only engine");
  JB_REQUIRE(h->planned, JATTS_E_STATE, "fs2_run called without a successful fs2_plan");
  JB_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const jatts_fs2_config& c = h->cfg;
  const int d = c.adim;
  h->planned = false;
  const int text_total = h->text.total;
  JB_CUDA_OK(cudaMemcpyAsync(d_durations, h->o_dur, sizeof(long long) * text_total, cudaMemcpyDeviceToDevice, s));
  JB_CUDA_OK(cudaMemcpyAsync(d_pitch, h->o_pitch, sizeof(float) * text_total, cudaMemcpyDeviceToDevice, s));
  JB_CUDA_OK(cudaMemcpyAsync(d_energy, h->o_energy, sizeof(float) * text_total, cudaMemcpyDeviceToDevice, s));
  if (h->frame.total == 0) return 0;
  if (h->frame.n_rows > h->cap_rows) {
    // growing the arena would drop phase-1 state (hs, cum, pitch/energy): keep them across the move
    const int tr = h->text.n_rows;
    struct Keep {   // frees on every exit path (an early JB_CUDA_OK return used to leak the four buffers)
      float *hs = nullptr, *sp = nullptr, *se = nullptr; int* cum = nullptr;
      ~Keep() { cudaFree(hs); cudaFree(sp); cudaFree(se); cudaFree(cum); }
    } k;
    JB_CUDA_OK(cudaMalloc(&k.hs, sizeof(float) * tr * d));
    JB_CUDA_OK(cudaMalloc(&k.sp, sizeof(float) * tr));
    JB_CUDA_OK(cudaMalloc(&k.se, sizeof(float) * tr));
    JB_CUDA_OK(cudaMalloc(&k.cum, sizeof(int) * tr));
    JB_CUDA_OK(cudaMemcpyAsync(k.hs, h->hs, sizeof(float) * tr * d, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(k.sp, h->s_pitch, sizeof(float) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(k.se, h->s_energy, sizeof(float) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(k.cum, h->cum, sizeof(int) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaStreamSynchronize(s));   // the old arena is freed by the reserve below
    JB_PROPAGATE(ensure_workspace(h, h->frame.n_rows, h->n_utt, text_total));
    JB_CUDA_OK(cudaMemcpyAsync(h->hs, k.hs, sizeof(float) * tr * d, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(h->s_pitch, k.sp, sizeof(float) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(h->s_energy, k.se, sizeof(float) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaMemcpyAsync(h->cum, k.cum, sizeof(int) * tr, cudaMemcpyDeviceToDevice, s));
    JB_CUDA_OK(cudaStreamSynchronize(s));   // the temporaries are freed when `k` leaves scope
  }
  // The text-level layout tables (phase 0) stay valid; the text mask/seg arrays are about to be
  // overwritten by the frame layout, so the regulator only uses seg_start/seg_len of the text layout.
  RowLayout Lt = fs2_device_layout(h, h->text, 0);
  JB_PROPAGATE(upload_layout(h, h->frame, 1, s, &h->Lf, &h->d_frame_off));
  const RowLayout& L = h->Lf;
  JB_PROPAGATE(zero_operand_gaps(h, L, s));
  // LengthRegulator + pitch/energy embeddings + decoder input scaling (fastspeech2.py:614-617, encoder.py:138-141)
  JB_PROPAGATE(length_regulate(h->hs, h->s_pitch, h->s_energy, h->pitch_w, h->pitch_b, h->energy_w, h->energy_b, d,
                               std::sqrt(static_cast<float>(d)), Lt, h->cum, L, h->d_frame_off, h->x, h->lr_index, s));
  JB_PROPAGATE(conformer_stack(h, h->dec, L, h->frame.max_len, s));
  JB_PROPAGATE(layernorm_rows(h->x, d, h->dec.after_g, h->dec.after_b, 1e-12f, L, nullptr, h->h_hi, h->h_lo, d, s));
  // feat_out (fastspeech2.py:641-643): before (fp32, for the residual) + its operand copy for the postnet
  const int od_pad = round_up(c.odim, 64), pn = round_up(c.postnet_chans, 64);
  ConvGemmEpilogue ef{};
  ef.out_f32 = h->before; ef.out_f32_ld = c.odim; ef.out_hi = h->b_hi; ef.out_lo = h->b_lo; ef.out_bf_ld = od_pad;
  JB_PROPAGATE(split_conv(h->feat_out, h->h_hi, h->h_lo, d, L, ef, s));
  // postnet (pre_postnets.py:108-185), BatchNorm folded into the conv weights/bias on the host
  const bf16 *in_hi = h->b_hi, *in_lo = h->b_lo;
  int in_ld = od_pad;
  for (int i = 0; i < c.postnet_layers; ++i) {
    ConvGemmEpilogue e{};
    const bool last = i == c.postnet_layers - 1;
    bf16* o_hi = (i & 1) ? h->pb_hi : h->pa_hi;
    bf16* o_lo = (i & 1) ? h->pb_lo : h->pa_lo;
    if (!last) {
      e.act = ACT_TANH; e.out_hi = o_hi; e.out_lo = o_lo; e.out_bf_ld = pn;
    } else {
      e.res_f32 = h->before; e.res_ld = c.odim; e.out_f32 = h->after; e.out_f32_ld = c.odim;  // fastspeech2.py:649
    }
    JB_PROPAGATE(split_conv(h->postnet[i], in_hi, in_lo, in_ld, L, e, s));
    in_hi = o_hi; in_lo = o_lo; in_ld = pn;
  }
  JB_PROPAGATE(unpack_rows(h->after, c.odim, c.odim, L, h->d_frame_off, d_mel, s));
  if (d_lr_index)
    JB_CUDA_OK(cudaMemcpyAsync(d_lr_index, h->lr_index, sizeof(int) * h->frame.total, cudaMemcpyDeviceToDevice, s));
  return 0;
}
