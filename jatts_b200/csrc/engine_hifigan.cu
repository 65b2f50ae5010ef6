// HiFi-GAN V1 generator engine: batched, B200-native restatement of
// parallel_wavegan HiFiGANGenerator.inference() as reached from jatts/vocoder/vocoder.py:64
// (architecture: oracle/hifigan.py header).  All Conv1d / ConvTranspose1d layers are tcgen05
// implicit GEMMs over the packed-with-gaps channels-last layout with bf16 operands and fp32 TMEM
// accumulation; LeakyReLU, bias, residual add, the MRF branch mean and the next layer's operand
// conversion are fused into the epilogues.  ConvTranspose1d(k = 2s, stride s) is a 2-tap polyphase
// GEMM with N = s * C_out whose epilogue scatters phase q of input row j to output row j*s + q - p.
#include <cstdlib>

#include "engine_common.cuh"
#include "mrf_pair.cuh"

using namespace jb;

struct jatts_hifigan {
  jatts_hifigan_config cfg;
  int device = 0;
  WeightTable wt;
  ConvW input_conv;
  std::vector<ConvW> ups, ups3;   // ups3: narrow stages as one 3-tap convolution (hi == null: not available)
  std::vector<std::vector<std::vector<ConvW>>> c1, c2;  // [stage][block][dilation]
  std::vector<std::vector<float>> host_bias;            // host copies of the residual-unit biases (ConvW::h_bias)
  const float *out_w, *mel_scale, *mel_shift;
  float out_b = 0.f;
  int hop = 1;
  std::vector<int> rate, chans;  // per stage output rate / channels

  Arena arena;
  int cap_rows = 0, cap_utt = 0;
  bf16 *mel, *xa0, *x, *xa, *t, *y[2], *sum;   // xa0 / x / xa hold lrelu(x) of the stage input and of the units' outputs
  uint8_t* mask;
  int *seg, *d_small = nullptr, *h_small = nullptr;
  cudaEvent_t staged = nullptr;
  bool staged_pending = false;
};

namespace jb {

static int ensure_workspace(jatts_hifigan* h, int rows, int n_utt) {
  if (n_utt > h->cap_utt) {
    const int cu = round_up(n_utt, 64);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->d_small) cudaFree(h->d_small);
    h->h_small = nullptr; h->d_small = nullptr;
    JB_CUDA_OK(cudaMallocHost(&h->h_small, sizeof(int) * 3 * cu));
    JB_CUDA_OK(cudaMalloc(&h->d_small, sizeof(int) * 3 * cu));
    h->cap_utt = cu;
  }
  if (rows <= h->cap_rows) return 0;
  const jatts_hifigan_config& c = h->cfg;
  const size_t R = round_up(rows + 64, 256);
  size_t per_row = static_cast<size_t>(c.channels);  // stage input of stage 0: [rows, channels]
  for (size_t i = 0; i < h->rate.size(); ++i) per_row = std::max(per_row, static_cast<size_t>(h->rate[i]) * h->chans[i]);
  const size_t in_pad = round_up(c.in_channels, 64);
  size_t bytes = Arena::padded(2 * R * in_pad) + 7 * Arena::padded(2 * R * per_row) +
                 Arena::padded(R) + Arena::padded(4 * R);
  JB_PROPAGATE(h->arena.reserve(bytes));
  Arena& a = h->arena;
  a.reset();
  h->mel = a.take<bf16>(R * in_pad);
  h->xa0 = a.take<bf16>(R * per_row);
  h->x = a.take<bf16>(R * per_row); h->xa = a.take<bf16>(R * per_row);
  h->t = a.take<bf16>(R * per_row);
  h->y[0] = a.take<bf16>(R * per_row); h->y[1] = a.take<bf16>(R * per_row);
  h->sum = a.take<bf16>(R * per_row);
  h->mask = a.take<uint8_t>(R);
  h->seg = a.take<int>(R);
  h->cap_rows = static_cast<int>(R);
  return 0;
}

struct StageIO {
  const uint8_t* mask;
  int rate;       // rows per mel frame of the tensors this conv reads and writes
  long long rows; // rows at that rate
};

static int run_conv(const ConvW& w, const bf16* a, int a_cols, const StageIO& io, int dilation, ConvGemmEpilogue ep,
                    cudaStream_t s) {
  ConvGemmProblem p{};
  p.a_hi = a; p.a_lo = nullptr; p.a_rows = static_cast<int>(io.rows); p.a_ld = a_cols; p.a_cols = a_cols;
  p.w_hi = w.hi; p.w_lo = nullptr; p.taps = w.taps; p.n_pad = w.n_pad; p.k_pad = w.k_pad;
  p.tap_off0 = -((w.taps - 1) / 2) * dilation; p.tap_stride = dilation;
  p.n = w.n; p.m_rows = static_cast<int>(io.rows); p.block_n = w.block_n;
  static const int bn_cap = getenv("JATTS_B200_BN_CAP") ? atoi(getenv("JATTS_B200_BN_CAP")) : 128;
  if (w.taps > 1 && p.block_n > bn_cap && w.n_pad % bn_cap == 0) p.block_n = bn_cap;
  p.frame_mask = io.mask; p.rate = io.rate; p.out_rows = static_cast<int>(io.rows);
  ep.bias = w.bias;
  if (ep.scale == 0.f) ep.scale = 1.f;
  if (ep.post_scale == 0.f) ep.post_scale = 1.f;
  p.ep = ep;
  // the engine relies on the TMA epilogue storing every row (gap rows as zeros): no silent fallback
  JB_REQUIRE(conv_gemm_tc2_eligible(p), JATTS_E_UNSUPPORTED, "convolution not eligible for the TMA-epilogue kernel");
  return conv_gemm_tc2(p, s);
}

// One residual unit x + conv(k,1)(lrelu(conv(k,d)(lrelu(x)))) of a narrow stage in one launch (mrf_pair.cu):
// only lrelu(x) is kept in HBM, the kernel recovers x from it.
static MrfPairProblem pair_problem(const ConvW& w1, const ConvW& w2, const bf16* xa, int c, const StageIO& io, int dilation,
                                   float slope) {
  MrfPairProblem p{};
  p.xa = xa; p.rows = static_cast<int>(io.rows); p.ld = c; p.c = c;
  p.w1 = w1.hi; p.w2 = w2.hi; p.taps = w1.taps; p.n_pad = w1.n_pad; p.k_pad = w1.k_pad; p.dilation = dilation;
  p.h_b1 = w1.h_bias; p.h_b2 = w2.h_bias; p.slope = slope;
  p.frame_mask = io.mask; p.rate = io.rate;
  p.post_scale = 1.f; p.out_slope = slope;
  p.out_ld = c;
  return p;
}

}  // namespace jb

extern "C" int jatts_hifigan_create(const jatts_hifigan_config* cfg, const jatts_tensor* weights, int32_t n_weights,
                                    jatts_hifigan** out) {
  JB_REQUIRE(cfg && weights && out, JATTS_E_INVALID, "hifigan_create: null argument");
  JB_REQUIRE(cfg->out_channels == 1, JATTS_E_UNSUPPORTED, "only out_channels == 1 is implemented");
  JB_REQUIRE(cfg->n_upsamples >= 1 && cfg->n_upsamples <= 8 && cfg->n_resblocks >= 1 && cfg->n_resblocks <= 8 &&
                 cfg->n_dilations >= 1 && cfg->n_dilations <= 8,
             JATTS_E_INVALID, "hifigan_create: bad stage/block counts");
  JB_REQUIRE(cfg->channels % (1 << cfg->n_upsamples) == 0 && (cfg->channels >> cfg->n_upsamples) % 32 == 0,
             JATTS_E_UNSUPPORTED, "every stage needs a channel count that is a multiple of 32");
  JB_REQUIRE((cfg->kernel_size & 1) == 1, JATTS_E_UNSUPPORTED, "kernel_size must be odd");
  JB_REQUIRE(cfg->lrelu_slope > 0.f && cfg->lrelu_slope <= 1.f, JATTS_E_UNSUPPORTED,
             "lrelu_slope must be in (0, 1] (activations are stored as LeakyReLU(x) and inverted with 1/slope)");
  jatts_hifigan* h = new jatts_hifigan();
  h->cfg = *cfg;
  auto fail = [&](int rc) { delete h; return rc; };
  if (cudaGetDevice(&h->device) != cudaSuccess) { set_last_error("cudaGetDevice failed (no CUDA device?)"); return fail(JATTS_E_CUDA); }
  int rc = h->wt.init(weights, n_weights);
  if (rc) return fail(rc);
  if ((rc = load_conv(h->wt, "input_conv", cfg->kernel_size, cfg->channels, cfg->in_channels, false, true, cfg->channels, &h->input_conv))) return fail(rc);
  int r = 1;
  h->ups.resize(cfg->n_upsamples);
  h->ups3.resize(cfg->n_upsamples);
  h->c1.resize(cfg->n_upsamples);
  h->c2.resize(cfg->n_upsamples);
  for (int i = 0; i < cfg->n_upsamples; ++i) {
    const int s = cfg->upsample_scales[i];
    const int ci = cfg->channels >> i, co = cfg->channels >> (i + 1);
    if (s < 1) { set_last_error("upsample scale < 1"); return fail(JATTS_E_INVALID); }
    r *= s;
    h->rate.push_back(r);
    h->chans.push_back(co);
    if ((rc = load_conv(h->wt, "ups" + std::to_string(i), 2, s * co, ci, false, true, s * co, &h->ups[i]))) return fail(rc);
    // narrow stages also come as ONE 3-tap convolution with N = s * C_out (<= 128) columns (_pack.py::pack_hifigan)
    if (s * co <= 128 && h->wt.has("ups" + std::to_string(i) + ".t3.hi")) {
      if ((rc = load_conv(h->wt, "ups" + std::to_string(i) + ".t3", 3, 128, ci, false, true, s * co, &h->ups3[i]))) return fail(rc);
    }
    h->c1[i].resize(cfg->n_resblocks);
    h->c2[i].resize(cfg->n_resblocks);
    for (int j = 0; j < cfg->n_resblocks; ++j) {
      const int k = cfg->resblock_kernels[j];
      if ((k & 1) == 0) { set_last_error("resblock kernel must be odd"); return fail(JATTS_E_UNSUPPORTED); }
      h->c1[i][j].resize(cfg->n_dilations);
      h->c2[i][j].resize(cfg->n_dilations);
      for (int d = 0; d < cfg->n_dilations; ++d) {
        const std::string p = "rb" + std::to_string(i) + "_" + std::to_string(j);
        if ((rc = load_conv(h->wt, p + ".c1_" + std::to_string(d), k, co, co, false, true, co, &h->c1[i][j][d]))) return fail(rc);
        if ((rc = load_conv(h->wt, p + ".c2_" + std::to_string(d), k, co, co, false, true, co, &h->c2[i][j][d]))) return fail(rc);
        for (ConvW* w : {&h->c1[i][j][d], &h->c2[i][j][d]}) {
          h->host_bias.emplace_back(co);
          if (cudaMemcpy(h->host_bias.back().data(), w->bias, sizeof(float) * co, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_last_error("cudaMemcpy(bias) failed");
            return fail(JATTS_E_CUDA);
          }
          w->h_bias = h->host_bias.back().data();   // the inner vector's buffer does not move when host_bias grows
        }
        // the gap between utterances must cover the widest receptive field at this rate
        if ((k - 1) / 2 * cfg->resblock_dilations[j][d] > kGapRows * r) {
          set_last_error("dilated receptive field exceeds the inter-utterance gap");
          return fail(JATTS_E_UNSUPPORTED);
        }
      }
    }
  }
  h->hop = r;
  const int cl = cfg->channels >> cfg->n_upsamples;
  if ((rc = h->wt.f32("output.w", static_cast<long long>(cfg->kernel_size) * cl, &h->out_w))) return fail(rc);
  const float* ob;
  if ((rc = h->wt.f32("output.b", 1, &ob))) return fail(rc);
  if (cudaMemcpy(&h->out_b, ob, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) { set_last_error("cudaMemcpy(output.b) failed"); return fail(JATTS_E_CUDA); }
  if ((rc = h->wt.f32("mel_scale", cfg->in_channels, &h->mel_scale))) return fail(rc);
  if ((rc = h->wt.f32("mel_shift", cfg->in_channels, &h->mel_shift))) return fail(rc);
  if (cudaEventCreateWithFlags(&h->staged, cudaEventDisableTiming) != cudaSuccess) { set_last_error("cudaEventCreate failed"); return fail(JATTS_E_CUDA); }
  *out = h;
  return 0;
}

extern "C" void jatts_hifigan_destroy(jatts_hifigan* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  h->arena.release();
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->d_small) cudaFree(h->d_small);
  if (h->staged) cudaEventDestroy(h->staged);
  delete h;
}

static int hifigan_run_impl(jatts_hifigan* h, const float* d_mel, const int32_t* h_mel_lens, int32_t n_utt,
                            float* d_wave, int16_t* d_pcm, void* stream) {
  JB_REQUIRE(h && d_mel && h_mel_lens && (d_wave || d_pcm) && n_utt > 0, JATTS_E_INVALID, "hifigan_run: bad argument");
  JB_CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const jatts_hifigan_config& c = h->cfg;
  for (int i = 0; i < n_utt; ++i) JB_REQUIRE(h_mel_lens[i] > 0, JATTS_E_INVALID, "every clip needs >= 1 frame");
  HostLayout hl;
  hl.build(h_mel_lens, n_utt);
  JB_REQUIRE(static_cast<long long>(hl.n_rows) * h->hop < (1ll << 31), JATTS_E_UNSUPPORTED,
             "batch too large for 32-bit row indices; split it");
  JB_PROPAGATE(ensure_workspace(h, hl.n_rows, n_utt));
  // stage the layout tables (pinned buffer is reused: wait for the previous call's copy first)
  if (h->staged_pending) JB_CUDA_OK(cudaEventSynchronize(h->staged));
  const int cu = h->cap_utt;
  for (int i = 0; i < n_utt; ++i) {
    h->h_small[i] = hl.seg_start[i];
    h->h_small[cu + i] = hl.seg_len[i];
    h->h_small[2 * cu + i] = hl.off[i];
  }
  JB_CUDA_OK(cudaMemcpyAsync(h->d_small, h->h_small, sizeof(int) * 3 * cu, cudaMemcpyHostToDevice, s));
  JB_CUDA_OK(cudaEventRecord(h->staged, s));
  h->staged_pending = true;
  RowLayout L;
  L.seg_start = h->d_small; L.seg_len = h->d_small + cu; L.frame_mask = h->mask; L.frame_seg = h->seg;
  L.nseg = n_utt; L.n_rows = hl.n_rows;
  const int* d_off = h->d_small + 2 * cu;
  JB_PROPAGATE(fill_layout(L.seg_start, L.seg_len, n_utt, hl.n_rows, h->mask, h->seg, s));

  const int in_pad = round_up(c.in_channels, 64);
  const float slope = c.lrelu_slope;
  // mel re-normalisation (vocoder.py:57-61) fused into the operand conversion
  // Gap rows: every tensor-core launch below stores EVERY row of its output (rows outside an utterance
  // as zeros, through the TMA epilogue) and pack_mel_affine zero-fills the gap rows of the mel operand,
  // so no buffer needs a separate zeroing pass.
  JB_PROPAGATE(pack_mel_affine(d_mel, c.in_channels, h->mel_scale, h->mel_shift, L, d_off, h->mel, in_pad, s));
  StageIO io{h->mask, 1, hl.n_rows};
  int cur = 0;  // y[cur] holds leaky_relu(stage input)
  {
    ConvGemmEpilogue e{};
    e.out_act = h->y[cur]; e.out_act_slope = slope; e.out_act_ld = c.channels;
    JB_PROPAGATE(run_conv(h->input_conv, h->mel, in_pad, io, 1, e, s));
  }
  int c_in = c.channels;
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int sc = c.upsample_scales[i];
    const int co = h->chans[i];
    const StageIO in_io = io;
    // narrow stages (HBM bound when unfused): one launch per residual unit, lrelu(x) is the only stored tensor
    static const bool no_fuse = getenv("JATTS_B200_NO_FUSE") != nullptr;
    bool fused = !no_fuse && (co == 32 || co == 64);
    if (fused) {
      StageIO probe{h->mask, h->rate[i], static_cast<long long>(hl.n_rows) * h->rate[i]};
      for (int j = 0; j < c.n_resblocks && fused; ++j)
        for (int d = 0; d < c.n_dilations && fused; ++d) {
          MrfPairProblem pp = pair_problem(h->c1[i][j][d], h->c2[i][j][d], h->xa0, co, probe, c.resblock_dilations[j][d], slope);
          pp.out = h->x;
          fused = h->c1[i][j][d].taps == h->c2[i][j][d].taps && mrf_pair_eligible(pp);
        }
    }
    io.rate = h->rate[i];
    io.rows = static_cast<long long>(hl.n_rows) * io.rate;
    const int nxt = cur ^ 1;
    {
      // ConvTranspose1d(k = 2s, stride s, padding p): out[j*s + q - p] = x[j] W[:,:,q] + x[j-1] W[:,:,q+s]
      // (SURVEY appendix C).  Each output phase q is a plain 2-tap convolution over the INPUT rows whose
      // result lands on every s-th output row, so it runs on the TMA-epilogue kernel with a strided
      // output view per phase.
      const ConvW& w = h->ups[i];
      const int pp = sc / 2 + sc % 2;
      auto phase_problem = [&](int q) {
        ConvGemmProblem p{};
        p.a_hi = h->y[cur]; p.a_rows = static_cast<int>(in_io.rows); p.a_ld = c_in; p.a_cols = c_in;
        p.w_hi = w.hi; p.taps = 2; p.n_pad = w.n_pad; p.k_pad = w.k_pad;
        // tap 0 reads x[j-1], tap 1 reads x[j]; GEMM row r is j (q >= p) or j-1 (q < p, whose j = 0 row
        // would land on a negative output row), so that view row == GEMM row and no coordinate is negative
        p.tap_off0 = q >= pp ? -1 : 0; p.tap_stride = 1;
        p.w_row0 = w.n_pad + q * co; p.w_tap_stride = -w.n_pad; p.w_rows_total = 2 * w.n_pad;  // packed [x[j] taps | x[j-1] taps]
        p.n = co; p.m_rows = static_cast<int>(in_io.rows); p.block_n = co > 256 ? 256 : co;
        p.frame_mask = h->mask; p.rate = io.rate; p.out_rows = static_cast<int>(io.rows);
        // output view = every s-th row starting at the first non-negative output row of this phase
        const long long first = q >= pp ? q - pp : q - pp + sc;
        p.mask_mul = sc; p.mask_add = static_cast<int>(first);
        p.store_row_off = 0; p.out_pitch_mul = sc;
        p.out_view_rows = static_cast<int>((io.rows - first + sc - 1) / sc);
        ConvGemmEpilogue e{};
        e.bias = w.bias; e.scale = 1.f; e.post_scale = 1.f;
        e.out_act = h->xa0 + first * co; e.out_act_slope = slope; e.out_act_ld = co;
        p.ep = e;
        return p;
      };
      // all phases in ONE launch (units walked input-tile major: the s phases of a tile run concurrently and its
      // activation slab leaves HBM once instead of s times; measured on the hop-300 generator: 17 launches -> 4,
      // convolution family 10.77 -> 10.6 ms per 64-utterance step), or -- scale > 5 (the canonical hop-256 V1 has 8)
      // or C_out > 256 -- one launch per phase
      bool done = false;
      if (h->ups3[i].hi != nullptr) {
        // narrow stage: the whole transposed convolution as one 3-tap convolution whose output row i is frames
        // i*s .. i*s + s - 1 (N = s * C_out <= 128: one tile per input tile; the phase tiles of N = C_out were too
        // small to pay for their own pipeline hand-offs)
        const ConvW& w3 = h->ups3[i];
        ConvGemmProblem p{};
        p.a_hi = h->y[cur]; p.a_rows = static_cast<int>(in_io.rows); p.a_ld = c_in; p.a_cols = c_in;
        p.w_hi = w3.hi; p.taps = 3; p.n_pad = w3.n_pad; p.k_pad = w3.k_pad;
        p.tap_off0 = -1; p.tap_stride = 1;
        p.n = sc * co; p.m_rows = static_cast<int>(in_io.rows); p.block_n = w3.block_n;
        p.frame_mask = h->mask; p.rate = in_io.rate; p.out_rows = static_cast<int>(in_io.rows);
        ConvGemmEpilogue e{};
        e.bias = w3.bias; e.scale = 1.f; e.post_scale = 1.f;
        e.out_act = h->xa0; e.out_act_slope = slope; e.out_act_ld = sc * co;
        p.ep = e;
        if (conv_gemm_tc2_eligible(p)) {
          JB_PROPAGATE(conv_gemm_tc2(p, s));
          done = true;
        }
      }
      ConvGemmProblem all = phase_problem(0);
      all.phases = sc; all.phase_pp = pp;
      all.w_row0 = w.n_pad; all.tap_off0 = 0; all.mask_add = 0; all.out_view_rows = 0;
      all.ep.out_act = h->xa0;
      if (done) {
      } else if (co <= 256 && sc >= 2 && sc <= 5 && conv_gemm_tc2_eligible(all)) {
        JB_PROPAGATE(conv_gemm_tc2(all, s));
      } else {
        for (int q = 0; q < sc; ++q) {
          ConvGemmProblem p = phase_problem(q);
          JB_REQUIRE(conv_gemm_tc2_eligible(p), JATTS_E_UNSUPPORTED, "transposed-conv phase not eligible for the TMA kernel");
          JB_PROPAGATE(conv_gemm_tc2(p, s));
        }
      }
    }
    const bool last_stage = i == c.n_upsamples - 1;
    const float next_slope = last_stage ? 0.01f : slope;  // torch.nn.LeakyReLU() default before output_conv
    for (int j = 0; j < c.n_resblocks && fused; ++j) {
      const bf16* in = h->xa0;
      for (int d = 0; d < c.n_dilations; ++d) {
        MrfPairProblem pp = pair_problem(h->c1[i][j][d], h->c2[i][j][d], in, co, io, c.resblock_dilations[j][d], slope);
        if (d + 1 < c.n_dilations) {
          pp.out = (d & 1) ? h->xa : h->x;   // ping-pong: other CTAs still read the input's halo rows
        } else {
          // branch output: accumulate the sum over residual blocks (in place, a tile reads its own rows before
          // storing them); the last block emits the next layer's operand leaky_relu(mean) directly
          if (j > 0) { pp.accum = h->sum; pp.accum_ld = co; }
          if (j + 1 < c.n_resblocks) {
            pp.out = h->sum; pp.out_slope = 1.f;
          } else {
            pp.out = h->y[nxt]; pp.post_scale = 1.0f / c.n_resblocks; pp.out_slope = next_slope;
          }
        }
        JB_PROPAGATE(mrf_pair(pp, s));
        in = pp.out;
      }
    }
    for (int j = 0; j < c.n_resblocks && !fused; ++j) {
      // wide stages: two launches per residual unit.  Only lrelu(x) is stored (as in the fused path): conv2's
      // epilogue recovers x from it for the residual add, so a unit moves 5 activation tensors instead of 6
      // and its epilogue ring entries are one slab wide.
      const bf16* xa = h->xa0;
      for (int d = 0; d < c.n_dilations; ++d) {
        ConvGemmEpilogue e1{};
        e1.act = ACT_LRELU; e1.slope = slope; e1.out_hi = h->t; e1.out_bf_ld = co;
        JB_PROPAGATE(run_conv(h->c1[i][j][d], xa, co, io, c.resblock_dilations[j][d], e1, s));
        ConvGemmEpilogue e2{};
        e2.res_bf16 = xa; e2.res_ld = co; e2.res_inv_slope = 1.0f / slope;
        if (d + 1 < c.n_dilations) {
          bf16* out = (d & 1) ? h->xa : h->x;
          e2.out_act = out; e2.out_act_slope = slope; e2.out_act_ld = co;
          xa = out;
        } else {
          // branch output: accumulate the mean over residual blocks; the last block emits the next
          // layer's operand leaky_relu(mean) directly
          if (j > 0) e2.accum_bf16 = h->sum;
          if (j + 1 < c.n_resblocks) {
            e2.out_hi = h->sum; e2.out_bf_ld = co;
          } else {
            e2.post_scale = 1.0f / c.n_resblocks;
            e2.out_act = h->y[nxt]; e2.out_act_slope = next_slope; e2.out_act_ld = co;
          }
        }
        JB_PROPAGATE(run_conv(h->c2[i][j][d], h->t, co, io, 1, e2, s));
      }
    }
    cur = nxt;
    c_in = co;
  }
  JB_PROPAGATE(output_conv_tanh(h->y[cur], c_in, c_in, h->out_w, h->out_b, c.kernel_size, L, h->hop, d_off, d_wave,
                                reinterpret_cast<short*>(d_pcm), s));
  return 0;
}

extern "C" int jatts_hifigan_run(jatts_hifigan* h, const float* d_mel, const int32_t* h_mel_lens, int32_t n_utt,
                                 float* d_wave, void* stream) {
  JB_REQUIRE(d_wave != nullptr, JATTS_E_INVALID, "hifigan_run: null output");
  return hifigan_run_impl(h, d_mel, h_mel_lens, n_utt, d_wave, nullptr, stream);
}

extern "C" int jatts_hifigan_run_pcm16(jatts_hifigan* h, const float* d_mel, const int32_t* h_mel_lens, int32_t n_utt,
                                       int16_t* d_pcm, void* stream) {
  JB_REQUIRE(d_pcm != nullptr, JATTS_E_INVALID, "hifigan_run_pcm16: null output");
  return hifigan_run_impl(h, d_mel, h_mel_lens, n_utt, nullptr, d_pcm, stream);
}
