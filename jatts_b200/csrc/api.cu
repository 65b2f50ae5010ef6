// C-ABI glue: version / error / launch counter and the op-level test entry point.
#include "engine_common.cuh"
#include "mrf_pair.cuh"

using namespace jb;

extern "C" int jatts_abi_version(void) { return JATTS_B200_ABI_VERSION; }
extern "C" const char* jatts_last_error(void) { return get_last_error(); }
extern "C" int64_t jatts_launch_count(void) { return g_launch_count; }

extern "C" int jatts_op_conv_gemm(const jatts_conv_gemm_args* a, int32_t impl, void* stream) {
  JB_REQUIRE(a != nullptr, JATTS_E_INVALID, "op_conv_gemm: null args");
  ConvGemmProblem p{};
  p.a_hi = static_cast<const bf16*>(a->d_a_hi);
  p.a_lo = static_cast<const bf16*>(a->d_a_lo);
  p.a_rows = a->a_rows; p.a_ld = a->a_ld; p.a_cols = a->a_cols;
  p.w_hi = static_cast<const bf16*>(a->d_w_hi);
  p.w_lo = static_cast<const bf16*>(a->d_w_lo);
  p.taps = a->taps; p.n_pad = a->n_pad; p.k_pad = a->k_pad; p.tap_off0 = a->tap_off0; p.tap_stride = a->tap_stride;
  p.n = a->n; p.m_rows = a->m_rows; p.block_n = a->block_n;
  p.frame_mask = a->d_frame_mask; p.rate = a->rate; p.out_rows = a->out_rows;
  p.up_s = a->up_s; p.up_p = a->up_p; p.up_cout = a->up_cout;
  ConvGemmEpilogue& e = p.ep;
  e.bias = a->d_bias; e.act = a->act; e.slope = a->slope; e.scale = a->scale;
  e.res_f32 = a->d_res_f32; e.res_bf16 = static_cast<const bf16*>(a->d_res_bf16); e.res_ld = a->res_ld;
  e.accum_in = a->d_accum_in; e.accum_bf16 = static_cast<const bf16*>(a->d_accum_bf16); e.post_scale = a->post_scale;
  e.out_f32 = a->d_out_f32; e.out_f32_ld = a->out_f32_ld;
  e.out_hi = static_cast<bf16*>(a->d_out_hi); e.out_lo = static_cast<bf16*>(a->d_out_lo); e.out_bf_ld = a->out_bf_ld;
  e.out_act = static_cast<bf16*>(a->d_out_act); e.out_act_slope = a->out_act_slope; e.out_act_ld = a->out_act_ld;
  e.snake_a = a->d_snake_a; e.snake_ib = a->d_snake_ib;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return impl == 0 ? conv_gemm_tc(p, s) : conv_gemm_simt_debug(p, s);
}

extern "C" int jatts_op_mrf_pair(const jatts_mrf_pair_args* a, void* stream) {
  JB_REQUIRE(a != nullptr, JATTS_E_INVALID, "op_mrf_pair: null args");
  MrfPairProblem p{};
  p.xa = static_cast<const bf16*>(a->d_xa); p.rows = a->rows; p.ld = a->ld; p.c = a->c;
  p.w1 = static_cast<const bf16*>(a->d_w1); p.w2 = static_cast<const bf16*>(a->d_w2);
  p.taps = a->taps; p.n_pad = a->n_pad; p.k_pad = a->k_pad; p.dilation = a->dilation;
  JB_REQUIRE(a->d_b1 && a->d_b2 && a->c > 0 && a->c <= 64, JATTS_E_INVALID, "op_mrf_pair: bad bias / channel count");
  float hb[2][64];   // the engine keeps host copies of the biases; the test entry fetches them here
  JB_CUDA_OK(cudaMemcpy(hb[0], a->d_b1, sizeof(float) * a->c, cudaMemcpyDeviceToHost));
  JB_CUDA_OK(cudaMemcpy(hb[1], a->d_b2, sizeof(float) * a->c, cudaMemcpyDeviceToHost));
  p.h_b1 = hb[0]; p.h_b2 = hb[1]; p.slope = a->slope;
  p.frame_mask = a->d_frame_mask; p.rate = a->rate;
  p.accum = static_cast<const bf16*>(a->d_accum); p.accum_ld = a->accum_ld;
  p.post_scale = a->post_scale; p.out_slope = a->out_slope;
  p.out = static_cast<bf16*>(a->d_out); p.out_ld = a->out_ld;
  return mrf_pair(p, static_cast<cudaStream_t>(stream));
}

extern "C" int jatts_op_relpos_attention(const jatts_relpos_attention_args* a, void* stream) {
  JB_REQUIRE(a != nullptr, JATTS_E_INVALID, "op_relpos_attention: null args");
  RowLayout L{};
  L.seg_start = a->d_seg_start; L.seg_len = a->d_seg_len; L.nseg = a->nseg; L.n_rows = static_cast<int>(a->x_rows);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t bytes = relpos_attention_scratch_bytes(a->max_len, a->nseg, a->n_head);
  float* scratch = nullptr;   // the engine keeps its scratch; the test entry allocates one per call
  if (bytes) JB_CUDA_OK(cudaMalloc(&scratch, bytes));
  int rc;
  if (a->d_pos_hi == nullptr && a->d_pos_lo == nullptr)   // plain attention: x = [q | k | v]
    rc = plain_attention(static_cast<const bf16*>(a->d_x_hi), static_cast<const bf16*>(a->d_x_lo), a->x_rows, a->n_head,
                         a->d_model, L, a->max_len, scratch, bytes, static_cast<bf16*>(a->d_out_hi),
                         static_cast<bf16*>(a->d_out_lo), a->out_ld, s);
  else
    rc = relpos_attention(static_cast<const bf16*>(a->d_x_hi), static_cast<const bf16*>(a->d_x_lo), a->x_rows,
                          static_cast<const bf16*>(a->d_pos_hi), static_cast<const bf16*>(a->d_pos_lo), a->pos_rows,
                          a->n_head, a->d_model, L, a->max_len, scratch, bytes, static_cast<bf16*>(a->d_out_hi),
                          static_cast<bf16*>(a->d_out_lo), a->out_ld, s);
  const cudaError_t e = cudaStreamSynchronize(s);
  if (scratch) cudaFree(scratch);
  if (rc == 0 && e != cudaSuccess) {
    set_last_error(std::string("op_relpos_attention: ") + cudaGetErrorString(e));
    return JATTS_E_CUDA;
  }
  return rc;
}

extern "C" int jatts_profile_begin(void) {
  for (auto& e : g_profile_events) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
  g_profile_events.clear();
  g_profile_on = true;
  return 0;
}

extern "C" int jatts_profile_end(double* ms_bf16, int64_t* n_bf16, double* ms_split, int64_t* n_split) {
  g_profile_on = false;
  double ms[2] = {0.0, 0.0};
  int64_t n[2] = {0, 0};
  for (auto& e : g_profile_events) {
    JB_CUDA_OK(cudaEventSynchronize(e.e1));
    float t = 0.f;
    JB_CUDA_OK(cudaEventElapsedTime(&t, e.e0, e.e1));
    if (e.split == 0 || e.split == 1) {
      ms[e.split] += t;
      n[e.split] += 1;
    }
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  g_profile_events.clear();
  if (ms_bf16) *ms_bf16 = ms[0];
  if (n_bf16) *n_bf16 = n[0];
  if (ms_split) *ms_split = ms[1];
  if (n_split) *n_split = n[1];
  return 0;
}

extern "C" int jatts_profile_end_classes(double* ms, int64_t* n, int32_t n_classes) {
  JB_REQUIRE(ms && n && n_classes >= PROF_CLASSES, JATTS_E_INVALID, "profile_end_classes: need room for every class");
  g_profile_on = false;
  for (int i = 0; i < n_classes; ++i) { ms[i] = 0.0; n[i] = 0; }
  for (auto& e : g_profile_events) {
    JB_CUDA_OK(cudaEventSynchronize(e.e1));
    float t = 0.f;
    JB_CUDA_OK(cudaEventElapsedTime(&t, e.e0, e.e1));
    if (e.split >= 0 && e.split < n_classes) { ms[e.split] += t; n[e.split] += 1; }
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
  }
  g_profile_events.clear();
  return 0;
}

namespace jb { extern long long* g_trace_ptr; }
// debug only: device buffer of 5*8*64 int64 that CTA 0 of the TMA-epilogue conv kernel fills with clock64 stamps
extern "C" int jatts_debug_set_trace(void* d_buf) {
  jb::g_trace_ptr = static_cast<long long*>(d_buf);
  return 0;
}
