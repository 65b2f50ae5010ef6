// Host-side plumbing shared by the two engines: weight lookup, workspace arena, layout upload,
// convolution descriptors.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jatts_b200.h"
#include "conv_gemm.cuh"
#include "kernels.cuh"

namespace jb {

struct WeightTable {
  std::unordered_map<std::string, jatts_tensor> m;
  int init(const jatts_tensor* w, int n) {
    for (int i = 0; i < n; ++i) {
      JB_REQUIRE(w[i].name && w[i].d_ptr, JATTS_E_INVALID, "weight table: null entry");
      m[w[i].name] = w[i];
    }
    return 0;
  }
  bool has(const std::string& name) const { return m.count(name) != 0; }
  int get(const std::string& name, int dtype, long long numel, const void** out) const {
    auto it = m.find(name);
    JB_REQUIRE(it != m.end(), JATTS_E_INVALID, "missing weight '" + name + "'");
    JB_REQUIRE(it->second.dtype == dtype, JATTS_E_INVALID, "weight '" + name + "' has the wrong dtype");
    JB_REQUIRE(numel < 0 || it->second.numel == numel, JATTS_E_INVALID,
               "weight '" + name + "' has " + std::to_string(it->second.numel) + " elements, expected " +
                   std::to_string(numel));
    *out = it->second.d_ptr;
    return 0;
  }
  int f32(const std::string& name, long long numel, const float** out) const {
    return get(name, JATTS_F32, numel, reinterpret_cast<const void**>(out));
  }
};

// One convolution's repacked weights: [taps][n_pad][k_pad] bf16, or an fp16 (hi, lo*2^11) pair in split mode.
struct ConvW {
  const bf16* hi = nullptr;
  const bf16* lo = nullptr;
  const float* bias = nullptr;
  const float* h_bias = nullptr;   // optional host copy (owned by the engine), for kernels that take the bias by value
  int taps = 1, n = 0, n_pad = 0, k_pad = 0, block_n = 128;
};

// tile width rule -- mirrored by jatts_b200/_pack.py::pick_block_n
inline int pick_block_n(int n_cols, bool split) {
  if (split) return 128;
  if (n_cols % 256 == 0) return 256;
  if (n_cols % 128 == 0) return 128;
  if (n_cols % 64 == 0) return 64;
  return 32;
}

inline int load_conv(const WeightTable& wt, const std::string& name, int taps, int n_cols, int c_in, bool split,
                     bool has_bias, int n_out, ConvW* w) {
  w->taps = taps;
  w->n = n_out;
  w->block_n = pick_block_n(n_cols, split);
  w->n_pad = round_up(n_cols, w->block_n);
  w->k_pad = round_up(c_in, 64);
  const long long numel = static_cast<long long>(taps) * w->n_pad * w->k_pad;
  // split weights are fp16 pairs (hi, lo * 2^11), plain weights bf16
  JB_PROPAGATE(wt.get(name + ".hi", split ? JATTS_F16 : JATTS_BF16, numel, reinterpret_cast<const void**>(&w->hi)));
  if (split) JB_PROPAGATE(wt.get(name + ".lo", JATTS_F16, numel, reinterpret_cast<const void**>(&w->lo)));
  if (has_bias) JB_PROPAGATE(wt.f32(name + ".b", -1, &w->bias));
  return 0;
}

struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (base) JB_CUDA_OK(cudaFree(base));
    base = nullptr;
    cap = 0;
    JB_CUDA_OK(cudaMalloc(&base, bytes));
    JB_CUDA_OK(cudaMemset(base, 0, bytes));
    cap = bytes;
    return 0;
  }
  void reset() { off = 0; }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 1023) & ~size_t(1023);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
  static size_t padded(size_t bytes) { return (bytes + 1023) & ~size_t(1023); }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
  }
};

// Host description of a packed-with-gaps layout + its device mirror.
struct HostLayout {
  std::vector<int> seg_start, seg_len, off;  // off = utterance-contiguous offsets (exclusive scan of seg_len)
  int n_rows = 0, total = 0, max_len = 0;
  void build(const int* lens, int n) {
    seg_start.resize(n);
    seg_len.assign(lens, lens + n);
    off.resize(n);
    int row = 0, acc = 0;
    max_len = 0;
    for (int i = 0; i < n; ++i) {
      seg_start[i] = row;
      off[i] = acc;
      row += lens[i] + kGapRows;
      acc += lens[i];
      if (lens[i] > max_len) max_len = lens[i];
    }
    n_rows = row;
    total = acc;
  }
};

}  // namespace jb
