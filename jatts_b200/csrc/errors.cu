#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
namespace jb {
long long g_launch_count = 0;
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function IN ONE DEVICE CONTEXT: the opt-in is
// remembered per (device, kernel), so a second GPU used from the same process gets its own cudaFuncSetAttribute.
int ensure_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> granted;
  int dev = 0;
  JB_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  int& have = granted[std::make_pair(dev, kernel)];
  if (bytes > have) {
    JB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    have = bytes;
  }
  return 0;
}
}  // namespace jb
