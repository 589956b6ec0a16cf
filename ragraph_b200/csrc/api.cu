// Status / error plumbing and small host utilities of the C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace rag {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}
int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return RAG_ECUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int dev_attr(cudaDeviceAttr a) {
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, a, dev);
  return v;
}
int sm_count() {
  static int v = dev_attr(cudaDevAttrMultiProcessorCount);
  return v > 0 ? v : 148;
}
int max_smem_optin() {
  static int v = dev_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin);
  return v;
}
}  // namespace rag

extern "C" {
int rag_abi_version(void) { return RAG_ABI_VERSION; }
const char* rag_last_error(void) { return rag::g_err; }
const char* rag_status_string(int s) {
  switch (s) {
    case RAG_OK: return "RAG_OK";
    case RAG_EINVAL: return "RAG_EINVAL";
    case RAG_EALIGN: return "RAG_EALIGN";
    case RAG_EUNSUPPORTED: return "RAG_EUNSUPPORTED";
    case RAG_ECUDA: return "RAG_ECUDA";
    case RAG_EWORKSPACE: return "RAG_EWORKSPACE";
    default: return "RAG_E?";
  }
}
int64_t rag_launch_count(void) { return rag::g_launches.load(std::memory_order_relaxed); }
}
