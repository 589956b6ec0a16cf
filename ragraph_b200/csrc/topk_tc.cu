// placeholder until the tcgen05 filter lands (next commit)
#include "common.cuh"
namespace rag {
bool topk_tc_available() { return false; }
size_t topk_tc_workspace(int64_t, int64_t, int, int, int) { return 256; }
int topk_tc_run(const float*, int64_t, const float*, const float*, const uint16_t*, int64_t, int, int, int mode,
                uint32_t, int64_t, float*, int64_t*, void*, size_t, cudaStream_t) {
  return fail(RAG_EUNSUPPORTED, "cosine_topk: mode %d not built yet", mode);
}
}  // namespace rag
