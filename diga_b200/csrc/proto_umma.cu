// a5 tensor-core path (tcgen05 / TMEM) — placeholder until the UMMA kernel lands; reports "unsupported"
// so diga_proto_distance uses the FP32 kernel in proto.cu.
#include "common.cuh"

namespace diga {

int proto_umma_supported(int64_t, int64_t, int64_t, int64_t) { return 0; }
size_t proto_umma_workspace_bytes(int64_t, int64_t) { return 256; }
int proto_umma_launch(const float*, const float*, int64_t, int64_t, int64_t, int64_t, float*, float*, void*, cudaStream_t) {
  set_error("proto_umma: not available");
  return DIGA_ERR_INVALID;
}

}  // namespace diga
