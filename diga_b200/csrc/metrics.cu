// f5 (last row) — evaluation confusion matrix.
// Replaces runningScore._fast_hist / update (util/metrics.py:32-41 of the reference): the reference copies prediction and
// ground truth to the host and runs np.bincount(n * true[mask] + pred[mask]) per image on one core.  Here the label maps
// stay on the GPU: 2 label reads per pixel (uint8 or int64, as the producers deliver them), nothing written but the
// n x n counters.  Integer counting: exact, order-independent, bit-equal to the reference matrix.
//
// A CTA keeps the matrix in shared memory (<= 32 x 32 counters); a warp first groups its lanes by bin
// (__match_any_sync: label maps are piecewise constant, so most of a warp lands in one or two bins) and issues one
// shared atomic per distinct bin; the CTA adds its non-zero counters to the global int64 matrix at the end.
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

template <typename TT, typename TP, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
confusion_kernel(const TT* __restrict__ label_true, const TP* __restrict__ label_pred, int64_t total, int n_class,
                 unsigned long long* __restrict__ hist, unsigned int* __restrict__ flags) {
  __shared__ unsigned int sh[DIGA_MAX_CLASSES * DIGA_MAX_CLASSES];
  const int bins = n_class * n_class;
  for (int i = threadIdx.x; i < bins; i += BLOCK) sh[i] = 0;
  __syncthreads();
  bool bad_pred = false;
  // one label pair per thread and iteration, consecutive threads on consecutive pixels (coalesced for either type)
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  const int64_t rounds = (total + stride - 1) / stride;      // every lane runs every round: __match_any_sync needs them all
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    int bin = -1;
    if (i < total) {
      const long long t = (long long)label_true[i];
      const long long p = (long long)label_pred[i];
      if (t >= 0 && t < n_class) {                               // metrics.py:33  mask = (true >= 0) & (true < n_class)
        if (p >= 0 && p < n_class) bin = (int)(n_class * t + p); // :35
        else bad_pred = true;                                    // np.bincount(...).reshape would raise in the reference
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[bin], (unsigned)__popc(peers));
  }
  if (bad_pred) atomicOr(flags, 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += BLOCK)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

template <typename TT, typename TP>
static int launch_confusion(const void* lt, const void* lp, int64_t total, int n_class, int64_t* hist, unsigned int* flags,
                            cudaStream_t st) {
  constexpr int BLOCK = 256;
  int64_t grid = (total + BLOCK - 1) / BLOCK;
  // a CTA must not count more than 2^32 - 1 pixels into one 32-bit shared counter
  const int64_t cap = (int64_t)sm_count() * tunable("confusion_ctas_per_sm", 8);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  while ((total + grid - 1) / grid >= ((int64_t)1 << 32)) grid *= 2;
  confusion_kernel<TT, TP, BLOCK><<<(unsigned)grid, BLOCK, 0, st>>>(reinterpret_cast<const TT*>(lt), reinterpret_cast<const TP*>(lp),
                                                                    total, n_class, reinterpret_cast<unsigned long long*>(hist), flags);
  DIGA_CHECK_LAUNCH("confusion_kernel");
  return DIGA_OK;
}

}  // namespace diga

extern "C" {

int diga_confusion_matrix(const void* label_true, int true_is_u8, const void* label_pred, int pred_is_u8, int64_t total,
                          int64_t n_class, int64_t* hist, uint32_t* flags, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(label_true && label_pred && hist && flags, DIGA_ERR_INVALID, "confusion_matrix: null pointer");
  DIGA_REQUIRE(n_class >= 1 && n_class <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "confusion_matrix: n_class=%lld outside [1,%d]",
               (long long)n_class, DIGA_MAX_CLASSES);
  DIGA_REQUIRE(total >= 0, DIGA_ERR_INVALID, "confusion_matrix: negative size");
  DIGA_REQUIRE(aligned(hist, 8) && aligned(flags, 4) && (true_is_u8 || aligned(label_true, 8)) && (pred_is_u8 || aligned(label_pred, 8)),
               DIGA_ERR_MISALIGNED, "confusion_matrix: misaligned pointer");
  if (total == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (true_is_u8 && pred_is_u8) return launch_confusion<uint8_t, uint8_t>(label_true, label_pred, total, (int)n_class, hist, flags, st);
  if (true_is_u8) return launch_confusion<uint8_t, int64_t>(label_true, label_pred, total, (int)n_class, hist, flags, st);
  if (pred_is_u8) return launch_confusion<int64_t, uint8_t>(label_true, label_pred, total, (int)n_class, hist, flags, st);
  return launch_confusion<int64_t, int64_t>(label_true, label_pred, total, (int)n_class, hist, flags, st);
}

}  // extern "C"
