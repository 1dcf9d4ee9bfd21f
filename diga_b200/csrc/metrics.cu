// f5 (last row) — evaluation confusion matrix.
// Replaces runningScore._fast_hist / update (util/metrics.py:32-41 of the reference): the reference copies prediction and
// ground truth to the host and runs np.bincount(n * true[mask] + pred[mask]) per image on one core.  Here the label maps
// stay on the GPU: 2 label reads per pixel (uint8 or int64, as the producers deliver them), nothing written but the
// n x n counters.  Integer counting: exact, order-independent, bit-equal to the reference matrix.
//
// A CTA keeps the matrix in shared memory (<= 32 x 32 counters); a warp issues one shared atomic per distinct bin (label
// maps are piecewise constant, so most warps land in one bin); the CTA adds its non-zero counters to the global int64
// matrix at the end.
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

// Four consecutive labels of one map as 32-bit values, anything outside [0, 2^31) folded to -1 (neither a class nor a
// countable prediction): one 32-bit load (uint8) / two 128-bit loads (int64) when the map's base is aligned for it (`vec`)
// and the group lies inside the map, else element by element (-1 past the end).
template <typename T>
__device__ __forceinline__ int fold_label(T x) {
  if constexpr (sizeof(T) == 1) return (int)x;
  else return (unsigned long long)x < 0x80000000ull ? (int)x : -1;
}
template <typename T>
__device__ __forceinline__ void load_labels4(const T* __restrict__ p, int64_t i, int64_t total, bool vec, int (&v)[4]) {
  if (vec && i + 3 < total) {
    if constexpr (sizeof(T) == 1) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + i));
      v[0] = w & 0xffu, v[1] = (w >> 8) & 0xffu, v[2] = (w >> 16) & 0xffu, v[3] = w >> 24;
    } else {
      const longlong2 a = ld_stream_i64x2(reinterpret_cast<const int64_t*>(p + i));
      const longlong2 b = ld_stream_i64x2(reinterpret_cast<const int64_t*>(p + i + 2));
      v[0] = fold_label<long long>(a.x), v[1] = fold_label<long long>(a.y);
      v[2] = fold_label<long long>(b.x), v[3] = fold_label<long long>(b.y);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i + k < total ? fold_label<T>(p[i + k]) : -1;
  }
}

// One shared atomic per RUN of equal bins over consecutive lanes (bin < 0: not counted): a lane whose lower neighbour holds
// another bin heads a run and adds weight x (distance to the next head).  Label maps are piecewise constant, so a warp
// holds one or two runs; no __match_any_sync (whose cost grows with the number of distinct values in the warp).
__device__ __forceinline__ void add_runs(unsigned int* sh, int bin, int lane, unsigned weight) {
  const int prev = __shfl_up_sync(0xffffffffu, bin, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || bin != prev);
  if (bin >= 0 && ((heads >> lane) & 1u)) {
    const unsigned above = heads & ~((2u << lane) - 1u);          // heads above this lane (lane 31: none)
    const int next = above ? __ffs(above) - 1 : 32;
    atomicAdd(&sh[bin], weight * (unsigned)(next - lane));
  }
}

// A thread owns four consecutive pixels per round (both maps' loads issued before the first use: at one pixel per thread the
// kernel was bound by the bytes in flight, 2.4 TB/s), consecutive threads consecutive groups.  Counting: when every
// thread's four pixels agree — the common case on piecewise-constant maps — one pass of `add_runs` with weight 4, otherwise
// one pass per pixel slot.
template <typename TT, typename TP, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
confusion_kernel(const TT* __restrict__ label_true, const TP* __restrict__ label_pred, int64_t total, int n_class,
                 unsigned long long* __restrict__ hist, unsigned int* __restrict__ flags) {
  __shared__ unsigned int sh[DIGA_MAX_CLASSES * DIGA_MAX_CLASSES];
  const int bins = n_class * n_class;
  for (int i = threadIdx.x; i < bins; i += BLOCK) sh[i] = 0;
  __syncthreads();
  bool bad_pred = false;
  const bool vec_t = (reinterpret_cast<uintptr_t>(label_true) & (sizeof(TT) == 1 ? 3 : 15)) == 0;   // groups start at i % 4 == 0
  const bool vec_p = (reinterpret_cast<uintptr_t>(label_pred) & (sizeof(TP) == 1 ? 3 : 15)) == 0;
  const int lane = threadIdx.x & 31;
  const int64_t groups = (total + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * BLOCK;
  const int64_t rounds = (groups + stride - 1) / stride;     // every lane runs every round: the warp votes need them all
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i = (r * stride + (int64_t)blockIdx.x * BLOCK + threadIdx.x) * 4;
    int t[4], p[4];
    load_labels4(label_true, i, total, vec_t, t);          // groups past the end load nothing and come back as -1
    load_labels4(label_pred, i, total, vec_p, p);
    int bin[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool counted = (unsigned)t[k] < (unsigned)n_class;       // metrics.py:33  mask = (true >= 0) & (true < n_class)
      const bool in_range = (unsigned)p[k] < (unsigned)n_class;
      bin[k] = counted && in_range ? n_class * t[k] + p[k] : -1;     // :35
      bad_pred |= counted && !in_range;                              // np.bincount(...).reshape would raise in the reference
    }
    if (__all_sync(0xffffffffu, bin[0] == bin[1] && bin[1] == bin[2] && bin[2] == bin[3])) {
      add_runs(sh, bin[0], lane, 4u);                     // every thread's four pixels agree: one pass over the lanes
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) add_runs(sh, bin[k], lane, 1u);
    }
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
  int64_t grid = ((total + 3) / 4 + BLOCK - 1) / BLOCK;
  // a CTA must not count more than 2^32 - 1 pixels into one 32-bit shared counter
  const int64_t cap = (int64_t)sm_count() * tunable("confusion_ctas_per_sm", 8);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  while ((total + grid - 1) / grid + 4 * BLOCK >= ((int64_t)1 << 32)) grid *= 2;
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
