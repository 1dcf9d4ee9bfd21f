// a6 / a7 — per-class centroid accumulation and running update.
// Replaces Class_Features.calculate_mean_vector / _by_output / update_objective_SingleVector
// (calc_centroids.py:97-164 of the reference) and the one-hot helper util/utils.py:158-163.
//
// Pipeline per batch of N images (all on one stream, no host sync):
//   assign : logits [N,C,hw] (+ optional fp32 labels [N,1,hw]) -> cls u8 [N,hw], counts i32 [N,C]
//   accum  : feat [N,D,hw], cls -> sums [N,C,D]               (the HBM-bound kernel: D*4 B per feature px)
//   means  : sums, counts -> vec [N,C,D], vecsum [N,C], valid [N,C]
//   update : vec -> objective_vectors [C,D], objective_num [C]  (sequential over images, parallel over (c,d))
//
// accum is a segmented reduction keyed by the per-pixel class.  Features are NCHW, so for one channel
// row the pixels are contiguous: a warp reads 32 adjacent pixels of R=4 consecutive channel rows
// (4 coalesced 128-byte requests) and each lane adds its 4 values, as one 128-bit shared-memory
// read-modify-write, into a LANE-PRIVATE accumulator acc[warp][class][lane] (float4).  Lane-private
// columns make the update conflict-free whatever the label pattern is — segmentation maps are
// piecewise constant, so a shared per-class cell would see 32-way same-address collisions.
// A CTA owns (image, 4-row group) and covers all pixels of that image, so every sums[n][c][d] has
// exactly one writer: no global atomics, bitwise deterministic.
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

// ------------------------------------------------------------------------------------------------
// assign
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_assign_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int nclass, int64_t hw,
                       uint8_t* __restrict__ cls, int32_t* __restrict__ counts) {
  __shared__ int hist[DIGA_MAX_CLASSES];
  if (threadIdx.x < DIGA_MAX_CLASSES) hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t img = blockIdx.y;
  const float* lg = logits + img * nclass * hw;
  for (int64_t base = (int64_t)blockIdx.x * BLOCK; base < hw; base += (int64_t)gridDim.x * BLOCK) {
    const int64_t p = base + threadIdx.x;
    int c = 255;
    if (p < hw) {
      float m = __ldg(lg + p);
      int am = 0;
      for (int k = 1; k < nclass; ++k) {
        const float v = __ldg(lg + (int64_t)k * hw + p);
        if (v > m) {   // argmax(softmax(out)) == first index of the max logit (calc_centroids.py:121-122)
          m = v;
          am = k;
        }
      }
      c = am;
      if (labels != nullptr) {
        // process_label(labels) * process_label(argmax): the pixel counts for class t iff
        // long(label) == t == argmax, t < C (utils.py:161-162, calc_centroids.py:126-127).
        const float lf = __ldg(labels + img * hw + p);
        const bool in_range = lf < (float)nclass;
        if (!(in_range && (long long)lf == (long long)am)) c = 255;
      }
      cls[img * hw + p] = (uint8_t)c;
    }
    // warp-aggregated histogram: one shared atomic per distinct class per warp
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c != 255 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[c], __popc(peers));
  }
  __syncthreads();
  if ((int)threadIdx.x < nclass && hist[threadIdx.x]) atomicAdd(&counts[img * nclass + threadIdx.x], hist[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// accum
// ------------------------------------------------------------------------------------------------
template <int R> struct AccT;
template <> struct AccT<4> { using type = float4; };
template <> struct AccT<2> { using type = float2; };

// One batch = U pixel stripes x R rows of loads, all issued before any is consumed.
template <int R, int U>
struct AccumBatch {
  float x[U][R];
  int cid[U];
};

template <int R, int WARPS, int U>
__device__ __forceinline__ void accum_load(AccumBatch<R, U>& b, const float* const (&rows)[R], const uint8_t* cl, int64_t base,
                                           int64_t hw) {
  constexpr int STRIPE = WARPS * 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t p = base + (int64_t)u * STRIPE;
    const bool ok = p < hw;
    b.cid[u] = ok ? (int)__ldg(cl + p) : 255;
#pragma unroll
    for (int r = 0; r < R; ++r) b.x[u][r] = ok ? ld_stream<1>(rows[r] + p).v[0] : 0.f;
  }
}

template <int R, int U, typename acc_t>
__device__ __forceinline__ void accum_apply(const AccumBatch<R, U>& b, acc_t* my, int nclass) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (b.cid[u] < nclass) {
      acc_t v = my[b.cid[u] * 32];
      if constexpr (R == 4) {
        v.x += b.x[u][0]; v.y += b.x[u][1]; v.z += b.x[u][2]; v.w += b.x[u][3];
      } else {
        v.x += b.x[u][0]; v.y += b.x[u][1];
      }
      my[b.cid[u] * 32] = v;
    }
  }
}

// PIPE: software pipelining — the loads of batch i+1 are issued before the shared-memory updates of batch i, so a
// warp keeps R*U..2*R*U 32-bit loads in flight instead of draining to zero every iteration.
template <int R, int WARPS, int U, bool PIPE>
__global__ void __launch_bounds__(WARPS * 32)
centroid_accum_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ cls, int nclass, int64_t n, int64_t D,
                      int64_t hw, float* __restrict__ sums) {
  using acc_t = typename AccT<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  acc_t* acc = reinterpret_cast<acc_t*>(smem_raw);   // [WARPS][nclass][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t groups = (D + R - 1) / R;
  const int64_t items = n * groups;
  acc_t* my = acc + (size_t)warp * nclass * 32 + lane;
  constexpr int64_t STEP = (int64_t)WARPS * 32 * U;

  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t img = item / groups;
    const int64_t d0 = (item - img * groups) * R;
    const uint8_t* cl = cls + img * hw;
    const float* rows[R];
    bool row_ok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row_ok[r] = d0 + r < D;
      rows[r] = feat + (img * D + (row_ok[r] ? d0 + r : d0)) * hw;
    }
    const int64_t first = warp * 32 + lane;
    AccumBatch<R, U> a, b;
    accum_load<R, WARPS, U>(a, rows, cl, first, hw);          // in flight while the accumulators are cleared
    acc_t zero;
    if constexpr (R == 4) zero = make_float4(0.f, 0.f, 0.f, 0.f); else zero = make_float2(0.f, 0.f);
    for (int c = 0; c < nclass; ++c) my[c * 32] = zero;
    // (own lane-private cells only: no barrier needed before the main loop)

    if constexpr (PIPE) {
      // bases are warp-uniform up to +lane, so the loop trip count is uniform within a warp
      for (int64_t base = first; base - lane < hw; base += 2 * STEP) {
        accum_load<R, WARPS, U>(b, rows, cl, base + STEP, hw);
        accum_apply<R, U>(a, my, nclass);
        accum_load<R, WARPS, U>(a, rows, cl, base + 2 * STEP, hw);
        accum_apply<R, U>(b, my, nclass);
      }
    } else {
      for (int64_t base = first; base - lane < hw; base += STEP) {
        accum_apply<R, U>(a, my, nclass);
        accum_load<R, WARPS, U>(a, rows, cl, base + STEP, hw);
      }
    }
    __syncthreads();
    // cross-warp then cross-lane reduction; classes are dealt round-robin to warps
    for (int c = warp; c < nclass; c += WARPS) {
      float s[R];
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const acc_t v = acc[((size_t)w * nclass + c) * 32 + lane];
        if constexpr (R == 4) {
          s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        } else {
          s[0] += v.x; s[1] += v.y;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = warp_sum(s[r]);
      if (lane == 0) {
        float* o = sums + (img * nclass + c) * D + d0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (row_ok[r]) o[r] = s[r];
      }
    }
    __syncthreads();
  }
}

template <int R, int WARPS, int U, bool PIPE>
static int launch_accum(const float* feat, const uint8_t* cls, int nclass, int64_t n, int64_t D, int64_t hw, float* sums,
                        cudaStream_t st) {
  auto kern = centroid_accum_kernel<R, WARPS, U, PIPE>;
  const size_t smem = (size_t)WARPS * nclass * 32 * sizeof(typename AccT<R>::type);
  static size_t configured = 0;
  static int blocks_per_sm = 0;
  if (configured != smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("centroid_accum: cannot reserve %zu bytes of shared memory", smem);
      return DIGA_ERR_CUDA;
    }
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, WARPS * 32, smem);
    blocks_per_sm = b > 0 ? b : 1;
    configured = smem;
  }
  const int64_t items = n * ((D + R - 1) / R);
  int64_t grid = (int64_t)sm_count() * blocks_per_sm;
  if (grid > items) grid = items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(feat, cls, nclass, n, D, hw, sums);
  DIGA_CHECK_LAUNCH("centroid_accum_kernel");
  return DIGA_OK;
}

// ------------------------------------------------------------------------------------------------
// means
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_means_kernel(const float* __restrict__ sums, const int32_t* __restrict__ counts, int64_t C, int64_t D,
                      int64_t hw, float* __restrict__ vec, float* __restrict__ vecsum, uint8_t* __restrict__ valid) {
  __shared__ float red[BLOCK / 32];
  const int64_t nc = (int64_t)blockIdx.y * C + blockIdx.x;
  const int cnt = counts[nc];
  // calc_centroids.py:129,141: avgpool(feat*mask) / avgpool(mask) == (sum/hw) / (count/hw)
  const float frac = (float)cnt / (float)hw;
  float part = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) {
    float v = 0.f;
    if (cnt > 0) v = (sums[nc * D + d] / (float)hw) / frac;
    vec[nc * D + d] = v;
    part += v;
  }
  const float tot = block_sum<BLOCK>(part, red);
  if (threadIdx.x == 0) {
    vecsum[nc] = tot;
    valid[nc] = (cnt >= 5) ? 1 : 0;   // :134 (frac == 0) and :136 (< 5 pixels) skips
  }
}

// ------------------------------------------------------------------------------------------------
// update (calc_centroids.py:147-164), arithmetic mirrored op for op (separately rounded mul/add/div)
// ------------------------------------------------------------------------------------------------
struct UpdateRule {
  int mode;          // DIGA_UPDATE_*
  int start_mean;
  float m, one_minus_m;
};

__device__ __forceinline__ bool rule_is_mean(const UpdateRule& r, float num) {
  return r.mode == DIGA_UPDATE_MEAN || (r.start_mean && num < 100.f);   // :150
}
__device__ __forceinline__ float rule_apply(const UpdateRule& r, bool mean, float obj, float num, float v) {
  if (mean) {
    const float t = __fadd_rn(__fmul_rn(obj, num), v);                  // :158
    return __fdiv_rn(t, __fadd_rn(num, 1.f));                           // :159-160
  }
  return __fadd_rn(__fmul_rn(obj, r.one_minus_m), __fmul_rn(r.m, v));   // :153-154
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_kernel(const float* __restrict__ vec, const float* __restrict__ vecsum, const uint8_t* __restrict__ valid,
                       int64_t n, int64_t C, int64_t D, float* __restrict__ obj, float* __restrict__ objnum, UpdateRule rule) {
  const int64_t c = blockIdx.x;
  float num = objnum[c];
  float* o = obj + c * D;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t nc = i * C + c;
    if (valid != nullptr && !valid[nc]) continue;
    if (vecsum[nc] == 0.f) continue;                                    // :148
    const bool mean = rule_is_mean(rule, num);
    const float* v = vec + nc * D;
    for (int64_t d = threadIdx.x; d < D; d += BLOCK) o[d] = rule_apply(rule, mean, o[d], num, v[d]);
    num = fminf(__fadd_rn(num, 1.f), 3000.f);                           // :155-156 / :159,161
  }
  if (threadIdx.x == 0) objnum[c] = num;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_single_kernel(const float* __restrict__ v, int64_t id, int64_t D, float* __restrict__ obj,
                              float* __restrict__ objnum, UpdateRule rule) {
  __shared__ float red[BLOCK / 32];
  __shared__ float s_tot;
  float part = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) part += v[d];
  const float tot = block_sum<BLOCK>(part, red);
  if (threadIdx.x == 0) s_tot = tot;
  __syncthreads();
  if (s_tot == 0.f) return;
  const float num = objnum[id];
  const bool mean = rule_is_mean(rule, num);
  float* o = obj + id * D;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) o[d] = rule_apply(rule, mean, o[d], num, v[d]);
  __syncthreads();
  if (threadIdx.x == 0) objnum[id] = fminf(__fadd_rn(num, 1.f), 3000.f);
}

// acc[c][0..D) += sum over valid images of vec[n][c][:]; acc[c][D] += number of such images.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_reduce_images_kernel(const float* __restrict__ vec, const float* __restrict__ vecsum,
                              const uint8_t* __restrict__ valid, int64_t n, int64_t C, int64_t D, float* __restrict__ acc) {
  const int64_t c = blockIdx.x;
  float* a = acc + c * (D + 1);
  int used = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t nc = i * C + c;
    if (valid != nullptr && !valid[nc]) continue;
    if (vecsum[nc] == 0.f) continue;
    ++used;
    const float* v = vec + nc * D;
    for (int64_t d = threadIdx.x; d < D; d += BLOCK) a[d] += v[d];
  }
  if (threadIdx.x == 0) a[D] += (float)used;
}

// process_label (util/utils.py:158-163): onehot[b][k][p] = (k == (label < C ? long(label) : C)), k in [0, C].
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
onehot_labels_kernel(const float* __restrict__ label, int64_t C, int64_t hw, float* __restrict__ onehot) {
  const int64_t b = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * BLOCK + threadIdx.x; p < hw; p += (int64_t)gridDim.x * BLOCK) {
    const float lf = __ldg(label + b * hw + p);
    const long long id = lf < (float)C ? (long long)lf : (long long)C;
    float* o = onehot + b * (C + 1) * hw + p;
    for (int64_t k = 0; k <= C; ++k) o[k * hw] = (k == id) ? 1.f : 0.f;
  }
}

static UpdateRule make_rule(int mode, int start_mean, double momentum) {
  UpdateRule r;
  r.mode = mode;
  r.start_mean = start_mean;
  r.m = (float)momentum;
  r.one_minus_m = (float)(1.0 - momentum);   // Python evaluates (1 - momentum) in double, then torch rounds to fp32
  return r;
}

}  // namespace diga

extern "C" {

int diga_centroid_assign(const float* logits, const float* labels, int64_t n, int64_t C, int64_t hw, uint8_t* cls,
                         int32_t* counts, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(logits && cls && counts, DIGA_ERR_INVALID, "centroid_assign: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "centroid_assign: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && n <= 65535 && hw >= 0, DIGA_ERR_INVALID, "centroid_assign: bad sizes");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(labels, 4) && aligned(counts, 4), DIGA_ERR_MISALIGNED,
               "centroid_assign: misaligned pointer");
  if (n == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts, 0, (size_t)n * C * sizeof(int32_t), st);
  if (hw == 0) return DIGA_OK;
  constexpr int BLOCK = 256;
  int64_t gx = (hw + BLOCK - 1) / BLOCK;
  const int64_t cap = ((int64_t)sm_count() * 8 + n - 1) / n;
  if (gx > cap) gx = cap;
  centroid_assign_kernel<BLOCK><<<dim3((unsigned)gx, (unsigned)n), BLOCK, 0, st>>>(logits, labels, (int)C, hw, cls, counts);
  DIGA_CHECK_LAUNCH("centroid_assign_kernel");
  return DIGA_OK;
}

int diga_centroid_accum(const float* feat, const uint8_t* cls, int64_t n, int64_t D, int64_t C, int64_t hw, float* sums,
                        diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(feat && cls && sums, DIGA_ERR_INVALID, "centroid_accum: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "centroid_accum: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && D >= 0 && hw >= 0, DIGA_ERR_INVALID, "centroid_accum: bad sizes");
  DIGA_REQUIRE(aligned(feat, 4) && aligned(sums, 4), DIGA_ERR_MISALIGNED, "centroid_accum: misaligned pointer");
  if (n == 0 || D == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int variant = tunable("accum_variant", 0);
  switch (variant) {
    case 1: return launch_accum<4, 4, 4, false>(feat, cls, (int)C, n, D, hw, sums, st);
    case 2: return launch_accum<4, 4, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 3: return launch_accum<4, 8, 4, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 4: return launch_accum<2, 4, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 5: return launch_accum<4, 2, 4, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 6: return launch_accum<4, 2, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 7: return launch_accum<2, 8, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    default: return launch_accum<4, 4, 4, true>(feat, cls, (int)C, n, D, hw, sums, st);
  }
}

int diga_centroid_means(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw, float* vec,
                        float* vecsum, uint8_t* valid, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(sums && counts && vec && vecsum && valid, DIGA_ERR_INVALID, "centroid_means: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && n >= 0 && n <= 65535 && D >= 0 && hw > 0, DIGA_ERR_INVALID,
               "centroid_means: bad sizes");
  if (n == 0) return DIGA_OK;
  centroid_means_kernel<256><<<dim3((unsigned)C, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(sums, counts, C, D, hw, vec,
                                                                                                vecsum, valid);
  DIGA_CHECK_LAUNCH("centroid_means_kernel");
  return DIGA_OK;
}

int diga_centroid_update(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C, int64_t D,
                         float* objective_vectors, float* objective_num, int mode, int start_mean, double momentum,
                         diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vec && vecsum && objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_update: null pointer");
  DIGA_REQUIRE(mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  DIGA_REQUIRE(C >= 1 && n >= 0 && D >= 0, DIGA_ERR_INVALID, "centroid_update: bad sizes");
  if (n == 0) return DIGA_OK;
  centroid_update_kernel<256><<<(unsigned)C, 256, 0, (cudaStream_t)stream>>>(
      vec, vecsum, valid, n, C, D, objective_vectors, objective_num, make_rule(mode, start_mean, momentum));
  DIGA_CHECK_LAUNCH("centroid_update_kernel");
  return DIGA_OK;
}

int diga_centroid_update_single(const float* vector, int64_t id, int64_t C, int64_t D, float* objective_vectors,
                                float* objective_num, int mode, int start_mean, double momentum, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vector && objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_update_single: null pointer");
  DIGA_REQUIRE(mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  DIGA_REQUIRE(id >= 0 && id < C && D >= 0, DIGA_ERR_INVALID, "centroid_update_single: class id %lld outside [0,%lld)",
               (long long)id, (long long)C);
  centroid_update_single_kernel<256><<<1, 256, 0, (cudaStream_t)stream>>>(vector, id, D, objective_vectors, objective_num,
                                                                          make_rule(mode, start_mean, momentum));
  DIGA_CHECK_LAUNCH("centroid_update_single_kernel");
  return DIGA_OK;
}

int diga_onehot_labels(const float* label, int64_t B, int64_t C, int64_t hw, float* onehot, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(label && onehot, DIGA_ERR_INVALID, "onehot_labels: null pointer");
  DIGA_REQUIRE(B >= 0 && B <= 65535 && C >= 1 && hw >= 0, DIGA_ERR_INVALID, "onehot_labels: bad sizes");
  if (B == 0 || hw == 0) return DIGA_OK;
  int64_t gx = (hw + 255) / 256;
  const int64_t cap = ((int64_t)sm_count() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  onehot_labels_kernel<256><<<dim3((unsigned)gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(label, C, hw, onehot);
  DIGA_CHECK_LAUNCH("onehot_labels_kernel");
  return DIGA_OK;
}

int diga_centroid_reduce_images(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C, int64_t D,
                                float* acc, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vec && vecsum && acc, DIGA_ERR_INVALID, "centroid_reduce_images: null pointer");
  DIGA_REQUIRE(C >= 1 && n >= 0 && D >= 0, DIGA_ERR_INVALID, "centroid_reduce_images: bad sizes");
  if (n == 0) return DIGA_OK;
  centroid_reduce_images_kernel<256><<<(unsigned)C, 256, 0, (cudaStream_t)stream>>>(vec, vecsum, valid, n, C, D, acc);
  DIGA_CHECK_LAUNCH("centroid_reduce_images_kernel");
  return DIGA_OK;
}

}  // extern "C"
