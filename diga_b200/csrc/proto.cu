// a5 — prototype rectification: per-pixel L2 distance to the C class centroids and softmax(-dist).
// Replaces Class_Features.feat_centroid_distance / get_centroid_weight / get_centroid_distance
// (calc_centroids.py:166-180 of the reference), which re-reads the feature map 19 times.
//
// This file holds the FP32 CUDA-core kernel: exact difference form sqrt(sum_d (c_d - f_d)^2), one pass
// over the features (D*4 B per feature pixel).  It is the path for shapes the tensor-core kernel
// (proto_umma.cu) does not take, and the numerical yard-stick that kernel is tested against.
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);
int proto_umma_supported(int64_t n, int64_t D, int64_t C, int64_t hw);
int proto_umma_launch(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw, float* dist,
                      float* weight, void* workspace, int prepared, cudaStream_t st);
int proto_umma_prepare(const float* centroids, int64_t D, int64_t C, void* workspace, cudaStream_t st);
size_t proto_umma_workspace_bytes(int64_t C, int64_t D);

// Epilogue shared by both kernels: dist -> (dist, softmax(-dist)) for one pixel.
template <int C, bool PAD>
__device__ __forceinline__ void proto_epilogue(float (&d)[C], int nclass, float* dist, float* weight, int64_t plane) {
  float dmin = d[0];
#pragma unroll
  for (int c = 1; c < C; ++c)
    if (!PAD || c < nclass) dmin = fminf(dmin, d[c]);
  if (dist) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) dist[c * plane] = d[c];
  }
  if (weight) {
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        d[c] = fast_exp(dmin - d[c]);   // softmax(-dist), shifted by its maximum -dmin
        S += d[c];
      }
    const float inv = 1.0f / S;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) weight[c * plane] = d[c] * inv;
  }
}

// One thread per pixel, DK channels of all centroids staged in shared memory per step.
template <int C, bool PAD, int BLOCK, int DK>
__global__ void __launch_bounds__(BLOCK)
proto_distance_fp32_kernel(const float* __restrict__ feat, const float* __restrict__ cen, int nclass, int64_t D, int64_t hw,
                           float* __restrict__ dist, float* __restrict__ weight) {
  __shared__ float s_cen[C][DK];
  const int64_t img = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  const bool ok = p < hw;
  const float* f = feat + img * D * hw + (ok ? p : 0);
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  for (int64_t d0 = 0; d0 < D; d0 += DK) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * DK; i += BLOCK) {
      const int c = i / DK, k = i - c * DK;
      s_cen[c][k] = (c < nclass && d0 + k < D) ? __ldg(cen + (int64_t)c * D + d0 + k) : 0.f;
    }
    __syncthreads();
    float x[DK];
#pragma unroll
    for (int k = 0; k < DK; ++k) x[k] = (ok && d0 + k < D) ? ld_stream<1>(f + (d0 + k) * hw).v[0] : 0.f;
    // two-level summation: a fresh partial per DK-channel step, then one add into the running sum.  A single
    // 2048-term fp32 chain would be ~15x less accurate than torch's tree reduction (measured: weight error
    // 1.2e-5 vs 8e-7 at D=2048).
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < DK; ++k) {
          const float df = s_cen[c][k] - x[k];
          part = fmaf(df, df, part);
        }
        acc[c] += part;
      }
  }
  if (!ok) return;
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = sqrtf(acc[c]);
  const int64_t o = img * nclass * hw + p;
  proto_epilogue<C, PAD>(acc, nclass, dist ? dist + o : nullptr, weight ? weight + o : nullptr, hw);
}

}  // namespace diga

extern "C" {

size_t diga_proto_workspace_bytes(int64_t C, int64_t D) { return diga::proto_umma_workspace_bytes(C, D); }

static int proto_distance_impl(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw, float* dist,
                               float* weight, void* workspace, int prepared, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(feat && centroids, DIGA_ERR_INVALID, "proto_distance: null input");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "proto_distance: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && n <= 65535 && D >= 1 && hw >= 0, DIGA_ERR_INVALID, "proto_distance: bad sizes");
  DIGA_REQUIRE(aligned(feat, 4) && aligned(centroids, 4) && aligned(dist, 4) && aligned(weight, 4), DIGA_ERR_MISALIGNED,
               "proto_distance: misaligned pointer");
  if (n == 0 || hw == 0 || (!dist && !weight)) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (tunable("proto_path", 0) != 1 && workspace != nullptr && proto_umma_supported(n, D, C, hw))
    return proto_umma_launch(feat, centroids, n, D, C, hw, dist, weight, workspace, prepared, st);
  constexpr int BLOCK = 128, DK = 16;
  dim3 grid((unsigned)((hw + BLOCK - 1) / BLOCK), (unsigned)n);
  DIGA_DISPATCH_C(C, {
    proto_distance_fp32_kernel<kC, kPad, BLOCK, DK><<<grid, BLOCK, 0, st>>>(feat, centroids, (int)C, D, hw, dist, weight);
  });
  DIGA_CHECK_LAUNCH("proto_distance_fp32_kernel");
  return DIGA_OK;
}

int diga_proto_distance(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw, float* dist,
                        float* weight, void* workspace, diga_stream_t stream) {
  return proto_distance_impl(feat, centroids, n, D, C, hw, dist, weight, workspace, 0, stream);
}

int diga_proto_prepare(const float* centroids, int64_t C, int64_t D, void* workspace, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(centroids && workspace, DIGA_ERR_INVALID, "proto_prepare: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && D >= 1, DIGA_ERR_INVALID, "proto_prepare: bad sizes");
  if (!proto_umma_supported(1, D, C, 1)) return DIGA_OK;      // the FP32 kernel reads the centroids directly
  return proto_umma_prepare(centroids, D, C, workspace, (cudaStream_t)stream);
}

int diga_proto_distance_prepared(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw,
                                 float* dist, float* weight, void* workspace, diga_stream_t stream) {
  return proto_distance_impl(feat, centroids, n, D, C, hw, dist, weight, workspace, 1, stream);
}

}  // extern "C"
