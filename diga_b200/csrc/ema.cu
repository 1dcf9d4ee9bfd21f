// f4 — EMA teacher update as a multi-tensor kernel.
// Replaces util/utils.py:103-116 of the reference (`update_teacher_params`): a Python loop over ~500 parameter tensors,
// three tiny launches each.  Here up to 512 tensors share one launch: the pointer table travels as a 14 KB kernel argument
// (CUDA >= 12.1 accepts 32 KB of parameters), so a ResNet-101 teacher is ONE launch and a block finds its tensor by
// binary search in the block-offset table.
//   teacher = alpha * teacher + (1 - alpha) * student      (separately rounded mul, mul, add — bit-exact with torch)
// Traffic: 12 B per parameter element.
#include "common.cuh"

namespace diga {

constexpr int kEmaTensorsPerLaunch = 512;
constexpr int kEmaBlock = 256;
constexpr int kEmaElemsPerBlock = kEmaBlock * 4 * 4;   // 4 float4 per thread

struct EmaTable {
  float* teacher[kEmaTensorsPerLaunch];
  const float* student[kEmaTensorsPerLaunch];
  int64_t numel[kEmaTensorsPerLaunch];
  int block_start[kEmaTensorsPerLaunch + 1];
  int count;
};

__device__ __forceinline__ float ema1(float t, float s, float a, float oma) { return __fadd_rn(__fmul_rn(a, t), __fmul_rn(oma, s)); }

__global__ void __launch_bounds__(kEmaBlock)
ema_update_kernel(const __grid_constant__ EmaTable tab, float alpha, float one_minus_alpha) {
  int i = 0, hi = tab.count - 1;                           // largest i with block_start[i] <= blockIdx.x
  while (i < hi) {
    const int mid = (i + hi + 1) >> 1;
    if ((int)blockIdx.x >= tab.block_start[mid]) i = mid;
    else hi = mid - 1;
  }
  float* __restrict__ t = tab.teacher[i];
  const float* __restrict__ s = tab.student[i];
  const int64_t n = tab.numel[i];
  const int64_t base = (int64_t)(blockIdx.x - tab.block_start[i]) * kEmaElemsPerBlock;
  const bool vec = ((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(s)) & 15) == 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t e = base + ((int64_t)k * kEmaBlock + threadIdx.x) * 4;
    if (vec && e + 4 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(t + e);
      const Vec<4> b = ld_stream<4>(s + e);
      float4 r;
      r.x = ema1(a.x, b.v[0], alpha, one_minus_alpha);
      r.y = ema1(a.y, b.v[1], alpha, one_minus_alpha);
      r.z = ema1(a.z, b.v[2], alpha, one_minus_alpha);
      r.w = ema1(a.w, b.v[3], alpha, one_minus_alpha);
      *reinterpret_cast<float4*>(t + e) = r;
    } else {
      for (int64_t j = e; j < e + 4 && j < n; ++j) t[j] = ema1(t[j], s[j], alpha, one_minus_alpha);
    }
  }
}

}  // namespace diga

extern "C" int diga_ema_update(float* const* teacher_host, const float* const* student_host, const int64_t* numel_host,
                               int64_t count, double alpha, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(count >= 0 && (count == 0 || (teacher_host && student_host && numel_host)), DIGA_ERR_INVALID,
               "ema_update: null table");
  const float a = (float)alpha, oma = (float)(1.0 - alpha);   // Python evaluates (1 - alpha) in double; torch rounds both to fp32
  // `first` advances to wherever the inner loop stopped: empty tensors are skipped without taking a table slot, so a
  // fixed stride of kEmaTensorsPerLaunch would visit (and update) the tensors past the stride twice.
  for (int64_t first = 0; first < count;) {
    EmaTable tab;
    int nt = 0, blocks = 0;
    int64_t i = first;
    for (; i < count && nt < kEmaTensorsPerLaunch; ++i) {
      DIGA_REQUIRE(numel_host[i] >= 0, DIGA_ERR_INVALID, "ema_update: negative size");
      if (numel_host[i] == 0) continue;
      DIGA_REQUIRE(teacher_host[i] && student_host[i], DIGA_ERR_INVALID, "ema_update: null tensor %lld", (long long)i);
      DIGA_REQUIRE(aligned(teacher_host[i], 4) && aligned(student_host[i], 4), DIGA_ERR_MISALIGNED, "ema_update: misaligned tensor");
      tab.teacher[nt] = teacher_host[i];
      tab.student[nt] = student_host[i];
      tab.numel[nt] = numel_host[i];
      tab.block_start[nt] = blocks;
      blocks += (int)((numel_host[i] + kEmaElemsPerBlock - 1) / kEmaElemsPerBlock);
      ++nt;
    }
    first = i;
    if (nt == 0) continue;
    tab.block_start[nt] = blocks;
    tab.count = nt;
    ema_update_kernel<<<(unsigned)blocks, kEmaBlock, 0, (cudaStream_t)stream>>>(tab, a, oma);
    DIGA_CHECK_LAUNCH("ema_update_kernel");
  }
  return DIGA_OK;
}
