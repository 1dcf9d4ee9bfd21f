// a2 — cross-domain ClassMix: class-presence bitmap (device half of torch.unique) and LUT blend.
// Replaces train_DiGA_gta2city_self_training.py:259-275 / :306-325, warm_up.py:240-259 and
// calc_centroids.py:47-58 of the reference.  The class choice itself (random.sample on the sorted
// list of present classes) stays on the host so that the Python `random` stream is consumed exactly
// as the reference consumes it; only a 32-byte bitmap per image crosses the bus.
//
// Blend traffic per pixel: 8 (label) + 4*ch*3 (a, b, mix) [+4 mask] [+16 tlabel in / mixlabel out].
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

// ------------------------------------------------------------------------------------------------
// presence: bitmap[b][8] |= 1 << label
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
class_presence_kernel(const int64_t* __restrict__ slabel, int64_t hw, uint32_t* __restrict__ bitmap,
                      uint32_t* __restrict__ flags) {
  __shared__ uint32_t bm[8];
  __shared__ uint32_t bad;
  if (threadIdx.x < 8) bm[threadIdx.x] = 0u;
  if (threadIdx.x == 8) bad = 0u;
  __syncthreads();
  const int64_t b = blockIdx.y;
  const int64_t* lab = slabel + b * hw;
  auto mark = [&](int64_t v) {
    if ((uint64_t)v > 255ull) {
      bad = 1u;   // benign race: every writer stores 1
      return;
    }
    const uint32_t word = (uint32_t)v >> 5, bit = 1u << ((uint32_t)v & 31u);
    // Labels are piecewise constant, so after the first few pixels the bit is already set and the
    // shared-memory atomic (which would serialise 32 lanes on one word) is skipped.
    if (!(((volatile uint32_t*)bm)[word] & bit)) atomicOr(&bm[word], bit);
  };
  const int64_t pairs = hw >> 1;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(lab) & 15) == 0;
  if (vec_ok) {
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * BLOCK) {
      const longlong2 v = ld_stream_i64x2(lab + 2 * i);
      mark(v.x);
      mark(v.y);
    }
    if ((hw & 1) && blockIdx.x == 0 && threadIdx.x == 0) mark(ld_stream_i64(lab + hw - 1));
  } else {
    for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < hw; i += (int64_t)gridDim.x * BLOCK)
      mark(ld_stream_i64(lab + i));
  }
  __syncthreads();
  if (threadIdx.x < 8 && bm[threadIdx.x]) atomicOr(&bitmap[b * 8 + threadIdx.x], bm[threadIdx.x]);
  if (threadIdx.x == 8 && bad) atomicOr(flags, 1u);
}

// ------------------------------------------------------------------------------------------------
// blend
// ------------------------------------------------------------------------------------------------
// mix = a*(1-m) + b*m evaluated with separately rounded multiplies and add, exactly like
// torch.mul(a, 1 - mask) + torch.mul(b, mask): bit-exact including -0.0, inf and NaN propagation.
__device__ __forceinline__ float blend(float a, float b, float m) {
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, m)), __fmul_rn(b, m));
}

// The per-image class selection travels as a kernel argument (256-bit bitmap per image, up to 16 images per
// launch): no device-side LUT buffer, no H2D copy.
constexpr int kBlendImagesPerLaunch = 16;
struct BlendSelection {
  uint32_t bits[kBlendImagesPerLaunch][8];
};

template <int VEC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
classmix_blend_kernel(const int64_t* __restrict__ slabel, const __grid_constant__ BlendSelection sel, int64_t img0,
                      const float* __restrict__ a, const float* __restrict__ b, const int64_t* __restrict__ tlabel,
                      int channels, int64_t hw, float* __restrict__ mask, float* __restrict__ mix,
                      int64_t* __restrict__ mixlabel) {
  __shared__ uint8_t s_lut[256];
  const int64_t img = img0 + blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += BLOCK) s_lut[i] = (sel.bits[blockIdx.y][i >> 5] >> (i & 31)) & 1u;
  __syncthreads();
  const int64_t groups = hw / VEC;
  const int64_t* sl = slabel + img * hw;
  for (int64_t g = (int64_t)blockIdx.x * BLOCK + threadIdx.x; g < groups; g += (int64_t)gridDim.x * BLOCK) {
    const int64_t p = g * VEC;
    int64_t lab[VEC];
    if constexpr (VEC >= 2) {
#pragma unroll
      for (int v = 0; v < VEC; v += 2) {
        const longlong2 t = ld_stream_i64x2(sl + p + v);
        lab[v] = t.x;
        lab[v + 1] = t.y;
      }
    } else {
      lab[0] = ld_stream_i64(sl + p);
    }
    Vec<VEC> m;
#pragma unroll
    for (int v = 0; v < VEC; ++v) m.v[v] = ((uint64_t)lab[v] <= 255ull && s_lut[lab[v]]) ? 1.0f : 0.0f;
    if (mask) st_stream<VEC>(mask + img * hw + p, m);
    if (mix) {
      for (int c = 0; c < channels; ++c) {
        const int64_t o = (img * channels + c) * hw + p;
        const Vec<VEC> va = ld_stream<VEC>(a + o);
        const Vec<VEC> vb = ld_stream<VEC>(b + o);
        Vec<VEC> r;
#pragma unroll
        for (int v = 0; v < VEC; ++v) r.v[v] = blend(va.v[v], vb.v[v], m.v[v]);
        st_stream<VEC>(mix + o, r);
      }
    }
    if (mixlabel) {
      const int64_t* tl = tlabel + img * hw + p;
      int64_t out[VEC];
      if constexpr (VEC >= 2) {
#pragma unroll
        for (int v = 0; v < VEC; v += 2) {
          const longlong2 t = ld_stream_i64x2(tl + v);
          out[v] = m.v[v] != 0.f ? lab[v] : t.x;
          out[v + 1] = m.v[v + 1] != 0.f ? lab[v + 1] : t.y;
        }
#pragma unroll
        for (int v = 0; v < VEC; v += 2) st_stream_i64x2(mixlabel + img * hw + p + v, out[v], out[v + 1]);
      } else {
        const int64_t t = ld_stream_i64(tl);
        st_stream_i64(mixlabel + img * hw + p, m.v[0] != 0.f ? lab[0] : t);
      }
    }
  }
}

}  // namespace diga

extern "C" {

int diga_class_presence(const int64_t* slabel, int64_t B, int64_t hw, uint32_t* bitmap, uint32_t* flags,
                        diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(slabel && bitmap && flags, DIGA_ERR_INVALID, "class_presence: null pointer");
  DIGA_REQUIRE(B >= 0 && hw >= 0 && B <= 65535, DIGA_ERR_INVALID, "class_presence: bad batch %lld", (long long)B);
  DIGA_REQUIRE(aligned(slabel, 8) && aligned(bitmap, 4) && aligned(flags, 4), DIGA_ERR_MISALIGNED,
               "class_presence: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return DIGA_OK;
  cudaMemsetAsync(bitmap, 0, (size_t)B * 8 * sizeof(uint32_t), st);
  cudaMemsetAsync(flags, 0, sizeof(uint32_t), st);
  if (hw == 0) return DIGA_OK;
  constexpr int BLOCK = 256;
  // ~4 label pairs per thread keeps >= 8 loads in flight per warp; cap the x-grid per image.
  int64_t gx = ((hw >> 1) + BLOCK * 4 - 1) / (BLOCK * 4);
  const int64_t cap = ((int64_t)sm_count() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  class_presence_kernel<BLOCK><<<dim3((unsigned)gx, (unsigned)B), BLOCK, 0, st>>>(slabel, hw, bitmap, flags);
  DIGA_CHECK_LAUNCH("class_presence_kernel");
  return DIGA_OK;
}

int diga_classmix_blend(const int64_t* slabel, const uint8_t* lut_host, const float* a, const float* b,
                        const int64_t* tlabel, int64_t B, int64_t channels, int64_t hw, float* mask, float* mix,
                        int64_t* mixlabel, diga_stream_t stream) {
  using namespace diga;
  if (B == 0 || hw == 0) return DIGA_OK;      // empty batch: pointers may be NULL
  DIGA_REQUIRE(slabel && lut_host, DIGA_ERR_INVALID, "classmix_blend: null label/lut");
  DIGA_REQUIRE(!mix || (a && b), DIGA_ERR_INVALID, "classmix_blend: mix needs both images");
  DIGA_REQUIRE(!mixlabel || tlabel, DIGA_ERR_INVALID, "classmix_blend: mixlabel needs tlabel");
  DIGA_REQUIRE(B >= 0 && hw >= 0 && channels >= 0, DIGA_ERR_INVALID, "classmix_blend: bad sizes");
  DIGA_REQUIRE(aligned(slabel, 8) && aligned(tlabel, 8) && aligned(mixlabel, 8) && aligned(a, 4) && aligned(b, 4) &&
                   aligned(mix, 4) && aligned(mask, 4),
               DIGA_ERR_MISALIGNED, "classmix_blend: misaligned pointer");
  if (B == 0 || hw == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int BLOCK = 256;
  const bool a16 = aligned(slabel, 16) && aligned(tlabel, 16) && aligned(mixlabel, 16) && aligned(a, 16) &&
                   aligned(b, 16) && aligned(mix, 16) && aligned(mask, 16);
  int vec = ((hw % 4) == 0 && a16) ? tunable("cm_vec", 4) : 1;
  if (vec == 2 && (hw % 2) != 0) vec = 1;
  const int64_t groups = hw / vec;
  for (int64_t img0 = 0; img0 < B; img0 += kBlendImagesPerLaunch) {
    const int64_t nb = B - img0 < kBlendImagesPerLaunch ? B - img0 : kBlendImagesPerLaunch;
    BlendSelection sel;
    for (int i = 0; i < kBlendImagesPerLaunch; ++i)
      for (int wd = 0; wd < 8; ++wd) {
        uint32_t bits = 0;
        if (i < nb)
          for (int k = 0; k < 32; ++k) bits |= (lut_host[(img0 + i) * 256 + wd * 32 + k] ? 1u : 0u) << k;
        sel.bits[i][wd] = bits;
      }
    int64_t gx = (groups + BLOCK - 1) / BLOCK;
    const int64_t cap = ((int64_t)sm_count() * 8 * tunable("cm_waves", 2) + nb - 1) / nb;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)nb);
    if (vec == 4)
      classmix_blend_kernel<4, BLOCK><<<grid, BLOCK, 0, st>>>(slabel, sel, img0, a, b, tlabel, (int)channels, hw, mask, mix, mixlabel);
    else if (vec == 2)
      classmix_blend_kernel<2, BLOCK><<<grid, BLOCK, 0, st>>>(slabel, sel, img0, a, b, tlabel, (int)channels, hw, mask, mix, mixlabel);
    else
      classmix_blend_kernel<1, BLOCK><<<grid, BLOCK, 0, st>>>(slabel, sel, img0, a, b, tlabel, (int)channels, hw, mask, mix, mixlabel);
    DIGA_CHECK_LAUNCH("classmix_blend_kernel");
  }
  return DIGA_OK;
}

}  // extern "C"
