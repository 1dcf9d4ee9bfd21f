// f3, reader half — label maps as the Cityscapes loader hands them to the training step.
// Replaces util/loader/CityLoader.py:93-95 (PIL NEAREST resize of the label / pseudo-label image to crop_size) and
// :115-132 (re-assignment `label_copy[label == k] = v` for 19-34 values of k: one full-image pass per value) of the
// reference, after the PNG has been decoded: one gather through the resize tables plus a 256-entry look-up, uint8 in,
// int64 out (the dtype the loss and ClassMix kernels take).  The source-index tables reproduce Pillow's NEAREST
// transform (ImagingScaleAffine: xo = scale * 0.5, xin = (int)xo, xo += scale in double) and are built on the host once
// per geometry (diga_b200/util/labels.py); the kernel is a pure byte gather: 1 B read (cached), 8 B written per pixel.
#include "common.cuh"

namespace diga {

struct Lut256 {
  uint8_t v[256];
};

// A CTA walks output rows r = blockIdx.x, blockIdx.x + gridDim.x, .. (one look-up table set-up per CTA, not per row): no
// index division per pixel; a thread owns four adjacent output pixels per iteration (two 128-bit stores).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
label_resize_remap_kernel(const uint8_t* __restrict__ src, int64_t w0, const int32_t* __restrict__ ytab,
                          const int32_t* __restrict__ xtab, int H, int W, int64_t rows, int64_t plane0, const Lut256 lut,
                          int64_t* __restrict__ out) {
  __shared__ uint8_t slut[256];
  for (int i = threadIdx.x; i < 256; i += BLOCK) slut[i] = lut.v[i];
  __syncthreads();
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {   // r = img * H + y
    const int y = (int)(r % H);
    const int64_t img = r / H;
    const uint8_t* row = src + img * plane0 + (int64_t)__ldg(ytab + y) * w0;
    int64_t* orow = out + r * W;
    const bool vec = (reinterpret_cast<uintptr_t>(orow) & 15) == 0;
    for (int x = threadIdx.x * 4; x < W; x += BLOCK * 4) {
      int64_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (x + k < W) ? (int64_t)slut[__ldg(row + __ldg(xtab + x + k))] : 0;
      if (vec && x + 3 < W) {
        st_stream_i64x2(orow + x, v[0], v[1]);
        st_stream_i64x2(orow + x + 2, v[2], v[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (x + k < W) orow[x + k] = v[k];
      }
    }
  }
}

}  // namespace diga

extern "C" int diga_label_resize_remap(const uint8_t* src, int64_t n, int64_t h0, int64_t w0, const int32_t* ytab,
                                       const int32_t* xtab, int64_t H, int64_t W, const uint8_t* lut_host, int64_t* out,
                                       diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(src && ytab && xtab && lut_host && out, DIGA_ERR_INVALID, "label_resize_remap: null pointer");
  DIGA_REQUIRE(n >= 0 && h0 >= 1 && w0 >= 1 && H >= 1 && W >= 1, DIGA_ERR_INVALID, "label_resize_remap: bad sizes");
  DIGA_REQUIRE(aligned(ytab, 4) && aligned(xtab, 4) && aligned(out, 8), DIGA_ERR_MISALIGNED, "label_resize_remap: misaligned pointer");
  if (n == 0) return DIGA_OK;
  Lut256 lut;
  for (int i = 0; i < 256; ++i) lut.v[i] = lut_host[i];
  DIGA_REQUIRE(n * H < ((int64_t)1 << 31) && H < (1 << 30) && W < (1 << 30), DIGA_ERR_INVALID, "label_resize_remap: too many rows");
  constexpr int BLOCK = 256;
  int64_t grid = n * H;
  const int64_t cap = (int64_t)sm_count() * (2048 / BLOCK) * 2;          // two waves of resident CTAs
  if (grid > cap) grid = cap;
  label_resize_remap_kernel<BLOCK><<<(unsigned)grid, BLOCK, 0, (cudaStream_t)stream>>>(src, w0, ytab, xtab, (int)H, (int)W, n * H,
                                                                                     h0 * w0, lut, out);
  DIGA_CHECK_LAUNCH("label_resize_remap_kernel");
  return DIGA_OK;
}
