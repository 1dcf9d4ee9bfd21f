// a4 — bilateral-consensus ("threshold-free dynamic") pseudo-label selection.
// Replaces train_DiGA_gta2city_self_training.py:298-304 of the reference: the stride-8 prototype weights
// [B,C,h,w] are bilinearly up-sampled (align_corners=True) to [B,C,H,W], arg-maxed, and the stored
// pseudo-label is kept only where it agrees.  The up-sampled tensor (76 B/px) is never materialised:
// this kernel reads the low-resolution weights through L1/L2 (~1.2 B/px amortised), the int64
// pseudo-labels (8 B/px) and writes the two int64 maps (16 B/px).
#include "bilinear.cuh"
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

// Generic up-sampler (materialising); used by tests to pin the interpolation against torch bit for bit
// and by callers that need the up-sampled map itself.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
upsample_bilinear_kernel(const float* __restrict__ in, int64_t planes, int h, int w, int H, int W, float sh, float sw,
                         float* __restrict__ out) {
  const int64_t total = planes * H * W;
  for (int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x; i < total; i += (int64_t)gridDim.x * BLOCK) {
    const int X = (int)(i % W);
    const int64_t r = i / W;
    const int Y = (int)(r % H);
    const int64_t pl = r / H;
    const Tap ty = bilinear_tap(sh, Y, h), tx = bilinear_tap(sw, X, w);
    const float* src = in + pl * h * w;
    const float top = bilinear_row(tx, __ldg(src + (int64_t)ty.i0 * w + tx.i0), __ldg(src + (int64_t)ty.i0 * w + tx.i1));
    const float bot = bilinear_row(tx, __ldg(src + (int64_t)ty.i1 * w + tx.i0), __ldg(src + (int64_t)ty.i1 * w + tx.i1));
    out[i] = bilinear_col(ty, top, bot);
  }
}

// One thread = PX adjacent output pixels x RY consecutive output rows (ColumnInterp keeps the horizontally
// interpolated source rows of all classes in registers while it walks down).  A warp covers 32*PX adjacent
// pixels of a row, so the int64 accesses are 128-bit and fully coalesced.  torch.max(dim=1): first index on ties.
// Label I/O of PX adjacent pixels: int64 (the training step's LongTensors; 128-bit accesses for PX = 2) or uint8 (the
// offline pseudo-label path, where the maps come from and go to palette PNGs).
template <typename T, int PX>
__device__ __forceinline__ void load_px(const T* __restrict__ p, int64_t (&dst)[PX]) {
  if constexpr (sizeof(T) == 8 && PX == 2) {
    const longlong2 t = ld_stream_i64x2(reinterpret_cast<const int64_t*>(p));
    dst[0] = t.x;
    dst[1] = t.y;
  } else {
#pragma unroll
    for (int v = 0; v < PX; ++v) dst[v] = (int64_t)p[v];
  }
}
template <typename T, int PX>
__device__ __forceinline__ void store_px(T* __restrict__ p, const int64_t (&src)[PX]) {
  if constexpr (sizeof(T) == 8 && PX == 2) {
    st_stream_i64x2(reinterpret_cast<int64_t*>(p), src[0], src[1]);
  } else if constexpr (sizeof(T) == 1 && PX == 2) {
    *reinterpret_cast<uint16_t*>(p) = (uint16_t)((uint8_t)src[0] | ((uint16_t)(uint8_t)src[1] << 8));
  } else {
#pragma unroll
    for (int v = 0; v < PX; ++v) p[v] = (T)src[v];
  }
}

template <int C, bool PAD, int PX, int BLOCK, typename T>
__device__ __forceinline__ void consensus_select_body(const float* __restrict__ wl, const T* __restrict__ pseudo, int nclass, int h, int w,
                                                      int H, int W, float sh, float sw, int RY, T* __restrict__ kept,
                                                      T* __restrict__ feat_pseudo) {
  const int64_t img = blockIdx.z;
  const int Y0 = blockIdx.y * RY;
  const int X0 = (blockIdx.x * BLOCK + threadIdx.x) * PX;
  if (X0 >= W) return;
  const float* base = wl + img * nclass * h * w;
  const int64_t plane = (int64_t)h * w;
  Tap tx[PX];
#pragma unroll
  for (int v = 0; v < PX; ++v) tx[v] = bilinear_tap(sw, X0 + v, w);
  ColumnInterp<C, PAD, PX> ci;
  const int Yend = min(Y0 + RY, H);
  // software prefetch of the label loads: one row ahead
  int64_t lab[PX], nxt[PX];
  auto load_labels = [&](int64_t (&dst)[PX], int Y) { load_px<T, PX>(pseudo + (img * H + Y) * W + X0, dst); };
  load_labels(lab, Y0);
  for (int Y = Y0; Y < Yend; ++Y) {
    if (Y + 1 < Yend) load_labels(nxt, Y + 1);
    const Tap ty = bilinear_tap(sh, Y, h);
    ci.seek(ty, base, plane, w, tx, nclass);
    int64_t am[PX];
#pragma unroll
    for (int v = 0; v < PX; ++v) {
      float val[C];
#pragma unroll
      for (int c = 0; c < C; ++c) val[c] = (!PAD || c < nclass) ? ci.value(ty, v, c) : -INFINITY;
      float best;
      int arg;
      argmax_first<C>(val, best, arg);
      am[v] = arg;
    }
    const int64_t o = (img * H + Y) * W + X0;
    int64_t keep[PX];
#pragma unroll
    for (int v = 0; v < PX; ++v) keep[v] = lab[v] == am[v] ? lab[v] : (int64_t)DIGA_IGNORE_LABEL;                 // :304
    store_px<T, PX>(kept + o, keep);
    if (feat_pseudo) store_px<T, PX>(feat_pseudo + o, am);
#pragma unroll
    for (int v = 0; v < PX; ++v) lab[v] = nxt[v];
  }
}

template <int C, bool PAD, int PX, int BLOCK, typename T>
__global__ void __launch_bounds__(BLOCK)
consensus_select_kernel(const float* __restrict__ wl, const T* __restrict__ pseudo, int nclass, int h, int w, int H, int W, float sh,
                        float sw, int RY, T* __restrict__ kept, T* __restrict__ feat_pseudo) {
  consensus_select_body<C, PAD, PX, BLOCK, T>(wl, pseudo, nclass, h, w, H, W, sh, sw, RY, kept, feat_pseudo);
}
// uint8 labels, two columns per thread (config 5, the PNG path): left alone the compiler takes 136 registers, i.e. three CTAs
// per SM; capped at 128 it fits a fourth without spilling.
template <int C, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 4)
consensus_select_u8x2_kernel(const float* __restrict__ wl, const uint8_t* __restrict__ pseudo, int nclass, int h, int w, int H, int W,
                             float sh, float sw, int RY, uint8_t* __restrict__ kept, uint8_t* __restrict__ feat_pseudo) {
  consensus_select_body<C, false, 2, BLOCK, uint8_t>(wl, pseudo, nclass, h, w, H, W, sh, sw, RY, kept, feat_pseudo);
}

}  // namespace diga

template <typename T>
static int consensus_select_impl(const float* weights_lowres, const T* pseudo, int64_t B, int64_t C, int64_t h, int64_t w, int64_t H,
                                 int64_t W, T* kept, T* feat_pseudo, cudaStream_t st) {
  using namespace diga;
  DIGA_REQUIRE(weights_lowres && pseudo && kept, DIGA_ERR_INVALID, "consensus_select: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "consensus_select: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(B >= 0 && B <= 65535 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && H <= 65535 && h < (1 << 24) && w < (1 << 24) &&
                   W < (1 << 24),
               DIGA_ERR_INVALID, "consensus_select: bad sizes");
  DIGA_REQUIRE(aligned(weights_lowres, 4) && aligned(pseudo, sizeof(T)) && aligned(kept, sizeof(T)) && aligned(feat_pseudo, sizeof(T)),
               DIGA_ERR_MISALIGNED, "consensus_select: misaligned pointer");
  if (B == 0) return DIGA_OK;
  const float sh = bilinear_scale_host(h, H), sw = bilinear_scale_host(w, W);
  const bool pair = (W % 2) == 0 && aligned(pseudo, 2 * sizeof(T)) && aligned(kept, 2 * sizeof(T)) && aligned(feat_pseudo, 2 * sizeof(T));
  // Launch shape (tools/tune.py select, profiles/r01_tune_select.jsonl): a thread walks RY output rows of PX adjacent
  // columns.  Longer walks amortise the two source rows a strip interpolates first but leave too few warps (RY = 64:
  // 48 us, 128: 80 us); two columns x 16 rows is the best of the sweep (35.7 us).  The kernel is ALU-bound: per pixel
  // 19 x (bit-exact 2-op interpolation + 3-op arg-max) plus the strip set-up, 190 instructions at 67 % issue.
  constexpr int BLOCK = 128;
  const int RY = tunable("select_ry", 16) < 1 ? 1 : tunable("select_ry", 16);
  const bool two = pair && tunable("select_px", 2) == 2;
  const unsigned gy = (unsigned)((H + RY - 1) / RY);
  DIGA_DISPATCH_C(C, {
    if (two) {
      dim3 grid((unsigned)((W / 2 + BLOCK - 1) / BLOCK), gy, (unsigned)B);
      if constexpr (sizeof(T) == 1 && !kPad)
        consensus_select_u8x2_kernel<kC, BLOCK><<<grid, BLOCK, 0, st>>>(weights_lowres, pseudo, (int)C, (int)h, (int)w, (int)H, (int)W, sh,
                                                                        sw, RY, kept, feat_pseudo);
      else
        consensus_select_kernel<kC, kPad, 2, BLOCK, T><<<grid, BLOCK, 0, st>>>(weights_lowres, pseudo, (int)C, (int)h, (int)w, (int)H,
                                                                              (int)W, sh, sw, RY, kept, feat_pseudo);
    } else {
      dim3 grid((unsigned)((W + BLOCK - 1) / BLOCK), gy, (unsigned)B);
      consensus_select_kernel<kC, kPad, 1, BLOCK, T><<<grid, BLOCK, 0, st>>>(weights_lowres, pseudo, (int)C, (int)h, (int)w, (int)H,
                                                                            (int)W, sh, sw, RY, kept, feat_pseudo);
    }
  });
  DIGA_CHECK_LAUNCH("consensus_select_kernel");
  return DIGA_OK;
}

extern "C" {

int diga_upsample_bilinear(const float* in, int64_t planes, int64_t h, int64_t w, int64_t H, int64_t W, float* out,
                           diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(in && out, DIGA_ERR_INVALID, "upsample_bilinear: null pointer");
  DIGA_REQUIRE(planes >= 0 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && h < (1 << 24) && w < (1 << 24) && H < (1 << 24) &&
                   W < (1 << 24),
               DIGA_ERR_INVALID, "upsample_bilinear: bad sizes");
  if (planes == 0) return DIGA_OK;
  const int64_t total = planes * H * W;
  int64_t grid = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  upsample_bilinear_kernel<256><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      in, planes, (int)h, (int)w, (int)H, (int)W, bilinear_scale_host(h, H), bilinear_scale_host(w, W), out);
  DIGA_CHECK_LAUNCH("upsample_bilinear_kernel");
  return DIGA_OK;
}

int diga_consensus_select(const float* weights_lowres, const int64_t* pseudo, int64_t B, int64_t C, int64_t h, int64_t w,
                          int64_t H, int64_t W, int64_t* kept, int64_t* feat_pseudo, diga_stream_t stream) {
  return consensus_select_impl<int64_t>(weights_lowres, pseudo, B, C, h, w, H, W, kept, feat_pseudo, (cudaStream_t)stream);
}

int diga_consensus_select_u8(const float* weights_lowres, const uint8_t* pseudo, int64_t B, int64_t C, int64_t h, int64_t w,
                             int64_t H, int64_t W, uint8_t* kept, uint8_t* feat_pseudo, diga_stream_t stream) {
  return consensus_select_impl<uint8_t>(weights_lowres, pseudo, B, C, h, w, H, W, kept, feat_pseudo, (cudaStream_t)stream);
}

}  // extern "C"
