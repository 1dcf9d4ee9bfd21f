// a3 — pseudo-label generation: (two-scale max) -> softmax -> argmax + confidence, one pass.
// Replaces pseudolabel_generator.py:80-85 of the reference; the *_upsampled entry point also folds in
// the two bilinear up-samplings of :77-78.
//
// Layout: logits [n, C, hw] fp32 NCHW.  A thread owns VEC adjacent pixels, loads one VEC-wide vector
// per class plane (and per scale), keeps max / argmax / sum-exp in registers, and writes 1 B (u8 label,
// VEC labels packed into one store) + 4 B (confidence) per pixel.  Algorithmic traffic: 4C + 5 B/px
// for one scale, 8C + 5 B/px for two.
//
// Semantics: label = first index of the maximum *logit* (== numpy argmax of the softmax except where
// two softmax values round to the same float); confidence = 1 / sum_c exp(z_c - max).
// NaN logits are not supported (the reference would propagate them).
#include "bilinear.cuh"
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

template <int VEC>
__device__ __forceinline__ void store_labels_u8(uint8_t* p, const int (&lab)[VEC]) {
  if constexpr (VEC == 4) {
    const uint32_t packed = (uint32_t)lab[0] | ((uint32_t)lab[1] << 8) | ((uint32_t)lab[2] << 16) | ((uint32_t)lab[3] << 24);
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(packed) : "memory");
  } else if constexpr (VEC == 2) {
    const uint16_t packed = (uint16_t)(lab[0] | (lab[1] << 8));
    asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"(packed) : "memory");
  } else {
    p[0] = (uint8_t)lab[0];
  }
}

template <int VEC>
__device__ __forceinline__ void store_labels_i64(int64_t* p, const int (&lab)[VEC]) {
  if constexpr (VEC >= 2) {
#pragma unroll
    for (int v = 0; v < VEC; v += 2) st_stream_i64x2(p + v, (int64_t)lab[v], (int64_t)lab[v + 1]);
  } else {
    st_stream_i64(p, (int64_t)lab[0]);
  }
}

// Register budget: C*VEC floats per scale.  One scale: VEC=4 fits 128 registers (2 CTAs/SM); two scales
// use VEC=2 for the same budget.
template <int C, bool PAD, int VEC, int BLOCK, bool HAS2>
__global__ void __launch_bounds__(BLOCK, 2)
pseudo_label_kernel(const float* __restrict__ z1, const float* __restrict__ z2, int nclass, int64_t n, int64_t hw,
                    uint8_t* __restrict__ lab8, int64_t* __restrict__ lab64, float* __restrict__ conf) {
  const int64_t groups_per_img = hw / VEC;
  const int64_t total = n * groups_per_img;
  for (int64_t gidx = (int64_t)blockIdx.x * BLOCK + threadIdx.x; gidx < total; gidx += (int64_t)gridDim.x * BLOCK) {
    const int64_t img = gidx / groups_per_img;
    const int64_t p = (gidx - img * groups_per_img) * VEC;
    const float* p1 = z1 + img * nclass * hw + p;
    Vec<VEC> z[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) z[c] = ld_stream<VEC>(p1 + c * hw);
    if constexpr (HAS2) {
      const float* p2 = z2 + img * nclass * hw + p;
      Vec<VEC> y[C];
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) y[c] = ld_stream<VEC>(p2 + c * hw);
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) z[c].v[v] = fmaxf(z[c].v[v], y[c].v[v]);   // torch.max(output_ds, output)
        }
    }
    int lab[VEC];
    Vec<VEC> cf;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float m = z[0].v[v];
      int am = 0;
#pragma unroll
      for (int c = 1; c < C; ++c)
        if (!PAD || c < nclass) {
          const bool gt = z[c].v[v] > m;   // strict: first index wins ties
          m = gt ? z[c].v[v] : m;
          am = gt ? c : am;
        }
      float S = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) S += fast_exp(z[c].v[v] - m);
      lab[v] = am;
      cf.v[v] = 1.0f / S;
    }
    const int64_t o = img * hw + p;
    if (lab8) store_labels_u8<VEC>(lab8 + o, lab);
    if (lab64) store_labels_i64<VEC>(lab64 + o, lab);
    if (conf) st_stream<VEC>(conf + o, cf);
  }
}

template <int C, bool PAD, int VEC, int BLOCK, bool HAS2>
static int launch_pl(const float* z1, const float* z2, int nclass, int64_t n, int64_t hw, uint8_t* lab8, int64_t* lab64,
                     float* conf, cudaStream_t st) {
  auto kern = pseudo_label_kernel<C, PAD, VEC, BLOCK, HAS2>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BLOCK, 0);
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int64_t total = n * (hw / VEC);
  int64_t grid = (total + BLOCK - 1) / BLOCK;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm * (HAS2 ? tunable("pl_waves2", 8) : tunable("pl_waves", 2));
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, BLOCK, 0, st>>>(z1, z2, nclass, n, hw, lab8, lab64, conf);
  DIGA_CHECK_LAUNCH("pseudo_label_kernel");
  return DIGA_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused two-scale variant: the bilinear up-samplings of pseudolabel_generator.py:77-78 are evaluated
// on the fly from the stride-8 logits (L1/L2 resident), so HBM traffic drops from 8C+5 to ~5 B/px.
// Interpolation mirrors ATen (bilinear.cuh).
// ------------------------------------------------------------------------------------------------
// One thread = one output column x RY consecutive rows; both scales are walked with a ColumnInterp each.
// CONF=false drops the 19 exponentials per pixel: the reference computes the confidence and throws it away
// (`label, _ = np.argmax(...), np.max(...)`, pseudolabel_generator.py:85), and this kernel is instruction-bound.
template <int C, bool PAD, int RY, int BLOCK, bool HAS2, bool CONF>
__global__ void __launch_bounds__(BLOCK)
pseudo_label_upsampled_kernel(const float* __restrict__ z1, int h1, int w1, float sh1, float sw1,
                              const float* __restrict__ z2, int h2, int w2, float sh2, float sw2, int nclass, int H, int W,
                              uint8_t* __restrict__ lab8, int64_t* __restrict__ lab64, float* __restrict__ conf) {
  const int64_t img = blockIdx.z;
  const int Y0 = blockIdx.y * RY;
  const int X = blockIdx.x * BLOCK + threadIdx.x;
  if (X >= W) return;
  const float* b1 = z1 + img * nclass * h1 * w1;
  const int64_t pl1 = (int64_t)h1 * w1;
  const Tap tx1[1] = {bilinear_tap(sw1, X, w1)};
  ColumnInterp<C, PAD, 1> c1;
  const float* b2 = nullptr;
  int64_t pl2 = 0;
  Tap tx2[1] = {tx1[0]};
  ColumnInterp<HAS2 ? C : 1, PAD, 1> c2;
  if constexpr (HAS2) {
    b2 = z2 + img * nclass * h2 * w2;
    pl2 = (int64_t)h2 * w2;
    tx2[0] = bilinear_tap(sw2, X, w2);
  }
  const int Yend = min(Y0 + RY, H);
  for (int Y = Y0; Y < Yend; ++Y) {
    const Tap ty1 = bilinear_tap(sh1, Y, h1);
    c1.seek(ty1, b1, pl1, w1, tx1, nclass);
    Tap ty2 = ty1;
    if constexpr (HAS2) {
      ty2 = bilinear_tap(sh2, Y, h2);
      c2.seek(ty2, b2, pl2, w2, tx2, nclass);
    }
    float z[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float val = -INFINITY;
      if (!PAD || c < nclass) {
        val = c1.value(ty1, 0, c);
        if constexpr (HAS2) val = fmaxf(val, c2.value(ty2, 0, c));   // torch.max(output_ds, output), :80
      }
      z[c] = val;
    }
    float m;
    int am;
    argmax_first<C>(z, m, am);
    const int64_t o = (img * H + Y) * W + X;
    if (lab8) lab8[o] = (uint8_t)am;
    if (lab64) lab64[o] = am;
    if constexpr (CONF) {
      float S = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) S += fast_exp(z[c] - m);
      conf[o] = 1.0f / S;
    }
  }
}

}  // namespace diga

extern "C" int diga_pseudo_label_upsampled(const float* logits, int64_t h1, int64_t w1, const float* logits_ds, int64_t h2,
                                           int64_t w2, int64_t n, int64_t C, int64_t H, int64_t W, uint8_t* label_u8,
                                           int64_t* label_i64, float* conf, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(logits, DIGA_ERR_INVALID, "pseudo_label_upsampled: null logits");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "pseudo_label_upsampled: C=%lld outside [1,%d]",
               (long long)C, DIGA_MAX_CLASSES);
  const int64_t lim = 1 << 24;
  DIGA_REQUIRE(n >= 0 && n <= 65535 && h1 >= 1 && w1 >= 1 && H >= 1 && W >= 1 && H <= 65535 && h1 < lim && w1 < lim && W < lim,
               DIGA_ERR_INVALID, "pseudo_label_upsampled: bad sizes");
  DIGA_REQUIRE(!logits_ds || (h2 >= 1 && w2 >= 1 && h2 < lim && w2 < lim), DIGA_ERR_INVALID,
               "pseudo_label_upsampled: bad low-res sizes");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(logits_ds, 4) && aligned(conf, 4) && aligned(label_i64, 8), DIGA_ERR_MISALIGNED,
               "pseudo_label_upsampled: misaligned pointer");
  if (n == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int BLOCK = 128, RY = 16;
  dim3 grid((unsigned)((W + BLOCK - 1) / BLOCK), (unsigned)((H + RY - 1) / RY), (unsigned)n);
  const float sh1 = bilinear_scale_host(h1, H), sw1 = bilinear_scale_host(w1, W);
  const float sh2 = logits_ds ? bilinear_scale_host(h2, H) : 0.f, sw2 = logits_ds ? bilinear_scale_host(w2, W) : 0.f;
#define DIGA_PLU_GO(HAS2, CONF)                                                                                      \
  pseudo_label_upsampled_kernel<kC, kPad, RY, BLOCK, HAS2, CONF><<<grid, BLOCK, 0, st>>>(                              \
      logits, (int)h1, (int)w1, sh1, sw1, logits_ds, (int)h2, (int)w2, sh2, sw2, (int)C, (int)H, (int)W, label_u8, label_i64, conf)
  DIGA_DISPATCH_C(C, {
    if (logits_ds) {
      if (conf) DIGA_PLU_GO(true, true); else DIGA_PLU_GO(true, false);
    } else {
      if (conf) DIGA_PLU_GO(false, true); else DIGA_PLU_GO(false, false);
    }
  });
#undef DIGA_PLU_GO
  DIGA_CHECK_LAUNCH("pseudo_label_upsampled_kernel");
  return DIGA_OK;
}

namespace diga {
}  // namespace diga

extern "C" int diga_pseudo_label(const float* logits, const float* logits_ds, int64_t n, int64_t C, int64_t hw,
                                 uint8_t* label_u8, int64_t* label_i64, float* conf, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(n >= 0 && hw >= 0, DIGA_ERR_INVALID, "pseudo_label: negative size");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "pseudo_label: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  if (n == 0 || hw == 0) return DIGA_OK;      // empty batch: nothing to read, pointers may be NULL
  DIGA_REQUIRE(logits, DIGA_ERR_INVALID, "pseudo_label: null logits");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(logits_ds, 4) && aligned(conf, 4) && aligned(label_i64, 8),
               DIGA_ERR_MISALIGNED, "pseudo_label: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool has2 = logits_ds != nullptr;
  int vec = has2 ? tunable("pl_vec2", 2) : tunable("pl_vec", 4);
  const bool a16 = aligned(logits, 16) && aligned(logits_ds, 16) && aligned(conf, 16) && aligned(label_i64, 16) &&
                   aligned(label_u8, 4);
  const bool a8 = aligned(logits, 8) && aligned(logits_ds, 8) && aligned(conf, 8) && aligned(label_i64, 16) &&
                  aligned(label_u8, 2);
  if (vec == 4 && !((hw % 4) == 0 && a16)) vec = 2;
  if (vec == 2 && !((hw % 2) == 0 && a8)) vec = 1;
#define DIGA_PL_GO(V)                                                                                         \
  do {                                                                                                        \
    if (has2) return launch_pl<kC, kPad, V, 256, true>(logits, logits_ds, (int)C, n, hw, label_u8, label_i64, conf, st); \
    return launch_pl<kC, kPad, V, 256, false>(logits, nullptr, (int)C, n, hw, label_u8, label_i64, conf, st);  \
  } while (0)
  DIGA_DISPATCH_C(C, {
    if (vec == 4) DIGA_PL_GO(4);
    if (vec == 2) DIGA_PL_GO(2);
    DIGA_PL_GO(1);
  });
#undef DIGA_PL_GO
  return DIGA_OK;
}
