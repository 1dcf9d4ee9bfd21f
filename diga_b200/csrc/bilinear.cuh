// Bilinear interpolation with align_corners=True, mirroring ATen's CUDA kernel bit for bit.
//
// The reference takes argmax over bilinearly up-sampled maps (self_training.py:302-303,
// pseudolabel_generator.py:77-78), so the interpolated VALUES sit inside a bit-exact label path and the
// expression tree must equal the one torch runs on the same GPU.  It was read off the SASS of
// upsample_bilinear2d_out_frame<float,float> in the libtorch_cuda.so of this image (torch 2.11+cu128,
// sm_100 cubin; see DESIGN.md §"Bilinear parity"):
//     scale   = (float)(in - 1) / (float)(out - 1)          (host, fp32 division; 0 when out == 1)
//     src     = scale * (float)dst                           FMUL
//     i0      = (int)src ; i1 = i0 + (i0 < in - 1)           F2I.TRUNC
//     l1      = src - (float)i0 ; l0 = 1 - l1                FADD, FADD
//     top     = fma(w0, v00, w1 * v01)                       FMUL, FFMA
//     bottom  = fma(w0, v10, w1 * v11)                       FMUL, FFMA
//     val     = fma(h0, top, h1 * bottom)                    FMUL, FFMA
#pragma once

#include "common.cuh"

namespace diga {

struct Tap {
  int i0, i1;
  float l0, l1;
};

static inline float bilinear_scale_host(long long in, long long out) {
  return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
}

__device__ __forceinline__ Tap bilinear_tap(float scale, int dst, int in) {
  Tap t;
  const float src = __fmul_rn(scale, (float)dst);
  t.i0 = (int)src;
  t.i1 = t.i0 + ((t.i0 < in - 1) ? 1 : 0);
  t.l1 = __fsub_rn(src, (float)t.i0);
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}

__device__ __forceinline__ float bilinear_row(const Tap& tx, float v0, float v1) {
  return __fmaf_rn(tx.l0, v0, __fmul_rn(tx.l1, v1));
}

__device__ __forceinline__ float bilinear_col(const Tap& ty, float top, float bottom) {
  return __fmaf_rn(ty.l0, top, __fmul_rn(ty.l1, bottom));
}

// Vertical-run evaluator.  A thread that walks down a column of output pixels keeps, for every class, the two
// horizontally interpolated source rows (`top`, `bot`) of the current source cell in registers.  Moving to the
// next output row inside the same cell costs one FMUL + FFMA per class; crossing into the next cell re-uses `bot`
// as the new `top` (same inputs, same expression => bit-identical) and interpolates one new source row.  The
// values are exactly those of ATen's per-pixel expression (common sub-expressions are merely not recomputed).
// The classes are held as PAIRS and interpolated with the packed FMUL2 / FFMA2 of sm_100 (common.cuh): each lane is the
// same IEEE multiply / fused multiply-add as the scalar instruction, so the bit pattern is unchanged while the two
// instructions per class and row become one.
template <int C, bool PAD, int PX>
struct ColumnInterp {
  static constexpr int P = (C + 1) / 2;
  float2 top[PX][P], bot[PX][P];
  int i0 = -1, i1 = -1;

  // One pointer pair per pixel, bumped by `plane` per class: two 64-bit adds per class and pixel instead of a fresh
  // index -> address computation per load (the setup code was a third of the selection kernel's instructions).
  __device__ __forceinline__ void row(float2 (&dst)[PX][P], const float* __restrict__ base, int64_t plane, int w, int r,
                                      const Tap (&tx)[PX], int nclass) {
    const float* q0[PX];
    const float* q1[PX];
#pragma unroll
    for (int v = 0; v < PX; ++v) {
      q0[v] = base + (int64_t)r * w + tx[v].i0;
      q1[v] = base + (int64_t)r * w + tx[v].i1;
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
      for (int v = 0; v < PX; ++v) {
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
        if (!PAD || 2 * p < nclass) {
          a.x = __ldg(q0[v]);
          b.x = __ldg(q1[v]);
          q0[v] += plane;
          q1[v] += plane;
        }
        if (2 * p + 1 < C && (!PAD || 2 * p + 1 < nclass)) {
          a.y = __ldg(q0[v]);
          b.y = __ldg(q1[v]);
          q0[v] += plane;
          q1[v] += plane;
        }
        dst[v][p] = ffma2(splat2(tx[v].l0), a, fmul2(splat2(tx[v].l1), b));      // bilinear_row on both lanes
      }
    }
  }

  __device__ __forceinline__ void seek(const Tap& ty, const float* __restrict__ base, int64_t plane, int w, const Tap (&tx)[PX],
                                       int nclass) {
    if (ty.i0 == i0 && ty.i1 == i1) return;
    if (ty.i0 == i1 && i1 >= 0) {
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int v = 0; v < PX; ++v) top[v][p] = bot[v][p];
    } else if (ty.i0 != i0) {
      row(top, base, plane, w, ty.i0, tx, nclass);
    }
    if (ty.i1 == ty.i0) {
#pragma unroll
      for (int p = 0; p < P; ++p)
#pragma unroll
        for (int v = 0; v < PX; ++v) bot[v][p] = top[v][p];
    } else {
      row(bot, base, plane, w, ty.i1, tx, nclass);
    }
    i0 = ty.i0;
    i1 = ty.i1;
  }

  // the interpolated values of classes (2p, 2p + 1) of pixel v: bilinear_col on both lanes
  __device__ __forceinline__ float2 value2(const Tap& ty, int v, int p) const {
    return ffma2(splat2(ty.l0), top[v][p], fmul2(splat2(ty.l1), bot[v][p]));
  }
  // all classes of pixel v into z[0..C) (-inf for c >= nclass in the padded variant)
  __device__ __forceinline__ void values(const Tap& ty, int v, int nclass, float (&z)[C]) const {
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float2 val = value2(ty, v, p);
      z[2 * p] = (!PAD || 2 * p < nclass) ? val.x : -INFINITY;
      if (2 * p + 1 < C) z[2 * p + 1] = (!PAD || 2 * p + 1 < nclass) ? val.y : -INFINITY;
    }
  }
};

}  // namespace diga
