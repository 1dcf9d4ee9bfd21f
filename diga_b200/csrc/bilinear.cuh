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
template <int C, bool PAD, int PX>
struct ColumnInterp {
  float top[PX][C], bot[PX][C];
  int i0 = -1, i1 = -1;

  // One pointer pair per pixel, bumped by `plane` per class: two 64-bit adds per class and pixel instead of a fresh
  // index -> address computation per load (the setup code was a third of the selection kernel's instructions).
  __device__ __forceinline__ void row(float (&dst)[PX][C], const float* __restrict__ base, int64_t plane, int w, int r,
                                      const Tap (&tx)[PX], int nclass) {
    const float* q0[PX];
    const float* q1[PX];
#pragma unroll
    for (int v = 0; v < PX; ++v) {
      q0[v] = base + (int64_t)r * w + tx[v].i0;
      q1[v] = base + (int64_t)r * w + tx[v].i1;
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
#pragma unroll
        for (int v = 0; v < PX; ++v) {
          dst[v][c] = bilinear_row(tx[v], __ldg(q0[v]), __ldg(q1[v]));
          q0[v] += plane;
          q1[v] += plane;
        }
      }
  }

  __device__ __forceinline__ void seek(const Tap& ty, const float* __restrict__ base, int64_t plane, int w, const Tap (&tx)[PX],
                                       int nclass) {
    if (ty.i0 == i0 && ty.i1 == i1) return;
    if (ty.i0 == i1 && i1 >= 0) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int v = 0; v < PX; ++v) top[v][c] = bot[v][c];
    } else if (ty.i0 != i0) {
      row(top, base, plane, w, ty.i0, tx, nclass);
    }
    if (ty.i1 == ty.i0) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int v = 0; v < PX; ++v) bot[v][c] = top[v][c];
    } else {
      row(bot, base, plane, w, ty.i1, tx, nclass);
    }
    i0 = ty.i0;
    i1 = ty.i1;
  }

  __device__ __forceinline__ float value(const Tap& ty, int v, int c) const { return bilinear_col(ty, top[v][c], bot[v][c]); }
};

}  // namespace diga
