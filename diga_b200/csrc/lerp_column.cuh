// Shared pieces of the fused up-sampling loss kernels (csrc/loss_up.cu, csrc/ohem_up.cu): the vertical walk of one output
// column in the cell-referenced log2 domain.
#pragma once

#include "bilinear.cuh"
#include "common.cuh"

namespace diga {

// Three-input max (sm_100: one FMNMX3 instead of two FMNMX).
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Vertical walk of one column in the log2 domain.  For the current source cell the thread keeps, per class,
//     top[c] = (row_i0[c] - ref) * log2(e)     dif[c] = (row_i1[c] - row_i0[c]) * log2(e)
// where row_* are the horizontally interpolated source rows and `ref` is the largest value either row holds in any
// class, so that every interpolated value  v_c = top[c] + l1 * dif[c]  (ONE FFMA per class and output row) is <= 0 and
// can go straight into ex2: no per-pixel max, no per-pixel rescaling.  softmax / log-sum-exp / soft-target cross
// entropy are invariant to the choice of ref; `ref` only has to keep the sums away from underflow, which the caller
// checks per pixel (it falls back to the exact per-pixel max when a sum drops below 2^-60).
// The loss path is held to 1e-5, not to the bit pattern of ATen's up-sampler (the label paths keep ColumnInterp).
template <int C, bool PAD>
struct LerpColumn {
  float top[C], dif[C];
  float ref2 = 0.f;       // reference, log2 units

  // dst[c] = l0s * v[c][k] + l1s * v[c][k + 1] - sub.  `q` points at v[0][k]; one 64-bit pointer bump per class, the
  // second load is the same register with an immediate offset (PAIR = false: single-column source, w == 1).
  template <bool PAIR>
  __device__ __forceinline__ void hrow(float (&dst)[C], const float* __restrict__ q, int64_t class_stride, float l0s, float l1s,
                                       float sub, int nclass) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        dst[c] = fmaf(l0s, __ldg(q), fmaf(l1s, __ldg(q + (PAIR ? 1 : 0)), -sub));
        q += class_stride;
      }
  }
  __device__ __forceinline__ float rowmax(const float (&v)[C], int nclass) {
    float m0 = v[0], m1 = v[0], m2 = v[0];
#pragma unroll
    for (int c = 1; c + 1 < C; c += 2) {
      float& m = ((c >> 1) % 3 == 0) ? m0 : ((c >> 1) % 3 == 1) ? m1 : m2;
      if (!PAD || c + 1 < nclass) m = fmax3(m, v[c], v[c + 1]);
      else if (c < nclass) m = fmaxf(m, v[c]);
    }
    if constexpr ((C & 1) == 0) {
      if (!PAD || C - 1 < nclass) m0 = fmaxf(m0, v[C - 1]);
    }
    return fmax3(m0, m1, m2);
  }
  // rows r0 (-> top) and r1 (-> bottom) of the cell; `fresh` = first cell of the strip, otherwise the old bottom row
  // (top + dif) becomes the new top row.
  __device__ __forceinline__ void enter(bool fresh, const float* __restrict__ base, int64_t row_stride, int64_t class_stride, int r0,
                                        int r1, bool pair, float l0s, float l1s, int nclass) {
    if (fresh) {
      if (pair) hrow<true>(top, base + r0 * row_stride, class_stride, l0s, l1s, 0.f, nclass);
      else hrow<false>(top, base + r0 * row_stride, class_stride, l0s, l1s, 0.f, nclass);
      ref2 = 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) top[c] += dif[c];
    }
    float m = rowmax(top, nclass);
    if (r1 != r0) {
      if (pair) hrow<true>(dif, base + r1 * row_stride, class_stride, l0s, l1s, ref2, nclass);
      else hrow<false>(dif, base + r1 * row_stride, class_stride, l0s, l1s, ref2, nclass);
      m = fmaxf(m, rowmax(dif, nclass));
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dif[c] -= top[c];
        top[c] -= m;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dif[c] = 0.f;
        top[c] -= m;
      }
    }
    ref2 += m;
  }
  __device__ __forceinline__ float value(float l1, int c) const { return fmaf(l1, dif[c], top[c]); }
};

// ---- packed single precision (sm_100: FFMA2 / FADD2 / FMUL2 — two IEEE fp32 operations per issue slot) ---------------
// The loss kernels are bound by issue slots, not by the FMA pipe: the per-class arithmetic runs on class PAIRS.
#define DIGA_F32X2_3(name, op)                                                                                          \
  __device__ __forceinline__ float2 name(float2 a, float2 b, float2 c) {                                               \
    float2 d;                                                                                                          \
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t" op \
        " rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"                                                              \
        : "=f"(d.x), "=f"(d.y)                                                                                         \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));                                                 \
    return d;                                                                                                          \
  }
#define DIGA_F32X2_2(name, op)                                                                                          \
  __device__ __forceinline__ float2 name(float2 a, float2 b) {                                                         \
    float2 d;                                                                                                          \
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t" op                          \
        " rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"                                                                  \
        : "=f"(d.x), "=f"(d.y)                                                                                         \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                                     \
    return d;                                                                                                          \
  }
DIGA_F32X2_3(ffma2, "fma.rn.f32x2")
DIGA_F32X2_2(fadd2, "add.rn.f32x2")
DIGA_F32X2_2(fmul2, "mul.rn.f32x2")
#undef DIGA_F32X2_3
#undef DIGA_F32X2_2
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// LerpColumn on class pairs: pair p holds classes (2p, 2p + 1).  A lane without a class (odd C, or c >= nclass in the
// padded variant) carries finite don't-care values; the callers force its exponential to zero and skip it in maxima.
template <int C, bool PAD>
struct LerpColumn2 {
  static constexpr int P = (C + 1) / 2;
  float2 top[P], dif[P];
  float ref2 = 0.f;

  static __device__ __forceinline__ bool on(int c, int nclass) { return c < C && (!PAD || c < nclass); }

  template <bool PAIR>
  __device__ __forceinline__ void hrow(float2 (&dst)[P], const float* q, int64_t class_stride, float l0s, float l1s,
                                       float sub, int nclass) {
    const float2 w0 = splat2(l0s), w1 = splat2(l1s), ns = splat2(-sub);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float2 va = make_float2(0.f, 0.f), vb = make_float2(0.f, 0.f);
      if (on(2 * p, nclass)) {                      // plain loads: `q` is the CTA's shared-memory tile or global memory
        va.x = q[0];
        vb.x = q[PAIR ? 1 : 0];
        q += class_stride;
      }
      if (on(2 * p + 1, nclass)) {
        va.y = q[0];
        vb.y = q[PAIR ? 1 : 0];
        q += class_stride;
      }
      dst[p] = ffma2(w0, va, ffma2(w1, vb, ns));
    }
  }
  __device__ __forceinline__ float rowmax(const float2 (&v)[P], int nclass) {
    float m0 = v[0].x, m1 = v[0].x, m2 = v[0].x;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float& m = (p % 3 == 0) ? m0 : (p % 3 == 1) ? m1 : m2;
      if (on(2 * p + 1, nclass)) m = fmax3(m, v[p].x, v[p].y);
      else if (on(2 * p, nclass)) m = fmaxf(m, v[p].x);
    }
    return fmax3(m0, m1, m2);
  }
  __device__ __forceinline__ void enter(bool fresh, const float* base, int64_t row_stride, int64_t class_stride, int r0,
                                        int r1, bool pair, float l0s, float l1s, int nclass) {
    if (fresh) {
      if (pair) hrow<true>(top, base + r0 * row_stride, class_stride, l0s, l1s, 0.f, nclass);
      else hrow<false>(top, base + r0 * row_stride, class_stride, l0s, l1s, 0.f, nclass);
      ref2 = 0.f;
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) top[p] = fadd2(top[p], dif[p]);
    }
    float m = rowmax(top, nclass);
    if (r1 != r0) {
      if (pair) hrow<true>(dif, base + r1 * row_stride, class_stride, l0s, l1s, ref2, nclass);
      else hrow<false>(dif, base + r1 * row_stride, class_stride, l0s, l1s, ref2, nclass);
      m = fmaxf(m, rowmax(dif, nclass));
      const float2 neg1 = splat2(-1.f), nm = splat2(-m);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        dif[p] = ffma2(top[p], neg1, dif[p]);          // dif - top, one rounding
        top[p] = fadd2(top[p], nm);
      }
    } else {
      const float2 nm = splat2(-m);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        dif[p] = make_float2(0.f, 0.f);
        top[p] = fadd2(top[p], nm);
      }
    }
    ref2 += m;
  }
  __device__ __forceinline__ float2 value2(float2 l1, int p) const { return ffma2(l1, dif[p], top[p]); }
};

}  // namespace diga
