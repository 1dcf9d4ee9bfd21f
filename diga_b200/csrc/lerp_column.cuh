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

// LerpColumn on class pairs (NP pairs = 2 NP class slots per thread).  A slot without a class (odd C, or c >= nclass in
// the padded variant) is fed kLerpPad from both source rows: its value stays hugely negative (dif == 0 exactly, top
// absorbs every re-basing), so its exponential is an exact zero and it never wins a maximum — no special case in the
// per-pixel loops.
constexpr float kLerpPad = -1e30f;

template <int NP>
struct LerpColumnP {
  float2 top[NP], dif[NP];
  float ref2 = 0.f;

  // dst[p] = l0s * v[c][k] + l1s * v[c][k + 1] - sub.
  // TILE: `q` points at the class vector of source column k in the CTA's shared-memory tile [row][column][2 NP] (absent
  // classes pre-filled with kLerpPad); the next column is `col_stride` floats on: two 64-bit loads per class pair.
  __device__ __forceinline__ void hrow_tile(float2 (&dst)[NP], const float* q, int col_stride, float l0s, float l1s, float sub) {
    const float2 w0 = splat2(l0s), w1 = splat2(l1s), ns = splat2(-sub);
    const float2* qa = reinterpret_cast<const float2*>(q);
    const float2* qb = reinterpret_cast<const float2*>(q + col_stride);
#pragma unroll
    for (int p = 0; p < NP; ++p) dst[p] = ffma2(w0, qa[p], ffma2(w1, qb[p], ns));
  }
  // global memory: `q` points at v[0][k] of the source row, classes `class_stride` apart (an absent class re-reads the last one)
  __device__ __forceinline__ void hrow_global(float2 (&dst)[NP], const float* __restrict__ q, int64_t class_stride, int second, float l0s,
                                              float l1s, float sub, int nlimit) {
    const float2 w0 = splat2(l0s), w1 = splat2(l1s), ns = splat2(-sub);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int c0 = 2 * p, c1 = c0 + 1;
      const float* q0 = q + (int64_t)min(c0, nlimit - 1) * class_stride;
      const float* q1 = q + (int64_t)min(c1, nlimit - 1) * class_stride;
      float2 va = make_float2(__ldg(q0), __ldg(q1)), vb = make_float2(__ldg(q0 + second), __ldg(q1 + second));
      if (c0 >= nlimit) va.x = vb.x = kLerpPad;
      if (c1 >= nlimit) va.y = vb.y = kLerpPad;
      dst[p] = ffma2(w0, va, ffma2(w1, vb, ns));
    }
  }
  __device__ __forceinline__ float rowmax(const float2 (&v)[NP]) {
    float m0 = v[0].x, m1 = v[0].y, m2 = v[0].x;
#pragma unroll
    for (int p = 1; p < NP; ++p) {
      float& m = (p % 3 == 0) ? m0 : (p % 3 == 1) ? m1 : m2;
      m = fmax3(m, v[p].x, v[p].y);
    }
    return fmax3(m0, m1, m2);
  }
  // rows r0 (-> top) and r1 (-> bottom) of the cell; `fresh` = first cell of the strip, otherwise the old bottom row
  // (top + dif) becomes the new top row.  `load(dst, row, sub)` fetches one horizontally interpolated source row.
  template <typename Load>
  __device__ __forceinline__ void enter(bool fresh, int r0, int r1, Load&& load) {
    if (fresh) {
      load(top, r0, 0.f);
      ref2 = 0.f;
    } else {
#pragma unroll
      for (int p = 0; p < NP; ++p) top[p] = fadd2(top[p], dif[p]);
    }
    float m = rowmax(top);
    if (r1 != r0) {
      load(dif, r1, ref2);
      m = fmaxf(m, rowmax(dif));
      const float2 neg1 = splat2(-1.f), nm = splat2(-m);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        dif[p] = ffma2(top[p], neg1, dif[p]);          // dif - top, one rounding
        top[p] = fadd2(top[p], nm);
      }
    } else {
      const float2 nm = splat2(-m);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        dif[p] = make_float2(0.f, 0.f);
        top[p] = fadd2(top[p], nm);
      }
    }
    ref2 += m;
  }
  __device__ __forceinline__ float2 value2(float2 l1, int p) const { return ffma2(l1, dif[p], top[p]); }
};

}  // namespace diga
