// Shared device/host helpers for the diga_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/diga_b200.h"

namespace diga {

// ---------------------------------------------------------------------------------------------
// host side: error reporting, launch accounting, device properties
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();
int device_slot();   // current device, clamped to [0, 64): key of per-device launch-configuration caches

#define DIGA_REQUIRE(cond, code, ...)     \
  do {                                    \
    if (!(cond)) {                        \
      ::diga::set_error(__VA_ARGS__);     \
      return (code);                      \
    }                                     \
  } while (0)

// Call after a kernel launch.  cudaPeekAtLastError does not synchronise.
#define DIGA_CHECK_LAUNCH(name)                                                        \
  do {                                                                                 \
    cudaError_t e__ = cudaPeekAtLastError();                                           \
    if (e__ != cudaSuccess) {                                                          \
      (void)cudaGetLastError();                                                        \
      ::diga::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));     \
      return DIGA_ERR_CUDA;                                                            \
    }                                                                                  \
    ::diga::count_launch();                                                            \
  } while (0)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// Dispatch on the class count: the two the reference uses (19: GTA5/DG/SS trees, 16: Synthia) are
// fully unrolled; any other C <= 32 runs a padded generic instantiation.
#define DIGA_DISPATCH_C(C, ...)                          \
  do {                                                   \
    if ((C) == 19) {                                     \
      constexpr int kC = 19; constexpr bool kPad = false; \
      __VA_ARGS__                                        \
    } else if ((C) == 16) {                              \
      constexpr int kC = 16; constexpr bool kPad = false; \
      __VA_ARGS__                                        \
    } else {                                             \
      constexpr int kC = 32; constexpr bool kPad = true;  \
      __VA_ARGS__                                        \
    }                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Streaming global accesses: every byte on this path is touched exactly once, so loads bypass L1
// allocation and stores are marked evict-first (cache-streaming).
template <int VEC> struct Vec;
template <> struct Vec<1> { float v[1]; };
template <> struct alignas(8) Vec<2> { float v[2]; };
template <> struct alignas(16) Vec<4> { float v[4]; };

template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_stream(const float* p) {
  Vec<VEC> r;
  if constexpr (VEC == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
  } else if constexpr (VEC == 2) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void st_stream(float* p, const Vec<VEC>& r) {
  if constexpr (VEC == 4) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                 "f"(r.v[3]) : "memory");
  } else if constexpr (VEC == 2) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]) : "memory");
  } else {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(r.v[0]) : "memory");
  }
}

__device__ __forceinline__ longlong2 ld_stream_i64x2(const int64_t* p) {
  longlong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int64_t ld_stream_i64(const int64_t* p) {
  int64_t r;
  asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_i64x2(int64_t* p, int64_t a, int64_t b) {
  asm volatile("st.global.cs.v2.s64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_stream_i64(int64_t* p, int64_t a) {
  asm volatile("st.global.cs.s64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}

// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (rel. error 2^-22); exp2(0) == 1 exactly.
__device__ __forceinline__ float fast_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * kLog2e));
  return r;
}
__device__ __forceinline__ float fast_log(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * kLn2;
}

// The raw MUFU forms (log2 domain), for kernels that fold the log2(e) factor into an FFMA.
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {   // MUFU.RCP, 1 ulp
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ---- packed single precision (sm_100: FFMA2 / FADD2 / FMUL2 — two IEEE fp32 operations per issue slot) ---------------
// Every lane is one IEEE operation: results are bit-identical to the scalar instructions (the label kernels rely on it).
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

// First index of the maximum of v[0..C) (torch.max / np.argmax tie rule) as a pairwise tournament: in every comparison
// the left operand covers the lower indices and the right one wins only if STRICTLY greater, so ties resolve to the lowest
// index exactly like the left-to-right scan, but the dependency depth is log2(C) instead of C (the scan left the
// ALU-bound label kernels waiting on their own compare -> select chain).  Entries c >= nclass must hold -inf.
template <int C>
__device__ __forceinline__ void argmax_first(const float (&v)[C], float& best, int& arg) {
  float bv[C];
  int bi[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    bv[c] = v[c];
    bi[c] = c;
  }
#pragma unroll
  for (int stride = 1; stride < C; stride *= 2) {
#pragma unroll
    for (int i = 0; i + stride < C; i += 2 * stride) {
      const bool gt = bv[i + stride] > bv[i];
      bv[i] = gt ? bv[i + stride] : bv[i];
      bi[i] = gt ? bi[i + stride] : bi[i];
    }
  }
  best = bv[0];
  arg = bi[0];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` holds one float per warp.
template <int BLOCK>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < BLOCK / 32) ? red[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace diga
