// a6 / a7 — per-class centroid accumulation and running update.
// Replaces Class_Features.calculate_mean_vector / _by_output / update_objective_SingleVector
// (calc_centroids.py:97-164 of the reference) and the one-hot helper util/utils.py:158-163.
//
// Pipeline per batch of N images (all on one stream, no host sync):
//   assign : logits [N,C,hw] (+ optional fp32 labels [N,1,hw]) -> cls u8 [N,hw], counts i32 [N,C]
//   accum  : feat [N,D,hw], cls -> sums [N,C,D]               (the HBM-bound kernel: D*4 B per feature px)
//   means  : sums, counts -> vec [N,C,D], vecsum [N,C], valid [N,C]
//   update : vec -> objective_vectors [C,D], objective_num [C]  (sequential over images, parallel over (c,d))
//
// accum is a segmented reduction keyed by the per-pixel class.  Features are NCHW, so for one channel
// row the pixels are contiguous: a warp reads 32 adjacent pixels of R=4 consecutive channel rows
// (4 coalesced 128-byte requests) and each lane adds its 4 values, as one 128-bit shared-memory
// read-modify-write, into a LANE-PRIVATE accumulator acc[warp][class][lane] (float4).  Lane-private
// columns make the update conflict-free whatever the label pattern is — segmentation maps are
// piecewise constant, so a shared per-class cell would see 32-way same-address collisions.
// A CTA owns (image, 4-row group) and covers all pixels of that image, so every sums[n][c][d] has
// exactly one writer: no global atomics, bitwise deterministic.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace diga {

int tunable(const char* name, int dflt);
__device__ __forceinline__ void pdl_wait();
__device__ __forceinline__ void pdl_launch_dependents();

// ------------------------------------------------------------------------------------------------
// assign
// ------------------------------------------------------------------------------------------------
// Phase-shifted class words for the 128-bit accumulation kernel.  That kernel reads a channel row from the 128-byte line
// below its first element, i.e. shifted by s floats (s = 0..31), so the four pixels of its aligned quad t are
// 4t-s .. 4t-s+3: which bytes of the class map belong to one quad depends on s mod 4.  M_phi (phi = 0..3) is the class map
// delayed by phi bytes behind kClswFront words of padding, so quad t of a row with shift s reads ONE aligned 32-bit word,
// word kClswFront + t - (s >> 2) of M_(s & 3).  Pixels outside the row and gated-out pixels (255) carry the value `nclass`,
// the index of a dummy accumulator slot — the accumulation loop needs no bounds or validity test at all.
constexpr int kClswFront = 8;                          // words of front padding: covers t < (s >> 2), s <= 31
__host__ __device__ inline int64_t clsw_words(int64_t hw) { return kClswFront + (hw + 6) / 4 + 1; }
__device__ __forceinline__ void clsw_put(uint8_t* m, int64_t wp, int64_t p, int c) {
#pragma unroll
  for (int phi = 0; phi < 4; ++phi) m[phi * wp * 4 + 4 * kClswFront + p + phi] = (uint8_t)c;
}
__device__ __forceinline__ void clsw_pads(uint8_t* m, int64_t wp, int64_t hw, int nclass, int tid, int nthreads) {
  for (int phi = 0; phi < 4; ++phi) {
    uint8_t* mp = m + phi * wp * 4;
    const int64_t head = 4 * kClswFront + phi, tail0 = 4 * kClswFront + hw + phi, total = wp * 4;
    for (int64_t j = tid; j < head + (total - tail0); j += nthreads) mp[j < head ? j : tail0 + (j - head)] = (uint8_t)nclass;
  }
}

// kC: compile-time class count (19, 16, or 32 = padded generic, see DIGA_DISPATCH_C): the per-pixel loads are unrolled, so
// all of a pixel's class planes are requested before the first comparison (the rolled loop exposed one latency per class:
// 6 us for [8,19,65,129]).
template <int BLOCK, int kC, bool kPad>
__global__ void __launch_bounds__(BLOCK)
centroid_assign_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int nclass, int64_t hw,
                       uint8_t* __restrict__ cls, int32_t* __restrict__ counts, uint32_t* __restrict__ clsw,
                       const int64_t* __restrict__ labels_full = nullptr,
                       int w = 0, int HH = 0, int WW = 0, float sy = 0.f, float sx = 0.f) {
  __shared__ int hist[DIGA_MAX_CLASSES];
  pdl_launch_dependents();      // the accumulation kernel may start fetching features (which this kernel does not write)
  if (threadIdx.x < DIGA_MAX_CLASSES) hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t img = blockIdx.y;
  const float* lg = logits + img * nclass * hw;
  const int64_t wp = clsw_words(hw);
  uint8_t* mw = clsw ? reinterpret_cast<uint8_t*>(clsw + img * 4 * wp) : nullptr;
  if (mw && blockIdx.x == 0) clsw_pads(mw, wp, hw, nclass, threadIdx.x, BLOCK);
  for (int64_t base = (int64_t)blockIdx.x * BLOCK; base < hw; base += (int64_t)gridDim.x * BLOCK) {
    const int64_t p = base + threadIdx.x;
    int c = 255;
    if (p < hw) {
      float z[kC];
#pragma unroll
      for (int k = 0; k < kC; ++k) z[k] = (!kPad || k < nclass) ? __ldg(lg + (int64_t)k * hw + p) : -INFINITY;
      float m = z[0];
      int am = 0;
#pragma unroll
      for (int k = 1; k < kC; ++k) {
        if (z[k] > m) {   // argmax(softmax(out)) == first index of the max logit (calc_centroids.py:121-122)
          m = z[k];
          am = k;
        }
      }
      c = am;
      if (labels_full != nullptr) {
        // the reference's label down-sampling folded in (self_training.py:327-330, :336-337: .float() + F.interpolate(
        // mode='nearest') of the [B,H,W] int64 map): ATen's nearest rule src = min((int)floorf(dst * scale), in - 1) with
        // scale = (float)in / out, then the same gate as below on the fp32-rounded label value.
        const int y = (int)(p / w), x = (int)(p - (int64_t)y * w);
        const int Y = min((int)floorf((float)y * sy), HH - 1), X = min((int)floorf((float)x * sx), WW - 1);
        const float lf = (float)__ldg(labels_full + ((int64_t)img * HH + Y) * WW + X);
        const bool in_range = lf < (float)nclass;
        if (!(in_range && (long long)lf == (long long)am)) c = 255;
      } else if (labels != nullptr) {
        // process_label(labels) * process_label(argmax): the pixel counts for class t iff
        // long(label) == t == argmax, t < C (utils.py:161-162, calc_centroids.py:126-127).
        const float lf = __ldg(labels + img * hw + p);
        const bool in_range = lf < (float)nclass;
        if (!(in_range && (long long)lf == (long long)am)) c = 255;
      }
      cls[img * hw + p] = (uint8_t)c;
      if (mw) clsw_put(mw, wp, p, c == 255 ? nclass : c);
    }
    // warp-aggregated histogram: one shared atomic per distinct class per warp
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    if (c != 255 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[c], __popc(peers));
  }
  __syncthreads();
  if ((int)threadIdx.x < nclass && hist[threadIdx.x]) atomicAdd(&counts[img * nclass + threadIdx.x], hist[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// accum
// ------------------------------------------------------------------------------------------------
// Persistent grid for `items` equal work items: among 3..max resident CTAs per SM pick the count whose last round is
// fullest (4096 items on 148 SMs: 5 CTAs/SM leaves a 5.53 -> 6 round, 8 % idle; 4 CTAs/SM gives 6.92 -> 7, 1 %).
static int64_t balanced_grid(int64_t items, int blocks_per_sm) {
  const int64_t sms = sm_count();
  if (items <= sms * blocks_per_sm) return items < 1 ? 1 : items;
  int64_t best = sms * blocks_per_sm;
  double best_eff = 0.0;
  for (int k = blocks_per_sm; k >= (blocks_per_sm > 3 ? 3 : 1); --k) {
    const int64_t g = sms * k;
    const int64_t rounds = (items + g - 1) / g;
    const double eff = (double)items / (double)(rounds * g);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = g;
    }
  }
  return best;
}

template <int R> struct AccT;
template <> struct AccT<4> { using type = float4; };
template <> struct AccT<2> { using type = float2; };

// One batch = U pixel stripes x R rows of loads, all issued before any is consumed.
template <int R, int U>
struct AccumBatch {
  float x[U][R];
  int cid[U];
};

template <int R, int WARPS, int U>
__device__ __forceinline__ void accum_load(AccumBatch<R, U>& b, const float* const (&rows)[R], const uint8_t* cl, int64_t base,
                                           int64_t hw) {
  constexpr int STRIPE = WARPS * 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t p = base + (int64_t)u * STRIPE;
    const bool ok = p < hw;
    b.cid[u] = ok ? (int)__ldg(cl + p) : 255;
#pragma unroll
    for (int r = 0; r < R; ++r) b.x[u][r] = ok ? ld_stream<1>(rows[r] + p).v[0] : 0.f;
  }
}

// Run combining: the pixels one lane visits (p, p+STRIPE, p+2*STRIPE, ...) are vertical neighbours in the image, and
// segmentation maps are piecewise constant, so consecutive visits usually hit the same class.  The values of a run are
// summed in registers and flushed to the lane-private shared-memory cell only when the class changes: on realistic
// maps that removes most of the shared-memory read-modify-writes (the L1/shared data pipe is this kernel's busiest
// unit, 65 % in ncu), on i.i.d. random classes it degenerates to one RMW per pixel as before.
template <int R, typename acc_t>
struct AccumRun {
  int cls = 255;
  float v[R];
  __device__ __forceinline__ void flush(acc_t* my, int nclass) {
    if (cls < nclass) {
      acc_t a = my[cls * 32];
      if constexpr (R == 4) {
        a.x += v[0]; a.y += v[1]; a.z += v[2]; a.w += v[3];
      } else {
        a.x += v[0]; a.y += v[1];
      }
      my[cls * 32] = a;
    }
  }
};

template <int R, int U, typename acc_t>
__device__ __forceinline__ void accum_apply(const AccumBatch<R, U>& b, acc_t* my, int nclass, AccumRun<R, acc_t>& run) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (b.cid[u] == run.cls) {
#pragma unroll
      for (int r = 0; r < R; ++r) run.v[r] += b.x[u][r];
    } else {
      run.flush(my, nclass);
      run.cls = b.cid[u];
#pragma unroll
      for (int r = 0; r < R; ++r) run.v[r] = b.x[u][r];
    }
  }
}

// PIPE: software pipelining — the loads of batch i+1 are issued before the shared-memory updates of batch i, so a
// warp keeps R*U..2*R*U 32-bit loads in flight instead of draining to zero every iteration.
template <int R, int WARPS, int U, bool PIPE>
__global__ void __launch_bounds__(WARPS * 32)
centroid_accum_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ cls, int nclass, int64_t n, int64_t D,
                      int64_t hw, float* __restrict__ sums) {
  using acc_t = typename AccT<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  acc_t* acc = reinterpret_cast<acc_t*>(smem_raw);   // [WARPS][nclass][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t groups = (D + R - 1) / R;
  const int64_t items = n * groups;
  acc_t* my = acc + (size_t)warp * nclass * 32 + lane;
  constexpr int64_t STEP = (int64_t)WARPS * 32 * U;

  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t img = item / groups;
    const int64_t d0 = (item - img * groups) * R;
    const uint8_t* cl = cls + img * hw;
    const float* rows[R];
    bool row_ok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      row_ok[r] = d0 + r < D;
      rows[r] = feat + (img * D + (row_ok[r] ? d0 + r : d0)) * hw;
    }
    const int64_t first = warp * 32 + lane;
    AccumBatch<R, U> a, b;
    accum_load<R, WARPS, U>(a, rows, cl, first, hw);          // in flight while the accumulators are cleared
    acc_t zero;
    if constexpr (R == 4) zero = make_float4(0.f, 0.f, 0.f, 0.f); else zero = make_float2(0.f, 0.f);
    for (int c = 0; c < nclass; ++c) my[c * 32] = zero;
    // (own lane-private cells only: no barrier needed before the main loop)

    AccumRun<R, acc_t> run;
    if constexpr (PIPE) {
      // bases are warp-uniform up to +lane, so the loop trip count is uniform within a warp
      for (int64_t base = first; base - lane < hw; base += 2 * STEP) {
        accum_load<R, WARPS, U>(b, rows, cl, base + STEP, hw);
        accum_apply<R, U>(a, my, nclass, run);
        accum_load<R, WARPS, U>(a, rows, cl, base + 2 * STEP, hw);
        accum_apply<R, U>(b, my, nclass, run);
      }
    } else {
      for (int64_t base = first; base - lane < hw; base += STEP) {
        accum_apply<R, U>(a, my, nclass, run);
        accum_load<R, WARPS, U>(a, rows, cl, base + STEP, hw);
      }
    }
    run.flush(my, nclass);
    __syncthreads();
    // cross-warp then cross-lane reduction; classes are dealt round-robin to warps
    for (int c = warp; c < nclass; c += WARPS) {
      float s[R];
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const acc_t v = acc[((size_t)w * nclass + c) * 32 + lane];
        if constexpr (R == 4) {
          s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        } else {
          s[0] += v.x; s[1] += v.y;
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = warp_sum(s[r]);
      if (lane == 0) {
        float* o = sums + (img * nclass + c) * D + d0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (row_ok[r]) o[r] = s[r];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// accum, 128-bit variant.  A lane owns 4 ADJACENT pixels of 4 channel rows (four aligned LDG.128).  The four rows
// are chosen RS apart so that they share one alignment: with hw odd (65x129) row d starts (d*hw) mod 32 floats past a
// 128-byte line, rows d, d+32, d+64, d+96 start at the same phase s, and the aligned quad t of every row covers
// pixels 4t-s .. 4t-s+3, so every warp request is exactly four whole cache lines.  Adjacent pixels usually carry the same class (segmentation maps are piecewise constant):
// when all four agree their values are summed in registers and ONE shared-memory read-modify-write is issued instead
// of four, which takes the kernel off the shared-memory pipe on realistic label maps.
// ------------------------------------------------------------------------------------------------
// Row stride RS (a power of two) such that rows d and d+RS start at the same phase within a 128-byte line
// (RS*hw % 32 == 0), falling back to 16-byte phase equality (RS*hw % 4 == 0); 0 if D cannot be tiled by 4*RS rows.
// `unit_out`: the alignment unit found (32 or 4 floats); the phase of row q is (q*hw) mod unit.
static int quad_row_stride(int64_t hw, int64_t D, int* unit_out = nullptr) {
  for (int unit = 32; unit >= 4; unit /= 8) {       // 32 floats = one line, then 4 floats = one 128-bit access
    int rs = unit;
    while (rs > 1 && ((rs / 2) * hw) % unit == 0) rs /= 2;
    if (D % (4 * rs) == 0) {
      if (unit_out) *unit_out = unit;
      return rs;
    }
  }
  return 0;
}

template <int U>
struct QuadBatch {
  float4 x[U][4];
  uint32_t cls4[U];   // 4 class bytes, 0xff = skip
};

template <int WARPS, int U>
__device__ __forceinline__ void quad_load(QuadBatch<U>& b, const float* const (&rows)[4], const uint8_t* cl, int64_t t0,
                                          int64_t T, int s, int64_t hw) {
  constexpr int STRIPE = WARPS * 32;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t t = t0 + (int64_t)u * STRIPE;
    const bool ok = t < T;
    uint32_t c4 = 0xffffffffu;
    if (ok) {
      const int64_t p0 = 4 * t - s;
      c4 = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t p = p0 + i;
        const uint32_t c = (p >= 0 && p < hw) ? (uint32_t)__ldg(cl + p) : 0xffu;
        c4 |= c << (8 * i);
      }
    }
    b.cls4[u] = c4;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (ok) {
        const Vec<4> v = ld_stream<4>(rows[r] + 4 * t);
        b.x[u][r] = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
      } else {
        b.x[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

__device__ __forceinline__ float f4_get(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <int U>
__device__ __forceinline__ void quad_apply(const QuadBatch<U>& b, float4* my, int nclass) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t c4 = b.cls4[u];
    const uint32_t c0 = c4 & 0xffu;
    if (c4 == c0 * 0x01010101u) {            // all four pixels in one class (or all skipped)
      if ((int)c0 < nclass) {
        float4 v = my[c0 * 32];
        v.x += (b.x[u][0].x + b.x[u][0].y) + (b.x[u][0].z + b.x[u][0].w);
        v.y += (b.x[u][1].x + b.x[u][1].y) + (b.x[u][1].z + b.x[u][1].w);
        v.z += (b.x[u][2].x + b.x[u][2].y) + (b.x[u][2].z + b.x[u][2].w);
        v.w += (b.x[u][3].x + b.x[u][3].y) + (b.x[u][3].z + b.x[u][3].w);
        my[c0 * 32] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (int)((c4 >> (8 * i)) & 0xffu);
        if (c < nclass) {
          float4 v = my[c * 32];
          v.x += f4_get(b.x[u][0], i);
          v.y += f4_get(b.x[u][1], i);
          v.z += f4_get(b.x[u][2], i);
          v.w += f4_get(b.x[u][3], i);
          my[c * 32] = v;
        }
      }
    }
  }
}

template <int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32)
centroid_accum_quad_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ cls, int nclass, int64_t n, int64_t D,
                           int64_t hw, int RS, int unit_mask, float* __restrict__ sums) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* acc = reinterpret_cast<float4*>(smem_raw);   // [WARPS][nclass][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t groups = D / 4;
  const int64_t items = n * groups;
  float4* my = acc + (size_t)warp * nclass * 32 + lane;
  constexpr int64_t STEP = (int64_t)WARPS * 32 * U;

  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t img = item / groups;
    const int64_t g = item - img * groups;
    const int64_t blk = g / RS, q = g - blk * RS;        // rows blk*4*RS + q + j*RS, j = 0..3
    const int64_t d0 = blk * 4 * RS + q;
    const int s = (int)((q * hw) & unit_mask);            // common alignment phase (floats) of the four rows
    const int64_t T = (hw + s + 3) >> 2;
    const uint8_t* cl = cls + img * hw;
    const float* rows[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) rows[j] = feat + (img * D + d0 + (int64_t)j * RS) * hw - s;
    const int64_t first = warp * 32 + lane;
    QuadBatch<U> a, b;
    quad_load<WARPS, U>(a, rows, cl, first, T, s, hw);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nclass; ++c) my[c * 32] = zero;
    for (int64_t t0 = first; t0 - lane < T; t0 += 2 * STEP) {
      quad_load<WARPS, U>(b, rows, cl, t0 + STEP, T, s, hw);
      quad_apply<U>(a, my, nclass);
      quad_load<WARPS, U>(a, rows, cl, t0 + 2 * STEP, T, s, hw);
      quad_apply<U>(b, my, nclass);
    }
    __syncthreads();
    for (int c = warp; c < nclass; c += WARPS) {
      float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const float4 v = acc[((size_t)w * nclass + c) * 32 + lane];
        sx += v.x; sy += v.y; sz += v.z; sw += v.w;
      }
      sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sw = warp_sum(sw);
      if (lane == 0) {
        float* o = sums + (img * nclass + c) * D + d0;
        o[0] = sx; o[RS] = sy; o[2 * RS] = sz; o[3 * RS] = sw;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// accum, lean 128-bit variant (round 2).  Same decomposition as the quad kernel above — CTA = (image, four channel rows
// that share an alignment phase), lane = four adjacent pixels, lane-private float4 accumulators — with everything that is
// not a load or an add taken out of the loop:
//   * the class bytes of a quad arrive as ONE aligned 32-bit word (phase-shifted class words, see clsw_words) in which
//     out-of-row and gated-out pixels name a dummy accumulator slot: no bounds test, no validity test, no byte loads;
//   * quads past the end of the row re-read the last full quad against an all-dummy class word (clamped address), the
//     <= 3 pixels left over behind the last full quad are added by a scalar tail — the 128-bit loads never leave the row
//     of the LAST channel, so nothing is read beyond the tensor;
//   * the first batch of the NEXT work item is requested before the cross-warp reduction of the current one, so the
//     memory pipe does not drain at every item boundary;
//   * the per-class reduction uses a 4-value butterfly (6 shuffles instead of 20) and skips classes absent from the
//     image (per-image counts from the assign kernel), as does the clearing of the accumulators.
// ncu of the round-1 kernel: 374 warp instructions per 4 KB batch, issue slots 41 % busy with every warp waiting on a
// load; this loop issues ~100.
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (the a6 -> a7 chain, diga_centroid_chain): a kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor drains; pdl_wait() blocks until the predecessor grid has completed
// and its writes are visible (a no-op for a normally launched kernel), pdl_launch_dependents() lets the successor's CTAs be
// scheduled as soon as every CTA of this grid has passed it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg128_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Sum of a float4 over the 32 lanes with 6 shuffles: lanes 0-7 end with the x total, 8-15 y, 16-23 z, 24-31 w.
__device__ __forceinline__ float warp_sum4(const float4& v, int lane) {
  const bool hi16 = lane & 16, hi8 = lane & 8;
  float a = hi16 ? v.z : v.x, b = hi16 ? v.w : v.y;
  a += __shfl_xor_sync(0xffffffffu, hi16 ? v.x : v.z, 16);
  b += __shfl_xor_sync(0xffffffffu, hi16 ? v.y : v.w, 16);
  float k = hi8 ? b : a;
  k += __shfl_xor_sync(0xffffffffu, hi8 ? a : b, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

struct LeanUnit {            // one quad of one lane: 4 rows x 4 adjacent pixels + their 4 class bytes
  float4 x[4];
  uint32_t cw;
  bool in_row;               // false: the lane is past the end of the row (clamped re-read, counts for the dummy class)
};

struct LeanItem {
  const float4* row[4];   // the four rows, shifted back by s floats (to the 16-byte / 128-byte boundary below their start)
  const uint32_t* cw;     // class word of quad t at cw[t]
  int T;                  // full quads in the row
};

template <bool FEATURES = true, bool CLASSES = true>
__device__ __forceinline__ void lean_load(LeanUnit& u, const LeanItem& it, int t, uint32_t dummy4) {
  const unsigned tt = (unsigned)min(t, it.T - 1);  // lanes past the end re-read the last quad against a dummy class word
  if constexpr (CLASSES) {
    u.cw = __ldg(it.cw + tt);                      // (selected against the dummy word at apply time: no wait here)
    u.in_row = t < it.T;
  }
  if constexpr (FEATURES) {
#pragma unroll
    for (int j = 0; j < 4; ++j) u.x[j] = ldg128_stream(reinterpret_cast<const float*>(it.row[j] + tt));
  }
}

__device__ __forceinline__ void f4_add_if(float4& a, const float4& b, bool p) {
  if (p) {
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
}

// Branch-free form: the four accumulator cells are read first (four independent LDS), every pixel's column is folded
// into the FIRST pixel of the quad that has the same class (predicated adds), and only first occurrences are written
// back — the written cells are distinct, so the four read-modify-writes no longer form a dependent chain.
__device__ __forceinline__ void lean_apply_merged(const LeanUnit& u, uint32_t my, uint32_t dummy4) {
  const uint32_t cw = u.in_row ? u.cw : dummy4;
  const uint32_t c0 = cw & 0xffu, c1 = (cw >> 8) & 0xffu, c2 = (cw >> 16) & 0xffu, c3 = cw >> 24;
  const uint32_t a0 = my + (c0 << 9), a1 = my + (c1 << 9), a2 = my + (c2 << 9), a3 = my + (c3 << 9);
  float4 s0 = lds128(a0), s1 = lds128(a1), s2 = lds128(a2), s3 = lds128(a3);
  // pixel i as a column over the four rows
  float4 v0 = make_float4(u.x[0].x, u.x[1].x, u.x[2].x, u.x[3].x), v1 = make_float4(u.x[0].y, u.x[1].y, u.x[2].y, u.x[3].y);
  float4 v2 = make_float4(u.x[0].z, u.x[1].z, u.x[2].z, u.x[3].z), v3 = make_float4(u.x[0].w, u.x[1].w, u.x[2].w, u.x[3].w);
  const bool e01 = c0 == c1, e02 = c0 == c2, e03 = c0 == c3, e12 = c1 == c2, e13 = c1 == c3, e23 = c2 == c3;
  f4_add_if(v2, v3, e23);                      // fold from the back so that chains (c1 == c2 == c3) accumulate
  f4_add_if(v1, v3, e13 && !e23);
  f4_add_if(v1, v2, e12);
  f4_add_if(v0, v3, e03 && !e13 && !e23);
  f4_add_if(v0, v2, e02 && !e12);
  f4_add_if(v0, v1, e01);
  s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
  s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
  s2.x += v2.x; s2.y += v2.y; s2.z += v2.z; s2.w += v2.w;
  s3.x += v3.x; s3.y += v3.y; s3.z += v3.z; s3.w += v3.w;
  sts128(a0, s0);
  if (!e01) sts128(a1, s1);
  if (!(e02 || e12)) sts128(a2, s2);
  if (!(e03 || e13 || e23)) sts128(a3, s3);
}

// MODE 0: four sequential read-modify-writes; 1: + one-RMW fast path when the four pixels share a class (a per-lane
// branch: pays on single-class maps, costs on mixed ones); 2: the branch-free merged form above; 3 (default): as 0 but
// pixels of the dummy class branch around their RMW and the dummy slot is dropped (19 instead of 20 slots: four CTAs then
// fit the 164 KB shared-memory carve-out, which leaves the L1 that buffers the in-flight loads 92 KB instead of 60).
template <int MODE>
__device__ __forceinline__ void lean_apply(const LeanUnit& u, uint32_t my, uint32_t dummy4) {
  if constexpr (MODE == 2) {
    lean_apply_merged(u, my, dummy4);
    return;
  }
  constexpr bool SPLAT = MODE == 1;
  const uint32_t cw = u.in_row ? u.cw : dummy4;
  if constexpr (MODE == 3) {                   // no dummy slot: pixels of the dummy class are predicated off
    const uint32_t dummy = dummy4 & 0xffu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t c = (cw >> (8 * i)) & 0xffu;
      if (c != dummy) {
        const uint32_t a = my + (c << 9);
        float4 v = lds128(a);
        v.x += f4_get(u.x[0], i);
        v.y += f4_get(u.x[1], i);
        v.z += f4_get(u.x[2], i);
        v.w += f4_get(u.x[3], i);
        sts128(a, v);
      }
    }
    return;
  }
  if (SPLAT && cw == (cw & 0xffu) * 0x01010101u) {      // the four pixels share a class: one read-modify-write
    const uint32_t a = my + ((cw & 0xffu) << 9);
    float4 v = lds128(a);
    v.x += (u.x[0].x + u.x[0].y) + (u.x[0].z + u.x[0].w);
    v.y += (u.x[1].x + u.x[1].y) + (u.x[1].z + u.x[1].w);
    v.z += (u.x[2].x + u.x[2].y) + (u.x[2].z + u.x[2].w);
    v.w += (u.x[3].x + u.x[3].y) + (u.x[3].z + u.x[3].w);
    sts128(a, v);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = my + (((cw >> (8 * i)) & 0xffu) << 9);
      float4 v = lds128(a);
      v.x += f4_get(u.x[0], i);
      v.y += f4_get(u.x[1], i);
      v.z += f4_get(u.x[2], i);
      v.w += f4_get(u.x[3], i);
      sts128(a, v);
    }
  }
}

// P = quads per lane and batch.  RING = false (default): a batch of P quads (4*P 128-bit loads) is requested, then
// consumed, then the next batch is requested — ptxas puts every LDG.128 of the loop on ONE scoreboard and a scoreboard
// wait returns only when ALL its loads have landed, so any scheme that re-requests a unit right after consuming it
// (RING = true, the register ring this kernel started with; also the double buffer of the round-1 kernel) still pays a
// full memory latency per UNIT: the wait for the oldest unit also waits for the one requested a moment ago.  Measured
// (profiles/r02_accum_sweep.jsonl): ring 0.73, batches 0.8x.  Latency is hidden across warps, not inside one.
template <int WARPS, int P, int MINB, int MODE, bool RING>
__global__ void __launch_bounds__(WARPS * 32, MINB)
centroid_accum_lean_kernel(const float* __restrict__ feat, const uint32_t* __restrict__ clsw, const uint8_t* __restrict__ cls,
                           const int32_t* __restrict__ counts, int nclass, int64_t n, int64_t D, int64_t hw, int RS, int unit_mask,
                           float* __restrict__ sums) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int STRIDE = WARPS * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slots = MODE == 3 ? nclass : nclass + 1;              // + the dummy slot
  const uint32_t acc0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t my = acc0 + (uint32_t)((warp * slots * 32 + lane) * 16);   // slot c of this lane: my + c * 512
  const uint32_t dummy4 = (uint32_t)nclass * 0x01010101u;
  const int64_t groups = D / 4, items = n * groups, wp = clsw_words(hw);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  auto setup = [&](int64_t item, LeanItem& it, int64_t& img, int64_t& d0, int& s) {
    img = item / groups;
    const int64_t g = item - img * groups;
    const int64_t blk = g / RS, q = g - blk * RS;                 // rows blk*4*RS + q + j*RS, j = 0..3
    d0 = blk * 4 * RS + q;
    s = (int)((q * hw) & unit_mask);                              // common alignment phase (floats) of the four rows
    const float* r0 = feat + (img * D + d0) * hw - s;
#pragma unroll
    for (int j = 0; j < 4; ++j) it.row[j] = reinterpret_cast<const float4*>(r0 + (int64_t)j * RS * hw);
    it.cw = clsw + (img * 4 + (s & 3)) * wp + kClswFront - (s >> 2);
    it.T = (int)((hw + s) >> 2);
  };

  int64_t item = blockIdx.x;
  if (item >= items) return;
  LeanItem it;
  int64_t img, d0;
  int s;
  setup(item, it, img, d0, s);
  LeanUnit ring[P];
  const int first = warp * 32 + lane;
  // first batch: the features are requested BEFORE the grid dependency is awaited (in the chain the predecessor is the
  // assign kernel, which writes the class words and counts but not the features), the class words after it
#pragma unroll
  for (int k = 0; k < P; ++k) lean_load<true, false>(ring[k], it, first + k * STRIDE, dummy4);   // (rounds past the row: clamped)
  pdl_wait();
  pdl_launch_dependents();
#pragma unroll
  for (int k = 0; k < P; ++k) lean_load<false, true>(ring[k], it, first + k * STRIDE, dummy4);

  while (true) {
    // classes present in this image (per-image counts): absent ones are neither cleared nor reduced nor written
    uint32_t present = 0xffffffffu;
    if (counts != nullptr) present = __ballot_sync(0xffffffffu, lane < nclass && __ldg(counts + img * nclass + lane) > 0);
    for (int c = 0; c < nclass; ++c)
      if ((present >> c) & 1u) sts128(my + (c << 9), zero);
    if (MODE != 3) sts128(my + (nclass << 9), zero);
    // (own lane-private cells only: no barrier needed before the main loop)

    const int rounds = (it.T + STRIDE - 1) / STRIDE;                // CTA-uniform
    if constexpr (RING) {
      for (int r = 0; r < rounds; r += P) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
          if (r + k < rounds) {
            lean_apply<MODE>(ring[k], my, dummy4);
            if (r + k + P < rounds) lean_load(ring[k], it, first + (r + k + P) * STRIDE, dummy4);
          }
        }
      }
    } else {
      for (int r = 0; r < rounds; r += P) {
#pragma unroll
        for (int k = 0; k < P; ++k)
          if (r + k < rounds) lean_apply<MODE>(ring[k], my, dummy4);
#pragma unroll
        for (int k = 0; k < P; ++k)
          if (r + P + k < rounds) lean_load(ring[k], it, first + (r + P + k) * STRIDE, dummy4);
      }
    }
    // the (hw + s) & 3 pixels behind the last full quad
    const int rem = (int)((hw + s) & 3);
    if (warp == 0 && lane < rem) {
      const int64_t p = hw - rem + lane;
      const int c = cls[img * hw + p];
      if (c < nclass) {
        const float* f = feat + (img * D + d0) * hw + p;
        const int64_t row_stride = (int64_t)RS * hw;
        const uint32_t ad = my + (c << 9);
        float4 v = lds128(ad);
        v.x += __ldg(f);
        v.y += __ldg(f + row_stride);
        v.z += __ldg(f + 2 * row_stride);
        v.w += __ldg(f + 3 * row_stride);
        sts128(ad, v);
      }
    }
    // next item: request its first P rounds now, consume them after the reduction below
    const int64_t out_base = (img * nclass) * D + d0;
    const int64_t next = item + gridDim.x;
    const bool more = next < items;
    if (more) {
      setup(next, it, img, d0, s);
#pragma unroll
      for (int k = 0; k < P; ++k) lean_load(ring[k], it, first + k * STRIDE, dummy4);
    }
    __syncthreads();
    for (int c = warp; c < nclass; c += WARPS) {
      if (!((present >> c) & 1u)) continue;
      float4 t = zero;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const float4 v = lds128(acc0 + (uint32_t)(((w * slots + c) * 32 + lane) * 16));
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      const float tot = warp_sum4(t, lane);
      if ((lane & 7) == 0) sums[out_base + (int64_t)c * D + (int64_t)(lane >> 3) * RS] = tot;
    }
    if (!more) break;
    item = next;
    __syncthreads();
  }
}

// class words from a plain u8 class map (callers that did not get them from the assign kernel)
__global__ void centroid_clsw_build_kernel(const uint8_t* __restrict__ cls, int nclass, int64_t hw, uint32_t* __restrict__ clsw) {
  const int64_t wp = clsw_words(hw);
  const int64_t img = blockIdx.z;
  const int phi = blockIdx.y;
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= wp) return;
  uint32_t word = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int64_t p = wi * 4 + b - 4 * kClswFront - phi;
    int c = nclass;
    if (p >= 0 && p < hw) {
      c = cls[img * hw + p];
      if (c >= nclass) c = nclass;
    }
    word |= (uint32_t)c << (8 * b);
  }
  clsw[(img * 4 + phi) * wp + wi] = word;
}

static thread_local bool g_chain_pdl = false;     // set by diga_centroid_chain around its launches (same thread): launch as programmatic dependents

template <int WARPS, int P, int MINB, int MODE, bool RING = false>
static int launch_accum_lean(const float* feat, const uint32_t* clsw, const uint8_t* cls, const int32_t* counts, int nclass,
                             int64_t n, int64_t D, int64_t hw, float* sums, cudaStream_t st) {
  auto kern = centroid_accum_lean_kernel<WARPS, P, MINB, MODE, RING>;
  const size_t smem = (size_t)WARPS * (nclass + (MODE == 3 ? 0 : 1)) * 32 * sizeof(float4);
  static size_t configured_dev[64] = {0};
  static int blocks_dev[64] = {0};
  const int slot = device_slot();
  if (configured_dev[slot] != smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("centroid_accum: cannot reserve %zu bytes of shared memory", smem);
      return DIGA_ERR_CUDA;
    }
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, WARPS * 32, smem);
    blocks_dev[slot] = b > 0 ? b : 1;
    configured_dev[slot] = smem;
  }
  int unit = 0;
  const int RS = quad_row_stride(hw, D, &unit);
  const int64_t items = n * (D / 4);
  const int64_t full = (int64_t)sm_count() * blocks_dev[slot];
  const int64_t grid = tunable("accum_balance", 1) ? balanced_grid(items, blocks_dev[slot]) : (items < full ? items : full);
  if (g_chain_pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(WARPS * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, feat, clsw, cls, counts, nclass, n, D, hw, RS, unit - 1, sums);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("centroid_accum_lean_kernel: launch failed: %s", cudaGetErrorString(e));
      return DIGA_ERR_CUDA;
    }
  } else {
    kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(feat, clsw, cls, counts, nclass, n, D, hw, RS, unit - 1, sums);
  }
  DIGA_CHECK_LAUNCH("centroid_accum_lean_kernel");
  return DIGA_OK;
}

template <int WARPS, int U>
static int launch_accum_quad(const float* feat, const uint8_t* cls, int nclass, int64_t n, int64_t D, int64_t hw, float* sums,
                             cudaStream_t st) {
  auto kern = centroid_accum_quad_kernel<WARPS, U>;
  const size_t smem = (size_t)WARPS * nclass * 32 * sizeof(float4);
  static size_t configured = 0;
  static int blocks_per_sm = 0;
  if (configured != smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("centroid_accum: cannot reserve %zu bytes of shared memory", smem);
      return DIGA_ERR_CUDA;
    }
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, WARPS * 32, smem);
    blocks_per_sm = b > 0 ? b : 1;
    configured = smem;
  }
  int unit = 0;
  const int RS = quad_row_stride(hw, D, &unit);
  const int64_t items = n * (D / 4);
  const int64_t grid = tunable("accum_balance", 1) ? balanced_grid(items, blocks_per_sm) : (items < (int64_t)sm_count() * blocks_per_sm ? items : (int64_t)sm_count() * blocks_per_sm);
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(feat, cls, nclass, n, D, hw, RS, unit - 1, sums);
  DIGA_CHECK_LAUNCH("centroid_accum_quad_kernel");
  return DIGA_OK;
}

template <int R, int WARPS, int U, bool PIPE>
static int launch_accum(const float* feat, const uint8_t* cls, int nclass, int64_t n, int64_t D, int64_t hw, float* sums,
                        cudaStream_t st) {
  auto kern = centroid_accum_kernel<R, WARPS, U, PIPE>;
  const size_t smem = (size_t)WARPS * nclass * 32 * sizeof(typename AccT<R>::type);
  static size_t configured = 0;
  static int blocks_per_sm = 0;
  if (configured != smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("centroid_accum: cannot reserve %zu bytes of shared memory", smem);
      return DIGA_ERR_CUDA;
    }
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, WARPS * 32, smem);
    blocks_per_sm = b > 0 ? b : 1;
    configured = smem;
  }
  const int64_t items = n * ((D + R - 1) / R);
  const int64_t grid = tunable("accum_balance", 1) ? balanced_grid(items, blocks_per_sm) : (items < (int64_t)sm_count() * blocks_per_sm ? items : (int64_t)sm_count() * blocks_per_sm);
  kern<<<(unsigned)grid, WARPS * 32, smem, st>>>(feat, cls, nclass, n, D, hw, sums);
  DIGA_CHECK_LAUNCH("centroid_accum_kernel");
  return DIGA_OK;
}

// ------------------------------------------------------------------------------------------------
// means
// ------------------------------------------------------------------------------------------------
// Per-image class mean of one channel: the reference forms avgpool(feat * mask) / avgpool(mask) = (sum / hw) / (count / hw)
// (calc_centroids.py:129,141) — three fp32 roundings around the exact value sum / count, on top of torch's own summation
// order.  One correctly rounded division is at least as close to it (parity bar for the vectors: 1e-5) and is the ONE
// expression every kernel below uses, so that all paths (means, finish, means + scatter) produce the same bits.  Round 1 used
// the two-division form; the finish kernel is a chain of dependent arithmetic on two warps per scheduler, and the divisions
// were a third of its instructions.
__device__ __forceinline__ float class_mean(float sum, int cnt) { return __fdiv_rn(sum, (float)cnt); }

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_means_kernel(const float* __restrict__ sums, const int32_t* __restrict__ counts, int64_t C, int64_t D,
                      int64_t hw, float* __restrict__ vec, float* __restrict__ vecsum, uint8_t* __restrict__ valid) {
  __shared__ float red[BLOCK / 32];
  const int64_t nc = (int64_t)blockIdx.y * C + blockIdx.x;
  const int cnt = counts[nc];
  // calc_centroids.py:129,141: avgpool(feat*mask) / avgpool(mask) == (sum/hw) / (count/hw)
  float part = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) {
    float v = 0.f;
    if (cnt > 0) v = class_mean(sums[nc * D + d], cnt);
    vec[nc * D + d] = v;
    part += v;
  }
  const float tot = block_sum<BLOCK>(part, red);
  if (threadIdx.x == 0) {
    vecsum[nc] = tot;
    valid[nc] = (cnt >= 5) ? 1 : 0;   // :134 (frac == 0) and :136 (< 5 pixels) skips
  }
}

// means fused with the exchange of the image-sharded exact mode: the per-image class vectors are written
// straight into the row block this rank owns in EVERY rank's gathered buffer — plain stores to peer memory over NVLink
// (symmetric allocation: the same layout at every `peers[r]`), fire-and-forget while the accumulation of the next batch
// already runs.  What used to be a separate all-gather of 463 MB at the end of the pass (1.4 ms at 8 ranks) is gone; the pass
// ends with a barrier and the ordered replay.
struct PeerBuffers {
  unsigned char* base[DIGA_MAX_PEERS];
  unsigned char* multicast;                   // NVLS multicast mapping of the same allocation (one store reaches every rank), or null
  int world;
  int64_t off_vec, off_vecsum, off_valid;     // byte offsets of vec [rows,C,D] f32, vecsum [rows,C] f32, valid [rows,C] u8
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_means_scatter_kernel(const float* __restrict__ sums, const int32_t* __restrict__ counts, int64_t C, int64_t D,
                              int64_t hw, int64_t row0, PeerBuffers peers) {
  __shared__ float red[BLOCK / 32];
  const int64_t nc = (int64_t)blockIdx.y * C + blockIdx.x;                 // local (image, class)
  const int64_t gnc = (row0 + blockIdx.y) * C + blockIdx.x;                // its place in the gathered buffers
  const int cnt = counts[nc];
  float part = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) {
    float v = 0.f;
    if (cnt > 0) v = class_mean(sums[nc * D + d], cnt);                    // same expression as centroid_means_kernel
    if (peers.multicast != nullptr) {
      // one store, replicated by the NVSwitch into every rank's copy (the 155 KB of a row leave this GPU once, not `world` times)
      asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(peers.multicast + peers.off_vec) + gnc * D + d),
                   "f"(v)
                   : "memory");
    } else {
      for (int r = 0; r < peers.world; ++r) reinterpret_cast<float*>(peers.base[r] + peers.off_vec)[gnc * D + d] = v;
    }
    part += v;
  }
  const float tot = block_sum<BLOCK>(part, red);
  if (threadIdx.x == 0) {
    for (int r = 0; r < peers.world; ++r) {                                // (two scalars per row: plain peer stores)
      reinterpret_cast<float*>(peers.base[r] + peers.off_vecsum)[gnc] = tot;
      (peers.base[r] + peers.off_valid)[gnc] = (cnt >= 5) ? 1 : 0;
    }
  }
  __threadfence_system();      // the peer / multicast stores are ordered before this kernel's completion at system scope
}

// ------------------------------------------------------------------------------------------------
// update (calc_centroids.py:147-164), arithmetic mirrored op for op (separately rounded mul/add/div)
// ------------------------------------------------------------------------------------------------
constexpr int kUpdateSum = 2;   // internal: the sum-mode accumulator of the multi-GPU pass (not a reference update rule)

struct UpdateRule {
  int mode;          // DIGA_UPDATE_* or kUpdateSum
  int start_mean;
  float m, one_minus_m;
};

__device__ __forceinline__ bool rule_is_mean(const UpdateRule& r, float num) {
  return r.mode == DIGA_UPDATE_MEAN || (r.start_mean && num < 100.f);   // :150
}
__device__ __forceinline__ float rule_apply(const UpdateRule& r, bool mean, float obj, float num, float v) {
  if (mean) {
    const float t = __fadd_rn(__fmul_rn(obj, num), v);                  // :158
    return __fdiv_rn(t, __fadd_rn(num, 1.f));                           // :159-160
  }
  return __fadd_rn(__fmul_rn(obj, r.one_minus_m), __fmul_rn(r.m, v));   // :153-154
}

// Order in which the rows of vec / vecsum / valid are visited.  Identity (world == 1) for a single process.  For the
// image-sharded pass (exact mode) the buffer is the all-gather of every rank's rows, [world][per_shard], and
// rank r holds the loader batches r, r + world, ... (`group` images each): the g-th image of the GLOBAL sequence
// (calc_centroids.py:67-78 visits images in loader order) sits at row (k % world) * per_shard + (k / world) * group + j
// with k = g / group, j = g % group.
struct ImageOrder {
  int64_t group, world, per_shard;
};
__device__ __forceinline__ int64_t order_row(const ImageOrder& o, int64_t g) {
  if (o.world == 1) return g;
  const int64_t k = g / o.group, j = g - k * o.group;
  return (k % o.world) * o.per_shard + (k / o.world) * o.group + j;
}

// grid (C, ceil(D / BLOCK)): one thread per (class, channel).  The sequential recurrence over images runs in
// registers (the centroid element is read once and written once); every block of a class replays the same scalar
// `num` recurrence.  (A first version with one block per class and global read-modify-writes
// per image took 37 us for N=8, D=2048 — a third of the accumulation kernel it follows.)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_kernel(const float* __restrict__ vec, const float* __restrict__ vecsum, const uint8_t* __restrict__ valid,
                       int64_t n, int64_t C, int64_t D, float* __restrict__ obj, const float* __restrict__ objnum, UpdateRule rule) {
  const int64_t c = blockIdx.x;
  const int64_t d = (int64_t)blockIdx.y * BLOCK + threadIdx.x;
  const bool live = d < D;
  float num = objnum[c];
  float o = live ? obj[c * D + d] : 0.f;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t nc = i * C + c;
    if (valid != nullptr && !valid[nc]) continue;
    if (vecsum[nc] == 0.f) continue;                                    // :148
    const bool mean = rule_is_mean(rule, num);
    const float v = live ? vec[nc * D + d] : 0.f;
    o = rule_apply(rule, mean, o, num, v);
    num = fminf(__fadd_rn(num, 1.f), 3000.f);                           // :155-156 / :159,161
  }
  if (live) obj[c * D + d] = o;
}

// Long sequences (a whole pass over the target set: thousands of images, most of them without a valid vector for a
// given class): the block first compacts, in order, the rows that update class c into a shared-memory list, then walks
// the list eight rows at a time with the eight loads issued before the eight dependent updates.  The loop above would
// pay one exposed DRAM latency per image (2975 x 19 x D/256 blocks: milliseconds).
template <int BLOCK, int CHUNK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_long_kernel(const float* __restrict__ vec, const float* __restrict__ vecsum, const uint8_t* __restrict__ valid,
                            int64_t n, int64_t C, int64_t D, float* __restrict__ obj, const float* __restrict__ objnum,
                            UpdateRule rule, ImageOrder order) {
  static_assert(CHUNK % BLOCK == 0, "CHUNK must be a multiple of BLOCK");
  constexpr int PER = CHUNK / BLOCK;
  __shared__ int32_t list[CHUNK];
  __shared__ int warp_tot[BLOCK / 32];
  const int64_t c = blockIdx.x;
  const int64_t d = (int64_t)blockIdx.y * BLOCK + threadIdx.x;
  const bool live = d < D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float num = objnum[c];
  float o = live ? obj[c * D + d] : 0.f;
  for (int64_t base = 0; base < n; base += CHUNK) {
    // ordered compaction: thread t owns the PER consecutive sequence positions base + t*PER ..
    int32_t rows[PER];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int64_t g = base + (int64_t)threadIdx.x * PER + u;
      rows[u] = -1;
      if (g < n) {
        const int64_t r = order_row(order, g);
        const int64_t nc = r * C + c;
        if ((valid == nullptr || valid[nc]) && vecsum[nc] != 0.f) {      // :134/:136 skips and :148
          rows[u] = (int32_t)r;
          ++mine;
        }
      }
    }
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = incl - mine, total = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) {
      if (w < warp) before += warp_tot[w];
      total += warp_tot[w];
    }
#pragma unroll
    for (int u = 0; u < PER; ++u)
      if (rows[u] >= 0) list[before++] = rows[u];
    __syncthreads();
    for (int i = 0; i < total; i += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (live && i + u < total) ? __ldg(vec + ((int64_t)list[i + u] * C + c) * D + d) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (i + u < total) {
          const bool mean = rule_is_mean(rule, num);
          o = rule_apply(rule, mean, o, num, v[u]);
          num = fminf(__fadd_rn(num, 1.f), 3000.f);
        }
      }
    }
    __syncthreads();          // the list is rebuilt by the next chunk
  }
  if (live) obj[c * D + d] = o;
}

// The counts are advanced by a second, tiny launch so that no block of the kernels above can observe a count that
// another block of the same class has already updated.  One block per class: parallel count of the updating rows,
// then the `min(num + 1, 3000)` recurrence (:155-156) replayed by one thread, stopping at the clamp.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_num_kernel(const float* __restrict__ vecsum, const uint8_t* __restrict__ valid, int64_t n, int64_t C,
                           float* __restrict__ objnum) {
  __shared__ float red[BLOCK / 32];
  const int64_t c = blockIdx.x;
  float cnt = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += BLOCK) {
    const int64_t nc = i * C + c;
    if ((valid == nullptr || valid[nc]) && vecsum[nc] != 0.f) cnt += 1.f;
  }
  const float total = block_sum<BLOCK>(cnt, red);      // integer-valued, exact below 2^24 rows
  if (threadIdx.x == 0) {
    float num = objnum[c];
    for (int64_t k = 0; k < (int64_t)total; ++k) {
      const float nx = fminf(__fadd_rn(num, 1.f), 3000.f);
      if (nx == num) break;                            // at the clamp (or beyond fp32's integer range): a fixed point
      num = nx;
    }
    objnum[c] = num;
  }
}

// means + update + count in ONE launch for the online path (a batch of <= kFinishMaxRows image rows per thread): one
// thread-block CLUSTER per class, CTA y of the cluster owns channels [y*BLOCK, (y+1)*BLOCK) (+ k * Y*BLOCK).  The thread
// turns the class sums of its channel into the per-image mean vectors (calc_centroids.py:129,141, same expression as
// centroid_means_kernel), the CTAs exchange their partial channel sums through distributed shared memory to obtain
// vector.sum() of every image (:148) in a fixed order, and the recurrence of :150-161 runs in registers.  Exactly one
// cluster touches objnum[c]: every CTA reads it before the first cluster barrier, rank 0 writes it after the last.
constexpr int kFinishMaxRows = 16;

template <int BLOCK, int K, int NR>
__global__ void __launch_bounds__(BLOCK)
centroid_finish_kernel(const float* __restrict__ sums, const int32_t* __restrict__ counts, int n, int64_t C, int64_t D,
                       int64_t hw, float* __restrict__ vec, float* __restrict__ vecsum, uint8_t* __restrict__ valid,
                       float* __restrict__ obj, float* __restrict__ objnum, UpdateRule rule, int do_update, int64_t obj_stride,
                       int64_t num_stride) {
  // obj_stride / num_stride: row pitch of the centroid matrix and element pitch of the counts — D and 1 for the state of
  // Class_Features, D+1 and D+1 when the target is the sum-mode accumulator acc[C, D+1] (count in column D, rule.mode = SUM)
  static_assert(NR * K <= kFinishMaxRows, "register budget");   // NR: image rows a thread holds (>= n), K: channels per row
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float warp_part[NR][BLOCK / 32];
  __shared__ float cta_part[NR];
  __shared__ float tot[NR];
  const int64_t c = blockIdx.x;
  const int Y = gridDim.y;                                   // == cluster size along y
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();                                                // (chain: the accumulation kernel's sums must have landed)
  float num = do_update ? objnum[c * num_stride] : 0.f;
  float v[NR][K];                                            // image i, channel dbase + k*Y*BLOCK
  const int64_t dbase = (int64_t)blockIdx.y * BLOCK + threadIdx.x;
  // all loads first (counts, then the sums unconditionally), arithmetic afterwards: the divisions below carry a slow-path
  // branch each, and ptxas does not move loads across them — interleaved, every row paid two exposed L2 latencies
  // (ncu: 14.8 us for 8 rows; now one latency for the whole batch)
  int cnt[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) cnt[i] = i < n ? __ldg(counts + i * C + c) : 0;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int64_t d = dbase + (int64_t)k * Y * BLOCK;
      v[i][k] = (i < n && d < D) ? __ldg(sums + (i * C + c) * D + d) : 0.f;      // (garbage where cnt == 0: masked below)
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int64_t d = dbase + (int64_t)k * Y * BLOCK;
      v[i][k] = (i < n && d < D && cnt[i] > 0) ? class_mean(v[i][k], cnt[i]) : 0.f;
      if (i < n && d < D && vec != nullptr) vec[(i * C + c) * D + d] = v[i][k];
    }
  }
  // vector.sum() per image: warp -> CTA -> cluster, every level in a fixed order
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (i < n) {                                             // CTA-uniform
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) p += v[i][k];
      p = warp_sum(p);
      if (lane == 0) warp_part[i][warp] = p;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    float p = 0.f;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) p += warp_part[threadIdx.x][w];
    cta_part[threadIdx.x] = p;
  }
  cluster.sync();
  if ((int)threadIdx.x < n) {
    float part[8];                                           // cluster size <= 8: all remote loads in flight together
#pragma unroll
    for (int r = 0; r < 8; ++r) part[r] = r < Y ? *cluster.map_shared_rank(&cta_part[threadIdx.x], r) : 0.f;
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r];                // fixed order
    tot[threadIdx.x] = t;
    if (blockIdx.y == 0) {
      if (vecsum != nullptr) vecsum[threadIdx.x * C + c] = t;
      if (valid != nullptr) valid[threadIdx.x * C + c] = __ldg(counts + threadIdx.x * C + c) >= 5 ? 1 : 0;   // :134, :136
    }
  }
  cluster.sync();                                            // remote reads done before any CTA may exit; tot[] visible
  if (!do_update) return;
  float o[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int64_t d = dbase + (int64_t)k * Y * BLOCK;
    o[k] = d < D ? obj[c * obj_stride + d] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (i < n && cnt[i] >= 5 && tot[i] != 0.f) {                          // :134/:136 skips, :148
      if (rule.mode == kUpdateSum) {                                      // multi-GPU sum mode: acc += (vector, 1)
#pragma unroll
        for (int k = 0; k < K; ++k) o[k] += v[i][k];
        num += 1.f;
      } else {
        const bool mean = rule_is_mean(rule, num);
#pragma unroll
        for (int k = 0; k < K; ++k) o[k] = rule_apply(rule, mean, o[k], num, v[i][k]);
        num = fminf(__fadd_rn(num, 1.f), 3000.f);                         // :155-156 / :159,161
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int64_t d = dbase + (int64_t)k * Y * BLOCK;
    if (d < D) obj[c * obj_stride + d] = o[k];
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) objnum[c * num_stride] = num;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_update_single_kernel(const float* __restrict__ v, int64_t id, int64_t D, float* __restrict__ obj,
                              float* __restrict__ objnum, UpdateRule rule) {
  __shared__ float red[BLOCK / 32];
  __shared__ float s_tot;
  float part = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) part += v[d];
  const float tot = block_sum<BLOCK>(part, red);
  if (threadIdx.x == 0) s_tot = tot;
  __syncthreads();
  if (s_tot == 0.f) return;
  const float num = objnum[id];
  const bool mean = rule_is_mean(rule, num);
  float* o = obj + id * D;
  for (int64_t d = threadIdx.x; d < D; d += BLOCK) o[d] = rule_apply(rule, mean, o[d], num, v[d]);
  __syncthreads();
  if (threadIdx.x == 0) objnum[id] = fminf(__fadd_rn(num, 1.f), 3000.f);
}

// acc[c][0..D) += sum over valid images of vec[n][c][:]; acc[c][D] += number of such images.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
centroid_reduce_images_kernel(const float* __restrict__ vec, const float* __restrict__ vecsum,
                              const uint8_t* __restrict__ valid, int64_t n, int64_t C, int64_t D, float* __restrict__ acc) {
  const int64_t c = blockIdx.x;
  const int64_t d = (int64_t)blockIdx.y * BLOCK + threadIdx.x;   // d == D is the image-count column
  if (d > D) return;
  float a = acc[c * (D + 1) + d];
  for (int64_t i = 0; i < n; ++i) {
    const int64_t nc = i * C + c;
    if (valid != nullptr && !valid[nc]) continue;
    if (vecsum[nc] == 0.f) continue;
    a += d < D ? vec[nc * D + d] : 1.f;
  }
  acc[c * (D + 1) + d] = a;
}

// process_label (util/utils.py:158-163): onehot[b][k][p] = (k == (label < C ? long(label) : C)), k in [0, C].
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
onehot_labels_kernel(const float* __restrict__ label, int64_t C, int64_t hw, float* __restrict__ onehot) {
  const int64_t b = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * BLOCK + threadIdx.x; p < hw; p += (int64_t)gridDim.x * BLOCK) {
    const float lf = __ldg(label + b * hw + p);
    const long long id = lf < (float)C ? (long long)lf : (long long)C;
    float* o = onehot + b * (C + 1) * hw + p;
    for (int64_t k = 0; k <= C; ++k) o[k * hw] = (k == id) ? 1.f : 0.f;
  }
}

static UpdateRule make_rule(int mode, int start_mean, double momentum) {
  UpdateRule r;
  r.mode = mode;
  r.start_mean = start_mean;
  r.m = (float)momentum;
  r.one_minus_m = (float)(1.0 - momentum);   // Python evaluates (1 - momentum) in double, then torch rounds to fp32
  return r;
}

}  // namespace diga

extern "C" {

int64_t diga_centroid_clsw_bytes(int64_t n, int64_t hw) {
  if (n < 0 || hw < 0) return 0;
  return n * 4 * diga::clsw_words(hw) * (int64_t)sizeof(uint32_t);
}

int diga_centroid_clsw_build(const uint8_t* cls, int64_t n, int64_t C, int64_t hw, uint32_t* clsw, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(cls && clsw, DIGA_ERR_INVALID, "centroid_clsw_build: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && n >= 0 && n <= 65535 && hw >= 0, DIGA_ERR_INVALID, "centroid_clsw_build: bad sizes");
  DIGA_REQUIRE(aligned(clsw, 4), DIGA_ERR_MISALIGNED, "centroid_clsw_build: misaligned pointer");
  if (n == 0) return DIGA_OK;
  const int64_t wp = clsw_words(hw);
  centroid_clsw_build_kernel<<<dim3((unsigned)((wp + 255) / 256), 4, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(cls, (int)C, hw, clsw);
  DIGA_CHECK_LAUNCH("centroid_clsw_build_kernel");
  return DIGA_OK;
}

int diga_centroid_assign(const float* logits, const float* labels, int64_t n, int64_t C, int64_t hw, uint8_t* cls,
                         int32_t* counts, uint32_t* clsw, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(logits && cls && counts, DIGA_ERR_INVALID, "centroid_assign: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "centroid_assign: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && n <= 65535 && hw >= 0, DIGA_ERR_INVALID, "centroid_assign: bad sizes");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(labels, 4) && aligned(counts, 4) && aligned(clsw, 4), DIGA_ERR_MISALIGNED,
               "centroid_assign: misaligned pointer");
  if (n == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts, 0, (size_t)n * C * sizeof(int32_t), st);
  if (hw == 0) return DIGA_OK;
  constexpr int BLOCK = 256;
  int64_t gx = (hw + BLOCK - 1) / BLOCK;
  const int64_t cap = ((int64_t)sm_count() * 8 + n - 1) / n;
  if (gx > cap) gx = cap;
  DIGA_DISPATCH_C(C, {
    centroid_assign_kernel<BLOCK, kC, kPad><<<dim3((unsigned)gx, (unsigned)n), BLOCK, 0, st>>>(logits, labels, (int)C, hw, cls, counts, clsw);
  });
  DIGA_CHECK_LAUNCH("centroid_assign_kernel");
  return DIGA_OK;
}

int diga_centroid_assign_fullres(const float* logits, const int64_t* labels_full, int64_t n, int64_t C, int64_t h, int64_t w,
                                 int64_t H, int64_t W, uint8_t* cls, int32_t* counts, uint32_t* clsw, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(logits && labels_full && cls && counts, DIGA_ERR_INVALID, "centroid_assign_fullres: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "centroid_assign_fullres: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && n <= 65535 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && h < (1 << 24) && w < (1 << 24) && H < (1 << 24) &&
                   W < (1 << 24),
               DIGA_ERR_INVALID, "centroid_assign_fullres: bad sizes");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(labels_full, 8) && aligned(counts, 4) && aligned(clsw, 4), DIGA_ERR_MISALIGNED,
               "centroid_assign_fullres: misaligned pointer");
  if (n == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t hw = h * w;
  cudaMemsetAsync(counts, 0, (size_t)n * C * sizeof(int32_t), st);
  constexpr int BLOCK = 256;
  int64_t gx = (hw + BLOCK - 1) / BLOCK;
  const int64_t cap = ((int64_t)sm_count() * 8 + n - 1) / n;
  if (gx > cap) gx = cap;
  DIGA_DISPATCH_C(C, {
    centroid_assign_kernel<BLOCK, kC, kPad><<<dim3((unsigned)gx, (unsigned)n), BLOCK, 0, st>>>(
        logits, nullptr, (int)C, hw, cls, counts, clsw, labels_full, (int)w, (int)H, (int)W, (float)H / (float)h, (float)W / (float)w);
  });
  DIGA_CHECK_LAUNCH("centroid_assign_kernel");
  return DIGA_OK;
}

int diga_centroid_accum(const float* feat, const uint8_t* cls, const int32_t* counts, const uint32_t* clsw, int64_t n, int64_t D,
                        int64_t C, int64_t hw, float* sums, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(feat && cls && sums, DIGA_ERR_INVALID, "centroid_accum: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "centroid_accum: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 0 && D >= 0 && hw >= 0, DIGA_ERR_INVALID, "centroid_accum: bad sizes");
  DIGA_REQUIRE(aligned(feat, 4) && aligned(sums, 4) && aligned(counts, 4) && aligned(clsw, 4), DIGA_ERR_MISALIGNED,
               "centroid_accum: misaligned pointer");
  if (n == 0 || D == 0) return DIGA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int variant = tunable("accum_variant", 0);
  const int RS = quad_row_stride(hw, D);
  const bool quad_ok = RS > 0 && aligned(feat, 128) && hw >= 4;
  // Default: the lean 128-bit kernel when the caller supplies the phase-shifted class words (diga_centroid_assign /
  // diga_centroid_clsw_build), else the round-1 128-bit kernel on the plain byte map; the scalar kernels take the shapes
  // neither can tile (variants 1..7 force them, 8..12 the round-1 128-bit shapes, 13.. the lean shapes).
  if (quad_ok && clsw != nullptr && (variant == 0 || variant >= 13)) {
    const int nc = (int)C;
    switch (variant) {
      case 13: return launch_accum_lean<4, 4, 4, 1, true>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);   // register ring
      case 14: return launch_accum_lean<4, 4, 4, 1>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);         // splat fast path
      case 15: return launch_accum_lean<4, 3, 5, 0>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);
      case 16: return launch_accum_lean<4, 4, 4, 2>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);         // merged RMWs
      case 17: return launch_accum_lean<4, 4, 4, 0>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);         // dummy slot
      case 18: return launch_accum_lean<2, 4, 8, 0>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);
      case 19: return launch_accum_lean<8, 4, 2, 0>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);
      case 20: return launch_accum_lean<3, 4, 5, 3>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);
      default: return launch_accum_lean<4, 4, 4, 3>(feat, clsw, cls, counts, nc, n, D, hw, sums, st);
    }
  }
  if (quad_ok && (variant == 0 || variant >= 8)) {
    switch (variant) {
      case 8: return launch_accum_quad<4, 1>(feat, cls, (int)C, n, D, hw, sums, st);
      case 10: return launch_accum_quad<2, 1>(feat, cls, (int)C, n, D, hw, sums, st);
      case 11: return launch_accum_quad<2, 2>(feat, cls, (int)C, n, D, hw, sums, st);
      case 12: return launch_accum_quad<8, 1>(feat, cls, (int)C, n, D, hw, sums, st);
      default: return launch_accum_quad<4, 2>(feat, cls, (int)C, n, D, hw, sums, st);
    }
  }
  switch (variant) {
    case 1: return launch_accum<4, 4, 4, false>(feat, cls, (int)C, n, D, hw, sums, st);
    case 2: return launch_accum<4, 4, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 3: return launch_accum<4, 8, 4, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 4: return launch_accum<2, 4, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 5: return launch_accum<4, 2, 4, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 6: return launch_accum<4, 2, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    case 7: return launch_accum<2, 8, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
    default: return launch_accum<4, 4, 8, true>(feat, cls, (int)C, n, D, hw, sums, st);
  }
}

int diga_centroid_means(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw, float* vec,
                        float* vecsum, uint8_t* valid, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(sums && counts && vec && vecsum && valid, DIGA_ERR_INVALID, "centroid_means: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && n >= 0 && n <= 65535 && D >= 0 && hw > 0, DIGA_ERR_INVALID,
               "centroid_means: bad sizes");
  if (n == 0) return DIGA_OK;
  centroid_means_kernel<256><<<dim3((unsigned)C, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(sums, counts, C, D, hw, vec,
                                                                                                vecsum, valid);
  DIGA_CHECK_LAUNCH("centroid_means_kernel");
  return DIGA_OK;
}

static int launch_centroid_update(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C, int64_t D,
                                  float* objective_vectors, float* objective_num, int mode, int start_mean, double momentum,
                                  diga::ImageOrder order, cudaStream_t st) {
  using namespace diga;
  if (n == 0) return DIGA_OK;
  if (D > 0) {
    const dim3 grid((unsigned)C, (unsigned)((D + 255) / 256));
    if (n <= 32 && order.world == 1) {
      centroid_update_kernel<256><<<grid, 256, 0, st>>>(vec, vecsum, valid, n, C, D, objective_vectors, objective_num,
                                                        make_rule(mode, start_mean, momentum));
      DIGA_CHECK_LAUNCH("centroid_update_kernel");
    } else {
      centroid_update_long_kernel<256, 1024><<<grid, 256, 0, st>>>(vec, vecsum, valid, n, C, D, objective_vectors, objective_num,
                                                                   make_rule(mode, start_mean, momentum), order);
      DIGA_CHECK_LAUNCH("centroid_update_long_kernel");
    }
  }
  // rows outside the visited sequence (padding of a sharded buffer) must not count: the sharded caller zeroes their vecsum
  const int64_t rows = order.world == 1 ? n : order.world * order.per_shard;
  centroid_update_num_kernel<256><<<(unsigned)C, 256, 0, st>>>(vecsum, valid, rows, C, objective_num);
  DIGA_CHECK_LAUNCH("centroid_update_num_kernel");
  return DIGA_OK;
}

int diga_centroid_means_scatter(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw,
                                void* const* peer_bases, void* multicast_base, int64_t world, int64_t off_vec, int64_t off_vecsum,
                                int64_t off_valid, int64_t row0, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(sums && counts && peer_bases, DIGA_ERR_INVALID, "centroid_means_scatter: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && n >= 0 && n <= 65535 && D >= 0 && hw > 0 && row0 >= 0, DIGA_ERR_INVALID,
               "centroid_means_scatter: bad sizes");
  DIGA_REQUIRE(world >= 1 && world <= DIGA_MAX_PEERS, DIGA_ERR_INVALID, "centroid_means_scatter: world=%lld outside [1,%d]",
               (long long)world, DIGA_MAX_PEERS);
  DIGA_REQUIRE(off_vec % 16 == 0 && off_vecsum % 4 == 0 && off_vec >= 0 && off_vecsum >= 0 && off_valid >= 0, DIGA_ERR_MISALIGNED,
               "centroid_means_scatter: misaligned offsets");
  if (n == 0) return DIGA_OK;
  PeerBuffers pb;
  pb.multicast = tunable("scatter_multicast", 1) ? static_cast<unsigned char*>(multicast_base) : nullptr;
  pb.world = (int)world;
  pb.off_vec = off_vec;
  pb.off_vecsum = off_vecsum;
  pb.off_valid = off_valid;
  for (int r = 0; r < DIGA_MAX_PEERS; ++r) pb.base[r] = r < world ? static_cast<unsigned char*>(peer_bases[r]) : nullptr;
  for (int r = 0; r < world; ++r) DIGA_REQUIRE(pb.base[r] != nullptr, DIGA_ERR_INVALID, "centroid_means_scatter: null peer buffer %d", r);
  centroid_means_scatter_kernel<256><<<dim3((unsigned)C, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(sums, counts, C, D, hw, row0, pb);
  DIGA_CHECK_LAUNCH("centroid_means_scatter_kernel");
  return DIGA_OK;
}

int diga_centroid_update(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C, int64_t D,
                         float* objective_vectors, float* objective_num, int mode, int start_mean, double momentum,
                         diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vec && vecsum && objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_update: null pointer");
  DIGA_REQUIRE(mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  DIGA_REQUIRE(C >= 1 && C <= 65535 && n >= 0 && n < (1ll << 31) && D >= 0, DIGA_ERR_INVALID, "centroid_update: bad sizes");
  return launch_centroid_update(vec, vecsum, valid, n, C, D, objective_vectors, objective_num, mode, start_mean, momentum,
                                ImageOrder{1, 1, n}, (cudaStream_t)stream);
}

int diga_centroid_update_sharded(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n_total, int64_t group,
                                 int64_t world, int64_t per_shard, int64_t C, int64_t D, float* objective_vectors,
                                 float* objective_num, int mode, int start_mean, double momentum, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vec && vecsum && objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_update_sharded: null pointer");
  DIGA_REQUIRE(mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  DIGA_REQUIRE(C >= 1 && C <= 65535 && D >= 0 && n_total >= 0 && group >= 1 && world >= 1 && per_shard >= 0 &&
                   world * per_shard < (1ll << 31),
               DIGA_ERR_INVALID, "centroid_update_sharded: bad sizes");
  // every image of the global sequence must have a row: rank r holds ceil((K - r) / world) batches, K = ceil(n_total / group)
  const int64_t batches = (n_total + group - 1) / group;
  DIGA_REQUIRE(((batches + world - 1) / world) * group <= per_shard || n_total == 0, DIGA_ERR_INVALID,
               "centroid_update_sharded: per_shard=%lld rows cannot hold %lld batches of %lld images over %lld ranks",
               (long long)per_shard, (long long)batches, (long long)group, (long long)world);
  return launch_centroid_update(vec, vecsum, valid, n_total, C, D, objective_vectors, objective_num, mode, start_mean, momentum,
                                ImageOrder{group, world, per_shard}, (cudaStream_t)stream);
}

int diga_centroid_finish_supported(int64_t n, int64_t D) {
  if (n < 1 || D < 1) return 0;
  const int64_t Y = (D + 255) / 256 < 8 ? (D + 255) / 256 : 8;
  int64_t K = (D + Y * 256 - 1) / (Y * 256), Kp = 1, Np = 1;
  while (Kp < K) Kp *= 2;                                  // kernel instantiations: K in {1, 2, 4}, rows in {1, 2, 4, 8, 16}
  while (Np < n) Np *= 2;
  return (Kp <= 4 && Np * Kp <= diga::kFinishMaxRows) ? 1 : 0;
}

static int centroid_finish_impl(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw, float* vec,
                                float* vecsum, uint8_t* valid, float* objective_vectors, float* objective_num, int mode,
                                int start_mean, double momentum, int64_t obj_stride, int64_t num_stride, diga_stream_t stream);

int diga_centroid_finish(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw, float* vec,
                         float* vecsum, uint8_t* valid, float* objective_vectors, float* objective_num, int mode, int start_mean,
                         double momentum, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(objective_vectors == nullptr || mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  return centroid_finish_impl(sums, counts, n, C, D, hw, vec, vecsum, valid, objective_vectors, objective_num, mode, start_mean,
                              momentum, D, 1, stream);
}

static int centroid_finish_impl(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw, float* vec,
                                float* vecsum, uint8_t* valid, float* objective_vectors, float* objective_num, int mode,
                                int start_mean, double momentum, int64_t obj_stride, int64_t num_stride, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(sums && counts, DIGA_ERR_INVALID, "centroid_finish: null pointer");
  const int do_update = objective_vectors != nullptr;
  DIGA_REQUIRE(!do_update || objective_num != nullptr, DIGA_ERR_INVALID, "centroid_finish: objective_num missing");
  DIGA_REQUIRE(C >= 1 && C <= 65535 && hw > 0, DIGA_ERR_INVALID, "centroid_finish: bad sizes");
  DIGA_REQUIRE(diga_centroid_finish_supported(n, D), DIGA_ERR_INVALID,
               "centroid_finish: n=%lld rows of D=%lld exceed the register budget (use centroid_means + centroid_update)",
               (long long)n, (long long)D);
  const int64_t Y = (D + 255) / 256 < 8 ? (D + 255) / 256 : 8;
  const int64_t K = (D + Y * 256 - 1) / (Y * 256);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C, (unsigned)Y, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)Y;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_chain_pdl ? 2 : 1;
  const UpdateRule rule = make_rule(do_update ? mode : DIGA_UPDATE_MEAN, start_mean, momentum);
  cudaError_t e = cudaSuccess;
  int np = 1;
  while (np < n) np *= 2;
#define DIGA_FINISH_KN(KK, NN)                                                                                                \
  e = cudaLaunchKernelEx(&cfg, centroid_finish_kernel<256, KK, NN>, sums, counts, (int)n, C, D, hw, vec, vecsum, valid,      \
                         objective_vectors, objective_num, rule, do_update, obj_stride, num_stride)
#define DIGA_FINISH_K(KK)                                                                                                     \
  do {                                                                                                                        \
    if (np == 1) DIGA_FINISH_KN(KK, 1);                                                                                       \
    else if (np == 2) DIGA_FINISH_KN(KK, 2);                                                                                  \
    else if (np == 4) DIGA_FINISH_KN(KK, 4);                                                                                  \
    else if (np == 8) { if constexpr (KK <= 2) DIGA_FINISH_KN(KK, (KK <= 2 ? 8 : 1)); }                                       \
    else { if constexpr (KK <= 1) DIGA_FINISH_KN(KK, (KK <= 1 ? 16 : 1)); }                                                   \
  } while (0)
  if (K <= 1) DIGA_FINISH_K(1);
  else if (K <= 2) DIGA_FINISH_K(2);
  else DIGA_FINISH_K(4);
#undef DIGA_FINISH_K
#undef DIGA_FINISH_KN
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("centroid_finish_kernel: launch failed: %s", cudaGetErrorString(e));
    return DIGA_ERR_CUDA;
  }
  DIGA_CHECK_LAUNCH("centroid_finish_kernel");
  return DIGA_OK;
}

// ---- a6 -> a7 as ONE call: assign -> accum -> finish (or means + update for batches finish cannot hold) on `stream`.
// The per-image call sequence of the reference's loops (calc_centroids.py:67-78: one image per call) is host-bound when every
// kernel is its own FFI call with its own scratch tensors; this entry point takes one workspace and queues the whole chain.
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int64_t diga_centroid_chain_workspace_bytes(int64_t n, int64_t C, int64_t D, int64_t hw) {
  if (n < 0 || C < 1 || D < 0 || hw < 0) return 0;
  size_t b = align_up((size_t)n * C * sizeof(int32_t), 256);                       // counts
  b += align_up((size_t)diga_centroid_clsw_bytes(n, hw), 256);                     // phase-shifted class words
  b += align_up((size_t)n * hw, 256);                                              // class map
  b += align_up((size_t)n * C * D * sizeof(float), 256);                           // class sums
  b += align_up((size_t)n * C * D * sizeof(float), 256);                           // vec     (paths that do not end in `finish`)
  b += align_up((size_t)n * C * sizeof(float), 256) + align_up((size_t)n * C, 256);   // vecsum, valid
  return (int64_t)b;
}

namespace {
struct ChainBufs {
  int32_t* counts;
  uint32_t* clsw;
  uint8_t* cls;
  float* sums;
  float* vec;
  float* vecsum;
  uint8_t* valid;
};

ChainBufs chain_layout(void* workspace, int64_t n, int64_t C, int64_t D, int64_t hw) {
  ChainBufs b;
  unsigned char* p = static_cast<unsigned char*>(workspace);
  b.counts = reinterpret_cast<int32_t*>(p);
  p += align_up((size_t)n * C * sizeof(int32_t), 256);
  b.clsw = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)diga_centroid_clsw_bytes(n, hw), 256);
  b.cls = p;
  p += align_up((size_t)n * hw, 256);
  b.sums = reinterpret_cast<float*>(p);
  p += align_up((size_t)n * C * D * sizeof(float), 256);
  b.vec = reinterpret_cast<float*>(p);
  p += align_up((size_t)n * C * D * sizeof(float), 256);
  b.vecsum = reinterpret_cast<float*>(p);
  p += align_up((size_t)n * C * sizeof(float), 256);
  b.valid = p;
  return b;
}

struct PdlScope {
  explicit PdlScope(bool on) { diga::g_chain_pdl = on; }
  ~PdlScope() { diga::g_chain_pdl = false; }
};

// assign -> accum into the workspace.  Short chains (the per-image calls of the reference's loops) launch accum (and the
// finish kernel of diga_centroid_chain) as programmatic dependents of their predecessor: accum fetches its first feature
// batch while assign still runs, finish is resident when the last accumulation CTA leaves — 24.5 -> 22.6 us per
// [1,2048,65,129] call.  Long accumulations gain nothing (0.103 -> 0.105 ms at [8,2048,65,129]: the early feature requests only
// compete with the assign kernel's own loads) and launch plainly.  Tunable chain_pdl: 0 = never, 1 = short chains, 2 = always.
int chain_sums(const float* feat, const float* logits, const float* labels, const int64_t* labels_full, int64_t H, int64_t W,
               int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, const ChainBufs& b, diga_stream_t stream) {
  using namespace diga;
  const int64_t hw = h * w;
  int rc = labels_full ? diga_centroid_assign_fullres(logits, labels_full, n, C, h, w, H, W, b.cls, b.counts, b.clsw, stream)
                       : diga_centroid_assign(logits, labels, n, C, hw, b.cls, b.counts, b.clsw, stream);
  if (rc != DIGA_OK || hw == 0 || D == 0) return rc;
  const int pdl_mode = tunable("chain_pdl", 1);
  g_chain_pdl = pdl_mode == 2 || (pdl_mode == 1 && n * ((D + 3) / 4) <= (int64_t)sm_count() * 8);
  return diga_centroid_accum(feat, b.cls, b.counts, b.clsw, n, D, C, hw, b.sums, stream);
}

int chain_check(const void* feat, const void* logits, const void* labels, const void* labels_full, void* workspace, int64_t n,
                int64_t h, int64_t w, const char* what) {
  using namespace diga;
  DIGA_REQUIRE(feat && logits && workspace, DIGA_ERR_INVALID, "%s: null pointer", what);
  DIGA_REQUIRE(!(labels && labels_full), DIGA_ERR_INVALID, "%s: pass labels or labels_full, not both", what);
  DIGA_REQUIRE(aligned(workspace, 256), DIGA_ERR_MISALIGNED, "%s: workspace must be 256-byte aligned", what);
  DIGA_REQUIRE(n >= 0 && h >= 0 && w >= 0, DIGA_ERR_INVALID, "%s: bad sizes", what);
  return DIGA_OK;
}
}  // namespace

int diga_centroid_chain(const float* feat, const float* logits, const float* labels, const int64_t* labels_full, int64_t H,
                        int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, void* workspace,
                        float* objective_vectors, float* objective_num, int mode, int start_mean, double momentum,
                        diga_stream_t stream) {
  using namespace diga;
  int rc = chain_check(feat, logits, labels, labels_full, workspace, n, h, w, "centroid_chain");
  if (rc != DIGA_OK) return rc;
  DIGA_REQUIRE(objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_chain: null pointer");
  if (n == 0) return DIGA_OK;
  const int64_t hw = h * w;
  const ChainBufs b = chain_layout(workspace, n, C, D, hw);
  PdlScope pdl_scope(false);
  rc = chain_sums(feat, logits, labels, labels_full, H, W, n, C, D, h, w, b, stream);
  if (rc != DIGA_OK || hw == 0 || D == 0) return rc;
  if (diga_centroid_finish_supported(n, D))
    return diga_centroid_finish(b.sums, b.counts, n, C, D, hw, nullptr, nullptr, nullptr, objective_vectors, objective_num, mode,
                                start_mean, momentum, stream);
  g_chain_pdl = false;
  rc = diga_centroid_means(b.sums, b.counts, n, C, D, hw, b.vec, b.vecsum, b.valid, stream);
  if (rc != DIGA_OK) return rc;
  return diga_centroid_update(b.vec, b.vecsum, b.valid, n, C, D, objective_vectors, objective_num, mode, start_mean, momentum, stream);
}

int diga_centroid_chain_sums(const float* feat, const float* logits, const float* labels, const int64_t* labels_full, int64_t H,
                             int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, void* workspace,
                             float** sums_out, int32_t** counts_out, diga_stream_t stream) {
  using namespace diga;
  int rc = chain_check(feat, logits, labels, labels_full, workspace, n, h, w, "centroid_chain_sums");
  if (rc != DIGA_OK) return rc;
  DIGA_REQUIRE(sums_out && counts_out, DIGA_ERR_INVALID, "centroid_chain_sums: null pointer");
  const ChainBufs b = chain_layout(workspace, n, C, D, h * w);
  *sums_out = b.sums;
  *counts_out = b.counts;
  if (n == 0) return DIGA_OK;
  PdlScope pdl_scope(false);
  return chain_sums(feat, logits, labels, labels_full, H, W, n, C, D, h, w, b, stream);
}

int diga_centroid_chain_reduce(const float* feat, const float* logits, const float* labels, const int64_t* labels_full, int64_t H,
                               int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, void* workspace, float* acc,
                               diga_stream_t stream) {
  using namespace diga;
  int rc = chain_check(feat, logits, labels, labels_full, workspace, n, h, w, "centroid_chain_reduce");
  if (rc != DIGA_OK) return rc;
  DIGA_REQUIRE(acc, DIGA_ERR_INVALID, "centroid_chain_reduce: null pointer");
  if (n == 0) return DIGA_OK;
  const int64_t hw = h * w;
  const ChainBufs b = chain_layout(workspace, n, C, D, hw);
  {
    PdlScope pdl_scope(false);
    rc = chain_sums(feat, logits, labels, labels_full, H, W, n, C, D, h, w, b, stream);
  }
  if (rc != DIGA_OK || hw == 0 || D == 0) return rc;
  // means + reduce in ONE cluster launch (the finish kernel with the accumulator as its target: acc[c][:D] += vector,
  // acc[c][D] += 1 for every valid image) — 10.5 us of two small kernels -> one, per batch
  if (diga_centroid_finish_supported(n, D) && tunable("chain_reduce_fused", 1))
    return centroid_finish_impl(b.sums, b.counts, n, C, D, hw, nullptr, nullptr, nullptr, acc, acc + D, kUpdateSum, 0, 0.0, D + 1, D + 1,
                                stream);
  rc = diga_centroid_means(b.sums, b.counts, n, C, D, hw, b.vec, b.vecsum, b.valid, stream);
  if (rc != DIGA_OK) return rc;
  return diga_centroid_reduce_images(b.vec, b.vecsum, b.valid, n, C, D, acc, stream);
}

int diga_centroid_update_single(const float* vector, int64_t id, int64_t C, int64_t D, float* objective_vectors,
                                float* objective_num, int mode, int start_mean, double momentum, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vector && objective_vectors && objective_num, DIGA_ERR_INVALID, "centroid_update_single: null pointer");
  DIGA_REQUIRE(mode == DIGA_UPDATE_MEAN || mode == DIGA_UPDATE_MOVING_AVERAGE, DIGA_ERR_INVALID,
               "no such updating way of objective vectors %d", mode);
  DIGA_REQUIRE(id >= 0 && id < C && D >= 0, DIGA_ERR_INVALID, "centroid_update_single: class id %lld outside [0,%lld)",
               (long long)id, (long long)C);
  centroid_update_single_kernel<256><<<1, 256, 0, (cudaStream_t)stream>>>(vector, id, D, objective_vectors, objective_num,
                                                                          make_rule(mode, start_mean, momentum));
  DIGA_CHECK_LAUNCH("centroid_update_single_kernel");
  return DIGA_OK;
}

int diga_onehot_labels(const float* label, int64_t B, int64_t C, int64_t hw, float* onehot, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(label && onehot, DIGA_ERR_INVALID, "onehot_labels: null pointer");
  DIGA_REQUIRE(B >= 0 && B <= 65535 && C >= 1 && hw >= 0, DIGA_ERR_INVALID, "onehot_labels: bad sizes");
  if (B == 0 || hw == 0) return DIGA_OK;
  int64_t gx = (hw + 255) / 256;
  const int64_t cap = ((int64_t)sm_count() * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  onehot_labels_kernel<256><<<dim3((unsigned)gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(label, C, hw, onehot);
  DIGA_CHECK_LAUNCH("onehot_labels_kernel");
  return DIGA_OK;
}

int diga_centroid_reduce_images(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C, int64_t D,
                                float* acc, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(vec && vecsum && acc, DIGA_ERR_INVALID, "centroid_reduce_images: null pointer");
  DIGA_REQUIRE(C >= 1 && n >= 0 && D >= 0, DIGA_ERR_INVALID, "centroid_reduce_images: bad sizes");
  if (n == 0) return DIGA_OK;
  centroid_reduce_images_kernel<256><<<dim3((unsigned)C, (unsigned)((D + 1 + 255) / 256)), 256, 0, (cudaStream_t)stream>>>(
      vec, vecsum, valid, n, C, D, acc);
  DIGA_CHECK_LAUNCH("centroid_reduce_images_kernel");
  return DIGA_OK;
}

}  // extern "C"
