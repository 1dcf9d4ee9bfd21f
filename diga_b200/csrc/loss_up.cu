// f1 for the loss consumers: symmetric KD loss and cross_entropy2d evaluated straight
// from the stride-8 logits, the bilinear `align_corners=True` up-sampling (train_DiGA_gta2city_self_training.py
// :289,:344,:348,:351) fused in front of the loss (util/loss.py:125-143, :48-62) and its transpose fused behind the
// gradient.  The reference materialises three [N,19,H,W] tensors per loss (up-sampled logits, their gradient, the
// saved softmax): 76 B/px each, written and re-read.  Here the HBM traffic is the low-resolution logits (~1.2 B/px),
// the int64 targets (8 B/px, CE only) and the low-resolution gradient: the kernels are ALU-bound.
//
// Forward (per output pixel): a thread walks down `ry` output rows of one column, keeping the horizontally
// interpolated source rows of every class in registers (ColumnInterp, bit-identical to ATen's CUDA up-sampler), and
// evaluates the same per-pixel expressions as csrc/kd.cu and csrc/ce.cu.  One kernel serves three call shapes:
//   KD only (teacher + student, 2B images), CE only (logits + targets), and CE on the first n_ce images of the
//   student batch + KD on all of them (self_training.py:349 + :352 share s_pred_cat_stu).
//
// Backward = transpose of the interpolation, without atomics and bitwise deterministic:
//   dlow[y,x] = sum_{Y,X} wy(Y,y) wx(X,x) g[Y,X],  g = per-pixel logit gradient (never materialised).
//   vertical   : the thread accumulates l0(Y)*g into `Gt` (source row i0) and l1(Y)*g into `Gb` (row i1) while it
//                walks down; when the walk crosses into the next source cell (a CTA-uniform event, all threads share
//                the rows) row i0 is complete for this strip and is flushed;
//   horizontal : a flush stages the 128 column values per class in shared memory, and the CTA reduces the runs of
//                columns that share a source column (i0 is monotone in X, so a run is a contiguous range);
//   the CTA writes its partial low-resolution patch [R rows][K cols][C padded to 4] to a scratch slot of its own;
//   gather     : a second, tiny kernel adds the <= 2x2 patches that overlap each low-resolution element, in fixed order.
#include <type_traits>

#include "lerp_column.cuh"

namespace diga {

int tunable(const char* name, int dflt);

constexpr int kLuBlock = 128;                       // threads per CTA = output columns per CTA
constexpr int kLuCols = kLuBlock;
constexpr int kLuLd = kLuCols + kLuCols / 8 + 2;    // rows of the swizzled staging arrays

struct LossUpPlan {
  int ry, SX, SY, R, K;
  int64_t ctas;
  size_t off_partial, off_scratch, bytes;
};
constexpr size_t kLuTileMax = 64 * 1024;   // largest source tile a CTA stages in shared memory (else it reads global memory)

// The tap arithmetic of bilinear_tap() on the host (IEEE single multiply + truncation: identical results).
static inline void host_tap(float scale, int dst, int in, int* i0, int* i1) {
  const float src = scale * (float)dst;
  *i0 = (int)src;
  *i1 = *i0 + ((*i0 < in - 1) ? 1 : 0);
}

static LossUpPlan make_plan(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W) {
  LossUpPlan p;
  p.ry = tunable("lossup_ry", 32);
  if (p.ry < 1) p.ry = 1;
  p.SX = (int)((W + kLuCols - 1) / kLuCols);
  p.SY = (int)((H + p.ry - 1) / p.ry);
  const float sh = bilinear_scale_host(h, H), sw = bilinear_scale_host(w, W);
  p.R = 1;
  for (int ky = 0; ky < p.SY; ++ky) {
    int lo, hi, t;
    const int ye = (int)((int64_t)(ky + 1) * p.ry < H ? (int64_t)(ky + 1) * p.ry : H) - 1;
    host_tap(sh, ky * p.ry, (int)h, &lo, &t);
    host_tap(sh, ye, (int)h, &t, &hi);
    if (hi - lo + 1 > p.R) p.R = hi - lo + 1;
  }
  p.K = 1;
  for (int kx = 0; kx < p.SX; ++kx) {
    int lo, hi, t;
    const int xe = (int)((int64_t)(kx + 1) * kLuCols < W ? (int64_t)(kx + 1) * kLuCols : W) - 1;
    host_tap(sw, kx * kLuCols, (int)w, &lo, &t);
    host_tap(sw, xe, (int)w, &t, &hi);
    if (hi - lo + 1 > p.K) p.K = hi - lo + 1;
  }
  p.ctas = n * p.SY * p.SX;
  p.off_partial = 32;                                   // [0,16): ticket, [16,32): target counter + its ticket
  p.off_scratch = (p.off_partial + (size_t)p.ctas * 3 * sizeof(double) + 15) & ~(size_t)15;
  const int64_t cp = C == 19 ? 20 : (C == 16 ? 16 : 32);   // class pitch of a patch column (DIGA_DISPATCH_C: 19, 16, padded 32)
  p.bytes = p.off_scratch + (size_t)p.ctas * p.R * cp * p.K * sizeof(float);
  return p;
}

struct LossUpArgs {
  const float* tea;        // [2B,C,h,w] or null
  const float* stu;        // [n,C,h,w]
  const int64_t* target;   // [n_ce,H,W] or null
  const float* weight;     // [C] or null
  int nclass, n, B, n_ce, h, w, H, W;
  float sh, sw, scale, inv_count_kd;
  int size_average;
  const float* up_kd;      // device scalars (backward)
  const float* up_ce;
  const float* denom;
  float up_kd_host;        // used when up_kd == null (single-pass variants)
  float up_ce_host;        // used when up_ce == null
  float denom_host;        // > 0: the caller's #(target >= 0) (size_average, denom == null); checked against the counted value
  const float* sel_pred;   // OHEM (csrc/ohem_up.cu): per-pixel target probability [n_ce,H,W], < 0 = ignored; null = plain CE
  const float* sel_thr;    // OHEM: device scalar, a pixel is kept iff 0 <= sel_pred < sel_thr[0]
  int ignore_label;        // OHEM: target value that is masked out (CE: 255 is >= nclass anyway)
  int ry, R, K;
  int tile;                // 1: the CTA stages its source rows [views][R][nclass][K + 1] in shared memory (cp.async) first
  float* scratch;
  double* partial;
  unsigned int* ticket;
  float* loss_kd;
  float* loss_ce;
  float* denom_out;
  float* loss_total;       // single-pass KD + CE variant: up_ce_host * loss_ce + up_kd_host * loss_kd (:356,:382), or null
};

__device__ __forceinline__ int lu_swz(int x) { return x + (x >> 3); }

#ifndef LU_MINB_LOSS
#define LU_MINB_LOSS 2      // resident CTAs per SM the loss-only kernels are compiled for
#endif
#ifndef LU_MINB_GRAD
#define LU_MINB_GRAD 2      // ... and the gradient kernels (up to 255 registers)
#endif
constexpr float kLuRecurrenceMax = 24.f;   // largest |dif| (log2 units per source row) a warp advances by multiplication

// Dynamic shared memory of the loss kernel, in this order:
//   ytab  [ry] int2          vertical taps of the strip's rows
//   tile                     the strip's source rows [view][row][column][CP] (see the kernel); 0 bytes when the CTA reads
//                            global memory
//   tabT / tabD [CP][128]    (CE) the student's (top, dif) of the current cell: the target-class logit of a pixel is two
//                            shared loads + one FFMA instead of C compares and selects
//   oh [2][CP][128]          (CE, gradient) the one-hot term of the CE gradient per source-row parity; subtracted from the
//                            class gradients when the row is flushed
//   wtab [CP]                (CE) class weights (1 when the caller passes none)
//   tcode [ry][128] uint8    (CE) the strip's targets: class id, 254 = counted but ignored, 255 = not counted
// A per-column table entry is written and read by the thread that owns the column (tcode / wtab: before the first barrier).
struct LuSmem {
  size_t tile, tab, oh, wtab, tcode, total;
};
__host__ __device__ static inline LuSmem lu_smem(int ry, int cp, bool kd, bool ce, bool grad, size_t tile_bytes) {
  LuSmem m;
  m.tile = ((size_t)ry * sizeof(int2) + 15) & ~(size_t)15;
  m.tab = m.tile + tile_bytes;
  m.oh = m.tab + (ce ? (size_t)2 * cp * kLuBlock * sizeof(float) : 0);
  m.wtab = m.oh + ((ce && grad) ? (size_t)2 * cp * kLuBlock * sizeof(float) : 0);
  m.tcode = m.wtab + (ce ? (size_t)cp * sizeof(float) : 0);
  m.total = m.tcode + (ce ? (((size_t)ry * kLuBlock + 15) & ~(size_t)15) : 0);
  return m;
}
__host__ __device__ static inline size_t lu_tile_bytes(int views, int R, int cp, int K) {
  return ((size_t)views * R * (K + 1) * cp * sizeof(float) + 15) & ~(size_t)15;
}
__device__ __forceinline__ void lu_cp_async4(float* dst_shared, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}

// One CTA = 128 output columns x `ry` output rows of one image, one thread per column.
//
// Exponentials: inside a source cell the interpolated logit of a class is linear in the output row, v(Y+1) = v(Y) +
// dif * sh, hence 2^v(Y+1) = 2^v(Y) * 2^(dif * sh): the kernel seeds 2^v and the per-row factor with MUFU.EX2 when the
// walk enters a cell and advances with one FMUL2 per class pair and row (<= 8 rows at stride 8: a few ulp of drift, re-
// seeded at every crossing; the vertical tap fl(sh * Y) - i0 deviates from an exact progression by <= ulp(h), i.e. a
// relative 1e-6 in the exponentials at |dif| = 1).  A warp whose cell holds |dif| > 24 (adjacent source rows > 16 logits
// apart in some class) evaluates every row with MUFU.EX2 and, if a sum still underflows, re-bases on the per-pixel max.
template <int C, bool PAD, bool KD, bool CE, bool LOSS, bool GRAD>
__global__ void __launch_bounds__(kLuBlock, GRAD ? LU_MINB_GRAD : LU_MINB_LOSS)
loss_up_kernel(const LossUpArgs a) {
  extern __shared__ __align__(16) unsigned char lu_dyn[];
  int2* ytab = reinterpret_cast<int2*>(lu_dyn);            // per-row vertical taps of the strip
  constexpr int CP = (C + 3) & ~3, CQ = CP / 4;            // classes padded to a multiple of four
  constexpr int P = CP / 2;                                // class pairs per thread
  using Col = LerpColumnP<P>;
  const int n = blockIdx.z, ky = blockIdx.y, kx = blockIdx.x, tid = threadIdx.x;
  const int X0 = kx * kLuBlock, X = X0 + tid;
  const bool in_range = X < a.W;
  const int Y0 = ky * a.ry, Yend = min(Y0 + a.ry, a.H);
  const int64_t plane = (int64_t)a.h * a.w;
  const int nclass = a.nclass;
  const Tap tx = bilinear_tap(a.sw, in_range ? X : a.W - 1, a.w);
  // Horizontal tap as the column pair (kc, kc + 1): a lane clamped at the last source column (i1 == i0) reads the pair
  // (i0 - 1, i0) with weights (0, l0 + l1) so that the second load is always "first + 1" (a 1-column source has no pair).
  const bool clamped = tx.i1 == tx.i0, pair = a.w > 1;
  const int kc = (clamped && pair) ? tx.i0 - 1 : tx.i0;
  const float l0s = (clamped ? (pair ? 0.f : tx.l0 + tx.l1) : tx.l0) * kLog2e;
  const float l1s = (clamped ? (pair ? tx.l0 + tx.l1 : 0.f) : tx.l1) * kLog2e;
  float wkd = 0.f;
  const float* tbase = nullptr;
  if constexpr (KD) {
    const int nt = n < a.B ? n + a.B : n - a.B;          // the other view supervises this one (loss.py:130-133)
    tbase = a.tea + (int64_t)nt * nclass * plane;
    wkd = n < a.B ? a.scale : 1.f;
  }
  const bool ce_img = CE && n < a.n_ce;
  const LuSmem lay = lu_smem(a.ry, CP, KD, CE, GRAD, a.tile ? lu_tile_bytes(KD ? 2 : 1, a.R, CP, a.K) : 0);
  float* tile = reinterpret_cast<float*>(lu_dyn + lay.tile);
  float* tabT = reinterpret_cast<float*>(lu_dyn + lay.tab) + tid;                  // this column of [CP][128]
  float* tabD = tabT + CP * kLuBlock;
  float* oh = reinterpret_cast<float*>(lu_dyn + lay.oh) + tid;                     // [2][CP][128]
  float* wtab = reinterpret_cast<float*>(lu_dyn + lay.wtab);
  unsigned char* tcode = lu_dyn + lay.tcode;

  float ckd = 0.f, cce = 0.f;
  if constexpr (GRAD) {
    if constexpr (KD) ckd = (a.up_kd != nullptr ? __ldg(a.up_kd) : a.up_kd_host) * a.inv_count_kd * wkd;
    if (ce_img) cce = (a.up_ce != nullptr ? __ldg(a.up_ce) : a.up_ce_host) / (!a.size_average ? 1.0f : a.denom != nullptr ? __ldg(a.denom) : a.denom_host > 0.f ? a.denom_host : 1.0f);
  }

  const bool ohem = CE && a.sel_pred != nullptr;
  const float sel_thr = ohem ? __ldg(a.sel_thr) : 0.f;
  const float* prow = (ohem && ce_img) ? a.sel_pred + ((int64_t)n * a.H) * a.W + (in_range ? X : a.W - 1) : nullptr;

  // ---- CTA geometry: source rows from ylo, source columns xlo..xhi --------------------------------------------------
  const int nvalid = min(kLuBlock, a.W - X0);
  const int xlo = bilinear_tap(a.sw, X0, a.w).i0;
  const int Kt = bilinear_tap(a.sw, X0 + nvalid - 1, a.w).i1 - xlo + 1;
  const int ylo = bilinear_tap(a.sh, Y0, a.h).i0;
  const float* scol = a.stu + (int64_t)n * nclass * plane + (int64_t)ylo * a.w + kc;   // (row ylo, class 0, column kc)
  const float* tcol = KD ? tbase + (int64_t)ylo * a.w + kc : nullptr;
  // The strip's source rows, both views, go to shared memory with ONE round of asynchronous copies (`a.tile`; geometries
  // whose tile would not fit keep reading global memory): a cell crossing then costs shared-memory latency instead of two
  // dependent trips to L2 / DRAM — every source row is first touched by the CTAs that interpolate from it.
  // Layout [view][row][column][CP], first column = kc of the CTA's first thread; the class slots >= nclass hold kLerpPad.
  const Tap t0 = bilinear_tap(a.sw, X0, a.w);
  const int xbase = (t0.i1 == t0.i0 && pair) ? t0.i0 - 1 : t0.i0;
  const int Kp = a.K + 1;
  const int Rn = bilinear_tap(a.sh, Yend - 1, a.h).i1 - ylo + 1;
  if (a.tile) {
    // one (view, row) per warp and step; lanes = columns, the loop walks the classes (source address + plane, tile slot + 1)
    const int ncols = min(Kp, a.w - xbase), lane = tid & 31;
    for (int vr = tid >> 5; vr < (KD ? 2 : 1) * Rn; vr += kLuBlock / 32) {
      const int view = vr >= Rn, r = vr - view * Rn;
      for (int col = lane; col < ncols; col += 32) {
        const float* src = (view ? tcol : scol) - kc + xbase + (int64_t)r * a.w + col;
        float* dst = tile + (vr * Kp + col) * CP;
        for (int c = 0; c < nclass; ++c, src += plane) lu_cp_async4(dst + c, src);
        for (int c = nclass; c < CP; ++c) dst[c] = kLerpPad;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const float* stile = tile + (kc - xbase) * CP;           // (row 0, column kc) of the student view
  const float* ttile = stile + Rn * Kp * CP;
  const int tile_row = Kp * CP, tile_col = pair ? CP : 0;

  // per-row vertical taps, computed once per CTA: {local row of i0 | (i1 - i0) << 16, l1}
  for (int i = tid; i < Yend - Y0; i += kLuBlock) {
    const Tap t = bilinear_tap(a.sh, Y0 + i, a.h);
    ytab[i] = make_int2((t.i0 - ylo) | ((t.i1 - t.i0) << 16), __float_as_int(t.l1));
  }
  if constexpr (CE) {
    if (ce_img) {
      // the strip's targets as one byte per pixel: the int64 loads of the whole strip are in flight together, off the
      // row loop (loss.py:56 counts target >= 0; nll_loss ignores 255 — any id >= C — and the OHEM ignore label)
      const int64_t* tg = a.target + ((int64_t)n * a.H + Y0) * a.W + (in_range ? X : a.W - 1);
      const int rows = Yend - Y0;
      for (int row0 = 0; row0 < rows; row0 += 16) {          // 16 loads in flight per thread, then the conversions
        int64_t t[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) t[k] = ld_stream_i64(tg + (int64_t)min(row0 + k, rows - 1) * a.W);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (row0 + k < rows)
            tcode[(row0 + k) * kLuBlock + tid] = t[k] < 0 ? 255 : (t[k] >= nclass || t[k] == a.ignore_label) ? 254 : (unsigned char)t[k];
      }
      if (tid < CP) wtab[tid] = (a.weight != nullptr && tid < nclass) ? __ldg(a.weight + tid) : 1.f;
      if constexpr (GRAD) {
#pragma unroll
        for (int k = 0; k < 2 * CP; ++k) oh[k * kLuBlock] = 0.f;
      }
    }
  }

  // ---- backward staging (shared memory): [column (swizzled)][class], classes padded to a multiple of four ----------
  __shared__ __align__(16) float sa[GRAD ? kLuLd : 1][GRAD ? CP : 4];
  __shared__ __align__(16) float sb[GRAD ? kLuLd : 1][GRAD ? CP : 4];
  __shared__ int st[GRAD ? kLuBlock + 4 : 1];
  const int64_t cta = ((int64_t)n * gridDim.y + ky) * gridDim.x + kx;
  if constexpr (GRAD) {
    for (int i = tid; i <= Kt; i += kLuBlock) st[i] = nvalid;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  if constexpr (GRAD) {
    if (in_range) {
      const int prev = tid == 0 ? -1 : bilinear_tap(a.sw, X - 1, a.w).i0;
      if (tx.i0 != prev) st[tx.i0 - xlo] = tid;            // first column of the run that maps to source column i0
    }
    __syncthreads();
  }
  float2 Gt[GRAD ? P : 1], Gb[GRAD ? P : 1];
  if constexpr (GRAD) {
#pragma unroll
    for (int p = 0; p < P; ++p) Gt[p] = Gb[p] = make_float2(0.f, 0.f);
  }
  // flush: horizontal half of the transposed interpolation for one completed source row.  Every column stages l0*G
  // (-> source column i0) and l1*G (-> i0 + 1) as float4 class quads; a work item (source column, class quad) sums the
  // contiguous run of output columns that map to it and stores one float4 of the CTA's patch [row][column][CP].
  auto flush = [&](const float2 (&G)[GRAD ? P : 1], int row) {
    if constexpr (GRAD) {
      const int sx = lu_swz(tid);
      const float2 wa = splat2(in_range ? (clamped ? tx.l0 + tx.l1 : tx.l0) : 0.f);
      const float2 wb = splat2((in_range && !clamped) ? tx.l1 : 0.f);
      float* ohs = oh + (row & 1) * (CP * kLuBlock);       // the row's one-hot accumulator (CE images)
#pragma unroll
      for (int q = 0; q < CQ; ++q) {
        float2 g0 = G[2 * q], g1 = G[2 * q + 1];
        if constexpr (CE) {
          if (ce_img) {
            g0.x -= ohs[(4 * q) * kLuBlock], g0.y -= ohs[(4 * q + 1) * kLuBlock];
            g1.x -= ohs[(4 * q + 2) * kLuBlock], g1.y -= ohs[(4 * q + 3) * kLuBlock];
            ohs[(4 * q) * kLuBlock] = 0.f, ohs[(4 * q + 1) * kLuBlock] = 0.f;
            ohs[(4 * q + 2) * kLuBlock] = 0.f, ohs[(4 * q + 3) * kLuBlock] = 0.f;
          }
        }
        const float2 a0 = fmul2(wa, g0), a1 = fmul2(wa, g1), b0 = fmul2(wb, g0), b1 = fmul2(wb, g1);
        *reinterpret_cast<float4*>(&sa[sx][4 * q]) = make_float4(a0.x, a0.y, a1.x, a1.y);
        *reinterpret_cast<float4*>(&sb[sx][4 * q]) = make_float4(b0.x, b0.y, b1.x, b1.y);
      }
      __syncthreads();
      float4* dst = reinterpret_cast<float4*>(a.scratch + ((cta * a.R + row) * a.K) * CP);
      for (int j = tid; j < Kt * CQ; j += kLuBlock) {
        const int xl = j / CQ, q = j - xl * CQ;
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = st[xl]; i < st[xl + 1]; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(&sa[lu_swz(i)][4 * q]);
          sum.x += v.x, sum.y += v.y, sum.z += v.z, sum.w += v.w;
        }
        if (xl > 0)
          for (int i = st[xl - 1]; i < st[xl]; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(&sb[lu_swz(i)][4 * q]);
            sum.x += v.x, sum.y += v.y, sum.z += v.z, sum.w += v.w;
          }
        dst[xl * CQ + q] = sum;
      }
      __syncthreads();
    }
  };
  // one horizontally interpolated source row of a view: from the tile, or (no tile) from global memory
  auto load_row = [&](bool teacher) {
    return [=](float2 (&dst)[P], int row, float sub) {
      Col* none = nullptr;
      if (a.tile) none->hrow_tile(dst, (teacher ? ttile : stile) + row * tile_row, tile_col, l0s, l1s, sub);
      else none->hrow_global(dst, (teacher ? tcol : scol) + (int64_t)row * a.w, plane, pair ? 1 : 0, l0s, l1s, sub, nclass);
    };
  };

  Col cs;                                                    // student (top, dif); the teacher's live only through a crossing
  float2 es[P], ss[P];                                       // 2^v of the current row and its per-row factor
  float2 et[KD ? P : 1], ts[KD ? P : 1];
  bool bigcell = false;                                      // this warp evaluates the cell's rows with MUFU.EX2
  float acc_kd = 0.f, acc_ce = 0.f, acc_cnt = 0.f;
  int cur_r0 = -1, cur_r1 = -1;
  const float sh_step = a.sh;

  int2 yt_next = ytab[0];
  for (int Y = Y0; Y < Yend; ++Y) {
    const int2 yt = yt_next;
    yt_next = ytab[min(Y + 1, Yend - 1) - Y0];               // one row ahead: the tap is not on the row's critical path
    const int r0 = yt.x & 0xffff, r1 = r0 + (yt.x >> 16);
    const float yl1 = __int_as_float(yt.y), yl0 = 1.0f - yl1;
    const float2 yl1v = splat2(yl1);
    if (r0 != cur_r0) {                                      // first row, or crossed into the next source cell (CTA-uniform)
      const bool fresh = cur_r0 < 0;
      if constexpr (GRAD) {
        if (!fresh) {
          flush(Gt, cur_r0);                                 // source row cur_r0 is complete for this strip
#pragma unroll
          for (int p = 0; p < P; ++p) {
            Gt[p] = Gb[p];
            Gb[p] = make_float2(0.f, 0.f);
          }
        }
      }
      cs.enter(fresh, r0, r1, load_row(false));
      const float2 lseed = splat2(yl1 - sh_step), stepv = splat2(sh_step);   // the row loop multiplies before it uses
      float mx = 0.f;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float2 v = ffma2(lseed, cs.dif[p], cs.top[p]), d = fmul2(stepv, cs.dif[p]);
        es[p] = make_float2(fast_ex2(v.x), fast_ex2(v.y));
        ss[p] = make_float2(fast_ex2(d.x), fast_ex2(d.y));
        mx = fmax3(mx, fabsf(cs.dif[p].x), fabsf(cs.dif[p].y));
      }
      if constexpr (KD) {
        Col ct;                                              // re-based on its own cell: nothing carried between cells
        ct.enter(true, r0, r1, load_row(true));
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const float2 u = ffma2(lseed, ct.dif[p], ct.top[p]), d = fmul2(stepv, ct.dif[p]);
          et[p] = make_float2(fast_ex2(u.x), fast_ex2(u.y));
          ts[p] = make_float2(fast_ex2(d.x), fast_ex2(d.y));
          mx = fmax3(mx, fabsf(ct.dif[p].x), fabsf(ct.dif[p].y));
        }
      }
      bigcell = __any_sync(0xffffffffu, mx > kLuRecurrenceMax);
      if constexpr (CE) {
        if (ce_img) {
#pragma unroll
          for (int p = 0; p < P; ++p) {
            tabT[(2 * p) * kLuBlock] = cs.top[p].x, tabD[(2 * p) * kLuBlock] = cs.dif[p].x;
            tabT[(2 * p + 1) * kLuBlock] = cs.top[p].y, tabD[(2 * p + 1) * kLuBlock] = cs.dif[p].y;
          }
        }
      }
      cur_r0 = r0;
      cur_r1 = r1;
    }

    // Per-pixel statistics in the log2 domain, relative to the cell reference (see LerpColumn):
    //   es_c = 2^v_c, Ss = sum es_c, (KD) et_c = 2^u_c, St = sum et_c, cross2 = sum et_c v_c,
    // on class pairs (FMUL2 / FFMA2 / FADD2), two interleaved chains of pairs.
    float Ss = 0.f, St = 0.f, cross2 = 0.f, ms = 0.f;
    if (!bigcell) {
      float2 S2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      [[maybe_unused]] float2 T2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      [[maybe_unused]] float2 X2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int p = 0; p < P; ++p) {
        es[p] = fmul2(es[p], ss[p]);
        S2[p & 1] = fadd2(S2[p & 1], es[p]);
        if constexpr (KD) {
          et[p] = fmul2(et[p], ts[p]);
          T2[p & 1] = fadd2(T2[p & 1], et[p]);
          X2[p & 1] = ffma2(et[p], cs.value2(yl1v, p), X2[p & 1]);
        }
      }
      const float2 sv = fadd2(S2[0], S2[1]);
      Ss = sv.x + sv.y;
      if constexpr (KD) {
        const float2 tv = fadd2(T2[0], T2[1]), xv = fadd2(X2[0], X2[1]);
        St = tv.x + tv.y, cross2 = xv.x + xv.y;
      }
    } else {
      // exact path: MUFU.EX2 per class and row; EXACT re-bases on the per-pixel max (taken only after an underflow)
      Col ct;
      if constexpr (KD) ct.enter(true, cur_r0, cur_r1, load_row(true));
      auto stats = [&](auto exact_tag) {
        constexpr bool EXACT = decltype(exact_tag)::value;
        float mt = 0.f;
        if constexpr (EXACT) {
          ms = mt = -INFINITY;
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const float2 v = cs.value2(yl1v, p);
            ms = fmax3(ms, v.x, v.y);
            if constexpr (KD) {
              const float2 u = ct.value2(yl1v, p);
              mt = fmax3(mt, u.x, u.y);
            }
          }
        }
        const float2 nms = splat2(-ms), nmt = splat2(-mt);
        float2 S2 = make_float2(0.f, 0.f);
        [[maybe_unused]] float2 T2 = make_float2(0.f, 0.f), X2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          float2 v = cs.value2(yl1v, p);
          if constexpr (EXACT) v = fadd2(v, nms);
          es[p] = make_float2(fast_ex2(v.x), fast_ex2(v.y));
          S2 = fadd2(S2, es[p]);
          if constexpr (KD) {
            float2 u = ct.value2(yl1v, p);
            if constexpr (EXACT) u = fadd2(u, nmt);
            et[p] = make_float2(fast_ex2(u.x), fast_ex2(u.y));
            T2 = fadd2(T2, et[p]);
            X2 = ffma2(et[p], v, X2);
          }
        }
        Ss = S2.x + S2.y;
        if constexpr (KD) St = T2.x + T2.y, cross2 = X2.x + X2.y;
      };
      stats(std::false_type{});
      if (Ss < 0x1p-60f || (KD && St < 0x1p-60f)) stats(std::true_type{});   // adjacent source rows > 41 logits apart
    }

    float inv_t = 0.f;
    if constexpr (KD) inv_t = fast_rcp(St);
    float wt = 1.f, dtgt = 0.f;
    int code = 255;
    bool counted = false, valid = false;
    if constexpr (CE) {
      if (ce_img) {
        code = tcode[(Y - Y0) * kLuBlock + tid];
        counted = in_range && code != 255;                            // loss.py:56  mask = target >= 0
        valid = in_range && code < 254;
        if (ohem && valid) {                                          // OhemCrossEntropy keeps the hard pixels only
          const float pv = __ldg(prow + (int64_t)Y * a.W);
          valid = pv >= 0.f && pv < sel_thr;
        }
        if (valid) {
          wt = wtab[code];
          dtgt = fmaf(yl1, tabD[code * kLuBlock], tabT[code * kLuBlock]) - ms;   // same frame as the sums
        }
      }
    }
    if constexpr (LOSS) {                                            // accumulated in log2 units (x ln 2 at the end)
      const float lse2 = fast_lg2(Ss);
      if constexpr (KD) {
        if (in_range) acc_kd += wkd * (lse2 - cross2 * inv_t);
      }
      if constexpr (CE) {
        if (valid) acc_ce += wt * (lse2 - dtgt);
        if (counted) acc_cnt += 1.f;
      }
    }
    if constexpr (GRAD) {
      const float cpx = (CE && valid) ? wt * cce : 0.f;
      const float2 ga = splat2((ckd + cpx) * fast_rcp(Ss)), ngb = splat2(-(ckd * inv_t));
      const float2 yl0v = splat2(yl0);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float2 g = fmul2(ga, es[p]);
        if constexpr (KD) g = ffma2(ngb, et[p], g);
        Gt[p] = ffma2(yl0v, g, Gt[p]);
        Gb[p] = ffma2(yl1v, g, Gb[p]);
      }
      if constexpr (CE) {
        if (valid) {                                                 // one-hot term, per source-row parity (see flush)
          float* o0 = oh + ((r0 & 1) * CP + code) * kLuBlock;
          *o0 += yl0 * cpx;
          float* o1 = oh + ((r1 & 1) * CP + code) * kLuBlock;
          *o1 += yl1 * cpx;
        }
      }
    }
  }

  if constexpr (GRAD) {
    if (cur_r1 == cur_r0) {                                  // clamped at the last source row: both taps hit it
#pragma unroll
      for (int p = 0; p < P; ++p) Gt[p] = fadd2(Gt[p], Gb[p]);
      flush(Gt, cur_r0);
    } else {
      flush(Gt, cur_r0);
      flush(Gb, cur_r1);
    }
  }

  if constexpr (LOSS) {
    __shared__ float red[3][kLuBlock / 32];
    __shared__ bool is_last;
    const float wk = warp_sum(acc_kd), wc = CE ? warp_sum(acc_ce) : 0.f, wn = CE ? warp_sum(acc_cnt) : 0.f;
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = wk;
      red[1][tid >> 5] = wc;
      red[2][tid >> 5] = wn;
    }
    __syncthreads();
    if (tid == 0) {
      float bk = 0.f, bc = 0.f, bn = 0.f;
#pragma unroll
      for (int i = 0; i < kLuBlock / 32; ++i) {
        bk += red[0][i];
        bc += red[1][i];
        bn += red[2][i];
      }
      a.partial[3 * cta + 0] = (double)bk;
      a.partial[3 * cta + 1] = (double)bc;
      a.partial[3 * cta + 2] = (double)bn;
      __threadfence();
      const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
      is_last = (atomicAdd(a.ticket, 1u) == total - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      __shared__ double dr[3][kLuBlock];
      const int64_t total = (int64_t)gridDim.x * gridDim.y * gridDim.z;
      double v0 = 0.0, v1 = 0.0, v2 = 0.0;
      for (int64_t i = tid; i < total; i += kLuBlock) {    // fixed partition, fixed order: deterministic
        v0 += __ldcg(&a.partial[3 * i + 0]);
        v1 += __ldcg(&a.partial[3 * i + 1]);
        v2 += __ldcg(&a.partial[3 * i + 2]);
      }
      dr[0][tid] = v0;
      dr[1][tid] = v1;
      dr[2][tid] = v2;
      __syncthreads();
      for (int o = kLuBlock / 2; o > 0; o >>= 1) {
        if (tid < o) {
          dr[0][tid] += dr[0][tid + o];
          dr[1][tid] += dr[1][tid + o];
          dr[2][tid] += dr[2][tid + o];
        }
        __syncthreads();
      }
      if (tid == 0) {
        if (KD && a.loss_kd) a.loss_kd[0] = (float)(dr[0][0] * 0.6931471805599453 * (double)a.inv_count_kd);
        if constexpr (CE) {
          const float cnt = (float)dr[2][0], tot = (float)(dr[1][0] * 0.6931471805599453);
          if (a.loss_ce) a.loss_ce[0] = a.size_average ? tot / cnt : tot;   // fp32 division like `loss /= mask.data.sum()`
          if (a.denom_out) a.denom_out[0] = cnt;
          if constexpr (KD && LOSS && GRAD) {                // the weighted sum the call site would form with three scalar ops
            if (a.loss_total)
              a.loss_total[0] = __fadd_rn(__fmul_rn(a.up_ce_host, a.size_average ? tot / cnt : tot),
                                          __fmul_rn(a.up_kd_host, (float)(dr[0][0] * 0.6931471805599453 * (double)a.inv_count_kd)));
            // a promised denominator that the count contradicts scaled the CE gradient wrongly: fail loudly, not quietly
            if (a.size_average && a.denom == nullptr && a.denom_host > 0.f && cnt != a.denom_host) {
              const float bad = __int_as_float(0x7fc00000);
              if (a.loss_ce) a.loss_ce[0] = bad;
              if (a.loss_total) a.loss_total[0] = bad;
            }
          }
        }
        *a.ticket = 0;                                       // leave the workspace ready for the next launch
      }
    }
  }
}

// dlow[n,c,y,x] = sum of the scratch patches that cover (y,x): strips in increasing ky, tiles in increasing kx (fixed
// order: deterministic).  One CTA per (image, source row): the matching strips are CTA-uniform, a thread owns one
// source column, resolves its patches (at most 2 x 2, else the generic loop) and adds their class vectors as float4s;
// the stores are coalesced per class plane.  `num` != null (the deferred gather of the autograd backward): every sum is
// multiplied by num[0] / den[0] (den null: num[0]) — the upstream scalar, read on the device — before it is stored, the
// same two roundings as gathering first and scaling the tensor afterwards.
template <int C, bool PAD>
__global__ void __launch_bounds__(128)
loss_up_gather_kernel(const float* __restrict__ scratch, float* __restrict__ dlow, int nclass, int h, int w, int H, int W,
                      float sh, float sw, float inv_sh_ry, float inv_sw_bx, int ry, int R, int K, int SX, int SY,
                      const float* __restrict__ num, const float* __restrict__ den) {
  constexpr int CP = (C + 3) & ~3, CQ = CP / 4;
  const int y = blockIdx.x, img = blockIdx.y;
  const bool scaled = num != nullptr;
  const float coef = scaled ? (den != nullptr ? __fdiv_rn(__ldg(num), __ldg(den)) : __ldg(num)) : 1.f;
  // candidate strips: output rows with i0 in {y-1, y} lie in [(y-1)/sh, (y+1)/sh]; one strip of slack either side,
  // then the exact test with the kernel's own tap arithmetic (matching strips are contiguous: ylo, yhi are monotone)
  int kyA = 0, kyB = SY - 1;
  if (sh > 0.f) {
    kyA = max(0, (int)((float)max(y - 1, 0) * inv_sh_ry) - 1);
    kyB = min(SY - 1, (int)((float)(y + 1) * inv_sh_ry) + 1);
  }
  while (kyA <= kyB && bilinear_tap(sh, min((kyA + 1) * ry, H) - 1, h).i1 < y) ++kyA;
  while (kyB >= kyA && bilinear_tap(sh, kyB * ry, h).i0 > y) --kyB;
  const int64_t plane = (int64_t)h * w;
  for (int x = threadIdx.x; x < w; x += 128) {
    int kxA = 0, kxB = SX - 1;
    if (sw > 0.f) {
      kxA = max(0, (int)((float)max(x - 1, 0) * inv_sw_bx) - 1);
      kxB = min(SX - 1, (int)((float)(x + 1) * inv_sw_bx) + 1);
    }
    while (kxA <= kxB && bilinear_tap(sw, min((kxA + 1) * kLuCols, W) - 1, w).i1 < x) ++kxA;
    while (kxB >= kxA && bilinear_tap(sw, kxB * kLuCols, w).i0 > x) --kxB;
    auto patch_ptr = [&](int ky, int kx) {
      const int ylo = bilinear_tap(sh, ky * ry, h).i0, xlo = bilinear_tap(sw, kx * kLuCols, w).i0;
      const int64_t cta = ((int64_t)img * SY + ky) * SX + kx;
      return reinterpret_cast<const float4*>(scratch + (((cta * R + (y - ylo)) * K) + (x - xlo)) * CP);
    };
    float4 acc[CQ];
#pragma unroll
    for (int q = 0; q < CQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add = [&](const float4* p) {
#pragma unroll
      for (int q = 0; q < CQ; ++q) {
        const float4 v = __ldcg(p + q);
        acc[q].x += v.x, acc[q].y += v.y, acc[q].z += v.z, acc[q].w += v.w;
      }
    };
    const int ny = kyB - kyA + 1, nx = kxB - kxA + 1;
    if (ny >= 1 && ny <= 2 && nx >= 1 && nx <= 2) {
      add(patch_ptr(kyA, kxA));
      if (nx == 2) add(patch_ptr(kyA, kxB));
      if (ny == 2) add(patch_ptr(kyB, kxA));
      if (nx == 2 && ny == 2) add(patch_ptr(kyB, kxB));
    } else {
      for (int ky = kyA; ky <= kyB; ++ky)
        for (int kx = kxA; kx <= kxB; ++kx) add(patch_ptr(ky, kx));
    }
    float* dst = dlow + (int64_t)img * nclass * plane + (int64_t)y * w + x;
    const float* av = reinterpret_cast<const float*>(acc);
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) dst[c * plane] = scaled ? __fmul_rn(av[c], coef) : av[c];
  }
}

// denom = #(target >= 0) (util/loss.py:56,:60) ahead of a single-pass loss+gradient launch, whose CE gradient needs it.
// Integer atomics: exact and order-independent.  `counter` = {count, ticket}, left zeroed.
__global__ void __launch_bounds__(256)
count_targets_kernel(const int64_t* __restrict__ target, int64_t total, unsigned long long* __restrict__ counter,
                     float* __restrict__ denom_out) {
  unsigned int cnt = 0;
  const int64_t pairs = total / 2;
  const bool vec = (reinterpret_cast<uintptr_t>(target) & 15) == 0;
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * 256) {
      const longlong2 t = ld_stream_i64x2(target + 2 * i);
      cnt += (t.x >= 0) + (t.y >= 0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (total & 1)) cnt += target[total - 1] >= 0;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) cnt += target[i] >= 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  __shared__ unsigned int wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int b = 0;
    for (int i = 0; i < 8; ++i) b += wsum[i];
    atomicAdd(&counter[0], (unsigned long long)b);
    __threadfence();
    if (atomicAdd(&counter[1], 1ull) == gridDim.x - 1) {
      __threadfence();
      denom_out[0] = (float)atomicExch(&counter[0], 0ull);
      counter[1] = 0ull;
    }
  }
}

static int check_common(const char* who, const float* stu, int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W,
                        const void* workspace) {
  DIGA_REQUIRE(stu && workspace, DIGA_ERR_INVALID, "%s: null input / workspace", who);
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "%s: C=%lld outside [1,%d]", who, (long long)C, DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 1 && n <= 65535 && h >= 1 && w >= 1 && H >= h && W >= w && H < (1 << 24) && W < (1 << 24), DIGA_ERR_INVALID,
               "%s: needs 1 <= n <= 65535 and an up-sampling geometry (H >= h, W >= w); got n=%lld %lldx%lld -> %lldx%lld", who,
               (long long)n, (long long)h, (long long)w, (long long)H, (long long)W);
  DIGA_REQUIRE(aligned(stu, 4) && aligned(workspace, 16), DIGA_ERR_MISALIGNED, "%s: misaligned pointer", who);
  return DIGA_OK;
}

template <bool KD, bool CE, bool LOSS, bool GRAD>
static int launch_loss_up(LossUpArgs a, const LossUpPlan& p, int64_t C, float* dlow, cudaStream_t st) {
  dim3 grid((unsigned)p.SX, (unsigned)p.SY, (unsigned)a.n);
  DIGA_REQUIRE((size_t)p.ry * sizeof(int2) <= 32 * 1024, DIGA_ERR_INVALID, "loss_up: strip height %d too large", p.ry);
  const size_t tile_bytes = lu_tile_bytes(KD ? 2 : 1, p.R, ((int)C == 19 || (int)C == 16) ? (((int)C + 3) & ~3) : 32, p.K);
  a.tile = tile_bytes <= kLuTileMax && tunable("lossup_tile", 1) != 0;
  DIGA_DISPATCH_C(C, {
    const size_t dyn = lu_smem(p.ry, (kC + 3) & ~3, KD, CE, GRAD, a.tile ? tile_bytes : 0).total;
    auto kernel = loss_up_kernel<kC, kPad, KD, CE, LOSS, GRAD>;
    static size_t configured_dev[64] = {0};            // per instantiation and device
    size_t& configured = configured_dev[device_slot()];
    if (configured < dyn) {                              // static staging + dynamic tables can pass the 48 KB default
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("loss_up: cannot reserve %zu bytes of shared memory", dyn);
        return DIGA_ERR_CUDA;
      }
      configured = dyn;
    }
    kernel<<<grid, kLuBlock, dyn, st>>>(a);
    DIGA_CHECK_LAUNCH("loss_up_kernel");
    if (GRAD && dlow != nullptr) {                       // dlow null: the patches stay in the caller's scratch (deferred gather)
      const float inv_sh_ry = a.sh > 0.f ? 1.0f / (a.sh * (float)p.ry) : 0.f;
      const float inv_sw_bx = a.sw > 0.f ? 1.0f / (a.sw * (float)kLuCols) : 0.f;
      loss_up_gather_kernel<kC, kPad><<<dim3((unsigned)a.h, (unsigned)a.n), 128, 0, st>>>(
          a.scratch, dlow, a.nclass, a.h, a.w, a.H, a.W, a.sh, a.sw, inv_sh_ry, inv_sw_bx, p.ry, p.R, p.K, p.SX, p.SY, nullptr, nullptr);
      DIGA_CHECK_LAUNCH("loss_up_gather_kernel");
    }
  });
  return DIGA_OK;
}

static LossUpArgs fill_args(const LossUpPlan& p, void* workspace, const float* tea, const float* stu, const int64_t* target,
                            const float* weight, int64_t n, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W,
                            float scale, int size_average) {
  LossUpArgs a{};
  a.tea = tea;
  a.stu = stu;
  a.target = target;
  a.weight = weight;
  a.nclass = (int)C;
  a.n = (int)n;
  a.B = (int)(n / 2);
  a.n_ce = (int)n_ce;
  a.h = (int)h;
  a.w = (int)w;
  a.H = (int)H;
  a.W = (int)W;
  a.sh = bilinear_scale_host(h, H);
  a.sw = bilinear_scale_host(w, W);
  a.scale = scale;
  a.inv_count_kd = (float)(1.0 / ((double)(n / 2 > 0 ? n / 2 : 1) * (double)H * (double)W));
  a.size_average = size_average;
  a.ignore_label = DIGA_IGNORE_LABEL;
  a.ry = p.ry;
  a.R = p.R;
  a.K = p.K;
  char* ws = reinterpret_cast<char*>(workspace);
  a.ticket = reinterpret_cast<unsigned int*>(ws);
  a.partial = reinterpret_cast<double*>(ws + p.off_partial);
  a.scratch = reinterpret_cast<float*>(ws + p.off_scratch);
  return a;
}

}  // namespace diga

extern "C" {

size_t diga_loss_up_workspace_bytes(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W) {
  if (n < 1 || C < 1 || h < 1 || w < 1 || H < 1 || W < 1) return 0;
  return diga::make_plan(n, C, h, w, H, W).bytes;
}

int diga_loss_up_fwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                     int64_t n, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                     int size_average, float* loss_kd, float* loss_ce, float* denom_out, void* workspace,
                     diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("loss_up_fwd", student_low, n, C, h, w, H, W, workspace)) return rc;
  const bool kd = teacher_low != nullptr, ce = target != nullptr;
  DIGA_REQUIRE(kd || ce, DIGA_ERR_INVALID, "loss_up_fwd: neither teacher nor target given");
  DIGA_REQUIRE(!kd || ((n % 2) == 0 && loss_kd), DIGA_ERR_INVALID, "loss_up_fwd: KD needs an even batch (two views) and loss_kd");
  DIGA_REQUIRE(!ce || (n_ce >= 1 && n_ce <= n && loss_ce && denom_out), DIGA_ERR_INVALID,
               "loss_up_fwd: CE needs 1 <= n_ce <= n, loss_ce and denom_out");
  DIGA_REQUIRE(aligned(teacher_low, 4) && aligned(target, 8) && aligned(weight, 4), DIGA_ERR_MISALIGNED,
               "loss_up_fwd: misaligned pointer");
  const LossUpPlan p = make_plan(n, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, teacher_low, student_low, target, weight, n, ce ? n_ce : 0, C, h, w, H, W, scale, size_average);
  a.loss_kd = loss_kd;
  a.loss_ce = loss_ce;
  a.denom_out = denom_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (kd && ce) return launch_loss_up<true, true, true, false>(a, p, C, nullptr, st);
  if (kd) return launch_loss_up<true, false, true, false>(a, p, C, nullptr, st);
  return launch_loss_up<false, true, true, false>(a, p, C, nullptr, st);
}

int diga_loss_up_bwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                     int64_t n, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                     int size_average, const float* upstream_kd, const float* upstream_ce, const float* denom,
                     float* dstudent_low, void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("loss_up_bwd", student_low, n, C, h, w, H, W, workspace)) return rc;
  const bool kd = teacher_low != nullptr, ce = target != nullptr;
  DIGA_REQUIRE(kd || ce, DIGA_ERR_INVALID, "loss_up_bwd: neither teacher nor target given");
  DIGA_REQUIRE(dstudent_low, DIGA_ERR_INVALID, "loss_up_bwd: dstudent_low required");
  DIGA_REQUIRE(!kd || ((n % 2) == 0 && upstream_kd), DIGA_ERR_INVALID, "loss_up_bwd: KD needs an even batch and upstream_kd");
  DIGA_REQUIRE(!ce || (n_ce >= 1 && n_ce <= n && upstream_ce && (!size_average || denom)), DIGA_ERR_INVALID,
               "loss_up_bwd: CE needs 1 <= n_ce <= n, upstream_ce and denom");
  DIGA_REQUIRE(aligned(teacher_low, 4) && aligned(target, 8) && aligned(weight, 4) && aligned(dstudent_low, 4), DIGA_ERR_MISALIGNED,
               "loss_up_bwd: misaligned pointer");
  const LossUpPlan p = make_plan(n, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, teacher_low, student_low, target, weight, n, ce ? n_ce : 0, C, h, w, H, W, scale, size_average);
  a.up_kd = upstream_kd;
  a.up_ce = upstream_ce;
  a.denom = denom;
  cudaStream_t st = (cudaStream_t)stream;
  if (kd && ce) return launch_loss_up<true, true, false, true>(a, p, C, dstudent_low, st);
  if (kd) return launch_loss_up<true, false, false, true>(a, p, C, dstudent_low, st);
  return launch_loss_up<false, true, false, true>(a, p, C, dstudent_low, st);
}

int diga_kd_up_fwd_bwd(const float* teacher_low, const float* student_low, int64_t n2, int64_t C, int64_t h, int64_t w,
                       int64_t H, int64_t W, float scale, float upstream_host, float* loss_out, float* dstudent_low,
                       float* scratch, void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("kd_up_fwd_bwd", student_low, n2, C, h, w, H, W, workspace)) return rc;
  DIGA_REQUIRE(teacher_low && loss_out && (dstudent_low || scratch) && (n2 % 2) == 0, DIGA_ERR_INVALID,
               "kd_up_fwd_bwd: teacher, loss_out, dstudent_low (or scratch) and an even batch are required");
  DIGA_REQUIRE(aligned(teacher_low, 4) && aligned(dstudent_low, 4) && aligned(scratch, 16), DIGA_ERR_MISALIGNED,
               "kd_up_fwd_bwd: misaligned pointer");
  const LossUpPlan p = make_plan(n2, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, teacher_low, student_low, nullptr, nullptr, n2, 0, C, h, w, H, W, scale, 1);
  if (scratch) a.scratch = scratch;
  a.up_kd_host = upstream_host;
  a.loss_kd = loss_out;
  return launch_loss_up<true, false, true, true>(a, p, C, dstudent_low, (cudaStream_t)stream);
}

int diga_ce_up_fwd_bwd(const float* logits_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h,
                       int64_t w, int64_t H, int64_t W, int size_average, float* loss_out, float* denom_out,
                       float* dlogits_sum, float* scratch, void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("ce_up_fwd_bwd", logits_low, n, C, h, w, H, W, workspace)) return rc;
  DIGA_REQUIRE(target && loss_out && denom_out && (dlogits_sum || scratch), DIGA_ERR_INVALID, "ce_up_fwd_bwd: null pointer");
  DIGA_REQUIRE(aligned(target, 8) && aligned(weight, 4) && aligned(dlogits_sum, 4) && aligned(scratch, 16), DIGA_ERR_MISALIGNED,
               "ce_up_fwd_bwd: misaligned pointer");
  const LossUpPlan p = make_plan(n, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, nullptr, logits_low, target, weight, n, n, C, h, w, H, W, 0.f, size_average);
  if (scratch) a.scratch = scratch;
  a.loss_ce = loss_out;
  a.denom_out = denom_out;
  a.up_ce_host = 1.0f;              // unit upstream, no denominator: the gradient of the SUMMED loss (the caller scales it)
  return launch_loss_up<false, true, true, true>(a, p, C, dlogits_sum, (cudaStream_t)stream);
}

int diga_seg_kd_up_fwd_bwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                           int64_t n2, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                           int size_average, float lambda_ce_host, float lambda_kd_host, float denom_known, float* loss_kd,
                           float* loss_ce, float* denom_out, float* loss_total, float* dstudent_low, float* scratch,
                           void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("seg_kd_up_fwd_bwd", student_low, n2, C, h, w, H, W, workspace)) return rc;
  DIGA_REQUIRE(teacher_low && target && loss_kd && loss_ce && denom_out && (dstudent_low || scratch) && (n2 % 2) == 0 && n_ce >= 1 &&
                   n_ce <= n2,
               DIGA_ERR_INVALID, "seg_kd_up_fwd_bwd: all pointers, an even batch and 1 <= n_ce <= n2 are required");
  DIGA_REQUIRE(aligned(teacher_low, 4) && aligned(target, 8) && aligned(weight, 4) && aligned(dstudent_low, 4) && aligned(scratch, 16),
               DIGA_ERR_MISALIGNED, "seg_kd_up_fwd_bwd: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const LossUpPlan p = make_plan(n2, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, teacher_low, student_low, target, weight, n2, n_ce, C, h, w, H, W, scale, size_average);
  if (scratch) a.scratch = scratch;
  DIGA_REQUIRE(denom_known >= 0.f, DIGA_ERR_INVALID, "seg_kd_up_fwd_bwd: negative denom_known");
  if (size_average && denom_known > 0.f) {                // the caller knows #(target >= 0) (e.g. loader labels: trainIds or 255)
    a.denom_host = denom_known;
  } else if (size_average) {                              // the CE gradient is divided by #(target >= 0): count it first
    const int64_t total = n_ce * H * W;
    int64_t grid = (total / 2 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    count_targets_kernel<<<(unsigned)grid, 256, 0, st>>>(target, total, reinterpret_cast<unsigned long long*>(
                                                             reinterpret_cast<char*>(workspace) + 16), denom_out);
    DIGA_CHECK_LAUNCH("count_targets_kernel");
    a.denom = denom_out;
  }
  a.up_kd_host = lambda_kd_host;
  a.up_ce_host = lambda_ce_host;
  a.loss_kd = loss_kd;
  a.loss_ce = loss_ce;
  a.denom_out = denom_out;                                // rewritten with the same value by the loss reduction
  a.loss_total = loss_total;
  return launch_loss_up<true, true, true, true>(a, p, C, dstudent_low, st);
}

size_t diga_loss_up_scratch_bytes(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W) {
  if (n < 1 || C < 1 || h < 1 || w < 1 || H < 1 || W < 1) return 0;
  const diga::LossUpPlan p = diga::make_plan(n, C, h, w, H, W);
  return p.bytes - p.off_scratch;
}

int diga_loss_up_gather(const float* scratch, const float* num, const float* den, int64_t n, int64_t C, int64_t h, int64_t w,
                        int64_t H, int64_t W, float* dlow, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(scratch && num && dlow, DIGA_ERR_INVALID, "loss_up_gather: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES && n >= 1 && n <= 65535 && h >= 1 && w >= 1 && H >= h && W >= w && H < (1 << 24) &&
                   W < (1 << 24),
               DIGA_ERR_INVALID, "loss_up_gather: bad geometry");
  DIGA_REQUIRE(aligned(scratch, 16) && aligned(num, 4) && aligned(den, 4) && aligned(dlow, 4), DIGA_ERR_MISALIGNED,
               "loss_up_gather: misaligned pointer");
  const LossUpPlan p = make_plan(n, C, h, w, H, W);
  const float sh = bilinear_scale_host(h, H), sw = bilinear_scale_host(w, W);
  const float inv_sh_ry = sh > 0.f ? 1.0f / (sh * (float)p.ry) : 0.f;
  const float inv_sw_bx = sw > 0.f ? 1.0f / (sw * (float)kLuCols) : 0.f;
  cudaStream_t st = (cudaStream_t)stream;
  DIGA_DISPATCH_C(C, {
    loss_up_gather_kernel<kC, kPad><<<dim3((unsigned)h, (unsigned)n), 128, 0, st>>>(
        scratch, dlow, (int)C, (int)h, (int)w, (int)H, (int)W, sh, sw, inv_sh_ry, inv_sw_bx, p.ry, p.R, p.K, p.SX, p.SY, num, den);
    DIGA_CHECK_LAUNCH("loss_up_gather_kernel");
  });
  return DIGA_OK;
}

/* OhemCrossEntropy backward (csrc/ohem_up.cu holds the forward): the CE gradient pass restricted to the kept pixels
 * (0 <= pred < thr[0]), divided by their number `count[0]`. */
int diga_ohem_up_bwd(const float* logits_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h, int64_t w,
                     int64_t H, int64_t W, int64_t ignore_label, const float* pred, const float* thr, const float* count,
                     const float* upstream, float* dlogits_low, void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("ohem_up_bwd", logits_low, n, C, h, w, H, W, workspace)) return rc;
  DIGA_REQUIRE(target && pred && thr && count && upstream && dlogits_low, DIGA_ERR_INVALID, "ohem_up_bwd: null pointer");
  DIGA_REQUIRE(aligned(target, 8) && aligned(weight, 4) && aligned(pred, 4) && aligned(dlogits_low, 4), DIGA_ERR_MISALIGNED,
               "ohem_up_bwd: misaligned pointer");
  const LossUpPlan p = make_plan(n, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, nullptr, logits_low, target, weight, n, n, C, h, w, H, W, 0.f, 1);
  a.up_ce = upstream;
  a.denom = count;
  a.sel_pred = pred;
  a.sel_thr = thr;
  a.ignore_label = (int)ignore_label;
  return launch_loss_up<false, true, false, true>(a, p, C, dlogits_low, (cudaStream_t)stream);
}

}  // extern "C"
