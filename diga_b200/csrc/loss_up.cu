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
//   the CTA writes its partial low-resolution patch [R rows][C][K cols] to a scratch slot of its own;
//   gather     : a second, tiny kernel adds the <= 2x2 patches that overlap each low-resolution element, in fixed order.
#include "bilinear.cuh"
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

constexpr int kLuBlock = 128;
constexpr int kLuLd = kLuBlock + kLuBlock / 8 + 2;  // row pitch of the swizzled staging arrays

struct LossUpPlan {
  int ry, SX, SY, R, K;
  int64_t ctas;
  size_t off_partial, off_scratch, bytes;
};

// The tap arithmetic of bilinear_tap() on the host (IEEE single multiply + truncation: identical results).
static inline void host_tap(float scale, int dst, int in, int* i0, int* i1) {
  const float src = scale * (float)dst;
  *i0 = (int)src;
  *i1 = *i0 + ((*i0 < in - 1) ? 1 : 0);
}

static LossUpPlan make_plan(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W) {
  LossUpPlan p;
  p.ry = tunable("lossup_ry", 16);
  if (p.ry < 1) p.ry = 1;
  p.SX = (int)((W + kLuBlock - 1) / kLuBlock);
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
    const int xe = (int)((int64_t)(kx + 1) * kLuBlock < W ? (int64_t)(kx + 1) * kLuBlock : W) - 1;
    host_tap(sw, kx * kLuBlock, (int)w, &lo, &t);
    host_tap(sw, xe, (int)w, &t, &hi);
    if (hi - lo + 1 > p.K) p.K = hi - lo + 1;
  }
  p.ctas = n * p.SY * p.SX;
  p.off_partial = 16;
  p.off_scratch = (p.off_partial + (size_t)p.ctas * 3 * sizeof(double) + 15) & ~(size_t)15;
  p.bytes = p.off_scratch + (size_t)p.ctas * p.R * C * p.K * sizeof(float);
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
  float up_kd_host;        // used when up_kd == null (single-pass KD)
  int ry, R, K;
  float* scratch;
  double* partial;
  unsigned int* ticket;
  float* loss_kd;
  float* loss_ce;
  float* denom_out;
};

__device__ __forceinline__ int lu_swz(int x) { return x + (x >> 3); }

template <int C, bool PAD, bool KD, bool CE, bool LOSS, bool GRAD>
__global__ void __launch_bounds__(kLuBlock, GRAD ? 2 : 3)
loss_up_kernel(const LossUpArgs a) {
  const int n = blockIdx.z, ky = blockIdx.y, kx = blockIdx.x, tid = threadIdx.x;
  const int X0 = kx * kLuBlock, X = X0 + tid;
  const bool in_range = X < a.W;
  const int Y0 = ky * a.ry, Yend = min(Y0 + a.ry, a.H);
  const int64_t plane = (int64_t)a.h * a.w;
  const int nclass = a.nclass;
  Tap tx[1];
  tx[0] = bilinear_tap(a.sw, in_range ? X : a.W - 1, a.w);
  const float* sbase = a.stu + (int64_t)n * nclass * plane;
  const float* tbase = nullptr;
  float wkd = 0.f;
  if constexpr (KD) {
    const int nt = n < a.B ? n + a.B : n - a.B;          // the other view supervises this one (loss.py:130-133)
    tbase = a.tea + (int64_t)nt * nclass * plane;
    wkd = n < a.B ? a.scale : 1.f;
  }
  const bool ce_img = CE && n < a.n_ce;
  const int64_t* trow = ce_img ? a.target + ((int64_t)n * a.H) * a.W + (in_range ? X : a.W - 1) : nullptr;

  float ckd = 0.f, cce = 0.f;
  if constexpr (GRAD) {
    if constexpr (KD) ckd = (a.up_kd != nullptr ? __ldg(a.up_kd) : a.up_kd_host) * a.inv_count_kd * wkd;
    if (ce_img) cce = __ldg(a.up_ce) / (a.size_average ? __ldg(a.denom) : 1.0f);
  }

  // ---- backward staging (shared memory) ------------------------------------------------------------------------------
  __shared__ float sa[GRAD ? C : 1][GRAD ? kLuLd : 1];
  __shared__ float sb[GRAD ? C : 1][GRAD ? kLuLd : 1];
  __shared__ int st[GRAD ? kLuBlock + 4 : 1];
  int xlo = 0, Kt = 0, ylo = 0;
  const int64_t cta = ((int64_t)n * gridDim.y + ky) * gridDim.x + kx;
  if constexpr (GRAD) {
    const int nvalid = min(kLuBlock, a.W - X0);
    xlo = bilinear_tap(a.sw, X0, a.w).i0;
    Kt = bilinear_tap(a.sw, X0 + nvalid - 1, a.w).i1 - xlo + 1;
    ylo = bilinear_tap(a.sh, Y0, a.h).i0;
    for (int i = tid; i <= Kt; i += kLuBlock) st[i] = nvalid;
    __syncthreads();
    if (in_range) {
      const int prev = tid == 0 ? -1 : bilinear_tap(a.sw, X - 1, a.w).i0;
      if (tx[0].i0 != prev) st[tx[0].i0 - xlo] = tid;      // first column of the run that maps to source column i0
    }
    __syncthreads();
  }
  float Gt[GRAD ? C : 1], Gb[GRAD ? C : 1];
  if constexpr (GRAD) {
#pragma unroll
    for (int c = 0; c < C; ++c) Gt[c] = Gb[c] = 0.f;
  }
  auto flush = [&](const float (&G)[GRAD ? C : 1], int row) {
    if constexpr (GRAD) {
      const bool clamped = tx[0].i1 == tx[0].i0;
      const int sx = lu_swz(tid);
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          const float gv = in_range ? G[c] : 0.f;
          float va = tx[0].l0 * gv, vb = tx[0].l1 * gv;
          if (clamped) {
            va += vb;
            vb = 0.f;
          }
          sa[c][sx] = va;
          sb[c][sx] = vb;
        }
      __syncthreads();
      float* dst = a.scratch + ((cta * a.R + row) * nclass) * a.K;
      for (int j = tid; j < nclass * Kt; j += kLuBlock) {
        const int c = j / Kt, xl = j - c * Kt;
        float sum = 0.f;
        for (int i = st[xl]; i < st[xl + 1]; ++i) sum += sa[c][lu_swz(i)];
        if (xl > 0)
          for (int i = st[xl - 1]; i < st[xl]; ++i) sum += sb[c][lu_swz(i)];
        dst[c * a.K + xl] = sum;
      }
      __syncthreads();
    }
  };

  ColumnInterp<C, PAD, 1> cs;
  ColumnInterp<KD ? C : 1, PAD, 1> ct;
  float acc_kd = 0.f, acc_ce = 0.f, acc_cnt = 0.f;
  int cur_i0 = -1, cur_i1 = -1;
  int64_t tgt = 0, tgt_next = 0;
  if (ce_img) tgt = ld_stream_i64(trow + (int64_t)Y0 * a.W);

  for (int Y = Y0; Y < Yend; ++Y) {
    if (ce_img && Y + 1 < Yend) tgt_next = ld_stream_i64(trow + (int64_t)(Y + 1) * a.W);
    const Tap ty = bilinear_tap(a.sh, Y, a.h);
    if constexpr (GRAD) {
      if (cur_i0 >= 0 && ty.i0 != cur_i0) {                // crossed into the next source cell: row cur_i0 is complete
        flush(Gt, cur_i0 - ylo);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          Gt[c] = Gb[c];
          Gb[c] = 0.f;
        }
      }
    }
    cur_i0 = ty.i0;
    cur_i1 = ty.i1;
    cs.seek(ty, sbase, plane, a.w, tx, nclass);
    if constexpr (KD) ct.seek(ty, tbase, plane, a.w, tx, nclass);

    float s[C], t[KD ? C : 1];
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        s[c] = cs.value(ty, 0, c);
        if constexpr (KD) t[c] = ct.value(ty, 0, c);
      }
    float ms = s[0], mt = KD ? t[0] : 0.f;
#pragma unroll
    for (int c = 1; c < C; ++c)
      if (!PAD || c < nclass) {
        ms = fmaxf(ms, s[c]);
        if constexpr (KD) mt = fmaxf(mt, t[c]);
      }
    float Ss = 0.f, St = 0.f, cross = 0.f, dtgt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        const float d = s[c] - ms;
        const float es = fast_exp(d);
        Ss += es;
        if constexpr (KD) {
          const float e = fast_exp(t[c] - mt);
          St += e;
          cross = fmaf(e, d, cross);
          t[c] = e;
        }
        if constexpr (CE) {
          if (tgt == c) dtgt = d;
        }
        s[c] = es;
      }
    float inv_t = 0.f;
    if constexpr (KD) inv_t = 1.0f / St;
    float wt = 1.f;
    bool counted = false, valid = false;
    if constexpr (CE) {
      counted = ce_img && in_range && tgt >= 0;                     // loss.py:56  mask = target >= 0
      valid = counted && tgt < nclass;                              // 255 (any id >= C) is ignored by nll_loss
      if (valid && a.weight != nullptr) wt = __ldg(a.weight + tgt);
    }
    if constexpr (LOSS) {
      const float lse = fast_log(Ss);
      if constexpr (KD) {
        if (in_range) acc_kd += wkd * (lse - cross * inv_t);
      }
      if constexpr (CE) {
        if (valid) acc_ce += wt * (lse - dtgt);
        if (counted) acc_cnt += 1.f;
      }
    }
    if constexpr (GRAD) {
      const float cpx = (CE && valid) ? wt * cce : 0.f;
      const float ga = (ckd + cpx) / Ss;
      const float gb = ckd * inv_t;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          float g = ga * s[c];
          if constexpr (KD) g = fmaf(-gb, t[c], g);
          if constexpr (CE) g -= (tgt == c) ? cpx : 0.f;
          Gt[c] = fmaf(ty.l0, g, Gt[c]);
          Gb[c] = fmaf(ty.l1, g, Gb[c]);
        }
    }
    tgt = tgt_next;
  }

  if constexpr (GRAD) {
    if (cur_i1 == cur_i0) {                                  // clamped at the last source row: both taps hit it
#pragma unroll
      for (int c = 0; c < C; ++c) Gt[c] += Gb[c];
      flush(Gt, cur_i0 - ylo);
    } else {
      flush(Gt, cur_i0 - ylo);
      flush(Gb, cur_i1 - ylo);
    }
  }

  if constexpr (LOSS) {
    __shared__ float red[kLuBlock / 32];
    __shared__ bool is_last;
    const float bk = block_sum<kLuBlock>(acc_kd, red);
    __syncthreads();
    const float bc = block_sum<kLuBlock>(acc_ce, red);
    __syncthreads();
    const float bn = block_sum<kLuBlock>(acc_cnt, red);
    if (tid == 0) {
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
        if (KD && a.loss_kd) a.loss_kd[0] = (float)(dr[0][0] * (double)a.inv_count_kd);
        if constexpr (CE) {
          const float cnt = (float)dr[2][0], tot = (float)dr[1][0];
          if (a.loss_ce) a.loss_ce[0] = a.size_average ? tot / cnt : tot;   // fp32 division like `loss /= mask.data.sum()`
          if (a.denom_out) a.denom_out[0] = cnt;
        }
        *a.ticket = 0;                                       // leave the workspace ready for the next launch
      }
    }
  }
}

// dlow[n,c,y,x] = sum of the scratch patches that cover (y,x): strips in increasing ky, tiles in increasing kx.
template <int C, bool PAD>
__global__ void __launch_bounds__(128)
loss_up_gather_kernel(const float* __restrict__ scratch, float* __restrict__ dlow, int nclass, int n, int h, int w, int H,
                      int W, float sh, float sw, int ry, int R, int K, int SX, int SY) {
  const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (idx >= (int64_t)n * h * w) return;
  const int x = (int)(idx % w);
  const int y = (int)((idx / w) % h);
  const int img = (int)(idx / ((int64_t)h * w));
  float acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = 0.f;
  // strips / tiles that can hold (y, x): output rows with i0 in {y-1, y} lie in [(y-1)/sh, (y+1)/sh]
  int ky0 = 0, ky1 = SY - 1, kx0 = 0, kx1 = SX - 1;
  if (sh > 0.f) {
    ky0 = max(0, (int)(((float)y - 1.f) / sh) / ry - 1);
    ky1 = min(SY - 1, (int)(((float)y + 1.f) / sh) / ry + 1);
  }
  if (sw > 0.f) {
    kx0 = max(0, (int)(((float)x - 1.f) / sw) / kLuBlock - 1);
    kx1 = min(SX - 1, (int)(((float)x + 1.f) / sw) / kLuBlock + 1);
  }
  for (int ky = ky0; ky <= ky1; ++ky) {
    const int ylo = bilinear_tap(sh, ky * ry, h).i0;
    const int yhi = bilinear_tap(sh, min((ky + 1) * ry, H) - 1, h).i1;
    if (y < ylo || y > yhi) continue;
    for (int kx = kx0; kx <= kx1; ++kx) {
      const int xlo = bilinear_tap(sw, kx * kLuBlock, w).i0;
      const int xhi = bilinear_tap(sw, min((kx + 1) * kLuBlock, W) - 1, w).i1;
      if (x < xlo || x > xhi) continue;
      const int64_t cta = ((int64_t)img * SY + ky) * SX + kx;
      const float* src = scratch + ((cta * R + (y - ylo)) * nclass) * K + (x - xlo);
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) acc[c] += __ldcg(src + c * K);
    }
  }
  float* dst = dlow + ((int64_t)img * nclass * h + y) * w + x;
#pragma unroll
  for (int c = 0; c < C; ++c)
    if (!PAD || c < nclass) dst[(int64_t)c * h * w] = acc[c];
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
  DIGA_DISPATCH_C(C, {
    loss_up_kernel<kC, kPad, KD, CE, LOSS, GRAD><<<grid, kLuBlock, 0, st>>>(a);
    DIGA_CHECK_LAUNCH("loss_up_kernel");
    if (GRAD) {
      const int64_t total = (int64_t)a.n * a.h * a.w;
      loss_up_gather_kernel<kC, kPad><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
          a.scratch, dlow, a.nclass, a.n, a.h, a.w, a.H, a.W, a.sh, a.sw, p.ry, p.R, p.K, p.SX, p.SY);
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
                       void* workspace, diga_stream_t stream) {
  using namespace diga;
  if (int rc = check_common("kd_up_fwd_bwd", student_low, n2, C, h, w, H, W, workspace)) return rc;
  DIGA_REQUIRE(teacher_low && loss_out && dstudent_low && (n2 % 2) == 0, DIGA_ERR_INVALID,
               "kd_up_fwd_bwd: teacher, loss_out, dstudent_low and an even batch are required");
  DIGA_REQUIRE(aligned(teacher_low, 4) && aligned(dstudent_low, 4), DIGA_ERR_MISALIGNED, "kd_up_fwd_bwd: misaligned pointer");
  const LossUpPlan p = make_plan(n2, C, h, w, H, W);
  LossUpArgs a = fill_args(p, workspace, teacher_low, student_low, nullptr, nullptr, n2, 0, C, h, w, H, W, scale, 1);
  a.up_kd_host = upstream_host;
  a.loss_kd = loss_out;
  return launch_loss_up<true, false, true, true>(a, p, C, dstudent_low, (cudaStream_t)stream);
}

}  // extern "C"
