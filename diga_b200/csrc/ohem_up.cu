// OhemCrossEntropy (util/loss.py:65-122 of the reference; the segmentation loss of the Synthia tree,
// Synthia/train_DiGA_syn2city_*.py) evaluated straight from the stride-8 scores: `_ohem_forward` itself up-samples the
// score to the label size with bilinear `align_corners=True` when the two differ (:91-96), so the low-resolution call is the
// reference's own signature and the fused up-sampling of csrc/loss_up.cu applies unchanged.
//
//   pred      = softmax(score)[target]                for pixels with target != ignore_label            (:97-104)
//   min_value = sorted(pred)[min(min_kept, M - 1)]    M = number of such pixels                          (:105-106)
//   threshold = max(min_value, thresh)                                                                   (:107)
//   loss      = mean over { pixels with pred < threshold } of  -w[target] * log_softmax(score)[target]   (:109-111)
//
// The reference sorts all M probabilities to read one order statistic.  Here:
//   1. ohem_pred_kernel walks the output pixels like the loss kernels (LerpColumn) and stores pred and the per-pixel loss
//      (8 B/px);
//   2. an exact radix select over the float bits of pred (non-negative floats order like their bit patterns): three
//      histogram passes over the 4 B/px buffer (12 + 12 + 8 bits) with a one-CTA pick after each;
//   3. ohem_sum_kernel adds the kept losses and counts them (two-stage, deterministic);
//   4. the backward is the CE gradient pass of csrc/loss_up.cu restricted to the kept pixels (diga_ohem_up_bwd).
#include "lerp_column.cuh"

namespace diga {

int tunable(const char* name, int dflt);

constexpr int kOhBlock = 128;
constexpr int kOhBins = 4096;
constexpr int kOhMaxPartials = 2048;

struct OhemState {
  unsigned int hist[kOhBins];
  unsigned int prefix;        // key bits fixed so far
  unsigned int k;             // rank still to resolve inside the current prefix
  unsigned int M;             // pixels with target != ignore_label
  unsigned int empty;         // M == 0
  unsigned int ticket;
  unsigned int pad[3];
  double part_sum[kOhMaxPartials];
  double part_cnt[kOhMaxPartials];
};

template <int C, bool PAD>
__global__ void __launch_bounds__(kOhBlock, 4)
ohem_pred_kernel(const float* __restrict__ score, const int64_t* __restrict__ target, const float* __restrict__ weight, int nclass,
                 int h, int w, int H, int W, float sh, float sw, int ry, int ignore_label, float* __restrict__ pred,
                 float* __restrict__ losspx) {
  extern __shared__ __align__(16) int2 ytab[];
  const int n = blockIdx.z, tid = threadIdx.x;
  const int X0 = blockIdx.x * kOhBlock, X = X0 + tid;
  const bool in_range = X < W;
  const int Y0 = blockIdx.y * ry, Yend = min(Y0 + ry, H);
  const int64_t plane = (int64_t)h * w;
  const Tap tx = bilinear_tap(sw, in_range ? X : W - 1, w);
  const bool clamped = tx.i1 == tx.i0, pair = w > 1;
  const int kc = (clamped && pair) ? tx.i0 - 1 : tx.i0;
  const float l0s = (clamped ? (pair ? 0.f : tx.l0 + tx.l1) : tx.l0) * kLog2e;
  const float l1s = (clamped ? (pair ? tx.l0 + tx.l1 : 0.f) : tx.l1) * kLog2e;
  const int ylo = bilinear_tap(sh, Y0, h).i0;
  const float* scol = score + (int64_t)n * nclass * plane + (int64_t)ylo * w + kc;
  for (int i = tid; i < Yend - Y0; i += kOhBlock) {
    const Tap t = bilinear_tap(sh, Y0 + i, h);
    ytab[i] = make_int2((t.i0 - ylo) | ((t.i1 - t.i0) << 16), __float_as_int(t.l1));
  }
  __syncthreads();
  LerpColumn<C, PAD> cs;
  int cur_r0 = -1;
  const int64_t col = ((int64_t)n * H) * W + (in_range ? X : W - 1);
  for (int Y = Y0; Y < Yend; ++Y) {
    const int64_t tgt = ld_stream_i64(target + col + (int64_t)Y * W);
    const int2 yt = ytab[Y - Y0];
    const int r0 = yt.x & 0xffff, r1 = r0 + (yt.x >> 16);
    const float yl1 = __int_as_float(yt.y);
    if (r0 != cur_r0) {
      cs.enter(cur_r0 < 0, scol, w, plane, r0, r1, pair, l0s, l1s, nclass);
      cur_r0 = r0;
    }
    const int t32 = (tgt >= 0 && tgt < nclass && tgt != ignore_label) ? (int)tgt : -1;
    float S4[4] = {0.f, 0.f, 0.f, 0.f};
    float dtgt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) {
        const float v = cs.value(yl1, c);
        S4[c & 3] += fast_ex2(v);
        if (t32 == c) dtgt = v;
      }
    float S = (S4[0] + S4[1]) + (S4[2] + S4[3]);
    if (S < 0x1p-60f) {                                            // cell reference too loose (see LerpColumn): exact max
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) m = fmaxf(m, cs.value(yl1, c));
      S = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) S += fast_ex2(cs.value(yl1, c) - m);
      dtgt -= m;
    }
    if (in_range) {
      const int64_t o = col + (int64_t)Y * W;
      if (t32 >= 0) {
        const float lse2 = fast_lg2(S);
        const float wt = weight != nullptr ? __ldg(weight + t32) : 1.f;
        pred[o] = fminf(fast_ex2(dtgt - lse2), 1.0f);              // softmax(score)[target]
        losspx[o] = wt * (lse2 - dtgt) * kLn2;                     // nn.CrossEntropyLoss(weight, reduction='none')
      } else {
        pred[o] = -1.f;                                            // masked out (:99)
        losspx[o] = 0.f;
      }
    }
  }
}

// key of a kept probability = its float bits (>= 0); level 0: bits 31..20, level 1: 19..8, level 2: 7..0
__device__ __forceinline__ bool ohem_key_bin(float p, int level, unsigned int prefix, unsigned int* bin) {
  if (!(p >= 0.f)) return false;
  const unsigned int key = __float_as_uint(p);
  if (level == 0) {
    *bin = key >> 20;
    return true;
  }
  if (level == 1) {
    *bin = (key >> 8) & 0xfffu;
    return (key >> 20) == prefix;
  }
  *bin = key & 0xffu;
  return (key >> 8) == prefix;
}

__global__ void __launch_bounds__(256)
ohem_hist_kernel(const float* __restrict__ pred, int64_t total, int level, OhemState* __restrict__ st) {
  __shared__ unsigned int sh[kOhBins];
  for (int i = threadIdx.x; i < kOhBins; i += 256) sh[i] = 0;
  __syncthreads();
  const unsigned int prefix = level == 0 ? 0u : st->prefix;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t rounds = (total + stride - 1) / stride;
  for (int64_t r = 0; r < rounds; ++r) {                          // every lane runs every round (__match_any_sync)
    const int64_t i = r * stride + (int64_t)blockIdx.x * 256 + threadIdx.x;
    unsigned int bin = 0;
    const bool on = i < total && ohem_key_bin(__ldg(pred + i), level, prefix, &bin);
    const unsigned int key = on ? bin : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (on && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kOhBins; i += 256)
    if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
}

// One CTA: find the bin that holds rank k, descend into it, clear the histogram for the next level.
__global__ void __launch_bounds__(256)
ohem_pick_kernel(OhemState* __restrict__ st, int level, unsigned int min_kept, float thresh, float* __restrict__ thr_out) {
  __shared__ unsigned int part[256], incl[256];
  __shared__ unsigned int chosen_bin, chosen_rank, rank_k;
  const int tid = threadIdx.x;
  constexpr int PER = kOhBins / 256;
  unsigned int mine = 0;
  for (int i = 0; i < PER; ++i) mine += st->hist[tid * PER + i];
  part[tid] = mine;
  incl[tid] = mine;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {                                 // inclusive prefix sum of the 256 chunk totals
    const unsigned int add = tid >= o ? incl[tid - o] : 0u;
    __syncthreads();
    incl[tid] += add;
    __syncthreads();
  }
  if (tid == 0) {
    unsigned int k = st->k;
    if (level == 0) {
      const unsigned int M = incl[255];
      st->M = M;
      st->empty = (M == 0);
      k = M == 0 ? 0u : (min_kept < M - 1 ? min_kept : M - 1);       // loss.py:104  min(self.min_kept, pred.numel() - 1)
    }
    rank_k = k;
    chosen_bin = kOhBins - 1;
    chosen_rank = 0;
  }
  __syncthreads();
  {
    const unsigned int k = rank_k, before = incl[tid] - part[tid];
    if (before <= k && k < incl[tid]) {                               // exactly one chunk holds rank k (when M > 0)
      unsigned int run = before;
      int b = tid * PER;
      while (b < tid * PER + PER - 1 && run + st->hist[b] <= k) run += st->hist[b++];
      chosen_bin = (unsigned)b;
      chosen_rank = k - run;
    }
  }
  __syncthreads();
  for (int i = tid; i < kOhBins; i += 256) st->hist[i] = 0;
  if (tid == 0) {
    const unsigned int prefix = level == 0 ? chosen_bin : (level == 1 ? ((st->prefix << 12) | chosen_bin) : ((st->prefix << 8) | chosen_bin));
    st->prefix = prefix;
    st->k = chosen_rank;
    if (level == 2) {
      const float kth = __uint_as_float(prefix);                      // the exact order statistic
      thr_out[0] = st->empty ? thresh : fmaxf(kth, thresh);           // :107
    }
  }
}

__global__ void __launch_bounds__(256)
ohem_sum_kernel(const float* __restrict__ pred, const float* __restrict__ losspx, int64_t total, const float* __restrict__ thr,
                OhemState* __restrict__ st, float* __restrict__ loss_out, float* __restrict__ count_out) {
  __shared__ double rs[8], rc[8];
  __shared__ bool is_last;
  const float t = __ldg(thr);
  double s = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const float p = __ldg(pred + i);
    if (p >= 0.f && p < t) {                                           // :110  pixel_losses[pred < threshold]
      s += (double)__ldg(losspx + i);
      c += 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rc[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) {
      a += rs[i];
      b += rc[i];
    }
    st->part_sum[blockIdx.x] = a;
    st->part_cnt[blockIdx.x] = b;
    __threadfence();
    is_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    __shared__ double fs[256], fc[256];
    double a = 0.0, b = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) {      // fixed partition, fixed order: deterministic
      a += __ldcg(&st->part_sum[i]);
      b += __ldcg(&st->part_cnt[i]);
    }
    fs[threadIdx.x] = a;
    fc[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) {
        fs[threadIdx.x] += fs[threadIdx.x + o];
        fc[threadIdx.x] += fc[threadIdx.x + o];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      loss_out[0] = (float)fs[0] / (float)fc[0];                       // :109 .mean()  (0 / 0 = NaN like torch on an empty selection)
      count_out[0] = (float)fc[0];
      st->ticket = 0;
      st->prefix = 0;
      st->k = 0;
    }
  }
}

}  // namespace diga

extern "C" {

size_t diga_ohem_up_workspace_bytes(void) { return sizeof(diga::OhemState); }

int diga_ohem_up_fwd(const float* score_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h,
                     int64_t w, int64_t H, int64_t W, int64_t ignore_label, float thresh, int64_t min_kept, float* pred,
                     float* losspx, float* loss_out, float* count_out, float* thr_out, void* workspace, diga_stream_t stream) {
  using namespace diga;
  DIGA_REQUIRE(score_low && target && pred && losspx && loss_out && count_out && thr_out && workspace, DIGA_ERR_INVALID,
               "ohem_up_fwd: null pointer");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "ohem_up_fwd: C=%lld outside [1,%d]", (long long)C, DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n >= 1 && n <= 65535 && h >= 1 && w >= 1 && H >= h && W >= w && H < (1 << 24) && W < (1 << 24) &&
                   n * H * W < ((int64_t)1 << 32),
               DIGA_ERR_INVALID, "ohem_up_fwd: needs an up-sampling geometry (H >= h, W >= w) and fewer than 2^32 pixels");
  DIGA_REQUIRE(min_kept >= 0, DIGA_ERR_INVALID, "ohem_up_fwd: negative min_kept");
  DIGA_REQUIRE(aligned(score_low, 4) && aligned(target, 8) && aligned(weight, 4) && aligned(pred, 4) && aligned(losspx, 4) &&
                   aligned(workspace, 8),
               DIGA_ERR_MISALIGNED, "ohem_up_fwd: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  OhemState* state = reinterpret_cast<OhemState*>(workspace);
  const int ry = 32;
  dim3 grid((unsigned)((W + kOhBlock - 1) / kOhBlock), (unsigned)((H + ry - 1) / ry), (unsigned)n);
  DIGA_DISPATCH_C(C, {
    ohem_pred_kernel<kC, kPad><<<grid, kOhBlock, ry * sizeof(int2), st>>>(score_low, target, weight, (int)C, (int)h, (int)w, (int)H,
                                                                         (int)W, bilinear_scale_host(h, H), bilinear_scale_host(w, W),
                                                                         ry, (int)ignore_label, pred, losspx);
  });
  DIGA_CHECK_LAUNCH("ohem_pred_kernel");
  const int64_t total = n * H * W;
  int64_t hg = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (hg > cap) hg = cap;
  if (hg > kOhMaxPartials) hg = kOhMaxPartials;
  const unsigned int kept = (unsigned int)(min_kept > 0xfffffffell ? 0xfffffffell : min_kept);
  for (int level = 0; level < 3; ++level) {
    ohem_hist_kernel<<<(unsigned)hg, 256, 0, st>>>(pred, total, level, state);
    DIGA_CHECK_LAUNCH("ohem_hist_kernel");
    ohem_pick_kernel<<<1, 256, 0, st>>>(state, level, kept, thresh, thr_out);
    DIGA_CHECK_LAUNCH("ohem_pick_kernel");
  }
  ohem_sum_kernel<<<(unsigned)hg, 256, 0, st>>>(pred, losspx, total, thr_out, state, loss_out, count_out);
  DIGA_CHECK_LAUNCH("ohem_sum_kernel");
  return DIGA_OK;
}

}  // extern "C"
