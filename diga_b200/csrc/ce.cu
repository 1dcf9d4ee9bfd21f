// f2 — 2-D cross entropy with ignore label, forward and backward.
// Replaces util/loss.py:48-62 of the reference (`cross_entropy2d`), which permutes the logits to NHWC, boolean-
// gathers them and calls nll_loss: ~6 extra passes over a 76 B/px tensor.  Here: one pass forward
// (4C + 8 B/px), one pass backward (8C + 8 B/px), same thread mapping as the KD kernels (csrc/kd.cu).
//
// Semantics (mirrors the reference exactly):
//   log_p = log_softmax(input, 1);  pixels with target < 0 are dropped;  loss = sum over remaining pixels with
//   target != 255 of -w[target] * log_p[target];  if size_average: loss /= #(target >= 0)   (ignore-255 pixels
//   DO count in the denominator, loss.py:56,60).  Gradient: w[t] * (softmax - onehot(t)) / denom, 0 elsewhere.
// Targets in [C, 255) are outside the reference's domain (nll_loss raises); they are treated as ignored.
#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

constexpr int kCeMaxPartials = 4096;
struct CeWorkspace {
  unsigned int ticket;
  unsigned int pad[3];
  double loss[kCeMaxPartials];
  double count[kCeMaxPartials];
};

template <int C, bool PAD, int VEC, int BLOCK, bool GRAD>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK)
ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, const float* __restrict__ weight, int nclass,
          int64_t n, int64_t hw, int size_average, const float* __restrict__ upstream, const float* __restrict__ denom_in,
          float* __restrict__ dlogits, float* __restrict__ loss_out, float* __restrict__ denom_out, CeWorkspace* __restrict__ ws) {
  const int64_t groups_per_img = hw / VEC;
  const int64_t total = n * groups_per_img;
  float acc_loss = 0.f, acc_cnt = 0.f;
  float g = 0.f;
  if constexpr (GRAD) g = __ldg(upstream) / (size_average ? __ldg(denom_in) : 1.0f);

  for (int64_t gidx = (int64_t)blockIdx.x * BLOCK + threadIdx.x; gidx < total; gidx += (int64_t)gridDim.x * BLOCK) {
    const int64_t img = gidx / groups_per_img;
    const int64_t p = (gidx - img * groups_per_img) * VEC;
    const float* zp = logits + img * nclass * hw + p;
    Vec<VEC> z[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) z[c] = ld_stream<VEC>(zp + c * hw);
    int64_t tgt[VEC];
    if constexpr (VEC == 2) {
      const longlong2 t = ld_stream_i64x2(target + img * hw + p);
      tgt[0] = t.x;
      tgt[1] = t.y;
    } else {
      tgt[0] = ld_stream_i64(target + img * hw + p);
    }
    float coef[VEC], inv_s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float m = z[0].v[v];
#pragma unroll
      for (int c = 1; c < C; ++c)
        if (!PAD || c < nclass) m = fmaxf(m, z[c].v[v]);
      float S = 0.f, dt = 0.f, wt = 1.f;
      const bool counted = tgt[v] >= 0;                              // loss.py:56  mask = target >= 0
      const bool valid = counted && tgt[v] < nclass;                 // 255 (and any id >= C) is ignored by nll_loss
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          const float d = z[c].v[v] - m;
          const float e = fast_exp(d);
          S += e;
          if (tgt[v] == c) dt = d;
          if constexpr (GRAD) z[c].v[v] = e;
        }
      if (weight != nullptr && valid) wt = __ldg(weight + tgt[v]);
      if constexpr (!GRAD) {
        if (valid) acc_loss += wt * (fast_log(S) - dt);              // -log_softmax[target]
        if (counted) acc_cnt += 1.f;
      } else {
        coef[v] = valid ? wt * g : 0.f;
        inv_s[v] = 1.0f / S;
      }
    }
    if constexpr (GRAD) {
      float* dp = dlogits + img * nclass * hw + p;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          Vec<VEC> o;
#pragma unroll
          for (int v = 0; v < VEC; ++v) o.v[v] = coef[v] * (z[c].v[v] * inv_s[v] - (tgt[v] == c ? 1.f : 0.f));
          st_stream<VEC>(dp + c * hw, o);
        }
    }
  }

  if constexpr (!GRAD) {
    __shared__ float red[BLOCK / 32];
    __shared__ bool is_last;
    const float bl = block_sum<BLOCK>(acc_loss, red);
    __syncthreads();
    const float bc = block_sum<BLOCK>(acc_cnt, red);
    if (threadIdx.x == 0) {
      ws->loss[blockIdx.x] = (double)bl;
      ws->count[blockIdx.x] = (double)bc;
      __threadfence();
      is_last = (atomicAdd(&ws->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      __shared__ double dl[BLOCK], dc[BLOCK];
      double a = 0.0, b = 0.0;
      for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK) {
        a += __ldcg(&ws->loss[i]);
        b += __ldcg(&ws->count[i]);
      }
      dl[threadIdx.x] = a;
      dc[threadIdx.x] = b;
      __syncthreads();
      for (int o = BLOCK / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
          dl[threadIdx.x] += dl[threadIdx.x + o];
          dc[threadIdx.x] += dc[threadIdx.x + o];
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        const float cnt = (float)dc[0];
        const float tot = (float)dl[0];
        loss_out[0] = size_average ? tot / cnt : tot;      // fp32 division like `loss /= mask.data.sum()`
        denom_out[0] = cnt;
        ws->ticket = 0;
      }
    }
  }
}

template <int C, bool PAD, int VEC, bool GRAD>
static int launch_ce(const float* logits, const int64_t* target, const float* weight, int nclass, int64_t n, int64_t hw,
                     int size_average, const float* upstream, const float* denom_in, float* dlogits, float* loss_out,
                     float* denom_out, CeWorkspace* ws, cudaStream_t st) {
  constexpr int BLOCK = 256;
  auto kern = ce_kernel<C, PAD, VEC, BLOCK, GRAD>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BLOCK, 0);
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int64_t total = n * (hw / VEC);
  int64_t grid = (total + BLOCK - 1) / BLOCK;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm * tunable(GRAD ? "ce_waves_bwd" : "ce_waves_fwd", GRAD ? 8 : 1);
  if (grid > cap) grid = cap;
  if (grid > kCeMaxPartials) grid = kCeMaxPartials;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, BLOCK, 0, st>>>(logits, target, weight, nclass, n, hw, size_average, upstream, denom_in, dlogits, loss_out,
                                         denom_out, ws);
  DIGA_CHECK_LAUNCH("ce_kernel");
  return DIGA_OK;
}

template <bool GRAD>
static int dispatch_ce(const float* logits, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t hw,
                       int size_average, const float* upstream, const float* denom_in, float* dlogits, float* loss_out,
                       float* denom_out, void* workspace, cudaStream_t st) {
  DIGA_REQUIRE(logits && target, DIGA_ERR_INVALID, "cross_entropy2d: null input");
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "cross_entropy2d: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(n > 0 && hw > 0, DIGA_ERR_INVALID, "cross_entropy2d: empty input");
  DIGA_REQUIRE(GRAD ? (upstream && dlogits && (!size_average || denom_in)) : (loss_out && denom_out && workspace), DIGA_ERR_INVALID,
               "cross_entropy2d: missing output / workspace");
  DIGA_REQUIRE(aligned(logits, 4) && aligned(target, 8) && aligned(dlogits, 4) && aligned(weight, 4), DIGA_ERR_MISALIGNED,
               "cross_entropy2d: misaligned pointer");
  const bool v2 = (hw % 2) == 0 && aligned(logits, 8) && aligned(target, 16) && aligned(dlogits, 8) && tunable("ce_vec", 2) == 2;
  CeWorkspace* ws = reinterpret_cast<CeWorkspace*>(workspace);
  DIGA_DISPATCH_C(C, {
    if (v2)
      return launch_ce<kC, kPad, 2, GRAD>(logits, target, weight, (int)C, n, hw, size_average, upstream, denom_in, dlogits, loss_out,
                                          denom_out, ws, st);
    return launch_ce<kC, kPad, 1, GRAD>(logits, target, weight, (int)C, n, hw, size_average, upstream, denom_in, dlogits, loss_out,
                                        denom_out, ws, st);
  });
  return DIGA_OK;
}

}  // namespace diga

extern "C" {

size_t diga_ce_workspace_bytes(void) { return sizeof(diga::CeWorkspace); }

int diga_cross_entropy2d_fwd(const float* logits, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t hw,
                             int size_average, float* loss_out, float* denom_out, void* workspace, diga_stream_t stream) {
  return diga::dispatch_ce<false>(logits, target, weight, n, C, hw, size_average, nullptr, nullptr, nullptr, loss_out, denom_out,
                                  workspace, (cudaStream_t)stream);
}

int diga_cross_entropy2d_bwd(const float* logits, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t hw,
                             int size_average, const float* upstream, const float* denom, float* dlogits, diga_stream_t stream) {
  return diga::dispatch_ce<true>(logits, target, weight, n, C, hw, size_average, upstream, denom, dlogits, nullptr, nullptr, nullptr,
                                 (cudaStream_t)stream);
}

}  // extern "C"
