// a1 — symmetric knowledge-distillation loss, forward / backward / fused.
// Replaces util/loss.py:125-143 of the reference (see include/diga_b200.h).
//
// Data layout: teacher, student [2B, C, hw] fp32 (NCHW, class axis strided by hw).  A thread owns
// VEC adjacent pixels of one image and issues one VEC-wide load per class plane for each tensor:
// every warp request is a contiguous 32*VEC*4-byte line, and the whole class-axis reduction
// (max, sum exp, sum p*log q) stays in that thread's registers.  Nothing is staged in shared memory
// because no byte is used twice.  Algorithmic traffic: fwd 2*C*4 B/px, bwd 3*C*4 B/px.
//
// Per pixel, with e_c = exp(t_c - max t), S_t = sum e_c, d_c = s_c - max s, S_s = sum exp(d_c):
//   loss_px = log S_s - (sum_c e_c d_c) / S_t              ( = -sum_c softmax(t)_c log_softmax(s)_c )
//   dL/ds_c = w / (B hw) * (exp(d_c)/S_s - e_c/S_t),  w = scale for student view 0, 1 for view 1.
#include "common.cuh"

namespace diga {

constexpr int kKdMaxPartials = 4096;
struct KdWorkspace {
  unsigned int ticket;
  unsigned int pad[3];
  double partial[kKdMaxPartials];
};

// Register budget: the two class vectors (2*C*VEC floats) must stay in registers.  VEC=2 needs 128
// registers (512 threads per SM).  VEC=4 would need >255 and spill, so 64-bit accesses are the widest
// this two-tensor kernel uses (ptxas -v); VEC=1 is the path for odd plane sizes.
constexpr int kd_min_blocks(int vec, int block) { return 512 / block; }

template <int C, bool PAD, int VEC, int BLOCK, bool LOSS, bool GRAD>
__global__ void __launch_bounds__(BLOCK, kd_min_blocks(VEC, BLOCK))
kd_kernel(const float* __restrict__ tea, const float* __restrict__ stu, float* __restrict__ dstu, int nclass,
          int64_t B, int64_t hw, float scale, float inv_count, const float* __restrict__ upstream_dev,
          float upstream_host, float* __restrict__ loss_out, KdWorkspace* __restrict__ ws) {
  const int64_t groups_per_img = hw / VEC;
  const int64_t total = 2 * B * groups_per_img;
  const int64_t plane = hw;
  float gcoef = 0.f;
  if constexpr (GRAD) gcoef = (upstream_dev != nullptr ? __ldg(upstream_dev) : upstream_host) * inv_count;
  float acc = 0.f;

  for (int64_t gidx = (int64_t)blockIdx.x * BLOCK + threadIdx.x; gidx < total; gidx += (int64_t)gridDim.x * BLOCK) {
    const int64_t n = gidx / groups_per_img;
    const int64_t p = (gidx - n * groups_per_img) * VEC;
    const int64_t nt = n < B ? n + B : n - B;   // the *other* view supervises this one
    const float w = n < B ? scale : 1.f;        // teacher view 1 -> student view 0 is the scaled pair
    const float* tp = tea + nt * nclass * plane + p;
    const float* sp = stu + n * nclass * plane + p;

    Vec<VEC> t[C], s[C];
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) t[c] = ld_stream<VEC>(tp + c * plane);
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (!PAD || c < nclass) s[c] = ld_stream<VEC>(sp + c * plane);

    float inv_t[VEC], inv_s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float mt = t[0].v[v], ms = s[0].v[v];
#pragma unroll
      for (int c = 1; c < C; ++c)
        if (!PAD || c < nclass) {
          mt = fmaxf(mt, t[c].v[v]);
          ms = fmaxf(ms, s[c].v[v]);
        }
      float St = 0.f, Ss = 0.f, cross = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          const float e = fast_exp(t[c].v[v] - mt);
          const float d = s[c].v[v] - ms;
          const float es = fast_exp(d);
          St += e;
          Ss += es;
          cross = fmaf(e, d, cross);
          t[c].v[v] = e;
          if constexpr (GRAD) s[c].v[v] = es;
        }
      inv_t[v] = 1.0f / St;
      if constexpr (LOSS) acc += w * (fast_log(Ss) - cross * inv_t[v]);
      if constexpr (GRAD) inv_s[v] = 1.0f / Ss;
    }

    if constexpr (GRAD) {
      const float gw = gcoef * w;
      float* dp = dstu + n * nclass * plane + p;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (!PAD || c < nclass) {
          Vec<VEC> o;
#pragma unroll
          for (int v = 0; v < VEC; ++v) o.v[v] = gw * (s[c].v[v] * inv_s[v] - t[c].v[v] * inv_t[v]);
          st_stream<VEC>(dp + c * plane, o);
        }
    }
  }

  if constexpr (LOSS) {
    __shared__ float red[BLOCK / 32];
    __shared__ bool is_last;
    const float bsum = block_sum<BLOCK>(acc, red);
    if (threadIdx.x == 0) {
      ws->partial[blockIdx.x] = (double)bsum;
      __threadfence();
      const unsigned int t = atomicAdd(&ws->ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      // Deterministic second stage: fixed partition of the per-CTA partials, fp64.
      __threadfence();
      __shared__ double dred[BLOCK];
      double a = 0.0;
      for (unsigned int i = threadIdx.x; i < gridDim.x; i += BLOCK) a += __ldcg(&ws->partial[i]);
      dred[threadIdx.x] = a;
      __syncthreads();
      for (int o = BLOCK / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) dred[threadIdx.x] += dred[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        loss_out[0] = (float)(dred[0] * (double)inv_count);
        ws->ticket = 0;  // leave the workspace ready for the next launch
      }
    }
  }
}

int tunable(const char* name, int dflt);

template <int C, bool PAD, int VEC, int BLOCK, bool LOSS, bool GRAD>
static int launch_kd(const float* tea, const float* stu, float* dstu, int nclass, int64_t B, int64_t hw, float scale,
                     const float* up_dev, float up_host, float* loss_out, KdWorkspace* ws, cudaStream_t st) {
  auto kern = kd_kernel<C, PAD, VEC, BLOCK, LOSS, GRAD>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, BLOCK, 0);
    blocks_per_sm = b > 0 ? b : 1;
  }
  const int64_t total = 2 * B * (hw / VEC);
  int64_t grid = (total + BLOCK - 1) / BLOCK;
  // Grid = resident CTAs x waves.  Measured on B200 (tools/tune.py, profiles/tune_r01.md): the read-only forward
  // is fastest as one persistent wave (1.00 of the copy peak); the kernels that also write 76 B/px want ~8 waves
  // of shorter-lived CTAs (0.94 -> 1.00).
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm * tunable(GRAD ? "kd_waves_bwd" : "kd_waves_fwd", GRAD ? 8 : 1);
  if (grid > cap) grid = cap;
  if (grid > kKdMaxPartials) grid = kKdMaxPartials;
  if (grid < 1) grid = 1;
  const float inv_count = (float)(1.0 / ((double)B * (double)hw));
  kern<<<(unsigned)grid, BLOCK, 0, st>>>(tea, stu, dstu, nclass, B, hw, scale, inv_count, up_dev, up_host, loss_out, ws);
  DIGA_CHECK_LAUNCH("kd_kernel");
  return DIGA_OK;
}

template <bool LOSS, bool GRAD>
static int dispatch_kd(const float* tea, const float* stu, float* dstu, int64_t n2, int64_t C, int64_t hw, float scale,
                       const float* up_dev, float up_host, float* loss_out, void* workspace, cudaStream_t st) {
  DIGA_REQUIRE(tea && stu, DIGA_ERR_INVALID, "kd: null input");
  DIGA_REQUIRE(n2 > 0 && (n2 % 2) == 0, DIGA_ERR_INVALID, "kd: batch %lld must be even and positive (two views)",
               (long long)n2);
  DIGA_REQUIRE(C >= 1 && C <= DIGA_MAX_CLASSES, DIGA_ERR_INVALID, "kd: C=%lld outside [1,%d]", (long long)C,
               DIGA_MAX_CLASSES);
  DIGA_REQUIRE(hw > 0, DIGA_ERR_INVALID, "kd: empty plane");
  DIGA_REQUIRE(!LOSS || (loss_out && workspace), DIGA_ERR_INVALID, "kd: loss_out/workspace required");
  DIGA_REQUIRE(!GRAD || dstu, DIGA_ERR_INVALID, "kd: dstudent required");
  DIGA_REQUIRE(aligned(tea, 4) && aligned(stu, 4) && (!GRAD || aligned(dstu, 4)), DIGA_ERR_MISALIGNED,
               "kd: pointers must be 4-byte aligned");
  const int64_t B = n2 / 2;
  int vec = tunable("kd_vec", 2);
  const bool a8 = aligned(tea, 8) && aligned(stu, 8) && (!GRAD || aligned(dstu, 8));
  if (vec != 1 && !((hw % 2) == 0 && a8)) vec = 1;
  const int block = tunable((LOSS && GRAD) ? "kd_block_fused" : "kd_block", (LOSS && GRAD) ? 128 : 256);
  KdWorkspace* ws = reinterpret_cast<KdWorkspace*>(workspace);
#define DIGA_KD_GO(V, BL) \
  return launch_kd<kC, kPad, V, BL, LOSS, GRAD>(tea, stu, dstu, (int)C, B, hw, scale, up_dev, up_host, loss_out, ws, st)
  DIGA_DISPATCH_C(C, {
    if (block == 128) {
      if (vec == 2) DIGA_KD_GO(2, 128);
      DIGA_KD_GO(1, 128);
    } else {
      if (vec == 2) DIGA_KD_GO(2, 256);
      DIGA_KD_GO(1, 256);
    }
  });
#undef DIGA_KD_GO
  return DIGA_OK;
}

}  // namespace diga

extern "C" {

size_t diga_kd_workspace_bytes(void) { return sizeof(diga::KdWorkspace); }

int diga_kd_fwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                float* loss_out, void* workspace, diga_stream_t stream) {
  return diga::dispatch_kd<true, false>(teacher, student, nullptr, n2, C, hw, scale, nullptr, 0.f, loss_out, workspace,
                                        (cudaStream_t)stream);
}

int diga_kd_bwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                const float* upstream, float* dstudent, diga_stream_t stream) {
  DIGA_REQUIRE(upstream, DIGA_ERR_INVALID, "kd_bwd: upstream (device scalar) required");
  return diga::dispatch_kd<false, true>(teacher, student, dstudent, n2, C, hw, scale, upstream, 0.f, nullptr, nullptr,
                                        (cudaStream_t)stream);
}

int diga_kd_fwd_bwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                    float upstream_host, float* loss_out, float* dstudent, void* workspace, diga_stream_t stream) {
  return diga::dispatch_kd<true, true>(teacher, student, dstudent, n2, C, hw, scale, nullptr, upstream_host, loss_out,
                                       workspace, (cudaStream_t)stream);
}

}  // extern "C"
