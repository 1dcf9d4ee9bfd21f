// a5 tensor-core path — feature-by-prototype distance on tcgen05 (5th-gen tensor cores, TMEM accumulators).
// Replaces Class_Features.feat_centroid_distance / get_centroid_weight (calc_centroids.py:166-176 of the
// reference) for D % 32 == 0, D >= 256 (256, 512, 2048 in the reference's trees).
//
//   dist^2[p,c] = ||f_p||^2 - 2 f_p.c_c + ||c_c||^2 ;  the only contraction is  f_p.c_c  (HW x D by D x C).
//
// The kernel is HBM-bound (D*4 bytes per feature pixel against 2*D*C flops: 9.4 flop/B), so everything is
// organised around streaming the feature map exactly once at full bandwidth; the tensor pipe is used because
// fp32 CUDA cores would sit at their own roof (77.8 kFLOP per 8 KB pixel) — not to chase tensor utilisation.
//
// Precision: fp32-faithful 3xTF32.  x = hi + lo with hi = tf32(x) (round to nearest), lo = x - hi; the product
// is hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM (the dropped lo*lo term is 2^-22 relative).  ||f||^2 is
// accumulated on the CUDA cores in fp32 from the same values; ||c||^2 is precomputed.
//
// Data flow per CTA (one CTA per SM, persistent over 128-pixel tiles; pixels are the MMA M dimension):
//   loader threads  : TMA tensor copies (cp.async.bulk.tensor.2d) of 32 channel rows x 128 pixels of the NCHW
//                     feature map into an 8-stage shared-memory ring, plus bulk copies of the pre-split centroid
//                     tiles.  h*w is odd for the reference's feature maps (65x129), so a channel row's pitch is
//                     not a multiple of 16 B and cannot be a TMA stride; the tensor map therefore views the
//                     features as [N*D/4 groups][4*h*w] (pitch 16*h*w bytes, always legal) and each chunk is
//                     four boxes of [8 groups x 132 px], one per row-in-group r, starting at the 16-byte aligned
//                     element at or below r*h*w + p0; the converters read with the 0..3 float shift.
//   converter warps : read their pixel's 32 channel values from shared memory (conflict-free), accumulate the
//                     squared norm, split hi/lo and write both as the A operand into TENSOR MEMORY with
//                     tcgen05.st — the [pixel x channel] transpose NCHW needs happens here for free, so the
//                     MN-major feature tile never has to be re-read from shared memory by the MMA.
//   MMA thread      : tcgen05.mma.kind::tf32, A from TMEM (TS form), B = centroid tile [32 x 8] from shared
//                     memory (K-major, no swizzle), D = [128 x 32] fp32 in TMEM; 3 MMAs per 8 channels.
//   epilogue warps  : tcgen05.ld the accumulator, dist = sqrt(max(n_f + n_c - 2 dot, 0)), softmax(-dist),
//                     coalesced stores.  Two accumulators let the epilogue of tile t overlap the MMAs of t+1.
// All hand-offs are mbarriers (full/empty rings); tcgen05.commit releases operand stages.
#include <cuda.h>

#include "common.cuh"

namespace diga {

int tunable(const char* name, int dflt);

namespace umma {

constexpr int TILE_M = 128;   // pixels per tile = MMA M
constexpr int NPAD = 32;      // classes padded to the MMA N
constexpr int KC = 32;        // channels per pipeline chunk
constexpr int KSTEP = 8;      // channels per tcgen05.mma.kind::tf32
constexpr int SA = 8;         // shared-memory feature stages (135 KB in flight per SM)
constexpr int CPS = 2;        // chunks per operand stage: the MMA warp hands over 64 channels at a time, which halves its
                              // per-chunk wait / commit overhead (it was the critical path at 32)
constexpr int ST = 3;         // operand stages: TMEM A operand (3 x 128 columns) paired with a shared-memory centroid stage
constexpr int SB = ST;        // (one full/empty barrier pair per stage serves both operands)
constexpr int NACC = 2 * NPAD;   // accumulator columns per tile: [hi*hi + lo*hi | hi*lo]; 6 x 64 + 2 x 64 = 512 columns
constexpr int NWG = 4;        // converter warpgroups; chunk c is converted by warpgroup c % NWG
constexpr int BOX_G = KC / 4;                           // channel groups (of 4 rows) per TMA box
constexpr int BOX_W = TILE_M + 4;                       // 128 pixels + the 0..3 floats below a 16-byte aligned start
constexpr int A_BOX_FLOATS = BOX_G * BOX_W;             // 1056
constexpr int A_STAGE_FLOATS = 4 * A_BOX_FLOATS;
constexpr int A_STAGE_BYTES = A_STAGE_FLOATS * 4;       // 16896
constexpr int B_TILE_BYTES = NPAD * KSTEP * 4;          // 1024: one [32 x 8] tf32 operand tile
constexpr int B_KSTEP_BYTES = 2 * B_TILE_BYTES;         // hi tile + lo tile
constexpr int B_STAGE_BYTES = CPS * (KC / KSTEP) * B_KSTEP_BYTES;   // 16384
constexpr int CONV_WARPS = 4 * NWG, EPI_WARPS = 4;
constexpr int WARP_MMA = CONV_WARPS + EPI_WARPS, WARP_LOAD = WARP_MMA + 1, WARP_LOAD_B = WARP_MMA + 2;
constexpr int THREADS = (WARP_LOAD_B + 1) * 32;         // 736
constexpr int TMEM_COLS = 512;
constexpr int TMEM_STAGE_COLS = CPS * 2 * KC;           // 128: per chunk 32 hi + 32 lo columns
constexpr int TMEM_ACC_COL = ST * TMEM_STAGE_COLS;      // accumulators behind the A stages

struct Barriers {
  uint64_t a_full[SA], a_empty[SA];
  uint64_t op_full[ST], op_empty[ST];     // A operand in TMEM (4 converter warps) + centroid tile in smem (TMA tx)
  uint64_t acc_full[2], acc_empty[2], norm_full[2];
};

constexpr int SMEM_A = 0;
constexpr int SMEM_B = SMEM_A + SA * A_STAGE_BYTES;
constexpr int SMEM_NORM = SMEM_B + SB * B_STAGE_BYTES;          // [2 acc][NWG][128] floats
constexpr int SMEM_CNORM = SMEM_NORM + 2 * NWG * TILE_M * 4;    // [32] floats
constexpr int SMEM_BAR = SMEM_CNORM + NPAD * 4;
constexpr int SMEM_TMEM_PTR = SMEM_BAR + (int)sizeof(Barriers);
constexpr int SMEM_TOTAL = SMEM_TMEM_PTR + 16;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Blocking wait with a watchdog: a protocol bug must end in a trap (reported as a CUDA error), never in a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// Optional in-kernel profile (umma_debug bit 5): cycles each role spends blocked on each barrier.
struct Prof {
  long long* out;
  bool on;
  long long acc[8];
  __device__ __forceinline__ void wait(int slot, uint64_t* bar, uint32_t parity) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc[slot] += clock64() - t0;
  }
};
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp.  Single-thread roles run their loops with the WHOLE warp (warp-uniform control flow
// and operands, so descriptors live in uniform registers) and predicate only the issue instructions on this; putting
// the loop inside `if (lane == 0)` instead makes ptxas wrap every tcgen05.mma in an ELECT/R2UR waterfall loop
// (~90 cycles per MMA issued, measured), which made the issuing thread the bottleneck of the whole kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, cta_group::1
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): core matrices of 8 rows x 16 B;
// LBO = byte distance between the two 16-byte K halves of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  constexpr uint64_t LBO = (2 * NPAD / 8) * 128;   // 1024: [k-half][8 row-groups: b_hi rows 0..31, b_lo rows 32..63][8 rows][16 B]
  constexpr uint64_t SBO = 128;
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((LBO >> 4) << 16) | ((SBO >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, M = 128, N = 32 or 64.
constexpr uint32_t idesc_n(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// ---- operand preparation: centroids [C, D] -> per-8-channel [hi tile | lo tile] in the exact smem layout ------------
__global__ void proto_prepare_kernel(const float* __restrict__ cen, int nclass, int D, unsigned char* __restrict__ ws,
                                     size_t cnorm_off) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (kstep, row, k)
  const int total = (D / KSTEP) * NPAD * KSTEP;
  if (idx < total) {
    const int k = idx % KSTEP, n = (idx / KSTEP) % NPAD, ks = idx / (KSTEP * NPAD);
    const float v = n < nclass ? cen[(size_t)n * D + ks * KSTEP + k] : 0.f;
    const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
    const float lo = v - hi;
    const size_t off = (size_t)ks * B_KSTEP_BYTES + (k >> 2) * 1024 + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4;
    *reinterpret_cast<float*>(ws + off) = hi;                       // rows 0..31 of the stacked operand
    *reinterpret_cast<float*>(ws + off + (NPAD / 8) * 128) = lo;    // rows 32..63
  }
  // ||c||^2: one warp per class row (blocks 0..3 hold 8 warps each), coalesced loads, two-level fp32 sum
  const int wglobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wglobal < NPAD) {
    float s = 0.f;
    if (wglobal < nclass) {
      for (int d0 = 0; d0 < D; d0 += 32 * 8) {
        float part = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int d = d0 + u * 32 + lane;
          const float v = d < D ? cen[(size_t)wglobal * D + d] : 0.f;
          part = fmaf(v, v, part);
        }
        s += part;
      }
      s = warp_sum(s);
    }
    if (lane == 0) reinterpret_cast<float*>(ws + cnorm_off)[wglobal] = s;
  }
}

// ---- main kernel ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
proto_umma_kernel(const __grid_constant__ CUtensorMap fmap, const unsigned char* __restrict__ ws, size_t cnorm_off, int nclass,
                  int64_t n_img, int D, int64_t hw, int tile_px, float* __restrict__ dist, float* __restrict__ weight, int dbg) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_ring = reinterpret_cast<float*>(smem + SMEM_A);
  unsigned char* b_ring = smem + SMEM_B;
  float* norm_part = reinterpret_cast<float*>(smem + SMEM_NORM);
  float* cnorm = reinterpret_cast<float*>(smem + SMEM_CNORM);
  Barriers* bar = reinterpret_cast<Barriers*>(smem + SMEM_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SMEM_TMEM_PTR);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int nch = D / KC;
  // A tile is `tile_px` <= 128 pixels wide (rows beyond it ride along in the MMA for free but are never loaded): the
  // host picks the width that makes the tile count fill whole rounds of CTAs (see proto_umma_launch).
  const int64_t tiles_per_img = (hw + tile_px - 1) / tile_px;
  const int box_w = tile_px + 4, box_floats = BOX_G * box_w;
  const int64_t n_tiles = n_img * tiles_per_img;
  const int hwm = (int)(hw & 3);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&fmap) : "memory");
    for (int i = 0; i < SA; ++i) { mbar_init(&bar->a_full[i], 1); mbar_init(&bar->a_empty[i], 4); }
    for (int i = 0; i < ST; ++i) { mbar_init(&bar->op_full[i], 4 * CPS + 1); mbar_init(&bar->op_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar->acc_full[i], 1); mbar_init(&bar->acc_empty[i], 4); mbar_init(&bar->norm_full[i], CONV_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < NPAD) cnorm[threadIdx.x] = reinterpret_cast<const float*>(ws + cnorm_off)[threadIdx.x];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  Prof prof;
  prof.on = (dbg & 32) != 0;
  prof.out = reinterpret_cast<long long*>(const_cast<unsigned char*>(ws) + cnorm_off + 256) + (size_t)blockIdx.x * 64;
#pragma unroll
  for (int i = 0; i < 8; ++i) prof.acc[i] = 0;
  const long long t_start = clock64();

  if (warp == WARP_LOAD) {
    // ===================== feature loader: one thread drives the TMA engine ============================================
    // Runs ahead of the converters as far as the shared-memory ring allows (SA stages = 135 KB in flight).
    uint32_t ga = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t img = tile / tiles_per_img;
      const int p0 = (int)((tile - img * tiles_per_img) * tile_px);
      for (int c = 0; c < nch; ++c, ++ga) {
        const uint32_t sa = ga % SA;
        prof.wait(0, &bar->a_empty[sa], ((ga / SA) & 1) ^ 1);
        const int grp = (int)((img * D + (int64_t)c * KC) >> 2);
        float* dst = a_ring + sa * A_STAGE_FLOATS;
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar->a_full[sa], 4 * box_floats * 4);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            tma_load_2d(dst + r * box_floats, &fmap, ((int)(r * hw) + p0) & ~3, grp, &bar->a_full[sa]);
        }
        __syncwarp();
      }
    }
  } else if (warp == WARP_LOAD_B) {
    // ===================== centroid loader: refills a stage as soon as the MMAs that read it have retired ==============
    uint32_t gb = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int c = 0; c < nch; c += CPS, ++gb) {
        const uint32_t sb = gb % SB;
        prof.wait(0, &bar->op_empty[sb], ((gb / SB) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bar->op_full[sb], B_STAGE_BYTES);
          bulk_g2s(b_ring + sb * B_STAGE_BYTES, ws + (size_t)c * (B_STAGE_BYTES / CPS), B_STAGE_BYTES, &bar->op_full[sb]);
        }
        __syncwarp();
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer: warp-uniform loop, one elected lane issues =====================================
    uint32_t gc = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      prof.wait(0, &bar->acc_empty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + TMEM_ACC_COL + acc * NACC;
      for (int c = 0; c < nch; c += CPS, ++gc) {                // gc counts operand stages (pairs of chunks) here
        const uint32_t st = gc % ST;
        prof.wait(1, &bar->op_full[st], (gc / ST) & 1);      // A operands stored to TMEM and centroid tiles landed
        tc_fence_after();
        const uint32_t a_base = tmem_base + st * TMEM_STAGE_COLS;
        const uint32_t b_base = smem_u32(b_ring + st * B_STAGE_BYTES);
        if (elect_one()) {
          if (!(dbg & 1)) {
#pragma unroll
            for (int ks = 0; ks < CPS * KC / KSTEP; ++ks) {
              const uint32_t a_hi = a_base + (ks / (KC / KSTEP)) * 2 * KC + (ks % (KC / KSTEP)) * KSTEP, a_lo = a_hi + KC;
              const uint64_t b_desc = make_b_desc(b_base + ks * B_KSTEP_BYTES);
              // D[:, 0:64) (+)= A_hi * [b_hi | b_lo]^T ;  D[:, 0:32) += A_lo * b_hi^T   (2 MMAs instead of 3 per 8 channels)
              tc_mma_tf32_ts(d_tmem, a_hi, b_desc, idesc_n(2 * NPAD), (c | ks) != 0);
              tc_mma_tf32_ts(d_tmem, a_lo, b_desc, idesc_n(NPAD), 1u);
            }
          }
          tc_commit(&bar->op_empty[st]);      // both operand stages are free when these MMAs retire
          if (c + CPS >= nch) tc_commit(&bar->acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp < CONV_WARPS) {
    // ===================== converters: smem -> registers -> (norm, hi/lo split) -> TMEM A operand =====================
    const int wg = warp >> 2;                       // this warpgroup converts chunks c % NWG == wg
    const int row = (warp & 3) * 32 + lane;         // pixel row == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      float fn = 0.f;
      for (int c = wg; c < nch; c += NWG) {
        const uint32_t gc = it * (uint32_t)nch + (uint32_t)c;
        const uint32_t sa = gc % SA, pr = gc / CPS, st = pr % ST;       // pr: operand stage use (pair of chunks)
        prof.wait(0, &bar->a_full[sa], (gc / SA) & 1);
        const float* src = a_ring + sa * A_STAGE_FLOATS + row;
        const long long t_lds = prof.on ? clock64() : 0;
        float x[KC];
        if (dbg & 4) {
#pragma unroll
          for (int j = 0; j < KC; ++j) x[j] = 1.0f;
        } else
#pragma unroll
        for (int j = 0; j < KC; ++j)   // channel c*32+j = box r = j&3, group j>>2; box r starts (r*hw)&3 floats early
          x[j] = src[(j & 3) * box_floats + (j >> 2) * box_w + (((j & 3) * hwm) & 3)];
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < KC; ++j) part = fmaf(x[j], x[j], part);
        fn += part;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->a_empty[sa]);
        if (prof.on) prof.acc[3] += clock64() - t_lds;
        prof.wait(1, &bar->op_empty[st], ((pr / ST) & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_hi = tmem_base + lane_addr + st * TMEM_STAGE_COLS + (gc % CPS) * 2 * KC;
#pragma unroll
        for (int q = 0; q < KC / 8; ++q) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float v = x[q * 8 + j];
            hi[j] = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
            lo[j] = __float_as_uint(v - __uint_as_float(hi[j]));
          }
          if (!(dbg & 2)) {
            tc_st8(t_hi + q * 8, hi);
            tc_st8(t_hi + KC + q * 8, lo);
          } else if (hi[0] == 0x12345u && lo[1] == 0x54321u) {
            tc_st8(t_hi + q * 8, hi);
          }
        }
        const long long t_st = prof.on ? clock64() : 0;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (prof.on) prof.acc[2] += clock64() - t_st;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->op_full[st]);
      }
      norm_part[(acc * NWG + wg) * TILE_M + row] = fn;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar->norm_full[acc]);
    }
  } else {
    // ===================== epilogue: TMEM accumulator -> dist / softmax(-dist) ========================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, par = (it >> 1) & 1;
      const int64_t img = tile / tiles_per_img;
      const int64_t p = row < tile_px ? (tile - img * tiles_per_img) * tile_px + row : hw;   // rows past the tile width are padding
      prof.wait(0, &bar->acc_full[acc], par);
      prof.wait(1, &bar->norm_full[acc], par);
      tc_fence_after();
      uint32_t r0[16], r1[16], r2[16], r3[16];
      const uint32_t t = tmem_base + lane_addr + TMEM_ACC_COL + acc * NACC;
      tc_ld16(t, r0);
      tc_ld16(t + 16, r1);
      tc_ld16(t + 32, r2);
      tc_ld16(t + 48, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float fn = 0.f;
#pragma unroll
      for (int g = 0; g < NWG; ++g) fn += norm_part[(acc * NWG + g) * TILE_M + row];
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar->acc_empty[acc]);
      if (p < hw) {
        float d[NPAD];
        float dmin = 3.4e38f;
#pragma unroll
        for (int c = 0; c < NPAD; ++c) {
          const float dot = __uint_as_float(c < 16 ? r0[c & 15] : r1[c & 15]) + __uint_as_float(c < 16 ? r2[c & 15] : r3[c & 15]);
          const float d2 = fmaf(-2.f, dot, fn + cnorm[c]);
          d[c] = sqrtf(fmaxf(d2, 0.f));
          if (c < nclass) dmin = fminf(dmin, d[c]);
        }
        const int64_t o = img * nclass * hw + p;
        if (dist) {
#pragma unroll
          for (int c = 0; c < NPAD; ++c)
            if (c < nclass) dist[o + c * hw] = d[c];
        }
        if (weight) {
          float S = 0.f;
#pragma unroll
          for (int c = 0; c < NPAD; ++c)
            if (c < nclass) {
              d[c] = fast_exp(dmin - d[c]);
              S += d[c];
            }
          const float inv = 1.0f / S;
#pragma unroll
          for (int c = 0; c < NPAD; ++c)
            if (c < nclass) weight[o + c * hw] = d[c] * inv;
        }
      }
    }
  }

  if (prof.on && lane == 0 && (warp == WARP_LOAD || warp == WARP_LOAD_B || warp == WARP_MMA || warp == 0 || warp == CONV_WARPS)) {
    const int role = warp == WARP_LOAD ? 0 : warp == WARP_LOAD_B ? 1 : warp == WARP_MMA ? 2 : warp == 0 ? 3 : 4;
    prof.out[role * 8 + 7] = clock64() - t_start;
    for (int i = 0; i < 7; ++i) prof.out[role * 8 + i] = prof.acc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace umma

size_t proto_umma_workspace_bytes(int64_t C, int64_t D) {
  (void)C;
  const int64_t ksteps = (D + umma::KSTEP - 1) / umma::KSTEP;
  return (size_t)ksteps * umma::B_KSTEP_BYTES + umma::NPAD * sizeof(float) + 256 + 256 * 64 * sizeof(long long);
}

int proto_umma_supported(int64_t n, int64_t D, int64_t C, int64_t hw) {
  return n >= 1 && hw >= 1 && C >= 1 && C <= umma::NPAD && (D % (umma::KC * umma::CPS)) == 0 && D >= 256 && D <= (1 << 20) &&
         4 * hw < ((int64_t)1 << 31) && n * D / 4 < ((int64_t)1 << 31);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int proto_umma_prepare(const float* centroids, int64_t D, int64_t C, void* workspace, cudaStream_t st) {
  using namespace umma;
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t cnorm_off = (size_t)(D / KSTEP) * B_KSTEP_BYTES;
  const int total = (int)(D / KSTEP) * NPAD * KSTEP;
  proto_prepare_kernel<<<(total + 255) / 256, 256, 0, st>>>(centroids, (int)C, (int)D, ws, cnorm_off);
  DIGA_CHECK_LAUNCH("proto_prepare_kernel");
  return DIGA_OK;
}

int proto_umma_launch(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw, float* dist,
                      float* weight, void* workspace, int prepared, cudaStream_t st) {
  using namespace umma;
  if (!aligned(feat, 16) || !aligned(workspace, 128)) {
    set_error("proto_umma: feature pointer must be 16-byte and workspace 128-byte aligned");
    return DIGA_ERR_MISALIGNED;
  }
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) {
    set_error("proto_umma: cuTensorMapEncodeTiled not available from the driver");
    return DIGA_ERR_CUDA;
  }
  // Tile width.  Time per tile is proportional to the pixels it loads, time per launch to rounds * width, so instead of
  // 128-pixel tiles whose count rarely fills the last round of CTAs (528 tiles on 148 SMs: 4 rounds for 3.57 of work)
  // the width is chosen, among the per-image tile counts up to one extra round, to minimise rounds * width
  // (8 x 65x129: 74 tiles of 116 px per image = 592 = 4 x 148 tiles, -9 % time).
  int tile_px = TILE_M;
  if (tunable("umma_fit_tiles", 1)) {
    const int64_t sms = sm_count();
    const int64_t t_min = (hw + TILE_M - 1) / TILE_M;
    double best = 1e30;
    for (int64_t t = t_min; t <= t_min + (sms + n - 1) / n + 1; ++t) {
      int64_t px = (hw + t - 1) / t;
      px = (px + 3) & ~(int64_t)3;
      if (px < 32) break;
      if (px > TILE_M) continue;
      const int64_t rounds = (n * ((hw + px - 1) / px) + sms - 1) / sms;
      const double cost = (double)rounds * (double)(px + 6);       // +6: box padding and per-tile fixed cost
      if (cost < best - 1e-9) {
        best = cost;
        tile_px = (int)px;
      }
    }
  }
  // features viewed as [n*D/4 groups][4*hw]: the group pitch 16*hw bytes is a legal TMA stride for every hw
  CUtensorMap fmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)(4 * hw), (cuuint64_t)(n * D / 4)};
  const cuuint64_t gstride[1] = {(cuuint64_t)(16 * hw)};
  const cuuint32_t box[2] = {(cuuint32_t)(tile_px + 4), (cuuint32_t)BOX_G};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("proto_umma: cuTensorMapEncodeTiled failed with %d", (int)r);
    return DIGA_ERR_CUDA;
  }
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  const size_t cnorm_off = (size_t)(D / KSTEP) * B_KSTEP_BYTES;
  if (!prepared) {
    const int rc = proto_umma_prepare(centroids, D, C, workspace, st);
    if (rc != DIGA_OK) return rc;
  }
  static bool configured_dev[64] = {false};          // the attribute is per device
  bool& configured = configured_dev[device_slot()];
  if (!configured) {
    if (cudaFuncSetAttribute(proto_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("proto_umma: cannot reserve %d bytes of shared memory", SMEM_TOTAL);
      return DIGA_ERR_CUDA;
    }
    configured = true;
  }
  const int64_t tiles = n * ((hw + tile_px - 1) / tile_px);
  int64_t grid = sm_count();
  if (grid > tiles) grid = tiles;
  proto_umma_kernel<<<(unsigned)grid, THREADS, SMEM_TOTAL, st>>>(fmap, ws, cnorm_off, (int)C, n, (int)D, hw, tile_px, dist, weight,
                                                                     tunable("umma_debug", 0));
  DIGA_CHECK_LAUNCH("proto_umma_kernel");
  return DIGA_OK;
}

}  // namespace diga
