/*
 * diga_b200 — C ABI of the B200 (sm_100a) implementation of DiGA's per-pixel adaptation hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (fy-vision/DiGA) is pure Python, so "the FFI a maintainer would bind" is a ctypes stub; it is
 * shown in INTEGRATION.md and implemented in diga_b200/_lib.py.  Each entry point names the
 * reference code it replaces; paths are relative to domain_adaptation/GTA5/ of the reference.
 *
 * Conventions
 *   - every pointer is DEVICE memory owned by the caller unless the name ends in _host;
 *   - tensors are contiguous NCHW fp32, labels are int64 (the reference's LongTensor);
 *   - `hw` is H*W of one plane, `stream` is a cudaStream_t (NULL = legacy default stream);
 *   - a call only enqueues work: no allocation, no host synchronisation, no exceptions;
 *   - return 0 on success, a negative diga_status otherwise; diga_last_error_string()
 *     describes the last failure on the calling thread.
 */
#ifndef DIGA_B200_H_
#define DIGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* diga_stream_t; /* cudaStream_t */

enum diga_status {
  DIGA_OK = 0,
  DIGA_ERR_INVALID = -1,     /* bad shape / null pointer / unsupported class count */
  DIGA_ERR_MISALIGNED = -2,  /* pointer not aligned to its element size */
  DIGA_ERR_CUDA = -3,        /* launch failed; see diga_last_error_string() */
  DIGA_ERR_WORKSPACE = -4,   /* workspace too small */
  DIGA_ERR_IO = -5           /* a host file could not be written (diga_png_write_file) */
};

enum diga_update_mode { DIGA_UPDATE_MEAN = 0, DIGA_UPDATE_MOVING_AVERAGE = 1 };

#define DIGA_MAX_CLASSES 32
#define DIGA_MAX_PEERS 16          /* ranks of one NVLink domain a peer-memory kernel can address */
#define DIGA_IGNORE_LABEL 255

int diga_version(void);
const char* diga_last_error_string(void);
/* Number of kernels this library has launched in the calling process (bench.py: gpu_launches). */
int64_t diga_launch_count(void);
/* Launch-shape tunables (vector width, waves per SM, kernel variant); not part of the reference-facing
 * surface — used by tools/tune.py and the tests.  DIGA_TUNE_<NAME> in the environment does the same. */
int diga_set_tunable(const char* name, int value);

/* ------------------------------------------------------------------------------------------
 * a1  symmetric KD loss — util/loss.py:125-143 (scale 0.25: Synthia/util/loss.py:52)
 *   teacher, student: [n2, C, hw] fp32, n2 = 2B even.  Student view v is supervised by
 *   softmax(teacher view 1-v); the pair whose teacher is view 1 is weighted by `scale`.
 *   loss = mean_{B,hw} CE(p0, s1) + scale * mean_{B,hw} CE(p1, s0).
 * workspace: diga_kd_workspace_bytes() bytes, zero-filled once by the caller before first use
 *   (the kernels leave it zeroed); one workspace per concurrently used stream.
 * ------------------------------------------------------------------------------------------ */
size_t diga_kd_workspace_bytes(void);
/* forward: loss_out[0] = loss (deterministic two-stage reduction). */
int diga_kd_fwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                float* loss_out, void* workspace, diga_stream_t stream);
/* backward (autograd of the above): dstudent = upstream[0] * dL/dstudent; `upstream` is a DEVICE
 * scalar (the 0-dim grad_output), so no host sync is needed. */
int diga_kd_bwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                const float* upstream, float* dstudent, diga_stream_t stream);
/* single pass: loss and upstream_host * gradient together (228 B/px instead of 152 + 228). */
int diga_kd_fwd_bwd(const float* teacher, const float* student, int64_t n2, int64_t C, int64_t hw, float scale,
                    float upstream_host, float* loss_out, float* dstudent, void* workspace, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a3  pseudo-label — pseudolabel_generator.py:80-85
 *   z = max(logits, logits_ds) (logits_ds may be NULL); label = argmax_c softmax(z) (first index
 *   on ties, taken on the logits); conf = max_c softmax(z).  Any output may be NULL.
 *   logits: [n, C, hw]; label_u8/label_i64/conf: [n, hw].
 * ------------------------------------------------------------------------------------------ */
int diga_pseudo_label(const float* logits, const float* logits_ds, int64_t n, int64_t C, int64_t hw,
                      uint8_t* label_u8, int64_t* label_i64, float* conf, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a2  ClassMix — train_DiGA_gta2city_self_training.py:259-275 (image), :306-325 (DACS),
 *     train_DiGA_gta2city_warm_up.py:240-259, calc_centroids.py:47-58
 *   class_presence: bitmap[b][8] (256 bits) of label values present in image b — the device half
 *     of torch.unique(slabel[b]); flags[0] |= 1 if a label lies outside [0,255].
 *     bitmap and flags are cleared by the call.
 *   classmix_blend: lut_host[b][256] (uint8 in HOST memory, 1 = class selected for image b; it is consumed
 *     before the call returns and travels to the GPU as a kernel argument) ->
 *     mask[b,p] = lut_host[b][slabel[b,p]]  (fp32 0/1, may be NULL)
 *     mix[b,c,p] = a*(1-mask) + b*mask  (evaluated exactly as written, no FMA contraction)
 *     mixlabel[b,p] = mask ? slabel : tlabel   (only if tlabel != NULL)
 * ------------------------------------------------------------------------------------------ */
int diga_class_presence(const int64_t* slabel, int64_t B, int64_t hw, uint32_t* bitmap, uint32_t* flags,
                        diga_stream_t stream);
int diga_classmix_blend(const int64_t* slabel, const uint8_t* lut_host, const float* a, const float* b,
                        const int64_t* tlabel, int64_t B, int64_t channels, int64_t hw,
                        float* mask, float* mix, int64_t* mixlabel, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6  per-image per-class masked feature mean — calc_centroids.py:120-145 (:97-118 by_output),
 *     one-hot helper util/utils.py:158-163
 *   assign: cls[n,p] = argmax_c logits[n,:,p] if (labels == NULL or (int64)labels[n,p] == argmax)
 *           else 255; counts[n][c] = number of pixels assigned to c (cleared by the call).
 *           labels: [n,1,hw] fp32 as produced by F.interpolate(mode='nearest') of the int64 map.
 *   accum : sums[n][c][d] = sum over pixels of class c of feat[n,d,p]  (fully overwritten,
 *           deterministic: every output element has exactly one writer, no atomics).
 *   means : vec[n][c][d] = sums / counts (one correctly rounded division; the reference's (sum/hw) / (count/hw) rounds
 *           three times around the same value); valid[n][c] = counts >= 5;
 *           vecsum[n][c] = sum_d vec (for the reference's `vector.sum() == 0` skip).
 * ------------------------------------------------------------------------------------------ */
/* clsw (optional, may be NULL): phase-shifted class words for the 128-bit accumulation kernel, diga_centroid_clsw_bytes(n, hw)
 * bytes — four copies of the class map delayed by 0..3 bytes so that an aligned quad of any channel row reads its four
 * class bytes as one aligned word; out-of-row and gated-out pixels name a dummy class C.  Written by the assign calls, or
 * from a plain class map by diga_centroid_clsw_build. */
int64_t diga_centroid_clsw_bytes(int64_t n, int64_t hw);
int diga_centroid_clsw_build(const uint8_t* cls, int64_t n, int64_t C, int64_t hw, uint32_t* clsw, diga_stream_t stream);
int diga_centroid_assign(const float* logits, const float* labels, int64_t n, int64_t C, int64_t hw,
                         uint8_t* cls, int32_t* counts, uint32_t* clsw, diga_stream_t stream);
/* assign with the reference's label down-sampling folded in (self_training.py:327-330, :336-337): labels_full is the
 * [n,H,W] int64 map; pixel (y,x) of the h x w feature grid is gated by labels_full[min(floor(y*H/h), H-1)][min(floor(x*W/w), W-1)]
 * (F.interpolate(mode='nearest') of the .float() map), same gate as diga_centroid_assign. */
int diga_centroid_assign_fullres(const float* logits, const int64_t* labels_full, int64_t n, int64_t C, int64_t h, int64_t w,
                                 int64_t H, int64_t W, uint8_t* cls, int32_t* counts, uint32_t* clsw, diga_stream_t stream);
/* process_label (util/utils.py:158-163): label [B,1,hw] fp32 -> onehot [B,C+1,hw] fp32, ids >= C in channel C.
 * Negative labels are outside the reference's domain (scatter_ would raise) and give an all-zero column. */
int diga_onehot_labels(const float* label, int64_t B, int64_t C, int64_t hw, float* onehot, diga_stream_t stream);
/* counts (optional): the per-image class counts of the assign call — classes absent from an image are then skipped and
 * their sums left unwritten (diga_centroid_means reads sums only where counts > 0).  clsw (optional): see above. */
int diga_centroid_accum(const float* feat, const uint8_t* cls, const int32_t* counts, const uint32_t* clsw, int64_t n,
                        int64_t D, int64_t C, int64_t hw, float* sums, diga_stream_t stream);
int diga_centroid_means(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw,
                        float* vec, float* vecsum, uint8_t* valid, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a7  centroid running update — calc_centroids.py:147-164
 *   Applies, in (image, class) order, update_objective_SingleVector(c, vec[n][c], mode, start_mean)
 *   for every (n, c) with valid[n][c] != 0 and vecsum[n][c] != 0.  valid may be NULL (all valid).
 *   objective_vectors [C, D], objective_num [C] are updated in place (fp32, clamp at 3000,
 *   start_mean forces 'mean' while num < 100).  `momentum` is a double because the reference
 *   evaluates (1 - centroid_momentum) in Python double precision before torch rounds it to fp32.
 * ------------------------------------------------------------------------------------------ */
int diga_centroid_update(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n, int64_t C,
                         int64_t D, float* objective_vectors, float* objective_num, int mode, int start_mean,
                         double momentum, diga_stream_t stream);
/* means + update in one launch (the online path: a batch of a few images per call): equals diga_centroid_means followed by
 * diga_centroid_update on the same rows.  One thread-block cluster per class; the per-image vector.sum() of :148 is reduced
 * over the cluster through distributed shared memory.  vec / vecsum / valid are optional outputs (NULL: not stored);
 * objective_vectors == NULL: only the means are produced.  Supported while diga_centroid_finish_supported(n, D) != 0
 * (n * ceil(D / 2048) <= 16 rows per thread), larger batches use the two separate calls. */
int diga_centroid_finish_supported(int64_t n, int64_t D);
int diga_centroid_finish(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw,
                         float* vec, float* vecsum, uint8_t* valid, float* objective_vectors, float* objective_num,
                         int mode, int start_mean, double momentum, diga_stream_t stream);
/* a6 -> a7 in ONE call (what Class_Features.update_from_features queues): assign (labels: [n,1,h*w] fp32 or NULL;
 * labels_full: [n,H,W] int64 or NULL) -> accum -> finish (means + update when the batch is too large for finish), all on
 * `stream`, scratch in one caller-allocated 256-byte aligned workspace of diga_centroid_chain_workspace_bytes bytes.
 * Equals calculate_mean_vector followed by update_objective_SingleVector on every vector in (image, class) order. */
int64_t diga_centroid_chain_workspace_bytes(int64_t n, int64_t C, int64_t D, int64_t hw);
int diga_centroid_chain(const float* feat, const float* logits, const float* labels, const int64_t* labels_full,
                        int64_t H, int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, void* workspace,
                        float* objective_vectors, float* objective_num, int mode, int start_mean, double momentum,
                        diga_stream_t stream);
/* The same chain up to the class sums (assign -> accum): *sums_out [n,C,D] and *counts_out [n,C] point into the workspace. */
int diga_centroid_chain_sums(const float* feat, const float* logits, const float* labels, const int64_t* labels_full,
                             int64_t H, int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w, void* workspace,
                             float** sums_out, int32_t** counts_out, diga_stream_t stream);
/* ... and ending in the multi-GPU sum-mode accumulator: assign -> accum -> means -> acc[C, D+1] += (vectors, their number). */
int diga_centroid_chain_reduce(const float* feat, const float* logits, const float* labels, const int64_t* labels_full,
                               int64_t H, int64_t W, int64_t n, int64_t C, int64_t D, int64_t h, int64_t w,
                               void* workspace, float* acc, diga_stream_t stream);
/* means fused with the exchange of the image-sharded exact mode: like diga_centroid_means, but the rows (vec, vecsum, valid of
 * the n local images) are stored at row row0.. of the gathered buffers of ALL `world` ranks through peer pointers
 * (peer_bases[r]: base of rank r's symmetric allocation, host array; the three arrays live at the given byte offsets in each).
 * multicast_base (optional): the NVLS multicast mapping of the same allocation — the vectors are then written with ONE
 * multimem.st per element that the NVSwitch replicates into every rank's copy.
 * NVLink stores from the kernel — the all-gather of the pass disappears.  Visibility at the peers is the caller's
 * barrier (the stores are complete when this kernel has finished on its stream). */
int diga_centroid_means_scatter(const float* sums, const int32_t* counts, int64_t n, int64_t C, int64_t D, int64_t hw,
                                void* const* peer_bases, void* multicast_base, int64_t world, int64_t off_vec,
                                int64_t off_vecsum, int64_t off_valid, int64_t row0, diga_stream_t stream);
/* Image-sharded exact replay (calc_centroids.py:20-23,67-78,147-164): vec/vecsum/valid are the all-gather
 * of every rank's rows, [world][per_shard] x C (x D); rank r holds loader batches r, r+world, ... of `group` images each.
 * The n_total images are visited in GLOBAL loader order (image g = batch g/group, position g%group), so every rank
 * reproduces the single-process sequence bit for bit, beyond the 3000 clamp and across passes.  Rows that no image maps to
 * (padding of the shorter shards) must carry vecsum == 0. */
int diga_centroid_update_sharded(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n_total,
                                 int64_t group, int64_t world, int64_t per_shard, int64_t C, int64_t D,
                                 float* objective_vectors, float* objective_num, int mode, int start_mean,
                                 double momentum, diga_stream_t stream);
/* single vector form (the reference's per-call API): class `id`, vector [D]. */
int diga_centroid_update_single(const float* vector, int64_t id, int64_t C, int64_t D, float* objective_vectors,
                                float* objective_num, int mode, int start_mean, double momentum,
                                diga_stream_t stream);
/* multi-GPU 'mean' pass: acc[c][0..D) += sum over valid images of vec, acc[c][D] += their number.
 * acc: [C, D+1] fp32 (the buffer that is all-reduced over NCCL). */
int diga_centroid_reduce_images(const float* vec, const float* vecsum, const uint8_t* valid, int64_t n,
                                int64_t C, int64_t D, float* acc, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a5  prototype distance + reweighted softmax — calc_centroids.py:166-180
 *   dist[n,c,p] = || centroids[c,:] - feat[n,:,p] ||_2 ; weight = softmax_c(-dist).
 *   Either output may be NULL.  feat: [n, D, hw]; centroids: [C, D]; outputs: [n, C, hw].
 *   workspace: diga_proto_workspace_bytes(C, D) bytes (split centroid operands for the tensor-core path).
 * ------------------------------------------------------------------------------------------ */
size_t diga_proto_workspace_bytes(int64_t C, int64_t D);
int diga_proto_distance(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C, int64_t hw,
                        float* dist, float* weight, void* workspace, diga_stream_t stream);
/* Same, split in two for many calls against unchanged centroids (offline pseudo-label rectification): prepare the
 * split centroid operands once, then call the _prepared form with the same workspace. */
int diga_proto_prepare(const float* centroids, int64_t C, int64_t D, void* workspace, diga_stream_t stream);
int diga_proto_distance_prepared(const float* feat, const float* centroids, int64_t n, int64_t D, int64_t C,
                                 int64_t hw, float* dist, float* weight, void* workspace, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4  bilateral-consensus selection — train_DiGA_gta2city_self_training.py:298-304
 *   W = bilinear(align_corners=True) up-sampling of weights_lowres [B,C,h,w] to [B,C,H,W]
 *   (never materialised); feat_pseudo = argmax_c W; kept = pseudo if pseudo == feat_pseudo else 255.
 *   feat_pseudo may be NULL.
 * ------------------------------------------------------------------------------------------ */
int diga_consensus_select(const float* weights_lowres, const int64_t* pseudo, int64_t B, int64_t C,
                          int64_t h, int64_t w, int64_t H, int64_t W, int64_t* kept, int64_t* feat_pseudo,
                          diga_stream_t stream);

/* Same on uint8 label maps (the offline pseudo-label path: labels come from diga_pseudo_label* as uint8 and go to palette
 * PNGs): 1 B/px read, 1-2 B/px written instead of 8 + 8-16. */
int diga_consensus_select_u8(const float* weights_lowres, const uint8_t* pseudo, int64_t B, int64_t C, int64_t h, int64_t w,
                             int64_t H, int64_t W, uint8_t* kept, uint8_t* feat_pseudo, diga_stream_t stream);

/* The interpolation routine shared by a4 and the fused a3 variant, materialising:
 * out[planes,H,W] = bilinear(align_corners=True) of in[planes,h,w], bit-identical to torch's CUDA kernel. */
int diga_upsample_bilinear(const float* in, int64_t planes, int64_t h, int64_t w, int64_t H, int64_t W, float* out,
                           diga_stream_t stream);

/* (f)-next row 1: pseudo-label straight from the stride-8 logits, both bilinear up-samplings fused
 * (pseudolabel_generator.py:77-85).  logits: [n,C,h1,w1], logits_ds: [n,C,h2,w2] or NULL. */
int diga_pseudo_label_upsampled(const float* logits, int64_t h1, int64_t w1, const float* logits_ds, int64_t h2,
                                int64_t w2, int64_t n, int64_t C, int64_t H, int64_t W, uint8_t* label_u8,
                                int64_t* label_i64, float* conf, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f2 (next row)  cross_entropy2d — util/loss.py:48-62
 *   logits [n, C, hw] fp32, target [n, hw] int64 (255 = ignore, negative = dropped), weight [C] fp32 or NULL.
 *   fwd: loss_out[0] = sum_{valid} -w[t] log_softmax(logits)[t]  (/ #(target >= 0) if size_average);
 *        denom_out[0] = #(target >= 0) (kept for the backward).  workspace: diga_ce_workspace_bytes(), zero-filled
 *        once by the caller (the kernel leaves it zeroed).
 *   bwd: dlogits = upstream[0] * dloss/dlogits; `upstream` and `denom` are DEVICE scalars.
 * ------------------------------------------------------------------------------------------ */
size_t diga_ce_workspace_bytes(void);
int diga_cross_entropy2d_fwd(const float* logits, const int64_t* target, const float* weight, int64_t n, int64_t C,
                             int64_t hw, int size_average, float* loss_out, float* denom_out, void* workspace,
                             diga_stream_t stream);
int diga_cross_entropy2d_bwd(const float* logits, const int64_t* target, const float* weight, int64_t n, int64_t C,
                             int64_t hw, int size_average, const float* upstream, const float* denom, float* dlogits,
                             diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f1 (next row 1) for the loss consumers — train_DiGA_gta2city_self_training.py:289,:344,:348-352,:355
 *   The losses evaluated on bilinear(align_corners=True) up-sampled logits, straight from the stride-8 logits:
 *   neither the up-sampled [n,C,H,W] tensors nor their gradients are materialised.
 *     KD (util/loss.py:125-143): teacher_low, student_low [n,C,h,w], n = 2B even; loss_kd as diga_kd_fwd on the
 *        up-sampled pair.  teacher_low == NULL switches KD off.
 *     CE (util/loss.py:48-62): target [n_ce,H,W] int64 supervises the FIRST n_ce images of student_low (n_ce == n for
 *        a plain cross_entropy2d; n_ce == B for the source half of s_pred_cat_stu, :348-349); semantics of
 *        diga_cross_entropy2d_fwd.  target == NULL switches CE off.
 *   bwd: dstudent_low [n,C,h,w] (fully overwritten) = upstream_kd[0]*dloss_kd/dstudent_low
 *        + upstream_ce[0]*dloss_ce/dstudent_low, i.e. the transposed interpolation applied to the per-pixel gradient;
 *        no atomics, bitwise deterministic.  upstream_* / denom are DEVICE scalars (denom = denom_out of the forward).
 *   kd_up_fwd_bwd: KD loss and upstream_host * gradient in a single pass.
 *   workspace: diga_loss_up_workspace_bytes(n, C, h, w, H, W) bytes, 16-byte aligned; its first 32 bytes zero-filled
 *        once by the caller (the kernels leave them zeroed).  Requires H >= h and W >= w.
 * ------------------------------------------------------------------------------------------ */
size_t diga_loss_up_workspace_bytes(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W);
int diga_loss_up_fwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                     int64_t n, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                     int size_average, float* loss_kd, float* loss_ce, float* denom_out, void* workspace,
                     diga_stream_t stream);
int diga_loss_up_bwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                     int64_t n, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                     int size_average, const float* upstream_kd, const float* upstream_ce, const float* denom,
                     float* dstudent_low, void* workspace, diga_stream_t stream);
/* CE loss and, in the same pass, the gradient of the SUMMED loss (size_average only affects loss_out): the caller scales
 * it by upstream / denom_out — what the autograd wrapper does when the logits require a gradient. */
int diga_ce_up_fwd_bwd(const float* logits_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h,
                       int64_t w, int64_t H, int64_t W, int size_average, float* loss_out, float* denom_out,
                       float* dlogits_sum, float* scratch, void* workspace, diga_stream_t stream);
/* Deferred gather (the three *_fwd_bwd entry points): with `scratch` (diga_loss_up_scratch_bytes bytes, 16-byte aligned,
 * need not be initialised) the per-CTA gradient patches go there instead of the workspace, and with the gradient pointer NULL
 * they stay there: diga_loss_up_gather sums them into dlow [n,C,h,w] later, multiplied by num[0] / den[0] (den NULL: num[0]) read
 * on the device — the autograd backward, where the upstream scalar of loss.backward() first becomes known.  One launch instead
 * of gather + scale; the same roundings. */
size_t diga_loss_up_scratch_bytes(int64_t n, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W);
int diga_loss_up_gather(const float* scratch, const float* num, const float* den, int64_t n, int64_t C, int64_t h, int64_t w,
                        int64_t H, int64_t W, float* dlow, diga_stream_t stream);
/* Both losses of self_training.py:348-352 and the gradient of lambda_ce * loss_ce + lambda_kd * loss_kd (:382) in ONE pass
 * over the stride-8 logits, for call sites that know the two loss weights when the losses are computed.  `loss_total`
 * (nullable) receives that weighted sum, rounded like the three fp32 scalar operations of :356 / :382.  The CE gradient is
 * divided by #(target >= 0) (util/loss.py:56,:60): `denom_known` = 0 counts it first (one pass over the targets);
 * `denom_known` > 0 is the caller's promise of that count (loader labels are trainIds or 255, so it is n_ce * H * W) — the
 * value counted by the loss reduction is compared with it and loss_ce / loss_total come back NaN if they differ. */
int diga_seg_kd_up_fwd_bwd(const float* teacher_low, const float* student_low, const int64_t* target, const float* weight,
                           int64_t n2, int64_t n_ce, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W, float scale,
                           int size_average, float lambda_ce_host, float lambda_kd_host, float denom_known, float* loss_kd,
                           float* loss_ce, float* denom_out, float* loss_total, float* dstudent_low, float* scratch,
                           void* workspace, diga_stream_t stream);
int diga_kd_up_fwd_bwd(const float* teacher_low, const float* student_low, int64_t n2, int64_t C, int64_t h, int64_t w,
                       int64_t H, int64_t W, float scale, float upstream_host, float* loss_out, float* dstudent_low,
                       float* scratch, void* workspace, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * OhemCrossEntropy — util/loss.py:65-122 (the segmentation loss of the Synthia tree), from the stride-8 scores
 *   (`_ohem_forward` up-samples the score to the label size itself, :91-96; H == h, W == w is the plain case).
 *   pred = softmax(score)[target] over pixels with target != ignore_label; threshold = max(sorted(pred)[min(min_kept, M-1)],
 *   thresh); loss = mean of -w[t] log_softmax(score)[t] over pixels with pred < threshold.
 *   fwd: pred, losspx [n,H,W] fp32 (caller-allocated; pred is needed by the backward), loss_out / count_out (number of
 *        kept pixels) / thr_out device scalars; the order statistic is exact (radix select on the float bits).
 *        workspace: diga_ohem_up_workspace_bytes() bytes, zero-filled once by the caller (left zeroed).
 *   bwd: dlogits_low = upstream[0] / count[0] * sum over kept pixels of w[t] (softmax - onehot), through the transposed
 *        interpolation of the f1 loss kernels; workspace = diga_loss_up_workspace_bytes(n, C, h, w, H, W).
 * ------------------------------------------------------------------------------------------ */
size_t diga_ohem_up_workspace_bytes(void);
int diga_ohem_up_fwd(const float* score_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h,
                     int64_t w, int64_t H, int64_t W, int64_t ignore_label, float thresh, int64_t min_kept, float* pred,
                     float* losspx, float* loss_out, float* count_out, float* thr_out, void* workspace, diga_stream_t stream);
int diga_ohem_up_bwd(const float* logits_low, const int64_t* target, const float* weight, int64_t n, int64_t C, int64_t h, int64_t w,
                     int64_t H, int64_t W, int64_t ignore_label, const float* pred, const float* thr, const float* count,
                     const float* upstream, float* dlogits_low, void* workspace, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f4 (next row)  EMA teacher update — util/utils.py:103-116
 *   For each of `count` parameter tensors: teacher = alpha * teacher + (1 - alpha) * student (fp32, in place,
 *   separately rounded like the torch expression).  The three tables live in HOST memory (device pointers and
 *   element counts) and are consumed before the call returns; 512 tensors per launch.
 * ------------------------------------------------------------------------------------------ */
int diga_ema_update(float* const* teacher_host, const float* const* student_host, const int64_t* numel_host,
                    int64_t count, double alpha, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f5 (next row)  evaluation confusion matrix — util/metrics.py:32-41 (runningScore._fast_hist / update)
 *   hist[n_class * t + p] += 1 for every pixel with 0 <= t < n_class (t = label_true, p = label_pred); hist is an int64
 *   [n_class, n_class] DEVICE matrix accumulated in place (exact, order-independent).  Labels are uint8 or int64 maps
 *   (`*_is_u8`), `total` pixels each.  A prediction outside [0, n_class) on a counted pixel (where the reference's
 *   bincount(...).reshape would raise) is skipped and reported as flags[0] |= 1 (flags is NOT cleared by the call).
 * ------------------------------------------------------------------------------------------ */
int diga_confusion_matrix(const void* label_true, int true_is_u8, const void* label_pred, int pred_is_u8, int64_t total,
                          int64_t n_class, int64_t* hist, uint32_t* flags, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f3 (next row), reader half — util/loader/CityLoader.py:93-95 (NEAREST resize) + :115-132 (id re-assignment)
 *   out[i, y, x] = lut[src[i, ytab[y], xtab[x]]]: src uint8 [n, h0, w0] (decoded label / pseudo-label PNGs), ytab [H] and
 *   xtab [W] int32 DEVICE tables of source indices (Pillow's NEAREST transform, built by the caller once per geometry),
 *   lut_host[256] uint8 in HOST memory (consumed before the call returns: id -> trainId, or v < 19 ? v : 255), out int64.
 * ------------------------------------------------------------------------------------------ */
int diga_label_resize_remap(const uint8_t* src, int64_t n, int64_t h0, int64_t w0, const int32_t* ytab, const int32_t* xtab,
                            int64_t H, int64_t W, const uint8_t* lut_host, int64_t* out, diga_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f3 (next row), writer half — pseudolabel_generator.py:45-49, :100-105 (`colorize_mask(...).save(...)`: 'P'-mode PNG)
 *   The IDAT payload of the PNG, made on the GPU: for each of n uint8 label maps [H, W] the zlib stream (RFC 1950) of the
 *   scanlines `filter byte 2 (Up) | label[y] - label[y-1]`, one deflate block with a static Huffman table (tuned on label maps) of literal / distance-1 match
 *   tokens, Adler-32 trailer.  out + i * capacity receives image i's stream, lengths[i] (DEVICE int64) its byte count;
 *   capacity >= diga_png_deflate_capacity(H, W) (worst case, a multiple of 16); scratch holds
 *   diga_png_deflate_scratch_bytes(n, H) bytes.  The host only frames the payload (signature, IHDR, PLTE, IDAT + CRC-32,
 *   IEND).  Decoded pixels, mode and palette equal the reference's files; the compressed bytes differ.
 * ------------------------------------------------------------------------------------------ */
int64_t diga_png_deflate_capacity(int64_t H, int64_t W);
int64_t diga_png_deflate_scratch_bytes(int64_t n, int64_t H);
int diga_png_deflate(const uint8_t* labels, int64_t n, int64_t H, int64_t W, uint8_t* out, int64_t capacity, void* scratch,
                     int64_t* lengths, diga_stream_t stream);
/* Host side of the same row: frame a finished payload (HOST memory, e.g. the pinned staging buffer) as a 'P'-mode PNG
 * file — signature, IHDR (8 bit, colour type 3), PLTE (palette_bytes = 3 * entries <= 768), IDAT with its CRC-32, IEND —
 * and write it to `path`.  No CUDA call; safe to call from several host threads at once. */
int diga_png_write_file(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                        const uint8_t* palette_host, int64_t palette_bytes);
/* CRC-32 of every image's IDAT chunk ("IDAT" + payload[i, :lengths[i]]) on the GPU (tree combination of 256 partial remainders
 * per image), and the host writer that takes it: the host threads then only frame and write bytes. */
int diga_png_crc(const uint8_t* payload, int64_t n, int64_t capacity, const int64_t* lengths, uint32_t* crc_out,
                 diga_stream_t stream);
int diga_png_write_file_crc(const char* path, const uint8_t* payload_host, int64_t length, int64_t H, int64_t W,
                            const uint8_t* palette_host, int64_t palette_bytes, uint32_t idat_crc);

#ifdef __cplusplus
}
#endif
#endif /* DIGA_B200_H_ */
