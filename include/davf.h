/*
 * davf.h -- C ABI of libdavf_sm100.so: the sm_100a kernels behind the DeepAVFusion
 * pre-training hot path (fwd + bwd of the early-fusion ViT-B masked auto-encoder).
 *
 * The reference (stoneMo/DeepAVFusion) has NO plugin / FFI layer: its seam is the Python class
 * API.  This header is therefore the interface a maintainer binds with ctypes
 * from the drop-in `DeepAVFusion` / `AVMAE` modules; every entry point
 * cites the reference code (file:line under the reference repo) whose library calls it replaces.
 *
 * Conventions
 *   - plain C types only: device pointers, int / int64_t sizes, a cudaStream_t passed as void*.
 *   - every function returns 0 on success or a negative DAVF_E* code; davf_last_error() returns
 *     a thread-local message.  Nothing throws, nothing synchronises the stream, nothing
 *     allocates or frees caller memory (the only internal state is a per-process cache of TMA
 *     descriptors keyed by pointer/shape).
 *   - "bf16" buffers are raw uint16 bfloat16; "f32" is IEEE float; indices are int64 exactly as
 *     torch produces them (avmae.py:128-133).
 *   - all matrices are row-major with an explicit leading dimension in ELEMENTS.
 */
#ifndef DAVF_H_
#define DAVF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAVF_OK 0
#define DAVF_EINVAL (-1)   /* bad argument (shape / alignment / enum)            */
#define DAVF_ECUDA (-2)    /* a CUDA runtime / driver call failed                */
#define DAVF_EUNSUPPORTED (-3)

typedef void* davf_stream_t;      /* cudaStream_t */
typedef uint16_t davf_bf16;

/* ---- library ------------------------------------------------------------------------------ */
const char* davf_last_error(void);
int davf_version(void);                       /* ABI version, currently 1                        */
int davf_device_sm(void);                     /* 100 for B200, <0 on error                       */
/* 1 (default): large launches use the CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles);
 * 0: single-CTA kernel only (A/B comparison in tests and benches). */
int davf_set_gemm_2cta(int on);
/* 0 = default: tcgen05 / TMEM / TMA attention for problems with at least 8 query rows and head dim 64 / 32, warp-level
 * mma.sync attention for the tiny fusion-token problems; 2 = mma.sync attention for every problem (A/B comparison in
 * tests and benches).  (The CUDA-core checker kernels the tests compare against live in tests/check/libdavf_check.so.) */
int davf_set_attn_impl(int impl);
/* Number of kernel launches issued by this library since process start (bench gpu_launches). */
int64_t davf_launch_count(void);
/* The same, by kernel family: 1 = CTA-pair (cta_group::2) tcgen05 GEMM, 2 = tcgen05/TMEM attention, 3 = mma.sync attention
 * (small-query cases); any other value = all.  Lets tests and benches prove which kernel variant served a call. */
int64_t davf_launch_count_kind(int kind);
/* 1 (env DAVF_PDL=1; default 0): kernels are launched with programmatic stream serialization -- each kernel's set-up
 * (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of its predecessor on the stream; memory ordering
 * is unchanged (griddepcontrol.wait precedes the first global access of every kernel).  Off by default: with the step's
 * concurrent streams it measured 2 % slower on B200 (early-resident dependents hold SMs the other streams could use). */
int davf_set_pdl(int on);

/* ---- K2: MAE random masking ----------------------------------------------------------------
 * Replaces avmae.py:127-140 (rand -> argsort -> argsort -> slice -> gather) given the noise.
 * rank-by-counting, ties lower-index-first.  noise f32 [B,L]; ids_restore i64 [B,L];
 * ids_keep i64 [B,len_keep]; mask f32 [B,L] (1 = removed). */
int davf_mask_rank(const float* noise, int B, int L, int len_keep,
                   int64_t* ids_restore, int64_t* ids_keep, float* mask, davf_stream_t s);

/* ---- K1/K3: patch rows ---------------------------------------------------------------------
 * Replaces timm PatchEmbed's conv lowering + the token gather at vits.py:93,100: writes the
 * im2col row (order c,py,px == conv weight order) of every KEPT patch as bf16.
 * img f32 [B,C,H,W]; ids_keep i64 [B,nK] or NULL (all patches, nK = gH*gW); out bf16 [B*nK, C*p*p]. */
int davf_patch_rows(const float* img, const int64_t* ids_keep, davf_bf16* out,
                    int B, int C, int H, int W, int p, int nK, davf_stream_t s);

/* f32 -> bf16 cast of M rows of width D.  Source row of output row m is
 *   (m / g) * G + off + (m % g)      (g = rows per sample in the output, G = rows per sample in
 * the source, off = first row); pass g = G = M, off = 0 for a plain cast. */
int davf_cast_rows_bf16(const float* src, davf_bf16* dst, int64_t M, int D, int g, int G, int off,
                        davf_stream_t s);

/* Gradient fan-in: out = a + b (+ c when non-NULL), f32 [n], and (when out_bf16 is non-NULL) its bf16 copy.  Replaces the
 * autograd accumulation of the gradients of a tensor with several consumers (deepavfusion.py:104-106: x_image / x_audio feed
 * their own block and the fusion block, x_fusion feeds all three) together with the cast that made the GEMM operand. */
int davf_sum_cast(const float* a, const float* b, const float* c, float* out, davf_bf16* out_bf16, int64_t n, davf_stream_t s);

/* Column sums (bias gradients; replaces autograd's sum-to-size for every Linear bias):
 * out[n] += sum_m x[m*ld + n], x bf16 [M,N], out f32 [N] (atomic accumulate). */
int davf_colsum_bf16(const davf_bf16* x, int64_t M, int N, int64_t ld, float* out, davf_stream_t s);
/* Same for an f32 source with the row window mapping of davf_cast_rows_bf16 (pos-embed /
 * fusion-token style batch reductions): out[r*D + d] (+)= sum_b x[(b*G + off + r)*D + d]. */
int davf_batchsum_f32(const float* x, int B, int G, int off, int g, int D, float* out, int accumulate,
                      davf_stream_t s);

/* ---- K4: LayerNorm -------------------------------------------------------------------------
 * Replaces every nn.LayerNorm on the path (timm Block.norm1/norm2, vits.py:116,
 * deepavfusion.py:111-113, fusion_blocks.py:281,287, avmae.py:179) and the torch.cat of
 * deepavfusion.py:104-105: the normalised rows of sample b are the n0 rows of x0 followed by
 * the n1 rows of x1 (n1 may be 0).  x0 f32 [B,n0,D] with batch stride bs0 elements (0 broadcasts
 * one copy), x1 f32 [B,n1,D] with batch stride bs1.
 * Outputs (each may be NULL): y_bf16 / y_f32 [B*(n0+n1), D]; mean, rstd f32 [B*(n0+n1)].
 * Segmented bf16 output (fusion-token groups, fusion_blocks.py:240): if nseg > 1 the bf16 row of
 * (b, r) with seg_start[k] <= r < seg_start[k+1] is  B*seg_start[k] + b*len_k + (r-seg_start[k]). */
typedef struct {
  const float* x0; int64_t bs0; int n0;
  const float* x1; int64_t bs1; int n1;
  int B; int D; float eps;
  const float* gamma; const float* beta;
  davf_bf16* y_bf16; float* y_f32; float* mean; float* rstd;
  int nseg; int seg_start[5];
} davf_ln_fwd_args;
int davf_layernorm_fwd(const davf_ln_fwd_args* a, davf_stream_t s);

/* Backward.  dy = dy_bf16 (same row layout as the forward y_bf16, segmented if nseg > 1; may be
 * NULL) + dy_f32 (natural layout; may be NULL).  dx rows go to dx0 [B,n0,D] / dx1 [B,n1,D]
 * (batch strides dbs0 / dbs1); if add0 / add1 != NULL those f32 rows (same layout as dx0 / dx1)
 * are added (fused residual-stream gradient).  dgamma / dbeta f32 [D] are atomically accumulated. */
typedef struct {
  const float* x0; int64_t bs0; int n0;
  const float* x1; int64_t bs1; int n1;
  int B; int D;
  const float* gamma; const float* mean; const float* rstd;
  const davf_bf16* dy_bf16; const float* dy_f32;
  float* dx0; int64_t dbs0; const float* add0;
  float* dx1; int64_t dbs1; const float* add1;
  float* dgamma; float* dbeta;
  int nseg; int seg_start[5];
  /* optional bf16 copies of the dx0 / dx1 rows, compact [B*n0, D] / [B*n1, D]: the next backward region consumes the
   * gradient as a bf16 GEMM operand, so the f32 -> bf16 pass over it is folded into this kernel */
  davf_bf16* dx0_bf16; davf_bf16* dx1_bf16;
} davf_ln_bwd_args;
int davf_layernorm_bwd(const davf_ln_bwd_args* a, davf_stream_t s);

/* ---- K1/K5/K9/K10: GEMM with fused epilogue --------------------------------------------------
 * Replaces every nn.Linear / conv-as-GEMM on the path (timm Attention.qkv/.proj, Mlp.fc1/.fc2,
 * fusion_blocks.py:41-42,228-230, avmae.py:32,58,62,88 and PatchEmbed.proj) and their autograd
 * backward (dgrad, wgrad).
 *
 *   acc[m,n] = sum_k A(m,k) * B(n,k)          bf16 operands, f32 accumulation (tcgen05 / TMEM)
 *   A(m,k) = a[m*lda + k]  (a_kmajor = 1)   or   a[k*lda + m]  (a_kmajor = 0, "MN-major")
 *   B(n,k) = b[n*ldb + k]  (b_kmajor = 1)   or   b[k*ldb + n]  (b_kmajor = 0)
 *     forward  y = x W^T : A = x (K-major),  B = W  (K-major)
 *     dgrad   dx = dy W  : A = dy (K-major), B = W  (MN-major, k = out-features)
 *     wgrad   dW = dy^T x: A = dy (MN-major), B = x (MN-major), k = rows
 *   z = acc + bias[n]                                   (bias may be NULL)
 *   aux_out[m*ldaux + n] = bf16(act == DAVF_ACT_GELU ? gelu_erf'(z) : z)     (if aux_out)
 *   z = gelu_erf(z)                                     (if act == DAVF_ACT_GELU)
 *   z = z * aux_in[m*ldaux + n]                         (if act == DAVF_ACT_DGELU; aux_in bf16 = the gelu_erf'(.)
 *                                                        the forward launch saved: the activation derivative costs
 *                                                        two FMAs next to GELU itself, and backward -- an epilogue-
 *                                                        bound K = 512 dgrad -- then needs no transcendental at all)
 *   row(m) = (m / g) * G + off + (m % g)                (output / residual row window; g = 0: row = m)
 *   z += res[rrow*ldres + n], rrow = res_idx ? res_idx[m] : row(m)      (if res; f32)
 *   out[row(m)*ldo + n] = z   as f32 / bf16,  or atomically += z (f32) when accumulate != 0
 * split_k > 1 requires accumulate (partial sums are reduced by f32 atomics; bias is added by
 * split 0 only).
 *   rowsum_out[m] += sum_k A(m,k)   (if rowsum_out; f32 [M], atomic).  This is the bias gradient of a
 * wgrad launch (A = dy^T): the tensor core computes it with one extra N=16 MMA per k-step against a
 * constant all-ones B tile, so no separate column-sum pass over dy is needed.
 */
enum { DAVF_ACT_NONE = 0, DAVF_ACT_GELU = 1, DAVF_ACT_DGELU = 2 };
typedef struct {
  const davf_bf16* a; int64_t lda; int a_kmajor;
  const davf_bf16* b; int64_t ldb; int b_kmajor;
  int64_t M; int64_t N; int64_t K;
  const float* bias;
  int act;
  davf_bf16* aux_out; const davf_bf16* aux_in; int64_t ldaux;
  const float* res; int64_t ldres; const int64_t* res_idx;
  void* out; int64_t ldo; int out_bf16; int accumulate;
  int g; int G; int off;
  int split_k;
  float* rowsum_out;
  int64_t* debug_clocks;   /* optional: 8 x i64 SM-clock timeline of CTA 0 (profiling aid), else NULL */
} davf_gemm_args;
/* Number of SMs the persistent GEMM grids are sized for (default: all 148; even, >= 2; returns the value in effect).
 * Data-parallel training sets 148 - k so that the NCCL all-reduce kernels overlapped with backward (util/misc.py:32-34
 * in the reference: DDP's bucketed reducer) do not push the last CTAs of a full-machine launch into a second wave. */
int davf_set_gemm_sms(int n);
int davf_gemm(const davf_gemm_args* a, davf_stream_t s);

/* Grouped launch: `count` (1..DAVF_GEMM_MAX_GROUP) INDEPENDENT problems of one operand-layout class (same a_kmajor /
 * b_kmajor) in a single kernel launch; results are identical to `count` davf_gemm calls.  Serves the many small
 * Linears of one fusion block that have no data dependence on each other: CrossAttention q / kv of both modalities
 * and the pair-attention q (fusion_blocks.py:46-52,235-252), their projections (:58,:262), and the matching dgrad /
 * wgrad launches of backward, which are latency-bound when launched one by one (M = 512 rows). */
#define DAVF_GEMM_MAX_GROUP 6
int davf_gemm_grouped(const davf_gemm_args* a, int count, davf_stream_t s);

/* ---- K6/K7/K8/K11: fused attention ----------------------------------------------------------
 * Replaces F.scaled_dot_product_attention in timm Attention (encoder d=64, decoder d=32) and the
 * explicit softmax(q k^T * scale) v of fusion_blocks.py:53-57,254-258.
 * One problem per (batch b, head h): q rows  q + b*q_bs + i*q_rs + h*dqk   (i < Nq),
 * k rows  k + b*k_bs + j*k_rs + h*dqk,  v rows  v + b*v_bs + j*v_rs + h*dv  (j < Nk),
 * o rows  o + b*o_bs + i*o_rs + h*dv.   All bf16; strides in elements, so packed qkv buffers and
 * query sub-ranges (live rows only) need no copies.
 * lse f32 [B,H,Nq] = log-sum-exp of the scaled logits (saved for backward).
 * accumulate != 0: o += result (used to add the two factorised pair attentions): 1 = read-modify-write
 * (launches adding into one buffer must be ordered by the caller), 2 = bf16x2 atomic adds into a zero-initialised buffer
 * (such launches may run concurrently on different streams).
 * Supported (dqk, dv): (64,64), (32,32), (16,64).  Nk <= 256.  q / k / v rows must be 16-byte aligned
 * (strides multiples of 8 elements). */
typedef struct {
  const davf_bf16* q; int64_t q_bs; int64_t q_rs;
  const davf_bf16* k; int64_t k_bs; int64_t k_rs;
  const davf_bf16* v; int64_t v_bs; int64_t v_rs;
  davf_bf16* o; int64_t o_bs; int64_t o_rs;
  float* lse;
  int B; int H; int Nq; int Nk; int dqk; int dv;
  float scale; int accumulate;
} davf_attn_fwd_args;
int davf_attention_fwd(const davf_attn_fwd_args* a, davf_stream_t s);

/* Backward: given dO (layout of o) writes dq / dk / dv with the layouts of q / k / v (separate
 * stride sets so gradients can land in a packed dqkv buffer).  accumulate_dq != 0: dq += (1 / 2 as for the forward).
 * The softmax-Jacobian row term D_i is computed from the recomputed probabilities (exact, used for the
 * tiny fusion-token attentions) unless the forward output `o` is supplied (see the struct). */
typedef struct {
  const davf_bf16* q; int64_t q_bs; int64_t q_rs;
  const davf_bf16* k; int64_t k_bs; int64_t k_rs;
  const davf_bf16* v; int64_t v_bs; int64_t v_rs;
  const davf_bf16* d_o; int64_t do_bs; int64_t do_rs;
  const float* lse;
  davf_bf16* dq; int64_t dq_bs; int64_t dq_rs;
  davf_bf16* dk; int64_t dk_bs; int64_t dk_rs;
  davf_bf16* dv_; int64_t dv_bs; int64_t dv_rs;
  int B; int H; int Nq; int Nk; int dqk; int dv;
  float scale; int accumulate_dq;
  /* optional forward output (layout o_bs / o_rs, heads packed): if given, D_i = dO_i . O_i (one pass, the
   * flash-attention form); if NULL, D_i = sum_j P_ij dP_ij from a first pass over the recomputed scores. */
  const davf_bf16* o; int64_t o_bs; int64_t o_rs;
  /* dq_dead_rows > 0: the rows [-dq_dead_rows, 0) in front of dq's first row (same strides, every head's dqk columns) are
   * zero-filled: the query slots of the fusion-token prefix rows, whose block outputs the reference discards
   * (deepavfusion.py:104-105), inside a packed dqkv buffer that the qkv dgrad GEMM reads in full. */
  int dq_dead_rows;
} davf_attn_bwd_args;
int davf_attention_bwd(const davf_attn_bwd_args* a, davf_stream_t s);

/* ---- K3: decoder sequence assembly ----------------------------------------------------------
 * Replaces avmae.py:161-169 (mask-token append, unshuffle gather, +pos_embed, cat fusion).
 * e f32 [B*nK, D] (embedded kept tokens), ef f32 [B*nF, D] (embedded fusion tokens),
 * mask_token f32 [D], pos f32 [L, D], ids_restore i64 [B, L]  ->  seq f32 [B, nF+L, D]. */
int davf_decoder_assemble_fwd(const float* e, const float* ef, const float* mask_token, const float* pos,
                              const int64_t* ids_restore, float* seq, int B, int nK, int nF, int L, int D,
                              davf_stream_t s);
/* dseq f32 [B,nF+L,D] -> de bf16 [B*nK, D] (rows gathered through ids_keep), def bf16 [B*nF, D],
 * dmask_token f32 [D] (+=), dpos f32 [L, D] (+=). */
int davf_decoder_assemble_bwd(const float* dseq, const int64_t* ids_keep, const int64_t* ids_restore,
                              davf_bf16* de, davf_bf16* def_, float* dmask_token, float* dpos,
                              int B, int nK, int nF, int L, int D, davf_stream_t s);

/* ---- input stage: log-mel spectrograms and frame normalisation on the GPU -------------------------
 * Replaces the CPU data-loader transforms of train.py:44-54: torchaudio MelSpectrogram(sample_rate, n_fft = 800,
 * hop_length = 250, n_mels) with torchaudio's defaults (periodic Hann window, centre / reflect padding, power 2, HTK mel
 * scale, no filter normalisation) followed by util/audio_transforms.py Log (log10(x + eps)), optionally preceded by
 * RandomVol's gain + clamp (audio_transforms.py:8-18) with the per-clip gain drawn by the caller; and torchvision
 * ToTensor + Normalize.  int16 PCM / uint8 frames cross PCIe instead of fp32 tensors.
 * workspace: davf_logmel_workspace_bytes() bytes of device memory owned by the caller, filled once by davf_logmel_init
 * (twiddles, window, filter bank).  wave: exactly one of wave_f32 / wave_i16 (PCM, / 32768), [B, T]; gain_db f32 [B] or
 * NULL; out f32 [B, 1, n_mels, frames], frames <= T / hop + 1 (datasets.py:242 keeps T / hop). */
int64_t davf_logmel_workspace_bytes(void);
int davf_logmel_init(void* workspace, int sample_rate, int n_fft, int hop, int n_mels, davf_stream_t s);
int davf_logmel_fwd(const void* workspace, const float* wave_f32, const int16_t* wave_i16, const float* gain_db, int B, int T,
                    int n_mels, int frames, float eps, float* out, davf_stream_t s);
/* src u8 [B, H, W, C] (device) -> dst f32 [B, C, H, W] = (src / 255 - mean[c]) / std[c]; mean / std are HOST arrays of C floats. */
int davf_image_normalize_u8(const uint8_t* src, float* dst, int B, int H, int W, int C, const float* mean, const float* std_,
                            davf_stream_t s);

/* ---- K12: normalised masked patch-MSE --------------------------------------------------------
 * Replaces avmae.py:201-214 (patchify) + :183-198 (forward_loss).  img f32 [B,C,H,W];
 * pred f32 rows: patch (b,l) is at pred + (b*pred_G + pred_off + l)*P, P = p*p*C (in-patch order
 * py,px,c); mask f32 [B,L].  loss_sum f32 [1] += sum over masked patches of mean((pred-t)^2);
 * the caller divides by the masked count (avmae.py:197). */
int davf_masked_mse_fwd(const float* img, const float* pred, const float* mask, float* loss_sum,
                        int B, int C, int H, int W, int p, int pred_G, int pred_off, int norm_pix,
                        davf_stream_t s);
/* dpred bf16 [B*L, P] = gscale[0] * inv_count * mask * 2 (pred - t) / P  (zeros for kept patches). */
int davf_masked_mse_bwd(const float* img, const float* pred, const float* mask, const float* gscale,
                        float inv_count, davf_bf16* dpred,
                        int B, int C, int H, int W, int p, int pred_G, int pred_off, int norm_pix,
                        davf_stream_t s);

/* ---- K13/K14: fused AdamW + grad-norm + zero_grad + bf16 weight refresh ---------------------------
 * Replaces torch.optim.AdamW (train.py:93; misc.py:126-130), the /accum_iter sweep (misc.py:114-119),
 * get_grad_norm_ (misc.py:151-163), zero_grad, and the per-step bf16 weight casts of autocast.
 * Works on flat f32 buffers p, g, m, v of n elements (n % 64 == 0).  Hyper-parameters live in DEVICE
 * tables so LR schedules do not re-capture CUDA graphs and any parameter grouping is free:
 *   chunk_group u8 [n/64]  group id of each 64-element chunk (255 = frozen: no update),
 *   hp f32 [ngroups*2] = {lr, weight_decay} per group,
 *   scal f32 [4] = {beta1^t, beta2^t, grad_scale (1/accum or 1/(accum*world)), unused}.
 * p_bf16 (may be NULL) receives bf16(p_new); g is zeroed when zero_grad != 0; sumsq_out (may be
 * NULL) += sum((g*grad_scale)^2) over the trainable chunks (the global grad-norm^2, no host sync). */
int davf_adamw_step(float* p, float* g, float* m, float* v, davf_bf16* p_bf16, int64_t n,
                    const uint8_t* chunk_group, const float* hp, const float* scal,
                    float beta1, float beta2, float eps, int zero_grad, float* sumsq_out, davf_stream_t s);
/* out[0] += sum(g^2) over n elements (global grad-norm; one pass, no host sync). */
int davf_sumsq_f32(const float* g, int64_t n, float* out, davf_stream_t s);
/* Plain f32 -> bf16 cast of a flat buffer (weight refresh after load_state_dict). */
int davf_cast_flat_bf16(const float* src, davf_bf16* dst, int64_t n, davf_stream_t s);

/* ---- a11: classifier tail (reference models/classifier.py:42-59) -------------------------------
 * x.mean(dim=1) over the tokens, BatchNorm1d(embed_dim, affine=False, eps=1e-6) on the pooled features
 * (classifier.py:15-18,50-54; training mode updates running_mean / running_var with `momentum`, unbiased variance,
 * as torch does) and the three nn.Linear heads (:20-22,56-58).  All f32 (the reference runs lin-probe / fine-tune
 * with use_amp = False).  C (classes) is arbitrary: 310, 527, ... */
int davf_meanpool_fwd(const float* x, int64_t batch_stride, int B, int n, int D, float* out, davf_stream_t s);
int davf_meanpool_bwd(const float* dy, int B, int n, int D, float* dx, davf_stream_t s);
int davf_batchnorm1d_fwd(const float* x, int B, int D, int training, float* running_mean, float* running_var,
                         float momentum, float eps, float* y, float* save_mean, float* save_rstd, davf_stream_t s);
int davf_batchnorm1d_bwd(const float* dy, const float* x, const float* mean, const float* rstd, int B, int D, int training,
                         float* dx, davf_stream_t s);
/* y[b,c] = x[b,:] . W[c,:] + bias[c];  backward: dW += dy^T x, db += colsum dy (both optional), dx = dy W (optional) */
int davf_head_fwd(const float* x, const float* W, const float* bias, int B, int C, int D, float* y, davf_stream_t s);
int davf_head_bwd(const float* dy, const float* x, const float* W, int B, int C, int D, float* dW, float* db, float* dx,
                  davf_stream_t s);

/* ---- stochastic depth (timm DropPath on the residual branches; models/vits.py:33, models/fusion_blocks.py:276,283,288;
 * only configs/finetune.yaml:36 sets drop_path > 0) --------------------------------------------------------------
 * scale f32 [B] = keep-mask / keep_prob per sample; rows are sample-major with rows_per_sample rows each.
 *   davf_scale_rows_add: out = res + scale[r / rows_per_sample] * y          (f32 [rows, D]; the residual add of a dropped branch)
 *   davf_scale_rows    : dst = scale[r / rows_per_sample] * src  as f32 and / or bf16 (either may be NULL): the branch gradient */
int davf_scale_rows_add(const float* res, const float* y, const float* scale, int rows_per_sample, int64_t rows, int D,
                        float* out, davf_stream_t s);
int davf_scale_rows(const float* src, const float* scale, int rows_per_sample, int64_t rows, int D, float* dst_f32,
                    davf_bf16* dst_bf16, davf_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* DAVF_H_ */
