/*
 * otb200.h — C ABI of libotb200.so, the sm_100a implementation of the per-step
 * training hot path that OneTrainer owns itself.
 *
 * The reference (sangoi-exe/OneTrainer) is 100 % Python and has no FFI today; its
 * plugin API for this path is three Python seams:
 *   S1  LoRAModule.forward                 modules/module/LoRAModule.py:319-323
 *   S2  ModelSetupNoiseMixin / DiffusionMixin / FlowMatchingMixin / DiffusionLossMixin
 *                                          modules/modelSetup/mixin/*.py
 *   S3  patch_adamw / EMAModuleWrapper     modules/util/optimizer/adamw_extensions.py:202-205,
 *                                          modules/module/EMAModule.py:37-54
 * Every entry point below cites the reference lines it replaces.  The thin Python host
 * (onetrainer_b200/) binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - no allocation, no host synchronisation, no implicit stream: the caller passes
 *     a cudaStream_t (as void*) and any workspace;
 *   - return 0 on success, a negative OTB_E* code on failure; otb_last_error()
 *     returns a thread-local message;
 *   - the library keeps no reference to caller memory after the call returns
 *     (kernels read it asynchronously on the given stream).
 */
#ifndef OTB200_H
#define OTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTB_ABI_VERSION 6   /* 2: otb_lora_linear_grouped; pack_w / pred_pack_w appended to the noise / loss argument structs.  3: otb_grad_exchange, otb_grad_norm_from_partials.
                               4: OTB_TS_TABLE (ts_cdf / ts_cdf_len appended to otb_noise_args); otb_lora_grad_batched; otb_lora_linear_grouped takes up to 32 members; otb_lora_linear_args.flags; otb_adafactor_multi / otb_came_multi; otb_dora_compose / otb_dora_backward.
                               5: otb_shard_reduce / otb_shard_gather / otb_peer_barrier / otb_peer_signal / otb_peer_wait, otb_adamw_sr_ema_multi_peers (K-z)
                               6: otb_stochastic_round (the helpers of bf16_stochastic_rounding.py as stand-alone launches) */

/* error codes */
#define OTB_OK            0
#define OTB_EINVAL       -1   /* bad argument (shape, alignment, dtype, null pointer) */
#define OTB_EUNSUPPORTED -2   /* legal in the reference but not implemented here: fails loudly, never falls back */
#define OTB_ECUDA        -3   /* CUDA runtime/driver error; message carries cudaGetErrorString */
#define OTB_EARCH        -4   /* device is not sm_100 */

/* dtype enum (matches the host-side mirror) */
#define OTB_F32  0
#define OTB_BF16 1
#define OTB_F16  2

/* Arithmetic flavour bits.  torch's CPU and CUDA eager kernels round the AdamW
 * update differently in three places; the oracle and the kernel implement both so
 * parity can be pinned against the reference run on CPU (golden fixtures) AND the
 * reference run on the GPU (torch CUDA eager).  See DESIGN.md §numerics. */
#define OTB_FLAVOR_DIV_RECIPROCAL   1u   /* sqrt(v)/c computed as sqrt(v)*(1/c)  (CUDA eager div-by-scalar)  */
#define OTB_FLAVOR_ADDCDIV_MUL_LAST 2u   /* p + a*(m/d) (CUDA)  vs  p + (a*m)/d (CPU)                          */
#define OTB_FLAVOR_FMA_LERP         4u   /* m + w*(g-m) as one fma (both torch back-ends do)                   */
#define OTB_FLAVOR_FMA_ADDCMUL      8u   /* v + ((1-b2)*g)*g as fma(.., g, v)                                  */
#define OTB_FLAVOR_FMA_ADDCDIV      16u  /* p + a*q as fma(a, q, p)                                            */
#define OTB_FLAVOR_FMA_ADD_ALPHA    32u  /* Adam L2: g + wd*p as fma(wd, p, g)                                 */
#define OTB_FLAVOR_ADDCMUL_SQUARE_FIRST 64u /* v + (1-b2)*(g*g) (CUDA) vs v + ((1-b2)*g)*g (CPU)                 */
#define OTB_FLAVOR_ALPHA_IN_DTYPE   128u /* Adam L2 on bf16 tensors: alpha itself is rounded to bf16 first (CPU) */
/* measured with tools/probe_torch_flavor.py against torch 2.11 eager on each back-end */
#define OTB_FLAVOR_TORCH_CPU   (OTB_FLAVOR_FMA_LERP | OTB_FLAVOR_FMA_ADDCMUL | OTB_FLAVOR_FMA_ADD_ALPHA | OTB_FLAVOR_ALPHA_IN_DTYPE)
#define OTB_FLAVOR_TORCH_CUDA  (OTB_FLAVOR_DIV_RECIPROCAL | OTB_FLAVOR_ADDCDIV_MUL_LAST | OTB_FLAVOR_FMA_LERP | \
                                OTB_FLAVOR_FMA_ADDCMUL | OTB_FLAVOR_FMA_ADDCDIV | OTB_FLAVOR_FMA_ADD_ALPHA | \
                                OTB_FLAVOR_ADDCMUL_SQUARE_FIRST)

const char* otb_last_error(void);
int         otb_abi_version(void);
/* sm count / compute capability of the current device; OTB_EARCH if not 10.x */
int         otb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------
 * Counter-based RNG shared by every kernel (and restated in oracle/otb_oracle.c):
 * Philox4x32-10, key = 64-bit seed, counter = (group_lo, group_hi, offset_lo,
 * stream<<24 | offset_hi[0:24)).  Streams keep the draws of one step disjoint.
 * Indexing by GLOBAL sample / element index makes an N-GPU run draw exactly what a
 * 1-GPU run at the same global batch draws (SURVEY.md §5 "RNG under DP").
 * ---------------------------------------------------------------------------------- */
#define OTB_STREAM_NOISE     0u
#define OTB_STREAM_OFFSET    1u
#define OTB_STREAM_PERTURB   2u
#define OTB_STREAM_TIMESTEP  3u
#define OTB_STREAM_SR        4u
#define OTB_STREAM_DROPOUT   5u

/* ------------------------------------------------------------------------------------
 * K-c  multi-tensor AdamW / Adam step with bf16 stochastic rounding, fused EMA update,
 *      fused gradient-clip scaling and optional bf16 shadow emission.
 * Replaces: step_adamw_parameter  modules/util/optimizer/adamw_extensions.py:17-150
 *           step_adam_parameter   modules/util/optimizer/adam_extensions.py (L2 variant)
 *           copy_stochastic_/addcdiv_stochastic_  modules/util/bf16_stochastic_rounding.py:5-61
 *           EMAModuleWrapper.step modules/module/EMAModule.py:37-54
 *           nn.utils.clip_grad_norm_ scaling  modules/trainer/GenericTrainer.py:712-713
 * One launch updates every tensor in the table.
 * ---------------------------------------------------------------------------------- */
typedef struct otb_optim_tensor {
    void*       p;            /* parameter, dtype p_dtype, updated in place                 */
    const void* g;            /* gradient, same dtype as p                                  */
    void*       exp_avg;      /* state, same dtype as p                                     */
    void*       exp_avg_sq;   /* state, same dtype as p                                     */
    void*       ema;          /* EMA shadow (same dtype as p) or NULL                       */
    void*       shadow_bf16;  /* optional bf16 copy of the updated p (fp32 masters) or NULL */
    int64_t     numel;
    int64_t     rng_base;     /* global element offset of this tensor in the SR stream      */
    /* per-tensor scalars, already rounded the way torch rounds python scalars to float     */
    float       decay_mul;    /* 1 - lr*weight_decay   (AdamW)   | 1 (Adam)                 */
    float       l2;           /* 0 (AdamW)             | weight_decay (Adam: g += l2*p)     */
    float       lerp_w;       /* 1 - beta1                                                  */
    float       beta2;
    float       one_minus_beta2;
    float       bc2_sqrt;     /* sqrt(1 - beta2^step)                                       */
    float       inv_bc2_sqrt; /* (float)(1.0 / bc2_sqrt_double): torch CUDA divides by a python scalar as a * (float)(1/b) */
    float       eps;
    float       neg_step_size;/* -lr / (1 - beta1^step)                                     */
    int32_t     p_dtype;      /* OTB_F32 | OTB_BF16                                         */
    int32_t     flags;        /* bit0: stochastic rounding of p; bit1: maximize (g = -g)    */
} otb_optim_tensor;

/* per-step values a captured CUDA graph must re-read on every replay (device memory) */
typedef struct otb_optim_dyn {
    uint64_t seed, offset;             /* SR Philox key / step number                        */
    float    ema_one_minus_decay;
    int32_t  do_ema;
} otb_optim_dyn;

typedef struct otb_optim_launch {
    const otb_optim_tensor* tensors;   /* device array [n_tensors]                          */
    const int32_t* chunk_tensor;       /* device [n_chunks] tensor index of each chunk      */
    const int32_t* chunk_index;        /* device [n_chunks] chunk number inside that tensor */
    int32_t  n_tensors;
    int32_t  n_chunks;
    int32_t  chunk_elems;              /* elements per chunk (multiple of 2048)             */
    uint32_t flavor;                   /* OTB_FLAVOR_* bits                                 */
    uint64_t seed;                     /* SR Philox key                                     */
    uint64_t offset;                   /* SR Philox offset: the optimizer step number       */
    float    ema_one_minus_decay;      /* 1 - min((1+s)/(10+s), decay); used iff do_ema     */
    int32_t  do_ema;                   /* 1 on EMA update steps ((s+1) % interval == 0)     */
    const float* clip_coef;            /* device scalar: g is scaled by it first, or NULL   */
    const otb_optim_dyn* dyn;          /* optional device override of seed/offset/ema_*     */
} otb_optim_launch;

int otb_adamw_sr_ema_multi(const otb_optim_launch* args, void* stream);

/* EMA-only multi-tensor step (optimizer not ours, e.g. third-party optimizers).
 * Replaces EMAModuleWrapper.step modules/module/EMAModule.py:43-54.  Uses p/ema/numel/p_dtype of the table. */
int otb_ema_multi(const otb_optim_launch* args, void* stream);

/* ------------------------------------------------------------------------------------
 * Global gradient norm + clip coefficient.
 * Replaces nn.utils.clip_grad_norm_(parameters, max_norm)  GenericTrainer.py:712-713
 *   total_norm = ||g||_2 over all tensors; coef = min(1, max_norm / (total_norm + 1e-6)).
 * partials: device float[n_chunks] workspace; out: device float[2] = {total_norm, coef * grad_prescale}.
 * Deterministic (fixed-order two-level reduction).  grad_prescale multiplies g before
 * the norm (1/world_size under data parallel when the allreduce summed) and is folded into
 * the returned scale factor, so one multiply both averages and clips.
 * ---------------------------------------------------------------------------------- */
int otb_grad_norm_multi(const otb_optim_launch* args, float max_norm, float grad_prescale,
                        float* partials, float* out_norm_coef, void* stream);
/* the same finish step over squared-norm partials produced elsewhere (otb_grad_exchange) */
int otb_grad_norm_from_partials(const float* partials, int32_t n_partials, float max_norm, float grad_prescale,
                                float* out_norm_coef, void* stream);
/* in-place g *= coef for callers that want the reference's two-pass behaviour */
int otb_grad_scale_multi(const otb_optim_launch* args, const float* coef, void* stream);

/* ------------------------------------------------------------------------------------
 * DoRA (weight-decomposed LoRA): the weight-sized part of
 * DoRAModule.forward  modules/module/LoRAModule.py:385-419  and of its backward.
 *     V = W + BA * scale ;  norm = ||V|| per input-channel group (or per output row) (+ eps), detached ;
 *     WP = dora_scale * (V / norm)
 * The activation-sized GEMMs (y = x WP^T, dX = dY WP) run on otb_lora_linear with rank 0; BA = B A, G = dY^T x and the
 * rank-r products of the backward are plain library GEMMs on the caller's side.  [N, J] matrices are row-major; column j
 * belongs to group j / group_size (group_size = kh * kw of a conv leaf, 1 for Linear).
 * ---------------------------------------------------------------------------------- */
typedef struct otb_dora_args {
    const void*  w;           /* [N, J] frozen base weight, w_dtype                                                    */
    const float* ba;          /* [N, J] fp32: lora_up . lora_down (make_weight, :76-91), unscaled                      */
    const float* dora_scale;  /* [J / group_size] (input axis) or [N] (output axis), fp32                              */
    float*       norm;        /* same shape: written by otb_dora_compose, read by otb_dora_backward                     */
    float*       scratch;     /* max(N, 64 * J) floats (row-slab partials of the column passes)                        */
    void*        wp;          /* compose: [N, J] bf16 out                                                              */
    void*        wpt;         /* compose: [J, N] bf16 out (the K-major operand of the dX launch), or NULL              */
    const void*  G;           /* backward: [N, J] bf16, dL/dWP                                                         */
    float*       dV;          /* backward: [N, J] fp32 out, dL/dV with the norm held constant (:396-400)               */
    float*       d_scale;     /* backward: gradient of dora_scale, fp32 out                                            */
    int32_t N, J, group_size, out_axis, w_dtype;
    float scale;              /* alpha / rank                                                                          */
    float eps;                /* torch.finfo(fp32).eps when norm_epsilon, else 0                                       */
} otb_dora_args;

int otb_dora_compose(const otb_dora_args* args, void* stream);
int otb_dora_backward(const otb_dora_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * Factored optimizers: multi-tensor Adafactor and CAME, one launch per step,
 * one CTA per tensor, every reduction in a fixed order, no host read-back.
 * Replaces: step_adafactor_parameter  modules/util/optimizer/adafactor_extensions.py:17-93
 *           CAME.step_parameter       modules/util/optimizer/CAME.py:80-176
 *           add_stochastic_           modules/util/bf16_stochastic_rounding.py:33-45
 * A tensor with >= 2 dimensions is a batch of [rows, cols] matrices (batch = product of the leading dimensions) with a
 * row factor [batch, rows] and a column factor [batch, cols]; a 1-D tensor keeps an unfactored second moment
 * (batch = rows = 1, cols = numel).  State tensors are fp32, as in the reference.
 * ---------------------------------------------------------------------------------- */
#define OTB_FACTORED_FACTORED        1   /* sq_row / sq_col (and res_row / res_col) instead of sq                     */
#define OTB_FACTORED_FIRST_MOMENT    2   /* Adafactor: beta1 is set (exp_avg present); CAME always has one            */
#define OTB_FACTORED_SCALE_PARAMETER 4   /* Adafactor: lr_t = max(eps1, RMS(p)) * lr                                  */
#define OTB_FACTORED_SR              8   /* CAME: bf16 parameter written with stochastic rounding (add_stochastic_)   */

typedef struct otb_factored_tensor {
    void*       p;          /* parameter, p_dtype                                                                     */
    const void* g;          /* gradient, p_dtype                                                                      */
    float*      exp_avg;    /* [numel] first moment, or NULL                                                          */
    float*      sq_row;     /* [batch, rows]  exp_avg_sq_row                                                          */
    float*      sq_col;     /* [batch, cols]  exp_avg_sq_col                                                          */
    float*      sq;         /* [numel]        exp_avg_sq (not factored)                                               */
    float*      res_row;    /* CAME: exp_avg_res_row                                                                  */
    float*      res_col;    /* CAME: exp_avg_res_col                                                                  */
    float*      rms_out;    /* optional [1]: RMS(p) of this step (state["RMS"])                                       */
    int64_t numel;
    int64_t rng_base;       /* global element index of element 0 in the stochastic-rounding stream                    */
    int64_t scratch_offset; /* this tensor's `batch` floats inside otb_factored_launch.scratch                        */
    int32_t batch, rows, cols;
    int32_t p_dtype, flags;
    float beta1;            /* Adafactor beta1 / CAME betas[0]                                                        */
    float beta2;            /* Adafactor beta2t = 1 - step^decay_rate / CAME betas[1]                                 */
    float beta3;            /* CAME betas[2]                                                                          */
    float one_minus_beta1, one_minus_beta2, one_minus_beta3;   /* formed in double on the host: 1 - (float)0.9999 is off by 1.7e-4 */
    float eps0, eps1, clip_threshold;
    float lr;               /* Adafactor: the (relative) step size before parameter scaling; CAME: lr                 */
    float weight_decay;
} otb_factored_tensor;

typedef struct otb_factored_launch {
    const otb_factored_tensor* tensors;   /* device array                                                             */
    float*   scratch;                     /* device, sum of `batch` over the tensors                                  */
    int32_t  n_tensors;
    int32_t  cluster;                     /* CTAs per tensor (one thread-block cluster each): 1, 2, 4 or 8; 0 = 1       */
    uint64_t seed, offset;                /* CAME stochastic rounding: the weight-decay draw uses `offset`, the update
                                             draw `offset + 1` (`offset` when weight_decay == 0)                      */
} otb_factored_launch;

int otb_adafactor_multi(const otb_factored_launch* args, void* stream);
int otb_came_multi(const otb_factored_launch* args, void* stream);

/* ------------------------------------------------------------------------------------
 * K-r  data-parallel gradient exchange fused with the squared gradient norm, over NVLink
 *      peer memory ("one exchange step"; the reference has no multi-GPU
 *      code, so this replaces what a DDP wrapper around GenericTrainer.py:692-714 would do
 *      with ncclAllReduce + nn.utils.clip_grad_norm_'s norm pass).
 * The flat fp32 gradient arena is symmetric memory: every replica is mapped into every
 * process and (optionally) all of them as one NVLS multicast object.  Rank r sums slice r
 * over all replicas (P2P loads in rank order, or multimem.ld_reduce inside the switch),
 * accumulates (grad_prescale * sum)^2, and writes the sum back into every replica in place
 * (P2P stores, or multimem.st); block partials of the squared norm land in
 * float[world * blocks] on every rank (feed them to otb_grad_norm_from_partials).  Peer
 * barriers on flag words (uint32[2 * OTB_MAX_RANKS] per rank, zero-initialised,
 * peer-mapped) bracket the kernel; every rank must make the same call in the same order.
 * No NCCL, no host sync; capturable in a CUDA graph.
 * ---------------------------------------------------------------------------------- */
#define OTB_MAX_RANKS 16
#define OTB_EXCHANGE_P2P  0    /* world in {2, 4, 8}: (n-1)/n of the arena per NVLink direction   */
#define OTB_EXCHANGE_NVLS 1    /* any world <= OTB_MAX_RANKS: (1 + 1/n) of the arena per direction */
typedef struct otb_grad_exchange_args {
    float*    peer_grads[OTB_MAX_RANKS];    /* arena of rank i as mapped into THIS process         */
    float*    peer_partials[OTB_MAX_RANKS]; /* float[world * blocks] of rank i, likewise           */
    uint32_t* peer_flags[OTB_MAX_RANKS];    /* flag words of rank i, likewise                      */
    void*     mc_grads;                     /* NVLS only: multicast address of the arena           */
    float*    mc_partials;                  /* NVLS only: multicast address of the partial table   */
    int64_t   n_elems;                      /* floats in the arena, multiple of 4                  */
    int32_t   rank, world;
    int32_t   blocks;                       /* otb_grad_exchange_blocks(n_elems, world)            */
    int32_t   mode;                         /* OTB_EXCHANGE_*                                      */
    float     grad_prescale;                /* folded into the norm only; the arena holds the SUM  */
} otb_grad_exchange_args;
int otb_grad_exchange_blocks(int64_t n_elems, int32_t world);
int otb_grad_exchange(const otb_grad_exchange_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * K-z  the SHARDED form of the exchange for full fine-tuning (SURVEY.md §8f rank 3:
 *      "optimizer-in-backward mode ... with bucketed reduce-scatter"; hook point
 *      GenericTrainer.py:531-558, optimizer adamw_extensions.py:17-150, EMA EMAModule.py:37-54):
 *      reduce-scatter -> K-c on the owner's tensors only -> all-gather of the updated
 *      parameters, all over NVLink peer memory, no NCCL.
 * The flat gradient arena AND the flat parameter arena are symmetric memory.  A contiguous,
 * tensor-aligned element range [begin, end) has one owner rank.
 *   otb_shard_reduce  (owner only) sums the range over every replica — P2P: one 16-byte load
 *       per peer, added in rank order in fp32; NVLS: multimem.ld_reduce (.acc::f32 for bf16) —
 *       rounds once, keeps the sum in the OWNER's replica and writes the block partials of
 *       (grad_prescale * sum)^2 into float[partial_slot + block] of EVERY rank's partial table
 *       (all tables complete after the following barrier -> otb_grad_norm_from_partials);
 *   otb_shard_gather  (owner only) copies the range of the owner's replica into every
 *       replica (P2P stores / one multimem.st stream);
 *   otb_peer_barrier  all ranks: flag-word barrier, `phase` in [0, OTB_PEER_PHASES) selects the
 *       word set (uint32[OTB_PEER_PHASES * OTB_MAX_RANKS] per rank, zero-initialised);
 *   otb_peer_signal / otb_peer_wait  one-directional: a non-owner raises its word in the
 *       owner's table ("my backward has produced this range"), the owner consumes one word per
 *       peer — lets the reduce of a bucket overlap the rest of backward.
 * Bounded spins (trap after ~60 s).  Capturable in a CUDA graph; no host sync.
 * ---------------------------------------------------------------------------------- */
#define OTB_PEER_PHASES 4
typedef struct otb_shard_exchange_args {
    void*     peer_base[OTB_MAX_RANKS];     /* the arena of rank i as mapped into THIS process       */
    float*    peer_partials[OTB_MAX_RANKS]; /* partial table of rank i, likewise (reduce, P2P mode)  */
    void*     mc_base;                      /* NVLS only: multicast address of the arena             */
    float*    mc_partials;                  /* NVLS only: multicast address of the partial table     */
    int64_t   begin, end;                   /* element range, whole 16-byte vectors                  */
    int32_t   dtype;                        /* OTB_BF16 | OTB_F32                                    */
    int32_t   rank, world;
    int32_t   mode;                         /* OTB_EXCHANGE_*                                        */
    int32_t   blocks;                       /* grid size; the launch writes `blocks` partials        */
    int32_t   partial_slot;                 /* first index of this launch's partials in the table    */
    float     grad_prescale;                /* folded into the norm only; the arena holds the SUM    */
} otb_shard_exchange_args;
/* K-c with the all-gather folded in: otb_adamw_sr_ema_multi on the owner's tensors, every parameter store replicated into
 * the other ranks' replicas of the symmetric parameter arena (P2P stores, or one multimem.st) — the step's updated
 * parameters need no otb_shard_gather pass afterwards, only the closing otb_peer_barrier.  Every tensor's `p` must point
 * into this rank's replica; OTB_FLAVOR_TORCH_CUDA only. */
typedef struct otb_peer_store {
    int64_t delta[OTB_MAX_RANKS];   /* byte offset from MY replica of the arena to rank i's mapping (0 for my own rank)   */
    int64_t mc_delta;               /* NVLS: byte offset from my replica to the multicast mapping                         */
    int32_t rank, world;
    int32_t mode;                   /* OTB_EXCHANGE_*                                                                      */
    int32_t reserved;
} otb_peer_store;
int otb_adamw_sr_ema_multi_peers(const otb_optim_launch* args, const otb_peer_store* peers, void* stream);
int otb_shard_reduce(const otb_shard_exchange_args* args, void* stream);
int otb_shard_gather(const otb_shard_exchange_args* args, void* stream);
int otb_peer_barrier(uint32_t* const* peer_flags, int32_t rank, int32_t world, int32_t phase, void* stream);
int otb_peer_signal(uint32_t* owner_word, void* stream);
int otb_peer_wait(uint32_t* my_words, int32_t rank, int32_t world, void* stream);

/* ------------------------------------------------------------------------------------
 * bf16 shadow refresh: for every LoRA tensor (fp32 or bf16 master, [rows, cols])
 * emit a bf16 copy and a bf16 transposed copy, one launch for all leaves.
 * (No reference counterpart: under autocast the reference re-casts A/B on every
 *  call, modules/util/dtype_util.py:44-48; here the cast is hoisted to once per step.)
 * ---------------------------------------------------------------------------------- */
typedef struct otb_shadow_tensor {
    const void* src;      /* master [rows, cols], dtype src_dtype */
    void*       dst;      /* bf16 [rows, cols] or NULL            */
    void*       dst_t;    /* bf16 [cols, rows] or NULL            */
    int32_t     rows, cols;
    int32_t     src_dtype;
    int32_t     tile_base; /* first 32x32 tile index of this tensor in the launch */
} otb_shadow_tensor;
int otb_shadow_refresh_multi(const otb_shadow_tensor* tensors, int32_t n_tensors,
                             int32_t n_tiles, void* stream);

/* ------------------------------------------------------------------------------------
 * K-b1  noise + timestep + add-noise (+ target), the pre-network half of the
 *       memory-bound step kernel.
 * Replaces: _create_noise           ModelSetupNoiseMixin.py:18-49
 *           _get_timestep_discrete  ModelSetupNoiseMixin.py:51-155 (UNIFORM, LOGIT_NORMAL,
 *                                   HEAVY_TAIL, static + dynamic shift, deterministic;
 *                                   COS_MAP / SIGMOID :122-153 as an inverse-CDF draw over
 *                                   the reference's weight table)
 *           _add_noise_discrete     ModelSetupDiffusionMixin.py:15-38 (DDPM tables)
 *                                   ModelSetupFlowMatchingMixin.py:14-39 (rectified flow)
 *           latent scaling          BaseStableDiffusionXLSetup.py:215, BaseFluxSetup.py:230
 *           v / flow targets        BaseStableDiffusionXLSetup.py:285, BaseFluxSetup.py:307
 * ---------------------------------------------------------------------------------- */
#define OTB_TS_UNIFORM       0
#define OTB_TS_LOGIT_NORMAL  1
#define OTB_TS_HEAVY_TAIL    2
#define OTB_TS_DETERMINISTIC 3   /* int(T*0.5)-1 */
#define OTB_TS_GIVEN         4   /* timesteps supplied by the caller in `timestep` */
#define OTB_TS_TABLE         5   /* COS_MAP / SIGMOID: t = min_timestep + first i with ts_cdf[i] > u, u ~ U[0,1) from the
                                    timestep stream at the GLOBAL sample index (the reference calls torch.multinomial
                                    on the same weights; ts_cdf is their normalised running sum)           */

#define OTB_TARGET_NONE    0
#define OTB_TARGET_EPSILON 1     /* target = noise                               */
#define OTB_TARGET_V       2     /* target = coef_x0*noise - coef_noise*x0       */
#define OTB_TARGET_FLOW    3     /* target = noise - x0                          */

typedef struct otb_noise_args {
    /* tensors; B = batch, chw = C*hw */
    const void* latent;        /* [B, chw] latent image, dtype latent_dtype                           */
    const void* noise_in;      /* optional externally supplied noise (parity mode), latent_dtype or NULL */
    void*       x0_out;        /* optional scaled latent (latent - shift)*scale, latent_dtype         */
    void*       noise_out;     /* optional noise, latent_dtype                                        */
    void*       noisy_out;     /* optional noisy latent, latent_dtype                                 */
    void*       noisy_cast_out;/* optional noisy latent in cast_dtype (the network's train dtype)     */
    void*       target_out;    /* optional target, latent_dtype                                       */
    int32_t*    timestep;      /* [B] int32: output, or input when ts_mode == OTB_TS_GIVEN            */
    float*      sigma_out;     /* optional [B]: coef_noise[t] (flow: sigma)                           */
    const float* coef_x0;      /* [T] DDPM: sqrt(alphas_cumprod);     flow: 1 - sigma                 */
    const float* coef_noise;   /* [T] DDPM: sqrt(1 - alphas_cumprod); flow: sigma = (t+1)/T           */
    int32_t B, C, hw;
    int32_t latent_dtype, cast_dtype;
    int32_t target, ts_mode;
    int32_t num_train_timesteps;   /* T */
    int32_t generate_noise;        /* 1: draw from Philox; 0: read noise_in                           */
    int32_t do_add_noise;          /* 0: only noise / timesteps are produced                          */
    /* latent transform: x0 = (latent - shift) * scale */
    float latent_scale, latent_shift;
    /* noise options (_create_noise) */
    float offset_noise_weight, perturbation_noise_weight;
    /* timestep options (_get_timestep_discrete), pre-reduced on the host exactly as python does */
    int32_t min_timestep, max_timestep;   /* int(T*min_noising_strength), int(T*max_noising_strength) */
    float noising_weight, noising_bias;
    float shift_c1;                /* (float)(T * shift)                                              */
    float shift_c2;                /* (float)(shift - 1)                                              */
    /* RNG */
    uint64_t seed, offset;
    int64_t  sample_base;          /* global index of sample 0 of this rank's batch (DP invariance)   */
    const uint64_t* rng_dev;       /* optional device {seed, offset}: overrides the two scalars above so a
                                      captured CUDA graph draws fresh noise on every replay              */
    int32_t pack_w;                /* 0: noisy_cast_out is [B, C, H, W].  W > 0 (the latent width; H = hw / W): it is
                                      written in FLUX's packed layout [B, (H/2)(W/2), 4C] — FluxModel.pack_latents,
                                      modules/model/FluxModel.py:319-331 — so the network input needs no permute copy  */
    const float* ts_cdf;           /* OTB_TS_TABLE: [ts_cdf_len] non-decreasing, last entry 1.0 (device memory)          */
    int32_t ts_cdf_len;            /* max_timestep - min_timestep                                                        */
} otb_noise_args;

int otb_noise_addnoise(const otb_noise_args* args, void* stream);

/* ------------------------------------------------------------------------------------
 * K-b2  loss forward + backward in one pass: per-sample masked MSE/MAE with optional
 *       mask-area normalisation and prior-preservation term, loss scaler, concept
 *       weight, timestep weighting, and d(loss)/d(predicted) for an ASSUMED upstream
 *       gradient (1/(B*GA) — what `.mean()` then `/GA` produces).
 * Replaces: _diffusion_losses / _flow_matching_losses  ModelSetupDiffusionLossMixin.py:233-321
 *           __masked_losses :36-117, __unmasked_losses :119-168
 *           masked_losses_with_prior  modules/util/loss/masked_loss.py:21-45
 *           __snr/__min_snr_weight/__debiased_estimation_weight/__p2_loss_weight/
 *           __sigma_loss_weight :170-231
 *           and the autograd backward of that chain (GenericTrainer.py:692-696)
 * ---------------------------------------------------------------------------------- */
#define OTB_LW_CONSTANT   0
#define OTB_LW_MIN_SNR    1
#define OTB_LW_DEBIASED   2
#define OTB_LW_P2         3
#define OTB_LW_SIGMA      4

typedef struct otb_loss_args {
    const void*  predicted;     /* [B, chw] pred_dtype                                     */
    const void*  target;        /* [B, chw] target_dtype                                   */
    const void*  prior_target;  /* [B, chw] target_dtype or NULL                           */
    const void*  mask;          /* [B, hw] mask_dtype or NULL (unmasked)                   */
    const float* loss_weight;   /* [B] per-concept weight or NULL (=1)                     */
    const int32_t* timestep;    /* [B] or NULL                                             */
    const float* sqrt_ac;       /* [T] (for snr weights) or NULL                           */
    const float* sqrt_1mac;     /* [T]                                                     */
    const float* sigmas;        /* [T] (t+1)/T for OTB_LW_SIGMA or NULL                    */
    float*       losses;        /* [B] fp32 out                                            */
    void*        dpred;         /* [B, chw] pred_dtype out, or NULL (forward only)         */
    float*       partials;      /* workspace float[B * blocks_per_sample]                  */
    uint32_t*    counters;      /* workspace uint32[B], zero on entry, zero on exit        */
    float*       mask_stats;    /* workspace float[B*2]: mean(clamp(mask)), mean(1-clamp); filled by the call when needed */
    int32_t B, C, hw;
    int32_t pred_dtype, target_dtype, mask_dtype;
    int32_t masked;             /* 1: __masked_losses; 0: __unmasked_losses                */
    int32_t normalize_masked_area_loss;
    int32_t loss_weight_fn;     /* OTB_LW_*                                                */
    int32_t v_prediction;
    int32_t num_train_timesteps;
    int32_t blocks_per_sample;
    float mse_strength, mae_strength;
    float unmasked_weight, masked_prior_preservation_weight;
    float loss_scale;           /* batch_size_scale * gradient_accumulation_steps_scale    */
    float loss_weight_strength; /* gamma                                                   */
    float upstream;             /* assumed d(total)/d(losses[b]); 1/(B*GA)                 */
    int32_t pred_pack_w;        /* 0: predicted / dpred are [B, C, H, W].  W > 0: both are in FLUX's packed layout
                                   [B, (H/2)(W/2), 4C] (the transformer's output, before FluxModel.unpack_latents
                                   modules/model/FluxModel.py:333-344); target / mask stay [B, C, H, W]           */
} otb_loss_args;

int otb_loss_fwd_bwd(const otb_loss_args* args, void* stream);
/* grid heuristic shared with the host so it can size `partials`: blocks per sample for a [B, C*hw] problem */
int otb_loss_blocks_per_sample(int32_t B, int32_t C, int32_t hw);
/* fix-up used by autograd.backward: dpred[b,:] *= grad_out[b] / assumed, skipped on the
 * device (no write) when they are equal — the normal case. */
int otb_loss_bwd_rescale(void* dpred, int32_t dtype, const float* grad_out, float assumed,
                         int32_t B, int64_t chw, void* stream);

/* ------------------------------------------------------------------------------------
 * K-a  LoRA down/up projections fused with the frozen base Linear / 1x1 conv.
 *      tcgen05 (UTCHMMA) tiles, fp32 accumulators in TMEM, operands fed by TMA.
 * Replaces LoRAModule.forward  modules/module/LoRAModule.py:319-323
 *      y = x W^T + bias + ((s * x A^T)_bf16 [dropout]) B^T
 * and its autograd backward:
 *      dx = dy W + ((s * dy B)_bf16 [dropout]) A     -- the same kernel, called with the
 *                                                       transposed bf16 shadows
 *      dA = Gs^T x,  dB = dy^T Ts                    -- otb_lora_grad_ab
 * All operands bf16, row-major, K-major ("x[M,K] . w[N,K]^T").
 * ---------------------------------------------------------------------------------- */
typedef struct otb_lora_linear_args {
    const void* x;        /* [M, K]  bf16                                                   */
    const void* w;        /* [N, K]  bf16 frozen base weight (fwd) / W^T shadow [K_in, N_out] (bwd) */
    const void* down;     /* [R, K]  bf16 LoRA down (fwd: A; bwd: B^T)                      */
    const void* up;       /* [N, R]  bf16 LoRA up   (fwd: B; bwd: A^T)                      */
    const void* bias;     /* [N] bias_dtype or NULL                                         */
    void*       y;        /* [M, N]  bf16 out                                               */
    void*       t_out;    /* [M, R]  bf16 out: (s * x down^T) after dropout, or NULL        */
    int32_t M, N, K, R;
    int32_t bias_dtype;
    float   scale;        /* alpha / rank                                                   */
    float   dropout_p;    /* 0 disables                                                     */
    uint64_t seed, offset;/* dropout Philox key (same in fwd and bwd => same mask)          */
    int32_t tile_n;       /* 0 = let the library choose                                     */
    int32_t stages;       /* 0 = let the library choose                                     */
    int32_t flags;        /* OTB_LINEAR_STATIC_WEIGHTS: w / down / up / bias were NOT written by the launch that
                             precedes this one on the stream (frozen base weights, shadows refreshed after the last
                             optimizer step).  The kernel may then fetch its first weight tiles BEFORE the stream
                             dependency on that launch resolves (programmatic dependent launch), hiding their HBM
                             latency behind the predecessor's tail; x is always read after it.              */
} otb_lora_linear_args;
#define OTB_LINEAR_STATIC_WEIGHTS 1

int otb_lora_linear(const otb_lora_linear_args* args, void* stream);

/* Up to OTB_LORA_GROUP_MAX problems of the SAME shape (M, N, K, rank, dropout_p) as ONE launch: members[i] is an ordinary
 * otb_lora_linear_args (own x, w, down, up, bias, y, t_out, scale, dropout key).  The reference calls the q / k / v
 * projections of an attention as three separate module forwards (diffusers Attention -> LoRAModule.forward
 * modules/module/LoRAModule.py:319-323, three times) and autograd runs their three dX GEMMs one after another;
 * batched, the tiles of all members share one persistent grid.  The wide form (more than three members) is for the K / V
 * projections of the text context, which read the SAME tensor in every cross-attention of the network (140 leaves in
 * SDXL, M = 77 * batch): one launch per 32 of them instead of one 18 us launch per pair.  Results are bit-identical to
 * `count` single calls. */
#define OTB_LORA_GROUP_MAX 32
int otb_lora_linear_grouped(const otb_lora_linear_args* members, int32_t count, void* stream);

/* dA[R,K] += Gs[M,R]^T x[M,K] ;  dB[N,R] += dy[M,N]^T Ts[M,R]   (fp32; the outputs must hold zero or the running
 * accumulation of earlier micro-batches).  Two of the ~7 GEMMs autograd runs per leaf for LoRAModule.forward
 * (modules/module/LoRAModule.py:319-323).  The sum over M is formed in a fixed order: results are bit-identical from
 * run to run. */
typedef struct otb_lora_grad_args {
    const void* x;     /* [M, K] bf16 */
    const void* dy;    /* [M, N] bf16 */
    const void* gs;    /* [M, R] bf16 */
    const void* ts;    /* [M, R] bf16 */
    float*      dA;    /* [R, K] fp32, or NULL */
    float*      dB;    /* [N, R] fp32, or NULL */
    int32_t M, N, K, R;
} otb_lora_grad_args;

/* one leaf, M not split (one CTA per 128 output columns) */
int otb_lora_grad_ab(const otb_lora_grad_args* args, void* stream);

/* Up to OTB_GRAD_BATCH_MAX leaves of ANY shapes as ONE persistent launch: dA / dB feed nothing until the optimizer, so
 * the host queues the leaves of a backward pass and flushes them in batches (722 latency-bound launches per SDXL step
 * in the first version).  `workspace` (device memory, 256-byte aligned, at least otb_lora_grad_workspace_bytes(leaves,
 * count) bytes, or NULL) lets small batches and very long M be split over more CTAs; split partial sums are combined
 * in split order by the last CTA to arrive.  Its first 64 KB are arrival counters: zero before the FIRST use, left
 * zero by every launch.  `chained` != 0 declares that the launch issued immediately before this one on `stream` is
 * another otb_lora_grad_batched launch whose outputs this one neither reads nor writes and which uses a DIFFERENT
 * workspace: the two may then overlap (the second fills the SMs the first vacates at its tail). */
#define OTB_GRAD_BATCH_MAX 32
int otb_lora_grad_batched(const otb_lora_grad_args* leaves, int32_t count, void* workspace, int64_t workspace_bytes,
                          int32_t chained, void* stream);
int64_t otb_lora_grad_workspace_bytes(const otb_lora_grad_args* leaves, int32_t count);

/* ------------------------------------------------------------------------------------
 * The reference's bf16 stochastic-rounding helpers as stand-alone launches — what an optimizer of the reference that is
 * NOT fused here (it imports these three functions) calls per tensor.
 * Replaces: copy_stochastic_      modules/util/bf16_stochastic_rounding.py:5-30
 *           add_stochastic_       modules/util/bf16_stochastic_rounding.py:33-45   (call sites CAME.py:167-176)
 *           addcdiv_stochastic_   modules/util/bf16_stochastic_rounding.py:48-61   (call site adamw_extensions.py:139-144)
 * One launch instead of 5-7, no fp32 / int32 temporaries:
 *   COPY     r = src (fp32)
 *   ADD      r = fp32(src) + alpha * fp32(target)        — alpha multiplies `target` (= the reference's `input`): the
 *            reference computes `result = other.float(); result.add_(input, alpha=alpha)`, whatever its docstring says
 *   ADDCDIV  r = fp32(target) + alpha * src / src2       — rounded as the `flavor` bits say (OTB_FLAVOR_FMA_ADD_ALPHA,
 *            OTB_FLAVOR_ADDCDIV_MUL_LAST, OTB_FLAVOR_FMA_ADDCDIV: torch's CPU and CUDA kernels differ, see above)
 * then target = high half of (int32(r) + 16 random bits) — element i draws the bits of element rng_base + i of the
 * OTB_STREAM_SR Philox stream keyed by (seed, offset), the assignment K-c uses.  `src` / `src2` may alias `target`.
 * ---------------------------------------------------------------------------------- */
#define OTB_SR_COPY    0
#define OTB_SR_ADD     1
#define OTB_SR_ADDCDIV 2
typedef struct otb_sr_args {
    void*       target;      /* bf16 [n]: read by ADD / ADDCDIV, always written                   */
    const void* src;         /* COPY: fp32 source; ADD: `other`; ADDCDIV: tensor1                 */
    const void* src2;        /* ADDCDIV: tensor2 (same dtype as src), else NULL                   */
    int64_t     n;
    int64_t     rng_base;    /* position of target[0] in the SR stream; a multiple of 8           */
    uint64_t    seed, offset;
    int32_t     mode;        /* OTB_SR_*                                                          */
    int32_t     src_dtype;   /* OTB_F32 | OTB_BF16 (COPY: OTB_F32)                                */
    float       alpha;       /* ADD: alpha; ADDCDIV: value                                        */
    uint32_t    flavor;      /* OTB_FLAVOR_* bits                                                 */
} otb_sr_args;
int otb_stochastic_round(const otb_sr_args* args, void* stream);

/* Utility used by tests and the data-parallel bucket path: fp32 <-> bf16 casts with
 * round-to-nearest-even, and a device-side Philox dump for known-answer tests. */
/* sizeof() of every argument struct, in declaration order: optim_tensor, optim_launch, shadow_tensor,
 * noise_args, loss_args, lora_linear_args, lora_grad_args, grad_exchange_args, factored_tensor, factored_launch, dora_args,
 * shard_exchange_args, peer_store, sr_args.  Bindings assert these at load time. */
int otb_struct_sizes(int32_t* out14);

int otb_philox_dump(uint64_t seed, uint64_t offset, uint32_t stream_id, int64_t first_group,
                    int64_t n_groups, uint32_t* out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTB200_H */
