/* cds.h -- C ABI of the B200 diffusion-sampling engine (libcds.so).
 *
 * The reference (CleanDiffuser) has no FFI of its own: its hot path is the Python loop
 * DiscreteDiffusionSDE.sample / ContinuousDiffusionSDE.sample (cleandiffuser/diffusion/diffusionsde.py:401-606,
 * :743-952) calling model["diffusion"](x_t, t, cond) -> ATen kernels once per op.  This header is the
 * boundary a maintainer would bind instead (ctypes stub in INTEGRATION.md): the per-iteration work of that
 * loop is described ONCE as a short program of fused operators, and the engine replays the program for
 * all reverse iterations from a CUDA graph with x_t resident on the device and no host round trip.
 *
 *   operator            replaces (reference file:line)
 *   ------------------  ------------------------------------------------------------------------------
 *   CDS_OP_CONV         nn.Conv1d / ConvTranspose1d / Linear (+bias) followed by GroupNorm1d + Mish, the
 *                       time/FiLM conditioning and the residual add of a ResidualBlock / ChiResidualBlock
 *                       (nn_diffusion/jannerunet.py:21-36,52-69, chiunet.py:13-45, utils/building_blocks.py:60-76),
 *                       the Linear(+Mish/SiLU/GELU) layers of DQLMlp and DiT1d (dqlmlp.py:22-29, dit.py:25-29,43-45)
 *   CDS_OP_LNMOD        LayerNorm(no affine) + adaLN modulate (dit.py:10-11,19,21,33,35,49)
 *   CDS_OP_ATTN         nn.MultiheadAttention core softmax(QK^T/sqrt(hd))V per (trajectory, head) (dit.py:20,34)
 *   CDS_OP_PREP         consistency-model re-noise + c_in pre-scale (consistency_model.py:257,423)
 *   CDS_OP_CAST         x.permute(0,2,1) entry of the UNets (jannerunet.py:169, chiunet.py:142): layout/dtype hand-over of x_t
 *   CDS_OP_UPDATE       CFG combine, clip_prediction, eps<->x0 conversion, the 8 solver updates, fix_mask
 *                       (diffusionsde.py:202,208-223,539-592) and the CM skip/out combine (consistency_model.py:257-262)
 *
 * Conventions
 *  - every function returns 0 on success, a negative cds_status otherwise; cds_last_error() gives the
 *    message (thread local).  No C++ exception or Python object crosses this ABI.
 *  - all pointers inside operator descriptors are DEVICE pointers owned by the caller (weights are the
 *    caller's parameter storage or caller-allocated packed copies; activations live in a caller-allocated
 *    workspace).  The library never allocates device memory on the hot path and never frees caller memory.
 *  - activations are "channels last": element (b, l, c) of a (batch, L, C) tensor sits at
 *    base + b*bstride + l*lstride + c (strides in elements), which is also the reference's public (b, horizon, dim)
 *    layout, so x_t needs no permute on entry or exit.  x_t, predictions and all tables are fp32; intermediate
 *    activations are fp32 (CDS_MATH_FP32 and CDS_MATH_TF32_TC programs) or bf16 (CDS_MATH_BF16_TC programs).
 *  - "per-iteration" operands are indexed by a device-resident iteration counter, so one captured graph
 *    serves every reverse iteration:  vec(b, c) = step[iter*step_stride + c] + sample[b*sample_stride + c]
 *    (either part may be NULL = 0).
 *  - a plan is not thread safe; distinct plans may be used concurrently.  All work is enqueued on the
 *    caller's stream (e.g. torch.cuda.current_stream()) and is asynchronous.
 */
#ifndef CDS_H_
#define CDS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDS_ABI_VERSION 5

typedef enum cds_status {
  CDS_OK = 0,
  CDS_ERR_INVALID = -1,      /* bad argument / unsupported shape (message says which) */
  CDS_ERR_CUDA = -2,         /* a CUDA runtime call failed */
  CDS_ERR_UNSUPPORTED = -3,  /* valid request the kernels cannot serve; caller should fall back */
  CDS_ERR_STATE = -4         /* call order violated (e.g. run before finalize) */
} cds_status;

typedef enum cds_op_kind { CDS_OP_CONV = 0, CDS_OP_UPDATE = 1, CDS_OP_LNMOD = 2, CDS_OP_ATTN = 3, CDS_OP_PREP = 4,
                           CDS_OP_CAST = 5 } cds_op_kind;
typedef enum cds_act { CDS_ACT_NONE = 0, CDS_ACT_MISH = 1, CDS_ACT_SILU = 2, CDS_ACT_GELU_TANH = 3,
                       CDS_ACT_MISH_SILU = 4 /* silu(mish(x)): DiT's map_emb tail feeding every adaLN (dit.py:26,43,71) */,
                       CDS_ACT_LEAKY_RELU = 5 /* negative slope 0.01 (nn.LeakyReLU default; pearcemlp.py:44) */,
                       CDS_ACT_GELU_ERF = 6 /* exact GELU 0.5 x (1 + erf(x / sqrt 2)) (nn.GELU default; pearcemlp.py:30) */ } cds_act;
/* math mode of CDS_OP_CONV:
 *   CDS_MATH_FP32     fp32 CUDA-core FMA (bit-faithful to the fp32 oracle up to summation order)
 *   CDS_MATH_BF16_TC  tcgen05 tensor cores, bf16 operands and bf16 inter-layer activations, fp32 TMEM accumulation
 *   CDS_MATH_TF32_TC  tcgen05 tensor cores kind::tf32: fp32 activations and weights in memory, read by the MMA as TF32
 *                     (10-bit mantissa), fp32 accumulation -- the arithmetic of the reference's own GPU path
 *                     (torch.backends.cudnn.allow_tf32 = True is PyTorch's default for convolutions)            */
typedef enum cds_math { CDS_MATH_FP32 = 0, CDS_MATH_BF16_TC = 1, CDS_MATH_TF32_TC = 2 } cds_math;
/* element type of an activation tensor.  CDS_TF32 is fp32 STORAGE whose values are rounded to the nearest TF32 (10-bit
 * mantissa, ties away from zero, like cvt.rna.tf32.f32) when an operator writes them: tcgen05 kind::tf32 ignores the low 13
 * mantissa bits of its fp32 operands (truncation), so an operand that is already rounded is read exactly and its rounding
 * error is unbiased and half as large.  Producers honour it on store; consumers read CDS_TF32 exactly like CDS_F32. */
typedef enum cds_dtype { CDS_F32 = 0, CDS_BF16 = 1, CDS_TF32 = 2 } cds_dtype;
/* update shapes; must match cleandiffuser_b200/diffusion/solvers.py */
typedef enum cds_update_kind { CDS_UPD_DDPM = 0, CDS_UPD_DDIM = 1, CDS_UPD_EPS = 2, CDS_UPD_X = 3, CDS_UPD_X2M = 4, CDS_UPD_CM = 5,
  /* ContinuousEDM (newedm.py:142-148, :411-431).  D = K0*x + K1*net (c_skip, c_out), clipped to [x_min, x_max] when final_clip;
   *   CDS_UPD_EDM       Euler step   d = (x - D)/SIGMA;  x <- x - d*K2 (K2 = sigma_i - sigma_{i-1});  when K4 != 0 the step is
   *                     the predictor of a Heun step: the old x goes to xhat_prev and d to aux
   *   CDS_UPD_EDM_HEUN  corrector    d' = (x - D)/SIGMA (x = the predictor's result, SIGMA = sigma_{i-1});
   *                     x <- xhat_prev - (aux + d')/2 * K2                                                              */
  CDS_UPD_EDM = 6, CDS_UPD_EDM_HEUN = 7 } cds_update_kind;

/* per-iteration coefficient row (floats), one row per reverse iteration */
enum { CDS_ROW_ALPHA = 0, CDS_ROW_SIGMA = 1, CDS_ROW_K0 = 2, CDS_ROW_K1 = 3, CDS_ROW_K2 = 4, CDS_ROW_K3 = 5,
       CDS_ROW_K4 = 6, CDS_ROW_KIND = 7, CDS_ROW_NOISE = 8 /* 1 + noise slot, 0 = no draw */, CDS_ROW_T = 9,
       CDS_ROW_XW = 10, CDS_ROW_DW = 11 /* EDM kinds: when XW != 0 the slope is XW*x - DW*D (the legacy EDM archetecture's
                                           x_weight / D_weight, edm.py:152) instead of (x - D)/SIGMA */,
       CDS_ROW_FLOATS = 12 };

/* vec(b, c) = step[iter*step_stride + c] + sample[b*sample_stride + c] */
typedef struct cds_vec {
  const float* step;   int64_t step_stride;
  const float* sample; int64_t sample_stride;
} cds_vec;

/* Fused 1-D convolution / linear layer as an implicit GEMM over rows (b, l_out):
 *   acc(b,l,n)  = sum_{tap,ci} in(b, l*stride + tap - pad, ci) * w[tap][ci][n]          (zero outside [0, L_in))
 *   y = acc + bias(b,n)
 *   if groups > 0:  y = GroupNorm(y; stats over L_out x (C_out/groups) per (b, group)) * gamma + beta
 *   y = act(y)
 *   y = y * scale(b,n) + shift(b,n)                 (each optional)
 *   y += res(b,l,n)                                  (identity shortcut, optional)
 *   y += sum_ci res_in(b,l,ci) * res_w[ci][n] + res_bias[n]   (1x1-conv shortcut, optional)
 *   out(b, l*phases + n / C_out, n % C_out) = y      (phases = 2 writes a stride-2 transposed conv's two
 *                                                     output phases; N = C_out*phases columns in w)           */
typedef struct cds_conv_op {
  int32_t batch, L_in, L_out, C_in, C_out, taps, stride, pad, phases;
  int32_t in_batch_mod;            /* >0: read in() at batch index b % in_batch_mod (CFG branches share x_t) */
  const void* in;   int64_t in_bstride;  int32_t in_lstride;   /* strides in ELEMENTS of the tensor's dtype */
  const void*  w;                  /* CDS_MATH_FP32: fp32 [taps*C_in][C_out*phases] (K rows, N contiguous)
                                      CDS_MATH_BF16_TC: bf16 [taps][C_out][C_in]     (K contiguous, TMA/UMMA K-major)
                                      CDS_MATH_TF32_TC: fp32 [taps][C_out][C_in]     (same packing, fp32 elements) */
  cds_vec bias;
  int32_t groups; const float* gn_gamma; const float* gn_beta; float gn_eps;
  int32_t act;
  cds_vec scale, shift;
  const void* res; int64_t res_bstride; int32_t res_lstride; int32_t res_batch_mod;
  const void* res_in; int64_t res_in_bstride; int32_t res_in_lstride; int32_t res_C;
  const void* res_w; const float* res_bias;   /* res_w: fp32 [res_C][C_out] (FP32) or bf16 / fp32 [C_out][res_C] (TC modes), as `w` */
  void* out; int64_t out_bstride; int32_t out_lstride;
  int32_t math;                    /* cds_math: which kernel family / weight layout */
  int32_t in_dtype, out_dtype, res_dtype, res_in_dtype;   /* cds_dtype of the activation tensors */
  /* > 1: the `sample` parts of bias / scale / shift belong to row-batch index b / sample_row_div (Linear layers over a token
   * stream flattened to batch*L rows of length-1 "sequences": one vector per trajectory = per L tokens).  Tensor-core path only. */
  int32_t sample_row_div;
} cds_conv_op;

/* out(b,l,:) = LayerNorm(in(b,l,:), eps, no affine) * (1 + scale(b,:)) + shift(b,:) */
typedef struct cds_lnmod_op {
  int32_t batch, L, C; float eps;
  const float* in; void* out;                   /* dense (batch, L, C); in fp32, out of out_dtype */
  const float* shift; const float* scale; int64_t mod_bstride;   /* per-trajectory vectors */
  int32_t out_dtype;                            /* cds_dtype: bf16 feeds the tensor-core Linear that follows */
} cds_lnmod_op;

/* qkv dense (batch, L, 3*C) with [q | k | v] column blocks, heads split C evenly; out dense (batch, L, C) */
typedef struct cds_attn_op {
  int32_t batch, L, C, heads;
  const void* qkv; void* out;
  int32_t out_dtype;                            /* cds_dtype of out */
  int32_t qkv_dtype;                            /* cds_dtype of qkv: bf16 (head_dim 32, L <= 128) runs on tensor cores (mma.sync) */
} cds_attn_op;

/* dense fp32 (batch, L, C_in) -> dense (batch, L, C_out) of out_dtype, channels [C_in, C_out) zero: gives x_t the
 * channel-padded form the tensor-core conv reads through TMA (the UNets' first conv has C_in = obs+act dims, e.g. 14:
 * 32 bf16 channels for CDS_MATH_BF16_TC, 16 fp32 channels for CDS_MATH_TF32_TC -- rows must be multiples of 16 bytes) */
typedef struct cds_cast_op {
  int32_t batch, L, C_in, C_out;
  const float* in; void* out;
  int32_t out_dtype;             /* cds_dtype of out */
} cds_cast_op;

/* consistency model: if row.NOISE: x += K2 * noise[slot];  xin = K3 * x */
typedef struct cds_prep_op {
  int32_t batch, row;
  float* x; float* xin; const float* noise; const float* coef;
} cds_prep_op;

/* One reverse-process update on x (batch, row) in place; `coef` is the [n_iters][CDS_ROW_FLOATS] table.
 *   p = w_cfg*pred + w_uncond*pred_uncond       (if pred_uncond; w_uncond = 1-w_cfg)
 *   p = clip(p)  (eps-prediction: to [(x-alpha*x_max)/sigma, (x-alpha*x_min)/sigma]; x0-prediction: [x_min,x_max])
 *   x <- solver update of kind row.KIND with noise slot row.NOISE  ;  x = x*(1-mask) + prior*mask            */
typedef struct cds_update_op {
  int32_t batch, row;
  float* x;
  const float* pred; const float* pred_uncond; float w_cfg; float w_uncond;  /* w and (1-w), both rounded from double by the host */
  const float* noise;            /* [n_slots][...] pre-drawn N(0,1): slot s of this op's rows starts at noise + s*noise_slot_stride */
  int64_t noise_slot_stride;     /* floats between slots; 0 = batch*row (the op covers the whole tape) */
  const float* prior; const float* mask;      /* mask: `row` floats or NULL */
  const float* x_min; const float* x_max;     /* `row` floats or NULL */
  float* xhat_prev;              /* (batch,row) history for the 2M solvers / the EDM Heun corrector, or NULL */
  float* aux;                    /* (batch,row) second history buffer (EDM Heun: the predictor's slope) or NULL */
  const float* coef;
  int32_t predict_noise;
  int32_t final_clip;            /* CM only: clip the combined prediction to [x_min, x_max] */
  /* optional bf16 copy of the new x_t in the channel-padded form the tensor-core UNets read (what CDS_OP_CAST produces):
   * element (r, c) of the dense (rows, cast_C_in) view of x goes to x_cast[r*cast_C_out + c]; pad channels are never
   * written (a CDS_OPF_ONCE cast zeroes them and converts the initial x_t).  NULL = no copy. */
  void* x_cast; int32_t cast_C_in; int32_t cast_C_out;
  int32_t x_cast_dtype;          /* cds_dtype of x_cast */
} cds_update_op;

/* cds_op.flags */
enum { CDS_OPF_ONCE = 1 /* run once per cds_plan_run, before its first iteration, instead of in every iteration */,
       CDS_OPF_BRANCH_SHIFT = 8, CDS_OPF_BRANCH_MASK = 0xff00
       /* bits 8..15: branch index.  Operators of one branch run in program order; different branches are independent
        * chains (disjoint trajectories) that the engine enqueues on parallel streams / parallel graph branches and joins at
        * the end of every iteration, so that one chain's kernel boundaries overlap with the other chain's kernels. */ };

typedef struct cds_op {
  int32_t kind;                  /* cds_op_kind */
  int32_t flags;
  union { cds_conv_op conv; cds_update_op update; cds_lnmod_op lnmod; cds_attn_op attn; cds_prep_op prep; cds_cast_op cast; } u;
} cds_op;

typedef struct cds_plan cds_plan;

int         cds_version(void);
/* sizeof(cds_op) as compiled into the library -- bindings assert their struct mirror against it */
int         cds_op_size(void);
const char* cds_last_error(void);
/* number of SMs etc. of `device`, -1 on error; used by the host to size workspaces */
int         cds_device_sm_count(int device);

/* 1 if the tensor-core kernel of op->math (CDS_MATH_BF16_TC or CDS_MATH_TF32_TC; anything else is read as BF16_TC) can run
 * `op` as described (dtypes, strides, shapes; `w` may still be NULL), else 0: the host lowering asks before choosing the
 * weight layout; ops that are not eligible run on the CUDA-core kernel with math = CDS_MATH_FP32 (which accepts fp32 or
 * bf16 activations).  Pure host logic. */
int cds_conv_tc_supported(const cds_conv_op* op);

/* A plan = the per-iteration operator program for one (model, shape, option set) on one device. */
int cds_plan_create(int device, cds_plan** out);
int cds_plan_destroy(cds_plan* plan);
/* Append operators (copied).  Pointers inside must stay valid until the plan is destroyed or rebuilt. */
int cds_plan_append(cds_plan* plan, const cds_op* ops, int32_t n_ops);
/* Validate shapes, choose kernels, allocate the 4-byte device iteration counter. n_iters = rows in coef tables. */
int cds_plan_finalize(cds_plan* plan, int32_t n_iters);
/* Enqueue iterations [first, first+count) on `stream` (a cudaStream_t): CDS_OPF_ONCE operators first, then `count`
 * replays of the iteration program.  The first call captures the program into a CUDA graph; later calls replay it.
 * use_graph = 0 launches kernels directly (debug/ncu).  Consecutive tensor-core conv kernels are chained with
 * programmatic dependent launch (the next kernel's prologue and weight prefetch overlap the previous kernel's tail);
 * CDS_PDL=0 in the environment switches that off. */
int cds_plan_run(cds_plan* plan, int32_t first, int32_t count, void* stream, int32_t use_graph);
/* Single-step entry ("cds_step"): enqueue operators [op_first, op_first + op_count) of the program (indices into the appended
 * operator list; CDS_OPF_ONCE operators inside the range are skipped) for iteration `iter` with direct launches -- the device
 * iteration counter is set to `iter` first.  This is how a host callback is interleaved with the loop (classifier guidance,
 * diffusionsde.py:153-173: the denoiser operators run, PyTorch adds the guidance term to the prediction in place, then the
 * update operator runs); `cds_plan_run(plan, first, 0, ...)` runs the CDS_OPF_ONCE operators alone, and
 * `cds_plan_run(plan, i, 1, ...)` is the whole iteration i. */
int cds_plan_run_range(cds_plan* plan, int32_t iter, int32_t op_first, int32_t op_count, void* stream);
/* Run iteration `iter` once with direct launches, bracketing every operator with CUDA events on `stream`;
 * ms_per_op[i] receives the device time of operator i (n_ops entries; CDS_OPF_ONCE operators are run and timed first).
 * Synchronises `stream`.  For bench.py's live per-kernel roofline; note that it advances x_t like a normal iteration. */
int cds_plan_profile(cds_plan* plan, int32_t iter, void* stream, float* ms_per_op, int32_t n_ops);
/* Number of kernel launches one iteration of the program performs (for bench.py's gpu_launches). */
int cds_plan_launches_per_iter(const cds_plan* plan);

/* Debug: record a clock64 timeline (64 slots per CTA, layout in csrc/conv_tc.cuh) of the `target_launch`-th tensor-core conv
 * launch issued from now on into `device_buffer` (int64 entries); NULL switches tracing off.  Returns the grid size of the
 * launch traced since the previous call (0 if none). */
int cds_debug_trace(void* device_buffer, int64_t capacity_entries, int32_t target_launch);

/* Stand-alone operator launch (parity tests of single kernels): runs `op` once with iteration index `iter`. */
int cds_run_op(int device, const cds_op* op, int32_t iter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CDS_H_ */
