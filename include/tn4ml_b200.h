/*
 * tn4ml_b200 -- C ABI of the B200-native tn4ml hot path.
 *
 * The reference (bsc-quantic/tn4ml v1.1.1) is pure Python and has no FFI of its
 * own; the seam this library replaces is the pair
 *     loss_fn(data, targets, *params) -> scalar          tn4ml/models/model.py:723-751
 *     value_and_grad(params, data, targets)              tn4ml/models/model.py:554-566
 * plus the forward-only closures of evaluate() (model.py:1048-1076) and
 * predict()/forward() (model.py:278-368).  Every entry point below is what a
 * jax.ffi / ctypes binding for that seam would bind.
 *
 * Rules of the boundary
 *   - plain pointers and sizes only; every buffer (inputs, outputs, workspace)
 *     is DEVICE memory owned by the caller; the library never allocates, frees
 *     or retains device memory and never synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - every call returns 0 on success, non-zero otherwise; the message is read
 *     with tnb_last_error() (thread-local); no exceptions cross the boundary;
 *   - tnb_init(device) sets every per-device kernel attribute; after it (or after the first call on that device,
 *     which does the same under a lock) tnb_forward / tnb_backward / tnb_adam_update are re-entrant (no mutable
 *     globals on the call path, no environment reads: measurement knobs are read once when the library is loaded)
 *     and CUDA-graph capturable (tests/test_gpu_boundary.py captures and replays a step).
 */
#ifndef TN4ML_B200_H
#define TN4ML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNB_MAX_PHYS 16 /* max physical (embedding) dimension d          */
#define TNB_MAX_OUT 32  /* max class dimension C / SMPO output-leg d_o   */

/* tnb_problem.kind */
enum { TNB_KIND_MPS = 0, TNB_KIND_SMPO = 1 };
/* tnb_problem.dtype */
enum { TNB_F32 = 0, TNB_F64 = 1 };
/* tnb_problem.precision (f32 only): arithmetic of the chain GEMMs.
 * TNB_PREC_TENSOR: fp16x3 on the tensor cores (fp32-class accuracy) for MPS chains with 8 <= chi <= 256 and d in {2, 3, 4, 6, 8}:
 * tcgen05 / TMEM kernels from bond 32 up, register-resident mma.sync kernels at bonds up to 32 with d = 2, 3; a bond that is not a
 * multiple of 16 runs zero-padded to the next one INSIDE the library (params / grads keep the caller's [chi][chi][d] slabs).
 * Any other shape falls back to the SIMT family silently (same results to float32 rounding). */
enum { TNB_PREC_SIMT = 0, TNB_PREC_TENSOR = 1 };
/* tnb_embedding.kind -- tn4ml/embeddings.py:287-351, 354-423, 605-717, 544-602, 426-484, 720-869, 872-928 */
enum {
  TNB_EMB_TRIG = 0,
  TNB_EMB_FOURIER = 1,
  TNB_EMB_POLY = 2,
  TNB_EMB_RBF = 3,
  TNB_EMB_LINCOMP = 4,  /* LinearComplementEmbedding  [x, 1-x] / [1, x, 1-x], unit norm   embeddings.py:426-484 */
  TNB_EMB_LEGENDRE = 5, /* LegendreEmbedding          P_0 .. P_degree                      embeddings.py:720-768 */
  TNB_EMB_LAGUERRE = 6, /* LaguerreEmbedding          exp(-x/2) L_k                        embeddings.py:771-820 */
  TNB_EMB_HERMITE = 7,  /* HermiteEmbedding           exp(-x^2/2) H_k                      embeddings.py:823-869 */
  TNB_EMB_ARRAYS = 8    /* JaxArraysEmbedding         pre-embedded inputs [1?] + x[0..n_in) embeddings.py:872-928 */
};
/* tnb_problem.head -- tn4ml/metrics.py */
enum {
  TNB_HEAD_NLL = 0,                 /* NegLogLikelihood            metrics.py:17-53   */
  TNB_HEAD_SOFTMAX_XENT = 1,        /* OptaxWrapper(softmax_cross_entropy) :423-487   */
  TNB_HEAD_SOFTMAX_XENT_WEIGHTED = 2, /* CrossEntropyWeighted      metrics.py:490-532 */
  TNB_HEAD_XENT_SOFTMAX_ARGMAX = 3, /* CrossEntropySoftmax         metrics.py:291-333 */
  TNB_HEAD_MSE = 4,                 /* MeanSquaredError            metrics.py:336-396 */
  TNB_HEAD_TSN = 5,                 /* TransformedSquaredNorm      metrics.py:56-81   */
  TNB_HEAD_LOGQUAD = 6,             /* LogQuadNorm                 metrics.py:171-187 */
  TNB_HEAD_QUAD = 7                 /* QuadNorm                    metrics.py:190-206 */
};

/* One feature map; replaces Embedding.__call__ + embed()'s normaliser
 * (tn4ml/embeddings.py:1321-1389).  The embedded tensor never reaches HBM. */
typedef struct tnb_embedding {
  int32_t kind;
  int32_t k;            /* trig: number of frequencies, d = 2k          */
  int32_t p;            /* fourier: d = p; lincomp: d = p (2 or 3)      */
  int32_t degree;       /* poly; legendre / laguerre / hermite: d = degree + 1 */
  int32_t include_bias; /* poly; arrays (add_bias)                      */
  int32_t n_centers;    /* rbf: d = n_centers                           */
  int32_t n_in;         /* raw input values per site (Embedding.input_dim; 0 or 1 = scalar features): poly n, arrays */
  int32_t cross;        /* poly: include_cross_terms                    */
  double gamma;         /* rbf                                          */
  double centers[TNB_MAX_PHYS];
} tnb_embedding;

/* Static description of one contraction problem (the jax.ffi attributes). */
typedef struct tnb_problem {
  int32_t kind;      /* TNB_KIND_*                                              */
  int32_t dtype;     /* TNB_F32 / TNB_F64 : type of x, params, grads, outputs   */
  int32_t precision; /* TNB_PREC_*                                              */
  int32_t L;         /* number of sites                                         */
  int32_t chi;       /* bond dimension of the packed layout (max over bonds)    */
  int32_t d;         /* physical dimension (== embedding dimension)             */
  int32_t n_out;     /* MPS: class dimension C (1 = no class leg); SMPO: d_o    */
  int32_t out_site;  /* MPS: 0-based class site (any site when n_out == 1)      */
  int32_t spacing;   /* SMPO: output leg on sites i % spacing == 0 (1 == MPO)   */
  int32_t head;      /* TNB_HEAD_*                                              */
  tnb_embedding emb;
  double class_weights[TNB_MAX_OUT]; /* TNB_HEAD_SOFTMAX_XENT_WEIGHTED          */
} tnb_problem;

/* Packed parameter layout (also the layout of the gradients):
 *   site i occupies chi*chi*d*(n_i) elements, C-contiguous [a][r][s][o] with
 *   a = left bond, r = right bond, s = physical, o = class/output leg (n_i = n_out on
 *   sites that carry one, else 1).  Sites with smaller bonds (boundaries, "noteven"
 *   shapes, tn4ml/models/mps.py:124-151) are zero padded.  Sites are concatenated
 *   in order 0..L-1. */

const char *tnb_version(void);
const char *tnb_last_error(void);

/* Per-device setup (opt-in shared-memory attributes of every kernel): call once per device before capturing a CUDA
 * graph or calling from several threads.  Idempotent.  The XLA-FFI shim calls it from its `instantiate` stage. */
int tnb_init(int32_t device);

/* Number of elements of the packed parameter (and gradient) buffer. */
int64_t tnb_param_count(const tnb_problem *prob);
/* Element offset of site `site` inside the packed buffer, and its n_i. */
int64_t tnb_site_offset(const tnb_problem *prob, int32_t site);
int32_t tnb_site_nout(const tnb_problem *prob, int32_t site);

/* Ragged reference-shaped site tensors <-> the packed layout.  sites[i] is a DEVICE pointer to site i,
 * C-contiguous (chi_l[i], chi_r[i], d, n_i) -- tn4ml/models/mps.py:106-151 shapes with the squeezed boundary bonds
 * given as 1; `packed` holds tnb_param_count elements.  pack zero-fills the padding; unpack reads the same region of
 * a packed gradient back into reference-shaped buffers (value_and_grad returns one array per site, model.py:554-566). */
int tnb_pack_params(const tnb_problem *prob, const void *const *sites, const int32_t *chi_l, const int32_t *chi_r,
                    void *packed, void *stream);
int tnb_unpack_grads(const tnb_problem *prob, const void *packed, void *const *sites, const int32_t *chi_l,
                     const int32_t *chi_r, void *stream);

/* Bytes of caller-provided device scratch for a batch of `batch` samples.
 * need_grad = 0: forward only (evaluate / predict); 1: forward + backward. */
int64_t tnb_workspace_bytes(const tnb_problem *prob, int64_t batch, int32_t need_grad);

/* Forward: embed + contract + head for `batch` samples.
 *   x        [batch, L(, n_in)] raw features (dtype); n_in = emb.n_in values per site when > 1
 *   targets  [batch, n_out]    one-hot / soft labels (dtype) or NULL
 *   params   packed site tensors (dtype)
 *   loss     [batch]           per-sample loss (dtype)            (evaluate(): model.py:1048-1076)
 *   loss_sum [1] float64       sum_b loss_b, ADDED to the existing value, or NULL
 *   out      [batch, n_out]    MPS: mantissa of y = out * 2^out_exp;  SMPO: [batch] mantissa of n
 *   out_exp  [batch] int32     binary exponent carried by the in-chain rescaling
 *   keep_stash != 0 keeps the environments in `workspace` for tnb_backward.
 */
int tnb_forward(const tnb_problem *prob, int64_t batch, const void *x, const void *targets,
                const void *params, void *loss, double *loss_sum, void *out, int32_t *out_exp,
                int32_t keep_stash, void *workspace, int64_t workspace_bytes, void *stream);

/* Backward: gradients of  loss_scale * sum_b loss_b  w.r.t. every site tensor, from the
 * environments left in `workspace` by tnb_forward(..., keep_stash=1, ...) on the same
 * arguments.  grads is the packed layout; accumulate != 0 adds to its contents
 * (micro-batching / gradient accumulation), otherwise it is overwritten.
 * TNB_KIND_SMPO (whose heads take no targets): `targets`, if not NULL, is a vector of per-sample WEIGHTS w[batch] (dtype) and
 * the gradient is that of loss_scale * sum_b w_b loss_b -- what a loss that post-processes the per-sample value needs
 * (SemiSupervisedLoss, tn4ml/metrics.py:209-233: the caller passes dLoss/dLogQuadNorm_b). */
int tnb_backward(const tnb_problem *prob, int64_t batch, const void *x, const void *targets,
                 const void *params, double loss_scale, void *grads, int32_t accumulate,
                 void *workspace, int64_t workspace_bytes, void *stream);

/* The same backward pass in `nparts` site ranges, for overlapping the data-parallel gradient all-reduce with the
 * rest of the weight-gradient work: part 0 runs the adjoint sweep and the gradients of the sites of range 0, part k > 0 the
 * gradients of range k.  After part k returns (stream order) the gradient of every site in
 * [site_begin, site_end) = tnb_part_sites(prob, k, nparts) is final: the caller records an event and all-reduces
 * grads[tnb_site_offset(site_begin) : tnb_site_offset(site_end)] on another stream while part k+1 runs.  Parts must be
 * issued in order 0 .. nparts-1 on the same stream.  tnb_backward == one part. */
int tnb_backward_part(const tnb_problem *prob, int64_t batch, const void *x, const void *targets, const void *params,
                      double loss_scale, void *grads, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                      void *stream, int32_t part, int32_t nparts);
int32_t tnb_part_sites(const tnb_problem *prob, int32_t part, int32_t nparts, int32_t *site_begin, int32_t *site_end);

/* Sweeps / DMRG-like strategy (tn4ml/strategy.py:51-258; the training loop around it: tn4ml/models/model.py:753-773, 829-866):
 * the pair (site, site + 1) is contracted into one tensor T[a, r, s, s'(, k)] = sum_m A_site[a, m, s(, k)] A_site+1[m, r, s'(, k)]
 * and the loss is differentiated with respect to T.  tnb_pair_grad writes that gradient,
 *     dT [chi][chi][d][d][n_k]   (dtype; n_k = n_out when one of the two sites carries the class leg, else 1),
 * of  loss_scale * sum_b loss_b  from the environments left in `workspace` by tnb_forward(keep_stash = 1) followed by
 * tnb_backward_part(part = 0, ...) (the adjoint sweep; its loss_scale is the one that applies) on the same arguments.
 * TNB_KIND_MPS with TNB_PREC_SIMT arithmetic (float32 or float64). */
int tnb_pair_grad(const tnb_problem *prob, int64_t batch, const void *x, const void *targets, int32_t site, void *dT,
                  void *workspace, int64_t workspace_bytes, void *stream);

/* One optax.adam step on the packed buffers (the update side of train_step, tn4ml/models/model.py:591-614, where the
 * reference runs eager optax over an L-leaf pytree):  mu, nu updated in place, params += -lr * mu_hat / (sqrt(nu_hat +
 * eps_root) + eps) with the bias corrections of step `count` (1-based).  grad_scale multiplies the gradient first
 * (global-norm clipping, util.py:131-155, passes its factor here; 1.0 otherwise).  All buffers: n elements of dtype. */
int tnb_adam_update(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr,
                    double b1, double b2, double eps, double eps_root, int64_t count, double grad_scale, void *stream);

/* Same update with one more gradient factor read from DEVICE memory (grad_scale_dev: dtype[1], may be NULL): a factor
 * computed on the device -- 1 / global batch after the all-reduce of the shard sizes, a global-norm clip -- reaches the
 * update without a device-to-host round trip. */
int tnb_adam_update_dev(int32_t dtype, int64_t n, void *params, const void *grads, void *mu, void *nu, double lr,
                        double b1, double b2, double eps, double eps_root, int64_t count, double grad_scale,
                        const void *grad_scale_dev, void *stream);

/* util.gradient_clip (tn4ml/util.py:131-155) on the packed gradient, in place: the gradient of every SITE tensor is scaled by
 * min(1, threshold / (||g_site||_2 + 1e-6)), after multiplying the whole buffer by pre_scale.  scratch: device memory of
 * tnb_gradient_clip_scratch_bytes(prob) bytes; its tail holds the L squared site norms (float64) afterwards. */
int64_t tnb_gradient_clip_scratch_bytes(const tnb_problem *prob);
int tnb_gradient_clip(const tnb_problem *prob, void *grads, double threshold, double pre_scale, void *scratch, void *stream);

/* ---- model-only operations of the training loop ------------------------------------------------------
 * tnb_problem.{kind, dtype, L, chi, d, n_out, out_site, spacing} describe the packed model; embedding / head are ignored. */
/* regulariser kinds -- tn4ml/metrics.py:89-168.  l = log(norm) with norm = sqrt(<A|A>); kind + TNB_REG_TRACE selects the
 * form the reference uses when type(model) is SpacedMatrixProductOperator, where "norm" is Tr(P^dag P) = <A|A> itself
 * (metrics.py:99-103: model.H.apply(model) contracted, smpo.py:406-438). */
enum {
  TNB_REG_LOG_FROB = 0,      /* LogFrobNorm      l                 metrics.py:89-105   */
  TNB_REG_LOG_POW_FROB = 1,  /* LogPowFrobNorm   log(norm^2) = 2 l metrics.py:108-124  */
  TNB_REG_LOG_RELU_FROB = 2, /* LogReLUFrobNorm  max(0, l)         metrics.py:127-146  */
  TNB_REG_QUAD_FROB = 3,     /* QuadFrobNorm     (l - 1)^2         metrics.py:149-168  */
  TNB_REG_TRACE = 4
};
int64_t tnb_norm_workspace_bytes(const tnb_problem *prob);
/* <A|A> of the model by a device transfer-matrix sweep (tn4ml/models/tn.py:66-80, smpo.py:234-248):
 *   log2_norm2 [1] float64 (device): log2 <A|A>  (a logarithm: models of several hundred sites leave the float range)
 * and, if n_regs > 0: reg_value [1] float64 (device, may be NULL) += sum_r reg_coefs[r] * reg_r(model), and grads (packed,
 * dtype, may be NULL) += the gradient of that sum -- what jax.grad produces for the `reg(model)` term of CombinedLoss
 * (metrics.py:535-592). */
int tnb_norm(const tnb_problem *prob, const void *params, double *log2_norm2, int32_t n_regs, const int32_t *reg_kinds,
             const double *reg_coefs, void *grads, double *reg_value, void *workspace, int64_t workspace_bytes, void *stream);
/* params[i] *= 2^(mult * *log2_value) for n elements: Model.normalize() is mult = -1 / (2 L) with log2_value = log2 <A|A>
 * (every site divided by norm^(1/L), tn.py:103-107), or mult = -1/2 on one site (insert=). */
int tnb_scale(int32_t dtype, int64_t n, void *params, const double *log2_value, double mult, void *stream);

/* Embedding.__call__ for n inputs (x [n, n_in]): phi [n, d] (dtype), before embed()'s normaliser. */
int tnb_embed(const tnb_embedding *emb, int32_t dtype, int64_t n, const void *x, void *phi,
              void *stream);

/* Measurement aid (bench.py's roofline leg).  tnb_profile(1) starts bracketing every kernel
 * launch with CUDA events on its stream; tnb_profile(0) stops and drops the records.
 * tnb_profile_read synchronises on the recorded events and returns up to max_records
 * (name[48], milliseconds) pairs in launch order.  Not for use on the hot path. */
int tnb_profile(int32_t enable);
int32_t tnb_profile_read(char *names, float *ms, int32_t max_records);

/* Number of kernel launches issued by this library since load (all threads). */
int64_t tnb_launch_count(void);

/* Deterministic gradient reduction for debugging: on != 0 makes the MPS weight-gradient GEMMs run ONE K-split per
 * output tile (every gradient entry has a single writer accumulating in a fixed order), so two runs on the same inputs give
 * bitwise identical gradients; slower, and on the tensor path the longer TMEM accumulation costs accuracy.  The
 * loss sum is an atomic reduction either way: sum the per-sample losses for a fixed-order value.  The SMPO backward kernel
 * reduces over CTAs with atomics and is not covered.  Also settable with TNB_DETERMINISTIC=1 at load. */
int tnb_set_deterministic(int32_t on);

/* 1 when the XLA-FFI handler symbols of csrc/xla_ffi_shim.cc (forward and backward) were compiled in (the build found
 * xla/ffi/api/ffi.h), else 0.  See INTEGRATION.md. */
int32_t tnb_ffi_available(void);

#ifdef __cplusplus
}
#endif
#endif /* TN4ML_B200_H */
