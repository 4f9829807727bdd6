/*
 * linna_b200 C ABI -- B200 (sm_100a) implementation of LINNA's emulator-likelihood hot path.
 *
 * Plain C: pointers, sizes, an opaque handle.  No torch / CUDA types in the signatures
 * (`stream` is a cudaStream_t passed as void*; NULL = the legacy default stream).
 * Every entry point returns 0 on success and a negative LINNA_E* code on failure;
 * linna_last_error() gives the message.  There is no CPU fallback: creating a model
 * without a CUDA device fails with LINNA_ENODEV.
 *
 * Threading: a model owns scratch state (activation arenas, relu masks, staging buffers, the training planes) that every
 * launch on it uses, so calls on ONE model must not overlap in time from several host threads; launches issued on different
 * streams are serialised on the device by the library itself (an event chain).  Different models are independent.  Every
 * entry point runs on the model's device and restores the caller's current device before it returns.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference checkout, chto/linna).  INTEGRATION.md shows the reference-side ctypes
 * binding.
 */
#ifndef LINNA_B200_H
#define LINNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LINNA_ABI_VERSION 1

enum {
    LINNA_OK = 0,
    LINNA_EINVAL = -1,  /* bad argument / unsupported shape */
    LINNA_ENODEV = -2,  /* no CUDA device / wrong architecture */
    LINNA_ECUDA = -3,   /* CUDA runtime error (message in linna_last_error) */
    LINNA_ENOMEM = -4,
    LINNA_ESTATE = -5   /* call order: e.g. lnP before linna_model_set_likelihood */
};

/* ---- network description: mirrors linna/nn.py ----------------------------------------- */
enum { LINNA_OP_LINEAR = 0, LINNA_OP_RES = 1 };
enum { LINNA_ACT_NONE = 0, LINNA_ACT_RELU = 1 };

/* One layer.  All weight pointers are HOST float32 in torch nn.Linear layout [out][in].
 *   LINEAR:  y = act(W x + b)                                     (linna/nn.py:121,125-130)
 *   RES:     h = relu(W x + b);  y = relu(alpha*(W2 h + b2) + Ws x)   (linna/nn.py:45-56)
 *            Ws == NULL means the identity skip (in_dim == out_dim, linna/nn.py:28-29). */
typedef struct {
    int32_t kind;
    int32_t in_dim, mid_dim, out_dim;
    int32_t act;
    float alpha;
    const float *w, *b;   /* LINEAR: [out][in],[out]   RES layer1: [mid][in],[mid] */
    const float *w2, *b2; /* RES layer2: [out][mid],[out] */
    const float *ws;      /* RES skip_layer.weight [out][in] or NULL */
} linna_op_desc_t;

/* Emulator = op list + the diagonal input/output transforms that Predictor.predict applies
 * (linna/predictor_gpu.py:461-504):
 *   X_transform_class  (linna/util.py:483-497): xhat = (theta' - x_mean)/x_std, theta'_i = log10(theta_i)
 *                                               for flagged i
 *   Y_transform_class  (linna/util.py:532-542): y = yhat*y_std + y_mean  (exp(.) if ypositive)
 *   Y_invtransform_data(linna/util.py:457-458): m = y*sigma
 * extra_linear_*: ChtoModelv2_linear's "+ 1e-3*linearlayer(xhat)" (linna/nn.py:193), or NULL. */
typedef struct {
    int32_t n_in, n_out, n_ops;
    const linna_op_desc_t *ops;
    const float *x_mean, *x_std;     /* [n_in]  */
    const uint8_t *log10_flag;       /* [n_in] or NULL */
    const float *y_mean, *y_std;     /* [n_out] */
    int32_t ypositive;
    const float *sigma;              /* [n_out] or NULL (=1) */
    const float *extra_linear_w, *extra_linear_b; /* [n_out][n_in], [n_out] or NULL */
    float extra_linear_scale;
} linna_model_desc_t;

/* Likelihood constants of Log_prob (linna/util.py:957-1021):
 *   priors: Transform.__call__ (linna/util.py:323-347): theta_i = u_i*arg2+arg1 (gauss, kind 0) or
 *           Phi(u_i)*(arg2-arg1)+arg1 (flat, kind 1); lnprior = -1/2 sum u^2 (linna/util.py:1160-1165)
 *   data, quadratic form: gaussianlogliklihood (linna/util.py:953-955): -1/2 d C^-1 d^T, d = m - data.
 *     quad_kind LINNA_QUAD_CHOL : `quad` is L, lower-triangular row-major with C^-1 = L L^T
 *                                (chi^2 = |L^T d|^2, the north-star form);
 *     quad_kind LINNA_QUAD_DENSE: `quad` is the dense (symmetrised) C^-1 itself, as the reference uses.
 *   temperature: lnL/T (linna/util.py:1013, linna/main.py:153). */
enum { LINNA_PRIOR_GAUSS = 0, LINNA_PRIOR_FLAT = 1 };
enum { LINNA_QUAD_CHOL = 0, LINNA_QUAD_DENSE = 1 };
typedef struct {
    const int32_t *prior_kind;       /* [n_in] */
    const float *prior_arg1, *prior_arg2;
    const float *data;               /* [n_out] */
    const float *quad;               /* [n_out][n_out] */
    int32_t quad_kind;
    float temperature;
} linna_like_desc_t;

typedef struct linna_model linna_model_t;

int linna_abi_version(void);
const char *linna_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t linna_launch_count(void);

/* Packs the weights (forward and transposed layouts, zero-padded rows) and constants into one
 * device blob on `device`.  Replaces retrieve_model()'s host-side model build (linna/util.py:611-639). */
int linna_model_create(const linna_model_desc_t *desc, int device, linna_model_t **out);
void linna_model_destroy(linna_model_t *m);
int linna_model_set_likelihood(linna_model_t *m, const linna_like_desc_t *like);
/* Re-upload weights after a training update; `ops` must have the shapes given at create time. */
int linna_model_set_weights(linna_model_t *m, const linna_op_desc_t *ops, int32_t n_ops,
                            const float *extra_linear_w, const float *extra_linear_b);

/* What linna_predict writes, row-major [n][n_out]. */
enum { LINNA_OUT_YHAT = 0,  /* network output, normalised space (model(X_transform(X)))            */
       LINNA_OUT_Y = 1,     /* Predictor.predict: y_transform applied   (predictor_gpu.py:500)      */
       LINNA_OUT_M = 2 };   /* y_invtransform_data(predict(.)): data-space model vector (util.py:1012) */

/* Batched Predictor.predict (linna/predictor_gpu.py:461-504).  theta, out: DEVICE pointers. */
int linna_predict(linna_model_t *m, const float *theta, int64_t n, float *out, int32_t out_kind, void *stream);

/* Batched Log_prob.__call__ (linna/util.py:990-1021): u [n][n_in] latent positions -> lnP [n].
 * NaN results are returned as -inf (linna/util.py:1015-1016).  DEVICE pointers. */
int linna_lnp(linna_model_t *m, const float *u, int64_t n, float *lnp, void *stream);

/* lnP and d lnP/du in one launch; replaces Log_prob(nograd=False) + torch.autograd.grad
 * (linna/HMCSampler.py:29-48, linna/util.py:1023-1035 Dlnp).  grad: [n][n_in].  DEVICE pointers. */
int linna_lnp_grad(linna_model_t *m, const float *u, int64_t n, float *lnp, float *grad, void *stream);

/* Vector-Jacobian product of linna_predict -- what torch.autograd computes when the reference calls
 * Predictor.predict(X, no_grad=False) (linna/predictor_gpu.py:495-496) or model(x) with parameters that require grad
 * (linna/predictor_gpu.py:279-283): for a cotangent cot [n][n_out] of the requested output kind,
 *   out    [n][n_out] (or NULL): the forward value itself,
 *   gtheta [n][n_in]  (or NULL): sum_j cot[b][j] d out[b][j] / d theta[b][i],
 *   gparams [n_params] (or NULL): sum_b,j cot[b][j] d out[b][j] / d p, flat state_dict order; needs linna_train_setup
 *           (its row-major activation store) with max_batch >= n.
 * One fused forward + backward-data launch (+ the weight-gradient launch).  DEVICE pointers. */
int linna_predict_vjp(linna_model_t *m, const float *theta, int64_t n, const float *cot, int32_t out_kind, float *out, float *gtheta,
                      float *gparams, void *stream);

/* Host-buffer forms of the three calls above: the arrays are HOST memory (any pageable or pinned
 * buffer); the library stages them through pinned memory, runs the kernel and copies the result
 * back before returning.  These are what a per-call numpy caller (emcee/zeus with vectorize=True)
 * binds, and what bench.py's `e2e` leg times.  Large lnP / lnP+gradient batches are cut into chunks
 * whose copies overlap the kernels; pageable arrays are staged by the calling thread and ONE helper
 * thread that the model starts at its first such call (it sleeps between calls and is joined by
 * linna_model_destroy).  Like every entry point of a model, these calls must not run concurrently on
 * the SAME model (its scratch arena and staging buffers are per-model state); different models are
 * independent. */
int linna_predict_host(linna_model_t *m, const float *theta, int64_t n, float *out, int32_t out_kind);
int linna_lnp_host(linna_model_t *m, const float *u, int64_t n, float *lnp);
int linna_lnp_grad_host(linna_model_t *m, const float *u, int64_t n, float *lnp, float *grad);

/* Introspection used by bench.py / tests. */
int linna_model_info(const linna_model_t *m, int32_t *n_in, int32_t *n_out, int64_t *n_params, int32_t *num_sms);
/* Kernel selection for linna_lnp / linna_lnp_grad / linna_predict: 0 = automatic (the default: tensor-core kernel
 * -- lnP, lnP + gradient and predict programs -- for n >= tc_min_rows, default 256; below that the small-batch cluster kernel, which splits every layer over the
 * CTAs of a thread-block cluster and keeps the activations in distributed shared memory; the FP32 FFMA kernel
 * wherever neither applies), 1 = FP32 FFMA kernel only, 2 = tensor-core (tcgen05, split-fp16) kernel only,
 * 3 = cluster kernel only.  tc_min_rows <= 0 keeps the current threshold. */
int linna_model_set_path(linna_model_t *m, int32_t path, int64_t tc_min_rows);
/* lnP evaluates the last linear layer, the inverse output transform and the Cholesky product as ONE folded
 * affine map (formed in float64 at pack time); 0 switches the folding off (unfolded reference order). */
int linna_model_set_fold(linna_model_t *m, int32_t on);
/* Which kernel served the last launch on this model: 0 none yet, 1 FP32 FFMA kernel, 2 tensor-core kernel,
 * 3 small-batch cluster kernel. */
int linna_model_last_kernel(const linna_model_t *m);
/* Profiling hook (environment LINNA_TC_DEBUG set when the tensor-core context is built): copies the per-CTA
 * cycle counters of the last tensor-core launch into out[max_ctas][128] and returns the number of CTAs
 * (0 when the counters are off).  [0..2] TMA producer: total, waiting for a free stage, waiting for
 * activations; [3..5] MMA issuer (leader CTA of a pair): total, waiting for operands, waiting for the
 * epilogue to drain TMEM; [6..9] / [10..13] first epilogue warp of column group 0 / 1: total, waiting for
 * the MMAs, draining TMEM, chunk epilogues; then per program step (up to 24): [16..] cycles of the MMA
 * issuer, [40..] of epilogue group 0, [64..] of which waiting for accumulators, [88..] of which chunk
 * epilogues. */
int linna_debug_tc_counters(linna_model_t *m, int64_t *out, int32_t max_ctas);
/* Profiling hook (environment LINNA_CLUSTER_DEBUG set): per program step, the cycles thread 0 of the first cluster spent
 * in the k-loop, the k-lane reduction, the epilogue + broadcast and the cluster barrier, summed over the launches since
 * the last read; out[max_values] receives up to 64 x 4 values. */
int linna_debug_cluster_counters(int64_t *out, int32_t max_values);
/* Profiling hook of the tensor-core training kernels (environment LINNA_TG_DEBUG set at linna_train_setup): clock64 stamps
 * of CTA 0 of every layer launch of the last step, out[step][8] = {kernel entry, set-up done, predecessor complete
 * (griddepcontrol.wait), first segment drained, contraction done, epilogue done, tensor memory freed, 0}; returns the
 * number of steps (0 when off). */
int linna_debug_tg_counters(linna_model_t *m, int64_t *out, int32_t max_steps);
/* Force the row-tile height (8, 16 or 32; 0 = automatic) -- test hook for the tiling variants. */
int linna_model_set_tile_rows(linna_model_t *m, int32_t rows);

/* ---- emulator training: Predictor.train inner loop (linna/predictor_gpu.py:268-288) ------------------
 * One optimiser step = fused forward + loss + backward-data pass over the batch, then ONE weight-
 * gradient launch for all layers with the AdamW update fused into its epilogue (or, for data-parallel
 * training, gradients written to `grads` for an NCCL all-reduce followed by linna_train_adamw).
 *   loss: Auxilleryfunc / Loss_fn (linna/util.py:1070-1116) in normalised space:
 *         mean_b chi2(target_b, pred_b) / max(chi2(target_b, data), n_out/2)
 *   data_hat, icov_hat: Auxilleryfunc.__init__ constants (linna/util.py:1060-1069), computed by the
 *         caller in float64 and cast.
 * Parameters, AdamW moments and gradients are flat DEVICE float32 vectors in the reference's
 * state_dict() order, owned by the caller. */
typedef struct {
    const float *data_hat;  /* [n_out] */
    const float *icov_hat;  /* [n_out][n_out] */
    int32_t max_batch;      /* largest batch linna_train_step will see */
} linna_train_desc_t;

int linna_train_setup(linna_model_t *m, const linna_train_desc_t *desc);
int64_t linna_train_num_params(const linna_model_t *m);

/* Per-row quadratic forms of the loss for rows (X physical parameters, Y physical targets), any n:
 * kind 0: chi2(target, pred)   1: chi2(target, data) (NOT clamped)   2: chi2(pred, data)
 * (chisqMnn, chisqMd, chisqnnd of linna/util.py:1077-1085; Val_metric_fn, :1124-1127). */
int linna_train_chisq(linna_model_t *m, const float *X, const float *Y, int64_t n, int32_t kind, float *chi2,
                      void *stream);

/* One optimiser step.  cmd[b] = max(chi2(target_b, data), n_out/2) (targets only, precomputed with
 * linna_train_chisq kind 1).  step counts from 1.  fuse_adam != 0: update params/adam_m/adam_v and the
 * packed weights in place; fuse_adam == 0: write the flat gradient to `grads` only.
 * loss_rows [B] and loss_mean [1] are DEVICE outputs (no host synchronisation). */
int linna_train_step(linna_model_t *m, const float *X, const float *Y, const float *cmd, int64_t B, float *params,
                     float *adam_m, float *adam_v, float *grads, int64_t step, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t fuse_adam, float *loss_rows, float *loss_mean,
                     void *stream);
/* Data-parallel optimiser step with the gradient all-reduce inside it (one process per GPU, NVLink peer memory; the
 * reference intends DistributedDataParallel + AdamW here, linna/predictor_gpu.py:246, :266-267).  Every rank runs
 * linna_train_step(fuse_adam = 0) with `grads` inside a buffer that all ranks allocated symmetrically and exchanged
 * pointers for (torch.distributed._symmetric_memory in this package), then this call: the kernel signals the peers that
 * this rank's gradient is complete, waits for theirs, averages the `world` gradients in rank order straight from the
 * peers' memory and applies AdamW -- no NCCL call, replicas stay bit-identical.  peer_grad_ptrs / signal_pad_ptrs: DEVICE
 * arrays of `world` pointers (rank r's buffer base / signal pad); grad_offset: element offset of this step's gradient in
 * every buffer -- the caller alternates between two halves of the buffer from step to step; avg_offset < 0: every
 * rank reads all the peers' gradients itself (world <= 2), avg_offset >= 0: two-phase form -- element offset of a third
 * region of every buffer in which each rank leaves the average of ITS slice of the vector, fetched by the others after a
 * second barrier (remote reads ~2 x the vector whatever the world size); signal_slot: first 32-bit word of the pad this
 * model may use (32 words).  All ranks must make the same sequence of calls. */
int linna_train_adamw_peer(linna_model_t *m, float *params, float *adam_m, float *adam_v, const void *peer_grad_ptrs,
                           int64_t grad_offset, int64_t avg_offset, const void *signal_pad_ptrs, int32_t signal_slot, int32_t world, int32_t rank,
                           int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay, void *stream);
/* Which kernels run linna_train_step / linna_train_chisq: 0 = automatic (the default: the tensor-core (tcgen05, bf16x3
 * split) kernels whenever the network shape is covered -- LINEAR / RES ops with a skip matrix, LINEAR last layer --, the
 * FP32 FFMA kernels otherwise), 1 = FP32 FFMA kernels only, 2 = tensor-core only (LINNA_EINVAL when unavailable). */
int linna_train_set_path(linna_model_t *m, int32_t path);
/* Kernel family that served the last training call on this model: 0 none yet, 1 FP32 FFMA, 2 tensor core. */
int linna_train_last_kernel(const linna_model_t *m);
/* torch.optim.AdamW update from an (all-reduced) flat gradient; refreshes the packed weights. */
int linna_train_adamw(linna_model_t *m, float *params, float *adam_m, float *adam_v, const float *grads, int64_t step,
                      float lr, float beta1, float beta2, float eps, float weight_decay, void *stream);
/* Overwrite the packed weights from a flat DEVICE parameter vector (after load_state_dict / re-init). */
int linna_train_load_params(linna_model_t *m, const float *params, void *stream);
/* Adopt a flat HOST parameter vector as the model's weights (end of training). */
int linna_train_commit(linna_model_t *m, const float *params_host);

/* Auxilleryfunc.__call__(y_pred, y_target) (linna/util.py:1070-1088) on free-standing DEVICE tensors: per row
 * loss = chi2(target, pred) / chisqMd, chisqMd = max(chi2(target, data), n_out/2), chisqnnd = chi2(pred, data), all in the
 * normalised space of the network output (y_pred [n][n_out] is the network output, y_target physical units), and, when
 * dloss != NULL, d loss_b / d y_pred_b [n][n_out] for autograd.  Constants as for linna_train_setup; icov_hat symmetric.
 * (Predictor.train does not come through here: the training kernels evaluate the loss in their epilogues.) */
int linna_loss_terms(const float *y_pred, const float *y_target, int64_t n, int32_t n_out, const float *data_hat,
                     const float *icov_hat, const float *sigma, const float *y_mean, const float *y_std, int32_t ypositive,
                     float *loss, float *chisq_md, float *chisq_nnd, float *dloss, void *stream);

/* Normalisation statistics of the training set on the device (device pointers): per column c of the row-major matrix
 * Y [n][d], median[c] = LOWER median (torch.median's convention) of v = f(Y[r][c] / sigma[c]) with f = log when take_log
 * (ypositive) else the identity, and, when mad != NULL, mad[c] = lower median of |v - median[c]| -- y_mean and y_std of
 * Y_transform_class as train_NN takes them (linna/util.py:1440-1450, median_absolute_deviation :1308-1313; two CPU sorts
 * of the whole set there).  Radix selection, exact: bit-identical to the reference for take_log == 0.  sigma may be NULL. */
int linna_column_median_mad(const float *Y, int64_t n, int32_t d, const float *sigma, int32_t take_log, float *median,
                            float *mad, void *stream);

/* ---- on-device ensemble-sampler step (emcee's stretch move, linna/sampler.py:493-503, 530) ---------------
 * One half-ensemble update is propose -> linna_lnp(y) -> accept, all on device pointers.
 * propose: for i < ns: partner = second[randint(n_second)], z ~ g(z) on [1/a, a],
 *          y[i,:] = x[partner,:] + z (x[first[i],:] - x[partner,:]);  z[i] is kept for the acceptance.
 * accept:  ln q = (d-1) ln z + lnp_y - lnp[first];  accept iff ln U < ln q and lnp_y is finite; accepted walkers
 *          get x[first[i],:] = y[i,:], lnp[first[i]] = lnp_y[i], naccepted[first[i]] += 1.
 * Random numbers are Philox4x32-10 streams keyed by (seed, i, offset): pass a different offset per call. */
int linna_stretch_propose(const float *x, int32_t d, const int64_t *first, const int64_t *second, int64_t ns,
                          int64_t n_second, float a, uint64_t seed, uint64_t offset, float *y, float *z, void *stream);
int linna_stretch_accept(float *x, float *lnp, float *naccepted, int32_t d, const int64_t *first, int64_t ns, const float *y,
                         const float *lnp_y, const float *z, uint64_t seed, uint64_t offset, void *stream);


/* ---- batched Hamiltonian Monte Carlo step (linna/HMCSampler.py:23-66; every chain has its own Metropolis test) -----------
 * One sample of every chain = linna_hmc_begin, then per leapfrog step linna_lnp_grad(xn) followed by linna_hmc_step
 * (all but the last step), then linna_hmc_end.  x, lnp, grad: current state of the chains [nc][d] / [nc] / [nc][d]; mass [d].
 * begin: p ~ N(0, m) (Philox keyed by (seed, chain, offset)), H0 = sum p^2/2m - lnP, p += eps/2 grad, xn = x + eps p/m.
 * step : p += eps grad_n, xn += eps p/m.
 * end  : p += eps/2 grad_n, H1 = sum p^2/2m - lnP(xn); accept iff U < exp(min(H0 - H1, 0)) and lnP(xn) is finite; accepted
 *        chains get x = xn, lnp = lnp_n, grad = grad_n, naccepted += 1. */
int linna_hmc_begin(const float *x, const float *lnp, const float *grad, const float *mass, int32_t d, int64_t nc, float eps,
                    uint64_t seed, uint64_t offset, float *p, float *xn, float *H0, void *stream);
int linna_hmc_step(float *p, float *xn, const float *grad_n, const float *mass, int32_t d, int64_t nc, float eps, void *stream);
int linna_hmc_end(float *x, float *lnp, float *grad, const float *xn, const float *lnp_n, const float *grad_n, const float *p,
                  const float *mass, const float *H0, int32_t d, int64_t nc, float eps, uint64_t seed, uint64_t offset,
                  float *naccepted, void *stream);

/* ---- convergence statistics (checkmeanstd, linna/sampler.py:370-387) -----------------------------------------------------
 * Mean and population standard deviation per column of rows [r0, r1) of a DEVICE matrix x[rows][d] (float32, or float64
 * when is_double != 0), two-pass reduction in float64; mean_host / std_host are HOST arrays [d]. */
int linna_column_moments(const void *x, int32_t is_double, int64_t r0, int64_t r1, int32_t d, double *mean_host,
                         double *std_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LINNA_B200_H */
