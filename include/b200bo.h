/*
 * b200bo.h -- C ABI of libb200bo.so: the B200 (sm_100a) GP-surrogate + acquisition engine.
 *
 * The reference (wangronin/Bayesian-Optimization, `bayes-optim` 0.3.0) is pure Python and has no FFI;
 * its plug-in boundary for this path is duck-typing on `model.fit / model.predict(X, eval_MSE)` and the
 * acquisition `__call__`.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/bayes_optim/).  The Python mirror that binds these with
 * ctypes is bayesian_optimization_b200/_lib.py; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: every function returns 0 on success or a negative B200BO_E_* code (no exceptions cross
 * the ABI; b200bo_last_error() gives the message).  All matrices are row-major float64.  The caller owns
 * every buffer.  One handle per (process, GPU); calls on one handle must be serialised by the caller.
 * Calls are synchronous: they return after the results are in the caller's buffers.
 */
#ifndef B200BO_H
#define B200BO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200bo_ctx* b200bo_handle;

/* error codes */
#define B200BO_OK 0
#define B200BO_E_ARG (-1)      /* bad argument (shape, id, NULL)                                     */
#define B200BO_E_CUDA (-2)     /* CUDA runtime error; message in b200bo_last_error()                  */
#define B200BO_E_STATE (-3)    /* call order: set_train -> factor -> predict/acq                      */
#define B200BO_E_NODEVICE (-4) /* no CUDA device / wrong architecture (library is sm_100a only)       */

/* correlation ids: surrogate/gaussian_process/gpr.py:201-207 (_correlation_types) + kernel.py */
#define B200BO_CORR_RBF 0      /* "squared_exponential": exp(-sum theta d^2)        kernel.py:289-329 */
#define B200BO_CORR_MATERN12 1 /* matern(nu=0.5): exp(-h), h = sqrt(sum theta d^2)  kernel.py:189-190 */
#define B200BO_CORR_MATERN32 2 /* "matern" (nu=1.5): (1+sqrt3 h) exp(-sqrt3 h)      kernel.py:192-195 */
#define B200BO_CORR_MATERN52 3 /* matern(nu=2.5): (1+sqrt5 h+5h^2/3) exp(-sqrt5 h)  kernel.py:197-200 */
#define B200BO_CORR_ABSEXP 4   /* "absolute_exponential": exp(-sum theta |d|)       kernel.py:247-286 */
#define B200BO_CORR_CUBIC 5    /* "cubic": prod (1 - 3 t^2 + 2 t^3), t=min(1,theta|d|) kernel.py:419-466 */
#define B200BO_CORR_MATERN_NU 7 /* matern(nu = anything else): 2^(1-nu)/Gamma(nu) t^nu K_nu(t), t = sqrt(2 nu) h; theta has
                                   n+1 (or 2) entries, the last one is nu   kernel.py:201-207.  K_nu by Temme's method on
                                   the device (scipy.special.kv upstream); float64 path only, no gradients */
#define B200BO_CORR_GENEXP 6   /* "generalized_exponential": exp(-sum theta |d|^p); theta has n+1 (or 2) entries, the
                                  last one is p   kernel.py:332-374.  float64 path only, no gradients (as upstream) */

/* estimation modes: gpr.py:258-263 */
#define B200BO_MODE_NOISELESS 0   /* R = R0                      par = theta          gpr.py:932-947 */
#define B200BO_MODE_NOISY 1       /* R = (s2 R0 + tau2 I)/(s2+tau2)  par = [theta, s2] gpr.py:963-979 */
#define B200BO_MODE_NOISE_ESTIM 2 /* R = a R0 + (1-a) I          par = [theta, a]     gpr.py:949-961 */

/* trend ids: surrogate/gaussian_process/trend.py */
#define B200BO_TREND_CONSTANT 0  /* constant_trend, p = 1                                   trend.py:69-91  */
#define B200BO_TREND_LINEAR 1    /* linear_trend, f = [1, x], p = D + 1                     trend.py:94-116 */
#define B200BO_TREND_QUADRATIC 2 /* quadratic_trend, f = [1, x, {x_k x_j, j >= k}], p = (D+1)(D+2)/2  :119-142 */
/* p <= 64.  p > 1 runs on the float64 path (B200BO_PREC_FAST falls back to it); beta_or_null then holds p values and
 * B200BO_STATE_FT / _BETA / _G have N*p / p / p*p elements (row-major; G = the R factor of the thin QR of Ft with
 * LAPACK's sign convention, gpr.py:805). */

/* factor status (out_status of b200bo_factor) */
#define B200BO_FIT_OK 0
#define B200BO_FIT_NOT_SPD 1  /* Cholesky pivot <= 0 or NaN: scipy raises LinAlgError -> llf = -inf (gpr.py:946,960,978) */
#define B200BO_FIT_REJECTED 2 /* llf > 0 is rejected as -inf                                   (gpr.py:981-982)     */

/* acquisition ids: acquisition/acquisition_fun.py */
#define B200BO_ACQ_EI 0   /* EI          :150-189                               parameter unused      */
#define B200BO_ACQ_PI 1   /* EpsilonPI / PI :192-235                            parameter = epsilon   */
#define B200BO_ACQ_UCB 2  /* UCB         :107-147                               parameter = alpha     */
#define B200BO_ACQ_MGFI 3 /* MGFI        :238-310                               parameter = t         */

/* memory-location flags for candidate / output pointers */
#define B200BO_HOST 0
#define B200BO_DEVICE 1

/* precision of the M-candidate predict path */
#define B200BO_PREC_FP64 0 /* fp64 DMMA, parity path (default)                                          */
#define B200BO_PREC_FAST 1 /* tcgen05 tensor-core pass + fp64 re-score of the arg-max band: b200bo_acq returns the
                              exact fp64 best_val / best_idx; b200bo_predict returns approximate moments (three
                              split-fp16 products per MAC on an fp32 cross-correlation).  Their stated tolerance is
                              per fit: |d yhat| <= dy_model, |d mse| <= ds_abs_3 + ds_rel_3 sqrt(sum rt^2) of
                              b200bo_get_band_info; measured <= 3.2e-4 and <= 5e-5 sigma2 (1.5e-4 for Matern-1/2) on
                              the test shapes and bench workloads (profiles/r02/fast_predict_errors_*.json, enforced by
                              tests/test_fast_gpu.py).  Calls that ask for all q x M values run on the fp64 path. */

/* state ids for b200bo_get_state */
#define B200BO_STATE_L 0     /* Cholesky factor "C"   (N,N) lower, zeros above   gpr.py:408, :795 */
#define B200BO_STATE_LINV 1  /* L^-1                  (N,N) lower                (replaces solve_triangular gpr.py:494) */
#define B200BO_STATE_GAMMA 2 /* gamma = L^-T rho      (N,)                       gpr.py:788 */
#define B200BO_STATE_YT 3    /* Yt = L^-1 y           (N,)                       gpr.py:799 */
#define B200BO_STATE_FT 4    /* Ft = L^-1 F           (N,p)                      gpr.py:804 */
#define B200BO_STATE_RHO 5   /* rho                   (N,)                       gpr.py:806, :808 */
#define B200BO_STATE_BETA 6  /* beta                  (p,)                       gpr.py:787 */
#define B200BO_STATE_G 7     /* G of the thin QR of Ft (p,p)                     gpr.py:805 */
#define B200BO_STATE_R 8     /* R as assembled, before factorisation (N,N)       gpr.py:772-782, :952, :966-967
                                (only valid if b200bo_set_keep_R(h,1) was called before factor) */

const char* b200bo_last_error(void);
int b200bo_version(void);

/* -- life cycle ---------------------------------------------------------------------------------------
 * replaces GaussianProcess.__init__'s implicit host state (gpr.py:211-277): one engine per model. */
int b200bo_create(int device, b200bo_handle* out);
int b200bo_destroy(b200bo_handle h);
/* run all kernels of this handle on an existing CUDA stream (cudaStream_t as void*), e.g. torch's */
int b200bo_set_stream(b200bo_handle h, void* cuda_stream);
int b200bo_set_precision(b200bo_handle h, int prec);
int b200bo_set_keep_R(b200bo_handle h, int keep);
/* which tensor-core kernel B200BO_PREC_FAST uses: 6 (default) = generation 5 with two CTA pairs sharing one candidate
 * tile: the r chunks are produced once per pair of pairs, the accumulator super-tiles are dealt to the two pairs and the
 * scratch is half the size (N % 256 == 0, N >= 4096 -- below that the tiles are too short for the hand-over between the
 * pairs to pay, B200BO_GEN6_MIN_LD -- and all CTAs co-resident; else 5); 5 = CTA pairs, every r chunk computed once per tile into a
 * scratch by producers that run a tile ahead, all A operands by TMA, per-block accumulator drain (N >= 512, else 4);
 * 4 = CTA pairs + replay of r from an L2-resident scratch, first uses written straight into the A ring;
 * 3 = CTA pairs (tcgen05 cta_group::2) sharing the B operands, r recomputed per accumulator super-tile;
 * 2 = single-CTA kernel with the Gram product on the tensor cores; both need a kernel that is a function of the L2
 * distance, else generation 1 runs; 1 = always the first-generation kernel (A/B comparisons).  Generations 2 and 3 are
 * superseded and only compiled into developer builds (B200BO_DEV_KERNELS=1 python -m bayesian_optimization_b200.build):
 * a product build answers B200BO_E_ARG for them. */
int b200bo_set_fast_kernel(b200bo_handle h, int generation);
/* generation 4 only: the fp16 cross-correlation chunks r[:, k:k+64] of a candidate tile are computed once, kept in a
 * per-SM scratch of at most budget_mb MB in total (default 64: it has to stay in L2 next to the fp16 L^-1) and replayed
 * by TMA for every later accumulator super-tile; chunks that do not fit are recomputed.  budget_mb = 0 recomputes
 * everything (generation 3 behaviour); max_chunks >= 0 additionally caps the stored chunks per tile (test knob). */
int b200bo_set_replay(b200bo_handle h, int budget_mb, int max_chunks);
/* fp16 products per MAC of the first tensor-core pass of b200bo_acq: 1 (default; operands rounded to fp16, ~1e-3 on
 * the variance -- the band it leaves is re-scored in fp64, or the call escalates to 3 when the band is too wide) or
 * 3 (split fp16: ~2^-22 per product; the fp32 cross-correlation then dominates the error).  b200bo_predict always uses 3. */
int b200bo_set_fast_products(b200bo_handle h, int products);

/* Cholesky trailing updates (the rank-64 SYRK of scipy.linalg.cholesky's blocked form, gpr.py:795) on the tcgen05
 * tensor cores: digits = 7 or 8 signed 8-bit digit planes of an exact (error-free) splitting of the float64 panel,
 * multiplied as int8 x int8 -> int32 (kind::i8) and re-assembled in integer arithmetic (7: float64-grade, 8: sub-ulp);
 * digits = 0 (default) keeps the fp64 DMMA kernel.  Trailing matrices below min_rows rows stay on DMMA (<= 0: keep). */
int b200bo_set_chol_tc(b200bo_handle h, int digits, int min_rows);
/* developer / test hook: C (rows x rows, row-major, in/out) -= P P^T on the lower tiles, P (rows x 64) host arrays;
 * digits as above (0 = the DMMA kernel); out_ms = best device time of `reps` runs (CUDA events). */
int b200bo_debug_oz_syrk(b200bo_handle h, const double* P_host, int rows, double* C_host, int digits, int reps, double* out_ms);

/* -- training data: GaussianProcess._check_data (gpr.py:279-310) --------------------------------------
 * X (N,D), y (N,) float64 host pointers.  The pairwise-distance pre-pass l1_cross_distances(X)
 * (gpr.py:48-61) is never materialised: distances are recomputed inside the assembly kernel. */
int b200bo_set_train(b200bo_handle h, const double* X, const double* y, int N, int D);

/* -- fixed-hyper-parameter fit: log_likelihood_concentrated(par, env) (gpr.py:920-991) followed by the
 * attribute copy and compute_beta_gamma of fit() (gpr.py:402-415, :784-788).
 *   theta[n_theta]  n_theta == 1 (isotropic) or D
 *   par_last        sigma2 (NOISY), alpha (NOISE_ESTIM), ignored (NOISELESS)
 *   noise_var       tau^2 (NOISY only)
 *   beta_or_null    NULL: ordinary kriging, beta estimated (mean.beta is None, gpr.py:273-275);
 *                   else p fixed coefficients (simple kriging)
 * outputs: llf (-inf when status != 0), sigma2, noise_var as env[] holds them. */
int b200bo_factor(b200bo_handle h, int corr, const double* theta, int n_theta, int mode, double par_last,
                  double noise_var, int trend, const double* beta_or_null, double* out_llf,
                  double* out_sigma2, double* out_noise_var, int* out_status);

/* analytic gradient of the likelihood at the last factor() point, as the reference computes it
 * (gpr.py:994-1038, incl. its quirks g2/g3).  n_par = n_theta (+1 for NOISY / NOISE_ESTIM). */
int b200bo_llf_grad(b200bo_handle h, double* out_grad, int n_par);

/* likelihood="restricted": log_likelihood_restricted(par, env, eval_grad) (gpr.py:813-918).  All three estimation
 * modes build R = (sigma2 R0 + noise_var I) / (sigma2 + noise_var) (:826-839: noise_var = 0, the nugget, or the last
 * parameter), so the caller passes sigma2 and noise_var explicitly.  exp(llf) > 1 is rejected as -inf (:872-875);
 * the simple-kriging branch keeps upstream's sign of the log-determinant term (:866).  The factorisation state
 * (predict, get_state, gradient) is left at these parameters like after b200bo_factor. */
int b200bo_factor_restricted(b200bo_handle h, int corr, const double* theta, int n_theta, double sigma2, double noise_var,
                             int trend, const double* beta_or_null, double* out_llf, int* out_status);
/* its gradient (gpr.py:876-902): out_grad[i] = slice i of [theta_0 .. theta_{D-1}, sigma2, noise_var], i < n_par --
 * the reference indexes its gradient tensor by the PARAMETER number, also for an isotropic theta. */
int b200bo_llf_grad_restricted(b200bo_handle h, double* out_grad, int n_par);

/* -- GaussianProcess.update(X, y) with the hyper-parameters kept: the bordering ("rank-m") update upstream left as a TODO
 * (gpr.py:419-422 "TODO: implement the rank-one update").  Appends m new training points X_new (m, D) to the model of
 * the last b200bo_factor / b200bo_factor_restricted call at the SAME parameters: new rows of L and of L^-1 from three
 * thin GEMMs against L^-1 (O(m N^2)) instead of a fresh N^3 factorisation, then Yt, Ft, rho, gamma, beta, sigma2 and
 * the likelihood for y_all (N + m,) -- every target, because the callers re-standardise y whenever a point is added
 * (base.py:437).  The state afterwards equals set_train(X_all, y_all) + factor(same parameters) to rounding
 * (tests/test_update_gpu.py: likelihood 1e-10, posterior 1e-9).  Constant trend; any m (blocks of 64). */
int b200bo_append(b200bo_handle h, const double* X_new, int m, const double* y_all, double* out_llf, double* out_sigma2,
                  double* out_noise_var, int* out_status);

int b200bo_get_state(b200bo_handle h, int what, double* out, size_t n_elems);

/* -- GaussianProcess.predict(X, eval_MSE) (gpr.py:424-512) --------------------------------------------
 * Xc (M,D); yhat (M,), mse (M,) (mse may be NULL when eval_mse == 0).  loc: B200BO_HOST / B200BO_DEVICE
 * applies to Xc, yhat and mse alike. */
int b200bo_predict(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat,
                   double* mse);

/* -- batched acquisition with the reference's per-row semantics + arg-max ------------------------------
 * replaces AcquisitionFunction._predict + EI/EpsilonPI/PI/UCB/MGFI.__call__ (acquisition_fun.py:52-64,
 * :127-137, :156-179, :209-218, :268-290) evaluated for q parameter values in ONE predict pass, and the
 * candidate-set arg-max that stands in for argmax_restart (acquisition/optim/__init__.py:55-153).
 *   plugin   already sign-adjusted f* as ImprovementBased.plugin stores it (acquisition_fun.py:96-104)
 *   params   q values of (epsilon | alpha | t); ignored for EI (q criteria still reported)
 *   vals     NULL or (q,M) criterion-major, same location as Xc
 *   best_val (q,), best_idx (q,)  HOST pointers: max value and its LOWEST index (numpy argmax rule) */
int b200bo_acq(b200bo_handle h, const double* Xc, int64_t M, int loc, int acq_id, int minimize,
               double plugin, const double* params, int q, double* vals, double* best_val,
               int64_t* best_idx);

/* Multi-GPU arg-max exchange without the host: after b200bo_acq, write this rank's q (value,
 * global index) pairs into row `rank` of a (world, 2 q) int64 block in DEVICE memory -- [value bits x q | index +
 * index_offset x q, -1 for an empty shard] -- and zeros into the other rows, stream-ordered on the handle's stream.  One
 * integer-sum ncclAllReduce of the block then delivers every rank's pairs bit for bit (each element has a single non-zero
 * contributor); the caller merges them with numpy's rule.  Replaces the joblib fan-out of bayes_opt.py:108-111. */
int b200bo_best_pairs_device(b200bo_handle h, int64_t index_offset, int rank, int world, void* out_dev);

/* acquisition values from given (yhat, mse) -- the elementwise kernel alone (host or device pointers) */
int b200bo_acq_from_moments(b200bo_handle h, const double* yhat, const double* mse, int64_t M, int loc,
                            int acq_id, int minimize, double plugin, const double* params, int q,
                            double* vals, double* best_val, int64_t* best_idx);

/* test hook of the tensor-core path: rt = L^-1 r^T exactly as the tcgen05 pipeline produced it, (M,N) float32,
 * plus the three per-candidate outputs of the fused kernel (each may be NULL).  Host pointers, M <= 65536.
 * What it checks against: solve_triangular(C, r.T) of gpr.py:494. */
int b200bo_debug_fast_rt(b200bo_handle h, const double* Xc, int64_t M, float* out_rt, double* yhat, double* sumsq,
                         double* dotf);

/* -- posterior gradient: GaussianProcess.gradient (gpr.py:537-576, corr_dx :600-661), batched -----------------
 * Row i of y_dx / mse_dx (M, D) is what the reference returns for Xc[i] (it accepts one point per call).  yhat / mse
 * (M,) may be NULL.  Host pointers.  Kernels: squared_exponential, matern (nu = 1.5), absolute_exponential as upstream
 * (incl. the all-zero Jacobian when a Matern point coincides with a training point, gpr.py:628-630, :660-661);
 * Matern-5/2 and -1/2, left unimplemented upstream, from their analytic derivatives; cubic: B200BO_E_ARG. */
int b200bo_gradient(b200bo_handle h, const double* Xc, int64_t M, double* yhat, double* mse, double* y_dx, double* mse_dx);
/* acquisition value and gradient, the return_dx=True path of acquisition_fun.py (UCB :139-146, EI :181-188,
 * EpsilonPI :220-229, MGFI :292-309; early-outs and failures give zeros as upstream): val (M,), dx (M, D). */
int b200bo_acq_grad(b200bo_handle h, const double* Xc, int64_t M, int acq_id, int minimize, double plugin, double param,
                    double* val, double* dx);

/* developer hook: average device time (ms, CUDA events on the handle's stream) of `reps` back-to-back launches of the
 * fused tensor-core kernel alone over M host candidates (no band stage, results discarded); products = 1 or 3. */
int b200bo_debug_fused_time(b200bo_handle h, const double* Xc_host, int64_t M, int products, int reps, double* out_ms);

/* -- exactness of the arg-max on the tensor-core path: what the band stage allowed for, what it observed --------------
 * The band of b200bo_acq (B200BO_PREC_FAST) holds every candidate whose criterion could still be the maximum once the
 * fast pass's error is allowed for; it is re-scored in float64.  Half-widths: yhat +- dy; mse of candidate i +-
 * max(ds_cal, ds_abs + ds_rel sqrt(sum rt_i^2)) + sigma2 (2 |u_i| du + du^2), where (dy, du, ds_abs, ds_rel) come from
 * an A-PRIORI rounding model built at factor() time from ||L^-1|| (row norms, spectral and Frobenius norm), ||gamma||,
 * ||L^-T Ft|| and the kernel's slope (8 modelled standard deviations) and (dy_cal, ds_cal) = 8 x the
 * largest fast-vs-float64 error on a strided 2048-candidate sample of the first call after factor().  Every pass also
 * compares the errors it sees inside the band with what it allowed for and widens x4 + repeats above one half.
 *   out[0] dy model  [1] du model  [2] ds_abs, [3] ds_rel (one product)  [4] ds_abs, [5] ds_rel (three products)
 *   [6] deterministic worst-case bound on the mse error (for the record: it is never the binding one)
 *   [7] max row 2-norm of L^-1  [8] ||L^-1||_2 (power iteration + 25 %)  [9] ||L^-1||_F  [10] max row 1-norm
 *   [11] ||gamma||_2  [12] ||L^-T Ft||_2  [13] max ||x_j||^2 in kernel units  [14] modelled sd of one fp32 r entry
 *   [15] dy_cal, [16] ds_cal (one product)  [17] dy_cal, [18] ds_cal (three)  [19..22] largest calibration errors
 *   (y, mse) x (one, three products)  [23..27] dy, ds, ds_abs, ds_rel, du of the LAST band pass
 *   [28] max |yhat error|, [29] max |mse error| seen inside the last band  [30] their largest ratio to the allowance
 *   [31] widening factor in force */
#define B200BO_N_BAND_INFO 32
int b200bo_get_band_info(b200bo_handle h, double* out, int n);
/* test / bench hook: float64 moments of every stride-th candidate (at most max_samples) of the LAST tensor-core call
 * (b200bo_acq or b200bo_predict with B200BO_PREC_FAST; device-resident input must still be alive) against the moments
 * that call produced:  out[0] candidates checked  [1] max |yhat_fast - yhat|  [2] max |mse_fast - mse| (unclipped)
 * [3] largest (error / allowed half-width)  [4] dy  [5] ds_cal  [6] ds_abs + ds_rel in force  [7] products per MAC.
 * What it checks against: gpr.py:486-505 evaluated in float64 on the device (itself pinned to the reference goldens). */
int b200bo_debug_fast_check(b200bo_handle h, int64_t stride, int64_t max_samples, double* out, int n);

/* -- environment knobs read at b200bo_create (developer / A-B switches; the defaults are the measured best) --------
 *   B200BO_FAST_KERNEL=1..6      generation of the fused tensor-core kernel (default 6), = b200bo_set_fast_kernel
 *   B200BO_FAST_PRODUCTS=1|3     fp16 products per MAC of the first acquisition pass (default 1)
 *   B200BO_REPLAY_MB=n           scratch budget of generation 4 (default 64)
 *   B200BO_CHOL_LOOKAHEAD=0|1|2  Cholesky: single stream | look-ahead, separate kernels | fused panel step where faster
 *   B200BO_GEN6_MIN_LD=n         smallest padded N for which generation 6 replaces generation 5 (default 4096)
 *   B200BO_ASSEMBLE_TMA=0|1      kernel-matrix assembly: first version | TMA-staged, specialised per kernel (default 1)
 *   B200BO_CHOL_TC=0|7|8         Cholesky trailing updates on tcgen05 int8 digit planes (= b200bo_set_chol_tc)
 *   B200BO_CHOL_TC_MIN_ROWS=n    smallest trailing matrix that takes the tensor-core update (default 1024)
 *   B200BO_GRAPHS=0|1            CUDA-graph replay of the factorisation stretches for N <= 2048 (default 1)
 *   B200BO_WAIT_HINT_NS=n        suspend-time hint of the mbarrier waits in the fused kernels
 *   B200BO_BAND_MODEL=0|1        0: band half-widths from the calibration sample only (round-1 behaviour)
 *   B200BO_DEV_CHUNK_TILES=n     device-resident input: fused launches of n candidate tiles per SM (default 8; 0 = one launch)
 *   B200BO_TRACE=path            dump a clock64 timeline of CTA 0 of the fused kernel (Matern-5/2, one product) */

/* -- instrumentation ------------------------------------------------------------------------------------
 * CUDA-event timings (ms) of the last predict/acq call, recorded on the handle's stream:
 *   [0] whole call on device  [1] k* build kernels  [2] L^-1 k* contraction kernels (the dominant kernel)
 *   [3] acquisition + arg-max kernels  [4] number of contraction launches  [5] number of all launches
 *   [6] fp64 re-scored candidates (FAST only)  [7] band passes (FAST only)
 * FAST: [1] = 0 (the k* build is fused), [2] = fused tensor-core kernel, [3] = band selection + exact re-score,
 *       [8] fp16 products per MAC of the pass that produced the result (1 or 3)  [9] 1 if this call escalated 1 -> 3
 *       [10] generation of the fused kernel that ran (b200bo_set_fast_kernel; the fall-backs for small / odd N apply) */
#define B200BO_N_TIMINGS 12
int b200bo_get_timings(b200bo_handle h, double* out, int n);
/* timings (ms) of the last factor(): [0] total [1] assembly [2] cholesky [3] trtri [4] solves  [5] launches */
int b200bo_get_fit_timings(b200bo_handle h, double* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* B200BO_H */
