/*
 * gpcc_b200.h -- C ABI of the B200-native (sm_100a) replacement for the GPCC.jl hot path.
 *
 * The reference (HITS-AIN/GPCC.jl v0.1.35) is pure Julia and has no FFI; this header is the contract a
 * thin Julia `ccall` shim (julia/GPCC_B200.jl) binds instead of the reference's Julia
 * functions.  Each entry cites the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain C, all floating point is FP64, all arrays are caller-owned contiguous HOST memory;
 *     per-band data are concatenated in band order (Julia `reduce(vcat, ...)`); matrices are
 *     column-major (Julia layout).  The library never keeps a caller pointer after returning.
 *   - every function returns 0 on success, <0 for an invalid argument, >0 for a CUDA/NCCL failure;
 *     gpcc_last_error() gives the text.  Numerical failure (matrix not positive definite) is reported
 *     per element in `info` (LAPACK convention: k>0 = leading minor k not PD) with loglik = -Inf; it is
 *     never a return code.  No exception, exit or signal crosses the boundary.
 *   - calls are synchronous; a context is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device gpcc_ctx_create fails.
 *   - lifetime: a problem and a fit state may only be USED while their context (and, for a fit state, its problem)
 *     is alive; they may be DESTROYED in any order, also after the context (garbage-collected hosts run
 *     finalizers in unspecified order): destroy functions only touch memory the object itself owns.
 */
#ifndef GPCC_B200_H
#define GPCC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPCC_MAX_BANDS 8

/* src/util.jl:15-52 -- the four scalar kernels; the Julia shim maps function identity to this enum. */
enum { GPCC_KERNEL_OU = 0, GPCC_KERNEL_RBF = 1, GPCC_KERNEL_MATERN32 = 2, GPCC_KERNEL_MATERN52 = 3 };

/* Parameter transforms of src/gpccfixdelay_marginaliseb.jl:112-114 (MiscUtil.makepositive /
 * transformbetween, un-vendored): id 0 = softplus + 1e-8 floor, rhomin+(rhomax-rhomin)*logistic.   */
enum { GPCC_TRANSFORM_SOFTPLUS_LOGISTIC = 0 };
enum { GPCC_OPT_LBFGS = 0, GPCC_OPT_NELDERMEAD = 1 };

typedef struct gpcc_ctx gpcc_ctx;
typedef struct gpcc_problem gpcc_problem;
typedef struct gpcc_fit_state gpcc_fit_state;

/* Keyword arguments / constants of gpcc (src/gpccfixdelay_marginaliseb.jl:46, :69, :112, :205). */
typedef struct gpcc_fit_options {
    int    max_iter;        /* `iterations`: cap on optimiser iterations per candidate            */
    double rhomin;          /* `rhomin` (default 0.1)                                              */
    double rhomax;          /* `rhomax`                                                            */
    double alpha_floor;     /* 1e-8  (:112)                                                        */
    double gtol;            /* stop when max|d(-logL)/dtheta| <= gtol   (L-BFGS; default 1e-7)      */
    double ftol;            /* stop when the relative decrease of -logL <= ftol twice (1e-13)      */
    int    history;         /* L-BFGS memory (default 8, max 16)                                   */
    int    transform_id;    /* GPCC_TRANSFORM_*                                                    */
    int    theta0_per_candidate; /* 0: theta0 is [P][L+1] shared by all candidates (the reference's
                                    seed=1 behaviour, :62); 1: theta0 is [M][P][L+1]               */
    int    optimizer;       /* GPCC_OPT_LBFGS (default) or GPCC_OPT_NELDERMEAD: the reference's own optimiser
                               (:205-211, g_tol = nm_gtol), on the device, for basin-for-basin comparisons;
                               fused small-N path only (N <= 199)                                  */
    double nm_gtol;         /* Optim.Options(g_tol = 1e-6) (:205): simplex spread at which Nelder-Mead stops */
} gpcc_fit_options;

/* Per-stage device timings of the most recent call, measured with CUDA events on the library's own
 * streams (only when enabled with gpcc_ctx_set_profiling).                                          */
typedef struct gpcc_stats {
    double ms_total;            /* whole call, host wall clock                                      */
    double ms_eval_kernels;     /* sum of likelihood(+gradient) kernel time, CUDA events            */
    double ms_assembly;         /* large-N path: covariance assembly kernels                        */
    double ms_factor;           /* large-N path: pivot + panel + trailing-update kernels            */
    double ms_gradreduce;       /* large-N path: gradient reduction kernel                          */
    long long n_eval_launches;  /* kernels launched by this library in the call                     */
    long long n_evals;          /* (candidate, theta) evaluations                                   */
    long long n_evals_grad;     /* ... of which with gradient                                       */
    int    path;                /* 0 = fused register-resident small-N kernel, 1 = tiled large-N    */
    int    n_devices;
    long long n_shared_prefix;  /* tiled path, logL only: evaluations whose leading block (all bands but the
                                   last) was factorised by another evaluation of their wave (structure reuse) */
    long long n_tau_cache;      /* tiled path, fixed-theta sweep over >= 3 bands: evaluations that took the band-1 steps of
                                   their last band (a function of the last delay alone) from the last-band cache            */
    long long assembly_bytes;   /* tiled path: HBM bytes the assembly kernels of the busiest device moved (tiles written;
                                   tiles imported from the last-band cache count read + written); pairs with ms_assembly      */
} gpcc_stats;

int         gpcc_version(void);   /* 101.  100 -> 101: gpcc_stats grew by n_tau_cache and assembly_bytes (appended at the end);
                                      a caller built against 100 must be rebuilt before it calls gpcc_ctx_get_stats                */
const char* gpcc_last_error(void);

/* CUDA contexts, streams, workspaces on `ndev` devices (dev_ids NULL = 0..ndev-1).  With ndev>1 the
 * candidate grid is sharded over the devices (README.md:185-206 `pmap`) and the per-candidate
 * log-likelihoods are combined with one NCCL allgather.                                             */
int gpcc_ctx_create(int ndev, const int* dev_ids, gpcc_ctx** out);
int gpcc_ctx_destroy(gpcc_ctx* ctx);
/* One process per GPU (torchrun; one Julia `Distributed` worker per device, README.md:185-189): rank 0 draws an id,
 * the host program hands its 128 bytes to every rank, and each rank attaches its ONE-device context.  From then on
 * gpcc_fit_batch fits only the candidates m with m mod world == rank, and gpcc_grid_posterior (called by EVERY rank with
 * the same arguments) all-gathers the per-candidate records over NCCL and returns the full outputs on every rank.   */
int gpcc_comm_unique_id(char* out_id128);
int gpcc_ctx_comm_init_rank(gpcc_ctx* ctx, int world, int rank, const char* id128);
int gpcc_ctx_set_profiling(gpcc_ctx* ctx, int enabled);
int gpcc_ctx_get_stats(const gpcc_ctx* ctx, gpcc_stats* out);
int gpcc_ctx_device_count(const gpcc_ctx* ctx);

/* Data of one gpcc call (src/gpccfixdelay_marginaliseb.jl:85-98): L bands, n_per_band[l] points each,
 * t / y / sigma concatenated (N = sum n).  mub / Sigmab (length L: prior mean and the *diagonal* of the
 * inflated prior covariance, :92-94) may be NULL, in which case the library computes mean(y_l) and
 * 100*var(y_l) (unbiased) itself.  The data are uploaded to every device of the context.            */
int gpcc_problem_create(gpcc_ctx* ctx, int L, const int* n_per_band, const double* t, const double* y,
                        const double* sigma, int kernel_id, const double* mub, const double* Sigmab,
                        gpcc_problem** out);
int gpcc_problem_destroy(gpcc_problem* p);
int gpcc_problem_get_prior(const gpcc_problem* p, double* mub, double* Sigmab);

/* objective(alpha, rho) of src/gpccfixdelay_marginaliseb.jl:133-141 for M (delay, hyper-parameter)
 * pairs in constrained space: K~ = delayedCovariance (src/delayedCovariance.jl:1-38) + Sobs + B,
 * logpdf(MvNormal(bbar, K~), Y).  want_grad adds d logL / d(alpha_1..alpha_L, rho) (not in the
 * reference, which is derivative free).  delays, alpha: [M][L]; rho, out_ll: [M];
 * out_grad: [M][L+1] (may be NULL when want_grad == 0); out_info: [M] (may be NULL).                */
int gpcc_loglik_batch(gpcc_problem* p, int M, const double* delays, const double* alpha, const double* rho,
                      int want_grad, double* out_ll, double* out_grad, int* out_info);

/* Same objective in the optimiser's unconstrained parameters theta = [L+1] (unpack, :116-126), with
 * the gradient chained through the transforms.                                                      */
int gpcc_loglik_theta_batch(gpcc_problem* p, int M, const double* delays, const double* theta,
                            const gpcc_fit_options* opt, int want_grad, double* out_ll, double* out_grad,
                            int* out_info);

/* getsolution (src/gpccfixdelay_marginaliseb.jl:203-215) for M candidate delay vectors at once:
 * stage 1 evaluates the P (`initialrandom`) start points theta0 and keeps the best (:207-209); stage 2
 * maximises logL from there (:211; host-driven batched L-BFGS with the analytic gradient instead of
 * Nelder-Mead).  Outputs (any may be NULL): out_ll[M] = -result.minimum (:351), out_theta[M][L+1],
 * out_alpha[M][L], out_rho[M], out_iters[M], out_nfev[M], out_info[M] (0 converged, 1 iteration cap,
 * 2 line search stalled, <0 no start point was positive definite).                                  */
int gpcc_fit_batch(gpcc_problem* p, int M, const double* delays, int P, const double* theta0,
                   const gpcc_fit_options* opt, double* out_ll, double* out_theta, double* out_alpha,
                   double* out_rho, int* out_iters, int* out_nfev, int* out_info);

/* The README's grid driver (README.md:170-178, 185-210, 285) followed by getprobabilities
 * (src/getprobabilities.jl:1-20): fit every candidate (sharded over the context's devices, candidate
 * m on device m mod ndev), allgather the log-likelihoods, normalise with log-sum-exp under `logprior`
 * (NULL = flat).  out_post[M] sums to 1; the other outputs are as in gpcc_fit_batch.                */
int gpcc_grid_posterior(gpcc_problem* p, int M, const double* delays, const double* logprior, int P,
                        const double* theta0, const gpcc_fit_options* opt, double* out_ll, double* out_post,
                        double* out_theta, double* out_alpha, double* out_rho, int* out_nfev, int* out_info);

/* src/getprobabilities.jl:10-20 on the device: exp(ll + logprior - logsumexp(ll + logprior));
 * logprior NULL = the 1-argument method (:1-6).                                                     */
int gpcc_getprobabilities(gpcc_ctx* ctx, int M, const double* loglik, const double* logprior, double* out_post);

/* What the reference's `pred` closures capture after the fit (src/gpccfixdelay_marginaliseb.jl:235-252: alpha, rho,
 * KSobsB, postb): ONE Cholesky factorisation of K + Sobs at the fitted (delays, alpha, rho), kept on the device
 * together with postb, and reused by every later prediction (the reference re-factorises on each call, :275, :283).
 * Returns -6 when K + Sobs is not positive definite.                                                   */
int gpcc_fit_state_create(gpcc_problem* p, const double* delays, const double* alpha, double rho, gpcc_fit_state** out);
int gpcc_fit_state_destroy(gpcc_fit_state* s);
/* postb (:248-252): out_mu[L], out_Sigma[L*L] (symmetrised, :252).                                     */
int gpcc_fit_state_postb(const gpcc_fit_state* s, double* out_mu, double* out_Sigma);
/* predictTest (:259-307) and the test log-likelihood (:311-343) from the cached factor; arguments as in
 * gpcc_predict / gpcc_predict_loglik below.                                                            */
int gpcc_fit_state_predict(gpcc_fit_state* s, const int* ntest_per_band, const double* ttest, double* out_mu,
                           double* out_sd, double* out_Sigma);
int gpcc_fit_state_predict_loglik(gpcc_fit_state* s, const int* ntest_per_band, const double* ttest,
                                  const double* ytest, const double* sigmatest, double* out_ll, int* out_info);
/* Draws from the process at the state's hyper-parameters, on the device: out_f[s][N] = Lc z_s with Lc Lc' = K + Sobs and
 * z_s ~ N(0, I) from a counter-based generator (Philox-4x32-10 + Box-Muller, reproducible per seed).  This is the sampling
 * step of the reference's simulator (src/simulatedata.jl:128-145: Y ~ MvNormal(0, C), C = delayedCovariance(OU, ...)) for
 * large synthetic benchmarks; the light curves around it (scaling, offsets, noise: :151-159) are O(N) host
 * arithmetic in both shims.  out_z (may be NULL) receives the deviates [nsamples][N].                                  */
int gpcc_fit_state_sample(gpcc_fit_state* s, unsigned long long seed, int nsamples, double* out_f, double* out_z);
/* diagnostics: number of N^3 factorisations this state has run (1 for its whole life).               */
long long gpcc_fit_state_factorisations(const gpcc_fit_state* s);

/* Stateless forms of the three calls above.  They keep the fit state of the most recent (delays, alpha, rho) on
 * the problem, so postb + any number of pred calls at the fitted hyper-parameters share one factorisation.
 * Posterior of the shifts b (src/gpccfixdelay_marginaliseb.jl:248-252): out_mu[L], out_Sigma[L*L].  */
int gpcc_postb(gpcc_problem* p, const double* delays, const double* alpha, double rho, double* out_mu,
               double* out_Sigma);

/* predictTest (src/gpccfixdelay_marginaliseb.jl:259-307).  ttest holds ntest_per_band[l] times per band,
 * concatenated (NT = sum).  out_mu[NT]; out_sd[NT] = sqrt(max(diag(Sigma_pred), 1e-6)) (:303) if not
 * NULL; out_Sigma[NT*NT] = full predictive covariance incl. the 1e-8 jitter (:279) if not NULL.      */
int gpcc_predict(gpcc_problem* p, const double* delays, const double* alpha, double rho,
                 const int* ntest_per_band, const double* ttest, double* out_mu, double* out_sd,
                 double* out_Sigma);

/* pred(ttest, ytest, sigmatest) (src/gpccfixdelay_marginaliseb.jl:311-343): test log-likelihood under
 * N(mu_pred, Sigma_pred + diag(sigmatest^2)); out_info = 0, or k>0 if that matrix is not PD (the
 * reference then repairs it with MiscUtil.nearestposdef on the host, which the caller keeps doing).  */
int gpcc_predict_loglik(gpcc_problem* p, const double* delays, const double* alpha, double rho,
                        const int* ntest_per_band, const double* ttest, const double* ytest,
                        const double* sigmatest, double* out_ll, int* out_info);

/* Fill `opt` with the reference defaults (seed-independent part of :46, :112, :205).               */
int gpcc_fit_options_default(gpcc_fit_options* opt);

#ifdef __cplusplus
}
#endif
#endif /* GPCC_B200_H */
