/*
 * b200nb.h -- C ABI of the B200-native batched negative-binomial GLM engine (libb200nb.so, sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of thelovelab/DESeq2: the three native routines the R
 * package reaches through .Call (registration table /root/reference/src/RcppExports.cpp:84-94):
 *
 *   b200nb_fit_disp       replaces  _DESeq2_fitDisp      src/RcppExports.cpp:16-38, body src/DESeq2.cpp:164-277
 *   b200nb_fit_disp_grid  replaces  _DESeq2_fitDispGrid  src/RcppExports.cpp:64-82, body src/DESeq2.cpp:469-513
 *   b200nb_fit_beta       replaces  _DESeq2_fitBeta      src/RcppExports.cpp:41-61, body src/DESeq2.cpp:283-465
 *
 * Argument order follows the reference signatures (SEXP list -> plain pointers + explicit dimensions,
 * named-list members -> caller-allocated output buffers with the reference's member names).  The R-side
 * .Call shim that re-creates the SEXP interface on top of these symbols is in INTEGRATION.md.
 *
 * Conventions
 *   - matrices are R layout: column-major, genes = rows: y[i + n*j] is gene i, sample j; x is m x p.
 *   - host entry points take HOST pointers, copy to the current CUDA device, run the kernels, copy the
 *     results back; they are synchronous.  Inputs are never modified.
 *   - *_dev entry points take DEVICE pointers in the engine's gene-major layout (one gene per row, sample
 *     axis contiguous, row stride ld elements, ld % 4 == 0, base 16-byte aligned) and enqueue on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) without synchronising.
 *   - every function returns 0 on success, non-zero on failure; b200nb_last_error() then returns a
 *     message (thread-local).  Per-gene numerical failure is in-band exactly as in the reference
 *     (iter == maxit sentinel, NaN coefficients), never an error (src/DESeq2.cpp:357-360,375-378).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - p (design columns) is limited to B200NB_MAX_P.
 */
#ifndef B200NB_H
#define B200NB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NB_MAX_P 32
#define B200NB_Y_INT32 0
#define B200NB_Y_F64 1

/* ---- fitDisp: src/DESeq2.cpp:164.  Outputs = list members at :268-276 (length n each). */
int b200nb_fit_disp(const void* y, int y_type, const double* x, const double* mu_hat, const double* log_alpha,
                    const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                    double kappa_0, double tol, int maxit, int use_prior, const double* weights, int use_weights,
                    double weight_threshold, int use_cr, int n, int m, int p,
                    double* out_log_alpha, int32_t* out_iter, int32_t* out_iter_accept, double* out_last_change,
                    double* out_initial_lp, double* out_initial_dlp, double* out_last_lp, double* out_last_dlp,
                    double* out_last_d2lp);

/* ---- fitDispGrid: src/DESeq2.cpp:469.  Output = list member log_alpha (:512). */
int b200nb_fit_disp_grid(const void* y, int y_type, const double* x, const double* mu_hat, const double* disp_grid,
                         int disp_grid_n, const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq,
                         int use_prior, const double* weights, int use_weights, double weight_threshold, int use_cr,
                         int n, int m, int p, double* out_log_alpha);

/* ---- fitBeta: src/DESeq2.cpp:283.  beta_mat (n x p) is the starting value; outputs = list members at
 * :458-464: out_beta_mat n x p, out_beta_var_mat n x p, out_iter n (double, NumericVector in the reference),
 * out_hat_diagonals n x m, out_contrast_num n, out_contrast_denom n, out_deviance n.
 * nf is the n x m normalisation-factor matrix the reference receives (R/core.R:2221-2228).
 * out_mu (n x m, may be NULL) is an extension: the fitted mean nf*exp(x beta) clamped at minmu, which the
 * reference recomputes in R right after the call (R/fitNbinomGLMs.R:180). */
int b200nb_fit_beta(const void* y, int y_type, const double* x, const double* nf, const double* alpha_hat,
                    const double* contrast, const double* beta_mat, const double* lambda, const double* weights,
                    int use_weights, double tol, int maxit, int use_qr, double minmu, int n, int m, int p,
                    double* out_beta_mat, double* out_beta_var_mat, double* out_iter, double* out_hat_diagonals,
                    double* out_contrast_num, double* out_contrast_denom, double* out_deviance, double* out_mu);

/* ---- device-resident variants (gene-major n x ld matrices, see header comment) */
int b200nb_fit_disp_dev(const void* y, int y_type, const double* x, const double* mu_hat, const double* log_alpha,
                        const double* log_alpha_prior_mean, double log_alpha_prior_sigmasq, double min_log_alpha,
                        double kappa_0, double tol, int maxit, int use_prior, const double* weights,
                        int use_weights, double weight_threshold, int use_cr, int n, int m, int p, long long ld,
                        double* out_log_alpha, int32_t* out_iter, int32_t* out_iter_accept,
                        double* out_last_change, double* out_initial_lp, double* out_initial_dlp,
                        double* out_last_lp, double* out_last_dlp, double* out_last_d2lp, void* stream);

int b200nb_fit_disp_grid_dev(const void* y, int y_type, const double* x, const double* mu_hat,
                             const double* disp_grid, int disp_grid_n, const double* log_alpha_prior_mean,
                             double log_alpha_prior_sigmasq, int use_prior, const double* weights, int use_weights,
                             double weight_threshold, int use_cr, int n, int m, int p, long long ld,
                             double* out_log_alpha, void* stream);

/* nf_is_vector != 0: nf is the length-m size-factor vector instead of the n x ld matrix.
 * beta_mat / out_beta_mat / out_beta_var_mat stay column-major n x p.  out_hat_diagonals / out_mu are
 * gene-major n x ld and may be NULL. */
int b200nb_fit_beta_dev(const void* y, int y_type, const double* x, const double* nf, int nf_is_vector,
                        const double* alpha_hat, const double* contrast, const double* beta_mat,
                        const double* lambda, const double* weights, int use_weights, double tol, int maxit,
                        int use_qr, double minmu, int n, int m, int p, long long ld, double* out_beta_mat,
                        double* out_beta_var_mat, double* out_iter, double* out_hat_diagonals,
                        double* out_contrast_num, double* out_contrast_denom, double* out_deviance, double* out_mu,
                        void* stream);

/* b200nb_nb_loglik_dev: what R recomputes right after fitBeta (R/fitNbinomGLMs.R:180-182): the fitted mean
 * mu = nf * exp(x beta) WITHOUT the minmu clamp (out_mu, gene-major n x ld, may be NULL) and
 * logLike = rowSums([w *] dnbinom(y, mu, size = 1/alpha, log = TRUE)) (nbinomLogLike, R/core.R:2208-2217; out_loglik[n]).
 * nbinomLRT's statistic (R/core.R:1877) and Cook's distances (R/core.R:1457) are built from these, not from the
 * clamped quantities inside the IRLS.  beta_mat: n x p column-major, natural-log scale (fitBeta's out_beta_mat).
 * minmu > 0 evaluates the log-likelihood (not out_mu) at max(mu, minmu): what fitNbinomGLMsOptim stores for the rows
 * it refits (R/fitNbinomGLMs.R:386-398); 0 = no clamp. */
int b200nb_nb_loglik_dev(const void* y, int y_type, const double* x, const double* nf, int nf_is_vector,
                         const double* alpha_hat, const double* beta_mat, const double* weights, int use_weights,
                         double minmu, int n, int m, int p, long long ld, double* out_loglik, double* out_mu,
                         void* stream);

/* b200nb_beta_optim_dev: fitNbinomGLMsOptim (R/fitNbinomGLMs.R:340-407) for the rows whose IRLS did not converge /
 * diverged / gave NA: box-constrained maximisation of logLike + log prior (ridge lambda, natural-log scale here:
 * lambda_log2 / ln(2)^2; box |beta| <= 30 ln 2 = |beta_log2| <= 30).  The objective is strictly concave, so the unique
 * maximiser is what optim(method = "L-BFGS-B") converges to; a projected Newton iteration finds it (maxit iterations
 * at most; out_converged = 1 plays optim's convergence == 0).  Pass ONLY the rows to refit (gathered gene-major rows).
 * beta_start / out_beta_mat: n x p column-major, natural-log scale.  Standard errors, fitted means and the
 * log-likelihood at the optimum come from b200nb_fit_beta_dev(maxit = 0) and b200nb_nb_loglik_dev. */
int b200nb_beta_optim_dev(const void* y, int y_type, const double* x, const double* nf, int nf_is_vector,
                          const double* alpha_hat, const double* lambda, const double* beta_start,
                          const double* weights, int use_weights, int maxit, int n, int m, int p, long long ld,
                          double* out_beta_mat, int32_t* out_converged, int32_t* out_iter, void* stream);

/* layout helpers on device buffers: R column-major n x m <-> gene-major n x ld.  elem_size 4 (int32) or 8. */
int b200nb_to_gene_major_dev(const void* src_colmajor, void* dst, int n, int m, long long ld, int elem_size,
                             void* stream);
int b200nb_to_col_major_dev(const double* src, double* dst_colmajor, int n, int m, long long ld, void* stream);

/* ---- pre-steps of the hot path on device buffers (the R glue around the native calls)
 * b200nb_prep_dev: per gene, from the counts (gene-major n x ld) and size factors:
 *   base_mean, base_var, all_zero      getBaseMeansAndVariances   R/core.R:2138-2157
 *   alpha0 = clamp(min(roughDispEstimate, momentsDispEstimate), min_disp, max_disp)   R/core.R:2422-2448, 716, 727
 *   mu_lin (optional) = linearModelMuNormalized clamped at minmu                      R/core.R:2454-2467, 764
 *   beta0 (optional, n x p column-major) = QR least squares of log(K/s + 0.1) on X    R/fitNbinomGLMs.R:139-145
 * x is m x p column-major; proj is (X'X)^-1 X' as p x m row-major; xim = mean(1/size_factor).
 * b200nb_trend_fit_dev: parametricDispersionFit (R/core.R:2166-2189) over genes with disps > 100*min_disp;
 *   out4 = {asymptDisp, extraPois, status (0 ok, 1 not converged, 2 non-positive, 3 no usable genes), rounds}. */
int b200nb_prep_dev(const void* y, int y_type, const double* x, const double* proj, const double* size_factors,
                    double xim, double min_disp, double max_disp, double minmu, int n, int m, int p, long long ld,
                    double* base_mean, double* base_var, int32_t* all_zero, double* alpha0, double* mu_lin,
                    double* beta0, void* stream);
int b200nb_trend_fit_dev(const double* means, const double* disps, int n, double min_disp, double* out4,
                         void* stream);

/* b200nb_cooks_dev: calculateCooksDistance + recordMaxCooks + robustMethodOfMomentsDisp (R/core.R:2277-2359).
 * cells = distinct design rows: cell_ptr (ncell+1) / cell_samples (m, sample indices grouped by cell), device ints.
 * cooks (gene-major n x ld) may be NULL; max_cooks[n] is NaN when no cell has >= 3 replicates or m <= p. */
int b200nb_cooks_dev(const void* y, int y_type, const double* mu, const double* hat, const double* size_factors,
                     const int32_t* cell_ptr, const int32_t* cell_samples, int ncell, int n, int m, int p,
                     long long ld, double* cooks, double* max_cooks, double* robust_disp, void* stream);

/* b200nb_size_factors_dev: estimateSizeFactorsForMatrix (R/core.R:535-578; locfunc = median, no geoMeans /
 * controlGenes).  poscounts = 0 is type "ratio", 1 is type "poscounts".  y gene-major (n x ld); loggeomeans[n] and
 * size_factors[m] are outputs; scratch_gm (n x ld doubles) and scratch_cm (n x m doubles) are caller-provided device
 * scratch; n_finite (device int) receives the number of genes with a finite log geometric mean -- R stops with
 * "every gene contains at least one zero" when it is 0, the caller must do the same.  A sample with no usable
 * ratio gets NaN (median of an empty set is NA in R). */
int b200nb_size_factors_dev(const void* y, int y_type, int poscounts, int n, int m, long long ld,
                            double* loggeomeans, double* scratch_gm, double* scratch_cm, double* size_factors,
                            int32_t* n_finite, void* stream);

/* ---- housekeeping */
const char* b200nb_last_error(void);
int b200nb_device_count(void);            /* number of visible CUDA devices (0 if none / no driver) */
long long b200nb_kernel_launches(void);   /* kernels this library has launched in this process */
void b200nb_release_workspace(void);      /* frees cached device / pinned buffers of the host entry points */
/* The host entry points keep the gene-major device copies of the large input matrices of recent calls, addressed by
 * CONTENT (dimensions + a 128-bit hash of every byte, recomputed from the caller's buffer on each call), so the count
 * matrix and the fitted means that one DESeq() run passes to fitDisp, fitDisp and fitBeta cross PCIe once.
 * b200nb_cache_clear() forgets them (the memory stays allocated for reuse); B200NB_CACHE_MB=0 disables the cache. */
void b200nb_cache_clear(void);
/* Page-locked memory for RESULT matrices (pooled).  A result buffer obtained here receives its data by one DMA; any
 * other buffer is filled through the library's pinned staging ring by host threads and pays first-touch page faults
 * when it is fresh (~3 ms for a 50 000 x 100 hat-diagonal matrix).  Returns NULL when page-locked memory cannot be had
 * (use ordinary memory then).  In R: allocVector3(REALSXP, n, &allocator) with these two as mem_alloc / mem_free
 * (deseq2_b200/r_shim/deseq2_b200_shim.c). */
void* b200nb_host_alloc(size_t bytes);
void b200nb_host_free(void* p);
/* cumulative counters of the host entry points: out[0] bytes copied host->device, [1] device->host, [2] cache hits,
 * [3] cache misses, [4] bytes served from the cache instead of being uploaded, [5] bytes hashed / scanned on the host.
 * Writes min(n, 6) values, returns 6. */
int b200nb_host_stats(long long* out, int n);
const char* b200nb_version(void);

/* test hook: lgamma / digamma / trigamma of the device math on host arrays x[0..n) (x > 0) */
int b200nb_test_special(const double* x, int n, double* out_lgamma, double* out_digamma, double* out_trigamma);
/* test hook: the host content hash of `count` elements of `elem` (4 | 8) bytes at p, canonical indices first_index...,
 * by every instruction-set variant this CPU can run (scalar, AVX2, AVX-512): out[0..1] receive the scalar value;
 * returns the number of variants compared, or -1 if two of them disagree. */
int b200nb_test_hash(const void* p, long long count, int elem, long long first_index, unsigned long long* out2);

#ifdef __cplusplus
}
#endif
#endif /* B200NB_H */
