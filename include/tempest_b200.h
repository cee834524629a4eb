/*
 * tempest_b200.h -- C ABI of libtempest_b200.so (sm_100a kernels for the Persistent
 * Sampling inner loop of minaskar/tempest).
 *
 * Contract: every entry point takes raw DEVICE pointers + sizes + a
 * cudaStream_t (passed as void*), enqueues work on that stream and returns immediately
 * with 0 or a TB_ERR_* / cudaError code.  No entry point allocates or frees caller-visible
 * memory, synchronises the device, or throws.  Scratch memory is supplied by the caller
 * (sizes come from the *_workspace_bytes queries).  All floating point is IEEE fp64;
 * indices are int64.  "ref:" cites the reference function (relative to /root/reference)
 * each entry point replaces; INTEGRATION.md shows the binding a maintainer would add.
 */
#ifndef TEMPEST_B200_H
#define TEMPEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tb_stream_t; /* cudaStream_t */

enum {
  TB_OK = 0,
  TB_ERR_ARG = -1,       /* bad argument (null pointer, negative size, unsupported n_dim ...) */
  TB_ERR_WORKSPACE = -2, /* workspace too small */
  TB_ERR_UNSUPPORTED = -3
};

/* likelihood / prior kernel ids (tempest_b200/registry.py) */
enum { TB_LIKE_EXTERNAL = -1, TB_LIKE_ROSENBROCK = 0, TB_LIKE_GAUSSIAN = 1, TB_LIKE_ISO_MIXTURE = 2, TB_LIKE_TWIN_SHELLS = 3 };
enum { TB_PRIOR_AFFINE = 0 };
enum { TB_SAMPLE_TPCN = 0, TB_SAMPLE_RWM = 1 };
enum { TB_RNG_PHILOX = 0, TB_RNG_TAPE = 1 };

int tb_version(void);
int tb_sm_count(void); /* SM count of the current device (grids are sized from it) */

/* ------------------------------------------------------------------------------------
 * (a) persistent-ensemble reweighting.   ref: tempest/state_manager.py:418-480
 *
 * C_s = LSE_t( l_s*beta_t - logZ_t + log n_t )  (sequential logaddexp in generation order).
 * The reference rebuilds the N_total x T matrix for every probe; C_s does not depend on the
 * probe beta, so it is cached and updated incrementally (bitwise equal to a full rebuild).
 * gen_beta/gen_logz/gen_logn are DEVICE arrays of length T.
 * ---------------------------------------------------------------------------------- */
int tb_mixture_build(const double* logl, double* C, int64_t n_total,
                     const double* gen_beta, const double* gen_logz, const double* gen_logn,
                     int32_t T, tb_stream_t stream);
/* particles [0,n_old) get the one new term (generation T_new-1) folded in; particles
 * [n_old, n_old+n_new) get the full T_new-term sum. */
int tb_mixture_append(const double* logl, double* C, int64_t n_old, int64_t n_new,
                      const double* gen_beta, const double* gen_logz, const double* gen_logn,
                      int32_t T_new, tb_stream_t stream);

/* (b) one ESS probe.   ref: steps/reweight.py:88-118 + tools.py:120-135
 * a_s = beta*l_s - C_s ; out = {m = max a, S1 = sum exp(a-m), S2 = sum exp(a-m)^2,
 * ESS = S1^2/S2, logZ = m + log S1, n_nonfinite}.  workspace: tb_probe_workspace_bytes(). */
size_t tb_probe_workspace_bytes(void);
int tb_probe(const double* logl, const double* C, int64_t n_total, double beta,
             void* workspace, double* out6, tb_stream_t stream);

/* normalised importance weights at beta: w_s = exp(a_s - m) / S1 with (m,S1) read from the
 * device-resident probe result `stats` (out6 of tb_probe at the same beta).
 * ref: steps/reweight.py:299-339 (w / sum w), state_manager.py:473 */
int tb_weights(const double* logl, const double* C, int64_t n_total, double beta,
               const double* stats, double* w, tb_stream_t stream);
/* normalised log-weights logw_s = a_s - (m + log S1)  (posterior(return_logw=True),
 * core.py:197,233-242) */
int tb_log_weights(const double* logl, const double* C, int64_t n_total, double beta,
                   const double* stats, double* logw, tb_stream_t stream);

/* (b) device-side next-beta search: ESS bracket + bisection with the reference's branch
 * logic and constants (steps/reweight.py:123-297, config.py:233-237) in ONE cooperative
 * launch -- no host round trip per probe.
 *   flags bit0: skip the "ESS(beta_prev) <= target -> stay" probe (host already decided it).
 *   flags bit1: speculative passes -- every pass evaluates the current beta and both possible successors (same probe
 *   sequence and result, half the passes; fp64-bound, measured slower on one B200: off by default).
 *   result16[9] = passes over the ensemble (= exchanges on > 1 GPU); result16[6] = probes consumed by the search.
 *   result16[0]=beta, [1]=m, [2]=S1, [3]=S2, [4]=ESS, [5]=logZ(beta), [6]=n_probes,
 *   [7]=beta_low==beta_high flag, [8]=non-finite count; probe_log (may be NULL) receives
 *   (beta,ESS) pairs, at most probe_log_cap pairs.  logl and C must be 16-byte aligned. */
size_t tb_next_beta_workspace_bytes(void);
int tb_next_beta(const double* logl, const double* C, int64_t n_total, double beta_prev,
                 double ess_target, int32_t flags, void* workspace, double* result16,
                 double* probe_log, int32_t probe_log_cap, tb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (c) resampling.   ref: steps/resample.py:52-99, tools.py:178-228, numpy legacy choice
 *
 * tb_cdf_exact reproduces numpy's strictly sequential fp64 cumsum BIT FOR BIT with a
 * parallel algorithm (binade-segmented integer scan); tb_cdf_sequential is the
 * one-thread literal form kept for cross-checks.  p must be non-negative.
 * ---------------------------------------------------------------------------------- */
size_t tb_cdf_workspace_bytes(int64_t n);
int tb_cdf_exact(const double* p, int64_t n, double* cdf, void* workspace, tb_stream_t stream);
/* Two single-GPU implementations of tb_cdf_exact, both bit for bit numpy's cumsum: the multi-kernel pipeline of
 * tb_cdf_exact_x (default) and a chained scan with decoupled look-back that carries the exact running sum (a binade-guess
 * pass + ONE streaming kernel; tb_cdf_set_chain(1) or the environment variable TB_CDF_CHAIN=1).  On the weight vector
 * of a C4 run (3.8e7 weights, ~55 multi-binade jumps of the running sum) the chain is the slower one: every jump stalls
 * all tiles in flight (profiles/r02_cdf_paths.txt).  tb_cdf_chain_diag_ptr: device uint64[24] of the last chained call on
 * `workspace` {tiles needing > 1 round, rounds in them, look-back retries, re-published aggregates, serial-regime
 * elements, late prefix publications, look-back rounds, ns summed over tiles ticket -> aggregate, aggregate -> start
 * value, start value -> done, failed look-back attempts by reason [10..17]}. */
void tb_cdf_set_chain(int32_t on);
uint64_t* tb_cdf_chain_diag_ptr(void* workspace, int64_t n);
/* The same over a SHARDED weight vector.  The global order is generation-major, rank-minor (the order of the
 * single-GPU ensemble): this rank stores one contiguous segment per generation, seg_begin[n_gen + 1] (device)
 * are their local start positions.  x: rank / world / call number seq (>= 1, +1 per call on every rank) and
 * peer[r] = rank r's table memory (tb_cdf_x_table_bytes(ntg_cap) bytes of zero-initialised symmetric /
 * peer-mapped memory) as addressed from this device; NULL or world 1 = one GPU, tables inside `workspace`.
 * n_global: the global element count (sizes the launches; the same value on every rank).
 * ntg_cap >= tb_cdf_tile_cap(n_global, n_gen * world), the same value on every rank and for every call that
 * shares the table memory; workspace: tb_cdf_x_workspace_bytes(ntg_cap).
 * Tile tables and the elements of the few tiles that cross a binade travel by peer stores between the
 * stages; every rank launches the same call.  cdf receives this rank's part of numpy's cumsum of the global
 * vector, bit for bit.  tb_cdf_status_ptr: device int32[16] {tiles, segments, runs, hard tiles, error code (0 ok,
 * 3 peer timeout, 4/5 table capacity, 6 refuted hypothesis, 7 chained kernel: a predecessor tile never published), ...};
 * tb_cdf_total_ptr: device double cdf[-1]. */
typedef struct tb_cdf_x {
  int32_t rank, world;
  uint64_t seq;
  void* peer[8];
} tb_cdf_x;
int64_t tb_cdf_tile_cap(int64_t n_global, int32_t n_segments);
size_t tb_cdf_x_workspace_bytes(int64_t ntg_cap);
size_t tb_cdf_x_table_bytes(int64_t ntg_cap);
int32_t* tb_cdf_status_ptr(void* workspace);
double* tb_cdf_total_ptr(void* workspace, int64_t ntg_cap);
int tb_cdf_exact_x(const double* p, int64_t n_local, const int64_t* seg_begin, int32_t n_gen, int64_t n_global,
                   int64_t ntg_cap, double* cdf, void* workspace, const tb_cdf_x* x, tb_stream_t stream);
/* searches in the cdf of the last tb_cdf_exact_x call on `workspace` (same p / cdf / x arguments): two levels,
 * exact tile-end values (identical on every rank) then the owner's local cdf.  systematic 0: m replicated
 * uniforms `draws`, idx_k = numpy searchsorted(cdf / cdf[-1], u_k, 'right'); systematic 1: positions
 * (u0 + k) / m, tools.py:217-226 (*overflow set when a position lies beyond the last weight).  idx_k = LOCAL
 * index of the ancestor, or -1 when it is stored on another rank. */
int tb_cdf_search_x(const double* p, int64_t n_local, double* cdf, void* workspace, int64_t ntg_cap,
                    const tb_cdf_x* x, const double* draws, int64_t m, int32_t systematic, double u0,
                    int64_t* idx, int32_t* overflow, tb_stream_t stream);
int tb_cdf_sequential(const double* p, int64_t n, double* cdf, tb_stream_t stream);
/* multinomial: idx_k = #{ j : cdf_j / cdf_{n-1} <= U_k }   (searchsorted side='right') */
int tb_search_right(const double* cdf, int64_t n, const double* draws, int64_t m,
                    int64_t* idx, tb_stream_t stream);
/* the same indices for many draws: a guide table over 2^bits uniform grid points brackets every draw so
 * the search needs ~2-3 probes; guide: tb_search_guide_bytes(bits) bytes of scratch */
size_t tb_search_guide_bytes(int32_t bits);
int tb_search_right_guided(const double* cdf, int64_t n, const double* draws, int64_t m, void* guide,
                           int32_t bits, int64_t* idx, tb_stream_t stream);
/* systematic: pos_k = (u0 + k)/m ; idx_k = first j with cdf_j >= pos_k ; *overflow is set
 * to 1 if some pos_k exceeds cdf_{n-1} (the reference raises IndexError there). */
int tb_systematic(const double* cdf, int64_t n, double u0, int64_t m, int64_t* idx,
                  int32_t* overflow, tb_stream_t stream);
/* gather rows of the persistent ensemble into the active set (row-major [m,d]) */
int tb_gather_rows(const double* hist_u, const double* hist_logl, int32_t d,
                   const int64_t* idx, int64_t m, double* act_u, double* act_logl,
                   tb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (e) moments / volume variation / trim / Student-t mode statistics
 * ---------------------------------------------------------------------------------- */
/* mean_j = sum_s w_s u_sj ; cov = sum_s w_s (u_s-mean)(u_s-mean)^T  (w normalised).
 * ref: tools.py:94-99.  workspace: tb_moments_workspace_bytes(d). */
size_t tb_moments_workspace_bytes(int32_t d);
int tb_weighted_moments(const double* u, const double* w, int64_t n, int32_t d,
                        void* workspace, double* mean, double* cov, tb_stream_t stream);
/* cv = 0.5*sqrt( sum_s w_s^2 clip(d2_s - d, +-1e6)^2 ), d2_s = (u_s-mean)^T cov_inv (u_s-mean).
 * ref: tools.py:111-115 */
int tb_mahalanobis_cv(const double* u, const double* w, int64_t n, int32_t d,
                      const double* mean, const double* cov_inv, void* workspace,
                      double* cv_out2 /* {cv, raw sum} */, tb_stream_t stream);
/* batched d x d Cholesky + inverse with the reference's regularise-on-failure rule
 * (modes.py:105-119): on a non-positive pivot add max(1e-6, 1e-6*|tr|) to the diagonal (the
 * regularised matrix is written back to `a`) and retry.  info[k] = 0 ok, 1 regularised,
 * 2 failed even after regularisation.  norms3[k] = {|A|_F, |A^-1|_F, tr A}: |A|_F*|A^-1|_F
 * bounds cond_2 from above and screens the rank test of tools.py:101. */
int tb_chol_inv(double* a, int32_t d, int32_t batch, double* chol, double* inv,
                int32_t* info, double* norms3, tb_stream_t stream);
/* Sigma = cov_ddof1*(M-1)/M + diag(var_ddof0)/M from the count-weighted scatter matrix
 * (student.py:63) */
int tb_student_sigma(const double* scatter, int32_t d, double m_total, double* sigma,
                     tb_stream_t stream);
/* med[c] = (pair[2c] + pair[2c+1]) / 2  (np.median of an even count, student.py:62) */
int tb_median_pairs(const double* pair, int32_t d, double* med, tb_stream_t stream);
/* cov += factor * trace(cov) * I   (tools.py:101-104 regularisation, factor = 1e-6) */
int tb_add_trace_reg(double* cov, int32_t d, double factor, tb_stream_t stream);

/* trim_weights (tools.py:10-55).  Step kernels; the <=1000-threshold scan is driven by the
 * host step object (tempest_b200/steps.py) as a binary search over the monotone criterion.
 *   tb_normalize_inplace : w /= sum(w) ; stats3 = {sum before, sum w^2 after, max after}
 *   tb_binade_hist       : per-binade (exponent) count / sum w / sum w^2, 2048 bins
 *   tb_select_ranks      : exact order statistics by 6 MSD radix passes of 11 bits
 *   tb_masked_sums       : {count, sum w, sum w^2} over w >= thr
 *   tb_compact_ge        : ordered stream compaction of {s : w_s >= thr} -> idx, w/denom
 */
size_t tb_reduce_workspace_bytes(void);
int tb_normalize_inplace(double* w, int64_t n, void* workspace, double* stats3, tb_stream_t stream);
int tb_binade_hist(const double* w, int64_t n, uint64_t* count2048, double* s1_2048,
                   double* s2_2048, tb_stream_t stream);
/* the same three sums over the 2048 linear sub-bins (top 11 mantissa bits) of one binade */
int tb_subbin_hist(const double* w, int64_t n, int32_t binade, uint64_t* count2048, double* s1_2048,
                   double* s2_2048, tb_stream_t stream);
int tb_masked_sums(const double* w, int64_t n, double thr, void* workspace, double* out3,
                   tb_stream_t stream);
size_t tb_compact_workspace_bytes(int64_t n);
int tb_compact_ge(const double* w, int64_t n, double thr, double denom, void* workspace,
                  int64_t* idx_out, double* w_out, int64_t* n_out, tb_stream_t stream);

/* exact k-th order statistics (0-based ranks, ascending) of non-negative doubles with
 * integer multiplicities.  keys are read as  base[ (rows ? rows[j] : j) * stride + col ],
 * j < n, col < ncols; mult may be NULL (all ones).  For each column the `nranks` ranks
 * (ascending) are resolved simultaneously.  workspace: tb_select_workspace_bytes(ncols,nranks).
 * Used for np.percentile order statistics (tools.py:46) and np.median (student.py:62). */
size_t tb_select_workspace_bytes(int32_t ncols, int32_t nranks);
int tb_select_ranks(const double* base, const int64_t* rows, int64_t stride, int64_t n,
                    int32_t ncols, const int32_t* mult, const int64_t* ranks, int32_t nranks,
                    void* workspace, double* out /*[ncols*nranks]*/, tb_stream_t stream);

/* the adjacent pair of order statistics (rank_lo, rank_lo+1) per column -- what np.percentile's
 * linear interpolation and np.median of an even count need -- in 4 radix passes of 16 bits plus one
 * min pass; `same` != 0 returns the rank_lo value twice (rank_lo is the last element).
 * out[2c], out[2c+1].  workspace: tb_select_pair_workspace_bytes(ncols). */
size_t tb_select_pair_workspace_bytes(int32_t ncols);
int tb_select_pair(const double* base, const int64_t* rows, int64_t stride, int64_t n, int32_t ncols,
                   const int32_t* mult, int64_t rank_lo, int32_t same, void* workspace, double* out,
                   tb_stream_t stream);

/* same pair for columns whose values lie in [0,1] (unit-cube coordinates): fixed-point bucket
 * histogram -> compaction of the bucket(s) holding the two ranks -> exact in-block select.
 * *overflow = 1 when a bucket held more than 65536 candidates (caller falls back to tb_select_pair). */
size_t tb_unit_median_workspace_bytes(int32_t d);
/* stage-by-stage form for sharded ensembles (same workspace): 0 local bucket histogram, 1 pick the bucket(s)
 * of the global rank from the (all-reduced) histogram, 2 compact the local candidates, 3 exact select over the
 * (merged) candidate lists.  tb_bucket_offsets: byte offsets {histogram u32[d][65536], selector records (24 B
 * per column: b1, b2, rank_local i64, count u32, overflow i32), candidates f64[d][65536], multiplicities
 * u32[d][65536]}.  unit_map != 0: keys in [0,1) bucketed by floor(v*65536); else by bits in [lo_value, hi_value]. */
int tb_bucket_offsets(int32_t d, int64_t* out4);
int tb_bucket_stage(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d,
                    int64_t rank_lo, int32_t same, double lo_value, double hi_value, int32_t unit_map,
                    int32_t stage, void* workspace, double* out, int32_t* overflow, tb_stream_t stream);
/* Sharded runs, between stage 2 and stage 3: concatenate the all-gathered candidate lists of the ranks
 * (gsel int32[world][d][2] = {count, overflow} of every rank's stage 2, gval / gmul [world][d][capx] = the first capx
 * candidate slots of every rank) into this rank's candidate arrays in rank order and set the global counts; a count
 * above capx, an overflow on any rank or more than 65 536 candidates in total raise the overflow flag stage 3 reports.
 * No count visits the host. */
int tb_bucket_merge(const int32_t* gsel, const double* gval, const uint32_t* gmul, int32_t world, int32_t d, int32_t capx,
                    void* workspace, tb_stream_t stream);
int tb_unit_median_pair(const double* u, const int64_t* rows, const int32_t* mult, int64_t n, int32_t d,
                        int64_t rank_lo, void* workspace, double* out, int32_t* overflow,
                        tb_stream_t stream);

/* same bucket -> compact -> exact-select scheme for ONE column of non-negative doubles known to lie in
 * [lo_value, hi_value] (weights above a binade edge): buckets are affine in the bit pattern.  The
 * histogram is kept in the workspace, so follow-up calls on the same data pass build_hist = 0.
 * workspace: tb_unit_median_workspace_bytes(1). */
int tb_bucket_select_pair(const double* v, int64_t n, int64_t rank_lo, int32_t same, double lo_value,
                          double hi_value, int32_t build_hist, void* workspace, double* out2,
                          int32_t* overflow, tb_stream_t stream);

/* multiplicity of each trimmed row among the 4n training draws: counts[idx[k]] += 1 */
int tb_count_indices(const int64_t* idx, int64_t m, int32_t* counts, int64_t n, tb_stream_t stream);
/* count-weighted mean and scatter of u[rows[j]] (multiplicity mult[j]):
 *   mean = sum c_j u_j / M ; scatter = sum c_j (u_j-mean)(u_j-mean)^T ; M = sum c_j
 * ref: student.py:63 (np.cov / np.var of the 4n-row resampled set, modes.py:272-275) */
int tb_counted_moments(const double* u, const int64_t* rows, const int32_t* mult, int64_t n,
                       int32_t d, double inv_total, void* workspace, double* mean,
                       double* scatter, tb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (d) mutation.   ref: steps/mutate.py:76-200, mcmc.py:104-323
 * ---------------------------------------------------------------------------------- */
/* In-kernel exchange across the GPUs of one node (sharded runs): peer[r] is rank r's exchange buffer
 * (tb_xgpu_buffer_bytes() bytes of zero-initialised symmetric / peer-mapped memory) as addressed from
 * this device; seq is the next unused exchange sequence number (>= 1, consecutive across calls). */
typedef struct tb_xgpu {
  int32_t rank, world;
  uint64_t seq;
  double* peer[8];
} tb_xgpu;
size_t tb_xgpu_buffer_bytes(void);

typedef struct tb_tape {
  /* TB_RNG_TAPE: variates recorded from the oracle's stream */
  const double* gamma;    /* [steps][n]  standard-gamma variates (tpCN) */
  const double* acc_u;    /* [steps][n]  accept uniforms */
  const double* z;        /* flat normals; attempt a of walker k at step t starts at z_off[t*n+k] + a*d */
  const int64_t* z_off;   /* [steps][n] */
  const int32_t* z_cnt;   /* [steps][n] attempts recorded */
  int32_t steps;          /* steps available on the tape */
} tb_tape;

typedef struct tb_mcmc_params {
  int32_t n_dim, n_modes, sampler, rng_mode;
  int32_t like_id, prior_id;
  int32_t n_steps, n_max;          /* per-dimension base / max step counts (config.py:80-84) */
  int32_t defer_update, reserved;  /* defer_update 1: sharded run, leave per-step totals for an all-reduce +
                                    * tb_mcmc_update.  reserved != 0: mode statistics / prior parameters are unchanged
                                    * since the previous tb_mcmc_steps call of this mutation (skip constant reloads) */
  double beta;
  uint64_t seed, iteration;         /* Philox key material */
  int64_t slot_offset;              /* global slot id of local walker 0 (multi-GPU) */
  int64_t n_global;                 /* global walker count (== n when not sharded) */
  const double* like_params;        /* device */
  const double* prior_params;       /* device */
  const double* mode_mean;          /* [K][d] device */
  const double* mode_chol;          /* [K][d][d] lower, row-major */
  const double* mode_inv;           /* [K][d][d] */
  const double* mode_dof;           /* [K] */
  const uint8_t* bc_kind;           /* [d] 0 strict, 1 periodic, 2 reflective; NULL = all strict */
  const tb_xgpu* xgpu;              /* HOST pointer or NULL: fuse the per-step all-reduce into the step kernel
                                     * (exchange number of step s of this call is xgpu->seq + s) */
} tb_mcmc_params;

/* warm-up draw at beta = 0: u = uniforms (tape `prior_u` [n][d] or Philox), x = prior(u),
 * logl = L(x).  ref: steps/mutate.py:100-120 */
int tb_prior_draw(int64_t n, const tb_mcmc_params* p, const double* prior_u_tape,
                  double* u, double* x, double* logl, tb_stream_t stream);
/* x = prior(u), optionally logl = L(x), for rows of u (posterior() / history export) */
int tb_transform(const double* u, int64_t n, const tb_mcmc_params* p, double* x, double* logl,
                 tb_stream_t stream);

/* control block layout (doubles), device resident, written by the step kernel:
 *  [0] steps done  [1] done flag  [2] sum accept (last step)  [3] mean alpha (last step)
 *  [4] error flag (1 = tape exhausted, 2 = non-finite)  [5] total proposals drawn
 *  [8 .. 8+K)       sigma_c
 *  [8+K .. 8+2K)    walkers per mode (filled by tb_mcmc_begin)
 *  [8+2K .. 8+3K)   sum alpha per mode of the last step */
size_t tb_mcmc_workspace_bytes(int64_t n, int32_t n_modes);
size_t tb_mcmc_ctrl_doubles(int32_t n_modes);
/* reset the control block (sigma_c init, mcmc.py:222-223/298-299), count walkers per mode and
 * cache q_k = (u_k-mu)^T Sigma^-1 (u_k-mu) of the starting state (tpCN).  qcur holds 2n doubles:
 * q_k in [0,n) and the Student-t term -(D+nu)/2 log(1+q_k/nu) of the current state in [n,2n). */
int tb_mcmc_begin(int64_t n, const tb_mcmc_params* p, const int32_t* assign, const double* u,
                  double* qcur, void* workspace, double* ctrl, tb_stream_t stream);
/* run up to `count` Metropolis steps in ONE persistent cooperative launch (registry likelihoods): the kernel
 * stops by itself when the stop rule of mcmc.py:104-135,192-194 fires and leaves steps / done / sigma /
 * acceptance in ctrl.  The workspace is private to the launch.  With p->xgpu the per-step totals of all
 * ranks are exchanged over peer memory inside the kernel (every rank must launch with the same count);
 * with defer_update it runs ONE step and leaves this rank's totals for an all-reduce + tb_mcmc_update. */
int tb_mcmc_steps(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign,
                  double* u, double* logl, double* qcur, void* workspace, double* ctrl,
                  int32_t count, tb_stream_t stream);

/* ---- entry points used when particles are sharded over GPUs ---------------------- */
/* multinomial search inside the GLOBAL cdf (generation-major, rank-minor): this rank holds one
 * segment per generation, seg_begin[n_seg+1] local positions; global value of local element j in
 * segment s is cdf[j] + seg_shift[s]; seg_start[s] is the global cdf value just before segment s.
 * idx = local ancestor index or -1 when the draw belongs to another rank. */
int tb_search_right_sharded(const double* cdf, int64_t n, const int64_t* seg_begin, const double* seg_shift,
                            const double* seg_start, int32_t n_seg, double total, const double* draws,
                            int64_t m, int64_t* idx, tb_stream_t stream);
/* the same with a guide table (tb_search_guide_bytes(bits) bytes): identical indices, ~3 probes per draw */
int tb_search_right_sharded_guided(const double* cdf, int64_t n, const int64_t* seg_begin, const double* seg_shift,
                                   const double* seg_start, int32_t n_seg, double total, const double* draws,
                                   int64_t m, void* guide, int32_t bits, int64_t* idx, tb_stream_t stream);
/* w /= denom */
int tb_scale_inplace(double* w, int64_t n, double denom, tb_stream_t stream);
/* the same with the divisor read from device memory (sharded runs: the all-reduced sum never visits the host) */
int tb_scale_inplace_dev(double* w, int64_t n, const double* denom, tb_stream_t stream);
/* one stage of tb_select_ranks: 0 init, 1 local histogram of `level`, 2 pick from the (all-reduced)
 * histogram, 3 write results; the histogram lives at workspace + tb_select_hist_offset() */
int tb_select_stage(const double* base, const int64_t* rows, int64_t stride, int64_t n, int32_t ncols,
                    const int32_t* mult, const int64_t* ranks, int32_t nranks, void* workspace, double* out,
                    int32_t stage, int32_t level, tb_stream_t stream);
size_t tb_select_hist_offset(int32_t ncols, int32_t nranks);
/* moments with explicit control: exactly one of w / mult is non-NULL; do_mean writes
 * mean = inv_norm * sum w x (a partial sum when sharded), do_cov the scatter about `mean` */
int tb_moments_partial(const double* u, const int64_t* rows, const double* w, const int32_t* mult, int64_t n,
                       int32_t d, double inv_norm, int32_t do_mean, int32_t do_cov, void* workspace,
                       double* mean, double* cov, tb_stream_t stream);
/* tb_next_beta with the per-probe merge of the ranks' (m,S1,S2) triples fused into the kernel over
 * peer memory: every rank launches it with the same arguments; consumes result16[6] exchanges */
int tb_next_beta_x(const double* logl, const double* C, int64_t n_total, double beta_prev,
                   double ess_target, int32_t flags, void* workspace, double* result16,
                   double* probe_log, int32_t probe_log_cap, const tb_xgpu* xgpu, tb_stream_t stream);
/* One Metropolis step split around a CALLER-evaluated likelihood (arbitrary prior_transform /
 * log_likelihood callables, core.py:317-358): tb_mcmc_propose writes the in-cube proposals u_prop[n][d]
 * and meta[n] (proposals drawn, or -error); the caller computes logl_prop = L(prior(u_prop)); then
 * tb_mcmc_accept does the Student-t ratio, accept/reject, sigma adaptation and the stop rule exactly
 * like tb_mcmc_steps.  Set like_id = TB_LIKE_EXTERNAL in tb_mcmc_begin's params. */
int tb_mcmc_propose(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign,
                    const double* u, const double* logl, const double* qcur, void* workspace, double* ctrl,
                    double* u_prop, int32_t* meta, tb_stream_t stream);
int tb_mcmc_accept(int64_t n, const tb_mcmc_params* p, const tb_tape* tape, const int32_t* assign,
                   double* u, double* logl, double* qcur, void* workspace, double* ctrl,
                   const double* u_prop, const double* logl_prop, const int32_t* meta, tb_stream_t stream);
/* apply sigma adaptation + stop rule from all-reduced per-step totals (defer_update = 1) */
int tb_mcmc_update(const tb_mcmc_params* p, double* ctrl, tb_stream_t stream);

/* ------------------------------------------------------------------------------------
 * (e') hierarchical Gaussian-mixture clustering.   ref: tempest/cluster.py
 *
 * One weighted mixture fit lives in a device PARAMETER BLOCK of tb_gmm_block_doubles(d,k) doubles:
 * a 16-double header {lower bound, done, completed passes (= n_iter once done), last bound, -, -, -,
 * bound-only result, ...} followed by weights[k], means[k,d], covariances[k,d,d], the inverse lower
 * Cholesky factors of cov + reg I, log normalisers, ok flags and M-step scratch; tb_gmm_offsets
 * returns {weights, means, covs, linv, lognorm, ok, mass, total} offsets in doubles.
 * x is row-major [.,d]; rows (nullable) lists the member rows of the cluster being fitted, in the
 * reference's order; sw are the members' normalised sample weights (position-indexed).
 * ---------------------------------------------------------------------------------- */
int64_t tb_gmm_block_doubles(int32_t d, int32_t k);
int tb_gmm_offsets(int32_t d, int32_t k, int64_t* out8);
size_t tb_gmm_workspace_bytes(void);
/* weighted k-means++ (cluster.py:139-158): p_i = min_c |x_i - c|^2 sw_i over the first k centres;
 * tb_kpp_pick: j = searchsorted(run, frac*run[n-1]) (left), centre <- x[rows[j]] */
int tb_kpp_prob(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d,
                const double* centres, int32_t k, double* p, tb_stream_t stream);
int tb_kpp_pick(const double* run, int64_t n, double frac, const double* x, const int64_t* rows,
                int32_t d, double* centre, int64_t* picked, tb_stream_t stream);
/* initial parameters from the centres (cluster.py:160-168): responsibilities exp(-|x-c_k|^2/2),
 * normalised, then one M-step.  wr: k*n doubles of scratch (r*sw columns). */
int tb_gmm_init(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k,
                const double* centres, double* block, double* wr, void* workspace,
                void* mom_workspace, double reg, tb_stream_t stream);
/* enqueue `passes` EM passes (cluster.py:103-121, 172-250, 264-283); passes after convergence
 * (new bound - bound < tol) or after max_iter return immediately; block[1] != 0 when finished,
 * block[2] = n_iter */
int tb_gmm_em(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k,
              double* block, double* wr, void* workspace, void* mom_workspace, double reg, double tol,
              int32_t max_iter, int32_t passes, tb_stream_t stream);
/* block[7] = sum_i sw_i log(sum_k w_k N_k(x_i) + 1e-10); sw NULL = 1/n (the BIC term, cluster.py:329-338) */
int tb_gmm_bound(const double* x, const int64_t* rows, const double* sw, int64_t n, int32_t d, int32_t k,
                 double* block, void* workspace, tb_stream_t stream);
/* factor cov + reg I of every component of a block whose weights/means/covs were written by the
 * caller; a failed factorisation falls back to `fallback` * I (cluster.py:185-188, 667-670) */
int tb_gmm_prepare(double* block, int32_t d, int32_t k, double reg, double fallback, tb_stream_t stream);
/* labels_i = argmax_k log(w_k + 1e-10) + log N(x_i; mu_k, Sigma_k + reg I) (cluster.py:285-308,
 * 633-696); lo/hi (nullable pair): min-max normalise x on the fly, (x - lo)/((hi - lo) + 1e-10) */
int tb_gmm_predict(const double* x, const int64_t* rows, int64_t n, int32_t d, int32_t k,
                   const double* block, const double* lo, const double* hi, int32_t skip_bad,
                   int32_t* labels, tb_stream_t stream);
/* column minima / maxima of x[rows] (cluster.py:437-438) and the normalised gather (:382) */
size_t tb_col_minmax_workspace_bytes(int32_t d);
int tb_col_minmax(const double* x, const int64_t* rows, int64_t n, int32_t d, void* workspace,
                  double* lo, double* hi, tb_stream_t stream);
int tb_gather_normalised(const double* x, const int64_t* rows, int64_t n, int32_t d, const double* lo,
                         const double* hi, double* out, tb_stream_t stream);
/* out_i = src[idx_i] */
int tb_take(const double* src, const int64_t* idx, int64_t n, double* out, tb_stream_t stream);
/* order-preserving split of a member list by a 0 / non-0 label (cluster.py:495-496):
 * out_zero[0 : n - *n_one], out_one[0 : *n_one]; members NULL = positions */
size_t tb_split_workspace_bytes(int64_t n);
int tb_split_by_label(const int32_t* labels, const int64_t* members, int64_t n, void* workspace,
                      int64_t* out_zero, int64_t* out_one, int64_t* n_one, tb_stream_t stream);

/* testing hook: route every n_dim through the generic (runtime-d) step kernel instead of the
 * compile-time-d fast path (tape mode must give identical decisions on both) */
int tb_set_mcmc_generic(int32_t on);
/* n_dim without a compile-time instantiation (17..128): use the warp-cooperative runtime-d step kernel
 * (tb_mcmc_wide.cu) instead of the per-thread one */
int tb_set_mcmc_wide(int32_t on);
/* testing hook: 0 routes single-mode runs through the multi-mode instantiation of the fused step kernel
 * (mode statistics in shared memory) instead of the constant-memory single-mode one (default 1) */
int tb_set_mcmc_kone(int32_t on);
/* testing hook: the three variates one Metropolis step of walker slot_offset+i consumes in Philox mode
 * (mcmc.py:236,243,169), exactly as the step kernels generate them: gamma[n] standard-gamma variate of
 * `shape`, z[n][d] the normals of redraw `attempt`, acc_u[n] the accept uniform (each nullable).
 * family 0: fused kernels (tb_mcmc_steps), family 1: split step (tb_mcmc_propose / tb_mcmc_accept) */
int tb_debug_variates(uint64_t seed, uint64_t iteration, int64_t slot_offset, int64_t n, int32_t step, int32_t d,
                      double shape, int32_t attempt, int32_t family, double* gamma, double* z, double* acc_u,
                      tb_stream_t stream);

/* out[i] = uniform [0,1) number i+offset of stream (seed, iteration, purpose): the draws the
 * host-driven resampling / training steps consume in Philox mode (purpose 4 / 5) */
int tb_philox_uniform(uint64_t seed, uint64_t iteration, uint32_t purpose, int64_t offset,
                      int64_t n, double* out, tb_stream_t stream);

/* volume_variation's rank test / regularisation / final value on the device (tools.py:101-115), so that the cv
 * diagnostic needs no host round trip: after tb_chol_inv of a copy of cov, tb_vv_regularise sets flags[0] =
 * rank-deficient and then adds 1e-6 trace(cov) I to cov, and copies cov to work for the second tb_chol_inv;
 * tb_vv_finish writes result[0] = 0.5 sqrt(raw[0]) or 1e10 when that inverse failed as well. */
int tb_vv_regularise(double* cov, double* work, int32_t d, const int32_t* info, const double* norms, int32_t* flags,
                     tb_stream_t stream);
int tb_vv_finish(const double* raw, const int32_t* info2, const double* norms2, const int32_t* flags, double* result,
                 tb_stream_t stream);

/* Small-message collectives over peer memory for sharded runs (csrc/tb_xcoll.cu): two kernels on the caller's
 * stream, no host synchronisation.  peer[r] = rank r's staging buffer (tb_xcoll_buffer_bytes(cap_bytes) bytes of
 * zero-initialised symmetric memory, cap_bytes a multiple of 16) as addressed from this device; seq = call number
 * (>= 1, +1 per call on every rank); ticket = a zeroed device uint32 private to this rank.  dtype 0 f64, 1 i64,
 * 2 i32.  All-reduce folds the ranks' payloads in rank order: bitwise identical results on every rank.
 * *err (device int32, nullable) receives 3 if a peer did not answer within the spin budget. */
typedef struct tb_xcoll {
  int32_t rank, world;
  uint64_t seq;
  uint64_t cap_bytes;
  void* ticket;
  void* peer[8];
} tb_xcoll;
size_t tb_xcoll_buffer_bytes(size_t cap_bytes);
int tb_xcoll_allreduce_sum(void* data, int64_t n, int32_t dtype, const tb_xcoll* x, int32_t* err, tb_stream_t stream);
int tb_xcoll_allgather(const void* src, int64_t nbytes, void* out, const tb_xcoll* x, int32_t* err, tb_stream_t stream);
/* Resampled rows of a sharded run (resample.py:86-96): idx[n_glob] holds, for every global draw k, the LOCAL index
 * of its ancestor or -1 (tb_cdf_search_x).  Row k = (u[idx k], logl[idx k]) is stored over NVLink straight into the
 * active-set buffer of the rank that owns walker slot k (k / per), then the stream is held until every rank's rows
 * have arrived.  x->peer[r]: rank r's row buffer (tb_xrows_buffer_bytes(per, d) bytes of zeroed symmetric memory);
 * the rows of call number seq land at double offset tb_xrows_offset(per, d, seq & 1): u[per][d] then logl[per].
 * (x->cap_bytes is ignored.) */
size_t tb_xrows_buffer_bytes(int64_t per, int32_t d);
int64_t tb_xrows_offset(int64_t per, int32_t d, int32_t parity);
int tb_xrows_scatter(const double* hu, const double* hl, int32_t d, const int64_t* idx, int64_t n_glob, int64_t per,
                     const tb_xcoll* x, int32_t* err, tb_stream_t stream);

/* measurement aid (bench.py): one launch of 8 * n_SM CTAs x 256 threads, each running eight independent
 * register-resident fp64 FMA chains for `iters` iterations = tb_fp64_peak_flops(iters) flop; timed by the
 * caller with CUDA events it gives the fp64 FMA peak the mutation kernel's roofline is quoted against */
/* Benchmark of the in-kernel synchronisation point of the persistent kernels (tb_xgpu.cuh grid_xreduce): `count`
 * back-to-back grid-wide (+ cross-GPU when xgpu->world > 1) reductions of a 4-double row in ONE cooperative launch of
 * `grid` CTAs; out2 = {ns per synchronisation point, checksum}.  Consumes `count` exchange sequence numbers. */
size_t tb_xgpu_bench_workspace_bytes(int32_t grid);
int tb_xgpu_bench(const tb_xgpu* xgpu, int32_t grid, int32_t count, void* workspace, double* out2, tb_stream_t stream);
int64_t tb_fp64_peak_flops(int32_t iters);
int tb_fp64_peak_run(int32_t iters, double* out, tb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TEMPEST_B200_H */
