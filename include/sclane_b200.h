/*
 * sclane_b200.h -- C ABI of libsclane_b200.so: the B200-native engine for scLANE's
 * per-gene trajectory-DE fitting path (testDynamic() -> marge2() -> NB fits -> LRT).
 *
 * Plain C: pointers + sizes, caller-owned buffers, no C++/torch types.  Every entry point
 * names the reference interface it replaces (file:line relative to the scLANE checkout,
 * v0.99.2).  The reference's own native boundary is `.Call("_scLANE_eigenMap*", ...)`
 * (R/RcppExports.R:4-14, src/RcppExports.cpp:14-63); the R-side stubs a maintainer would add
 * for the entry points below are shown in INTEGRATION.md.
 *
 * Conventions
 *   - return 0 = call completed; per-gene failures are DATA (status bits / NaN statistics),
 *     mirroring try(..., silent = TRUE) + .errorhandling = "pass" (testDynamic.R:176,188-199,387-420).
 *   - return != 0 = global failure (bad arguments, CUDA error, out of memory, no CUDA device);
 *     a message is written to errbuf (NUL terminated).  The library never exits, never throws
 *     across the boundary and has NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with SCLANE_ERR_CUDA.
 *   - NOT re-entrant: the per-device work buffers, streams and the last-call profile are process-wide.  Every compute
 *     entry point and sclane_release() take one internal lock, so concurrent callers are serialised (R is single
 *     threaded, testDynamic.R never calls the path concurrently); do not call the library from a signal handler.
 *   - NA_real_ inputs are NaN (tested with isnan); NaN is emitted for NA outputs.
 *   - all matrices of the eigenMap* drop-ins are column-major doubles, as R/Eigen::Map hand them over.
 */
#ifndef SCLANE_B200_H
#define SCLANE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCLANE_ABI_VERSION 4
#define SCLANE_PMAX 7          /* max columns of a model (intercept + hinges): M <= 7 (marge2.R:756) */
#define SCLANE_TMAX 32         /* max candidate knots per lineage (n.knot.max, marge2.R:66; default 25) */
#define SCLANE_MAX_GPUS 16

/* error codes */
#define SCLANE_OK 0
#define SCLANE_ERR_ARG 1
#define SCLANE_ERR_CUDA 2
#define SCLANE_ERR_NOMEM 3
#define SCLANE_ERR_UNSUPPORTED 4
#define SCLANE_ERR_INTERNAL 5      /* an unexpected C++ exception was caught at the boundary (message in errbuf) */

/* sclane_record.status bits (the four Model_Status strings of testDynamic.R:266-278) */
#define SCLANE_MARGE_OK 1
#define SCLANE_NULL_OK 2

/* sclane_record.*_notes bits */
#define SCLANE_NOTE_TH_ITER_LIMIT 1     /* theta.ml "iteration limit reached"           (th.warn) */
#define SCLANE_NOTE_ALT_LIMIT 2         /* glm.nb "alternation limit reached"            (th.warn) */
#define SCLANE_NOTE_TH_TRUNC 4          /* theta.ml "estimate truncated at zero"         (th.warn) */
#define SCLANE_NOTE_NOT_CONVERGED 8     /* last glm.fit2 did not converge                          */
#define SCLANE_NOTE_FWD_EMPTY 16        /* forward pass added < 2 columns -> marge2 subscript error (marge2.R:777) */
#define SCLANE_NOTE_ALL_ZERO 32         /* all counts zero in this lineage                         */
#define SCLANE_NOTE_NUMERIC 64          /* non-finite intermediate (R: "missing value where TRUE/FALSE needed") */
#define SCLANE_NOTE_BAD_COUNTS 128      /* a negative count in this lineage: both models fail (glm.nb: "negative values not allowed") */

/* sclane_opts.gee_cor: working correlation of the GEE mode (stat_out_score_gee_null.R:51-60) */
#define SCLANE_COR_INDEPENDENCE 0
#define SCLANE_COR_EXCHANGEABLE 1
#define SCLANE_COR_AR1 2
#define SCLANE_GEE_PMAX 5       /* GEE mode: M <= 5 */
#define SCLANE_GEE_SMAX 64      /* GEE mode: subjects per lineage */

/* hinge direction of a model term */
#define SCLANE_TERM_INTERCEPT 0
#define SCLANE_TERM_TP1 1               /* (x - t)+  "Right", name h_<Lineage>_<knot>  (tp1.R:17, marge2.R:593) */
#define SCLANE_TERM_TP2 2               /* (t - x)+  "Left",  name h_<knot>_<Lineage>  (tp2.R:17, marge2.R:594) */

/* Arguments of testDynamic() (testDynamic.R:60-75) and marge2() (marge2.R:56-72) that reach the path. */
typedef struct sclane_opts {
  int32_t abi_version;        /* must be SCLANE_ABI_VERSION */
  int32_t mode;               /* 0 = GLM (is.gee = FALSE, is.glmm = FALSE).  1 = GEE (is.gee = TRUE, gee.test = "wald");
                                 GEE needs subject ids: use sclane_test_dynamic_gee().  2 (GLMM) -> SCLANE_ERR_UNSUPPORTED */
  int32_t M;                  /* n.potential.basis.fns / marge2 M, default 5, 3 <= M <= SCLANE_PMAX */
  int32_t approx_knot;        /* approx.knot, default 1.  0 = every abscissa kept by min_span() / max_span() is a candidate knot
                                 (marge2.R:167-173): prefix-moment sweep over the abscissae, GLM mode only */
  int32_t n_knot_max;         /* n.knot.max, default 25, <= SCLANE_TMAX */
  int32_t minspan;            /* marge2 minspan; -1 = NULL (automatic, min_span.R:24-26) */
  double  tols_score;         /* marge2 tols_score, default 1e-5 */
  int32_t n_gpus;             /* devices to shard genes over (0 = all visible) */
  int32_t gpu_ids[SCLANE_MAX_GPUS]; /* used when n_gpus > 0 */
  int32_t genes_per_batch;    /* 0 = automatic (bounded by free device memory) */
  int32_t verbose;
  int32_t gee_cor;            /* cor.structure (testDynamic.R:66): SCLANE_COR_* ; default ar1 */
  int32_t gee_test;           /* gee.test (testDynamic.R:68): 0 = "wald" (waldTestGEE.R), 1 = "score" (scoreTestGEE.R) */
  int32_t force_per_cell;     /* 1 = never use the X-compressed kernels (testing / profiling of the per-cell kernels) */
  int32_t gee_bias_correction; /* gee.bias.correction.method (testDynamic.R:67): 0 = NULL (model-based variance), 1 = "kc",
                                 2 = "df"; != 0 switches every GEE fit of the path to the sandwich variance
                                 (testDynamic.R:195,243,286) and waldTestGEE() to biasCorrectGEE() (ar1 / exchangeable only) */
  int32_t h2d_mode;           /* host -> device transfer of the counts: 0 = automatic (staged for a large pageable buffer, else direct), 1 = direct (int32, straight from the caller's
                                 buffer), 2 = staged (host threads narrow each gene batch to uint16 into pinned staging buffers while
                                 the previous batch is in flight: half the PCIe bytes, and full PCIe speed from pageable R memory);
                                 a batch holding a count > 65534 or < 0 always goes direct */
  int32_t comp_warp_per_gene; /* testing / profiling only: 1 = run the X-compressed fits (no offset) on the warp-per-gene persistent
                                 kernels instead of the CTA-per-gene ones (same records; measured: 1.2-1.45x the steady-state
                                 throughput but the slowest gene runs 5x longer on one warp, so the launch is tail bound --
                                 profiles/r02_tail_cycles.txt).  The offset quadrature always runs warp per gene. */
  int32_t offset_per_cell;    /* testing / profiling only: 1 = fit the intercept-only models with an offset per cell instead of
                                 through the offset quadrature */
  int32_t basis_only;         /* 1 = stop after the backward pass and return the selected basis only (marge2(is.glmm = TRUE),
                                 marge2.R:871-879): records carry n_terms / term_* and SCLANE_MARGE_OK, every fitted quantity is NaN.
                                 Set by sclane_marge_basis(). */
  int32_t host_plan;          /* testing / profiling only: 1 = build the per-lineage plan (K0: sort, candidate knots, bucket and
                                 abscissa tables) on the host.  Default 0: on the device whenever it can be (approx_knot = 1, no
                                 size factors, GLM mode, pseudotime range that fits a 24-bit radix key) -- bit-identical tables */
  int32_t batch_streams;      /* gene batches of a shard in flight at once, each on its own stream (0 = default 2, max 4): the
                                 launch tails of one batch's kernels and the host work between batches hide behind the other
                                 batch.  1 = strictly sequential batches (clean per-kernel CUDA-event times for profiling);
                                 records do not depend on it */
} sclane_opts;

/* One record per (gene, lineage): everything testDynamic() returns for that pair except the
 * two per-cell data.frames, which are reproduced on demand by sclane_link_preds().
 * Field sources: testDynamic.R:320-342, pullMARGESummary.R:66-98, pullNullSummary.R:66-80,
 * modelLRT.R:43-53, marge2.R:860-866. */
typedef struct sclane_record {
  int32_t status;                       /* SCLANE_MARGE_OK | SCLANE_NULL_OK */
  int32_t marge_notes;                  /* SCLANE_NOTE_* of the alternative model */
  int32_t null_notes;                   /* SCLANE_NOTE_* of the null model */
  int32_t n_terms;                      /* columns of the final model incl. intercept; 0 if marge2 failed */
  int32_t term_kind[SCLANE_PMAX];       /* SCLANE_TERM_* ; term 0 is the intercept */
  int32_t term_knot_idx[SCLANE_PMAX];   /* index into the lineage's candidate knots (approx_knot = 0: into the gene's own table of
                                           selected knots); -1 for the intercept */
  double  term_knot[SCLANE_PMAX];       /* the knot itself (what tp1/tp2 were built with, marge2.R:589-590) */
  double  term_label[SCLANE_PMAX];      /* signif(knot, 4): the number printed in the coefficient name and parsed back
                                           by extractBreakpoints() (marge2.R:593-594, extractBreakpoints.R:27-28) */
  double  coef[SCLANE_PMAX];            /* MARGE_Summary$estimate  */
  double  se[SCLANE_PMAX];              /* MARGE_Summary$std.error (dispersion 1) */
  double  z[SCLANE_PMAX];               /* MARGE_Summary$statistic */
  double  p[SCLANE_PMAX];               /* MARGE_Summary$p.value = 2*pnorm(-|z|) */
  double  cov[SCLANE_PMAX * SCLANE_PMAX]; /* (X'WX)^-1 of the last IRLS solve, row-major n_terms x n_terms
                                             (vcov() for summarizeModel.R:37-39 and se.fit of predict()) */
  double  theta, se_theta;              /* final_mod$theta, $SE.theta */
  double  loglik, deviance;             /* LogLik_MARGE, Dev_MARGE */
  int32_t alternations, irls_iters;     /* glm.nb alternation count; total glm.fit2 iterations */
  double  null_coef, null_se, null_z, null_p;   /* Null_Summary row */
  double  null_theta, null_se_theta;
  double  null_loglik, null_deviance;   /* LogLik_Null, Dev_Null */
  int32_t null_alternations, null_irls_iters;
  double  test_stat;                    /* Test_Stat  = LRT_Stat = 2 (l1 - l0) */
  double  df;                           /* Degrees_Freedom (0 -> 1 rule, modelLRT.R:48-50) */
  double  p_val;                        /* P_Val = pchisq(LRT, DF, lower = FALSE) */
  /* trace of the discrete path, for parity tests and debugging */
  int32_t fwd_n_terms;                  /* columns after the forward pass */
  int32_t fwd_kind[SCLANE_PMAX];
  int32_t fwd_knot_idx[SCLANE_PMAX];
  double  fwd_score[4];                 /* score2 of each forward iteration (marge2.R:675) */
  double  fwd_sigma[4];                 /* sigma of each forward null fit; 0 = Poisson regime */
  double  wic[SCLANE_PMAX];             /* WIC_vec_2[-1] (marge2.R:778-797) */
  int32_t n_wic;
  int32_t fit_passes;                   /* passes over the cells spent on this record */
  double  gene_time;                    /* Gene_Time surrogate: seconds of device time / genes in the batch */
  /* GEE mode only (NaN / 0 in GLM mode).  In GEE mode se/z/p/cov are the model-based ("naive") ones of geeM::geem,
   * test_stat/df/p_val are waldTestGEE()'s (waldTestGEE.R:24-93) or scoreTestGEE()'s (scoreTestGEE.R:26-139) and
   * theta/loglik/deviance are NaN. */
  double  gee_alpha, gee_phi;           /* working correlation parameter and scale of the final model */
  double  null_gee_alpha, null_gee_phi;
  int32_t gee_converged, gee_iters;     /* geem outer iterations of the final model */
  int32_t null_gee_converged, null_gee_iters;
  /* profiling: SM clock cycles (clock64) the gene spent in the forward / backward / final kernels (whatever unit -- warp or
   * CTA -- fitted it); 0 where a stage did not run.  Used by tools/tail_analysis.py; not part of the reference's record. */
  double  stage_cycles[3];
} sclane_record;

/* Per-lineage description returned to the caller (knots, cell membership). */
typedef struct sclane_lineage_info {
  int64_t n_cells;                      /* cells with a non-NA pseudotime in this lineage (testDynamic.R:183) */
  int32_t n_knots;                      /* number of candidate knots (approx_knot = 0: may exceed SCLANE_TMAX) */
  int32_t minspan;
  double  knots[SCLANE_TMAX];           /* X_red of marge2.R:152-166, ascending (approx_knot = 0: the first SCLANE_TMAX of them) */
  double  pt_min, pt_max;               /* of the unrounded pseudotime (summarizeModel.R:71-72) */
} sclane_lineage_info;

/* ---------------------------------------------------------------------------------------
 * Library / device queries (no compute)
 * ------------------------------------------------------------------------------------- */
int32_t sclane_abi_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int32_t sclane_device_count(void);
/* fills *opts with the defaults of testDynamic.R:60-75 / marge2.R:56-72 */
void sclane_default_opts(sclane_opts* opts);
size_t sclane_record_size(void);
/* kernels launched by this process since load / since the last reset (the bench's gpu_launches claim) */
int64_t sclane_launch_count(void);
void sclane_reset_launch_count(void);

/* Device-side timing of the last sclane_test_dynamic*() call on this process (CUDA events on the
 * library's own stream, summed over batches / lineages / GPUs): what bench.py's roofline uses. */
#define SCLANE_NSTAGES 15
#define SCLANE_STAGE_GATHER 0      /* sort permutation + per-gene constants                                  */
#define SCLANE_STAGE_FWD_FIT 1     /* K1 forward null fits (stat_out_score_glm_null)                         */
#define SCLANE_STAGE_SWEEP 2       /* K2a score sweep, streaming kernel k_sweep_moments (HBM bound)          */
#define SCLANE_STAGE_BACKWARD 3    /* K4 backward WIC pass                                                   */
#define SCLANE_STAGE_FINAL 4       /* K5 final + null glm.nb                                                 */
#define SCLANE_STAGE_COPY 5        /* H2D of counts + D2H of records                                         */
#define SCLANE_STAGE_SELECT 6      /* K2b score sweep, per-gene algebra + selection tree (k_sweep_select)    */
#define SCLANE_STAGE_GEE 7         /* GEE mode: the whole per-gene path (k_gee)                              */
#define SCLANE_STAGE_AGGREGATE 8   /* X-compressed path: per-point sums + count histogram (k_aggregate)      */
#define SCLANE_STAGE_COMP_FWD 9    /* X-compressed forward pass: null fits + score sweep + selection         */
#define SCLANE_STAGE_COMP_BWD 10   /* X-compressed backward WIC pass                                         */
#define SCLANE_STAGE_COMP_FINAL 11 /* X-compressed final + null glm.nb (no offset)                            */
#define SCLANE_STAGE_OFF_AGG 12    /* offset quadrature: per-gene node weights (k_off_aggregate)               */
#define SCLANE_STAGE_OFF_NULL 13   /* offset quadrature: intercept-only glm.nb with the size-factor offset     */
#define SCLANE_STAGE_PLAN 14       /* K0 on the device: sort by round(x, 4), candidate knots, bucket / abscissa tables */
typedef struct sclane_profile {
  double  ms[SCLANE_NSTAGES];          /* device milliseconds per stage: CUDA-event intervals on the stream of each gene batch.  With
                                          batch_streams > 1 (the default is 2) batches run concurrently and their intervals overlap:
                                          the sum can exceed wall_ms; set batch_streams = 1 for a breakdown that adds up            */
  int64_t launches[SCLANE_NSTAGES];    /* kernel launches (memcpys for SCLANE_STAGE_COPY) per stage        */
  double  sweep_bytes;                 /* algorithmic bytes the sweep launches moved: sum of 12*N_l per
                                          (gene, forward iteration actually swept) + 8*N_l per launch       */
  int64_t sweep_gene_iters;            /* (gene, iteration) pairs swept                                    */
  double  h2d_bytes, d2h_bytes;
  double  wall_ms;                     /* host wall clock of the call                                      */
} sclane_profile;
void sclane_get_profile(sclane_profile* out);
/* frees the per-device work buffers the library keeps between calls */
void sclane_release(void);

/* ---------------------------------------------------------------------------------------
 * The hot path.  Replaces the cluster set-up, FBM creation and the foreach() of
 * testDynamic.R:135-385 (GLM mode): for every gene and lineage the marge2() forward pass
 * (marge2.R:109-759), backward WIC pass (:763-812), final glm.nb fit (:841-847), the
 * intercept-only null glm.nb (testDynamic.R:254-260) and modelLRT (modelLRT.R:43-53).
 *
 * counts        n_cells x n_genes column-major int32 == gene-major: gene g's cells are
 *               counts[g*n_cells .. (g+1)*n_cells) -- the bytes of the bigstatsr FBM the
 *               reference builds at testDynamic.R:145-147.  HOST pointer.
 * pt            n_cells x n_lineages column-major doubles, NaN = cell not in lineage (:183). HOST.
 * size_factor   n_cells doubles (createCellOffset(), enters as offset(log(1/sf)),
 *               marge2.R:826-829, testDynamic.R:228-231) or NULL.  HOST.
 * out           n_genes x n_lineages records, gene-major (record of gene g, lineage l at
 *               out[g*n_lineages + l]); caller allocated.
 * lineages      n_lineages entries or NULL.
 * ------------------------------------------------------------------------------------- */
int32_t sclane_test_dynamic(const int32_t* counts, int64_t n_cells, int64_t n_genes,
                            const double* pt, int32_t n_lineages,
                            const double* size_factor,
                            const sclane_opts* opts,
                            sclane_record* out, sclane_lineage_info* lineages,
                            char* errbuf, size_t errlen);

/* Same computation with `counts` already resident in device memory on `device`
 * (the "inputs in HBM" measurement of bench.py; also what a caller uses to keep a count
 * matrix on the GPU across calls).  pt / size_factor / out stay host pointers (small). */
int32_t sclane_test_dynamic_device(const int32_t* d_counts, int32_t device,
                                   int64_t n_cells, int64_t n_genes,
                                   const double* pt, int32_t n_lineages,
                                   const double* size_factor,
                                   const sclane_opts* opts,
                                   sclane_record* out, sclane_lineage_info* lineages,
                                   char* errbuf, size_t errlen);

/* Same computation from a SPARSE count matrix in the layout Seurat / SingleCellExperiment hold it: a genes x cells
 * dgCMatrix (compressed by cell column), i.e. what testDynamic.R:80-97 densifies and transposes on the host.
 * indptr  = dgCMatrix@p  (n_cells + 1 int32),  indices = dgCMatrix@i  (nnz int32 gene rows),  data = dgCMatrix@x (nnz
 * doubles holding integer counts).  HOST pointers.  12 B per non-zero cross PCIe instead of 4 B per cell and gene (real
 * scRNA-seq matrices are ~90 % zeros); every GPU expands its own gene shard to the dense gene-major layout on the device.
 * GLM mode only. */
int32_t sclane_test_dynamic_csc(const int32_t* indptr, const int32_t* indices, const double* data,
                                int64_t n_cells, int64_t n_genes,
                                const double* pt, int32_t n_lineages,
                                const double* size_factor,
                                const sclane_opts* opts,
                                sclane_record* out, sclane_lineage_info* lineages,
                                char* errbuf, size_t errlen);

/* marge2(..., is.glmm = TRUE) for every (gene, column of pt): forward pass + WIC backward pass only, no final fit -- the knot
 * harvesting of fitGLMM() (fitGLMM.R:54-64 calls it once per subject with M = M.glm = 3: pass one pt column per subject, NaN
 * outside the subject's cells).  Records: status & SCLANE_MARGE_OK when a basis was selected, n_terms / term_kind / term_knot /
 * term_label / term_knot_idx filled (marge_coef_names of marge2.R:877), all fitted quantities NaN.  The offset plays no role in
 * the forward / backward pass (marge2.R:826-829 adds it to the final fit only).  Same arguments as sclane_test_dynamic(). */
int32_t sclane_marge_basis(const int32_t* counts, int64_t n_cells, int64_t n_genes,
                           const double* pt, int32_t n_lineages,
                           const sclane_opts* opts,
                           sclane_record* out, sclane_lineage_info* lineages,
                           char* errbuf, size_t errlen);

/* GEE mode of the same path (testDynamic(is.gee = TRUE, gee.test = "wald"), testDynamic.R:66-70,235-244,
 * 296-301; marge2.R:258-268 (stat_out_score_gee_null), :438-448/:520-530 (score_fun_gee), :831-838 (final geem);
 * backward_sel_WIC.R:36-42; waldTestGEE.R:24-93).  Same arguments as sclane_test_dynamic() plus
 * subject_id   n_cells int32 subject labels; the cells MUST be ordered by subject (testDynamic.R:115 stops
 *              otherwise) -> SCLANE_ERR_ARG if a label re-appears after a different one.  HOST pointer.
 * opts->gee_cor selects the working correlation; opts->M <= SCLANE_GEE_PMAX; <= SCLANE_GEE_SMAX subjects per
 * lineage.  opts->gee_test selects waldTestGEE() or scoreTestGEE(); opts->gee_bias_correction the sandwich variance with
 * the "kc" / "df" small-sample correction of biasCorrectGEE.R.
 * Records: see the GEE note in sclane_record. */
int32_t sclane_test_dynamic_gee(const int32_t* counts, int64_t n_cells, int64_t n_genes,
                                const double* pt, int32_t n_lineages,
                                const int32_t* subject_id,
                                const double* size_factor,
                                const sclane_opts* opts,
                                sclane_record* out, sclane_lineage_info* lineages,
                                char* errbuf, size_t errlen);

/* Candidate knots of one lineage (marge2.R:150-173 with min_span.R / max_span.R): host-side,
 * no GPU needed.  x: n doubles without NaN.  Returns the number of knots (<= SCLANE_TMAX)
 * or a negative error code. */
int32_t sclane_candidate_knots(const double* x, int64_t n, const sclane_opts* opts,
                               double* knots_out, int32_t* minspan_out);

/* K0 in full: the per-lineage plan every gene of the lineage shares (marge2.R:150-166 + the sort, bucket and abscissa tables
 * of DESIGN.md section 2), built on the HOST (on_device = 0; no GPU needed) or by the device kernels of the batch path
 * (on_device = 1; SCLANE_ERR_UNSUPPORTED when the input is one the batch path would hand to the host: size factors,
 * approx_knot = 0, a pseudotime range beyond the radix key).  pt: n_cells doubles, NaN = not in the lineage.  Outputs (any may
 * be NULL): perm [N] sorted position -> cell; u [N] bucket-local coordinate; knots [T]; bstart [T + 2]; U [3 (T + 1)] =
 * U0 | U1 | U2; pstart [D + 1]; pn, pu [D]; pb [D]; pbstart [T + 2]; dims = {N, T, D, minspan}.  Arrays must hold n_cells + 1
 * entries (perm, u, pstart, pn, pu, pb) resp. SCLANE_TMAX + 2 / 3 (SCLANE_TMAX + 1).  The two builds are bit-identical
 * (tests/test_parity_r02.py). */
int32_t sclane_lineage_plan(const double* pt, int64_t n_cells, const sclane_opts* opts, int32_t on_device, int32_t device,
                            int32_t* perm, double* u, double* knots, int32_t* bstart, double* U, int32_t* pstart, double* pn,
                            double* pu, uint8_t* pb, int32_t* pbstart, int32_t* dims, char* errbuf, size_t errlen);

/* ---------------------------------------------------------------------------------------
 * Single-step entry points used by the unit tests (and by an R-side marge2() wrapper).
 * All pointers are HOST pointers; each call uploads, runs the same kernels as the batch
 * path on `device`, and downloads.
 * ------------------------------------------------------------------------------------- */

/* stat_out_score_glm_null() (stat_out_score_glm_null.R:18-48) for a batch of genes that share
 * one basis given as (kind, knot) terms: NB fit of Y ~ B - 1, returns per gene beta (n_terms),
 * sigma (0 = Poisson regime) and, if mu_out != NULL, the fitted means (n_genes x n_cells,
 * gene-major, in the CALLER's cell order). */
int32_t sclane_null_stats_glm(const int32_t* counts, int64_t n_cells, int64_t n_genes,
                              const double* x, const double* knots, int32_t n_knots,
                              const int32_t* term_kind, const int32_t* term_knot_idx, int32_t n_terms,
                              int32_t device,
                              double* beta_out, double* sigma_out, double* mu_out,
                              char* errbuf, size_t errlen);

/* The score sweep: score_fun_glm() (score_fun_glm.R:21-54) for every candidate knot and both
 * candidate shapes ("one" = tp1 only, "both" = tp1+tp2), given counts, fitted means mu and
 * sigma of the working null model with basis `terms`.  scores_out: n_genes x 2 x n_knots
 * ([g][0][k] = one, [g][1][k] = both), NaN = NA (rank rule).  This is the HBM-bound kernel
 * (12 bytes per cell per gene: int32 count + fp64 mu). */
int32_t sclane_score_glm(const int32_t* counts, const double* mu, int64_t n_cells, int64_t n_genes,
                         const double* x, const double* knots, int32_t n_knots,
                         const int32_t* term_kind, const int32_t* term_knot_idx, int32_t n_terms,
                         const double* sigma, int32_t device,
                         double* scores_out,
                         char* errbuf, size_t errlen);

/* Per-cell link-scale fit and standard error of a fitted record -- the MARGE_Preds /
 * Null_Preds data.frames (pullMARGESummary.R:66-69, pullNullSummary.R:69-72:
 * predict(type = "link", se.fit = TRUE)).  x: the lineage's pseudotime (n doubles, unrounded;
 * rounded internally as marge2.R:154), size_factor may be NULL.  Outputs: n doubles each,
 * any may be NULL. */
int32_t sclane_link_preds(const sclane_record* rec, const double* x, int64_t n,
                          const double* size_factor, int32_t device,
                          double* marge_link_fit, double* marge_link_se,
                          double* null_link_fit, double* null_link_se,
                          char* errbuf, size_t errlen);

/* smoothedCountsMatrix() (smoothedCountsMatrix.R:27-74) for one lineage, all genes in one launch: the fitted mean of every
 * cell under every gene's final model, exp(marge_link_fit) (* size_factor when one is given, :57-61), optionally log1p
 * (:65-67).  recs: the n_genes x n_lineages records of sclane_test_dynamic*(); lineage: which one; x / size_factor: the
 * lineage's n cells (pseudotime unrounded, rounded internally).  is_gee: the link fit of a GEE record has no offset
 * (predict.geem).  out: n x n_genes column-major (gene g's cells at out[g*n ..)); genes whose model failed get NaN (the
 * reference drops them from the matrix, :50-52).  HOST pointers. */
int32_t sclane_smoothed_counts(const sclane_record* recs, int64_t n_genes, int32_t n_lineages, int32_t lineage,
                               const double* x, int64_t n, const double* size_factor,
                               int32_t is_gee, int32_t log1p_norm, int32_t device,
                               double* out, char* errbuf, size_t errlen);

/* ---------------------------------------------------------------------------------------
 * Drop-ins for the reference's three RcppEigen helpers (column-major doubles, HOST pointers).
 * ------------------------------------------------------------------------------------- */
/* eigenMapMatMult(A, B, n_cores) -- src/eigenMapMatMult.cpp:6-15.  C (m x n) = A (m x k) * B (k x n). */
int32_t sclane_matmul(const double* A, const double* B, double* C, int64_t m, int64_t k, int64_t n,
                      int32_t device, char* errbuf, size_t errlen);
/* eigenMapMatrixInvert(A, n_cores) -- src/eigenMapInvert.cpp:6-11: A.llt().solve(I), n x n,
 * lower triangle read, NO positive-definiteness check (garbage in, garbage out, as Eigen). n <= 64. */
int32_t sclane_llt_inverse(const double* A, double* Ainv, int64_t n,
                           int32_t device, char* errbuf, size_t errlen);
/* eigenMapPseudoInverse(A, tolerance, n_cores) -- src/eigenMapPseudoInverse.cpp:6-21:
 * Jacobi SVD, singular values > tolerance (absolute) inverted.  A is m x n, out is n x m. m,n <= 48. */
int32_t sclane_pinv(const double* A, double* Apinv, int64_t m, int64_t n, double tolerance,
                    int32_t device, char* errbuf, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* SCLANE_B200_H */
