# testDynamic_b200.R -- R side of the sclane_b200 drop-in: replaces the cluster set-up, FBM creation and foreach() of
# testDynamic() (R/testDynamic.R:135-385) by ONE .Call into libsclane_b200.so, and keeps marge2() / modelLRT() as thin wrappers
# over the same library.  Source only: R is not installed in the build image.  The C half (r/sclane_b200_shim.c) builds the whole
# 21-field result list, so nothing below re-implements record assembly; that C code is compiled against a stand-in for R's C API
# and checked against the reference's bundled output in tests/test_r_shim.py.
#
# Drop into the package as R/testDynamic_b200.R + src/sclane_b200_shim.c, add `useDynLib(sclaneB200, .registration = TRUE)`.

.sclane_opts <- function(M = 5, approx.knot = TRUE, n.gpus = 0L, cor.structure = "ar1", gee.test = "wald", gee.bias.correction.method = NULL,
                         n.knot.max = 25, minspan = NULL) {
  cor_code <- match(tolower(cor.structure), c("independence", "exchangeable", "ar1")) - 1L
  if (is.na(cor_code)) stop("GEE models require a specified correlation structure.")
  bias <- if (is.null(gee.bias.correction.method)) 0L else match(tolower(gee.bias.correction.method), c("kc", "df"))
  if (is.na(bias)) stop("Unsupported bias correction method in waldTestGEE().")
  c(as.integer(M), as.integer(approx.knot), as.integer(n.gpus), cor_code, as.integer(tolower(gee.test) == "score"), bias,
    as.integer(n.knot.max), if (is.null(minspan)) -1L else as.integer(minspan))
}

# Per-cell predictions of one fitted (gene, lineage) record, on demand: the reference stores two N x 2 data.frames per gene and
# lineage (testDynamic.R:339-340; 128 GB at 20,000 genes x 200,000 cells), the drop-in computes them when they are asked for.
sclane_link_preds <- function(rec, pt.lineage, size.factor.lineage = NULL) {
  .Call("C_sclane_link_preds", attr(rec, "sclane_record"), as.double(pt.lineage), if (is.null(size.factor.lineage)) NULL else as.double(size.factor.lineage))
}

# fills MARGE_Preds / Null_Preds of every record eagerly (what the reference returns); call only when the result fits in memory
sclane_fill_preds <- function(test.dyn.res, pt, size.factor.offset = NULL, is.gee = FALSE) {
  pt <- as.matrix(pt)
  for (g in seq_along(test.dyn.res)) for (j in seq_len(ncol(pt))) {
    cells <- which(!is.na(pt[, j]))
    rec <- test.dyn.res[[g]][[j]]
    # predict.geem() returns X %*% beta without the offset (pullMARGESummary.R:38-41)
    p <- sclane_link_preds(rec, pt[cells, j], if (is.null(size.factor.offset) || is.gee) NULL else size.factor.offset[cells])
    if (!is.null(rec$MARGE_Summary)) test.dyn.res[[g]][[j]]$MARGE_Preds <- data.frame(marge_link_fit = p$marge_link_fit, marge_link_se = p$marge_link_se)
    if (!is.null(rec$Null_Summary)) test.dyn.res[[g]][[j]]$Null_Preds <- data.frame(null_link_fit = p$null_link_fit, null_link_se = p$null_link_se)
  }
  test.dyn.res
}

# testDynamic() with the same 16 arguments as R/testDynamic.R:60-75 (+ n.gpus, + eager.preds).  `n.cores` is kept for signature
# compatibility (the library sizes its own host staging threads); is.glmm = TRUE falls through to the reference implementation.
testDynamic <- function(expr.mat = NULL, pt = NULL, genes = NULL, size.factor.offset = NULL, is.gee = FALSE, cor.structure = "ar1",
                        gee.bias.correction.method = NULL, gee.test = "wald", is.glmm = FALSE, glmm.adaptive = TRUE, id.vec = NULL,
                        n.potential.basis.fns = 5, n.cores = 4L, approx.knot = TRUE, verbose = TRUE, random.seed = 312,
                        n.gpus = 0L, eager.preds = FALSE) {
  if (is.null(expr.mat) || is.null(pt)) { stop("You forgot some inputs to testDynamic().") }          # testDynamic.R:77
  if (is.glmm) return(scLANE::testDynamic(expr.mat = expr.mat, pt = pt, genes = genes, size.factor.offset = size.factor.offset, is.glmm = TRUE,
                                          glmm.adaptive = glmm.adaptive, id.vec = id.vec, n.potential.basis.fns = n.potential.basis.fns,
                                          n.cores = n.cores, approx.knot = approx.knot, verbose = verbose, random.seed = random.seed))
  # counts extraction (testDynamic.R:80-98), keeping sparse inputs sparse until the upload
  if (inherits(expr.mat, "SingleCellExperiment")) {
    expr.mat <- BiocGenerics::counts(expr.mat)
  } else if (inherits(expr.mat, "Seurat")) {
    expr.mat <- Seurat::GetAssayData(expr.mat, slot = "counts", assay = Seurat::DefaultAssay(expr.mat))
  } else if (inherits(expr.mat, "cell_data_set")) {
    expr.mat <- BiocGenerics::counts(expr.mat)
  }
  if (!(inherits(expr.mat, "matrix") || inherits(expr.mat, "array") || inherits(expr.mat, "dgCMatrix"))) {
    stop("Input expr.mat must be coerceable to a matrix of integer counts.")
  }
  if (is.null(genes)) genes <- rownames(expr.mat)
  expr.mat <- expr.mat[genes, , drop = FALSE]                                                             # genes x cells
  if (!inherits(pt, "data.frame")) { stop("pt must be of class data.frame.") }                           # testDynamic.R:98
  n_lineages <- ncol(pt)
  if (is.gee) {                                                                                           # testDynamic.R:107-120
    if (is.null(id.vec)) { stop("You must provide a vector of IDs if you're using GEE / GLMM modes.") }
    if (is.unsorted(id.vec)) { stop("Your data must be ordered by subject, please do so before running testDynamic() with the GEE / GLMM modes.") }
    cor.structure <- tolower(cor.structure)
    if (!cor.structure %in% c("ar1", "independence", "exchangeable")) { stop("GEE models require a specified correlation structure.") }
  }
  pt_mat <- as.matrix(pt); storage.mode(pt_mat) <- "double"
  sf <- if (is.null(size.factor.offset)) NULL else as.double(size.factor.offset)
  opts <- .sclane_opts(n.potential.basis.fns, approx.knot, n.gpus, cor.structure, gee.test, gee.bias.correction.method)
  t0 <- Sys.time()
  if (is.gee) {
    dense <- t(as.matrix(expr.mat)); storage.mode(dense) <- "integer"                                    # cells x genes, as testDynamic.R:97
    res <- .Call("C_sclane_test_dynamic_gee", dense, pt_mat, as.integer(factor(id.vec, levels = unique(id.vec))), sf, as.character(genes), opts)
  } else if (inherits(expr.mat, "dgCMatrix")) {
    res <- .Call("C_sclane_test_dynamic_csc", expr.mat@p, expr.mat@i, as.double(expr.mat@x), expr.mat@Dim, pt_mat, sf, as.character(genes), opts)
  } else {
    dense <- t(expr.mat); storage.mode(dense) <- "integer"
    res <- .Call("C_sclane_test_dynamic", dense, pt_mat, sf, as.character(genes), opts)
  }
  if (eager.preds) res <- sclane_fill_preds(res, pt_mat, sf, is.gee)
  if (verbose) {                                                                                          # testDynamic.R:428-443
    total_time <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
    message(paste0("scLANE testing in ", if (is.gee) "GEE" else "GLM", " mode completed for ", length(genes), " genes across ", n_lineages, " ",
                   ifelse(n_lineages == 1, "lineage ", "lineages "), "in ", round(total_time, 3), " secs"))
  }
  res                                                                                                     # class "scLANE" set by the C half
}

# marge2() (R/marge2.R:56-72) as a single-gene call of the same library: returns final_mod-like fields, coef_names and
# marge_coef_names.  is.glmm = TRUE returns the selected basis only (marge2.R:871-879), as fitGLMM() uses it.
marge2 <- function(X_pred = NULL, Y = NULL, Y.offset = NULL, M = 5, is.gee = FALSE, is.glmm = FALSE, id.vec = NULL, cor.structure = "ar1",
                   sandwich.var = FALSE, approx.knot = TRUE, n.knot.max = 25, glm.backend = "MASS", tols_score = 1e-5, minspan = NULL,
                   return.basis = FALSE, return.WIC = FALSE, return.GCV = FALSE) {
  if (is.null(X_pred) || is.null(Y)) { stop("Some required inputs to marge2() are missing.") }
  if (is.gee & is.null(id.vec)) { stop("id.vec in marge2() must be non-null if is.gee = TRUE.") }
  if (tols_score != 1e-5) { stop("tols_score other than the default: set sclane_opts.tols_score through the C ABI.") }
  x <- matrix(as.double(as.matrix(X_pred)[, 1]), ncol = 1)
  y <- matrix(as.integer(Y), ncol = 1)
  if (is.glmm) {
    nm <- .Call("C_sclane_marge_basis", y, x, .sclane_opts(M, approx.knot, n.knot.max = n.knot.max, minspan = minspan))[[1]][[1]]
    if (is.null(nm)) stop("marge2() failed: the forward pass added fewer than two basis functions")
    res <- list(final_mod = NULL, basis_mtx = NULL, WIC_mtx = NULL, GCV = NULL, model_type = "GLMM", coef_names = nm, marge_coef_names = nm)
    class(res) <- "marge"
    return(res)
  }
  opts <- .sclane_opts(M, approx.knot, 1L, cor.structure, "wald", if (sandwich.var) "df" else NULL, n.knot.max, minspan)
  rec <- if (is.gee) {
    .Call("C_sclane_test_dynamic_gee", y, x, as.integer(factor(id.vec, levels = unique(id.vec))), if (is.null(Y.offset)) NULL else as.double(Y.offset), "gene", opts)[[1]][[1]]
  } else {
    .Call("C_sclane_test_dynamic", y, x, if (is.null(Y.offset)) NULL else as.double(Y.offset), "gene", opts)[[1]][[1]]
  }
  if (is.null(rec$MARGE_Summary)) stop(rec$MARGE_Fit_Notes)
  final_mod <- list(coefficients = stats::setNames(rec$MARGE_Summary$estimate, rec$MARGE_Summary$term), std.error = rec$MARGE_Summary$std.error,
                    logLik = rec$LogLik_MARGE, df = nrow(rec$MARGE_Summary) + 1L, deviance = rec$Dev_MARGE, record = attr(rec, "sclane_record"))
  res <- list(final_mod = final_mod, basis_mtx = NULL, WIC_mtx = NULL, GCV = NULL, model_type = ifelse(is.gee, "GEE", "GLM"),
              coef_names = rec$MARGE_Summary$term, marge_coef_names = rec$MARGE_Summary$term)
  class(res) <- "marge"
  res
}

# modelLRT() (R/modelLRT.R:12-63) for the wrappers above: mod.1 = marge2() result, mod.0 = list(logLik, df) of the null model
modelLRT <- function(mod.1 = NULL, mod.0 = NULL, is.glmm = FALSE) {
  if (inherits(mod.1, "try-error") || inherits(mod.0, "try-error") || is.null(mod.1) || is.null(mod.0)) {
    return(list(Alt_Mod_LL = NA_real_, Null_Mod_LL = NA_real_, LRT_Stat = NA_real_, DF = NA_real_, P_Val = NA_real_, Model_Type = NA_character_,
                Notes = "No test performed due to model failure."))
  }
  ll1 <- mod.1$final_mod$logLik; ll0 <- mod.0$logLik
  lrt <- 2 * (ll1 - ll0)
  dgr_free <- mod.1$final_mod$df - mod.0$df
  if (dgr_free == 0) dgr_free <- 1                                                                        # modelLRT.R:48-50
  list(Alt_Mod_LL = ll1, Null_Mod_LL = ll0, LRT_Stat = lrt, DF = dgr_free, P_Val = stats::pchisq(lrt, dgr_free, lower = FALSE), Model_Type = "GLM",
       Notes = NA_character_)
}

# smoothedCountsMatrix() (R/smoothedCountsMatrix.R:27-74): one kernel launch per lineage over all genes
smoothedCountsMatrix <- function(test.dyn.res = NULL, size.factor.offset = NULL, pt = NULL, genes = NULL, log1p.norm = FALSE, n.cores = 2L, is.gee = FALSE) {
  if (is.null(test.dyn.res) || is.null(pt)) { stop("Please provide the scLANE output from testDynamic().") }
  if (is.null(genes)) genes <- names(test.dyn.res)
  pt <- as.matrix(pt)
  out <- lapply(seq_len(ncol(pt)), function(j) {
    cells <- which(!is.na(pt[, j]))
    recs <- lapply(genes, function(g) attr(test.dyn.res[[g]][[j]], "sclane_record"))
    m <- .Call("C_sclane_smoothed_counts", recs, as.double(pt[cells, j]), if (is.null(size.factor.offset)) NULL else as.double(size.factor.offset[cells]),
               as.integer(is.gee), as.integer(log1p.norm))
    colnames(m) <- genes
    m[, !is.nan(m[1, ]), drop = FALSE]                                                                    # failed models are dropped (:50-52)
  })
  names(out) <- paste0("Lineage_", LETTERS[seq_len(ncol(pt))])
  out
}

# the adaptive knot harvesting of fitGLMM() (R/fitGLMM.R:54-75): one library call with a pseudotime column per subject
sclane_glmm_knots <- function(X_pred, Y, id.vec, M.glm = 3, approx.knot = TRUE) {
  if (is.unsorted(id.vec)) { stop("Your data must be ordered by subject, please do so before running fitGLMM().") }
  x <- as.double(as.matrix(X_pred)[, 1])
  subj <- unique(id.vec)
  pt <- vapply(subj, function(s) ifelse(id.vec == s, x, NA_real_), numeric(length(x)))
  nm <- .Call("C_sclane_marge_basis", matrix(as.integer(Y), ncol = 1), pt, .sclane_opts(M.glm, approx.knot))[[1]]
  keep <- which(vapply(nm, function(v) length(v) > 1, logical(1)))
  coefs <- unlist(lapply(nm[keep], function(v) v[-1]))
  data.frame(knot = as.numeric(gsub("h_|_?Lineage_A_?", "", coefs)), coef = coefs, tp_fun = ifelse(grepl("h_[0-9]", coefs), "tp2", "tp1"))
}
