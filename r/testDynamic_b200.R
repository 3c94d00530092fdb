# testDynamic_b200.R -- R-side replacement of the cluster set-up, FBM creation and foreach() of testDynamic()
# (R/testDynamic.R:135-385) by one .Call into libsclane_b200.so.  Source only (no R in the build image); the same list
# assembly is implemented and tested in Python (sclane_b200/api.py::_record_to_list), which this mirrors field by field.
#
# Everything before (input coercion, :76-121) and after (class, message, :386-447) stays as in the reference, so the
# signature of testDynamic() is unchanged; `n.cores` is kept for compatibility and a new `n.gpus` picks the devices.

testDynamicB200 <- function(expr.mat, pt, genes, size.factor.offset = NULL, is.gee = FALSE, id.vec = NULL,
                            cor.structure = "ar1", gee.test = "wald", gee.bias.correction.method = NULL, n.potential.basis.fns = 5,
                            approx.knot = TRUE, n.gpus = 0L) {
  # expr.mat: dense cells x genes integer matrix exactly as built at testDynamic.R:97
  storage.mode(expr.mat) <- "integer"
  pt_mat <- as.matrix(pt); storage.mode(pt_mat) <- "double"
  if (is.gee) {
    # testDynamic.R:107-118 has already checked that the cells are ordered by subject
    cor_code <- match(tolower(cor.structure), c("independence", "exchangeable", "ar1")) - 1L
    if (is.na(cor_code)) stop("GEE models require a specified correlation structure.")
    raw <- .Call("C_sclane_test_dynamic_gee", expr.mat, pt_mat, as.integer(factor(id.vec, levels = unique(id.vec))),
                 size.factor.offset,
                 c(as.integer(n.potential.basis.fns), as.integer(approx.knot), as.integer(n.gpus), cor_code,
                   as.integer(tolower(gee.test) == "score"),
                   if (is.null(gee.bias.correction.method)) 0L else match(tolower(gee.bias.correction.method), c("kc", "df"))))
  } else {
    raw <- .Call("C_sclane_test_dynamic", expr.mat, pt_mat, size.factor.offset,
                 c(as.integer(n.potential.basis.fns), as.integer(approx.knot), as.integer(n.gpus)))
  }
  recs <- sclane_unpack_records(raw, n_genes = ncol(expr.mat), n_lineages = ncol(pt_mat))   # readBin() over the struct layout
  n_lineages <- ncol(pt_mat)
  test_stats <- lapply(seq_along(genes), function(i) {
    lineage_list <- lapply(seq(n_lineages), function(j) {
      r <- recs[[(i - 1) * n_lineages + j]]
      lin <- LETTERS[j]
      marge_ok <- bitwAnd(r$status, 1L) > 0; null_ok <- bitwAnd(r$status, 2L) > 0
      terms <- if (marge_ok) c("Intercept", vapply(seq_len(r$n_terms - 1) + 1, function(k) {
        lab <- as.character(r$term_label[k])                                  # signif(knot, 4), marge2.R:593-594
        if (r$term_kind[k] == 1L) paste0("h_Lineage_", lin, "_", lab) else paste0("h_", lab, "_Lineage_", lin)
      }, character(1))) else character(0)
      list(Gene = genes[i], Lineage = lin,
           Test_Stat = r$test_stat,
           Test_Stat_Type = if (is.gee) ifelse(tolower(gee.test) == "score", "Score", "Wald") else "LRT",   # testDynamic.R:324
           Test_Stat_Note = if (marge_ok && null_ok) NA_character_ else "No test performed due to model failure.",
           Degrees_Freedom = r$df, P_Val = r$p_val,
           LogLik_MARGE = r$loglik, LogLik_Null = r$null_loglik, Dev_MARGE = r$deviance, Dev_Null = r$null_deviance,
           Model_Status = paste0("MARGE model ", if (marge_ok) "OK" else "error", ", null model ", if (null_ok) "OK" else "error"),
           MARGE_Fit_Notes = NA_character_, Null_Fit_Notes = NA_character_, Gene_Time = r$gene_time,
           MARGE_Summary = if (marge_ok) data.frame(term = terms, estimate = r$coef[seq_len(r$n_terms)],
                                                    std.error = r$se[seq_len(r$n_terms)], statistic = r$z[seq_len(r$n_terms)],
                                                    p.value = r$p[seq_len(r$n_terms)]) else NULL,
           Null_Summary = if (null_ok) data.frame(term = "Intercept", estimate = r$null_coef, std.error = r$null_se,
                                                  statistic = r$null_z, p.value = r$null_p) else NULL,
           MARGE_Preds = NULL, Null_Preds = NULL,      # lazily: sclane_link_preds() on demand (see INTEGRATION.md)
           MARGE_Slope_Data = sclane_slope_data(r, terms, pt_mat[!is.na(pt_mat[, j]), j], genes[i], lin),   # createSlopeTestData.R:15-74
           Gene_Dynamics = sclane_gene_dynamics(r, range(pt_mat[, j], na.rm = TRUE), genes[i], lin))        # summarizeModel.R:28-140
    })
    names(lineage_list) <- paste0("Lineage_", LETTERS[seq(n_lineages)])
    lineage_list
  })
  names(test_stats) <- genes
  class(test_stats) <- "scLANE"
  test_stats
}
