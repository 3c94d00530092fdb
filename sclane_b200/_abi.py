"""ctypes mirror of include/sclane_b200.h (struct layouts + prototypes)."""
from __future__ import annotations

import ctypes as C

SCLANE_ABI_VERSION = 4
SCLANE_PMAX = 7
SCLANE_TMAX = 32
SCLANE_MAX_GPUS = 16

SCLANE_MARGE_OK = 1
SCLANE_NULL_OK = 2

NOTE_TH_ITER_LIMIT = 1
NOTE_ALT_LIMIT = 2
NOTE_TH_TRUNC = 4
NOTE_NOT_CONVERGED = 8
NOTE_FWD_EMPTY = 16
NOTE_ALL_ZERO = 32
NOTE_NUMERIC = 64
NOTE_BAD_COUNTS = 128

TERM_INTERCEPT, TERM_TP1, TERM_TP2 = 0, 1, 2

COR_INDEPENDENCE, COR_EXCHANGEABLE, COR_AR1 = 0, 1, 2
COR_CODES = {"independence": COR_INDEPENDENCE, "exchangeable": COR_EXCHANGEABLE, "ar1": COR_AR1}
SCLANE_GEE_PMAX = 5
SCLANE_GEE_SMAX = 64


class SclaneOpts(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("mode", C.c_int32),
        ("M", C.c_int32),
        ("approx_knot", C.c_int32),
        ("n_knot_max", C.c_int32),
        ("minspan", C.c_int32),
        ("tols_score", C.c_double),
        ("n_gpus", C.c_int32),
        ("gpu_ids", C.c_int32 * SCLANE_MAX_GPUS),
        ("genes_per_batch", C.c_int32),
        ("verbose", C.c_int32),
        ("gee_cor", C.c_int32),
        ("gee_test", C.c_int32),
        ("force_per_cell", C.c_int32),
        ("gee_bias_correction", C.c_int32),
        ("h2d_mode", C.c_int32),
        ("comp_warp_per_gene", C.c_int32),
        ("offset_per_cell", C.c_int32),
        ("basis_only", C.c_int32),
        ("host_plan", C.c_int32),
        ("batch_streams", C.c_int32),
    ]


class SclaneRecord(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("marge_notes", C.c_int32),
        ("null_notes", C.c_int32),
        ("n_terms", C.c_int32),
        ("term_kind", C.c_int32 * SCLANE_PMAX),
        ("term_knot_idx", C.c_int32 * SCLANE_PMAX),
        ("term_knot", C.c_double * SCLANE_PMAX),
        ("term_label", C.c_double * SCLANE_PMAX),
        ("coef", C.c_double * SCLANE_PMAX),
        ("se", C.c_double * SCLANE_PMAX),
        ("z", C.c_double * SCLANE_PMAX),
        ("p", C.c_double * SCLANE_PMAX),
        ("cov", C.c_double * (SCLANE_PMAX * SCLANE_PMAX)),
        ("theta", C.c_double),
        ("se_theta", C.c_double),
        ("loglik", C.c_double),
        ("deviance", C.c_double),
        ("alternations", C.c_int32),
        ("irls_iters", C.c_int32),
        ("null_coef", C.c_double),
        ("null_se", C.c_double),
        ("null_z", C.c_double),
        ("null_p", C.c_double),
        ("null_theta", C.c_double),
        ("null_se_theta", C.c_double),
        ("null_loglik", C.c_double),
        ("null_deviance", C.c_double),
        ("null_alternations", C.c_int32),
        ("null_irls_iters", C.c_int32),
        ("test_stat", C.c_double),
        ("df", C.c_double),
        ("p_val", C.c_double),
        ("fwd_n_terms", C.c_int32),
        ("fwd_kind", C.c_int32 * SCLANE_PMAX),
        ("fwd_knot_idx", C.c_int32 * SCLANE_PMAX),
        ("fwd_score", C.c_double * 4),
        ("fwd_sigma", C.c_double * 4),
        ("wic", C.c_double * SCLANE_PMAX),
        ("n_wic", C.c_int32),
        ("fit_passes", C.c_int32),
        ("gene_time", C.c_double),
        ("gee_alpha", C.c_double),
        ("gee_phi", C.c_double),
        ("null_gee_alpha", C.c_double),
        ("null_gee_phi", C.c_double),
        ("gee_converged", C.c_int32),
        ("gee_iters", C.c_int32),
        ("null_gee_converged", C.c_int32),
        ("null_gee_iters", C.c_int32),
        ("stage_cycles", C.c_double * 3),
    ]


class SclaneLineageInfo(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64),
        ("n_knots", C.c_int32),
        ("minspan", C.c_int32),
        ("knots", C.c_double * SCLANE_TMAX),
        ("pt_min", C.c_double),
        ("pt_max", C.c_double),
    ]


# every symbol include/sclane_b200.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = [
    "sclane_abi_version", "sclane_device_count", "sclane_default_opts", "sclane_record_size",
    "sclane_launch_count", "sclane_reset_launch_count", "sclane_get_profile", "sclane_release",
    "sclane_test_dynamic", "sclane_test_dynamic_device", "sclane_test_dynamic_gee", "sclane_test_dynamic_csc", "sclane_marge_basis", "sclane_candidate_knots", "sclane_lineage_plan",
    "sclane_null_stats_glm", "sclane_score_glm", "sclane_link_preds", "sclane_smoothed_counts",
    "sclane_matmul", "sclane_llt_inverse", "sclane_pinv",
]


def default_opts() -> SclaneOpts:
    """testDynamic.R:60-75 / marge2.R:56-72 defaults."""
    o = SclaneOpts()
    o.abi_version = SCLANE_ABI_VERSION
    o.mode = 0
    o.M = 5
    o.approx_knot = 1
    o.n_knot_max = 25
    o.minspan = -1
    o.tols_score = 1e-5
    o.n_gpus = 0
    o.genes_per_batch = 0
    o.verbose = 0
    o.gee_cor = COR_AR1
    return o
