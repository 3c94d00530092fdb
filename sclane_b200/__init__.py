"""sclane_b200 -- B200-native engine for scLANE's per-gene trajectory-DE fitting path.

The package is a thin host-side mirror of the reference's R interface (api.py) over the C ABI of
libsclane_b200.so (include/sclane_b200.h).  Importing it loads the CUDA library and raises if the
library is missing or stale: there is no CPU fallback."""
from ._lib import LIB_PATH, SclaneError, lib

lib()  # fail loudly at import time when the CUDA library has not been built

from .api import (  # noqa: E402
    ScLANE, candidate_knots, createCellOffset, device_count, eigenMapMatMult, eigenMapMatrixInvert,
    eigenMapPseudoInverse, get_profile, getFittedValues, glmmKnots, getResultsDE, launch_count, lineage_plan, link_preds, marge2, modelLRT,
    release, reset_launch_count, run_records, score_fun_glm_sweep, smoothedCountsMatrix, stat_out_score_glm_null, testDynamic,
    testSlope,
)

__all__ = [
    "LIB_PATH", "SclaneError", "ScLANE", "candidate_knots", "createCellOffset", "device_count", "eigenMapMatMult",
    "eigenMapMatrixInvert", "eigenMapPseudoInverse", "get_profile", "getFittedValues", "glmmKnots", "getResultsDE", "launch_count", "lineage_plan", "link_preds", "marge2",
    "modelLRT", "release", "reset_launch_count", "run_records", "score_fun_glm_sweep", "smoothedCountsMatrix", "stat_out_score_glm_null", "testDynamic", "testSlope",
]
