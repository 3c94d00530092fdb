"""CPU oracle for the scLANE testDynamic() per-gene fitting path.

TEST INFRASTRUCTURE ONLY.  This package is a dense fp64 numpy/scipy restatement of the
reference algorithm (jr-leary7/scLANE v0.99.2; citations are file:line relative to the
reference checkout).  It may be imported ONLY by tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py -- never by the product package
`sclane_b200`, which must fail loudly when its CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):
  * the reference's own tests pin no numbers (tests/testthat/test_scLANE.R asserts shapes only);
  * the oracle IS pinned against the reference's bundled golden run
    data/scLANE_models.rda (exported to tests/golden/ by tools/make_golden.py):
    final NB fits (coef/SE/logLik/deviance/LRT) reproduce the golden to <=1e-8 relative for
    every gene whose selected basis agrees; the discrete basis choice agrees for the
    count reported by tests/test_oracle_golden.py, the rest being the platform-dependent
    rank-check cases documented in DESIGN.md;
  * third-party arithmetic that is not in the reference tree (gamlss NBI fit, MASS::glm.nb,
    glm2::glm.fit2, RcppEigen::fastLmPure, Eigen LLT/JacobiSVD) is restated from the
    published algorithms; each restatement says so where it is defined.
"""
