"""CPU oracle for the MendelIHT.jl IHT hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy / C restatement of the
reference algorithm (OpenMendel/MendelIHT.jl v1.4.11, Julia) and of the SnpArrays.jl
`SnpLinAlg` genotype operator it calls.  It exists so the CUDA path can be checked
against the reference's arithmetic on the same inputs.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import or execute anything under `oracle/`; the product
(`mendeliht.jl_b200/`) never does, and fails loudly if its CUDA library is missing.

Pinning status
--------------
* Julia is not installed in this image and SnpArrays.jl is not vendored in
  /root/reference, so the reference itself cannot be executed here.
* Univariate Normal + covariates is PINNED to the reference's own published
  iteration trace (`docs/src/man/examples.md:230-268`): 5 log-likelihoods, 5 tol
  values, support, beta, c and PVE (see tests/test_oracle_golden.py).
* Bernoulli / Poisson / NegativeBinomial / MvNormal fits and all CV MSEs have no
  known-answer vector in the reference tree (its tests assert properties only,
  `test/cv_iht_test.jl:1-3`): for those the oracle is a line-by-line restatement
  and parity is "unpinned" beyond the shared code paths exercised by the Normal trace
  and the per-function closed-form checks the reference's unit tests use
  (`test/utilities_test.jl:20-92`).
"""
