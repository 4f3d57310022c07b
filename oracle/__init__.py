"""CPU oracle for the sparse-variational-GP node data sweep.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import, call, link or execute it,
and there only as the checker / reported baseline.  The product (``gaussianprocessnode_b200`` +
``libsgp.so``) never routes through this package and has no CPU fallback.

What it is: a NumPy float64 restatement, function by function, of the reference's arithmetic for the
hot path (biaslab/GaussianProcessNode: ``GPnode/UniSGPnode.jl``, ``GPnode/MultiSGPnode.jl``,
``helper_functions/gp_helperfunction.jl``, ``helper_functions/ut_approx.jl``) in two forms:

* the *per-point* rule form exactly as the reference schedules it (one message per data point, folded by
  ``prod``, flush on the N-th) -- ``oracle.unisgp`` / ``oracle.multisgp``;
* the *batched* identities those per-point sums collapse to (Psi0/Psi1/Psi2 -> posterior -> sum I1 / I2
  -> energy) -- ``oracle.batched`` -- which is what the CUDA path computes.

Pinning status.  The reference is Julia and cannot be run in the build container
(no ``julia``; its dependencies are unvendored and unpinned: no Manifest.toml, no [compat]).  The oracle is
pinned by the reference's own saved artefacts instead:

* SE-ARD kernel convention + predictive mean: the kin40k golden chain (``savefiles/qv_kin40k.jld``,
  ``Xu_kin40k.jld``, ``params_optimal_kin40k.jld``, ``data/kin40k``) reproduces the notebook's printed
  SMSE 0.08343114079545057 to ~1e-14 (``tests/test_oracle_golden.py``); the banana chain reproduces exactly
  125 / 1300 test errors.
* rule algebra: ``GPtest.jl``'s closed-form ground truths re-expressed with seeded inputs
  (``tests/test_oracle_rules.py``).
* Third-party pieces whose source is NOT in /root/reference and whose versions are unpinned --
  ReactiveMP's ``srcubature`` / ``ghcubature`` node placement, FastCholesky's ``cholinv`` on non-symmetric
  input (only met inside multivariate GenUnscented), ReactiveMP's probit moment matching --
  are restated from their published definitions: **parity unpinned** for those (the reference's own tests
  only pin them to Monte-Carlo tolerance, GPtest.jl:127-143, 376-382).
"""
