"""CPU oracle for the fbs CSMC / particle-Gibbs / pMCMC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``fbs_b200``) may import,
call, link or execute anything under ``oracle/``.  The only legitimate callers are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / CPU baseline.

What it is
----------
A NumPy (plus a small C port under ``oracle/csrc``) restatement of the reference's
algorithm for the path named in BASELINE.json:north_star, following, file by file,

* ``fbs/samplers/csmc/csmc.py``, ``fbs/samplers/csmc/resamplings.py``
* ``fbs/samplers/resampling.py``, ``fbs/samplers/gibbs.py``, ``fbs/samplers/smc.py``
* ``fbs/sdes/linear.py`` (:9-227, :397-457), ``fbs/sdes/simulators.py`` (:53-160)
* the closures of ``experiments/toy/gp_gibbs.py``, ``experiments/toy/gp_pmcmc.py``,
  ``experiments/sb/gibbs.py``

Part of the arithmetic lives in a third-party dependency that is NOT under
/root/reference: ``jax==0.4.26`` / ``jaxlib==0.4.26+cuda12.cudnn89``
(requirements_freeze.txt).  ``oracle/jax_random.py`` restates its published threefry2x32
PRNG (non-partitionable default of that version), ``split`` / ``uniform`` / ``normal`` /
``choice`` / ``randint`` and XLA's fp32 ``erf_inv`` polynomial.

PARITY PIN STATUS: **partially pinned**.
* bit level: threefry2x32-20 against the three Random123 known-answer vectors, and
  ``split(PRNGKey(0))``, ``uniform(PRNGKey(0))``, ``normal(PRNGKey(0), (1,))`` against
  the values printed in JAX's public documentation (tests/test_oracle_random.py);
* closed forms: the exact discretisations asserted in the reference's tests/test_sdes.py;
* distribution level: the reference's own acceptance criteria
  (tests/test_cond_resamplings.py, test_gibbs.py, test_pmcmc.py, test_filters.py,
  test_csmc.py) re-run against this restatement.
* everything else (``randint``/``choice`` streams, XLA summation order, whole sweeps) is
  **parity unpinned**: JAX cannot be imported in this image, the reference holds no golden
  vectors for the path, so no bit-level answer from the real reference exists.

Conventions fixed by this oracle where XLA leaves them open
-----------------------------------------------------------
* ``cumsum`` and the sums inside the resampling functions are *sequential* float32
  (``c[i] = fl(c[i-1] + w[i])``), which is what a non-reassociated XLA:CPU loop computes.
* float32 everywhere unless a function is handed float64 arrays (``jax_enable_x64`` off in
  the experiments, ``experiments/toy/gp_gibbs.py:28``).
"""
