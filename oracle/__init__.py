"""CPU oracle for the PTtools hot path (fluid-shell solve + Sound Shell Model GW spectrum).

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy/scipy restatement of the reference's algorithm for the
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import it; the product package ``pttools_b200`` never does and fails loudly if its CUDA library is missing.

The arithmetic of ODE stepping and root finding in the reference lives in un-vendored third-party code
(``scipy.integrate.odeint`` = ODEPACK LSODA, ``scipy.optimize.fsolve`` = MINPACK hybrd, ``root_scalar`` secant;
pinned ``scipy>=1.15.1`` in the reference's requirements.txt:27, scipy 1.18.1 in this image).  The oracle calls the
same scipy entry points at the reference's call sites (pttools/bubble/integrate.py:180, alpha.py:325,
speedup/solvers.py:25, fluid.py:619) so that index-based event decisions land on the same tau-grid points.

Parity is PINNED: tests/test_oracle_golden.py checks the oracle against the reference's own golden vectors
(re-derived into tests/golden/ by tools/make_golden.py, which imports the reference in the build container)
and against fixtures produced by running the reference itself on identical inputs.
"""
