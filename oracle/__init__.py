"""CPU oracle for the NLP-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or as the timed CPU baseline.
The product path (``nlotrajectories_b200``) never imports this package and has no
CPU fallback: it raises if the CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"):
* ``sdf_oracle``  - pinned against the reference's own TorchScript artefacts
  ``_l4c_generated/{nn_sdf,jac_nn_sdf,adj1_nn_sdf,jac_adj1_nn_sdf}.pt`` executed in the
  build container (``oracle/make_golden.py`` -> ``tests/golden/sdf_shipped_fourier128.npz``).
* ``nlp_oracle``  - pinned against the reference's own ``core/runner.py``,
  ``core/dynamics.py``, ``core/geometry.py``, ``core/utils.py`` and ``core/sdf/*.py``
  executed unmodified under the sympy-backed ``oracle/casadi_stub`` (real CasADi is not
  installable here) (``oracle/make_golden.py`` -> ``tests/golden/nlp_*.npz``).
The reference's own test-suite holds no golden vectors for this path.
"""
