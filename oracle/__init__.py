"""TEST INFRASTRUCTURE ONLY — CPU oracle for the GP-predict + acquisition hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
or as the timed CPU baseline -- never as a fallback for the CUDA path.

* ``oracle.gp_oracle``   -- float64 numpy/scipy restatement of the reference algorithm
  (``bayes_optim/surrogate/gaussian_process/{gpr,kernel,trend}.py`` and
  ``bayes_optim/acquisition/acquisition_fun.py``), each function citing the reference file:line.
* ``oracle.ref_loader``  -- imports the *real* reference from ``/root/reference`` when it is
  present (build container only; it does not exist on the GPU box) to pin the restatement and
  to generate ``tests/golden/*.npz``.

Parity pinning: the reference's own tests hold no numeric vectors for this path (SURVEY.md §4,
§8c).  The restatement is therefore pinned against outputs of the reference itself, run in the
build container by ``tests/golden/make_golden.py`` (committed) -> ``tests/golden/*.npz``.
"""
