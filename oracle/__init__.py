"""CPU oracle for the RCHQ batch-selection hot path of ma921/SOBER.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sober_b200/`` may import this package; the only
legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``, and there only as the checker / the timed CPU baseline,
never as the product path.

What it is
----------
A torch(CPU)/float64 restatement of the reference algorithm

* ``SOBER/_rchq.py``      (recombination, Nystrom basis, grouped barycentres, CAR elimination)
* ``SOBER/_utils.py``     (``is_psd`` / ``make_cov_psd`` PSD gate, lines 117-157)
* ``SOBER/_kernel.py``    (three ``Kernel`` modes, lines 16-47)
* ``SOBER/_gp.py``        (``predictive_covariance``, lines 281-295)
* ``SOBER/_drug_modelling.py`` (``batch_tanimoto_sim`` + clamp, lines 15-25, 36-38)
* gpytorch ``ScaleKernel`` / ``RBFKernel`` / ``MaternKernel`` arithmetic.  gpytorch is an
  un-vendored third-party dependency (``requirements.txt:2`` pins ``gpytorch==1.10``,
  ``pyproject.toml:27`` says ``>=1.11``) and is absent from ``/root/reference`` and from this
  image, so its published algorithm is restated in ``oracle/kernels.py``.

Parity status
-------------
The reference ships **no tests, golden vectors or known-answer fixtures** for this path
(SURVEY.md section 4).  The oracle is therefore pinned the second way the task allows: against
outputs of the reference itself.  ``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/SOBER/_rchq.py`` (+ ``_utils.py``, ``_settings.py``) by file path, runs it on
seeded synthetic inputs and stores stage-wise outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those fixtures bit-for-bit (and,
when ``/root/reference`` is present, against a live run of the reference).  The gpytorch kernel
arithmetic itself has no fixture (library absent): that part is "parity unpinned" and says so in
``oracle/kernels.py`` and DESIGN.md.
"""
