"""CPU oracle for the AIMNet-X2D hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference algorithm
(``/root/reference/src/models/{layers,pooling,gnn}.py``,
``src/datasets/{molecular,features}.py``) plus the published semantics of the
absent third-party wheel ``torch_scatter==2.1.2`` (``requirements.txt:6``).

It exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  The product package ``aimnet_x2d_b200`` never imports ``oracle`` and has
no CPU fallback: it raises if ``libax2d.so`` is missing.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the UNMODIFIED reference modules run in
the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` replays those fixtures on every CPU test run.
"""
