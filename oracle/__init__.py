"""CPU oracle for the EgoNN descriptor-extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``egonn_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import or execute it, and there only as the checker / the timed CPU arm.

What it is: a plain torch-CPU + numpy restatement of the MinkowskiEngine 0.5.4 semantics that the
reference's forward path (``models/minkgl.py:267-315``) relies on.  MinkowskiEngine itself is a
third-party dependency of the reference (``README.md:46,51``: "MinkowskiEngine 0.5.4"), it is not
vendored under ``/root/reference`` and cannot be installed in this image, so the arithmetic is
restated from its published behaviour and anchored on the reference's own
call sites.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 and MinkowskiEngine cannot be run here.  The pinning that exists is ours:
  * dense ``torch.nn.functional.conv3d`` / ``conv_transpose3d`` cross-checks (tests/test_oracle_dense.py),
  * the UNMODIFIED reference model code (``/root/reference/models/minkgl.py`` etc.) executed on top of
    ``oracle/me_shim/MinkowskiEngine`` to generate ``tests/golden/*.npz`` (tests/golden/make_golden.py),
  * invariance properties (permutation, batch independence, translation by multiples of 128).
"""
