"""CPU oracle for the eigen-pinns hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or the CPU baseline being timed.
The product package (``eigen-pinns_b200/``) never imports it and has no CPU
fallback.

Contents
--------
* ``step_port.py``      torch-CPU fp32 restatement of the training step
                        (reference src/multigrid_model.py:226-348, src/corrector_model.py,
                        src/utils.py:14-20)
* ``samplers_port.py``  numpy fp64 restatement of FPS / voxel down-sampling
                        (reference src/samplers.py:9-143)
* ``reference_loader.py`` imports the UNMODIFIED reference from /root/reference/src with
                        four empty stub modules (only works in the build container)
* ``make_golden.py``    runs the real reference and writes tests/golden/*.npz

Parity status: PINNED.  Every function in the two ports is checked in
``tests/test_oracle_golden.py`` against fixtures produced by the reference's own
code (``make_golden.py``), and against the reference's printed known answers
(bunny FEM eigenvalues, voxel level sizes 256/512/1003/2503).  The only part of the
reference pipeline that stays "parity unpinned" is the point-cloud Laplacian from the
third-party ``robust_laplacian`` package (not vendored, no pinned version, absent
here): all parity work uses the FEM operators of reference src/Mesh.py instead.
"""
