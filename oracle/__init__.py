"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DiGS hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker (or as the CPU arm being timed), never as the thing shipped.
The product package ``digs_b200`` never imports this package and fails loudly when its
CUDA library is missing.

Contents
--------
jet_spec.py        closed-form forward jet (value, gradient, Laplacian) and its adjoint,
                   numpy, any float dtype (fp64 = ground truth).  Restates
                   models/DiGS.py:122-157 + utils/utils.py:108-117 + models/losses.py:52-188.
ref_autograd.py    the reference's *mechanism* restated in torch: MLP forward, then
                   autograd.grad(create_graph=True) per axis, loss, .backward().  This is
                   the CPU "port" timed by bench.py.
reference_loader.py imports the real /root/reference modules (with 5 stub modules) when
                   that directory exists -- only in the build container, never on the GPU box.
make_golden.py     runs the real reference and writes tests/golden/*.npz.

Parity pinning: the reference has no tests / golden vectors, so the
oracle is pinned against *outputs of the reference itself run in the build container*
(tests/golden/*.npz, generator committed).  tests/test_oracle_cpu.py checks both oracle
forms against those fixtures on every CPU run.
"""
