"""CPU oracle for the port-Hamiltonian wave hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in numpy/scipy, the algorithm of the
reference's ``fenics_projects/wave_eqn/wave_2D.py::wave_2D_solve`` (assembled sparse
operators + sparse-direct solves, the way dolfin/PETSc executes the reference's UFL
forms).  Nothing in the product path (``porthamiltonian_fem_b200``) may import it; the
only legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: the real reference (dolfin 2019.1 + mshr + PETSc) cannot be imported in
this image, so the oracle is pinned against the reference's committed golden
``H_array.npy`` traces: exactly (13 digits) on the mesh-independent entries, and at
discretisation accuracy on the rest (mshr/CGAL meshes are not reproducible here).
See tests/test_oracle_golden.py and DESIGN.md section "Oracle".
"""
