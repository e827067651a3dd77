"""CPU oracle for the KNP-EMI per-timestep hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
The product (``knp-emi-cgx_b200``) never imports anything from here and fails
loudly when its CUDA library is missing.

The oracle is a numpy/scipy restatement of the arithmetic that the reference
(hherlyng/knp-emi-cgx, "CGx") delegates to FFCx / DOLFINx / multiphenicsx /
PETSc / MUMPS / hypre.  None of those third-party packages is vendored in the
reference tree and none is installable here (pins: fenics-dolfinx 0.9.0,
ffcx 0.9.0, basix 0.9.0, ufl 2024.2.0, multiphenicsx v0.3.9, petsc 3.23.5,
hypre 2.32.0, mumps 5.8.1 -- see /root/reference/conda-lock.yml), so the
restatement follows the reference's own call sites and forms:

* forms ``a``/``L``:  src/CGx/KNPEMI/KNPEMIx_problem.py:454-655
* preconditioner ``P``: src/CGx/KNPEMI/KNPEMIx_problem.py:657-744
* ionic models:        src/CGx/KNPEMI/KNPEMIx_ionic_model.py
* time loop / solve:   src/CGx/KNPEMI/KNPEMIx_solver.py:104-116,297-335,365-468

Parity pinning: the only results the reference's tests hold for this path are
four end-to-end L2 norms (tests/KNPEMI/electric_potential_norms_*_solver.py).
``tests/test_oracle_golden.py`` checks the oracle against them.  There are no
golden matrices / vectors / CSR patterns in the reference, so entry-level parity
("CSR structure", "entries within 1e-12") is *pinned only through those norms*:
with respect to the real DOLFINx numbering and per-entry values it is PARITY UNPINNED
(DESIGN.md section 3).  ``oracle/amg.py`` restates the product's OWN preconditioners
(SA-AMG cycle, charge-conservation Schur form); hypre is not restated.

``oracle/p2.py`` (``fem_order: 2``): PARITY UNPINNED against the reference -- no config, test or golden vector of the
reference uses order 2.  It is pinned against the pinned P1 restatement instead: the P2 forms must equal the P1 forms on
the P1 subspace (tests/test_oracle_p2.py), plus exact integrals of the P2 basis and convergence towards a fine P1 solution.
"""
