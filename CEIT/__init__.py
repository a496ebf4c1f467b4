"""Reference module paths (CEIT.EFEM.EFEM, CEIT.EJAC.EJAC, CEIT.Solver.Solver, ...) re-exported from
ceit_b200, so the compute calls of the reference's Example_02/03/04 scripts run unchanged against this repository
(plotting - show_JAC, CEIT.EITPlotter - is out of scope)."""
