from ceit_b200.Solver import Solver, reinitialize_solver  # noqa: F401
