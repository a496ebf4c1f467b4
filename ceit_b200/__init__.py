"""ceit_b200 - B200-native implementation of CEIT's Jacobian / Gauss-Newton hot path.

Drop-in mirrors of the reference's classes (same names, arguments, files and config keys):
    ceit_b200.EFEM.EFEM, ceit_b200.EJAC.EJAC, ceit_b200.Solver.Solver / reinitialize_solver,
    ceit_b200.models.mesh.MeshObj, ceit_b200.readmesh.init_mesh, ceit_b200.util.utilities.get_config
The top-level `CEIT` package in this repository re-exports them under the reference's module paths.
All arithmetic runs in libceit_b200.so (CUDA, sm_100a) through ctypes; there is no CPU fallback.
"""
__version__ = "0.1.0"
