from ceit_b200.models.mesh import MeshObj  # noqa: F401
