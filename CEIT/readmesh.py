from ceit_b200.readmesh import init_mesh, read_mesh_from_csv, ReadMesh, string_to_float  # noqa: F401
