"""TEST / BENCH INFRASTRUCTURE: run the UNMODIFIED reference (zehao99/CEIT) from the git-ignored copy `oracle/_ref/`.

`python oracle/ref_runner.py --install [/root/reference]` (called by `__graft_entry__.build()` where the reference
tree exists, i.e. in the build container) copies the reference's `CEIT/` package, `config.json` and `MESH_50_50/`
deck into `oracle/_ref/` and runs the reference's own `init_mesh()` once so that its mesh cache exists.  Nothing under
`oracle/_ref/` is tracked by git (the sources are never part of this repository's history); the directory travels to
the GPU box with the gpurun snapshot, where `bench.py --impl reference` times the reference's own forward solve
(`EFEM.calculation`, the body of the double loop in CEIT/EJAC.py:123-133) on the host cores.

The reference needs three accommodations to run on this image, none of which touches its code:
matplotlib / progressbar are not installed (stub modules), numpy >= 1.24 dropped np.float / np.int (restored), and
the tree must be writable (it is a copy).
"""
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
FOLDER = "MESH_50_50"


def install_shims():
    import numpy as np
    np.float = float
    np.int = int
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.font_manager"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    mpl.patches = sys.modules["matplotlib.patches"]
    mpl.font_manager = sys.modules["matplotlib.font_manager"]
    pb = types.ModuleType("progressbar")
    pb.progressbar = lambda it, *a, **k: it
    sys.modules["progressbar"] = pb


def available():
    return (os.path.isfile(os.path.join(REF, "CEIT", "EFEM.py"))
            and os.path.isfile(os.path.join(REF, FOLDER, "cache_Mesh_nodes.pkl")))


def import_reference():
    """Import the reference package from oracle/_ref (shadows this repository's own top-level `CEIT` shim)."""
    install_shims()
    for k in [k for k in sys.modules if k == "CEIT" or k.startswith("CEIT.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import CEIT  # noqa: F401
    assert os.path.realpath(os.path.dirname(CEIT.__file__)) == os.path.realpath(os.path.join(REF, "CEIT"))
    return CEIT


def install(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "CEIT")):
        return False
    os.makedirs(REF, exist_ok=True)
    for item in ("CEIT", "config.json", FOLDER):
        s, t = os.path.join(src, item), os.path.join(REF, item)
        if os.path.isdir(s):
            if not os.path.isdir(t):
                shutil.copytree(s, t)
        elif not os.path.exists(t):
            shutil.copy(s, t)
    for root, _, files in os.walk(REF):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    cfg_path = os.path.join(REF, "config.json")
    with open(cfg_path) as f:
        cfg = json.load(f)
    if cfg.get("folder_name") != FOLDER or cfg.get("device") != "cpu":
        cfg["folder_name"], cfg["device"] = FOLDER, "cpu"
        with open(cfg_path, "w") as f:
            json.dump(cfg, f, indent=1)
    if not os.path.isfile(os.path.join(REF, FOLDER, "cache_Mesh_nodes.pkl")):
        import_reference()
        from CEIT.readmesh import init_mesh
        init_mesh(draw=False)                    # the reference's own parser + O(N^2) clean-up, ~6 s, once
    return available()


def forward_solver():
    """(EFEM instance of the reference on MESH_50_50, L, E)."""
    import_reference()
    from CEIT.EFEM import EFEM
    fwd = EFEM()
    fwd.reset_variable(overall_variable=0)
    return fwd, fwd.mesh.electrode_num, fwd.elem_num


def jacobian_columns(fwd, pairs, c):
    """The body of the reference's double loop (CEIT/EJAC.py:123-133) for the given (excitation, element) pairs:
    returns an array (len(pairs), L - 1) of |electrode potential| rows."""
    import numpy as np
    L = fwd.mesh.electrode_num
    out = np.zeros((len(pairs), L - 1))
    for k, (i, j) in enumerate(pairs):
        fwd.elem_variable[j] += c
        _, _, electrode_potential = fwd.calculation(i)
        cnt = 0
        for m in range(L):
            if m != i:
                out[k, cnt] = np.abs(electrode_potential[m])
                cnt += 1
        fwd.elem_variable[j] -= c
    return out


if __name__ == "__main__":
    if "--install" in sys.argv:
        rest = [a for a in sys.argv[1:] if not a.startswith("--")]
        print("reference copy:", "ok" if install(*(rest[:1])) else "unavailable")
    else:
        print("available" if available() else "unavailable")
