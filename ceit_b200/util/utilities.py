"""config.json access and cache helpers (host-side mirror of CEIT/util/utilities.py:12-75).

`get_config()` keeps the reference's semantics: the file is `config.json` two directories above
this module (CEIT/util/utilities.py:20-23), it is re-read on every call, mm -> SI conversion of
`electrode_centers`, `electrode_radius`, `detection_bound` (utilities.py:28-35), `rootdir` injected
(utilities.py:27).  One addition: the environment variable CEIT_B200_CONFIG may point at another
config.json (its directory then becomes `rootdir`) so that tests and benchmarks can run several
configurations from one checkout.
"""
import os
import pickle
from json import loads

import numpy as np


def config_path():
    env = os.environ.get("CEIT_B200_CONFIG")
    if env:
        return os.path.abspath(env)
    path = os.path.dirname(os.path.realpath(__file__))
    path = os.path.dirname(os.path.dirname(path))
    return os.path.join(path, "config.json")


def get_config():
    """Get info in the config.json file and turn all data to SI units."""
    cfg_file = config_path()
    path = os.path.dirname(cfg_file)
    with open(cfg_file, "r", encoding="utf-8") as f:
        config = loads(f.read())
    param_unit = config["sensor_param_unit"]
    assert (isinstance(param_unit, str) and (param_unit == "mm" or param_unit == "SI")) or \
        isinstance(param_unit, float), "Please enter the accurate unit."
    config["rootdir"] = path
    if param_unit == "mm":
        config["electrode_centers"] = list(np.array(config["electrode_centers"]) / 1000)
        config["electrode_radius"] = config["electrode_radius"] / 1000
        config["detection_bound"] = config["detection_bound"] / 1000
    elif isinstance(param_unit, float):
        config["electrode_centers"] = list(np.array(config["electrode_centers"]) * param_unit)
        config["electrode_radius"] *= param_unit
        config["detection_bound"] *= param_unit
    return config


def insert_zero_to_delta_v(data, electrode_num):
    """utilities.py:39-45: put a 0 in front of every (L-1)-block."""
    data = np.asarray(data)
    out = np.zeros(len(data) + (len(data) + electrode_num - 2) // (electrode_num - 1), dtype=data.dtype)
    k = 0
    for i, v in enumerate(data):
        if i % (electrode_num - 1) == 0:
            out[k] = 0
            k += 1
        out[k] = v
        k += 1
    return out[:k]


def save_parameter(param, filename, path_name="."):
    """Pickle to `cache_<filename>.pkl` (utilities.py:48-59)."""
    with open(os.path.join(path_name, "cache_" + filename + ".pkl"), "wb") as file:
        pickle.dump(param, file)


def read_parameter(filename, path_name):
    """Read `cache_<filename>.pkl` (utilities.py:62-75)."""
    with open(os.path.join(path_name, "cache_" + filename + ".pkl"), "rb") as file:
        return pickle.load(file)


def _write_npy(path, arr):
    import io
    arr = np.ascontiguousarray(arr)
    hdr = io.BytesIO()
    np.lib.format.write_array_header_1_0(hdr, np.lib.format.header_data_from_array_1_0(arr))
    h = hdr.getvalue()
    try:
        if os.path.getsize(path) == len(h) + arr.nbytes:
            with open(path, "r+b") as f:
                f.write(h)
                f.write(arr.data)
            return
    except OSError:
        pass
    with open(path, "wb") as f:
        f.write(h)
        f.write(arr.data)


def save_npy(path, arr):
    """np.save-compatible writer (format 1.0, same bytes): when the file already exists with exactly the same
    size it is overwritten in place instead of being truncated and re-allocated (the Jacobian / inverse-model
    caches are rewritten with identical shapes on every run)."""
    wait_npy(path)
    _write_npy(path, arr)


# ---- write-behind for the cache files ---------------------------------------------------------------------
# JAC_cache.npy / inv_mat.npy are 8-10 MB memcpys into the page cache (~1 ms each at MESH_50_50, as long as the
# whole Jacobian on the GPU).  save_npy_async hands the write to a worker thread (file.write releases the GIL) so
# that it overlaps the next GPU stage; the caller must keep `arr` unchanged until wait_npy(path) - the classes
# join before they overwrite the pinned source buffer, before anything in this package reads the file back, and
# at interpreter exit.  OPT-IN (CEIT_B200_WRITE_BEHIND=1): by default every write is complete when the method
# returns, as in the reference - with write-behind a raw np.load of the cache by the caller, or an in-place edit of
# JAC_matrix, within the next millisecond would race the writer.
_pending = {}
_pool = None


def save_npy_async(path, arr):
    global _pool
    if os.environ.get("CEIT_B200_WRITE_BEHIND") != "1":
        return save_npy(path, arr)
    import atexit
    from concurrent.futures import ThreadPoolExecutor
    path = os.path.abspath(path)
    wait_npy(path)
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=2, thread_name_prefix="ceit-npy")
        atexit.register(wait_npy)
    _pending[path] = _pool.submit(_write_npy, path, arr)


def wait_npy(path=None):
    """Block until the pending write of `path` (default: all pending writes) is on disk; re-raises its error."""
    keys = list(_pending) if path is None else [os.path.abspath(path)]
    for k in keys:
        f = _pending.pop(k, None)
        if f is not None:
            f.result()


def load_npy(path):
    wait_npy(path)
    return np.load(path)
