#!/usr/bin/env python
"""bench.py - the hot-path benchmark (contract: see DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path on MESH_50_50 (BASELINE.json configs[1]): assemble K,
factorise the base system, build the full 240 x 5000 Jacobian (80 000 columns = forward solves of
the reference) and the Gauss-Newton inverse model (lambda = 289).  Metric: Jacobian columns/s.
`value` times the step with inputs resident in HBM; `e2e` times the same work through the drop-in
classes with host (numpy) inputs and outputs.  The realtime metric (frames/s) is reported under
"realtime" in the same JSON line.  N > 1 shards the excitations over ranks (strong scaling of one
mesh), all-gathers J over NCCL and replicates the inverse model.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
LAMBDA = 289
C_PERT = 1e-3
FREQ = 20000.0
N_FRAMES = 100000
WORKLOAD = ("MESH_50_50 full Jacobian (16 x 5000 = 80000 columns) + Gauss-Newton inverse model "
            "(lambda=289, n_det=4050, m=240)")


def load_mesh(tag="50_50"):
    m = dict(np.load(os.path.join(GOLDEN, f"mesh_{tag}.npz")))
    ptr, ee = m["electrode_ptr"], m["electrode_elems"]
    m["electrodes"] = [ee[ptr[i]:ptr[i + 1]] for i in range(len(ptr) - 1)]
    return m


CONFIGS = {
    # BASELINE.json configs[1] (the N=1 headline), [2] and [3]
    "50_50": dict(kind="fixture", tag="50_50", workload=None),
    "200": dict(kind="synthetic", n=200, per_side=5,
                workload="synthetic structured 200x200-node mesh (79202 elements, pitch 2 mm), 16 electrodes: full "
                         "Jacobian (1267232 columns) + Gauss-Newton inverse model (lambda=289)"),
    "400": dict(kind="synthetic", n=400, per_side=9,
                workload="synthetic structured 400x400-node mesh (318402 elements, pitch 2 mm), 32 electrodes: full "
                         "Jacobian (10188864 columns) + Gauss-Newton inverse model (lambda=289)"),
}


def load_workload(name):
    """mesh arrays + electrode element lists + detection index, built with the product's own mesh model."""
    cfg = CONFIGS[name]
    if cfg["kind"] == "fixture":
        m = load_mesh(cfg["tag"])
        return dict(nodes=m["nodes"], elem=m["elem"], elem_perm=m["elem_perm"], elem_param=m["elem_param"],
                    electrodes=m["electrodes"], detection_index=m["detection_index"], workload=WORKLOAD)
    from ceit_b200 import meshgen
    from ceit_b200.models.mesh import MeshObj
    td = tempfile.mkdtemp(prefix="ceit_bench_")
    old = os.environ.get("CEIT_B200_CONFIG")
    os.environ["CEIT_B200_CONFIG"] = meshgen.make_synthetic_workspace(td, cfg["n"], cfg["per_side"])
    try:
        mesh = MeshObj()
    finally:
        if old is None:
            del os.environ["CEIT_B200_CONFIG"]
        else:
            os.environ["CEIT_B200_CONFIG"] = old
        import shutil
        shutil.rmtree(td, ignore_errors=True)
    return dict(nodes=mesh.nodes, elem=mesh.elem, elem_perm=mesh.elem_perm,
                electrodes=[mesh.electrode_mesh[i] for i in range(mesh.electrode_num)],
                detection_index=mesh.detection_index, workload=cfg["workload"])


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's literal restatement of the reference's CPU path
# ------------------------------------------------------------------------------------------------------
def cpu_baseline_columns(mesh, n_cols, seed=0):
    """Time `n_cols` forward solves of the literal restatement of EFEM.calculation (dense K assembled
    in a Python loop, row/col swaps, np.linalg.inv) = n_cols Jacobian columns of the reference."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ceit_oracle as o
    N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
    omega = 2 * np.pi * FREQ
    rng = np.random.default_rng(seed)
    var = np.zeros(E)
    t0 = time.perf_counter()
    for _ in range(n_cols):
        i, j = int(rng.integers(L)), int(rng.integers(E))
        var[j] += C_PERT
        o.forward_dense(mesh["elem"], mesh["elem_param"], mesh["elem_perm"], var, omega, mesh["electrodes"], i, N)
        var[j] -= C_PERT
    return n_cols / (time.perf_counter() - t0)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([d.get("num_threads", 1) for d in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:  # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread it can use
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    mesh = load_mesh()
    per_step = 2
    for _ in range(args.warmup):
        cpu_baseline_columns(mesh, 1)
    t0 = time.perf_counter()
    tot = 0
    for s in range(args.steps):
        cpu_baseline_columns(mesh, per_step, seed=s)
        tot += per_step
    dt = time.perf_counter() - t0
    val = tot / dt
    cores = blas_threads()
    sample = (f"{per_step} forward solves (= Jacobian columns) of MESH_50_50 per step, literal numpy restatement "
              "of EFEM.calculation (oracle/ceit_oracle.py:forward_dense); the inverse model (1.4 s once per 80 000 "
              "columns) is not in the sample")
    print(json.dumps({
        "impl": "reference", "metric": "jacobian_columns_per_s", "value": val, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "MESH_50_50 mesh fixture (tests/golden), deterministic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": "columns/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_mhz_max_seen"] = float(np.max(sm))
            out["sm_max_mhz"] = float(np.max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ceit_b200 import _lib, meshgen
    from ceit_b200.sharding import gather_jacobian, shard_ranges

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")           # keep stdout to the single JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    mesh = load_workload(args.config)
    workload = mesh["workload"]
    is_headline = args.config == "50_50"
    N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
    m_rows = L * (L - 1)
    omega = 2 * np.pi * FREQ
    n_cols = L * E

    plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"])
    perm = torch.tensor(mesh["elem_perm"], device=dev)
    var = torch.zeros(E, dtype=torch.float64, device=dev)
    det = torch.tensor(mesh["detection_index"], device=dev)
    n_det = det.numel()
    J = torch.zeros((m_rows, E), dtype=torch.float64, device=dev)
    base = torch.zeros(m_rows, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    i0, i1, e0, e1 = shard_ranges(L, E, ws, rank)
    # inverse model: detection elements sharded over ranks (each rank produces its rows of inv_mat)
    dlo, dhi = (n_det * rank) // ws, (n_det * (rank + 1)) // ws
    det_local = det[dlo:dhi].contiguous()
    inv_holder = {}

    def local_part():                      # this rank's library calls, no collective: capturable
        plan.assemble(perm, var, omega)
        plan.factor(False)
        plan.jacobian_block(i0, i1, e0, e1, C_PERT, omega, J, E, base)
        if ws == 1:
            inv_holder["inv"] = _lib.inverse_model(J, det, LAMBDA, shift=1.0, form=0)

    def exchange_part():                   # N > 1: all-gather of J, sharded inverse model (one all-reduce)
        if ws > 1:
            gather_jacobian(J, L, E)
            inv_holder["inv"] = _lib.inverse_model_sharded(J, det_local, LAMBDA, shift=1.0)

    def step():
        local_part()
        exchange_part()

    def barrier():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
            torch.cuda.synchronize()

    step()                                 # first call allocates the library workspaces (not capturable)
    barrier()
    # the ~130 short dependent kernels of a step are replayed from ONE CUDA graph (no launch gaps);
    # --no-graph launches them on the stream one by one
    use_graph = not args.no_graph
    run_step = step
    if use_graph:                          # the NCCL exchange of N > 1 stays outside the graph
        graph = _lib.StepGraph(local_part)

        def run_step():
            graph.replay()
            exchange_part()
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        run_step()
    barrier()
    if plan.numeric_flag() != 0:
        raise SystemExit("numeric failure in the factorisation")
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        run_step()
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # second pass over the same K steps with CUDA events around every stage inside the library (stream
    # launches of the same kernels: event records are not part of the captured graph): stage breakdown and
    # the duration of the roofline kernel
    _lib.set_option("stage_timers", 1)
    _lib.stage_times()
    for k in range(args.steps):
        flush.zero_()
        step()
    barrier()
    stages = _lib.stage_times()
    _lib.set_option("stage_timers", 0)
    if ws > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = n_cols * args.steps / (ms * 1e-3)

    # ---- realtime reconstruction: 100k batched frames resident in HBM, and single frames ---------------
    inv = inv_holder["inv"]
    if ws > 1:   # the rows of inv_mat live on their ranks; the realtime leg below runs on rank 0 with the full matrix
        inv = _lib.inverse_model(J, det, LAMBDA, shift=1.0, form=0)
    realtime = None
    if rank == 0 and is_headline:
        g = torch.Generator(device=dev).manual_seed(0)
        dV = torch.rand((m_rows, N_FRAMES), dtype=torch.float64, device=dev, generator=g)
        out = torch.empty((n_det, N_FRAMES), dtype=torch.float64, device=dev)
        for _ in range(3):
            _lib.reconstruct(inv, dV, out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            _lib.reconstruct(inv, dV, out)
        b.record()
        torch.cuda.synchronize()
        t_batch = a.elapsed_time(b) / reps * 1e-3
        one = dV[:, 0].contiguous()
        o1 = torch.empty(n_det, dtype=torch.float64, device=dev)
        for _ in range(10):
            _lib.reconstruct(inv, one, o1)
        torch.cuda.synchronize()
        a.record()
        for _ in range(1000):
            _lib.reconstruct(inv, one, o1)
        b.record()
        torch.cuda.synchronize()
        t_one = a.elapsed_time(b) / 1000 * 1e-3
        flops = 2.0 * n_det * m_rows * N_FRAMES
        # FP64 tensor denominator: cuBLAS DGEMM measured live (MEASURED_PEAKS.json has no f64 entry)
        A = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
        B = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
        for _ in range(2):
            A @ B
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            A @ B
        b.record()
        torch.cuda.synchronize()
        dgemm_tf = 5 * 2 * 4096.0 ** 3 / (a.elapsed_time(b) * 1e-3) / 1e12
        # the same batch end to end: pinned host frames -> H2D -> GEMM -> D2H of the density maps
        dV_h = torch.empty((m_rows, N_FRAMES), dtype=torch.float64, pin_memory=True)
        dV_h.copy_(dV)
        out_h = torch.empty((n_det, N_FRAMES), dtype=torch.float64, pin_memory=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dV.copy_(dV_h, non_blocking=True)
        _lib.reconstruct(inv, dV, out)
        out_h.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        del dV_h, out_h
        realtime = {
            "frames_per_s_batched_100k_e2e": N_FRAMES / t_e2e, "ms_batched_100k_e2e": t_e2e * 1e3,
            "e2e_h2d_bytes": int(m_rows * N_FRAMES * 8), "e2e_d2h_bytes": int(n_det * N_FRAMES * 8),
            "frames_per_s_batched_100k": N_FRAMES / t_batch, "ms_batched_100k": t_batch * 1e3,
            "batched_tflops_f64": flops / t_batch / 1e12, "cublas_dgemm_4096_tflops": dgemm_tf,
            "batched_frac_of_cublas_dgemm": flops / t_batch / 1e12 / dgemm_tf,
            "frames_per_s_single": 1.0 / t_one, "us_single_frame": t_one * 1e6,
        }
        del dV, out, A, B

    # ---- e2e: the drop-in classes, host (numpy) in / out, cache files written like the reference -----
    e2e = None
    if rank == 0 and is_headline:
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=shm) as td:
            cfg = dict(meshgen.REFERENCE_CONFIG)
            cfg_path = meshgen.make_workspace(td, mesh["nodes"], mesh["elem"], cfg)
            os.environ["CEIT_B200_CONFIG"] = cfg_path
            from ceit_b200.EJAC import EJAC
            jac = EJAC()

            def e2e_step():
                jac.fwd_FEM.invalidate()                 # nothing cached across steps
                Jh = jac.JAC_calculation()               # H2D perm/var, K1-K3-K2, D2H J, np.save
                jac.save_inv_matrix(LAMBDA)              # H2D J, inverse model, D2H inv_mat, np.save
                return Jh

            for _ in range(3):
                e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e = {"value": n_cols * args.steps / dt, "unit": "columns/s",
                   "h2d_bytes_per_step": int(2 * E * 8 + m_rows * E * 8 + n_det * 8),
                   "d2h_bytes_per_step": int(m_rows * E * 8 + n_det * m_rows * 8),
                   "ms_per_step": 1e3 * dt / args.steps,
                   "api": "EJAC.JAC_calculation() + EJAC.save_inv_matrix(289) (writes JAC_cache.npy / inv_mat.npy)"}
            # informational: the same loop with the opt-in write-behind of the two cache files (worker thread,
            # joined inside the timed region); the headline above keeps the reference's synchronous writes
            from ceit_b200.util.utilities import wait_npy
            os.environ["CEIT_B200_WRITE_BEHIND"] = "1"
            for _ in range(3):
                e2e_step()
            wait_npy(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            wait_npy(); torch.cuda.synchronize()
            e2e["ms_per_step_write_behind"] = 1e3 * (time.perf_counter() - t0) / args.steps
            del os.environ["CEIT_B200_WRITE_BEHIND"]
            del os.environ["CEIT_B200_CONFIG"]
        if ws > 1:
            e2e["note"] = "single-rank API call on rank 0 (the class API is single-process)"

    # ---- roofline of the kernels (stage timers = CUDA events around the launches, on our stream) ------
    roofline, rooflines, cpu = None, None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"
        info = plan.info()
        nU = int(info[3])
        # K2 Woodbury: SURVEY.md §8d algorithmic bytes per column = 8 r + 96 + 8 (L-1), r = measurement nodes,
        # i.e. Yh/Xh rows read once per excitation (16 r N / E per column), the 3x3 block and the J write.
        r_nodes = nU - int(np.mean([len(np.unique(mesh["elem"][np.asarray(e)].reshape(-1))) for e in mesh["electrodes"]]))
        bytes_per_col = 8.0 * r_nodes + 96.0 + 8.0 * (L - 1)
        my_cols = (i1 - i0) * (e1 - e0)
        wb_ms, wb_n = stages["woodbury"]
        k2 = None
        if wb_n:
            per_launch_ms = wb_ms / wb_n
            cols_per_launch = my_cols * args.steps / wb_n
            ach = bytes_per_col * cols_per_launch / (per_launch_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")))["woodbury_lr_kernel"]
                if ws == 1 and is_headline:          # the capture is of the full 16 x 5000 launch
                    traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            except Exception:
                pass
            k2 = {"kernel": "woodbury_lr_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                  "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                  "ms_per_launch": per_launch_ms, "algorithmic_bytes_per_launch": bytes_per_col * cols_per_launch,
                  "timing": "CUDA events around the launch, second pass over the same K steps (stream launches; "
                            "the timed region replays the same kernels from a CUDA graph)",
                  "note": "algorithmic bytes = (8 r + 96 + 8 (L-1)) per column (SURVEY.md 8d); the stage timer wraps the "
                          "Woodbury kernel alone; at this size Yh is largely L2-resident and the kernel is bound by "
                          "shared-memory wavefronts (56 % of peak) and FP64 issue (33 %), see profiles/"}
        stage_ms = {k: v[0] / args.steps for k, v in stages.items() if v[1]}
        rooflines = {"stage_ms_per_step": stage_ms, "woodbury": k2}
        if stages.get("assemble", (0, 0))[1] and stages["assemble"][0] > 0:
            # K1: SURVEY.md 8d algorithmic bytes 16 N + 28 E + 16 nnz per launch
            a_bytes = 16.0 * N + 28.0 * E + 16.0 * int(info[7])
            a_ms = stages["assemble"][0] / stages["assemble"][1]
            rooflines["assemble"] = {
                "kernel": "assemble_csr_kernel", "bound": "hbm", "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "peak": hbm,
                "unit": "GB/s", "frac": a_bytes / (a_ms * 1e-3) / 1e9 / hbm, "traffic": None,
                "ms_per_launch": a_ms, "algorithmic_bytes_per_launch": a_bytes,
                "note": "launch-latency bound at this size (one ~9 us kernel over 0.5 MB); the byte count only "
                        "becomes meaningful on the 400x400 mesh (29 MB)"}
        if realtime:
            rooflines["reconstruct_batched"] = {
                "kernel": "gemm_dmma_kernel<0,0>", "bound": "tensor", "achieved": realtime["batched_tflops_f64"],
                "peak": realtime["cublas_dgemm_4096_tflops"], "unit": "TFLOP/s",
                "frac": realtime["batched_frac_of_cublas_dgemm"], "traffic": None,
                "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no f64 entry in MEASURED_PEAKS.json)"}
        roofline = k2
    if rank == 0 and is_headline:
        cpu_cols = 6
        v = cpu_baseline_columns(mesh, cpu_cols)
        cpu = {"value": v, "unit": "columns/s", "cores": blas_threads(), "kind": "port",
               "sample": f"{cpu_cols} forward solves of MESH_50_50 (literal numpy restatement of EFEM.calculation: "
                         "Python assembly loop + np.linalg.inv), = columns of the reference Jacobian"}
        # realtime CPU baseline: np.dot on a bounded batch
        inv_h = inv.cpu().numpy()
        dvh = np.random.default_rng(0).random((m_rows, 20000))
        t0 = time.perf_counter()
        np.dot(inv_h, dvh)
        cpu["realtime_frames_per_s_batched"] = 20000 / (time.perf_counter() - t0)

    if rank == 0:
        print(json.dumps({
            "metric": "jacobian_columns_per_s", "value": value, "unit": "columns/s", "n_gpus": ws,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": ("MESH_50_50 mesh fixture (tests/golden/mesh_50_50.npz), zero base variable; seeded synthetic frames"
                     if is_headline else "synthetic structured mesh (ceit_b200.meshgen), zero base variable"),
            "config": {"workload": workload,
                       "N": N, "E": E, "L": L, "columns_per_step": n_cols, "l2": "flushed between timed steps (256 MiB memset)",
                       "sharding": f"excitation blocks over {ws} rank(s), all-gather J (NCCL), inverse model sharded "
                                   f"over detection elements (one all-reduce of the m x m Gram matrix)"},
            "gpu_launches": int(launches), "launch_mode": ("cuda_graph" if ws == 1 else "cuda_graph(local part) + NCCL exchange") if use_graph else "stream",
            "wall_s_timed_region": wall,
            "clocks": clocks, "e2e": e2e, "roofline": roofline, "rooflines": rooflines, "cpu_baseline": cpu,
            "realtime": realtime,
        }))
    plan.close()
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--config", default="50_50", choices=sorted(CONFIGS), help="workload (default: the headline MESH_50_50)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
