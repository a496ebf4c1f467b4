#!/usr/bin/env python
"""bench.py - the hot-path benchmark (contract: see DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 50_50|21_21|200|400]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one mesh: assemble K, factorise the base system, build the full Jacobian
(L x E columns = forward solves of the reference) and the Gauss-Newton inverse model (lambda = 289).  Metric: Jacobian
columns/s.  The headline line is MESH_50_50 (BASELINE.json configs[1]): `value` times the step with inputs resident
in HBM, `e2e` times the same work through the drop-in classes with host (numpy) inputs and outputs, `realtime` is the
frames/s leg (configs[4]).  The same line carries, under `scaling_configs`, the other BASELINE.json configs measured
in the same run at the same number of GPUs: MESH_21_21 (configs[0]), the synthetic 200x200 / 16-electrode mesh
(configs[2]) and the 400x400 / 32-electrode mesh (configs[3]).  N > 1 shards the excitations over ranks (strong
scaling of one mesh): all-to-all of the detection-column slices + one all-reduce of the m x m Gram matrix on the
critical path, J replicated by an in-place all-gather that overlaps the inverse-model GEMMs; after the timed region
every rank re-computes the whole step alone and the gathered results are checked against it.
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
    # BASELINE.json configs[1] (the headline), [0], [2] and [3]
    "50_50": dict(kind="fixture", tag="50_50", workload=WORKLOAD, steps=None),
    "21_21": dict(kind="fixture", tag="21_21", steps=50,
                  workload="MESH_21_21 full Jacobian (16 x 882 = 14112 columns) + Gauss-Newton inverse model "
                           "(lambda=289, n_det=722, m=240)"),
    "200": dict(kind="synthetic", n=200, per_side=5, steps=20,
                workload="synthetic structured 200x200-node mesh (79202 elements, pitch 2 mm), 16 electrodes: full "
                         "Jacobian (1267232 columns) + Gauss-Newton inverse model (lambda=289)"),
    "400": dict(kind="synthetic", n=400, per_side=9, steps=5,
                workload="synthetic structured 400x400-node mesh (318402 elements, pitch 2 mm), 32 electrodes: full "
                         "Jacobian (10188864 columns) + Gauss-Newton inverse model (lambda=289)"),
}


def load_workload(name):
    """mesh arrays + electrode element lists + detection index, built with the product's own mesh model;
    `host_setup_s` = seconds of host work to get there (mesh model; the plan is timed by the caller)."""
    cfg = CONFIGS[name]
    t0 = time.perf_counter()
    if cfg["kind"] == "fixture":
        m = load_mesh(cfg["tag"])
        return dict(nodes=m["nodes"], elem=m["elem"], elem_perm=m["elem_perm"], elem_param=m["elem_param"],
                    electrodes=m["electrodes"], detection_index=m["detection_index"], workload=cfg["workload"],
                    mesh_model_s=time.perf_counter() - t0)
    from ceit_b200 import meshgen
    from ceit_b200.models.mesh import MeshObj
    td = tempfile.mkdtemp(prefix="ceit_bench_")
    old = os.environ.get("CEIT_B200_CONFIG")
    os.environ["CEIT_B200_CONFIG"] = meshgen.make_synthetic_workspace(td, cfg["n"], cfg["per_side"])
    try:
        t0 = time.perf_counter()
        mesh = MeshObj()
        dt = time.perf_counter() - t0
    finally:
        if old is None:
            del os.environ["CEIT_B200_CONFIG"]
        else:
            os.environ["CEIT_B200_CONFIG"] = old
        import shutil
        shutil.rmtree(td, ignore_errors=True)
    return dict(nodes=mesh.nodes, elem=mesh.elem, elem_perm=mesh.elem_perm,
                electrodes=[mesh.electrode_mesh[i] for i in range(mesh.electrode_num)],
                detection_index=mesh.detection_index, workload=cfg["workload"], mesh_model_s=dt)


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU forward solve (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------------------
class CpuArm:
    """Times Jacobian columns of MESH_50_50 on the host cores.  kind = "reference": the UNMODIFIED reference package
    from the git-ignored copy oracle/_ref (its own EFEM.calculation inside the loop body of CEIT/EJAC.py:123-133);
    kind = "port": the oracle's literal restatement (oracle/ceit_oracle.py:forward_dense) when the copy is absent."""

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_runner
        self.rr = ref_runner
        self.mesh = load_mesh()
        self.L, self.E = len(self.mesh["electrodes"]), len(self.mesh["elem"])
        if ref_runner.available():
            self.kind = "reference"
            self.fwd, L, E = ref_runner.forward_solver()
            assert (L, E) == (self.L, self.E)
            self.what = ("the unmodified reference package (oracle/_ref, git-ignored copy): EFEM.calculation inside "
                         "the loop body of CEIT/EJAC.py:123-133")
        else:
            self.kind = "port"
            import ceit_oracle
            self.o = ceit_oracle
            self.what = ("literal numpy restatement of EFEM.calculation (oracle/ceit_oracle.py:forward_dense); the "
                         "reference copy oracle/_ref is absent")

    def columns_per_s(self, n_cols, seed=0):
        rng = np.random.default_rng(seed)
        pairs = [(int(rng.integers(self.L)), int(rng.integers(self.E))) for _ in range(n_cols)]
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.rr.jacobian_columns(self.fwd, pairs, C_PERT)
        else:
            m, N = self.mesh, len(self.mesh["nodes"])
            var = np.zeros(self.E)
            for i, j in pairs:
                var[j] += C_PERT
                self.o.forward_dense(m["elem"], m["elem_param"], m["elem_perm"], var, 2 * np.pi * FREQ,
                                     m["electrodes"], i, N)
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
    arm = CpuArm()
    per_step = 2
    for _ in range(min(args.warmup, 2)):
        arm.columns_per_s(1)
    steps = min(args.steps, 20)               # bounded: ~1.2 s per column on 8-16 cores
    t0 = time.perf_counter()
    tot = 0
    for s in range(steps):
        arm.columns_per_s(per_step, seed=s)
        tot += per_step
    dt = time.perf_counter() - t0
    val = tot / dt
    cores = blas_threads()
    sample = (f"{per_step} forward solves (= Jacobian columns) of MESH_50_50 per step x {steps} steps, {arm.what}; the "
              "inverse model (1.4 s once per 80 000 columns) is not in the sample")
    print(json.dumps({
        "impl": "reference", "metric": "jacobian_columns_per_s", "value": val, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "MESH_50_50 (the reference's own mesh cache when oracle/_ref is present, else tests/golden), deterministic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": "columns/s", "cores": cores, "kind": arm.kind, "sample": sample},
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


class Ctx:
    """process-wide state of the GPU arm"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ws = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.ws > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
                os.environ.pop("NCCL_DEBUG")           # keep stdout to the single JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)     # > 126 MB L2
        self.dgemm_tf = None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.ws > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def cublas_dgemm_tflops(self):
        """FP64 tensor denominator: cuBLAS DGEMM 4096^3 measured live (MEASURED_PEAKS.json has no f64 entry)."""
        if self.dgemm_tf is None:
            torch = self.torch
            A = torch.randn((4096, 4096), dtype=torch.float64, device=self.dev)
            B = torch.randn((4096, 4096), dtype=torch.float64, device=self.dev)
            for _ in range(2):
                A @ B
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                A @ B
            b.record()
            torch.cuda.synchronize()
            self.dgemm_tf = 5 * 2 * 4096.0 ** 3 / (a.elapsed_time(b) * 1e-3) / 1e12
        return self.dgemm_tf


def _measure_domain(ctx, name, steps, warmup, use_graph, sample_clocks, mesh, plan, plan_s, dom, perm, var, det, J, base,
                    flag):
    """The timed loop of the domain-decomposed schedule (N > 1): two captured graphs per rank around the base-stage
    exchange, then the sharded inverse model with the replication of J overlapped on a second communicator."""
    torch, dist = ctx.torch, ctx.dist
    from ceit_b200 import _lib
    ws, rank, dev = ctx.ws, ctx.rank, ctx.dev
    N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
    m_rows, n_det = L * (L - 1), det.numel()
    omega = 2 * np.pi * FREQ
    n_cols = L * E
    holder = {}

    def part_a():
        plan.assemble(perm, var, omega)
        plan.factor_phase(1)

    def part_b():
        plan.factor_phase(2)
        plan.jacobian_block(0, L, 0, E, C_PERT, omega, J, E, base)

    def tail():
        dom.gather_jacobian(J, async_op=True)                 # replication of J: off the critical path
        holder["inv"] = dom.inverse_model(J, LAMBDA, shift=1.0, flag=flag)
        dom.finish()

    def step():
        part_a()
        dom.base_exchange()
        part_b()
        tail()

    step()
    ctx.barrier()
    run_step = step
    if use_graph:
        ga, gb = _lib.StepGraph(part_a), _lib.StepGraph(part_b)

        def run_step():
            ga.replay()
            dom.base_exchange()
            gb.replay()
            tail()
    for _ in range(max(warmup, 3)):
        ctx.flush.zero_()
        run_step()
    ctx.barrier()
    if plan.numeric_flag() != 0 or int(flag.item()) != 0:
        raise SystemExit(f"numeric failure in the factorisation ({name})")
    launches0 = _lib.launch_count()
    sampler = ClockSampler(ctx.local) if (rank == 0 and sample_clocks) else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    wall0 = time.perf_counter()
    for k in range(steps):
        ctx.flush.zero_()
        ev[k][0].record()
        run_step()
        ev[k][1].record()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    _lib.set_option("stage_timers", 1)
    _lib.stage_times()
    _lib.stage_flops()
    for k in range(steps):
        ctx.flush.zero_()
        step()
    ctx.barrier()
    stages = _lib.stage_times()
    flops = _lib.stage_flops()
    _lib.set_option("stage_timers", 0)
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    value = n_cols * steps / (ms * 1e-3)
    info = plan.info()
    xinfo = plan.exchange_info()[0]
    # ---- the decomposed result against an undecomposed single-rank recompute on every rank (outside the timed region)
    J_multi = J.clone()
    inv_rows = holder["inv"].clone()
    ref_plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"])
    J.zero_()
    ref_plan.assemble(perm, var, omega)
    ref_plan.factor(False)
    ref_plan.jacobian_block(0, L, 0, E, C_PERT, omega, J, E, base)
    inv_full = _lib.inverse_model(J, det, LAMBDA, shift=1.0, form=2)
    inv_same = _lib.inverse_model(J_multi, det, LAMBDA, shift=1.0, form=2)
    ref_plan.close()
    dJ = float((J_multi - J).abs().max().item())
    sig = float((J - 1).abs().max().item())
    dsig = float((torch.linalg.norm(J_multi - J) / torch.linalg.norm(J - 1)).item())
    pos = dom.det_pos
    scale = float(inv_full.abs().max().item())
    dinv = float(((inv_rows - inv_same[pos]).abs().max() / scale).item()) if pos.numel() else 0.0
    dinv0 = float(((inv_rows - inv_full[pos]).abs().max() / scale).item()) if pos.numel() else 0.0
    agg = torch.tensor([dJ, dinv, dsig, dinv0], dtype=torch.float64, device=dev)
    dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    dJ, dinv, dsig, dinv0 = (float(v) for v in agg.tolist())
    check = {"max_abs_J_vs_single_rank": dJ, "max_abs_signal_J_minus_1": sig, "rel_frobenius_on_signal_J_minus_1": dsig,
             "max_rel_inv_rows_vs_model_of_same_J": dinv, "max_rel_inv_rows_vs_single_rank_model": dinv0,
             "tolerance": {"J": 1e-11, "signal": 1e-4, "inv_same_J": 1e-9, "inv_single_rank": 1e-4},
             "what": "every rank recomputes the whole Jacobian + inverse model alone on an UNDECOMPOSED plan after the "
                     "timed region; the gathered J (a different elimination order: rounding-level differences) and "
                     "this rank's rows of inv_mat are compared with it (max over ranks)"}
    if not (dJ <= 1e-11 and dsig <= 1e-4 and dinv <= 1e-9 and dinv0 <= 1e-4):
        raise SystemExit(f"multi-GPU (domain-decomposed) result differs from the single-rank result ({name}): {check}")
    stage_ms = {k: v[0] / steps for k, v in stages.items() if v[1]}
    own_cols = int(dom.own_idx.numel()) * L
    res = {
        "workload": mesh["workload"], "N": N, "E": E, "L": L, "nU": int(info[3]), "n_det": int(n_det), "m": m_rows,
        "columns_per_step": n_cols, "n_gpus": ws, "mode": "domain", "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms / steps, "value": value, "unit": "columns/s",
        "stage_ms_per_step": stage_ms, "gpu_launches_per_step": launches / steps, "nd_levels": int(info[23]),
        "host_setup_s": {"mesh_model": mesh["mesh_model_s"], "plan_create": plan_s,
                         "note": "once per mesh, outside the timed region"},
        "signal_J_minus_1_min": float((J - 1).min().item()),
        "multi_gpu_check": check,
        "domain": {"rows_on_rank0": int(xinfo[4]), "rows_total": N, "elements_on_rank0": int(dom.own_idx.numel()),
                   "cut_depth": int(xinfo[6]),
                   "base_exchange_bytes_per_rank": int(dom.exchange_bytes),
                   "jacobian_replication_bytes_per_rank": int(8 * m_rows * dom.e_max * (ws - 1))},
        "nvlink_bytes_per_step_per_rank": int(dom.exchange_bytes + 8 * m_rows * dom.e_max * (ws - 1) + 8 * m_rows * m_rows),
    }
    state = dict(mesh=mesh, plan=plan, J=J, det=det, inv=inv_full, stages=stages, flops=flops, steps=steps,
                 launches=launches, wall=wall, clocks=clocks, shard=(0, L, 0, E), info=info, own_cols=own_cols)
    return res, state


def measure_config(ctx, name, steps, warmup, use_graph, sample_clocks=False, mode="sharded"):
    """W warm-up + K timed steps of one workload at ctx.ws ranks.  Returns (result dict, state for the headline legs).
    mode (N > 1 only): "domain" = the mesh is decomposed (ceit_b200.sharding.DomainStep: each rank factorises its own
    sub-trees of the elimination forest, computes the Jacobian columns of its own elements and its rows of the inverse
    model; one small exchange inside the base stage, one m x m all-reduce, J replicated off the critical path);
    "sharded" = excitation blocks per rank with the base stage replicated + the NCCL exchange of J (round-1 schedule);
    "replicated" = every rank computes the whole step, no exchange (what a job does with a mesh too small to amortise
    the exchange latency)."""
    torch, dist = ctx.torch, ctx.dist
    from ceit_b200 import _lib
    from ceit_b200.sharding import ShardedStep, DomainStep, shard_ranges
    ws, rank, dev = ctx.ws, ctx.rank, ctx.dev
    if ws == 1:
        mode = "single"
    domain = mode == "domain"
    mesh = load_workload(name)
    N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
    m_rows = L * (L - 1)
    omega = 2 * np.pi * FREQ
    n_cols = L * E
    t0 = time.perf_counter()
    if domain:
        plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"], parts=ws, part=rank)
    else:
        plan = _lib.Plan(mesh["nodes"], mesh["elem"], mesh["electrodes"])
    plan_s = time.perf_counter() - t0
    perm = torch.tensor(mesh["elem_perm"], device=dev)
    var = torch.zeros(E, dtype=torch.float64, device=dev)
    det = torch.tensor(mesh["detection_index"], device=dev)
    n_det = det.numel()
    J = torch.zeros((m_rows, E), dtype=torch.float64, device=dev)
    base = torch.zeros(m_rows, dtype=torch.float64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    shard_it = ws > 1 and mode == "sharded"
    i0, i1, e0, e1 = shard_ranges(L, E, ws, rank) if shard_it else (0, L, 0, E)
    sharded = ShardedStep(L, E, det) if shard_it else None
    dom = DomainStep(mesh["nodes"], mesh["elem"], mesh["electrodes"], det, plan=plan) if domain else None
    holder = {}

    if domain:
        return _measure_domain(ctx, name, steps, warmup, use_graph, sample_clocks, mesh, plan, plan_s, dom, perm, var,
                               det, J, base, flag)

    def local_part():                      # this rank's library calls, no collective: capturable
        plan.assemble(perm, var, omega)
        plan.factor(False)
        plan.jacobian_block(i0, i1, e0, e1, C_PERT, omega, J, E, base)
        if not shard_it:
            holder["inv"] = _lib.inverse_model(J, det, LAMBDA, shift=1.0, form=0, flag=flag)

    def exchange_part():                   # N > 1: all-to-all + Gram all-reduce (critical), J replication (overlapped)
        if shard_it:
            Jr, cols = sharded.exchange(J)
            holder["inv"] = _lib.inverse_model_sharded(Jr, cols, LAMBDA, shift=1.0, flag=flag)
            sharded.finish()

    def step():
        local_part()
        exchange_part()

    step()                                 # first call allocates the library workspaces (not capturable)
    ctx.barrier()
    run_step = step
    if use_graph:                          # the NCCL exchange of N > 1 stays outside the graph
        graph = _lib.StepGraph(local_part)

        def run_step():
            graph.replay()
            exchange_part()
    for _ in range(max(warmup, 3)):
        ctx.flush.zero_()
        run_step()
    ctx.barrier()
    if plan.numeric_flag() != 0 or int(flag.item()) != 0:
        raise SystemExit(f"numeric failure in the factorisation ({name})")
    launches0 = _lib.launch_count()
    sampler = ClockSampler(ctx.local) if (rank == 0 and sample_clocks) else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    wall0 = time.perf_counter()
    for k in range(steps):
        ctx.flush.zero_()                  # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        run_step()
        ev[k][1].record()
    ctx.barrier()
    wall = time.perf_counter() - wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    # second pass over the same K steps with CUDA events around every stage inside the library (stream launches of
    # the same kernels: event records are not part of the captured graph): stage times, stage flops, and the
    # duration of the roofline kernels
    _lib.set_option("stage_timers", 1)
    _lib.stage_times()
    _lib.stage_flops()
    for k in range(steps):
        ctx.flush.zero_()
        step()
    ctx.barrier()
    stages = _lib.stage_times()
    flops = _lib.stage_flops()
    _lib.set_option("stage_timers", 0)
    if ws > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = n_cols * steps / (ms * 1e-3)
    info = plan.info()

    # ---- N > 1: the sharded results against a single-rank recompute on every rank (outside the timed region) ----
    check = None
    if shard_it:
        J_multi = J.clone()
        inv_rows = holder["inv"].clone()
        J.zero_()
        plan.assemble(perm, var, omega)
        plan.factor(False)
        plan.jacobian_block(0, L, 0, E, C_PERT, omega, J, E, base)
        inv_full = _lib.inverse_model(J, det, LAMBDA, shift=1.0, form=2)
        lo, hi = sharded.bounds[rank], sharded.bounds[rank + 1]
        dJ = float((J_multi - J).abs().max().item())
        sig = float((J - 1).abs().max().item())
        dinv = float(((inv_rows - inv_full[lo:hi]).abs().max() / inv_full.abs().max()).item()) if hi > lo else 0.0
        agg = torch.tensor([dJ, dinv], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dJ, dinv = float(agg[0].item()), float(agg[1].item())
        check = {"max_abs_J_vs_single_rank": dJ, "max_abs_signal_J_minus_1": sig,
                 "max_rel_inv_rows_vs_single_rank": dinv, "tolerance": {"J": 1e-13, "inv": 1e-9},
                 "what": "every rank recomputes the whole Jacobian + inverse model alone after the timed region; the "
                         "gathered J and this rank's rows of inv_mat are compared with it (max over ranks)"}
        if not (dJ <= 1e-13 and dinv <= 1e-9):
            raise SystemExit(f"multi-GPU result differs from the single-rank result ({name}): {check}")
        holder["inv"] = inv_full

    stage_ms = {k: v[0] / steps for k, v in stages.items() if v[1]}
    res = {
        "workload": mesh["workload"], "N": N, "E": E, "L": L, "nU": int(info[3]), "n_det": int(n_det), "m": m_rows,
        "columns_per_step": n_cols, "n_gpus": ws, "mode": ("single GPU" if ws == 1 else mode), "steps": steps,
        "warmup": max(warmup, 3),
        "ms_per_step": ms / steps, "value": value, "unit": "columns/s",
        "stage_ms_per_step": stage_ms, "gpu_launches_per_step": launches / steps,
        "nd_levels": int(info[23]),
        "host_setup_s": {"mesh_model": mesh["mesh_model_s"], "plan_create": plan_s,
                         "note": "once per mesh, outside the timed region (MeshObj element/electrode/detection "
                                 "sets; symbolic nested dissection + table upload)"},
        "signal_J_minus_1_min": float((J - 1).min().item()),
    }
    if check:
        res["multi_gpu_check"] = check
    if sharded is not None and sharded.equal_rows:
        res["nvlink_bytes_per_step_per_rank"] = int(sharded.nvlink_bytes_per_step)
    state = dict(mesh=mesh, plan=plan, J=J, det=det, inv=holder["inv"], stages=stages, flops=flops, steps=steps,
                 launches=launches, wall=wall, clocks=clocks, shard=(i0, i1, e0, e1), info=info)
    return res, state


def stage_rooflines(ctx, st, name):
    """Per-stage roofline entries from the library's stage timers and flop counters (rank-local work)."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    tf = ctx.cublas_dgemm_tflops()
    tf_src = "cuBLAS DGEMM 4096^3 measured in this run (MEASURED_PEAKS.json has no f64 entry)"
    mesh, stages, flops, steps, info = st["mesh"], st["stages"], st["flops"], st["steps"], st["info"]
    N, E, L = len(mesh["nodes"]), len(mesh["elem"]), len(mesh["electrodes"])
    nU = int(info[3])
    i0, i1, e0, e1 = st["shard"]
    out = {}
    for stage, kernels in (("factor", "gemm_dmma_kernel + potrf_trtri_tile_kernel + extend_add / gather kernels"),
                           ("excitation", "rank_update_kernel + gemm_dmma_kernel (ZW, base potentials)"),
                           ("inverse_model", "gemm_dmma_kernel (Gram, split-K; Jt^T G^-1) + potrf_trtri_tile_kernel")):
        ms, n_ = stages.get(stage, (0.0, 0))
        if not n_ or ms <= 0:
            continue
        ach = flops[stage] / (ms * 1e-3) / 1e12
        out[stage] = {"kernel": kernels, "bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s",
                      "frac": ach / tf, "traffic": None, "peak_source": tf_src,
                      "ms_per_step": ms / steps, "algorithmic_flops_per_step": flops[stage] / steps,
                      "note": "stage = a dependent chain of launches on several streams; flops counted by the "
                              "library (2MNK per product, half for lower-only symmetric results, 2/3 n^3 per tile "
                              "factorisation + inverse), time = CUDA events around the stage"}
    # K2 Woodbury: SURVEY.md §8d algorithmic bytes per column = 8 r + 96 + 8 (L-1), r = measurement nodes
    r_nodes = nU - int(np.mean([len(np.unique(mesh["elem"][np.asarray(e)].reshape(-1))) for e in mesh["electrodes"]]))
    bytes_per_col = 8.0 * r_nodes + 96.0 + 8.0 * (L - 1)
    my_cols = st.get("own_cols", (i1 - i0) * (e1 - e0))
    wb_ms, wb_n = stages.get("woodbury", (0.0, 0))
    if wb_n:
        per_launch_ms = wb_ms / wb_n
        cols_per_launch = my_cols * steps / wb_n
        ach = bytes_per_col * cols_per_launch / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))["woodbury_lr_persist_kernel"][name]
            if ctx.ws == 1:                      # the capture is of the full single-GPU launch
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        out["woodbury"] = {
            "kernel": "woodbury_lr_persist_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
            "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
            "ms_per_launch": per_launch_ms, "algorithmic_bytes_per_launch": bytes_per_col * cols_per_launch,
            "timing": "CUDA events around the launch, second pass over the same K steps (stream launches; the timed "
                      "region replays the same kernels from a CUDA graph)",
            "note": "algorithmic bytes = (8 r + 96 + 8 (L-1)) per column (SURVEY.md 8d); the stage timer wraps the "
                    "Woodbury kernel alone"}
    a_ms, a_n = stages.get("assemble", (0.0, 0))
    if a_n and a_ms > 0:
        a_bytes = 16.0 * N + 28.0 * E + 16.0 * int(info[7])        # 16 N + 28 E + 16 nnz
        per = a_ms / a_n
        out["assemble"] = {"kernel": "assemble_elem_kernel + assemble_reduce_kernel", "bound": "hbm", "achieved": a_bytes / (per * 1e-3) / 1e9,
                           "peak": hbm, "unit": "GB/s", "frac": a_bytes / (per * 1e-3) / 1e9 / hbm, "traffic": None,
                           "ms_per_launch": per, "algorithmic_bytes_per_launch": a_bytes, "peak_source": hbm_src}
    return out


def realtime_leg(ctx, inv, n_det, m_rows):
    torch = ctx.torch
    from ceit_b200 import _lib
    dev = ctx.dev
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
    # parity of the full 100 k-frame product against torch.matmul in f64 (outside every timer)
    ref = inv @ dV[:, :4096]
    err = float(((out[:, :4096] - ref).abs().max() / ref.abs().max()).item())
    ref = inv @ dV[:, -4096:]
    err = max(err, float(((out[:, -4096:] - ref).abs().max() / ref.abs().max()).item()))
    chk = float((out.sum(dim=1) - inv @ dV.sum(dim=1)).abs().max().item() / max(1e-300, float(out.sum(dim=1).abs().max().item())))
    if not (err < 1e-12 and chk < 1e-9):
        raise SystemExit(f"batched reconstruction differs from torch.matmul: {err} {chk}")
    del ref
    # cuBLAS DGEMM on the SAME shape (library baseline for this product; the 4096^3 figure is the roofline peak)
    for _ in range(2):
        torch.matmul(inv, dV, out=out)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        torch.matmul(inv, dV, out=out)
    b.record()
    torch.cuda.synchronize()
    t_cublas_same = a.elapsed_time(b) / reps * 1e-3
    # single frame: eager launches and a replayed CUDA graph of 100 frames (the realtime loop of Example_04)
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
    graph = _lib.StepGraph(lambda: [_lib.reconstruct(inv, one, o1) for _ in range(100)] and None)
    graph.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    t_one_graph = a.elapsed_time(b) / 1000 * 1e-3
    flops = 2.0 * n_det * m_rows * N_FRAMES
    dgemm_tf = ctx.cublas_dgemm_tflops()
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
    del dV_h, out_h, dV, out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    gemv_bytes = 8.0 * n_det * m_rows + 8.0 * (n_det + m_rows)
    return {
        "frames_per_s_batched_100k_e2e": N_FRAMES / t_e2e, "ms_batched_100k_e2e": t_e2e * 1e3,
        "e2e_h2d_bytes": int(m_rows * N_FRAMES * 8), "e2e_d2h_bytes": int(n_det * N_FRAMES * 8),
        "frames_per_s_batched_100k": N_FRAMES / t_batch, "ms_batched_100k": t_batch * 1e3,
        "batched_tflops_f64": flops / t_batch / 1e12, "cublas_dgemm_4096_tflops": dgemm_tf,
        "batched_frac_of_cublas_dgemm": flops / t_batch / 1e12 / dgemm_tf,
        "cublas_dgemm_same_shape_tflops": flops / t_cublas_same / 1e12,
        "batched_frac_of_cublas_same_shape": t_cublas_same / t_batch,
        "batched_max_rel_err_vs_torch_matmul_f64": err, "batched_checksum_rel_err": chk,
        "frames_per_s_single": 1.0 / t_one, "us_single_frame": t_one * 1e6,
        "us_single_frame_graph_replay": t_one_graph * 1e6,
        "single_frame_weight_stream_gbs": gemv_bytes / t_one_graph / 1e9,
        "single_frame_frac_of_hbm": gemv_bytes / t_one_graph / 1e9 / hbm,
        "single_frame_note": "inv_mat (7.8 MB) is L2-resident between frames, so the GB/s figure is an L2 stream rate "
                             "quoted against the HBM peak; graph replay = 100 frames per captured graph",
    }


def e2e_leg(ctx, mesh, n_cols, steps):
    """the drop-in classes, host (numpy) in / out, cache files written like the reference"""
    torch = ctx.torch
    from ceit_b200 import meshgen
    E, L = len(mesh["elem"]), len(mesh["electrodes"])
    m_rows, n_det = L * (L - 1), len(mesh["detection_index"])
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
            jac.save_inv_matrix(LAMBDA)              # inverse model on the resident J, D2H inv_mat, np.save
            return Jh

        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": n_cols * steps / dt, "unit": "columns/s",
               "h2d_bytes_per_step": int(2 * E * 8),
               "d2h_bytes_per_step": int(m_rows * E * 8 + n_det * m_rows * 8 + m_rows * 8),
               "ms_per_step": 1e3 * dt / steps,
               "api": "EJAC.JAC_calculation() + EJAC.save_inv_matrix(289) (writes JAC_cache.npy / inv_mat.npy)"}
        # informational: the same loop with the opt-in write-behind of the two cache files (worker thread, joined
        # inside the timed region); the headline above keeps the reference's synchronous writes
        from ceit_b200.util.utilities import wait_npy
        os.environ["CEIT_B200_WRITE_BEHIND"] = "1"
        for _ in range(3):
            e2e_step()
        wait_npy()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        wait_npy()
        torch.cuda.synchronize()
        e2e["ms_per_step_write_behind"] = 1e3 * (time.perf_counter() - t0) / steps
        del os.environ["CEIT_B200_WRITE_BEHIND"]
        del os.environ["CEIT_B200_CONFIG"]
    if ctx.ws > 1:
        e2e["note"] = "single-rank API call on rank 0 (the class API is single-process)"
    return e2e


def _lib_error():
    from ceit_b200 import _lib
    return _lib.CeitError


def run_gpu(args):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    ctx = Ctx()
    ws, rank = ctx.ws, ctx.rank
    use_graph = not args.no_graph
    head_name = args.config
    def measure_best(name, steps, warmup, sample_clocks, candidates):
        """N > 1: measure the candidate schedules and keep the fastest (all are reported under `modes`)."""
        best, best_st, modes = None, None, {}
        for mode in candidates:
            try:
                r, s_ = measure_config(ctx, name, steps, warmup, use_graph, sample_clocks=sample_clocks, mode=mode)
            except _lib_error() as e:                      # e.g. fewer sub-trees than ranks on a tiny mesh
                modes[mode] = {"unavailable": str(e)[:200]}
                continue
            modes[mode] = {"ms_per_step": r["ms_per_step"], "value": r["value"],
                           "stage_ms_per_step": r["stage_ms_per_step"]}
            for k in ("multi_gpu_check", "nvlink_bytes_per_step_per_rank", "domain"):
                if k in r:
                    modes[mode][k] = r[k]
            if best is None or r["ms_per_step"] < best["ms_per_step"]:
                if best_st is not None:
                    best_st["plan"].close()
                best, best_st = r, s_
            else:
                s_["plan"].close()
            torch.cuda.empty_cache()
        modes["used"] = best["mode"]
        return best, best_st, modes

    modes = None
    if ws > 1:
        # a mesh this small may not amortise any exchange latency: the replicated schedule is measured too and the job
        # uses the fastest (all reported)
        head, st, modes = measure_best(head_name, args.steps, args.warmup, True, ("domain", "sharded", "replicated"))
    else:
        head, st = measure_config(ctx, head_name, args.steps, args.warmup, use_graph, sample_clocks=True)
    is_headline = head_name == "50_50"
    mesh = st["mesh"]
    n_det, m_rows = head["n_det"], head["m"]

    realtime = realtime_leg(ctx, st["inv"], n_det, m_rows) if (rank == 0 and is_headline) else None
    e2e = e2e_leg(ctx, mesh, head["columns_per_step"], min(args.steps, 100)) if (rank == 0 and is_headline) else None
    rooflines = None
    if rank == 0:
        rooflines = {"stage_ms_per_step": head["stage_ms_per_step"]}
        rooflines.update(stage_rooflines(ctx, st, head_name))
        if realtime:
            rooflines["reconstruct_batched"] = {
                "kernel": "gemm_dmma_kernel<64,64> (ceit_reconstruct, F = 100000)", "bound": "tensor",
                "achieved": realtime["batched_tflops_f64"], "peak": realtime["cublas_dgemm_4096_tflops"],
                "unit": "TFLOP/s", "frac": realtime["batched_frac_of_cublas_dgemm"], "traffic": None,
                "peak_source": "cuBLAS DGEMM 4096^3 measured in this run (no f64 entry in MEASURED_PEAKS.json)"}
            rooflines["reconstruct_single_frame"] = {
                "kernel": "gemv_rows_kernel (ceit_reconstruct, F = 1)", "bound": "hbm",
                "achieved": realtime["single_frame_weight_stream_gbs"], "unit": "GB/s",
                "frac": realtime["single_frame_frac_of_hbm"], "traffic": None,
                "us_per_frame": realtime["us_single_frame_graph_replay"], "note": realtime["single_frame_note"]}
    clocks = st["clocks"]
    st["plan"].close()
    del st

    # ---- the other BASELINE.json configs at the same number of GPUs (same line: `scaling_configs`) ----
    scaling = {}
    if is_headline and not args.headline_only:
        for name in ("21_21", "200", "400"):
            try:
                if ws > 1:
                    cands = ("domain", "sharded", "replicated") if name == "21_21" else ("domain", "sharded")
                    r, s2, md = measure_best(name, CONFIGS[name]["steps"], 3, False, cands)
                    r["modes"] = md
                else:
                    r, s2 = measure_config(ctx, name, CONFIGS[name]["steps"], 3, use_graph)
                if rank == 0:
                    r["rooflines"] = stage_rooflines(ctx, s2, name)
                s2["plan"].close()
                del s2
                torch.cuda.empty_cache()
                scaling[name] = r
            except SystemExit:
                raise
            except Exception as e:   # a config that cannot run (e.g. out of memory) is reported, not hidden
                scaling[name] = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and is_headline:
        arm = CpuArm()
        cpu_cols = 6
        v = arm.columns_per_s(cpu_cols)
        cpu = {"value": v, "unit": "columns/s", "cores": blas_threads(), "kind": arm.kind,
               "sample": f"{cpu_cols} forward solves of MESH_50_50 = columns of the reference Jacobian ({arm.what})"}
        inv_h = np.random.default_rng(1).random((n_det, m_rows))
        dvh = np.random.default_rng(0).random((m_rows, 20000))
        t0 = time.perf_counter()
        np.dot(inv_h, dvh)
        cpu["realtime_frames_per_s_batched"] = 20000 / (time.perf_counter() - t0)

    if rank == 0:
        line = {
            "metric": "jacobian_columns_per_s", "value": head["value"], "unit": "columns/s", "n_gpus": ws,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": ("MESH_50_50 mesh fixture (tests/golden/mesh_50_50.npz), zero base variable; seeded synthetic frames"
                     if is_headline else "synthetic structured mesh (ceit_b200.meshgen), zero base variable"),
            "config": {"workload": head["workload"], "N": head["N"], "E": head["E"], "L": head["L"],
                       "columns_per_step": head["columns_per_step"],
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "ordering": f"multi-level nested dissection, {head['nd_levels']} levels",
                       "sharding": ("single GPU" if ws == 1 else {
                           "domain": f"domain decomposition over {ws} ranks: each rank factorises its sub-trees of the "
                                     "elimination forest and computes the Jacobian columns of its own elements for all "
                                     "excitations; one all-gather (sub-tree roots' update matrices) + one all-reduce "
                                     "(top right-hand sides, partial S_U) inside the base stage, one all-reduce of the "
                                     "m x m Gram matrix; J replicated by an all-gather overlapped with the inverse model",
                           "sharded": f"excitation blocks over {ws} ranks, base stage replicated; all-to-all of the "
                                      "detection-column slices + all-reduce of the m x m Gram matrix (NCCL), J "
                                      "replicated by an in-place all-gather overlapped with the inverse model",
                           "replicated": f"replicated on {ws} ranks: this mesh does not amortise the exchange latency "
                                         "(every schedule was measured, see `modes`)"}[head["mode"]]),
                       "modes": modes},
            "gpu_launches": int(head["gpu_launches_per_step"] * args.steps),
            "launch_mode": (("cuda_graph" if ws == 1 else "cuda_graph(local part) + NCCL exchange") if use_graph
                            else "stream"),
            "host_setup_s": head["host_setup_s"],
            "clocks": None, "e2e": e2e, "roofline": (rooflines or {}).get("woodbury"), "rooflines": rooflines,
            "cpu_baseline": cpu, "realtime": realtime, "scaling_configs": scaling,
        }
        if "multi_gpu_check" in head:
            line["multi_gpu_check"] = head["multi_gpu_check"]
        if "nvlink_bytes_per_step_per_rank" in head:
            line["nvlink_bytes_per_step_per_rank"] = head["nvlink_bytes_per_step_per_rank"]
        line["clocks"] = clocks
        print(json.dumps(line))
    if ws > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--config", default="50_50", choices=sorted(CONFIGS), help="headline workload (default: MESH_50_50)")
    ap.add_argument("--headline-only", action="store_true", help="skip the scaling_configs legs (21_21, 200, 400)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
