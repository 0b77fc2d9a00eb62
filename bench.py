#!/usr/bin/env python
"""bench.py -- collision-checked configs/sec on the Franka validity sweep (BASELINE configs[1]), and
batched bi-RRT plans/sec (the second half of BASELINE's metric) as extra keys of the same line.

One "step" = one pass of the validity path (joint-limit mask + FK + group / capsule / OBB culls +
narrow phase + fp64 re-evaluation of uncertain items) over 1,000,000 synthetic Franka rows
(q ~ U[jnt_range], np.random.default_rng(rank), fp32; scene_with_obstacles, left/right finger pair
allowed).  `value` is timed with the rows resident in HBM; `e2e` is the same metric through the
public API with pinned HOST buffers (H2D of the rows and D2H of the mask inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); rows are independent, every rank checks its own
1M-row block (weak scaling), no data-path collective in the timed steps; time = max over ranks.
Extra keys (same JSON line): `plans` (4096 Franka planning queries sharded over the ranks, against a
measured all-core CPU planner baseline) and `sweep` (BASELINE configs[4]: a 1e9-row device-generated
sweep sharded over the ranks, bit-packed masks gathered with NCCL inside a timed sub-region).
`--impl reference` times the CPU implementation of the reference path on all host cores, pinned, on
a bounded sample of the same workload: real MuJoCo through the reference's exact call sequence if
`import mujoco` works on this box, else the fp64 oracle port (the `mujoco` wheel is not installable
in this image).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

ROWS_PER_STEP = 1_000_000
MODEL = "franka_scene_with_obstacles"
ALLOWED = [("left_finger", "right_finger")]
WORKLOAD = ("Franka Panda validity sweep: 1M uniformly sampled q in joint limits, FK + full self/scene "
            "collision (scene_with_obstacles.xml, left_finger/right_finger allowed), limits+FK+collision fused")
METRIC = "collision-checked configs/sec"
UNIT = "configs/s"
ALG_BYTES_PER_ROW = 37  # 9 fp32 joint values in + 1 validity byte out
PLAN_QUERIES = 4096
PLAN_JOINTS = [f"joint{i}" for i in range(1, 8)]
SWEEP_ROWS = 1_000_000_000


def config_dict(world):
    """identical keys and values in both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "rows_per_step_per_gpu": ROWS_PER_STEP,
            "l2": "256 MiB buffer written between timed steps (L2 flush); per-step CUDA events",
            "parallelism": f"{world} independent row blocks, no collective"}


def make_rows(model, n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(model.jnt_range[:, 0], model.jnt_range[:, 1], size=(n, model.nq)).astype(np.float32)


def host_cores():
    return sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))


def any_rank(flag, dist, device):
    """True on every rank iff `flag` is true on at least one (all-reduce MAX; `dist` None: single process)."""
    if dist is None:
        return bool(flag)
    import torch

    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


def bind_to_gpu_numa(local_rank):
    """Pin this rank to the CPU cores of its GPU's NUMA node (pinned-memory copies from the far node
    halve the host-side rate when 8 ranks all sit on node 0).  Best effort; returns a description."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip()) if path.exists() else -1
        if node < 0:
            return "numa: unknown node, affinity unchanged"
        cl = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        cores = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cores.update(range(int(a), int(b or a) + 1))
        cores &= set(os.sched_getaffinity(0))
        if cores:
            os.sched_setaffinity(0, cores)
            return f"numa node {node}: {len(cores)} cores"
        return f"numa node {node}: no allowed core, affinity unchanged"
    except Exception as e:  # noqa: BLE001
        return f"numa: {type(e).__name__}, affinity unchanged"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.skip = 0

    def start(self):
        """Launch the poller and wait for its first sample, so that NVML start-up (which can stall
        kernel launches for milliseconds) happens BEFORE the timed region, not inside it."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.01)
            self.skip = len(self.lines)  # samples taken while the GPU was still idle
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines[self.skip:]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU side: the reference path on the host cores
# ----------------------------------------------------------------------------------------------------
def mujoco_checker():
    """The reference's own arithmetic, if this box has it: `import mujoco` plus the scene XML (the
    reference checkout).  Returns (check(rows) -> mask, description) or None.  The call sequence is
    the reference's (collision_constraint.py:26-30, :83-95; joint_limit_constraint.py:19-20)."""
    try:
        import mujoco  # noqa: F401
    except Exception:
        return None
    for base in (os.environ.get("MJPL_REFERENCE_MODELS"), "/root/reference/examples/models", str(ROOT / "baseline/_ref/examples/models")):
        xml = Path(base) / "franka_emika_panda" / "scene_with_obstacles.xml" if base else None
        if xml is not None and xml.exists():
            break
    else:
        return None
    import mujoco

    model = mujoco.MjModel.from_xml_path(str(xml))
    allowed = {tuple(sorted((model.body(a).id, model.body(b).id))) for a, b in ALLOWED}
    lo, hi = model.jnt_range[:, 0].copy(), model.jnt_range[:, 1].copy()

    def check_block(rows):
        data = mujoco.MjData(model)
        out = np.zeros(len(rows), dtype=bool)
        for i, q in enumerate(rows):
            if not np.all((q >= lo) & (q <= hi)):
                continue
            data.qpos = q
            mujoco.mj_kinematics(model, data)
            mujoco.mj_collision(model, data)
            ok = True
            for g in data.contact.geom:
                if tuple(sorted(model.geom_bodyid[g])) not in allowed:
                    ok = False
                    break
            out[i] = ok
        return out

    return check_block, f"mujoco {mujoco.__version__}: data.qpos=q; mj_kinematics; mj_collision; allow-list test, one pinned process per core"


def _mujoco_worker(args):
    core, rows = args
    os.sched_setaffinity(0, {core})
    chk, _ = mujoco_checker()
    t0 = time.perf_counter()
    m = chk(rows)
    return m, time.perf_counter() - t0


def cpu_validity(model, rows, seconds_target):
    """Validity of a bounded sample of `rows` on ALL host cores, pinned -> (rate, mask, n, kind, description)."""
    cores = host_cores()
    mj_ref = mujoco_checker()
    if mj_ref is not None:
        import multiprocessing as mp

        chk, desc = mj_ref
        probe = rows[:2000].astype(np.float64)
        t0 = time.perf_counter()
        chk(probe)
        rate1 = len(probe) / (time.perf_counter() - t0)
        n = int(min(len(rows), max(2000 * len(cores), rate1 * len(cores) * seconds_target)))
        parts = np.array_split(rows[:n].astype(np.float64), len(cores))
        with mp.get_context("spawn").Pool(len(cores)) as pool:
            t0 = time.perf_counter()
            res = pool.map(_mujoco_worker, list(zip(cores, parts)))
            dt = time.perf_counter() - t0
        dt = max(dt_i for _, dt_i in res)
        return n / dt, np.concatenate([m for m, _ in res]), n, "reference", f"first {n} rows of the step's 1M-row block, {desc}, {dt:.1f} s"
    import oracle

    orc = oracle.Oracle(model, ALLOWED)
    oracle.Oracle.set_threads(len(cores))
    flags = oracle.CHECK_LIMITS | oracle.CHECK_COLLISION
    probe = rows[:20000].astype(np.float64)
    t0 = time.perf_counter()
    orc.check(probe, flags)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(min(len(rows), max(20000, rate * seconds_target)))
    sample = rows[:n].astype(np.float64)
    t0 = time.perf_counter()
    valid = orc.check(sample, flags)
    dt = time.perf_counter() - t0
    desc = (f"first {n} rows of the step's 1M-row block, fp64 C oracle port of the reference path (restated MuJoCo semantics; "
            f"`import mujoco` fails on this box), {len(cores)} pinned pthreads, {dt:.1f} s")
    return n / dt, valid, n, "port", desc


def run_reference(args):
    """--impl reference: the CPU path, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mjpl_b200 import models

    model = models.load(MODEL)
    rows = make_rows(model, ROWS_PER_STEP)
    cores = host_cores()
    budget = 120.0 / max(1, args.steps + args.warmup)
    rate, _, n, kind, desc = cpu_validity(model, rows, min(budget, 15.0))   # also sizes the per-step sample
    if kind == "port":
        import oracle

        orc = oracle.Oracle(model, ALLOWED)
        oracle.Oracle.set_threads(len(cores))
        flags = oracle.CHECK_LIMITS | oracle.CHECK_COLLISION
        sample = rows[:n].astype(np.float64)
        for _ in range(args.warmup):
            orc.check(sample, flags)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.check(sample, flags)
        dt = time.perf_counter() - t0
    else:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_validity(model, rows[:n], 1e9)
        dt = time.perf_counter() - t0
    value = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(cores), "kind": kind,
                         "sample": f"{n} of the step's 1M rows per step; " + desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- CPU planner baseline: the reference's sequential bi-RRT, one pinned process per core -----------
def _cpu_planner_worker():
    """`bench.py --cpu-planner-worker CORE FIRST STRIDE COUNT GOALS.npy`: plans queries FIRST, FIRST+STRIDE, ...
    with the sequential planner (reference control flow, rrt.py:141-237 / planning/utils.py:105-164, one
    configuration per constraint call) on the fp64 oracle; prints {"solved", "seconds", "queries"}."""
    i = sys.argv.index("--cpu-planner-worker")
    core, first, stride, count = (int(x) for x in sys.argv[i + 1:i + 5])
    goals = np.load(sys.argv[i + 5])
    os.sched_setaffinity(0, {core})
    import oracle
    from mjpl_b200 import models
    from mjpl_b200.planning.rrt import RRT
    from oracle.constraints import ReferenceCollisionConstraint, ReferenceJointLimitConstraint

    oracle.Oracle.set_threads(1)
    model = models.load(MODEL)
    cons = [ReferenceJointLimitConstraint(model), ReferenceCollisionConstraint(model, ALLOWED)]
    q_init = model.keyframe("home").qpos.copy()
    solved, done = 0, 0
    t0 = time.perf_counter()
    for k in range(count):
        qi = first + k * stride
        if qi >= len(goals):
            break
        r = RRT(model, PLAN_JOINTS, cons, max_planning_time=10.0, epsilon=0.05, seed=qi, goal_biasing_probability=0.1)
        solved += bool(r.plan_to_config(q_init, goals[qi]))
        done += 1
    print(json.dumps({"solved": solved, "queries": done, "seconds": time.perf_counter() - t0}))


def cpu_planner_baseline(goals, per_core=64):
    cores = host_cores()
    tmp = Path(os.environ.get("TMPDIR", "/tmp")) / f"bench_goals_{os.getpid()}.npy"
    np.save(tmp, goals)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, str(ROOT / "bench.py"), "--cpu-planner-worker", str(c), str(k), str(len(cores)),
                               str(per_core), str(tmp)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                              env={**os.environ, "OMP_NUM_THREADS": "1", "CUDA_VISIBLE_DEVICES": ""})
             for k, c in enumerate(cores)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    wall = time.perf_counter() - t0
    tmp.unlink(missing_ok=True)
    res = [json.loads(o.strip().splitlines()[-1]) for o in outs if o.strip()]
    solved, queries = sum(r["solved"] for r in res), sum(r["queries"] for r in res)
    busy = max(r["seconds"] for r in res) if res else float("nan")
    return {"plans_per_s": solved / busy if res else None, "solved": solved, "queries": queries, "cores": len(cores),
            "seconds": busy, "wall_seconds_incl_process_start": wall,
            "what": "the reference's sequential bi-RRT (one configuration per constraint call, 10 s budget per query) on the fp64 "
                    "oracle port, one pinned process per host core, all cores busy at once (measured, not extrapolated)"}


# ----------------------------------------------------------------------------------------------------
def plan_leg(model, eng, rank, world, dist, torch):
    """4096 independent Franka planning queries (home -> random valid goal, obstacle scene), sharded
    over the ranks; device-timed wall per rank, max over ranks."""
    import mjpl_b200 as mj

    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, ALLOWED)]
    q_init = model.keyframe("home").qpos.copy()
    rows = eng.sweep_rows(7, 0, 8 * PLAN_QUERIES).double().cpu().numpy()
    rows[:, 7:] = q_init[7:]
    ok = np.asarray(mj.obeys_constraints_batch(rows, c))
    goals = rows[ok][:PLAN_QUERIES]
    lo, hi = (len(goals) * rank) // world, (len(goals) * (rank + 1)) // world
    mine = goals[lo:hi]
    planner = mj.BatchedRRT(model, PLAN_JOINTS, c, max_planning_time=60.0, epsilon=0.05, seed=rank, goal_biasing_probability=0.1,
                            max_active=4096, max_iterations_per_query=2000, sync_every=32)
    # warm-up: the same batch once, untimed -- the tree arrays grow by doubling (hundreds of MB per tree at 4,096
    # slots) and the first pass pays a cudaMalloc for every size; nothing else is kept between the passes (plan()
    # rebuilds its trees and draws the same counter-based random stream)
    planner.plan(np.tile(q_init, (len(mine), 1)), mine)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    paths = planner.plan(np.tile(q_init, (len(mine), 1)), mine)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    solved = sum(1 for p in paths if p)
    t = torch.tensor([dt, float(solved), float(planner.stats.get("configs_checked", 0))], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt = float(mx[0])
    out = {"plans_per_s": float(t[1]) / dt, "queries": len(goals), "solved": int(t[1]), "seconds": dt,
           "configs_checked": int(t[2]), "scene": MODEL, "per_rank_queries": len(mine),
           "what": "batched bi-RRT (BatchedRRT.plan), home -> random valid goal, epsilon 0.05, goal bias 0.1, 2000 iterations per query; second pass over the batch (the first, untimed, warms the allocator)"}
    # every path of rank 0's first queries replays valid under the CPU checker
    if rank == 0:
        import oracle

        orc = oracle.Oracle(model, ALLOWED)
        oracle.Oracle.set_threads(len(host_cores()))
        bad = 0
        for i, p in enumerate(paths[:200]):
            if not p:
                continue
            P = np.array(p)
            good = orc.check(P, 3).all() and (np.linalg.norm(np.diff(P, axis=0), axis=1) <= 0.05 + 1e-9).all()
            bad += 0 if (good and np.array_equal(P[0], q_init) and np.array_equal(P[-1], mine[i])) else 1
        out["replay_failures_of_200"] = bad
    return out, goals


def sweep_leg(eng, rank, world, dist, torch):
    """BASELINE configs[4]: a 1e9-row validity sweep generated on the device from the global row id,
    sharded over the ranks; every rank bit-packs its mask and the packed masks are all-gathered with
    NCCL (timed on its own)."""
    lo, hi = (SWEEP_ROWS * rank) // world, (SWEEP_ROWS * (rank + 1)) // world
    lo, hi = (lo + 7) // 8 * 8, (hi + 7) // 8 * 8 if rank + 1 < world else hi
    block = 8_000_000
    packed = torch.empty(((hi - lo) + 7) // 8, dtype=torch.uint8, device="cuda")
    mask = torch.empty(block, dtype=torch.uint8, device="cuda")
    weights = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device="cuda")
    eng.sweep(2024, lo, min(block, hi - lo), out=mask[:min(block, hi - lo)])   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    nvalid = torch.zeros((), dtype=torch.int64, device="cuda")
    for r0 in range(lo, hi, block):
        n = min(block, hi - r0)
        eng.sweep(2024, r0, n, out=mask[:n])
        nvalid += mask[:n].sum(dtype=torch.int64)
        m8 = mask[:n]
        if n % 8:
            m8 = torch.cat([m8, torch.zeros(8 - n % 8, dtype=torch.uint8, device="cuda")])
        packed[(r0 - lo) // 8:(r0 - lo) // 8 + (n + 7) // 8] = (m8.view(-1, 8) * weights).sum(dim=1, dtype=torch.uint8)
    e1.record()
    gathered_bytes = int(packed.numel())
    if world > 1:
        per = (SWEEP_ROWS // world + 15) // 8
        buf = torch.zeros(per, dtype=torch.uint8, device="cuda")
        buf[:packed.numel()] = packed
        allm = torch.empty(per * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allm, buf)
        gathered_bytes = int(allm.numel())
    e2.record()
    torch.cuda.synchronize()
    # band accounting on the device (mjb_min_distance, fp64 signed distance): rows of this shard's first block
    # whose distance to contact lies within 1e-5 -- the only rows on which a checker may legitimately disagree
    nb = min(2_000_000, hi - lo)
    qb = eng.sweep_rows(2024, lo, nb)
    tb = time.perf_counter()
    sd, _ = eng.min_distance(qb)
    torch.cuda.synchronize()
    tb = time.perf_counter() - tb
    mb = eng.sweep(2024, lo, nb, flags=2)
    in_band = (sd.abs() < 1e-5)
    band = torch.tensor([float(in_band.sum()), float(((mb != 0) != (sd > 0))[~in_band].sum()), float(nb)], dtype=torch.float64, device="cuda")
    t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device="cuda")
    nv = nvalid.to(torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nv, op=dist.ReduceOp.SUM)
        dist.all_reduce(band, op=dist.ReduceOp.SUM)
    return {"rows": SWEEP_ROWS, "band": {"sample_rows": int(band[2]), "rows_within_1e-5_of_contact": int(band[0]),
                                         "collision_mask_vs_sign_of_distance_mismatches_outside_band": int(band[1]),
                                         "min_distance_rows_per_s_per_gpu": nb / tb,
                                         "what": "mjb_min_distance (fp64 GJK / EPA signed distance) on the first 2M rows of every shard"}, "seconds": float(t[0]) * 1e-3, "configs_per_s": SWEEP_ROWS / (float(t[0]) * 1e-3),
            "valid_fraction": float(nv) / SWEEP_ROWS, "mask_gather_ms": float(t[1]), "gathered_bytes": gathered_bytes,
            "what": "mjb_check_sweep in 8M-row blocks over this rank's shard of the global row range, masks bit-packed on the "
                    "device, packed shards all-gathered with NCCL (all_gather_into_tensor)" if world > 1 else
                    "mjb_check_sweep in 8M-row blocks, masks bit-packed on the device (single GPU: nothing to gather)"}


def main():
    if "--cpu-planner-worker" in sys.argv:
        return _cpu_planner_worker()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plans", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import mjpl_b200 as mj
    from mjpl_b200 import models
    from mjpl_b200.engine import fma_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else "single rank: affinity unchanged"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)

    model = models.load(MODEL)
    eng = mj.get_engine(model, ALLOWED)
    rows = make_rows(model, ROWS_PER_STEP, seed=rank)  # rank r checks its own independent block
    q_dev = torch.from_numpy(rows).cuda()
    q_pin = torch.from_numpy(rows).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    FLAGS = 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak = fma_peak()   # measured FP32 FMA throughput of this GPU (roofline denominator)

    # ---- device-resident timing: per-step CUDA events on the launch stream, L2 flushed between steps.
    # A pass whose event-timed steps add up to much more than its kernels (a host hiccup between the
    # launches of one step: the events then time an idle GPU; more than 4 % over the kernels) is taken again, at most three times;
    # the JSON line says how many passes were retaken.
    def timed_pass():
        for _ in range(warmup):
            eng.valid_configs(q_dev, FLAGS)
        torch.cuda.synchronize()
        eng.reset_stats()
        eng.kernel_timing(True)   # CUDA events around each kernel of the launch, on the launch stream
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.fill_(1)
            e0.record()
            m = eng.valid_configs(q_dev, FLAGS)
            e1.record()
        barrier()
        return m, [a.elapsed_time(b) for a, b in evs], eng.kernel_timing(False, read=True), sampler.stop(), eng.stats()

    retakes = 0
    while True:
        mask, step_ms, ktime, clocks, st = timed_pass()
        kernels_ms = sum(ktime[k] for k in ("first_ms", "mid_ms", "narrow_ms", "fp64_ms"))
        # the decision to take the pass again is COLLECTIVE: a pass contains barriers, so every rank has to take
        # the same number of passes (a rank deciding on its own numbers leaves the others waiting in a barrier)
        again = any_rank(sum(step_ms) > 1.04 * kernels_ms, dist if world > 1 else None, torch.device("cuda", local))
        if not again or retakes >= 3:
            break
        retakes += 1
    total_ms = float(sum(step_ms))
    launches = st["launches"]

    # ---- end to end through the public API: pinned host rows in, host mask out, every step
    for _ in range(3):
        eng.valid_configs(q_pin, FLAGS)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_mask = eng.valid_configs(q_pin, FLAGS)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    rows_all = ROWS_PER_STEP * world * args.steps
    value = rows_all / (total_ms * 1e-3)
    e2e_value = rows_all / (e2e_ms * 1e-3)

    plans = goals = sweep = None
    if not args.no_plans:
        plans, goals = plan_leg(model, eng, rank, world, dist, torch)
    if not args.no_sweep:
        sweep = sweep_leg(eng, rank, world, dist, torch)

    if rank == 0:
        step = total_ms / args.steps
        nl = max(1, ktime["launches"])
        first_ms, mid_ms, narrow_ms, fp64_ms = (ktime[k] / nl for k in ("first_ms", "mid_ms", "narrow_ms", "fp64_ms"))
        piped = ktime["pipeline"] != "single"
        per_kernel = ({"fk_cull_kernel": first_ms, "mid_kernel": mid_ms, "narrow_kernel": narrow_ms} if piped
                      else {"validity_kernel": first_ms})
        kernel_name = max(per_kernel, key=per_kernel.get)
        kernel_ms = per_kernel[kernel_name] if per_kernel[kernel_name] > 0 else step
        prof = {}
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            prof = json.loads(tf.read_text())
        kp = prof.get("kernels", {}).get(kernel_name, {})
        flop_row = kp.get("fp32_flop_per_row")
        inst_row = kp.get("thread_inst_per_row")
        flop_row_step = sum(v.get("fp32_flop_per_row", 0) for k, v in prof.get("kernels", {}).items() if k in per_kernel) or None
        traffic_step = sum(v.get("dram_bytes_per_launch", 0) for k, v in prof.get("kernels", {}).items() if k in per_kernel) or None
        hbm_file = ROOT / "MEASURED_PEAKS.json"
        if hbm_file.exists():
            hbm_peak, hbm_src = float(json.loads(hbm_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        achieved_tf = (flop_row * ROWS_PER_STEP / (kernel_ms * 1e-3) / 1e12) if flop_row else None
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * 32 * sm_mhz * 1e6   # thread-instructions per second: 148 SMs x 4 schedulers x 32 lanes
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(q_pin.numel() * 4),
                    "d2h_bytes_per_step": int(ROWS_PER_STEP), "api": "ValidityEngine.valid_configs(pinned CPU tensor)"},
            "gpu_launches": int(launches),
            # The path is bound by the CUDA cores (FP32 / instruction issue), not by HBM and not by tensor
            # cores: the roofline block is the FP32 roof of the dominant kernel against the FMA peak
            # MEASURED in this run; the HBM view (37 algorithmic bytes per row) is kept beside it.
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": peak["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak["tflops"]) if achieved_tf else None,
                         "traffic": traffic_step, "kernel": kernel_name,
                         "peak_source": f"mjb_fma_peak: FFMA micro-kernel measured in this run ({peak['ms']:.2f} ms best pass)",
                         "flop_per_row": flop_row, "flop_source": prof.get("source"),
                         "kernel_ms": {"dominant": kernel_ms, **per_kernel, "fp64_item_pass": fp64_ms,
                                       "share_of_step": kernel_ms / step, "step_over_kernels": sum(step_ms) / max(kernels_ms, 1e-9),
                                       "retaken_passes": retakes,
                                       "step_ms_min_median_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))]},
                         "step": {"fp32_flop_per_row": flop_row_step,
                                  "achieved_tflops": (flop_row_step * ROWS_PER_STEP / (step * 1e-3) / 1e12) if flop_row_step else None,
                                  "frac": (flop_row_step * ROWS_PER_STEP / (step * 1e-3) / 1e12 / peak["tflops"]) if flop_row_step else None},
                         "issue": {"thread_inst_per_row": inst_row,
                                   "frac_of_issue_peak": (inst_row * ROWS_PER_STEP / (kernel_ms * 1e-3) / issue_peak) if inst_row else None,
                                   "peak": "148 SMs x 4 schedulers x 32 lanes x SM clock under load"},
                         "hbm": {"bytes_per_row": ALG_BYTES_PER_ROW, "achieved_gbs": ALG_BYTES_PER_ROW * ROWS_PER_STEP / (step * 1e-3) / 1e9,
                                 "peak_gbs": hbm_peak, "frac": ALG_BYTES_PER_ROW * ROWS_PER_STEP / (step * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": hbm_src, "dram_bytes_per_step_ncu": traffic_step}},
            "stats": {"valid_fraction": float(mask.float().mean()), "narrow_items_per_row": st["narrow_items"] / max(1, st["rows"]),
                      "fp64_items_per_row": st["uncertain_rows"] / max(1, st["rows"]),
                      "queue_overflow_rows": st["queue_overflow"], "numa": numa},
        }
        if plans is not None:
            out["plans_per_s"] = plans["plans_per_s"]
            out["plans"] = plans
        if sweep is not None:
            out["sweep"] = sweep
        # the CPU legs run at N = 1 only (task contract): at N > 1 the other ranks would sit in the final barrier for
        # the half minute they take, and the host cores are shared with N - 1 busy rank processes anyway
        if not args.no_cpu_baseline and world == 1:
            rate, cpu_valid, n_cpu, kind, desc = cpu_validity(model, rows, 12.0)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": len(host_cores()), "kind": kind, "sample": desc}
            out["stats"]["agreement_with_cpu_checker_on_sample"] = float((cpu_valid == host_mask[:n_cpu].numpy()).mean())
            if plans is not None:
                cpu_plans = cpu_planner_baseline(goals)
                out["plans"]["cpu_baseline"] = cpu_plans
                if cpu_plans["plans_per_s"]:
                    out["plans"]["vs_cpu_all_cores"] = plans["plans_per_s"] / cpu_plans["plans_per_s"]
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
