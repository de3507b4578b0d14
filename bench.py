#!/usr/bin/env python
"""Benchmark of the hot path: min-snap solve + sample + collision check (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

One *step* = one pass of the whole pipeline over the per-GPU batch.  Workload (BASELINE.json
configs[4], the one the metric is quoted on): 1,048,576 independent trajectories per GPU x 10
pieces x 3 axes, one time vector per trajectory (T_i ~ U(0.5, 2) s), S = 100 samples each,
robot `custom_triangle_robot` vs obstacle `env-scene-ltu-experiment`, seed 20261019 (SURVEY
§8d).  Inputs + outputs are ~2.4 GB per step, far larger than the 126 MB L2.

Multi-GPU (torchrun, one process per GPU): trajectories are sharded by index, every rank
runs the same per-GPU batch (weak scaling), and the results are all-gathered over NVLink: by
default what the reference produces — path_to_pol's float32 polynomial matrix + the collision
flags — stored to every peer's buffer from inside the single-pass kernel (peer stores); the
FP64-coefficient gather (copy-engine push) and the flags-only gather are timed beside it, plus
the strong-scaling reading of configs[4] (1 M trajectories in total).

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEG, K_AX, S_SAMPLES = 10, 3, 100
TRAJ_PER_GPU = 1 << 20
SEED = 20261019
ROBOT, ENV = "custom_triangle_robot", "env-scene-ltu-experiment"
METRIC = "min-snap trajectories solved+collision-checked per second"
UNIT = "trajectories/s"
# compulsory HBM bytes per trajectory: waypoints + stamps in, coefficients +
# per-sample flags + any-flag out (durations and solver status are extra outputs we also write)
ALG_BYTES = (N_SEG + 1) * K_AX * 8 + (N_SEG + 1) * 8 + N_SEG * K_AX * 64 + S_SAMPLES + 1
BOUNDS_LO = np.array([-2.2, 2.8, 0.5])
BOUNDS_HI = np.array([2.2, 5.0, 2.5])


def make_workload(B, seed):
    """Config-5 generator: start uniform in the OMPL bounds, steps N(0, 0.3^2)."""
    rng = np.random.default_rng(seed)
    T = rng.uniform(0.5, 2.0, (B, N_SEG))
    t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = rng.normal(0.0, 0.3, (B, N_SEG + 1, K_AX))
    wp[:, 0, :] = rng.uniform(BOUNDS_LO, BOUNDS_HI, (B, K_AX))
    np.cumsum(wp, axis=1, out=wp)
    return wp, t


def mesh_soups():
    from drone_path_planning_python_b200 import meshio
    out = []
    for name in (ROBOT, ENV):
        verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
        out.append(meshio.triangle_soup(verts, tris))
    return out


# --------------------------------------------------------------------------- CPU arm
# The reference's own CPU implementation of the path: the UNMODIFIED package
# src/optimizations of the reference (placed under the git-ignored oracle/_ref/ by
# __graft_entry__.build(), oracle/build_oracle.populate_ref) — calculate_trajectory1D per axis
# (calculatingTrajectories.py:37-197) and PiecewisePolynomial.eval per sample and axis
# (uav_trajectory.py:154-169) — plus the C restatement of the collision test (python-fcl is not
# installable here).  If oracle/_ref is absent the oracle port is timed instead and the
# line says kind = "port".
_CPU_STATE = {}


def _cpu_init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import build_oracle
    _CPU_STATE["soups"] = mesh_soups()
    _CPU_STATE["ref"] = build_oracle.import_ref()


def _cpu_work(args):
    from oracle import build_oracle, minsnap_oracle as mo
    wp, t = args
    robot, env = _CPU_STATE["soups"]
    ref = _CPU_STATE["ref"]
    hits = 0
    for b in range(wp.shape[0]):
        if ref is not None:
            opt, ct = ref
            pts = [opt.Point_time(opt.Waypoint(float(wp[b, i, 0]), float(wp[b, i, 1]), float(wp[b, i, 2]), 0.0),
                                  t=float(t[b, i])) for i in range(wp.shape[1])]
            totals = [ct.calculate_trajectory1D(pts, k)[1] for k in range(K_AX)]
            dt = sum(totals[0].time_durations) / S_SAMPLES
            pos = np.array([[float(np.asarray(tot.eval(s * dt)).reshape(())) for tot in totals]
                            for s in range(S_SAMPLES)])
        else:
            coef, dur = mo.solve_waypoints(wp[b], t[b])
            ts = mo.uniform_sample_times(dur, S_SAMPLES)
            pos = mo.sample_trajectory(coef, dur, ts)
        poses = np.concatenate([pos, np.zeros((S_SAMPLES, 1))], axis=1)
        hits += int(build_oracle.c_collide_poses(robot, env, poses).any())
    return hits


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_kind():
    return "reference" if os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "optimizations", "calculatingTrajectories.py")) \
        else "port"


class CpuArm:
    def __init__(self, cores):
        import multiprocessing as mp
        from oracle import build_oracle
        build_oracle.build()
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init)

    def run(self, wp, t):
        parts = np.array_split(np.arange(wp.shape[0]), self.cores)
        t0 = time.perf_counter()
        self.pool.map(_cpu_work, [(wp[p], t[p]) for p in parts if len(p)])
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(wp_np, t_np, per_core):
    """All-core and one-core figures of the CPU arm on the first trajectories of the batch."""
    cores = host_cores()
    kind = cpu_kind()
    nsamp = per_core * cores
    arm = CpuArm(cores)
    arm.run(wp_np[:cores], t_np[:cores])
    secs = arm.run(wp_np[:nsamp], t_np[:nsamp])
    arm.close()
    one = CpuArm(1)
    one.run(wp_np[:1], t_np[:1])
    n1 = max(4, 2 * per_core)
    secs1 = one.run(wp_np[:n1], t_np[:n1])
    one.close()
    what = ("unmodified reference package (oracle/_ref/optimizations: calculate_trajectory1D x %d axes, "
            "PiecewisePolynomial.eval x %d samples) + C restatement of the collision test" % (K_AX, S_SAMPLES)) \
        if kind == "reference" else "oracle port (numpy restatement) + C restatement of the collision test"
    return {"value": nsamp / secs, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %d trajectories of the same batch, %.1f s wall on %d processes; %s"
                      % (nsamp, secs, cores, what),
            "one_core": {"value": n1 / secs1, "cores": 1, "sample": "first %d trajectories, %.1f s" % (n1, secs1)}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = 8 * cores
    wp, t = make_workload(per_step * (args.steps + args.warmup), SEED)
    arm = CpuArm(cores)
    for w in range(args.warmup):
        sl = slice(w * per_step, (w + 1) * per_step)
        arm.run(wp[sl], t[sl])
    elapsed = 0.0
    for s in range(args.steps):
        sl = slice((args.warmup + s) * per_step, (args.warmup + s + 1) * per_step)
        elapsed += arm.run(wp[sl], t[sl])
    arm.close()
    value = per_step * args.steps / elapsed
    kind = cpu_kind()
    sample = "%d trajectories per step (8 per core) of the same generator; %s" % (
        per_step, "unmodified reference package from oracle/_ref + C collision restatement" if kind == "reference"
        else "oracle port + C collision restatement")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] shape: %d pieces x %d axes, S=%d, %s vs %s; bounded sample"
                               % (N_SEG, K_AX, S_SAMPLES, ROBOT, ENV), "trajectories_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- GPU arm
def load_profile_counters():
    """Per-trajectory counters of the pipeline's kernels from the tracked ncu capture
    (profiles/r2_kernel_counters.json, written by tools/ncu_summary.py from the .ncu-rep of this same
    command): DRAM bytes and executed FP64 instructions per kernel.  {} when the file is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_kernel_counters.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200 import _abi
    from drone_path_planning_python_b200.host_pipeline import HostPipeline

    # NCCL_DEBUG=VERSION makes NCCL print its banner on stdout, next to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = _abi.load()
    B = args.traj
    robot_soup, env_soup = mesh_soups()
    robot, env = mst.Mesh(robot_soup), mst.Mesh(env_soup)
    wp_np, t_np = make_workload(B, SEED + rank)
    wp_host = torch.from_numpy(wp_np).pin_memory()
    t_host = torch.from_numpy(t_np).pin_memory()
    wp = wp_host.to(dev)
    t = t_host.to(dev)

    def new_result(count):
        return mst.PipelineResult(torch.empty((count, N_SEG, K_AX, 8), dtype=torch.float64, device=dev),
                                  torch.empty((count, N_SEG), dtype=torch.float64, device=dev),
                                  torch.empty((count,), dtype=torch.int32, device=dev),
                                  torch.empty((count, S_SAMPLES), dtype=torch.uint8, device=dev),
                                  torch.empty((count,), dtype=torch.uint8, device=dev))
    res = new_result(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- multi-GPU data planes: what is gathered, and how -------------------------------
    # WHAT: "f32-wire" = what the reference produces, path_to_pol's float32 polynomial matrix + the flags
    # (default); "flags-only"; "f64-full" = FP64 coefficients + flags (round 1's figure).
    # HOW: "push" = the two-launch pipeline writes into this rank's slot of a symmetric buffer and the copy
    # engines push the slot into every peer's buffer chunk by chunk; "store" = the single-pass kernel
    # stores every finished trajectory through all peer pointers itself (mst_pipeline_wire).
    from drone_path_planning_python_b200.distributed import PeerPushAllGather, PeerStoreGather
    modes = {}
    if world > 1:
        if args.gather_chunks <= 0:
            args.gather_chunks = 4 if world <= 4 else 2
        if args.gather_chunks_flags <= 0:
            args.gather_chunks_flags = 1 if world <= 2 else (2 if world <= 4 else 4)
        mat_t = torch.empty((1, N_SEG, 1 + 8 * K_AX), dtype=torch.float32, device=dev)

        def push_mode(what):
            tmpl = {"f64-full": [res.coef, res.hit, res.any_hit], "f32-wire": [mat_t, res.hit, res.any_hit],
                    "flags-only": [res.hit, res.any_hit]}[what]
            # chunks of the compute / push overlap: few large copies for the FP64 payload (round 1: 2 chunks
            # at 8 GPUs), more for the smaller payloads, whose last chunk's push is what is left exposed
            chunks = {"f64-full": args.gather_chunks, "f32-wire": args.gather_chunks_f32,
                      "flags-only": args.gather_chunks_flags}[what]
            g = PeerPushAllGather(B, world, rank, chunks, tmpl, streams=args.push_streams, taper={"f64-full": False, "f32-wire": "head" if world >= 4 else "tail", "flags-only": "tail"}[what])
            nb = len(g.buffers)

            def chunk(lo, hi):
                coef = g.local_slot(0, lo, hi) if what == "f64-full" else res.coef[lo:hi]
                view = mst.PipelineResult(coef, res.dur[lo:hi], res.info[lo:hi], g.local_slot(nb - 2, lo, hi),
                                          g.local_slot(nb - 1, lo, hi))
                # the float32 matrix is written by the solver kernel straight into this rank's slot
                mst.pipeline(wp[lo:hi], t[lo:hi], S_SAMPLES, robot, env, out=view,
                             pol_matrix=g.local_slot(0, lo, hi) if what == "f32-wire" else None)
            return g, (lambda: g.run(chunk))

        def store_mode(what):
            g = PeerStoreGather(B, world, rank, N_SEG, K_AX, S_SAMPLES, dev,
                                mode="pol_matrix_f32" if what == "f32-wire" else "flags")
            return g, (lambda: g.run(lambda wire: mst.pipeline_wire(wp, t, S_SAMPLES, robot, env, wire, out=res)))
        per_traj = {"f64-full": N_SEG * K_AX * 64 + S_SAMPLES + 1, "f32-wire": N_SEG * (1 + 8 * K_AX) * 4 + S_SAMPLES + 1,
                    "flags-only": S_SAMPLES + 1}
        for what, how in (("f32-wire", "push"), ("f32-wire", "store"), ("flags-only", "push"), ("flags-only", "store"),
                          ("f64-full", "push")):
            g, fn = push_mode(what) if how == "push" else store_mode(what)
            modes["%s/%s" % (what, how)] = {"gather": g, "step": fn, "bytes": per_traj[what]}
    if args.gather == "f64-full":
        args.gather_how = "push"        # the single-pass kernel's wire outputs are the float32 matrix and the flags
    default_mode = None if world == 1 else "%s/%s" % (args.gather, args.gather_how)

    def make_step(mode):
        if world == 1:
            return lambda: mst.pipeline(wp, t, S_SAMPLES, robot, env, out=res)
        return modes[mode]["step"]

    step = make_step(default_mode)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms = timed(step, args.steps)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    bad = int((res.info != 0).sum().item())
    if world > 1:   # the flags of this rank live in its slot of the gather buffer
        g = modes[default_mode]["gather"]
        own_any = (g.buffers["any_hit"] if isinstance(g.buffers, dict) else g.buffers[-1])[rank * B:(rank + 1) * B]
        hit_rate = float(own_any.float().mean().item())
    else:
        hit_rate = float(res.any_hit.float().mean().item())

    # ---- N > 1: the other gather modes, strong scaling, and a check of the gathered data -------------
    scaling_modes, strong, gather_ok = None, None, None
    if world > 1:
        side_steps = max(3, args.steps // 4)
        scaling_modes = {}
        for mode, m in modes.items():
            if mode == default_mode:
                ms = ms_per_step
            else:
                for _ in range(2):
                    m["step"]()
                ms = timed(m["step"], side_steps) / side_steps
            scaling_modes[mode] = {"ms_per_step": ms, "value": world * B / (ms * 1e-3),
                                   "gathered_bytes_per_trajectory": m["bytes"],
                                   "nvlink_in_gbs_per_gpu": (world - 1) * B * m["bytes"] / (ms * 1e-3) / 1e9}
        # out of the timed region: every rank's slot of this rank's gathered buffers must hold that
        # rank's results (checksums of the local results, exchanged with NCCL), for both data planes
        def sums(mat, hit, any_hit):
            return torch.stack([mat.view(torch.int32).to(torch.int64).sum(), hit.to(torch.int64).sum(),
                                (hit.to(torch.int64) * torch.arange(1, S_SAMPLES + 1, device=dev)).sum(),
                                any_hit.to(torch.int64).sum()])
        gather_ok = True
        for mode in ("f32-wire/push", "f32-wire/store"):
            g = modes[mode]["gather"]
            modes[mode]["step"]()
            torch.cuda.synchronize()
            bufs = [g.buffers["pol_matrix"], g.buffers["hit"], g.buffers["any_hit"]] if mode.endswith("store") else g.buffers
            own = slice(rank * B, (rank + 1) * B)
            mine = sums(bufs[0][own], bufs[1][own], bufs[2][own])
            ref = sums(mst.pack_pol_matrix(res.coef, res.dur), bufs[1][own], bufs[2][own])
            gather_ok = gather_ok and bool(torch.equal(mine, ref))
            allsums = torch.empty((world, 4), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allsums, mine.view(1, 4))
            for r in range(world):
                sl = slice(r * B, (r + 1) * B)
                gather_ok = gather_ok and bool(torch.equal(sums(bufs[0][sl], bufs[1][sl], bufs[2][sl]), allsums[r]))
        flag = torch.tensor([0 if gather_ok else 1], device=dev)
        dist.all_reduce(flag)
        gather_ok = int(flag.item()) == 0
        # strong scaling of configs[4]: 1,048,576 trajectories in total, sharded over the ranks
        Bs = TRAJ_PER_GPU // world
        res_s = new_result(Bs)
        mat_s = torch.empty((1, N_SEG, 1 + 8 * K_AX), dtype=torch.float32, device=dev)
        gs = PeerPushAllGather(Bs, world, rank, 2, [mat_s, res_s.hit, res_s.any_hit], streams=args.push_streams)

        def strong_chunk(lo, hi):
            view = mst.PipelineResult(res_s.coef[lo:hi], res_s.dur[lo:hi], res_s.info[lo:hi], gs.local_slot(1, lo, hi),
                                      gs.local_slot(2, lo, hi))
            mst.pipeline(wp[lo:hi], t[lo:hi], S_SAMPLES, robot, env, out=view, pol_matrix=gs.local_slot(0, lo, hi))
        fn = lambda: gs.run(strong_chunk)
        for _ in range(3):
            fn()
        ms = timed(fn, args.steps) / args.steps
        strong = {"total_trajectories": Bs * world, "ms_per_step": ms, "value": Bs * world / (ms * 1e-3),
                  "gather": "f32-wire/push"}

    # ---- the kernels of the step alone, timed live with CUDA events on the launching stream -----------
    stage_ms = {}
    ws = torch.empty((max(1, lib.mst_pipeline_workspace_bytes(B, N_SEG, K_AX, 1, S_SAMPLES)),), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def pipeline_call(solver):
        _abi.check(lib.mst_pipeline(wp.data_ptr(), t.data_ptr(), B, N_SEG, K_AX, 1, solver, S_SAMPLES, robot.handle,
                                    env.handle, res.coef.data_ptr(), res.dur.data_ptr(), res.info.data_ptr(),
                                    res.hit.data_ptr(), res.any_hit.data_ptr(), ws.data_ptr(), st), "mst_pipeline")

    def pipeline_only():
        pipeline_call(_abi.SOLVER_AUTO)

    def onepass_only():
        pipeline_call(_abi.SOLVER_AUTO_ONE_PASS)

    def stage_call(stage):
        _abi.check(lib.mst_pipeline_stage(stage, wp.data_ptr(), t.data_ptr(), B, N_SEG, K_AX, 1, _abi.SOLVER_AUTO, S_SAMPLES,
                                          robot.handle, env.handle, res.coef.data_ptr(), res.dur.data_ptr(),
                                          res.info.data_ptr(), res.hit.data_ptr(), res.any_hit.data_ptr(), ws.data_ptr(),
                                          st), "mst_pipeline_stage")

    def solve_only():      # the pipeline's solver launches (far-piece bounds included)
        stage_call(1)

    def collide_only():    # the pipeline's sampling / collision launch
        stage_call(2)

    def plain_collide():   # round 1's kernel: every sample evaluated (mst_collide_trajectories)
        _abi.check(lib.mst_collide_trajectories(res.coef.data_ptr(), res.dur.data_ptr(), B, N_SEG, K_AX, S_SAMPLES,
                                                robot.handle, env.handle, res.hit.data_ptr(), res.any_hit.data_ptr(),
                                                st), "mst_collide_trajectories")
    pipeline_only(); onepass_only(); solve_only(); collide_only(); plain_collide()
    side = max(3, args.steps // 2)
    stage_ms["pipeline, two launches (default): condensed_cols_kernel + sample_collide_cull_kernel"] = timed(pipeline_only, side) / side
    stage_ms["condensed_cols_kernel<cull> (+ banded_lu_kernel on the empty declined list)"] = timed(solve_only, side) / side
    stage_ms["sample_collide_cull_kernel"] = timed(collide_only, side) / side
    stage_ms["sample_collide_kernel (no culling, mst_collide_trajectories)"] = timed(plain_collide, side) / side
    stage_ms["pipeline, single pass (MST_SOLVER_AUTO_ONE_PASS): onepass_kernel"] = timed(onepass_only, side) / side

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # dominant kernel: sample_collide_kernel.  Its compulsory bytes per trajectory: coefficients and
    # durations in, per-sample flags + any-flag out; the whole step is reported beside it
    # against SURVEY §8d's ALG_BYTES.
    dom_ms = stage_ms["sample_collide_cull_kernel"]
    step_ms = stage_ms["pipeline, two launches (default): condensed_cols_kernel + sample_collide_cull_kernel"]
    one_ms = stage_ms["pipeline, single pass (MST_SOLVER_AUTO_ONE_PASS): onepass_kernel"]
    dom_bytes = N_SEG * K_AX * 64 + N_SEG * 8 + S_SAMPLES + 1
    achieved = B * dom_bytes / (dom_ms * 1e-3) / 1e9
    counters = load_profile_counters()
    c_dom, c_sol, c_one = (counters.get(k) for k in ("sample_collide_cull_kernel", "condensed_cols_kernel", "onepass_kernel"))
    traffic = c_dom["dram_bytes_per_trajectory"] * B if c_dom else None
    # FP64 roof: measured in this run (tools/fp64_peak.py)
    fp64 = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import fp64_peak
        fp64 = fp64_peak.measure()
    except Exception as exc:  # the probe library did not travel
        sys.stderr.write("fp64 peak probe unavailable: %r\n" % (exc,))
    traj_per_s = B / (step_ms * 1e-3)
    fp64_block = None
    if fp64:
        peak_tf = fp64["fp64_fma_tflops"]
        fp64_block = {"fp64_peak_tflops": peak_tf, "fp64_peak_source": "measured", "fp64_peak_how": fp64["how"],
                      "fp64_mul_add_tflops": fp64["fp64_mul_add_tflops"], "sm_mhz_after_probe": fp64.get("sm_mhz_after"),
                      # SURVEY §8d's credited flops (banded-LU formulation of the reference's own system)
                      "credited_flops_per_trajectory": 45600,
                      "fp64_frac_credited": traj_per_s * 45600 / (peak_tf * 1e12)}
        if c_dom and c_sol and c_dom.get("fp64_flops_per_trajectory") and c_sol.get("fp64_flops_per_trajectory"):
            ex = c_dom["fp64_flops_per_trajectory"] + c_sol["fp64_flops_per_trajectory"]
            ins = c_dom["fp64_instructions_per_trajectory"] + c_sol["fp64_instructions_per_trajectory"]
            fp64_block.update({"executed_flops_per_trajectory": ex, "fp64_frac": traj_per_s * ex / (peak_tf * 1e12),
                               "executed_fp64_instructions_per_trajectory": ins,
                               "fp64_pipe_frac": traj_per_s * ins / fp64["fma"]["thread_instructions_per_s"],
                               "fp64_note": "whole step (both kernels), executed flops from the tracked ncu capture"})

    # ---- end to end through HOST buffers (pinned), copies inside the timed region ----------
    hp = HostPipeline(N_SEG, K_AX, S_SAMPLES, robot, env, chunk=args.e2e_chunk)
    host_out = HostPipeline.alloc_host_result(B, N_SEG, K_AX, S_SAMPLES)

    def e2e_step():
        hp.run(wp_host, t_host, host_out)
    for _ in range(min(args.warmup, 3)):
        e2e_step()
    e2e_steps = max(3, args.steps // 4)
    e2e_ms = timed(e2e_step, e2e_steps) / e2e_steps
    h2d, d2h = hp.bytes_per_trajectory()
    assert np.array_equal(host_out.any_hit.numpy(), res.any_hit.cpu().numpy())
    # same call, results in the reference's own output format (float32 polynomial matrix)
    del hp, host_out
    hp32 = HostPipeline(N_SEG, K_AX, S_SAMPLES, robot, env, chunk=args.e2e_chunk, wire="pol_matrix_f32")
    host_out32 = HostPipeline.alloc_host_result(B, N_SEG, K_AX, S_SAMPLES, wire="pol_matrix_f32")

    def e2e32_step():
        hp32.run(wp_host, t_host, host_out32)
    for _ in range(min(args.warmup, 3)):
        e2e32_step()
    e2e32_ms = timed(e2e32_step, e2e_steps) / e2e_steps
    h2d32, d2h32 = hp32.bytes_per_trajectory()

    launches_per_step = lib.mst_pipeline_launch_count(B, N_SEG, K_AX, 1, _abi.SOLVER_AUTO, S_SAMPLES)
    if world > 1:
        nch = {"f64-full": args.gather_chunks, "f32-wire": args.gather_chunks_f32, "flags-only": args.gather_chunks_flags}[args.gather]
        launches_per_step = launches_per_step * nch + (nch if args.gather == "f32-wire" else 0) \
            if args.gather_how == "push" else launches_per_step + 1   # list-mode matrix rows per chunk / wire patch kernel
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "hbm_frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "kernel": "sample_collide_cull_kernel<3> (%.0f %% of the two-launch step)" % (100 * dom_ms / step_ms),
                "alg_bytes_per_trajectory": dom_bytes, "trajectories_per_launch": B, "kernel_ms": dom_ms,
                "kernels_ms": stage_ms,
                "traffic_source": "profiles/r2_kernel_counters.json (ncu --set full of this command)" if c_dom else None,
                "whole_step": {"alg_bytes_per_trajectory": ALG_BYTES, "ms": step_ms,
                               "achieved_gbs": B * ALG_BYTES / (step_ms * 1e-3) / 1e9,
                               "hbm_frac": B * ALG_BYTES / (step_ms * 1e-3) / 1e9 / peak_gbs,
                               "traffic": (c_dom["dram_bytes_per_trajectory"] + c_sol["dram_bytes_per_trajectory"]) * B
                               if c_dom and c_sol else None},
                "single_pass": {"kernel": "onepass_kernel<3>", "ms": one_ms, "alg_bytes_per_trajectory": ALG_BYTES,
                                "achieved_gbs": B * ALG_BYTES / (one_ms * 1e-3) / 1e9,
                                "hbm_frac": B * ALG_BYTES / (one_ms * 1e-3) / 1e9 / peak_gbs,
                                "traffic": c_one["dram_bytes_per_trajectory"] * B if c_one else None,
                                "note": "coefficients reach HBM once; slower than the two launches on B200 (10 resident "
                                        "warps per SM against 14-16), so not the default: profiles/r2_onepass_history.md"},
                "note": "instruction/latency bound (branchy FP64 geometry, short recurrences), not bandwidth bound: "
                        "both fractions are reported, see profiles/"}
    if fp64_block:
        roofline.update(fp64_block)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4]: %d trajectories/GPU x %d pieces x %d axes, own time vector "
                               "each (T~U(0.5,2)s), S=%d samples, %s vs %s, seed %d"
                               % (B, N_SEG, K_AX, S_SAMPLES, ROBOT, ENV, SEED),
                   "trajectories_per_gpu": B, "l2": "inputs+outputs %.2f GB per step >> 126 MB L2 (no flush needed)"
                   % (B * (ALG_BYTES + N_SEG * 8 + 4) / 1e9),
                   "gather": None if world == 1 else
                   "%s: %s" % (default_mode, "two-launch pipeline into a symmetric buffer, copy-engine pushes to every peer, "
                               "%d chunks" % {"f64-full": args.gather_chunks, "f32-wire": args.gather_chunks_f32,
                                                "flags-only": args.gather_chunks_flags}[args.gather]
                               if args.gather_how == "push" else
                               "single-pass kernel storing through all peer pointers (mst_pipeline_wire)"),
                   "solver": "auto (condensed LDL^T; banded pivoted LU for wide duration spreads)"},
        "clocks": clocks,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d * B), "d2h_bytes_per_step": int(d2h * B),
                "chunk": args.e2e_chunk, "note": "pinned host in/out (FP64 coefficients), 3-slot copy/compute overlap, "
                "host waits for the last device->host copy; PCIe-bound by the device->host copy",
                "pol_matrix_f32_wire": {"value": world * B / (e2e32_ms * 1e-3), "ms_per_step": e2e32_ms,
                                        "d2h_bytes_per_step": int(d2h32 * B),
                                        "note": "same call returning path_to_pol's float32 (n,33)-style matrix"}},
        "gpu_launches": args.steps * launches_per_step,
        "roofline": roofline,
        "checks": {"solver_failures": bad, "any_hit_rate": hit_rate, "gathered_buffers_verified": gather_ok},
    }
    if scaling_modes:
        line["gather_modes"] = scaling_modes
        line["strong_scaling"] = strong

    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wp_np, t_np, 4 if args.quick else 128)   # ~15-20 s of CPU work in all
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--traj", type=int, default=TRAJ_PER_GPU, help="trajectories per GPU per step")
    ap.add_argument("--gather-chunks", type=int, default=0,
                    help="chunks of the overlapped all-gather (0: 4 up to 4 GPUs where the kernels still matter, "
                         "2 at 8, where few large NVLink copies win; profiles/r1_scaling.md)")
    ap.add_argument("--gather", choices=["f32-wire", "flags-only", "f64-full"], default="f32-wire",
                    help="what the timed step gathers at N > 1 (the other modes are timed beside it)")
    ap.add_argument("--gather-chunks-f32", type=int, default=3, help="tapered: 4 : 2 : 1 at 2 GPUs (compute-bound), 1 : 2 : 4 from 4 GPUs (exchange-bound)")
    ap.add_argument("--gather-chunks-flags", type=int, default=0,
                    help="tapered (.. 4 : 2 : 1); 0: 1 chunk at 2 GPUs, 2 at 4, 4 at 8 (a chunk costs ~0.1 ms of launch "
                         "tails; the flags of one peer take 0.14 ms of NVLink time)")
    ap.add_argument("--gather-how", choices=["push", "store"], default="push",
                    help="copy-engine push behind the two-launch pipeline, or peer stores from the single-pass kernel")
    ap.add_argument("--push-streams", type=int, default=1)
    ap.add_argument("--e2e-chunk", type=int, default=1 << 16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="tiny CPU sample (smoke runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
