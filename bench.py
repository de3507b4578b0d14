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
runs the same per-GPU batch (weak scaling), and the final coefficients and collision flags
are all-gathered over NCCL, chunk by chunk, overlapped with the next chunk's kernels.

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
_CPU_STATE = {}


def _cpu_init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import build_oracle  # noqa: F401
    _CPU_STATE["soups"] = mesh_soups()


def _cpu_work(args):
    """Reference algorithm for a slice of trajectories: per-axis dense assembly + solve,
    Python Horner sampling (oracle/minsnap_oracle.py), C SAT collision (oracle/collision_oracle.c)."""
    from oracle import build_oracle, minsnap_oracle as mo
    wp, t = args
    robot, env = _CPU_STATE["soups"]
    hits = 0
    for b in range(wp.shape[0]):
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = 16 * cores
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
    sample = "%d trajectories per step (16 per core) of the same generator" % per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] shape: %d pieces x %d axes, S=%d, %s vs %s; bounded sample"
                               % (N_SEG, K_AX, S_SAMPLES, ROBOT, ENV), "trajectories_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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

    res = mst.PipelineResult(torch.empty((B, N_SEG, K_AX, 8), dtype=torch.float64, device=dev),
                             torch.empty((B, N_SEG), dtype=torch.float64, device=dev),
                             torch.empty((B,), dtype=torch.int32, device=dev),
                             torch.empty((B, S_SAMPLES), dtype=torch.uint8, device=dev),
                             torch.empty((B,), dtype=torch.uint8, device=dev))

    # multi-GPU: all-gather of coefficients + flags, chunked and overlapped with the next chunk.
    # Preferred: peer push over NVLink with the copy engines (distributed.PeerPushAllGather);
    # fallback: NCCL all_gather_into_tensor (distributed.ChunkedAllGather).
    from drone_path_planning_python_b200.distributed import ChunkedAllGather, PeerPushAllGather
    gather, gather_kind, n_chunks = None, None, 1
    if args.gather_chunks <= 0:
        args.gather_chunks = 4 if world <= 4 else 2
    if world > 1:
        if args.gather == "push":
            try:
                gather = PeerPushAllGather(B, world, rank, args.gather_chunks, [res.coef, res.hit, res.any_hit],
                                           streams=args.push_streams)
                gather_kind = "peer push over NVLink (copy engines, symmetric memory)"
                res = mst.PipelineResult(gather.local_slot(0), res.dur, res.info, gather.local_slot(1),
                                         gather.local_slot(2))
            except Exception as exc:  # symmetric memory unavailable on this box
                sys.stderr.write("peer-push gather unavailable (%r); using NCCL\n" % (exc,))
                gather = None
        if gather is None:
            gather = ChunkedAllGather(B, world, args.gather_chunks, [res.coef, res.hit, res.any_hit])
            gather_kind = "NCCL all_gather_into_tensor"
        n_chunks = len(gather.plan)

    def chunk_view(lo, hi):
        return mst.PipelineResult(res.coef[lo:hi], res.dur[lo:hi], res.info[lo:hi], res.hit[lo:hi], res.any_hit[lo:hi])

    def compute_chunk(lo, hi):
        view = chunk_view(lo, hi)
        mst.pipeline(wp[lo:hi], t[lo:hi], S_SAMPLES, robot, env, out=view)
        return view.coef, view.hit, view.any_hit

    def step():
        if world == 1:
            mst.pipeline(wp, t, S_SAMPLES, robot, env, out=res)
        elif isinstance(gather, PeerPushAllGather):
            gather.run(compute_chunk)
        else:
            gather.run(compute_chunk, assemble=False)   # results stay in [chunk][rank][...] staging buffers

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

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms = timed(step, args.steps)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    bad = int((res.info != 0).sum().item())
    hit_rate = float(res.any_hit.float().mean().item())

    # --- the two kernels of the step alone, timed live with CUDA events on the launching stream
    stage_ms = {}
    ws = torch.empty((max(1, lib.mst_solve_workspace_bytes(B, N_SEG, K_AX, 1)),), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def solve_only():
        _abi.check(lib.mst_solve_batch(wp.data_ptr(), t.data_ptr(), B, N_SEG, K_AX, 1, _abi.SOLVER_AUTO,
                                       res.coef.data_ptr(), res.dur.data_ptr(), res.info.data_ptr(), ws.data_ptr(),
                                       st), "mst_solve_batch")

    def collide_only():
        _abi.check(lib.mst_collide_trajectories(res.coef.data_ptr(), res.dur.data_ptr(), B, N_SEG, K_AX, S_SAMPLES,
                                                robot.handle, env.handle, res.hit.data_ptr(), res.any_hit.data_ptr(),
                                                st), "mst_collide_trajectories")
    solve_only(); collide_only()
    stage_ms["condensed_kernel (+ banded_lu_kernel on the declined list)"] = timed(solve_only, args.steps) / args.steps
    stage_ms["sample_collide_kernel"] = timed(collide_only, args.steps) / args.steps

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # dominant kernel: sample_collide_kernel.  Its compulsory bytes per trajectory: coefficients
    # and durations in, per-sample flags + any-flag out.
    dom_ms = stage_ms["sample_collide_kernel"]
    dom_bytes = N_SEG * K_AX * 64 + N_SEG * 8 + S_SAMPLES + 1
    achieved = B * dom_bytes / (dom_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from profiles/ (ncu --set full,
    # 262,144 trajectories per launch, cold L2), scaled to this launch's trajectory count
    traffic_per_traj = (541.113344e6 + 30.006528e6) / 262144

    # --- end to end through HOST buffers (pinned), copies inside the timed region ----------
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[4]: %d trajectories/GPU x %d pieces x %d axes, own time vector "
                               "each (T~U(0.5,2)s), S=%d samples, %s vs %s, seed %d"
                               % (B, N_SEG, K_AX, S_SAMPLES, ROBOT, ENV, SEED),
                   "trajectories_per_gpu": B, "l2": "inputs+outputs %.2f GB per step >> 126 MB L2 (no flush needed)"
                   % (B * (ALG_BYTES + N_SEG * 8 + 4) / 1e9),
                   "gather": None if world == 1 else "%s of coef(f64)+hit+any_hit, %d chunks" % (gather_kind, n_chunks),
                   "solver": "auto (condensed LDL^T; banded pivoted LU for wide duration spreads)"},
        "clocks": clocks,
        "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d * B), "d2h_bytes_per_step": int(d2h * B),
                "chunk": args.e2e_chunk, "note": "pinned host in/out (FP64 coefficients), 3-slot copy/compute overlap; "
                "PCIe-bound by the device->host copy",
                "pol_matrix_f32_wire": {"value": world * B / (e2e32_ms * 1e-3), "ms_per_step": e2e32_ms,
                                        "d2h_bytes_per_step": int(d2h32 * B),
                                        "note": "same call returning path_to_pol's float32 (n,33)-style matrix"}},
        "gpu_launches": args.steps * n_chunks * lib.mst_pipeline_launch_count(B // n_chunks, N_SEG, K_AX, 1, _abi.SOLVER_AUTO, S_SAMPLES),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic_per_traj * B,
                     "peak_source": peak_src, "kernel": "sample_collide_kernel<3> (%.0f %% of the step)"
                     % (100 * dom_ms / ms_per_step if world == 1 else 100 * dom_ms / sum(stage_ms.values())),
                     "alg_bytes_per_trajectory": dom_bytes, "trajectories_per_launch": B,
                     "kernel_ms": dom_ms, "kernels_ms": stage_ms,
                     "whole_step": {"alg_bytes_per_trajectory": ALG_BYTES,
                                    "achieved_gbs": (B * ALG_BYTES / (ms_per_step * 1e-3) / 1e9) if world == 1 else None},
                     "note": "instruction/latency bound (branchy FP64 geometry), not bandwidth bound: see profiles/"},
        "checks": {"solver_failures": bad, "any_hit_rate": hit_rate},
    }

    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        nsamp = 256 * cores if not args.quick else 16 * cores
        arm = CpuArm(cores)
        arm.run(wp_np[:cores], t_np[:cores])
        secs = arm.run(wp_np[:nsamp], t_np[:nsamp])
        arm.close()
        line["cpu_baseline"] = {"value": nsamp / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d trajectories of the same batch, %.1f s wall on %d processes"
                                          % (nsamp, secs, cores)}
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
    ap.add_argument("--gather", choices=["push", "nccl"], default="push")
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
