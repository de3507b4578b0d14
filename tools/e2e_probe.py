"""Where does the host-buffer pipeline's time go?  Chunk size / slot count sweep, and the same
loop with the kernels removed (copies only)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import drone_path_planning_python_b200 as mst
from drone_path_planning_python_b200 import host_pipeline as hpmod
from drone_path_planning_python_b200.host_pipeline import HostPipeline

B = 1 << 20
wp, t = bench.make_workload(B, 1)
wp_h = torch.from_numpy(wp).pin_memory(); t_h = torch.from_numpy(t).pin_memory()
rs, es = bench.mesh_soups()
robot, env = mst.Mesh(rs), mst.Mesh(es)

def measure(chunk, slots, wire, copy_only=False):
    hp = HostPipeline(bench.N_SEG, bench.K_AX, bench.S_SAMPLES, robot, env, chunk=chunk, slots=slots, wire=wire)
    out = HostPipeline.alloc_host_result(B, bench.N_SEG, bench.K_AX, bench.S_SAMPLES, wire=wire)
    real = hpmod.pipeline
    if copy_only:
        hpmod.pipeline = lambda *a, **k: None
    try:
        for _ in range(2): hp.run(wp_h, t_h, out)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4): hp.run(wp_h, t_h, out)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 4 * 1e3
    finally:
        hpmod.pipeline = real
    h2d, d2h = hp.bytes_per_trajectory()
    print("chunk %7d slots %d wire %-14s copy_only %d : %6.2f ms  %5.1f M/s  d2h %.1f GB/s" % (
        chunk, slots, wire, copy_only, ms, B / ms / 1e3, B * d2h / ms / 1e6), flush=True)

for wire in ("f64", "pol_matrix_f32"):
    measure(1 << 16, 3, wire)
    measure(1 << 16, 3, wire, copy_only=True)
    for chunk in (1 << 14, 1 << 15, 1 << 17, 1 << 18):
        measure(chunk, 3, wire)
    measure(1 << 16, 2, wire); measure(1 << 16, 4, wire)
