"""TEST SCRIPT (run under torchrun, one rank per GPU): every multi-GPU data plane of the benchmark
— peer push with the copy engines (FP64 coefficients), peer stores from inside the single-pass
kernel (float32 polynomial matrix + flags; flags only) and the NCCL all-gather — must leave, on
EVERY rank, the concatenation of all ranks' results in global trajectory order.  Each rank
recomputes the other ranks' shards locally (same seeds) and compares its gathered buffers bit for
bit.  Exit code 0 = all ranks agree."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import bench
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200.distributed import ChunkedAllGather, PeerPushAllGather, PeerStoreGather
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, n, K, S = int(os.environ.get("CHECK_TRAJ", "30000")), bench.N_SEG, bench.K_AX, bench.S_SAMPLES
    robot_soup, env_soup = bench.mesh_soups()
    robot, env = mst.Mesh(robot_soup), mst.Mesh(env_soup)

    def shard(r):
        wp, t = bench.make_workload(B, bench.SEED + r)
        if r % 2 == 1:      # some groups for the pivoted solver too (wire patch path)
            t[::97, 1:] = np.cumsum(np.clip((t[::97, 1:] - t[::97, :-1]) * np.exp(np.linspace(-2, 2, n)), 0.05, 5.0), axis=1)
        return torch.as_tensor(wp, device=dev), torch.as_tensor(t, device=dev)

    expected = [mst.pipeline(*shard(r), S, robot, env) for r in range(world)]
    exp_coef = torch.cat([e.coef for e in expected])
    exp_hit = torch.cat([e.hit for e in expected])
    exp_any = torch.cat([e.any_hit for e in expected])
    exp_mat = torch.cat([mst.pack_pol_matrix(e.coef, e.dur) for e in expected])
    wp, t = shard(rank)
    failures = []

    def check(name, got, want):
        same = torch.equal(got.view(torch.int64) if got.dtype == torch.float64 else got,
                           want.view(torch.int64) if want.dtype == torch.float64 else want)
        if not same:
            failures.append(name)

    # (a) copy-engine peer push of the FP64 coefficients + flags
    tmpl = [expected[rank].coef, expected[rank].hit, expected[rank].any_hit]
    push = PeerPushAllGather(B, world, rank, 3, tmpl)
    dur = torch.empty((B, n), dtype=torch.float64, device=dev)
    info = torch.empty((B,), dtype=torch.int32, device=dev)

    def compute_chunk(lo, hi):
        view = mst.PipelineResult(push.local_slot(0, lo, hi), dur[lo:hi], info[lo:hi], push.local_slot(1, lo, hi),
                                  push.local_slot(2, lo, hi))
        mst.pipeline(wp[lo:hi], t[lo:hi], S, robot, env, out=view)
    for _ in range(2):
        out = push.run(compute_chunk)
    torch.cuda.synchronize()
    check("push.coef", out[0], exp_coef); check("push.hit", out[1], exp_hit); check("push.any", out[2], exp_any)

    # (b) peer stores from inside the kernel: float32 polynomial matrix + flags, then flags only
    for mode in ("pol_matrix_f32", "flags"):
        store = PeerStoreGather(B, world, rank, n, K, S, dev, mode=mode)
        for buf in store.buffers.values():
            buf.fill_(77)
        dist.barrier()
        for _ in range(2):
            bufs = store.run(lambda wire: mst.pipeline_wire(wp, t, S, robot, env, wire))
        torch.cuda.synchronize()
        if mode == "pol_matrix_f32":
            check("store.mat", bufs["pol_matrix"], exp_mat)
        check("store.hit." + mode, bufs["hit"], exp_hit); check("store.any." + mode, bufs["any_hit"], exp_any)

    # (c) NCCL all-gather
    nccl = ChunkedAllGather(B, world, 3, tmpl)

    def compute_chunk2(lo, hi):
        r = mst.pipeline(wp[lo:hi], t[lo:hi], S, robot, env)
        return r.coef, r.hit, r.any_hit
    coef_g, hit_g, any_g = nccl.run(compute_chunk2)
    torch.cuda.synchronize()
    check("nccl.coef", coef_g, exp_coef); check("nccl.hit", hit_g, exp_hit); check("nccl.any", any_g, exp_any)

    flag = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(flag)
    print("rank %d: %s" % (rank, "OK" if not failures else "MISMATCH " + ",".join(failures)), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
