"""GPU tier: the host-buffer pipeline (pinned in / pinned out, chunked, copy/compute overlap)
returns exactly what the device-resident pipeline returns, in both wire formats."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
    return meshio.triangle_soup(verts, tris)


@pytest.mark.parametrize("K,G", [(3, 1), (4, 1), (3, 5)])
def test_host_pipeline_equals_device_pipeline(K, G):
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200.host_pipeline import HostPipeline
    from oracle import minsnap_oracle as mo
    rng = np.random.default_rng(10 * K + G)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    B, n, S = 10000 * G, 10, 100
    T = rng.uniform(0.5, 2.0, (B // G, n))
    t = np.concatenate([np.zeros((B // G, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = np.zeros((B, n + 1, K))
    wp[:, :, :3] = rng.uniform([-2.2, 2.8, 0.5], [2.2, 5.0, 2.5], (B, 1, 3)) + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), axis=1)
    if K == 4:
        wp[:, :, 3] = np.cumsum(rng.normal(0, 0.1, (B, n + 1)), axis=1)
    ref = mst.pipeline(wp, t, S, robot, env, share_time_group=G)
    wp_h, t_h = torch.from_numpy(wp).pin_memory(), torch.from_numpy(t).pin_memory()
    hp = HostPipeline(n, K, S, robot, env, chunk=3000 * G + 1, share_time_group=G)   # odd chunk: rounding to G, ragged tail
    out = hp.run(wp_h, t_h)
    torch.cuda.synchronize()
    assert torch.equal(out.coef, ref.coef.cpu()) and torch.equal(out.dur, ref.dur.cpu())
    assert torch.equal(out.hit, ref.hit.cpu()) and torch.equal(out.any_hit, ref.any_hit.cpu())
    assert torch.equal(out.info, ref.info.cpu())
    hp32 = HostPipeline(n, K, S, robot, env, chunk=4096 * G, share_time_group=G, wire="pol_matrix_f32")
    out32 = hp32.run(wp_h, t_h)
    torch.cuda.synchronize()
    assert out32.coef.shape == (B, n, 1 + 8 * K) and out32.coef.dtype == torch.float32 and out32.dur is None
    assert torch.equal(out32.hit, ref.hit.cpu())
    for b in (0, B - 1):
        assert np.array_equal(out32.coef[b].numpy(), mo.pack_pol_matrix(ref.coef[b].cpu().numpy(), ref.dur[b].cpu().numpy()))
    h2d, d2h = hp32.bytes_per_trajectory()
    assert d2h == n * (1 + 8 * K) * 4 + 4 + S + 1
