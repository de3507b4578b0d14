"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run as ``python oracle/make_golden.py`` in the build
container, where the reference checkout is mounted read-only at
``/root/reference``.  It

1. imports the reference's ``optimizations`` package as is (matplotlib /
   mpl_toolkits are absent from the image and only used for plotting, so empty
   stand-in modules are registered in ``sys.modules`` first — no reference file is
   touched);
2. runs the reference on seeded inputs and on its own shipped fixtures and stores
   inputs + the reference's outputs as small ``.npz`` files under ``tests/golden``;
3. asserts that ``oracle/minsnap_oracle.py`` reproduces every one of those outputs
   bit for bit (that is what "parity pinned" means for the trajectory half);
4. stores the shipped data fixtures (polynomial-matrix CSVs, ``src/traj.csv``, the
   STL obstacle/robot meshes) as arrays so the GPU box — which has no
   ``/root/reference`` — can run the parity tests and the benchmark.

Nothing here is imported by the product.
"""
from __future__ import annotations

import glob
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "drone_path_planning_python_b200", "data")
sys.path.insert(0, ROOT)


def import_reference():
    warnings.filterwarnings("ignore")
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(REF, "src"))
    import optimizations  # noqa: F401
    from optimizations import calculatingTrajectories
    return optimizations, calculatingTrajectories


def ref_solve(opt, ct, wp, t):
    """Reference coefficients ``[n, K, 8]`` for waypoints ``wp[m, K<=4]``."""
    m, K = wp.shape
    pts = []
    for i in range(m):
        vals = [float(v) for v in wp[i]] + [0.0] * (4 - K)
        pts.append(opt.Point_time(opt.Waypoint(*vals), t=float(t[i])))
    out = np.zeros((m - 1, K, 8))
    durs = None
    for k in range(K):
        pieces, total = ct.calculate_trajectory1D(pts, k)
        for i, piece in enumerate(pieces):
            out[i, k] = np.asarray(piece.p).reshape(8)
        durs = np.asarray(total.time_durations, dtype=np.float64)
    return out, durs


def main():
    from oracle import minsnap_oracle as mo

    opt, ct = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(DATA, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- a1: time-power rows ------------------------------------------------
    ts = np.array([0.0, 0.2, 0.2040816326530612, 0.5, 1.0, 1.37, 2.0, 5.0])
    rows = np.zeros((len(ts), 8, 8))
    for a, t in enumerate(ts):
        pol = opt.Polynomial([1, 1, 1, 1, 1, 1, 1, 1])
        for j in range(8):
            v = np.array(pol.pol_coeffs_at_t(float(t)))
            rows[a, j] = np.pad(v, (8 - len(v), 0), "constant")
            pol = pol.derivative()
            assert np.array_equal(rows[a, j], mo.time_power_row(t, j)), (t, j)
    np.savez(os.path.join(GOLD, "time_power_rows.npz"), t=ts, rows=rows)

    # ---- a2-a4: solves --------------------------------------------------------
    cases = {}

    def add_case(name, wp, t):
        coef, durs = ref_solve(opt, ct, wp, t)
        ocoef, odurs = mo.solve_waypoints(wp, t)
        assert np.array_equal(coef, ocoef), name
        assert np.array_equal(durs, odurs), name
        cases[name + "__wp"] = wp
        cases[name + "__t"] = t
        cases[name + "__coef"] = coef
        cases[name + "__dur"] = durs

    def walk(m, K, step=0.3):
        start = rng.uniform([-2.2, 2.8, 0.5, -1.0][:K], [2.2, 5.0, 2.5, 1.0][:K])
        inc = rng.normal(0.0, step, size=(m, K))
        if K == 4:
            inc[:, 3] = rng.normal(0.0, 0.1, size=m)
        inc[0] = 0
        return start + np.cumsum(inc, axis=0)

    for n in (1, 2, 3, 5, 10, 20):
        for rep in range(3):
            T = rng.uniform(0.5, 2.0, n)
            add_case("rand_n%d_r%d" % (n, rep), walk(n + 1, 4), np.concatenate([[0.0], np.cumsum(T)]))
    # shipped-example shape: 50 poses, uniform step 10/50
    add_case("uniform_n49", walk(50, 4, 0.05), mo.uniform_times(50))
    # ill-conditioned duration mixes (config 3 style)
    for rep in range(4):
        n = 20
        T = np.clip(rng.uniform(0.5, 2.0, n) * np.exp(0.25 * (2 * rep + 1) * rng.normal(size=n)), 0.05, 5.0)
        add_case("stress_n20_r%d" % rep, walk(n + 1, 3), np.concatenate([[0.0], np.cumsum(T)]))
    T = np.where(rng.uniform(size=20) < 0.25, 0.02, rng.uniform(0.05, 5.0, 20))
    add_case("stress_tail_n20", walk(21, 3), np.concatenate([[0.0], np.cumsum(T)]))
    # quirk (i): first stamp not zero
    T = rng.uniform(0.5, 2.0, 6)
    add_case("t0_nonzero_n6", walk(7, 3), 0.4 + np.concatenate([[0.0], np.cumsum(T)]))
    # the reference's own __main__ fixture (calculatingTrajectories.py:240-259)
    td = np.asarray(ct.test_data, dtype=np.float64)
    add_case("reference_test_data", td, np.arange(len(td)) * float(ct.timestep))
    np.savez(os.path.join(GOLD, "solve_cases.npz"), **cases)

    # ---- a5/a6: evaluation ----------------------------------------------------
    ev = {}
    wp = walk(8, 4)
    T = rng.uniform(0.5, 2.0, 7)
    tt = np.concatenate([[0.0], np.cumsum(T)])
    pts = [opt.Point_time(opt.Waypoint(*[float(v) for v in wp[i]]), t=float(tt[i])) for i in range(8)]
    _, pcs = opt.calculate_trajectory4D(pts)
    coef, durs = ref_solve(opt, ct, wp, tt)
    total = float(sum(durs))
    knots = [0.0]
    for d in durs:
        knots.append(knots[-1] + d)
    sample_t = np.concatenate([np.linspace(0, total, 41), np.asarray(knots),
                               np.nextafter(np.asarray(knots[1:]), 0), [total + 0.3, total + 2.0]])
    vals = np.zeros((len(sample_t), 4))
    for s, t in enumerate(sample_t):
        for k in range(4):
            vals[s, k] = float(np.asarray(pcs[k].eval(float(t))).reshape(()))
            assert vals[s, k] == mo.piecewise_eval(coef[:, k, :], durs, t), (s, k)
    # derivative levels through Polynomial.derivative()
    dvals = np.zeros((3, len(sample_t), 4))
    for level in range(1, 4):
        for s, t in enumerate(sample_t):
            i, local = mo.piece_lookup([float(d) for d in durs], t)
            for k in range(4):
                pol = pcs[k].pols[i]
                for _ in range(level):
                    pol = pol.derivative()
                dvals[level - 1, s, k] = float(np.asarray(pol.eval(local)).reshape(()))
                assert dvals[level - 1, s, k] == mo.piecewise_eval(coef[:, k, :], durs, t, level)
    ev.update(coef=coef, dur=durs, t=sample_t, values=vals, deriv_values=dvals)
    np.savez(os.path.join(GOLD, "piecewise_eval.npz"), **ev)

    # ---- Trajectory.loadcsv / eval + Polynomial4D.eval on shipped CSVs ----------
    tr_out = {}
    for label, path in (("traj", os.path.join(REF, "src", "traj.csv")),
                        ("pol1", os.path.join(REF, "resources", "trajectories", "Pol_matrix_1.csv"))):
        tr = opt.Trajectory()
        tr.loadcsv(path)
        raw = np.loadtxt(path, delimiter=",", skiprows=1 if label == "traj" else 0, usecols=range(33))
        tsamp = np.concatenate([np.arange(0, tr.duration, 0.1), [tr.duration]])
        pos = np.zeros((len(tsamp), 3)); vel = np.zeros((len(tsamp), 3)); acc = np.zeros((len(tsamp), 3))
        om = np.zeros((len(tsamp), 3)); yaw = np.zeros(len(tsamp))
        used = raw if label == "traj" else raw[1:]        # quirk (iii): first row skipped
        udur = used[:, 0]
        for s, t in enumerate(tsamp):
            o = tr.eval(float(t))
            pos[s], vel[s], acc[s], om[s], yaw[s] = o.pos, o.vel, o.acc, o.omega, o.yaw
            i, local = mo.trajectory_lookup(udur, t)
            f = mo.flat_output(used[i, 1:].reshape(4, 8), local)
            assert np.array_equal(f["pos"], o.pos) and np.array_equal(f["omega"], o.omega), (label, s)
            assert np.array_equal(f["vel"], o.vel) and np.array_equal(f["acc"], o.acc) and f["yaw"] == o.yaw
        tr_out.update({label + "__file_rows": raw, label + "__n_pieces": np.array(tr.n_pieces()),
                       label + "__duration": np.array(tr.duration), label + "__t": tsamp,
                       label + "__pos": pos, label + "__vel": vel, label + "__acc": acc,
                       label + "__omega": om, label + "__yaw": yaw})
    np.savez(os.path.join(GOLD, "trajectory_eval.npz"), **tr_out)

    # ---- shipped polynomial matrices (golden vectors of path_to_pol) -------------
    mats = {}
    for path in sorted(glob.glob(os.path.join(REF, "resources", "trajectories", "*.csv"))):
        name = os.path.splitext(os.path.basename(path))[0]
        mats[name] = np.loadtxt(path, delimiter=",").astype(np.float32)
        with open(path) as fh:
            mats[name + "__first_line"] = np.array(fh.readline().rstrip("\n"))
    np.savez(os.path.join(GOLD, "shipped_pol_matrices.npz"), **mats)

    # ---- meshes (data the product needs for the named obstacle configs) ----------
    from oracle import collision_oracle as co
    meshes = {}
    for path in sorted(glob.glob(os.path.join(REF, "resources", "stl", "*.stl"))):
        name = os.path.splitext(os.path.basename(path))[0]
        meshes[name] = co.read_stl_triangles(path)
    np.savez(os.path.join(DATA, "stl_meshes.npz"), **meshes)
    print("golden fixtures written to", GOLD, "and", DATA)


if __name__ == "__main__":
    main()
