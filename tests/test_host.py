"""Host-side mirror of the reference surface + the C-ABI library contract (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_of_the_header():
    from mpc_motion_planning_b200 import _lib, build

    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    declared = set(re.findall(r"\b(mpcb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mpcb_version() == 200


def test_cfg_struct_layout_matches_the_header_field_order():
    from mpc_motion_planning_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "mpcb200.h")).read()
    body = hdr[hdr.index("typedef struct mpcb_cfg {"): hdr.index("} mpcb_cfg;")]
    names = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0].strip()
        m = re.match(r"(int32_t|double)\s+(.*);", line)
        if m:
            names += [re.sub(r"\[.*\]", "", n).strip() for n in m.group(2).split(",")]
    assert names == [f[0] for f in _lib.MpcbCfg._fields_]


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mpc_motion_planning_b200 import _lib
    from mpc_motion_planning_b200.solver import BatchSolver

    with pytest.raises(_lib.MpcbError, match="no CUDA device"):
        BatchSolver("kin_cbf_pre")
    from mpc_motion_planning_b200.pipeline import PipelinedSolver

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PipelinedSolver(2, "kin_cbf_pre")
    lib = _lib.load()
    assert lib.mpcb_wait(None) < 0 and lib.mpcb_submit_batch_host(None, 1, *([None] * 10)) < 0   # argument errors, not crashes


def test_create_rejects_bad_configurations():
    from mpc_motion_planning_b200 import _lib
    from mpc_motion_planning_b200.helpers import PACKAGE_PARAMS, load_config
    from mpc_motion_planning_b200.problem import make_cfg

    lib = _lib.load()
    cfgd = load_config(PACKAGE_PARAMS)
    h = C.c_void_p()
    for mut in (dict(N=1), dict(N=1000), dict(M=9), dict(tol=0.0), dict(mu_init=-1.0)):
        c = make_cfg("kin_cbf_pre", cfgd)
        for k, v in mut.items():
            setattr(c, k, v)
        assert lib.mpcb_create(C.byref(c), C.byref(h)) == -1, mut
    c = make_cfg("kin_cbf_pre", cfgd)
    c.u_lo[0] = 1.0  # lower bound above upper bound
    assert lib.mpcb_create(C.byref(c), C.byref(h)) == -1
    assert lib.mpcb_create(None, C.byref(h)) == -1
    assert b"unsupported" in lib.mpcb_strerror(-1)
    sz = C.c_size_t(7)
    assert lib.mpcb_workspace_bytes(C.byref(make_cfg("kin_cbf_pre", cfgd)), 1000, C.byref(sz)) == 0


def test_problem_constants_match_the_oracle_restatement():
    """Two independent transcriptions of the reference's hard-coded numbers must agree."""
    from mpc_motion_planning_b200.helpers import PACKAGE_PARAMS, load_config
    from mpc_motion_planning_b200.problem import horizon_steps, make_cfg
    from oracle import c_oracle

    cfgd = load_config(PACKAGE_PARAMS)
    assert horizon_steps(cfgd) == 50
    assert cfgd["mpc_params"]["is_variable_time"] == "Flase"  # sic: parsed as a string, never == True
    for kind in ("kin_nocbf", "kin_cbf", "kin_cbf_pre", "dyn"):
        a = make_cfg(kind, cfgd)
        b = c_oracle.make_cfg(kind)
        for name, _ in a._fields_:
            if not hasattr(b, name):
                continue  # product-only fields (obs_input, reserved)
            va, vb = getattr(a, name), getattr(b, name)
            if hasattr(va, "__len__"):
                assert list(va) == list(vb), (kind, name)
            else:
                assert va == vb, (kind, name)


def test_initialize_constraints_lengths_and_values(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # no YAML in the CWD: falls back to the packaged copy
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn, MPC_CBF_optimize_kin, MPC_CBF_optimize_kin_pre, MPC_optimize_kin
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction

    m = MPC_CBF_optimize_kin.MPC_optimize()
    assert (m.N_p, m.T_S, m.num_states, m.num_controls) == (50, 0.1, 4, 2)
    assert len(m.t_vector) == 51
    lbg, ubg, lbx, ubx = m.initialize_constraints(np.array([[50, 3.5, 0, 8, 4.8, 1.8]]))
    assert (len(lbg), len(ubg), len(lbx), len(ubx)) == (303, 303, 304, 304)
    assert lbx[:2] == [m.df_min, m.ax_min] and ubx[100:104] == [np.inf, 5, np.inf, 40.0]
    assert lbg[204] == pytest.approx(-0.008726646259971648) and ubg[-1] == np.inf and lbg[-1] == 0.0
    mp = MPC_CBF_optimize_kin_pre.MPC_optimize()
    tr = obs_prediction([np.array([[50, 3.5, 0, 10, 4.8, 1.8]]), np.array([[70, 0, 0, 5, 4.8, 1.8]])], 0.1, 50)
    lbg, ubg, _, _ = mp.initialize_constraints(tr)
    assert len(lbg) == 204 + 49 + 100
    md = MPC_CBF_optimize_dyn.MPC_optimize()
    lbg, ubg, lbx, ubx = md.initialize_constraints()
    assert (len(lbg), len(lbx)) == (455, 406)
    # as shipped (default): the rate pair sits at 12,13 (on the x/y defect of stage 1); aligned: at the rate rows 18,19
    assert lbg[12:14] == [pytest.approx(-0.008726646259971648), pytest.approx(-0.3)] and lbg[18:20] == [0.0, 0.0] and lbg[-1] == 1
    md.dyn_bounds = "aligned"
    lbg, ubg, lbx, ubx = md.initialize_constraints()
    assert lbg[18:20] == [pytest.approx(-0.008726646259971648), pytest.approx(-0.3)] and lbg[12:14] == [0.0, 0.0] and lbg[-1] == 1
    mk = MPC_optimize_kin.MPC_optimize()
    lbg, ubg, lbx, ubx = mk.initialize_constraints()
    assert lbg == 0.0 and ubg == 0.0 and len(lbx) == 304


def test_model_function_matches_the_oracle_models():
    from mpc_motion_planning_b200 import MPC_CBF_optimize_dyn, MPC_CBF_optimize_kin
    from oracle.nlp import DynModel, KinModel, Params

    rng = np.random.default_rng(5)
    mk, md = MPC_CBF_optimize_kin.MPC_optimize(), MPC_CBF_optimize_dyn.MPC_optimize()
    for _ in range(5):
        x4 = np.array([rng.uniform(0, 100), rng.uniform(-1, 5), rng.uniform(-0.3, 0.3), rng.uniform(5, 30)])
        u = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-3, 3)])
        assert np.allclose(mk.f(x4.reshape(-1, 1), u).full().ravel(), KinModel(Params()).f(x4, u), rtol=1e-14)
        x6 = np.concatenate([x4, [rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)]])
        assert np.allclose(md.f(x6, u).full().ravel(), DynModel(Params()).f(x6, u), rtol=1e-12)


def test_obs_prediction_surface():
    from mpc_motion_planning_b200.Obs_prediction import obs_prediction, obs_prediction_batch

    obs_list = [np.array([[50, 3.5, 0, 15, 4.8, 1.8]]), np.array([[60, 10, np.pi / 4, 10, 3.6, 1.5]]),
                np.array([[70, -5, -np.pi / 6, 12, 5.0, 2.0]])]  # PKG/test.py:7-14
    tr = obs_prediction(obs_list, 0.1, 50)
    assert len(tr) == 3 and all(t.shape == (51, 6) for t in tr)
    for o, t in zip(obs_list, tr):
        x, y, th, v, l, w = o[0]
        for k in range(51):  # the reference's step-by-step recursion, bit for bit
            assert t[k, 0] == x and t[k, 1] == y and t[k, 4] == l and t[k, 5] == w
            x, y = x + v * np.cos(th) * 0.1, y + v * np.sin(th) * 0.1
    b = obs_prediction_batch(np.stack([o[0] for o in obs_list]), 0.1, 50)
    assert all(np.array_equal(b[i], tr[i]) for i in range(3))
    assert obs_prediction([], 0.1, 50) == []


def test_ref_path_generator_surface():
    from mpc_motion_planning_b200.RefPathGenerator import RefPathGenerator

    r = RefPathGenerator()
    x0 = np.array([0, 3, 0, 15.0]).reshape(-1, 1)
    xs = np.array([400, 3.5, 0, 30.0]).reshape(-1, 1)
    g = r.define_ref_path(x0, xs, 0.1)
    assert g.shape == (401, 4) and np.all(g[:, 1] == 3.5) and g[0, 0] == 0 and g[-1, 0] == 400
    traj, idx = r.find_ref_traj(x0, xs, 5, 0.1, 0)
    assert traj.shape == (51, 4) and idx == 0
    assert traj[-1, 0] == 112  # preview 0.5*(15+30)*5 = 112.5 m -> index 112
    back = RefPathGenerator().define_ref_path(xs, x0, 0.1)
    assert back[0, 0] == 400 and back[-1, 0] == 0


def test_scenario_generators_are_seeded_and_shaped():
    from mpc_motion_planning_b200 import scenarios

    a = scenarios.kin_cbf_moving(32)
    b = scenarios.kin_cbf_moving(32)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))
    x0, xs, obs = a
    assert x0.shape == (32, 4) and obs.shape == (32, 1, 51, 6)
    assert np.all(obs[:, 0, :, 4] == 4.8) and np.all((x0[:, 3] >= 10) & (x0[:, 3] <= 25))
    x0, xs, obs = scenarios.kin_cbf_static(8, N=20)
    assert obs.shape == (8, 1, 21, 6) and np.all(obs[:, 0, 0] == obs[:, 0, -1])
    x0, xs, obs = scenarios.dyn_static(8)
    assert x0.shape == (8, 6) and np.all(xs[:, 3] == 15)


def test_shard_ranges_cover_the_batch():
    from mpc_motion_planning_b200.sharding import balanced_permutation, shard_range

    for B in (0, 1, 7, 8, 1000, 1001):
        for G in (1, 2, 4, 8):
            spans = [shard_range(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
    p = balanced_permutation(100, 0)
    assert sorted(p) == list(range(100)) and np.array_equal(p, balanced_permutation(100, 0))


def test_install_reference_names_registers_flat_modules():
    import sys

    import mpc_motion_planning_b200 as pkg

    pkg.install_reference_names()
    import MPC_CBF_optimize_kin_pre  # noqa: F401  (the reference's flat module names)
    from Obs_prediction import obs_prediction  # noqa: F401

    for name in pkg.REFERENCE_MODULES:
        assert sys.modules[name].__name__ == f"mpc_motion_planning_b200.{name}"
    assert hasattr(sys.modules["MPC_optimize_kin"], "MPC_optimize")
    for name in pkg.REFERENCE_MODULES:
        sys.modules.pop(name, None)


def test_header_and_example_compile_as_plain_c(tmp_path):
    """include/mpcb200.h is a C header (no C++ in the boundary): the plain-C example compiles with -std=c11
    -pedantic and links against the library without a GPU."""
    import subprocess

    from mpc_motion_planning_b200 import _lib

    _lib.load()
    exe = str(tmp_path / "batch_solve")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "batch_solve.c"), "-L", os.path.dirname(_lib.SO_PATH), "-lmpcb200", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # without a device the program must fail loudly at mpcb_create (no CPU fallback)
    import torch

    if not torch.cuda.is_available():
        env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(_lib.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        run = subprocess.run([exe], capture_output=True, text=True, env=env)
        assert run.returncode == 2 and "mpcb_create" in run.stderr


def test_only_the_checkers_touch_the_oracle():
    """oracle/ is test infrastructure: the package, the mains, the examples and the scripts never import it
    (bench.py's CPU legs and __graft_entry__'s build/smoke are the two allowed callers outside tests/)."""
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    offenders = []
    for top in ("mpc_motion_planning_b200", "mains", "examples", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py") and pat.search(open(os.path.join(dirpath, f)).read()):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    assert "liboracle" not in open(os.path.join(ROOT, "mpc_motion_planning_b200", "_lib.py")).read()


def test_shard_permutation_is_cached_per_batch_size_and_seed():
    """solve_sharded keeps its shuffle on the device, one per (batch size, seed): two batch sizes in one process must not
    see each other's permutation, and the result comes back in input order."""
    import torch

    from mpc_motion_planning_b200 import sharding

    def solve_local(a, b, c, z):
        n = a.shape[0]
        return {"u0": torch.stack([a[:, 0], a[:, 0] * 2], 1), "cost": a[:, 0] + 0.5, "status": torch.zeros(n, dtype=torch.int32),
                "iters": a[:, 0].to(torch.int32)}

    for B in (37, 64, 37):
        x0 = torch.arange(B, dtype=torch.float64).reshape(B, 1)
        for seed in (0, 3, None):
            out = sharding.solve_sharded(solve_local, x0, x0, None, shuffle_seed=seed)
            assert torch.equal(out["u0"][:, 0], x0[:, 0]) and torch.equal(out["cost"], x0[:, 0] + 0.5)
            assert torch.equal(out["iters"], torch.arange(B, dtype=torch.int32))
            perm = sharding._device_permutation(B, seed, x0.device)
            assert perm.numel() == B and torch.equal(torch.sort(perm).values, torch.arange(B))
