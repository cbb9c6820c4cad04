"""The drop-in mirror of the reference's env class (marl_llm_b200/assembly_env.py + compat/gym) replayed against the golden
trajectories recorded from the real AssemblySwarmEnv: same seeds -> same reset() (NumPy-global RNG order), same 5-tuples."""
import argparse
import os
import sys

import numpy as np
import pytest

from tests.helpers import GOLDEN_CASES, REPO, load_golden, load_shapes, replay_golden

pytestmark = pytest.mark.gpu


def results_blob():
    sh = load_shapes()
    return {"l_cell": [float(v) for v in sh["l_cell"]],
            "grid_coords": [np.ascontiguousarray(g.T) for g in sh["grid_origin"]],
            "binary_image": [np.zeros((1, 1)) for _ in sh["l_cell"]],
            "shape_bound_points": [np.zeros(4) for _ in sh["l_cell"]]}


def make_args(n_a, **kw):
    d = dict(n_a=n_a, is_boundary=True, is_con_self_state=True, is_feature_norm=False, dynamics_mode="Cartesian",
             render_traj=False, traj_len=15, agent_strategy="input", training_method="llm_rl", is_collected=False,
             results_file=results_blob(), video=False)
    d.update(kw)
    return argparse.Namespace(**d)


@pytest.fixture()
def gym():
    sys.path.insert(0, os.path.join(REPO, "marl_llm_b200", "compat"))
    for m in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
        del sys.modules[m]
    import gym as g
    assert "swarm_b200" in g.__version__
    yield g
    sys.path.remove(os.path.join(REPO, "marl_llm_b200", "compat"))
    for m in [k for k in sys.modules if k == "gym" or k.startswith("gym.")]:
        del sys.modules[m]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_dropin_env_replays_reference_goldens(gym, case):
    g = load_golden(case)
    n_a = int(g["n_a"])
    base = gym.make("AssemblySwarm-v0").unwrapped                     # train_assembly.py:49
    env = gym.wrappers.AssemblySwarmWrapper(base, make_args(n_a, is_boundary=not bool(g.get("is_periodic", 0))))   # train_assembly.py:50
    assert env.num_agents == n_a and env.agent_types == ["agent"]
    assert env.observation_space.shape == (192, n_a) and env.action_space.shape == (2, n_a)
    assert env.r_avoid == float(g["r_avoid"])

    def snap(obs, rew=None, prior=None):
        e = env.env
        return dict(p=e.p, dp=e.dp, obs=obs, reward=rew, a_prior=prior, nbr=e.neighbor_index, in_flags=e.in_flags,
                    sensed=e.sensed_index, occupied=e.occupied_index)

    def reset_fn(g):
        np.random.seed(int(g["seed"]))
        obs = env.reset()
        assert obs.shape == (192, n_a) and obs.dtype == np.float64
        assert np.array_equal(env.env.grid_center, g["grid_center"]) and env.env.l_cell == float(g["l_cell"])
        assert np.array_equal(env.p, g["p0"]) and np.array_equal(env.dp, g["dp0"])
        return snap(obs)

    def step_fn(a):
        obs, rew, done, info, prior = env.step(a)
        assert rew.shape == (1, n_a) and done.shape == (1, n_a) and done.dtype == bool and not done.any()
        assert info.shape == (3, 1) and prior.shape == (2, n_a)
        return snap(obs, rew, prior)

    replay_golden(g, reset_fn, step_fn)
    env.close()


def test_eval_script_style_pokes_and_metrics(gym):
    """eval_assembly.py:34-57 swaps the target shape through env.env.* between steps and reads the wrapper metrics."""
    from oracle import oracle as orc
    sh = load_shapes()
    n_a = 30
    env = gym.wrappers.AssemblySwarmWrapper(gym.make("AssemblySwarm-v0").unwrapped, make_args(n_a))
    np.random.seed(3)
    env.reset()
    rng = np.random.RandomState(0)
    for t in range(12):
        if t == 5:
            k = 6
            env.env.l_cell = float(sh["l_cell"][k])
            env.env.grid_center_origin = sh["grid_origin"][k]
            env.env.n_g = sh["grid_origin"][k].shape[1]
            env.env.grid_center = sh["grid_origin"][k].copy() + np.zeros((2, 1))
        p0, dp0, nbr0 = env.p.copy(), env.dp.copy(), env.env.neighbor_index.copy()
        a = rng.uniform(-1, 1, (2, n_a)).astype(np.float32)
        obs, rew, done, info, prior = env.step(a)
        grid = env.env.grid_center
        P = orc.make_params(n_a, grid.shape[1], float(env.env.l_cell), env.r_avoid)
        ob = orc.OracleBatch([P])
        ob.p[0], ob.dp[0] = p0, dp0
        ob.set_grid(0, grid)
        ob.neighbor_index[0] = nbr0
        ob.step(a[None])
        assert np.array_equal(obs, ob.obs[0]) and np.array_equal(prior, ob.a_prior[0]) and np.array_equal(rew, ob.reward[0])
        assert np.array_equal(env.p, ob.p[0])
    assert 0.0 <= env.coverage_rate() <= 1.0
    assert np.isfinite(env.distribution_uniformity()) and np.isfinite(env.voronoi_based_uniformity())
    newp = env.p + 0.01
    env.env.p = newp                                                 # eval copies / overwrites env.p
    assert np.array_equal(env.p, newp)
    env.close()


def test_vectorised_variant_has_leading_batch_axis(gym):
    env = gym.wrappers.AssemblySwarmWrapper(gym.AssemblySwarmEnv(num_envs=4), make_args(30))
    np.random.seed(1)
    obs = env.reset()
    assert obs.shape == (4, 192, 30)
    out = env.step(np.zeros((4, 2, 30), np.float32))
    assert out[0].shape == (4, 192, 30) and out[1].shape == (4, 1, 30) and out[4].shape == (4, 2, 30)
    env.close()


def test_wrapper_metrics_match_the_reference_formulas(gym):
    """coverage_rate / distribution_uniformity / voronoi_based_uniformity (assembly_wrapper.py:48-129) computed on the device
    against a NumPy restatement of the reference loops."""
    n_a = 30
    env = gym.wrappers.AssemblySwarmWrapper(gym.make("AssemblySwarm-v0").unwrapped, make_args(n_a))
    np.random.seed(11)
    env.reset()
    rng = np.random.RandomState(1)
    from tests.helpers import goal_seeking_action
    for t in range(60):
        env.step(goal_seeking_action(env.env.obs, env.dp, rng))
        if t % 20 != 19:
            continue
        p, g, r = env.p, env.env.grid_center, env.r_avoid
        occupied = sum(1 for c in range(g.shape[1]) if (np.linalg.norm(p - g[:, [c]], axis=0) < r / 2).any())   # WRAP:59-66
        assert env.coverage_rate() == occupied / g.shape[1]
        mins = []
        for i in range(n_a):                                                                                   # WRAP:87-95
            d = np.linalg.norm(p - p[:, [i]], axis=0)
            mins.append(np.min(d[d != 0]))
        ref2 = (np.var(mins) - np.min(mins)) / (np.max(mins) - np.min(mins))
        assert abs(env.distribution_uniformity() - ref2) <= 1e-12 * max(1.0, abs(ref2))
        cnt = np.zeros(n_a)
        for c in range(g.shape[1]):                                                                            # WRAP:114-121
            cnt[np.argmin(np.linalg.norm(p - g[:, [c]], axis=0))] += 1
        ref3 = (np.var(cnt) - np.min(cnt)) / (np.max(cnt) - np.min(cnt))
        assert abs(env.voronoi_based_uniformity() - ref3) <= 1e-12 * max(1.0, abs(ref3))
    assert occupied > 0
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n_a,E", [("llm", 30, 48), ("rule", 30, 48), ("rule", 9, 16), ("llm", 64, 8)])
def test_strategy_actions_match_the_oracle_restatement(kind, n_a, E):
    """swarm_strategy_actions ('rule' assembly.py:530-601, 'llm' assembly.py:524-529/876-941) against the C restatement
    that tests/test_oracle_vs_reference.py pins to the NumPy original at 1e-12: bit-identical for 'llm'; 'rule' may
    differ in the last bits of the cosine weights (1e-13).  States come from a goal-seeking rollout (agents inside the
    shapes, occupancy filter live) and from a small-r_avoid batch that exercises the 80-cell half-even subsample."""
    import torch
    from marl_llm_b200.batched import BatchedAssemblySim
    from oracle import oracle as orc
    from tests.helpers import goal_seeking_action, load_shapes, reset_like_reference
    shapes = load_shapes()
    ngm = int(shapes["n_g"].max())
    for r_avoid, d_sen in ((orc.r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"]), 0.4), (0.05, 0.6)):
        rng = np.random.RandomState(17 + n_a)
        sim = BatchedAssemblySim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True, d_sen=d_sen)
        params, grids, P, DP = [], [], [], []
        for e in range(E):
            k, grid, p, dp = reset_like_reference(rng, n_a, shapes)
            params.append(orc.make_params(n_a, grid.shape[1], float(shapes["l_cell"][k]), r_avoid, d_sen=d_sen))
            grids.append(grid); P.append(p); DP.append(dp)
        ob = orc.OracleBatch(params, nthreads=8)
        for e in range(E):
            ob.set_grid(e, grids[e])
        ob.p[:], ob.dp[:] = np.stack(P), np.stack(DP)
        blocks, n_g = sim.pack_grids(grids, ngm)
        sim.set_grid(blocks, n_g, [q.l_cell for q in params])
        sim.set_state(ob.p, ob.dp)
        sim.observe(); ob.observe()
        worst, used_sub = 0.0, 0
        for t in range(40):
            want = orc.strategy_actions(ob, kind)
            got = sim.strategy_actions(kind).cpu().numpy()
            if kind == "llm":
                assert np.array_equal(got, want), t
            else:
                worst = max(worst, float(np.max(np.abs(got - want))))
            used_sub += int(((ob.sensed_index >= 0).sum(2) == 80).sum())
            a = goal_seeking_action(ob.obs, ob.dp, rng)
            sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        assert worst < 1e-13, worst
        assert ob.in_flags.sum() > 0
        if d_sen == 0.6:
            assert used_sub > 0


@pytest.mark.gpu
def test_dropin_env_runs_its_own_rule_strategy_like_collect_expert_data():
    """collect_expert_data.py:110-113 sets agent_strategy='rule', is_collected=True: step() ignores the caller's action,
    computes the controller on the device and returns it as the 5th output (assembly.py:663-664); the swarm assembles."""
    from marl_llm_b200.assembly_env import AssemblySwarmEnv
    args = make_args(30, agent_strategy="rule", is_collected=True)
    env = AssemblySwarmEnv()
    env.__reinit__(args)
    np.random.seed(3)
    env.reset()
    inside = []
    for t in range(250):
        obs, rew, done, info, u = env.step(np.zeros((2, 30)))
        assert u.shape == (2, 30) and np.all(np.abs(u) <= 1)
        inside.append(float(np.mean(np.all(obs[28:30] == 0, axis=0))))       # target offset is zero iff in the shape (CPP:136-137)
    # the reference with the same seed: in-shape fraction 0.45 over the first 20 steps, 1.0 over the last 20
    assert np.mean(inside[:20]) < 0.7 and np.mean(inside[-20:]) > 0.95


@pytest.mark.gpu
def test_device_metrics_match_values_recorded_from_the_real_wrapper():
    """(f2) swarm_metrics / k_metrics against the numbers the UNMODIFIED AssemblySwarmWrapper produced (assembly_wrapper.py:48-129;
    tests/golden/metrics.npz, recorded by tests/golden/make_metric_goldens.py): coverage exact, the two variance ratios to
    1e-12 (NumPy's np.var sums pairwise, the kernel sequentially)."""
    import torch
    from marl_llm_b200.batched import BatchedAssemblySim
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "metrics.npz"))
    for k in range(int(z["n_cases"])):
        grid, r, P, want = z[f"c{k}_grid"], float(z[f"c{k}_r_avoid"]), z[f"c{k}_p"], z[f"c{k}_metrics"]
        K, _, n_a = P.shape
        sim = BatchedAssemblySim(K, n_a, grid.shape[1], r, out_dtype=torch.float64)
        blocks, n_g = sim.pack_grids([grid] * K, grid.shape[1])
        sim.set_grid(blocks, n_g, [float(z[f"c{k}_l_cell"])] * K)
        sim.set_state(P, np.zeros_like(P))
        got = sim.metrics().cpu().numpy()
        assert np.array_equal(got[:, 0], want[:, 0]), k
        assert np.allclose(got[:, 1:], want[:, 1:], rtol=1e-12, atol=0), (k, np.abs(got[:, 1:] - want[:, 1:]).max())
        assert want[:, 0].max() > 0.3


@pytest.mark.gpu
def test_installing_a_larger_shape_between_steps_keeps_the_last_observation(gym):
    """eval_assembly.py:34-57 swaps the target shape mid-episode with no reset.  When the new shape has more cells than the
    handle was sized for, the drop-in rebuilds the handle; the next step must still use the LAST observation's neighbour
    list for its prior (assembly.py:613-624) and the new grid for everything else — compared with the oracle."""
    from oracle import oracle as orc
    sh = load_shapes()
    n_a = 30
    blob = results_blob()
    order = np.argsort(sh["n_g"])
    small = [int(k) for k in order[:3]]                       # the env only knows the three smallest shapes ...
    blob = {key: [val[k] for k in small] for key, val in blob.items()}
    env = gym.wrappers.AssemblySwarmWrapper(gym.make("AssemblySwarm-v0").unwrapped, make_args(n_a, results_file=blob))
    np.random.seed(4)
    env.reset()
    cap0 = env.env._n_g_cap
    rng = np.random.RandomState(1)
    for t in range(10):
        if t == 4:                                            # ... and the caller installs the largest one
            k = int(order[-1])
            assert sh["n_g"][k] > cap0
            env.env.l_cell = float(sh["l_cell"][k])
            env.env.grid_center_origin = sh["grid_origin"][k]
            env.env.n_g = sh["grid_origin"][k].shape[1]
            env.env.grid_center = sh["grid_origin"][k].copy() + np.array([[0.3], [-0.2]])
        p0, dp0, nbr0 = env.p.copy(), env.dp.copy(), env.env.neighbor_index.copy()
        a = rng.uniform(-1, 1, (2, n_a)).astype(np.float32)
        obs, rew, done, info, prior = env.step(a)
        grid = env.env.grid_center
        P = orc.make_params(n_a, grid.shape[1], float(env.env.l_cell), env.r_avoid)
        ob = orc.OracleBatch([P])
        ob.p[0], ob.dp[0] = p0, dp0
        ob.set_grid(0, grid)
        ob.neighbor_index[0] = nbr0
        ob.step(a[None])
        assert np.array_equal(obs, ob.obs[0]) and np.array_equal(prior, ob.a_prior[0]) and np.array_equal(rew, ob.reward[0]), t
        assert np.array_equal(env.p, ob.p[0]) and np.array_equal(env.env.neighbor_index, ob.neighbor_index[0])
    assert env.env._n_g_cap > cap0
    env.render()                                              # a no-op on the drop-in (train_assembly.py:93-94, eval_assembly.py:147)
    env.close()
