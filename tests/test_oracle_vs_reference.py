"""Pins oracle/assembly_oracle.c against the UNMODIFIED reference (real AssemblySwarmEnv driving its
own AssemblyEnv.cpp, compiled by oracle/Makefile into oracle/_ref/).  Bit-exact on every output of
every step.  Build-container only: skipped where /root/reference is not mounted (GPU box)."""
import numpy as np
import pytest

from oracle import live_reference as lr
from oracle import oracle as orc
from tests.helpers import goal_seeking_action

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not lr.available(), reason="reference checkout / oracle/_ref not available")]

FIELDS = ("p", "dp", "obs", "rew", "prior", "nbr", "inf", "sen", "occ")


def rollout_pair(n_a, seed, mode, steps, self_state=True, training_method="llm_rl"):
    env = lr.make_env(n_a, is_con_self_state=self_state, training_method=training_method)
    np.random.seed(seed)
    env.reset()
    e = env.env
    want_prior = training_method == "llm_rl"
    P = orc.make_params(n_a, e.n_g, float(e.l_cell), float(e.r_avoid), is_con_self_state=self_state, want_prior=want_prior)
    assert e.obs.shape == (P.obs_dim, n_a) and P.obs_dim == (192 if self_state else 188)
    ob = orc.OracleBatch([P])
    ob.p[0], ob.dp[0] = e.p, e.dp
    ob.set_grid(0, e.grid_center)
    ob.observe()
    assert np.array_equal(ob.obs[0], e.obs)
    assert np.array_equal(ob.neighbor_index[0], e.neighbor_index)
    assert np.array_equal(ob.sensed_index[0], e.sensed_index)
    rng = np.random.RandomState(seed + 1)
    stats = dict(in_shape=0, subsampled=0, occupied=0, reward=0.0)
    for t in range(steps):
        if mode == "random":
            a = rng.uniform(-1, 1, (2, n_a)).astype(np.float32)
        else:
            a = goal_seeking_action(e.obs, e.dp, rng, target_row=28 if self_state else 24)
        obs, rew, done, info, prior = env.step(a)
        ob.step(a[None])
        if not want_prior:                                   # assembly.py:666: the 5th output is None unless llm_rl
            assert prior is None and not ob.a_prior.any()
            prior = ob.a_prior[0]
        got = dict(p=ob.p[0], dp=ob.dp[0], obs=ob.obs[0], rew=ob.reward[0], prior=ob.a_prior[0],
                   nbr=ob.neighbor_index[0], inf=ob.in_flags[0], sen=ob.sensed_index[0], occ=ob.occupied_index[0])
        ref = dict(p=e.p, dp=e.dp, obs=obs, rew=rew, prior=prior, nbr=e.neighbor_index, inf=e.in_flags,
                   sen=e.sensed_index, occ=e.occupied_index)
        for k in FIELDS:
            assert np.array_equal(got[k], ref[k], equal_nan=True), f"{k} differs at step {t} (n_a={n_a}, seed={seed}, {mode})"
        assert not done.any()
        stats["in_shape"] += int(e.in_flags.sum())
        stats["occupied"] += int((e.occupied_index >= 0).sum())
        stats["subsampled"] += int(((e.sensed_index >= 0).sum(1) == 80).sum())
        stats["reward"] += float(rew.sum())
    return stats


@pytest.mark.parametrize("seed", [226, 1, 2])
def test_random_actions_30_agents_200_steps(seed):
    rollout_pair(30, seed, "random", 200)


@pytest.mark.parametrize("seed", [226, 3])
def test_goal_seeking_exercises_in_shape_branches(seed):
    st = rollout_pair(30, seed, "goal", 200)
    assert st["in_shape"] > 500 and st["subsampled"] >= 1 and st["occupied"] > 1000 and st["reward"] > 0, st


@pytest.mark.parametrize("n_a,steps", [(1, 20), (2, 50), (10, 100), (64, 60), (200, 10)])
def test_other_swarm_sizes(n_a, steps):
    rollout_pair(n_a, 11 + n_a, "goal", steps)


@pytest.mark.parametrize("n_a,seed,mode,steps", [(30, 21, "goal", 150), (31, 22, "goal", 100), (7, 23, "goal", 100), (30, 24, "random", 100)])
def test_without_self_state_obs_dim_188(n_a, seed, mode, steps):
    """is_con_self_state=False (assembly.py:105-108, 795-801; AssemblyEnv.cpp:103-126): the own-state column is dropped,
    obs_dim = 188, every later row moves up by four."""
    st = rollout_pair(n_a, seed, mode, steps, self_state=False)
    if mode == "goal":
        assert st["in_shape"] > 50 and st["occupied"] > 0, st


@pytest.mark.parametrize("method,self_state", [("manual_rl", True), ("irl", False)])
def test_without_prior_training_method_not_llm_rl(method, self_state):
    """training_method != 'llm_rl' (assembly.py:605, 666): calculateActionPrior is never called, the 5th output is None."""
    st = rollout_pair(30, 31, "goal", 120, self_state=self_state, training_method=method)
    assert st["in_shape"] > 50, st


def test_ref_glue_matches_the_real_env_class():
    """oracle/ref_glue.RefEnv (restated NumPy glue + the reference's compiled C++) == the real AssemblySwarmEnv,
    including reset()'s NumPy-global RNG call order.  This is what bench.py --impl reference times on the GPU box."""
    from oracle import ref_glue
    from tests.helpers import load_shapes
    env = lr.make_env(30)
    glue = ref_glue.RefEnv(30, load_shapes())
    assert glue.r_avoid == env.env.r_avoid
    for seed in (226, 5):
        np.random.seed(seed); o1 = env.reset().copy()
        np.random.seed(seed); o2 = glue.reset().copy()
        assert np.array_equal(o1, o2) and np.array_equal(env.env.grid_center, glue.grid_center)
        rng = np.random.RandomState(seed)
        for t in range(100):
            a = goal_seeking_action(env.env.obs, env.env.dp, rng)
            r1 = env.step(a); r2 = glue.step(a)
            for x, y in zip((r1[0], r1[1], r1[2], r1[4]), (r2[0], r2[1], r2[2], r2[4])):
                assert np.array_equal(x, y)
            assert np.array_equal(env.env.p, glue.p) and np.array_equal(env.env.occupied_index, glue.occupied_index)


@pytest.mark.parametrize("seed", [7, 8])
def test_periodic_boundaries_match_reference(seed):
    """is_boundary=False -> periodic wrap (assembly.py:99-103, 447-448, 651-652; AssemblyEnv.cpp:88-90, 474-477, 700-732),
    including the reference's quirk that _get_dist_b2b only wraps agent 0's rows."""
    n_a = 30
    env = lr.make_env(n_a, is_boundary=False)
    np.random.seed(seed)
    env.reset()
    e = env.env
    assert e.is_periodic
    P = orc.make_params(n_a, e.n_g, float(e.l_cell), float(e.r_avoid), is_periodic=True)
    ob = orc.OracleBatch([P])
    ob.p[0], ob.dp[0] = e.p, e.dp
    ob.set_grid(0, e.grid_center)
    ob.observe()
    assert np.array_equal(ob.obs[0], e.obs) and np.array_equal(ob.neighbor_index[0], e.neighbor_index)
    rng = np.random.RandomState(seed)
    wrapped = 0
    for t in range(150):
        # outward drift so that agents cross the box edges and wrap
        a = np.clip(0.8 * np.sign(e.p) + rng.normal(0, 0.5, (2, n_a)), -1, 1).astype(np.float32)
        before = e.p.copy()
        obs, rew, done, info, prior = env.step(a)
        ob.step(a[None])
        wrapped += int((np.abs(e.p - before) > 2.0).sum())
        for name, x, y in (("p", ob.p[0], e.p), ("dp", ob.dp[0], e.dp), ("obs", ob.obs[0], obs), ("rew", ob.reward[0], rew),
                           ("prior", ob.a_prior[0], prior), ("nbr", ob.neighbor_index[0], e.neighbor_index),
                           ("sen", ob.sensed_index[0], e.sensed_index), ("occ", ob.occupied_index[0], e.occupied_index)):
            assert np.array_equal(x, y, equal_nan=True), f"{name} differs at step {t}"
    assert wrapped > 10


@pytest.mark.parametrize("strategy,n_a,seed", [("rule", 30, 4), ("rule", 12, 9), ("llm", 30, 6), ("llm", 8, 2)])
def test_strategy_restatement_tracks_the_numpy_original(strategy, n_a, seed):
    """agent_strategy 'rule' / 'llm' (assembly.py:519-601): the reference computes the action itself and returns it as the
    5th output when is_collected (assembly.py:663-664).  The C restatement cannot be bit-identical to NumPy (BLAS dot with
    FMA, pairwise np.sum, NumPy's own cos); it must agree to 1e-12 on the action of every step when fed the reference's
    own state, including states inside the shape (the controllers drive the swarm there)."""
    env = lr.make_env(n_a, agent_strategy=strategy, is_collected=True)
    np.random.seed(seed)
    env.reset()
    e = env.env
    P = orc.make_params(n_a, e.n_g, float(e.l_cell), float(e.r_avoid))
    ob = orc.OracleBatch([P])
    ob.set_grid(0, e.grid_center)
    worst, in_shape, sub = 0.0, 0, 0
    for t in range(150):
        ob.p[0], ob.dp[0] = e.p, e.dp                      # the reference's own pre-step state
        ob.neighbor_index[0] = e.neighbor_index
        mine = orc.strategy_actions(ob, strategy)[0]
        obs, rew, done, info, u = env.step(np.zeros((2, n_a)))
        worst = max(worst, float(np.max(np.abs(mine - u))))
        in_shape += int(e.in_flags.sum()); sub += int(((e.sensed_index >= 0).sum(1) == 80).sum())
    assert worst < 1e-12, worst
    assert in_shape > 0
    # directed state for the 80-cell subsample of the rule controller (np.round = half-even, assembly.py:565): agents
    # sitting on cells deep inside the shape with a small avoidance radius keep > 80 free cells in range
    e.r_avoid, e.d_sen = 0.05, 0.6
    P2 = orc.make_params(n_a, e.n_g, float(e.l_cell), 0.05, d_sen=0.6)
    ob2 = orc.OracleBatch([P2]); ob2.set_grid(0, e.grid_center)
    centre = e.grid_center.mean(axis=1, keepdims=True)
    order = np.argsort(np.linalg.norm(e.grid_center - centre, axis=0))[:n_a * 3:3]
    e.p = e.grid_center[:, order] + np.random.RandomState(seed).normal(0, 0.004, (2, n_a))
    e.dp = np.random.RandomState(seed + 1).uniform(-0.2, 0.2, (2, n_a))
    e._get_obs()                                            # refresh neighbor_index for the 'llm' strategy
    ob2.p[0], ob2.dp[0], ob2.neighbor_index[0] = e.p, e.dp, e.neighbor_index
    mine = orc.strategy_actions(ob2, strategy)[0]
    n_free = [(np.linalg.norm(e.grid_center - e.p[:, [i]], axis=0) < e.d_sen).sum() for i in range(n_a)]
    _, _, _, _, u = env.step(np.zeros((2, n_a)))
    assert float(np.max(np.abs(mine - u))) < 1e-12
    assert max(n_free) > 120                                # far above the cap of 80 even after the occupancy filter
