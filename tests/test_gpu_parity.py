"""Parity tests proper: the CUDA path, called through the C ABI, against (a) the golden trajectories recorded from the
unmodified reference and (b) the C oracle on seeded inputs.  Integer/index outputs, done flags and — because the
kernels reproduce the reference's fp64 operation order without FMA — positions, velocities, observations, rewards
and priors are all required to be BIT-EXACT (the north star only asks for 1e-5 relative on the floating-point ones)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import GOLDEN_CASES, goal_seeking_action, load_golden, load_shapes, replay_golden, reset_like_reference

pytestmark = pytest.mark.gpu


SCAN = {"mode": "lookup"}


@pytest.fixture(params=["lookup", "culled"], autouse=True)
def scan_mode(request):
    """Every test of this module runs twice: with the shape library installed (grids set through set_grid are recognised as
    rigid transforms of library shapes -> the lookup-scan kernel, where eligible) and without it (general culled scan)."""
    SCAN["mode"] = request.param
    yield request.param
    SCAN["mode"] = "lookup"


def make_sim(E, n_a, n_g_max, r_avoid, **kw):
    from marl_llm_b200.batched import BatchedAssemblySim
    shapes = load_shapes()
    lookup = SCAN["mode"] == "lookup"
    if lookup:
        n_g_max = max(n_g_max, int(shapes["n_g"].max()))
    sim = BatchedAssemblySim(E, n_a, n_g_max, r_avoid, **kw)
    if lookup:
        sim.set_shapes(shapes["grid_origin"], shapes["l_cell"])
    import os
    big_fits = not (kw.get("emit_indices", False) and n_a > 512)     # 1024 agents + index arrays: shared memory is full without the records
    radii_ok = 0.5 * r_avoid + 1e-6 < kw.get("d_sen", 0.4)          # covered cells are looked for among the sensing candidates
    sim.expect_fast = lookup and big_fits and radii_ok and not kw.get("brute_force_scan", False) and not (n_a <= 32 and os.environ.get("SWARM_FUSED_STEP"))
    return sim


def check_scan_mode(sim):
    """After the grids are in: the kernel that will run is the one this test variant is about."""
    assert bool(sim.fast_path) == sim.expect_fast, (sim.fast_path, SCAN["mode"])


def sim_snapshot(sim, e=None):
    sl = slice(None) if e is None else e
    f = lambda t: None if t is None else t[sl].cpu().numpy()   # noqa: E731
    return dict(p=f(sim.p), dp=f(sim.dp), obs=f(sim.obs), reward=f(sim.reward), a_prior=f(sim.a_prior),
                nbr=f(sim.neighbor_index), in_flags=f(sim.in_flags), sensed=f(sim.sensed_index), occupied=f(sim.occupied_index))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_trajectories_bit_exact(case):
    g = load_golden(case)
    n_a, n_g = int(g["n_a"]), int(g["n_g"])
    sim = make_sim(1, n_a, n_g, float(g["r_avoid"]), out_dtype=torch.float64, emit_indices=True, d_sen=float(g["d_sen"]),
                   is_periodic=bool(g.get("is_periodic", 0)))

    def reset_fn(g):
        blocks, ng = sim.pack_grids([g["grid_center"]], sim.n_g_max)
        sim.set_grid(blocks, ng, [float(g["l_cell"])])
        check_scan_mode(sim)
        sim.set_state(g["p0"][None], g["dp0"][None])
        sim.observe()
        return sim_snapshot(sim, 0)

    def step_fn(a):
        sim.step(torch.from_numpy(a[None]).cuda())
        return sim_snapshot(sim, 0)

    replay_golden(g, reset_fn, step_fn)
    assert not sim.done.any()


def build_batch(E, n_a, seed, shapes=None, shape_of=None, **param_kw):
    shapes = shapes or load_shapes()
    rng = np.random.RandomState(seed)
    r_avoid = orc.r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"])
    params, grids, P, DP = [], [], [], []
    for e in range(E):
        k, grid, p, dp = reset_like_reference(rng, n_a, shapes)
        params.append(orc.make_params(n_a, grid.shape[1], float(shapes["l_cell"][k]), r_avoid, **param_kw))
        grids.append(grid); P.append(p); DP.append(dp)
    return shapes, r_avoid, params, grids, np.stack(P), np.stack(DP)


def load_batch(sim, ob, params, grids, P, DP):
    for e, g in enumerate(grids):
        ob.set_grid(e, g)
    ob.p[:], ob.dp[:] = P, DP
    blocks, n_g = sim.pack_grids(grids, sim.n_g_max)
    sim.set_grid(blocks, n_g, [p_.l_cell for p_ in params])
    check_scan_mode(sim)
    sim.set_state(P, DP)


def compare_all(sim, ob, t, fields=("p", "dp", "obs", "reward", "a_prior", "nbr", "in_flags", "sensed", "occupied")):
    got = sim_snapshot(sim)
    ref = dict(p=ob.p, dp=ob.dp, obs=ob.obs, reward=ob.reward, a_prior=ob.a_prior, nbr=ob.neighbor_index,
               in_flags=ob.in_flags, sensed=ob.sensed_index, occupied=ob.occupied_index)
    for k in fields:
        if got[k] is None:
            continue
        if not np.array_equal(got[k], ref[k], equal_nan=True):
            bad = np.argwhere(got[k] != ref[k])
            raise AssertionError(f"{k} differs at step {t}: {len(bad)} elements, first at {bad[0]} "
                                 f"got {got[k][tuple(bad[0])]} want {ref[k][tuple(bad[0])]}")


@pytest.mark.parametrize("n_a,E,steps,mode", [(30, 256, 200, "mixed"), (30, 64, 200, "goal"), (7, 64, 60, "goal"),
                                              (33, 32, 60, "goal"), (100, 16, 40, "goal"), (1, 8, 10, "random"),
                                              (32, 16, 40, "goal"), (31, 16, 40, "mixed")])
def test_batch_vs_oracle_every_step(n_a, E, steps, mode):
    """Seeded batch, every output of every step compared with the oracle (parity layout: fp64 + index arrays).
    'goal' drives agents into the shapes so the in-shape / occupancy / 80-cell subsample / reward branches run."""
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=100 + n_a)
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe(with_reward=True)
    compare_all(sim, ob, -1, fields=("obs", "reward", "nbr", "in_flags", "sensed", "occupied"))
    rng = np.random.RandomState(5)
    in_shape = rewards = 0
    for t in range(steps):
        a_rand = rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        a_goal = goal_seeking_action(ob.obs, ob.dp, rng)
        if mode == "random":
            a = a_rand
        elif mode == "goal":
            a = a_goal
        else:
            a = np.where((np.arange(E) % 2 == 0)[:, None, None], a_goal, a_rand)
        sim.step(torch.from_numpy(a).cuda())
        ob.step(a)
        compare_all(sim, ob, t)
        in_shape += int(ob.in_flags.sum()); rewards += float(ob.reward.sum())
    if mode != "random" and n_a >= 7:
        assert in_shape > 0
    assert not sim.done.any()


@pytest.mark.parametrize("n_a,E,emit", [(30, 96, True), (30, 96, False), (31, 32, True), (7, 48, True), (7, 48, False), (100, 8, True)])
def test_without_self_state_obs_dim_188(n_a, E, emit):
    """is_con_self_state=False (assembly.py:105-108, 801; AssemblyEnv.cpp:103-126): obs_dim 188, the own-state rows are
    dropped and the first sensed-cell row is 28 instead of 32.  Both layouts (parity: fp64 + index arrays; production: fp32),
    >= 100 steps with goal-seeking envs so the in-shape / occupancy / subsample branches fire; the observation buffer is
    poisoned before every step.  The oracle is pinned to the live reference for this flag (tests/test_oracle_vs_reference.py)."""
    steps = 120
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=300 + n_a, is_con_self_state=False)
    ngm = int(shapes["n_g"].max())
    dt = torch.float64 if emit else torch.float32
    cast = np.float64 if emit else np.float32
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit, is_con_self_state=False)
    assert sim.obs_dim == 188 and sim.obs.shape == (E, 188, n_a)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.obs.fill_(float("nan"))
    sim.observe(); ob.observe(with_reward=True)
    assert np.array_equal(sim.obs.cpu().numpy(), ob.obs.astype(cast))
    rng = np.random.RandomState(17)
    in_shape = sub = 0
    for t in range(steps):
        a_goal = goal_seeking_action(ob.obs, ob.dp, rng, target_row=24)
        a = np.where((np.arange(E) % 4 != 0)[:, None, None], a_goal, rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32))
        sim.obs.fill_(float("nan"))
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        if emit:
            compare_all(sim, ob, t)
        else:
            for name, got, ref in (("p", sim.p, ob.p), ("dp", sim.dp, ob.dp), ("obs", sim.obs, ob.obs.astype(cast)),
                                   ("reward", sim.reward, ob.reward.astype(cast)), ("a_prior", sim.a_prior, ob.a_prior.astype(cast)),
                                   ("nbr", sim.neighbor_index, ob.neighbor_index), ("in_flags", sim.in_flags, ob.in_flags)):
                assert np.array_equal(got.cpu().numpy(), ref), (name, t)
        in_shape += int(ob.in_flags.sum()); sub += int(((ob.sensed_index >= 0).sum(2) == 80).sum())
    assert in_shape > 100 and ob.reward.sum() >= 0
    if n_a >= 30:
        assert sub > 0                                          # the 80-cell subsample branch ran


@pytest.mark.parametrize("self_state,emit", [(True, True), (True, False), (False, True)])
def test_without_prior_training_method_not_llm_rl(self_state, emit):
    """want_prior=False (training_method != 'llm_rl': assembly.py:605, 666): step() returns None as its 5th output and the
    a_prior buffers are never written (they keep a sentinel); everything else equals the oracle every step."""
    E, n_a, steps = 64, 30, 110
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=41, is_con_self_state=self_state, want_prior=False)
    ngm = int(shapes["n_g"].max())
    dt = torch.float64 if emit else torch.float32
    cast = np.float64 if emit else np.float32
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit, is_con_self_state=self_state, want_prior=False)
    for b in sim._a_prior:
        b.fill_(-123.0)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe(with_reward=True)
    rng = np.random.RandomState(19)
    l0 = sim.launch_count
    for t in range(steps):
        a = goal_seeking_action(ob.obs, ob.dp, rng, target_row=28 if self_state else 24)
        out = sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        assert out[4] is None
        if t == 50:                                             # a state poke must not resurrect the stand-alone prior kernel
            sim.set_state(ob.p, ob.dp)
        fields = ("p", "dp", "obs", "reward", "nbr", "in_flags") + (("sensed", "occupied") if emit else ())
        got = sim_snapshot(sim)
        ref = dict(p=ob.p, dp=ob.dp, obs=ob.obs.astype(cast), reward=ob.reward.astype(cast), nbr=ob.neighbor_index,
                   in_flags=ob.in_flags, sensed=ob.sensed_index, occupied=ob.occupied_index)
        for k in fields:
            assert np.array_equal(got[k], ref[k]), (k, t)
    assert ob.in_flags.sum() > 0 and not ob.a_prior.any()
    assert sim.launch_count - l0 == 2 * steps                   # two launches per step, no k_prior launch
    for b in sim._a_prior:
        assert bool((b == -123.0).all())


def test_config2_4096_envs_200_steps():
    """BASELINE config 2: 30 agents x 4096 envs x 200 steps.  Every env is compared with the oracle at the steps
    the oracle budget allows (all envs at 10 checkpoints incl. the last, a 128-env slice at every step)."""
    E, n_a, steps = 4096, 30, 200
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=2)
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    ob = orc.OracleBatch(params, nthreads=16)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe()
    compare_all(sim, ob, -1, fields=("obs", "nbr", "in_flags", "sensed", "occupied"))
    act = torch.empty(E, 2, n_a, dtype=torch.float32, device="cuda")
    rng = np.random.RandomState(9)
    for t in range(steps):
        sim.fill_actions(act, seed=226, step=t)
        a = orc.fill_actions(E, n_a, 226, t)
        # half the envs follow a noisy goal-seeking controller so shapes get populated
        a_goal = goal_seeking_action(ob.obs, ob.dp, rng)
        a = np.where((np.arange(E) % 2 == 1)[:, None, None], a_goal, a)
        act.copy_(torch.from_numpy(a))
        sim.step(act)
        ob.step(a)
        if t % 20 == 19 or t == steps - 1:
            compare_all(sim, ob, t)
        else:
            got = sim_snapshot(sim, slice(0, 128))
            assert np.array_equal(got["p"], ob.p[:128]) and np.array_equal(got["obs"], ob.obs[:128]), t
            assert np.array_equal(got["nbr"], ob.neighbor_index[:128]) and np.array_equal(got["sensed"], ob.sensed_index[:128]), t
    assert ob.in_flags.sum() > 1000 and ob.reward.sum() > 0


def test_fp32_outputs_are_the_rounded_fp64_ones_and_state_stays_fp64():
    E, n_a = 64, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=31)
    ngm = int(shapes["n_g"].max())
    s64 = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    s32 = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float32, emit_indices=False)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(s64, ob, params, grids, P, DP)
    load_batch(s32, ob, params, grids, P, DP)
    s64.observe(); s32.observe(); ob.observe()
    rng = np.random.RandomState(1)
    for t in range(60):
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        ta = torch.from_numpy(a).cuda()
        s64.step(ta); s32.step(ta); ob.step(a)
        assert torch.equal(s32.p, s64.p) and torch.equal(s32.dp, s64.dp)
        assert torch.equal(s32.obs, s64.obs.float()) and torch.equal(s32.reward, s64.reward.float())
        assert torch.equal(s32.a_prior, s64.a_prior.float())
        assert torch.equal(s32.neighbor_index, s64.neighbor_index) and torch.equal(s32.in_flags, s64.in_flags)
    assert s32.sensed_index is None and s32.obs.dtype == torch.float32 and s32.p.dtype == torch.float64


def test_exact_occupancy_path_equals_shared_mask_path():
    """The literal per-agent sequential occupancy filter (taken when an agent sits in the rounding shell of the
    nearby radius) and the shared covered-mask shortcut must give identical lists."""
    E, n_a = 96, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=77)
    ngm = int(shapes["n_g"].max())
    fast = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    slow = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True, exact_occupancy=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(fast, ob, params, grids, P, DP); load_batch(slow, ob, params, grids, P, DP)
    fast.observe(); slow.observe(); ob.observe()
    rng = np.random.RandomState(3)
    for t in range(80):
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        ta = torch.from_numpy(a).cuda()
        fast.step(ta); slow.step(ta); ob.step(a)
        compare_all(slow, ob, t)
        for name in ("obs", "reward", "sensed_index", "occupied_index"):
            assert torch.equal(getattr(fast, name), getattr(slow, name)), (name, t)
    assert (ob.occupied_index >= 0).sum() > 1000


def test_reward_estimate_decides_like_the_exact_sums():
    """The reward predicate |v| < 0.05 (CPP:495-552) is normally decided by an fp32 estimate with an error bound and only
    falls back to the reference's fp64 psi sums when undecided; `exact_reward_sums` forces the fp64 path for every agent.
    Both must give the oracle's rewards on a batch where most agents sit inside their shapes for many steps."""
    E, n_a = 128, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=23)
    rng = np.random.RandomState(5)
    for e in range(E):      # start assembled: every agent on a cell of its shape
        idx = rng.choice(grids[e].shape[1], n_a, replace=False)
        P[e] = grids[e][:, idx] + rng.normal(0, 0.01, (2, n_a))
    ngm = int(shapes["n_g"].max())
    est = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    exact = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True, exact_reward_sums=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(est, ob, params, grids, P, DP); load_batch(exact, ob, params, grids, P, DP)
    est.observe(); exact.observe(); ob.observe(with_reward=True)
    assert np.array_equal(est.reward.cpu().numpy(), ob.reward) and np.array_equal(exact.reward.cpu().numpy(), ob.reward)
    ones = 0
    for t in range(80):
        a = goal_seeking_action(ob.obs, ob.dp, rng, noise=0.15)
        ta = torch.from_numpy(a).cuda()
        est.step(ta); exact.step(ta); ob.step(a)
        assert np.array_equal(est.reward.cpu().numpy(), ob.reward), t
        assert np.array_equal(exact.reward.cpu().numpy(), ob.reward), t
        assert torch.equal(est.obs, exact.obs)
        ones += int(ob.reward.sum())
    assert ob.in_flags.mean() > 0.5 and ones > 1000          # both outcomes of the predicate occurred many times


@pytest.mark.parametrize("n_a,E,emit,dt", [(30, 64, False, torch.float32), (30, 32, True, torch.float64), (100, 8, False, torch.float32), (7, 16, False, torch.float32)])
def test_agent_major_layout_is_the_transposed_reference_layout(n_a, E, emit, dt):
    """obs_layout='agent_major' ([E, n_a, obs_dim]: one contiguous row per agent, for device-side consumers) holds exactly the
    values of the reference layout [E, obs_dim, n_a] (CPP:324-328), every step; everything else is unaffected.  The buffer is
    poisoned before every step (zero-fill + scattered emission use different address arithmetic in this layout)."""
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=60 + n_a)
    ngm = int(shapes["n_g"].max())
    ref = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit)
    am = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit, obs_layout="agent_major")
    assert am.obs.shape == (E, n_a, ref.obs_dim)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(ref, ob, params, grids, P, DP); load_batch(am, ob, params, grids, P, DP)
    am.obs.fill_(float("nan"))
    ref.observe(); am.observe(); ob.observe()
    assert torch.equal(am.obs, ref.obs.transpose(1, 2))
    rng = np.random.RandomState(3)
    for t in range(70):
        a = goal_seeking_action(ob.obs, ob.dp, rng) if t % 3 else rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        ta = torch.from_numpy(a).cuda()
        am.obs.fill_(float("nan"))
        ref.step(ta); am.step(ta); ob.step(a)
        assert torch.equal(am.obs, ref.obs.transpose(1, 2)), t
        assert np.array_equal(ref.obs.cpu().numpy(), ob.obs.astype(np.float64 if emit else np.float32)), t
        for name in ("p", "dp", "reward", "a_prior", "neighbor_index", "in_flags"):
            assert torch.equal(getattr(am, name), getattr(ref, name)), (name, t)
        if emit:
            assert torch.equal(am.sensed_index, ref.sensed_index) and torch.equal(am.occupied_index, ref.occupied_index)
    assert ob.in_flags.sum() > 0


def test_lookup_scan_equals_culled_scan_at_scale():
    """Two independent algorithms for the second half of the step — the lookup scan (per-shape tables, lattice row records) and
    the word-box culled scan — on 16 384 envs x 100 steps, half of the envs driven into their shapes: every output identical at
    every 5th step, and 64 sampled envs follow the oracle."""
    if SCAN["mode"] != "lookup":
        pytest.skip("one run is enough: the test builds both kinds of simulator itself")
    import bench
    from marl_llm_b200.batched import BatchedAssemblySim
    E, n_a, steps = 16384, 30, 100
    shapes = load_shapes()
    ngm = int(shapes["n_g"].max())
    r_avoid = orc.r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"])
    blocks, n_g, l_cell, p, dp = bench.synth_batch(E, n_a, shapes, 13, "random")
    look = BatchedAssemblySim(E, n_a, ngm, r_avoid)
    look.set_shapes(shapes["grid_origin"], shapes["l_cell"])
    cull = BatchedAssemblySim(E, n_a, ngm, r_avoid)
    for s_ in (look, cull):
        s_.set_grid(blocks, n_g, l_cell); s_.set_state(p, dp); s_.observe()
    assert look.fast_path == 1 and cull.fast_path == 0
    pick = np.arange(0, E, E // 64)
    ob = orc.OracleBatch([orc.make_params(n_a, int(n_g[e]), float(l_cell[e]), r_avoid) for e in pick], nthreads=16)
    for k, e in enumerate(pick):
        ob.set_grid(k, blocks[e, :2 * n_g[e]].reshape(2, n_g[e]))
    ob.p[:], ob.dp[:] = p[pick], dp[pick]
    ob.observe()
    tp = torch.from_numpy(pick).cuda()
    act = torch.empty(E, 2, n_a, dtype=torch.float32, device="cuda")
    goal = (torch.arange(E, device="cuda") % 2 == 0)[:, None, None]
    for t in range(steps):
        look.fill_actions(act, seed=5, step=t)
        row = 28
        a_goal = (3.0 * look.obs[:, row:row + 2, :] - look.dp.float()).clamp(-1, 1)      # noise-free goal seeking from the device obs
        act = torch.where(goal, a_goal, act).contiguous()
        look.step(act); cull.step(act)
        ob.step(act[tp].cpu().numpy())
        assert np.array_equal(look.obs[tp].cpu().numpy(), ob.obs.astype(np.float32)), t
        assert np.array_equal(look.reward[tp].cpu().numpy(), ob.reward.astype(np.float32)), t
        if t % 5 == 4 or t == steps - 1:
            for name in ("p", "dp", "obs", "reward", "a_prior", "neighbor_index", "in_flags", "nearest_cell"):
                assert torch.equal(getattr(look, name), getattr(cull, name)), (name, t)
    assert float(look.in_flags.float().mean()) > 0.2 and float(look.reward.sum()) > 1000


def synthetic_shapes():
    """Lattice shapes that stress the lookup tables: a ring (an agent at its centre has dozens of equidistant nearest cells ->
    spilled candidate lists), scattered dots (holes inside rows), one long row, a 2-cell shape, a wide block (40 columns), a disc,
    denser dots, a tall block, an L, a small ring."""
    L = 0.06
    def cells(mask):
        iy, ix = np.nonzero(mask)                                  # row-major order = the numbering of assembly_cfg.py:60-79
        pts = np.stack([ix * L, iy * L]).astype(np.float64)
        return np.ascontiguousarray(pts - pts.mean(axis=1, keepdims=True))
    yy, xx = np.mgrid[0:33, 0:33]
    r = np.hypot(xx - 16, yy - 16)
    rng = np.random.RandomState(0)
    out = [cells((r > 11.5) & (r < 13.5)), cells(rng.rand(30, 30) < 0.25), cells(np.ones((1, 50), bool)),
           cells(np.array([[1, 0, 0, 1]], bool)), cells(np.ones((12, 40), bool))]
    # ten shapes in all: the kernels read the first eight tables from their parameters and the rest from global memory
    ell = np.zeros((24, 24), bool); ell[:, :5] = True; ell[-5:, :] = True
    out += [cells(r < 9.5), cells(rng.rand(20, 44) < 0.5), cells(np.ones((40, 12), bool)), cells(ell), cells((r > 5.5) & (r < 8.5))]
    return out, [L] * len(out)


def fine_shapes():
    """A fine lattice (l_cell = 0.03: the sensing disc spans 27 rows and ~560 cells): 32 row lanes per agent, more than 255
    candidates per agent (the provisional-slot field of a row record saturates), every agent in range senses more than 80 cells."""
    L = 0.03
    def cells(mask):
        iy, ix = np.nonzero(mask)
        pts = np.stack([ix * L, iy * L]).astype(np.float64)
        return np.ascontiguousarray(pts - pts.mean(axis=1, keepdims=True))
    yy, xx = np.mgrid[0:40, 0:40]
    r = np.hypot(xx - 19.5, yy - 19.5)
    out = [cells(np.ones((30, 30), bool)), cells((r > 9) & (r < 19.5)), cells(np.ones((16, 60), bool))]
    return out, [L] * len(out)


@pytest.mark.parametrize("n_a", [30, 64])
def test_lookup_scan_on_synthetic_lattice_shapes(n_a):
    """The lookup scan against the oracle on a shape library built to stress its tables (see synthetic_shapes): poses through
    set_grid (detected) and set_grid_pose (exact), agents spawned inside, on the rim and around each shape."""
    origins, l_cells = synthetic_shapes()
    run_synthetic_library(origins, l_cells, n_a, E=40, steps=40, min_sensed=1000)


def test_lookup_scan_on_a_fine_lattice():
    """l_cell = 0.03 (see fine_shapes): 32 row lanes per agent, saturated provisional slots, subsampled lists outside the shape."""
    origins, l_cells = fine_shapes()
    run_synthetic_library(origins, l_cells, 30, E=12, steps=15, min_sensed=5000)


def run_synthetic_library(origins, l_cells, n_a, E, steps, min_sensed):
    from marl_llm_b200.batched import BatchedAssemblySim
    ngm = max(o.shape[1] for o in origins)
    r_avoid = 0.2
    rng = np.random.RandomState(n_a)
    k = np.arange(E) % len(origins)
    ang = np.pi * rng.uniform(-1, 1, E)
    cs, sn, off = np.cos(ang), np.sin(ang), rng.uniform(-1.0, 1.0, (E, 2))
    grids = [BatchedAssemblySim.grid_from_pose(origins[k[e]], cs[e], sn[e], off[e, 0], off[e, 1]) for e in range(E)]
    params = [orc.make_params(n_a, grids[e].shape[1], l_cells[k[e]], r_avoid) for e in range(E)]
    P = np.empty((E, 2, n_a)); DP = rng.uniform(-0.3, 0.3, (E, 2, n_a))
    for e in range(E):
        centre = off[e][:, None]
        P[e] = centre + rng.normal(0, 0.5, (2, n_a))                                   # around / inside the shape
        pick = rng.choice(grids[e].shape[1], n_a // 3, replace=True)
        P[e][:, :n_a // 3] = grids[e][:, pick] + rng.normal(0, 0.01, (2, n_a // 3))      # on cells
        P[e][:, -1] = centre[:, 0]                                                      # exactly at the centre (ring: all cells equidistant-ish)
    sims = []
    for mode in ("detected", "exact"):
        sim = BatchedAssemblySim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
        sim.set_shapes(origins, l_cells)
        if mode == "exact":
            sim.set_grid_pose(k, cs, sn, off[:, 0], off[:, 1])
        else:
            blocks, n_g = sim.pack_grids(grids, ngm)
            sim.set_grid(blocks, n_g, [l_cells[j] for j in k])
        assert sim.fast_path == (2 if mode == "exact" else 1), (mode, sim.fast_path)
        sim.set_state(P, DP)
        sims.append(sim)
    ob = orc.OracleBatch(params, nthreads=8, ng_max=ngm)
    for e in range(E):
        ob.set_grid(e, grids[e])
    ob.p[:], ob.dp[:] = P, DP
    for s_ in sims:
        s_.observe()
    ob.observe(with_reward=True)
    for s_ in sims:
        compare_all(s_, ob, -1, fields=("obs", "reward", "nbr", "in_flags", "sensed", "occupied"))
    for t in range(steps):
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        ob.step(a)
        for s_ in sims:
            s_.step(torch.from_numpy(a).cuda())
            compare_all(s_, ob, t)
    assert ob.in_flags.sum() > 0 and (ob.sensed_index >= 0).sum() > min_sensed


def test_agents_far_outside_the_arena_take_the_literal_scan():
    """The lookup scan's bin table covers the arena plus a margin; a caller can still put agents anywhere (`env.p = ...`).
    Positions far outside (|p| up to 9) select the literal all-cells scan for those agents (CPP:869-885): same outputs as the
    oracle, including the walls' spring pulling them back for a few steps."""
    E, n_a = 32, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=71)
    rng = np.random.RandomState(4)
    P = P.copy()
    P[:, :, ::3] = rng.uniform(-9, 9, P[:, :, ::3].shape)            # every third agent anywhere in an 18 x 18 square
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe(with_reward=True)
    compare_all(sim, ob, -1, fields=("obs", "reward", "nbr", "in_flags", "sensed", "occupied"))
    for t in range(12):
        a = rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        compare_all(sim, ob, t)
    assert np.abs(ob.p).max() > 3.0


def test_grid_swap_between_steps_like_eval_script():
    """eval_assembly.py:34-57 overwrites env.grid_center / l_cell / n_g between steps; the next step's prior must come
    from the NEW grid and the OLD neighbour list (assembly.py:613-624), the observation from the new grid."""
    E, n_a = 16, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=55)
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    ob = orc.OracleBatch(params, nthreads=4, ng_max=ngm)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe()
    rng = np.random.RandomState(8)
    for t in range(30):
        if t in (10, 20):
            _, _, params2, grids2, _, _ = build_batch(E, n_a, seed=1000 + t)
            for e in range(E):
                ob.params[e].n_g, ob.params[e].l_cell = params2[e].n_g, params2[e].l_cell
                ob.set_grid(e, grids2[e])
            blocks, n_g = sim.pack_grids(grids2, ngm)
            sim.set_grid(blocks, n_g, [p_.l_cell for p_ in params2])
        if t == 15:   # the caller teleports the agents (env.p = ...): prior uses new p with the stale neighbour list
            newp = ob.p + rng.normal(0, 0.05, ob.p.shape)
            ob.p[:] = newp
            sim.set_state(newp, ob.dp)
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        compare_all(sim, ob, t)


def test_step_host_matches_device_step():
    E, n_a = 32, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=66)
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float32)
    ob = orc.OracleBatch(params, nthreads=4)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe()
    obs_h = torch.empty(E, 192, n_a, dtype=torch.float32).pin_memory()
    rew_h = torch.empty(E, 1, n_a, dtype=torch.float32).pin_memory()
    pri_h = torch.empty(E, 2, n_a, dtype=torch.float32).pin_memory()
    for t in range(10):
        a = orc.fill_actions(E, n_a, 4, t)
        sim.step_host(a, obs_h, rew_h, pri_h)
        ob.step(a)
        assert np.array_equal(obs_h.numpy(), ob.obs.astype(np.float32))
        assert np.array_equal(rew_h.numpy(), ob.reward.astype(np.float32))
        assert np.array_equal(pri_h.numpy(), ob.a_prior.astype(np.float32))


def test_action_generator_matches_oracle():
    sim = make_sim(128, 30, 64, 0.26)
    act = torch.empty(128, 2, 30, dtype=torch.float32, device="cuda")
    sim.fill_actions(act, seed=226, step=17, env_offset=4096)
    assert np.array_equal(act.cpu().numpy(), orc.fill_actions(128, 30, 226, 17, 4096))


# ------------------------------------------------------------------------------------------------------------------
# Legacy five-symbol ABI (host pointers), called exactly the way assembly.py calls the reference library.
# ------------------------------------------------------------------------------------------------------------------
def _c(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("n_a", [30, 5, 200])
def test_legacy_symbols_match_oracle(n_a):
    from marl_llm_b200 import _lib
    lib = _lib.load()
    shapes = load_shapes()
    rng = np.random.RandomState(n_a)
    r_avoid = orc.r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"])
    k, grid, p, dp = reset_like_reference(rng, n_a, shapes)
    # put a third of the agents onto cells so that in-shape branches fire
    idx = rng.choice(grid.shape[1], n_a // 3 + 1, replace=False)
    p[:, :len(idx)] = grid[:, idx] + rng.normal(0, 0.01, (2, len(idx)))
    P = orc.make_params(n_a, grid.shape[1], float(shapes["l_cell"][k]), r_avoid)
    ob = orc.OracleBatch([P])
    ob.p[0], ob.dp[0] = p, dp
    ob.set_grid(0, grid)
    ob.observe(with_reward=True)

    n_g = grid.shape[1]
    obs = np.zeros((192, n_a)); nbr = -np.ones((n_a, 6), np.int32); inf = np.zeros(n_a, np.int32)
    sen = -np.ones((n_a, 80), np.int32); occ = -np.ones((n_a, 200), np.int32)
    heading = np.zeros((2, n_a)); bp = np.array([-2.4, 2.4, 2.4, -2.4])
    cond = np.array([False, True, True, False])
    dbl, i32, bl = C.c_double, C.c_int32, C.c_bool
    lib._get_observation(_c(p, dbl), _c(dp, dbl), _c(heading, dbl), _c(obs, dbl), _c(bp, dbl), _c(grid, dbl), _c(nbr, i32),
                         _c(inf, i32), _c(sen, i32), _c(occ, i32), C.c_double(0.4), C.c_double(r_avoid), C.c_double(P.l_cell),
                         C.c_double(0.8), C.c_int(6), C.c_int(80), C.c_int(200), C.c_int(n_a), C.c_int(n_g), C.c_int(192),
                         C.c_int(2), _c(cond, bl))
    assert np.array_equal(obs, ob.obs[0]) and np.array_equal(nbr, ob.neighbor_index[0]) and np.array_equal(inf, ob.in_flags[0])
    assert np.array_equal(sen, ob.sensed_index[0]) and np.array_equal(occ, ob.occupied_index[0])
    assert inf.sum() > 0

    rew = np.zeros((1, n_a)); act = np.zeros((2, n_a)); coef = np.array([0.05])
    cond5 = np.array([False, True, True, True, True])
    cb2b = np.zeros((n_a, n_a), bool); cb2w = np.zeros((4, n_a), bool)
    lib._get_reward(_c(p, dbl), _c(dp, dbl), _c(heading, dbl), _c(act, dbl), _c(rew, dbl), _c(bp, dbl), _c(grid, dbl),
                    _c(nbr, i32), _c(inf, i32), _c(sen, i32), _c(occ, i32), C.c_double(0.4), C.c_double(r_avoid),
                    C.c_double(P.l_cell), C.c_int(6), C.c_int(80), C.c_int(200), C.c_int(n_a), C.c_int(n_g), C.c_int(2),
                    _c(cond5, bl), _c(cb2b, bl), _c(cb2w, bl), _c(coef, dbl))
    assert np.array_equal(rew, ob.reward[0])

    prior = np.zeros((2, n_a))
    lib.calculateActionPrior(_c(p, dbl), _c(dp, dbl), _c(prior, dbl), _c(grid, dbl), _c(nbr, i32), C.c_double(0.4),
                             C.c_double(r_avoid), C.c_double(P.l_cell), C.c_int(6), C.c_int(n_a), C.c_int(n_g), C.c_int(2))
    ref_prior = np.zeros((2, n_a))
    orc.lib().orc_prior(C.byref(P), _c(p, dbl), _c(dp, dbl), _c(np.ascontiguousarray(grid), dbl), _c(ob.neighbor_index[0], i32),
                        _c(ref_prior, dbl))
    assert np.array_equal(prior, ref_prior)

    # NumPy glue of assembly.py:442-457 feeding _sf_b2b_all, and _get_dist_b2w (assembly.py:460-466)
    pc = p.copy(); pc[:, 1] = pc[:, 0] + np.array([0.03, 0.02]); pc[:, 2] = pc[:, 0] - np.array([0.01, 0.05]) if n_a > 2 else pc[:, 2]
    pc[0, -1] = 2.39; pc[1, -1] = -2.38
    all_pos = np.tile(pc, (n_a, 1)); my_pos = np.tile(pc.T.reshape(2 * n_a, 1), (1, n_a))
    rel = all_pos - my_pos
    center = np.sqrt(rel[::2] ** 2 + rel[1::2] ** 2)
    size = np.full(n_a, 0.035); sizes = size[:, None] + size[None, :]; sizes[np.arange(n_a), np.arange(n_a)] = 0
    edge = center - sizes; coll = edge < 0; edge = np.abs(edge)
    sf = np.zeros((2, n_a))
    lib._sf_b2b_all(_c(pc, dbl), _c(sf, dbl), _c(edge, dbl), _c(coll, bl), _c(bp, dbl), _c(center, dbl), C.c_int(n_a), C.c_int(2),
                    C.c_double(30.0), C.c_bool(False))
    ref_sf = np.zeros((2, n_a))
    orc.lib().orc_ball_forces(C.byref(P), _c(pc, dbl), _c(ref_sf, dbl))
    assert np.array_equal(sf, ref_sf) and np.abs(sf).sum() > 0
    d_b2w = np.ones((4, n_a)); cw = np.zeros((4, n_a), bool)
    lib._get_dist_b2w(_c(pc, dbl), _c(size, dbl), _c(d_b2w, dbl), _c(cw, bl), C.c_int(2), C.c_int(n_a), _c(bp, dbl))
    raw = np.stack([pc[0] - size - bp[0], bp[1] - (pc[1] + size), bp[2] - (pc[0] + size), pc[1] - size - bp[3]])
    assert np.array_equal(d_b2w, np.abs(raw)) and np.array_equal(cw, raw < 0) and cw.any()


def test_large_swarm_1024_agents_matches_oracle():
    """BASELINE config 4 shape (1024 agents per env, r_avoid = 0.05): the CTA-per-env variant (1024 threads) against the
    oracle, every output, a few steps (the oracle is O(n_a^2) per env)."""
    E, n_a = 3, 1024
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=4)
    assert r_avoid == 0.05
    rng = np.random.RandomState(12)
    for e in range(E):      # put a third of each swarm onto the shape so in-shape / occupancy / reward branches run
        idx = rng.choice(grids[e].shape[1], n_a // 3, replace=False)
        P[e][:, :n_a // 3] = grids[e][:, idx] + rng.normal(0, 0.01, (2, n_a // 3))
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    ob = orc.OracleBatch(params, nthreads=E)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe(with_reward=True)
    compare_all(sim, ob, -1, fields=("obs", "reward", "nbr", "in_flags", "sensed", "occupied"))
    for t in range(4):
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        compare_all(sim, ob, t)
    assert ob.in_flags.sum() > 100 and (ob.occupied_index >= 0).sum() > 100


def test_large_swarm_1024_agents_production_layout():
    """BASELINE config 4 in the layout the bench times (fp32 outputs, no index arrays): every output the layout has, against
    the oracle, from a state with a third of each swarm on its shape (in-shape / occupancy / subsample / reward branches).  In
    the lookup variant this is the multi-warp lookup-scan kernel (32 warps per env, each with its own row records)."""
    E, n_a = 3, 1024
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=4)
    rng = np.random.RandomState(13)
    for e in range(E):
        idx = rng.choice(grids[e].shape[1], n_a // 3, replace=False)
        P[e][:, :n_a // 3] = grids[e][:, idx] + rng.normal(0, 0.01, (2, n_a // 3))
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float32, emit_indices=False)
    ob = orc.OracleBatch(params, nthreads=E)
    load_batch(sim, ob, params, grids, P, DP)
    sim.obs.fill_(float("nan"))
    sim.observe(); ob.observe(with_reward=True)
    assert np.array_equal(sim.obs.cpu().numpy(), ob.obs.astype(np.float32))
    for t in range(4):
        a = goal_seeking_action(ob.obs, ob.dp, rng)
        sim.obs.fill_(float("nan"))
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        for name, got, ref in (("p", sim.p, ob.p), ("dp", sim.dp, ob.dp), ("obs", sim.obs, ob.obs.astype(np.float32)),
                               ("reward", sim.reward, ob.reward.astype(np.float32)), ("a_prior", sim.a_prior, ob.a_prior.astype(np.float32)),
                               ("nbr", sim.neighbor_index, ob.neighbor_index), ("in_flags", sim.in_flags, ob.in_flags)):
            assert np.array_equal(got.cpu().numpy(), ref), (name, t)
    assert ob.in_flags.sum() > 100


def test_periodic_boundaries_batch_vs_oracle():
    """is_boundary=False: wrap of relative positions in k-NN / obs / reward and of p after the
    integrator, no wall forces.  The oracle's periodic mode is pinned to the live reference in the build container."""
    E, n_a = 64, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=91)
    for p_ in params:
        p_.is_periodic = 1
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True, is_periodic=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe(with_reward=True)
    compare_all(sim, ob, -1, fields=("obs", "reward", "nbr", "in_flags", "sensed", "occupied"))
    rng = np.random.RandomState(2)
    wrapped = 0
    for t in range(120):
        drift = np.clip(0.8 * np.sign(ob.p) + rng.normal(0, 0.5, ob.p.shape), -1, 1).astype(np.float32)
        a = np.where((np.arange(E) % 2 == 0)[:, None, None], drift, goal_seeking_action(ob.obs, ob.dp, rng))
        before = ob.p.copy()
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
        wrapped += int((np.abs(ob.p - before) > 2.0).sum())
        compare_all(sim, ob, t)
    assert wrapped > 100


@pytest.mark.parametrize("n_a", [30, 100])
def test_culled_scan_equals_brute_force_scan_and_ignores_seed_content(n_a):
    """The word-box culled grid scan (default) against the all-pairs scan and the oracle, with the nearest-cell seed buffer
    deliberately filled with garbage before some steps (any content must be a valid seed)."""
    E = 48
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=500 + n_a)
    ngm = int(shapes["n_g"].max())
    fast = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True)
    brute = make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True, brute_force_scan=True)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(fast, ob, params, grids, P, DP); load_batch(brute, ob, params, grids, P, DP)
    fast.nearest_cell.random_(-5, 5000)                       # garbage seeds, including out-of-range ones
    fast.observe(); brute.observe(); ob.observe()
    rng = np.random.RandomState(6)
    for t in range(60):
        a = np.where((np.arange(E) % 2 == 0)[:, None, None], goal_seeking_action(ob.obs, ob.dp, rng),
                     rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32))
        if t % 7 == 3:
            fast.nearest_cell.random_(-5, 5000)
        ta = torch.from_numpy(a).cuda()
        fast.step(ta); brute.step(ta); ob.step(a)
        compare_all(fast, ob, t)
        for name in ("obs", "reward", "sensed_index", "occupied_index", "nearest_cell", "in_flags"):
            assert torch.equal(getattr(fast, name), getattr(brute, name)), (name, t)


@pytest.mark.gpu
@pytest.mark.parametrize("emit", [False, True])
def test_every_output_element_is_rewritten_each_step(emit):
    """The observation buffer is poisoned before every step (NaN / garbage indices): every element must be rewritten.
    Guards the zero-fill + speculative in-scan emission of the sensed-cell rows (stale slots when an agent enters the
    shape, loses cells to the occupancy filter, or crosses the 80-cell subsample threshold).  Production layout
    (fp32, no index arrays) and parity layout, mixed random / goal-seeking envs, compared with the oracle every step."""
    E, n_a, steps = 192, 30, 150
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=77)
    ngm = int(shapes["n_g"].max())
    dt = torch.float64 if emit else torch.float32
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.obs.fill_(float("nan"))
    sim.observe(); ob.observe()
    cast = np.float64 if emit else np.float32
    assert np.array_equal(sim.obs.cpu().numpy(), ob.obs.astype(cast))
    rng = np.random.RandomState(3)
    for t in range(steps):
        a_rand = rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        a_goal = goal_seeking_action(ob.obs, ob.dp, rng)
        a = np.where((np.arange(E) % 3 != 0)[:, None, None], a_goal, a_rand)
        sim.obs.fill_(float("nan")); sim.reward.fill_(float("nan"))
        if emit:
            sim.sensed_index.fill_(-7); sim.occupied_index.fill_(-7); sim.neighbor_index.fill_(-7)
        sim.step(torch.from_numpy(a).cuda())
        ob.step(a)
        assert np.array_equal(sim.obs.cpu().numpy(), ob.obs.astype(cast)), t
        assert np.array_equal(sim.reward.cpu().numpy(), ob.reward.astype(cast)), t
        assert np.array_equal(sim.a_prior.cpu().numpy(), ob.a_prior.astype(cast)), t
        assert np.array_equal(sim.neighbor_index.cpu().numpy(), ob.neighbor_index), t
        if emit:
            assert np.array_equal(sim.sensed_index.cpu().numpy(), ob.sensed_index), t
            assert np.array_equal(sim.occupied_index.cpu().numpy(), ob.occupied_index), t
    assert ob.in_flags.sum() > 100 and ob.reward.sum() > 0


def test_two_launch_step_equals_fused_step(monkeypatch):
    """Single-warp envs run the step as two launches (k_step PH 1 + PH 2); SWARM_FUSED_STEP=1 keeps the one-launch kernel.
    Same inputs -> identical state, outputs and index arrays, every step (both layouts of the emission: parity + production)."""
    E, n_a = 96, 30
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=55)
    ngm = int(shapes["n_g"].max())
    sims = []
    for fused in (False, True):
        if fused:
            monkeypatch.setenv("SWARM_FUSED_STEP", "1")
        else:
            monkeypatch.delenv("SWARM_FUSED_STEP", raising=False)
        pair = [make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, emit_indices=True),
                make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float32, emit_indices=False)]
        sims.append(pair)
    monkeypatch.delenv("SWARM_FUSED_STEP", raising=False)
    ob = orc.OracleBatch(params, nthreads=8)
    for pair in sims:
        for s_ in pair:
            load_batch(s_, ob, params, grids, P, DP)
            s_.observe()
    ob.observe()
    rng = np.random.RandomState(2)
    l0 = [[s_.launch_count for s_ in pair] for pair in sims]
    for t in range(50):
        a = goal_seeking_action(ob.obs, ob.dp, rng) if t % 2 else rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        ta = torch.from_numpy(a).cuda()
        for pair in sims:
            for s_ in pair:
                s_.step(ta)
        ob.step(a)
        for k in range(2):
            a_, b_ = sims[0][k], sims[1][k]
            for name in ("p", "dp", "obs", "reward", "a_prior", "neighbor_index", "in_flags", "nearest_cell"):
                assert torch.equal(getattr(a_, name), getattr(b_, name)), (name, t, k)
        assert torch.equal(sims[0][0].sensed_index, sims[1][0].sensed_index) and torch.equal(sims[0][0].occupied_index, sims[1][0].occupied_index)
        assert np.array_equal(sims[0][0].obs.cpu().numpy(), ob.obs), t
    assert sims[0][0].launch_count - l0[0][0] == 100 and sims[1][0].launch_count - l0[1][0] == 50


def test_config3_full_size_65536_envs_properties():
    """BASELINE config 3 at its full size (65 536 envs x 30 agents, production layout), checked through properties that do
    not need 65 536 oracle envs: (1) 128 sampled envs follow the oracle bit for bit at every step (an env's result does not
    depend on its neighbours in the batch); (2) the same batch stepped by a second simulator gives identical tensors
    (determinism: checksum of every output); (3) observe() is idempotent; (4) done stays False, rewards are 0/1."""
    import bench
    E, n_a, steps = 65536, 30, 200                               # a whole episode (CFG:181), the north star's rollout length
    shapes = load_shapes()
    ngm = int(shapes["n_g"].max())
    r_avoid = orc.r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"])
    blocks, n_g, l_cell, p, dp = bench.synth_batch(E, n_a, shapes, 7, "random")
    sims = [make_sim(E, n_a, ngm, r_avoid, out_dtype=torch.float32) for _ in range(2)]
    for s_ in sims:
        s_.set_grid(blocks, n_g, l_cell); s_.set_state(p, dp); s_.observe()
        check_scan_mode(s_)
    pick = np.random.RandomState(0).choice(E, 128, replace=False)
    params = [orc.make_params(n_a, int(n_g[e]), float(l_cell[e]), r_avoid) for e in pick]
    ob = orc.OracleBatch(params, nthreads=16)
    for k, e in enumerate(pick):
        ob.set_grid(k, blocks[e, :2 * n_g[e]].reshape(2, n_g[e]))
    ob.p[:], ob.dp[:] = p[pick], dp[pick]
    ob.observe()
    tp = torch.from_numpy(pick).cuda()
    assert np.array_equal(sims[0].obs[tp].cpu().numpy(), ob.obs.astype(np.float32))
    o0 = sims[0].obs.clone(); sims[0].observe(); assert torch.equal(o0, sims[0].obs)          # idempotent
    act = torch.empty(E, 2, n_a, dtype=torch.float32, device="cuda")
    for t in range(steps):
        sims[0].fill_actions(act, seed=11, step=t)
        a = act[tp].cpu().numpy()
        if t % 2:                                                   # drive the sampled envs (and only them) into their shapes
            a = goal_seeking_action(ob.obs, ob.dp, np.random.RandomState(t), noise=0.1)
            act[tp] = torch.from_numpy(a).cuda()
        for s_ in sims:
            s_.step(act)
        ob.step(a)
        assert np.array_equal(sims[0].p[tp].cpu().numpy(), ob.p) and np.array_equal(sims[0].dp[tp].cpu().numpy(), ob.dp), t
        assert np.array_equal(sims[0].obs[tp].cpu().numpy(), ob.obs.astype(np.float32)), t
        assert np.array_equal(sims[0].reward[tp].cpu().numpy(), ob.reward.astype(np.float32)), t
        assert np.array_equal(sims[0].a_prior[tp].cpu().numpy(), ob.a_prior.astype(np.float32)), t
        assert np.array_equal(sims[0].neighbor_index[tp].cpu().numpy(), ob.neighbor_index), t
    for name in ("p", "dp", "obs", "reward", "a_prior", "neighbor_index", "in_flags", "nearest_cell"):
        assert torch.equal(getattr(sims[0], name), getattr(sims[1], name)), name
    assert not sims[0].done.any() and bool(((sims[0].reward == 0) | (sims[0].reward == 1)).all())
    assert torch.isfinite(sims[0].obs).all() and torch.isfinite(sims[0].p).all()


@pytest.mark.parametrize("n_a,E,emit,dt", [(30, 61, True, torch.float64), (30, 61, False, torch.float32), (7, 33, True, torch.float64),
                                           (100, 9, True, torch.float64), (1, 5, False, torch.float32)])
def test_no_kernel_writes_outside_its_buffers(n_a, E, emit, dt):
    """Every device buffer of the simulator is allocated between two 64 KB guard zones (compute-sanitizer is closed on this
    GPU pool); after set_grid / reset / observe / steps through every emission path the guards must be untouched."""
    shapes, r_avoid, params, grids, P, DP = build_batch(E, n_a, seed=900 + n_a)
    ngm = int(shapes["n_g"].max())
    sim = make_sim(E, n_a, ngm, r_avoid, out_dtype=dt, emit_indices=emit, guard_bytes=65536)
    ob = orc.OracleBatch(params, nthreads=8)
    load_batch(sim, ob, params, grids, P, DP)
    sim.observe(); ob.observe()
    rng = np.random.RandomState(4)
    for t in range(40):
        a = goal_seeking_action(ob.obs, ob.dp, rng) if t % 3 else rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)
        sim.step(torch.from_numpy(a).cuda()); ob.step(a)
    assert np.array_equal(sim.p.cpu().numpy(), ob.p)
    sim.set_shapes(shapes["grid_origin"], shapes["l_cell"])
    sim.reset(seed=5); sim.metrics()
    act = torch.empty(E, 2, n_a, dtype=torch.float32, device="cuda")
    for t in range(5):
        sim.step(sim.fill_actions(act, seed=1, step=t))
    torch.cuda.synchronize()
    assert sim.check_guards()


def test_device_cosine_kernel_matches_libm():
    """psi = 0.5 * (1 + cos(pi * z / r)) (CPP:1012-1020) through the kernels' own [0, pi] cosine: within 2 ulp of 1.0
    (2.3e-16 absolute) of the host libm value over the whole range, at the quadrant boundaries and at the ends."""
    from marl_llm_b200 import _lib
    lib = _lib.load()
    r = 0.4
    rng = np.random.RandomState(0)
    z = np.concatenate([rng.uniform(0, r, 200000), np.linspace(0, r, 4097), r * np.array([0.25, 0.5, 0.75]) * (1 + np.array([-1e-16, 0, 1e-16])),
                        np.nextafter(r * np.array([0.25, 0.5, 0.75, 1.0]), 0), [0.0, r, np.nextafter(r, 1), 1.0, 1e-300]])
    out = np.empty_like(z)
    _lib.check(lib.swarm_debug_rho(C.c_void_p(z.ctypes.data), len(z), r, C.c_void_p(out.ctypes.data)), "swarm_debug_rho")
    want = np.where(z < r, 0.5 * (1.0 + np.cos(np.pi * (z / r))), 0.0)
    assert np.all(out[z >= r] == 0.0)
    assert np.max(np.abs(out - want)) <= 2.3e-16, np.max(np.abs(out - want))
    assert np.mean(out == want) > 0.9                       # bit-identical for the vast majority


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_shared_reciprocal_division_is_correctly_rounded(kind):
    """The prior (CPP:1143-1189) divides two or three numerators by one distance; the second-half kernel shares the reciprocal
    refinement between them (div_shared).  Every quotient must be the correctly rounded a / b of the reference's plain divisions:
    2e8 pseudo-random operand pairs per kind (simulator magnitudes, wide exponents, the divisors 1..6 of the velocity mean)."""
    from marl_llm_b200 import _lib
    lib = _lib.load()
    bad = C.c_uint64(123)
    _lib.check(lib.swarm_selftest_division(0, 200_000_000, 17 + kind, kind, C.byref(bad)), "swarm_selftest_division")
    assert bad.value == 0
