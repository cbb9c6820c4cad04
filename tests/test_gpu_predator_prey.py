"""PredatorPreySwarm variant (VARIANTS.md §4).  The reference ships no source for it, so there is no oracle: these are
SELF-CONSISTENCY tests — (1) the kernel equals the written specification (`predator_prey.step_reference`, plain NumPy loops) bit
for bit, (2) its dynamics equal the shared pair core of the assembly / flocking simulators (pinned to the reference by the assembly
parity suite), (3) invariants: equivariance, speed conservation of the elastic wall, scripted strategies, batch independence."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dense_state(sim, seed, scale=0.3):
    sim.reset(seed=seed)
    sim.p.mul_(scale)                                     # a denser arena: contacts, captures and full neighbour lists occur
    return sim.observe()


@pytest.mark.parametrize("n_p,n_e,periodic,billiards,self_state,out", [
    (3, 10, False, False, True, torch.float64), (5, 25, True, False, True, torch.float64), (4, 9, False, True, False, torch.float64),
    (1, 1, False, False, True, torch.float64), (40, 60, False, False, True, torch.float32), (6, 0, False, False, True, torch.float64)])
def test_kernel_equals_the_written_specification(n_p, n_e, periodic, billiards, self_state, out):
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim, step_reference
    E, n = 6, n_p + n_e
    half = (0.8, 0.8)
    sim = BatchedPredatorPreySim(E, n_p, n_e, out_dtype=out, is_periodic=periodic, billiards=billiards, is_con_self_state=self_state,
                                 half_width=half[0], half_height=half[1])
    dense_state(sim, seed=n, scale=1.0)
    rng = np.random.RandomState(n)
    p, dp = sim.p.cpu().numpy().copy(), sim.dp.cpu().numpy().copy()
    saw_contact = saw_capture = False
    for t in range(12):
        a = rng.uniform(-1, 1, (E, 2, n))
        obs, rew, done, _ = sim.step(torch.from_numpy(a).cuda())
        for e in range(E):
            p[e], dp[e], want_obs, want_rew, want_nbr = step_reference(p[e], dp[e], a[e], n_p, periodic=periodic, billiards=billiards,
                                                                      self_state=self_state, half=half)
            assert np.array_equal(sim.p[e].cpu().numpy(), p[e]) and np.array_equal(sim.dp[e].cpu().numpy(), dp[e]), (t, e)
            assert np.array_equal(sim.neighbor_index[e].cpu().numpy(), want_nbr), (t, e)
            if out == torch.float64:
                assert np.array_equal(obs[e].cpu().numpy(), want_obs) and np.array_equal(rew[e, 0].cpu().numpy(), want_rew), (t, e)
            else:
                assert np.array_equal(obs[e].cpu().numpy(), want_obs.astype(np.float32))
                assert np.array_equal(rew[e, 0].cpu().numpy(), want_rew.astype(np.float32))
            saw_capture |= bool(np.abs(want_rew).max() >= 0.9)
        assert not done.any()
    if n >= 30:
        assert saw_capture                                # the dense arena exercises the capture term


def test_observe_without_dynamics_and_input_dtype():
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim, step_reference
    sim = BatchedPredatorPreySim(4, 3, 7, out_dtype=torch.float64)
    obs = dense_state(sim, seed=2).clone()
    for e in range(4):
        _, _, want_obs, _, want_nbr = step_reference(sim.p[e].cpu().numpy(), sim.dp[e].cpu().numpy(), None, 3, dyn=False)
        assert np.array_equal(obs[e].cpu().numpy(), want_obs) and np.array_equal(sim.neighbor_index[e].cpu().numpy(), want_nbr)
    a32 = (torch.rand(4, 2, 10, device="cuda") * 2 - 1).float()
    p0, dp0 = sim.p.clone(), sim.dp.clone()
    sim.step(a32); p32 = sim.p.clone()
    sim.set_state(p0, dp0); sim.step(a32.double())
    assert torch.equal(sim.p, p32)                        # fp32 actions are promoted exactly
    with pytest.raises(TypeError):
        sim.step(np.zeros((4, 2, 10)))


@pytest.mark.parametrize("n_a,periodic", [(30, False), (48, True)])
def test_dynamics_are_the_shared_pair_core(n_a, periodic):
    """Equal velocity limits for both types: positions and velocities must follow the flocking simulator's — i.e. the assembly
    step's first-half kernel — bit for bit."""
    from marl_llm_b200.flocking import BatchedFlockingSim
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim
    E = 16
    pp = BatchedPredatorPreySim(E, 10, n_a - 10, out_dtype=torch.float64, is_periodic=periodic, vel_max_p=0.8, vel_max_e=0.8)
    flk = BatchedFlockingSim(E, n_a, 0.26, out_dtype=torch.float64, is_periodic=periodic)
    dense_state(pp, seed=5)
    flk.set_state(pp.p, pp.dp); flk.observe()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(40):
        a = torch.rand(E, 2, n_a, device="cuda", generator=g) * 2 - 1
        pp.step(a); flk.step(a)
        assert torch.equal(pp.p, flk.p) and torch.equal(pp.dp, flk.dp), t
        # the self rows of the observation are the shared head's, too
        assert torch.equal(pp.obs[:, :4], flk.obs[:, :4])


def test_elastic_wall_conserves_speed_and_keeps_agents_near_the_arena():
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim
    E, n_p, n_e = 32, 4, 4
    sim = BatchedPredatorPreySim(E, n_p, n_e, out_dtype=torch.float64, billiards=True, pursuer_strategy="static", escaper_strategy="static",
                                 half_width=1.0, half_height=1.0, vel_max_p=10.0, vel_max_e=10.0, k_ball=0.0)
    sim.reset(seed=9)
    speed0 = sim.dp.abs().clone()
    for t in range(300):
        sim.step()
    assert torch.equal(sim.dp.abs(), speed0)              # no forces act (static strategies, k_ball = 0): reflections only flip signs
    assert float(sim.p.abs().max()) < 1.0 + 0.5 * 0.1 + 1e-9      # at most one step beyond the wall


def test_nearest_strategy_and_random_strategy():
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim
    E, n_p, n_e = 8, 3, 5
    kw = dict(out_dtype=torch.float64, k_ball=0.0, k_wall=0.0, c_wall=0.0, vel_max_p=100.0, vel_max_e=100.0)
    sim = BatchedPredatorPreySim(E, n_p, n_e, pursuer_strategy="nearest", escaper_strategy="nearest", **kw)
    sim.reset(seed=4); sim.dp.zero_()
    p0 = sim.p.cpu().numpy().copy()
    sim.step()
    u = sim.dp.cpu().numpy() / 0.1                         # v = 0 + u * dt
    for e in range(E):
        for i in range(n_p + n_e):
            others = range(n_p, n_p + n_e) if i < n_p else range(n_p)
            rel = np.stack([p0[e][:, j] - p0[e][:, i] for j in others])
            j = int(np.argmin((rel ** 2).sum(1)))
            want = rel[j] / np.linalg.norm(rel[j]) * (1.0 if i < n_p else -1.0)
            assert np.allclose(u[e][:, i], want, atol=1e-12), (e, i)
    # random: U(-1, 1), deterministic in (seed, step, env, agent), different between steps and envs
    a = BatchedPredatorPreySim(E, n_p, n_e, pursuer_strategy="random", escaper_strategy="random", seed=7, **kw)
    b = BatchedPredatorPreySim(E, n_p, n_e, pursuer_strategy="random", escaper_strategy="random", seed=7, **kw)
    a.reset(seed=1); b.reset(seed=1); a.dp.zero_(); b.dp.zero_()
    a.step(); b.step()
    assert torch.equal(a.dp, b.dp)
    ua = a.dp / 0.1
    assert float(ua.abs().max()) <= 1.0 + 1e-12 and float(ua.std()) > 0.4 and not torch.equal(ua[0], ua[1])
    v1 = a.dp.clone(); a.dp.zero_(); a.step()
    assert not torch.equal(a.dp, v1)


def test_rewards_are_zero_sum_on_captures_and_equivariant_and_batch_independent():
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim
    E, n_p, n_e = 64, 6, 18
    n = n_p + n_e
    sim = BatchedPredatorPreySim(E, n_p, n_e, out_dtype=torch.float64, is_periodic=True)      # periodic: no wall term
    dense_state(sim, seed=11, scale=0.25)
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand(E, 2, n, device="cuda", generator=g) * 2 - 1
    p0, dp0 = sim.p.clone(), sim.dp.clone()
    obs, rew, _, _ = sim.step(a)
    obs, rew, p1 = obs.clone(), rew.clone(), sim.p.clone()
    # capture counts: every cross-type contact is counted once by a pursuer (+1) and once by an escaper (-1); what is left of the
    # rewards are the +-0.1 d_nearest terms, bounded by 0.1 * arena diagonal
    rel = p1[:, :, :n_p, None] - p1[:, :, None, n_p:]
    rel = torch.where(rel < -2.4, rel + 4.8, torch.where(rel > 2.4, rel - 4.8, rel))
    caps = ((rel ** 2).sum(1).sqrt() < 0.07)
    assert int(caps.sum()) > 0
    assert torch.allclose(rew[:, 0, :n_p] + 0.1 * (rel ** 2).sum(1).sqrt().min(2).values, caps.sum(2).double(), atol=1e-12)
    assert torch.allclose(rew[:, 0, n_p:] - 0.1 * (rel ** 2).sum(1).sqrt().min(1).values, -caps.sum(1).double(), atol=1e-12)
    # permuting the agents within each type permutes rewards and self rows
    perm = torch.cat([torch.randperm(n_p, generator=torch.Generator().manual_seed(0)), n_p + torch.randperm(n_e, generator=torch.Generator().manual_seed(1))]).cuda()
    sim.set_state(p0[:, :, perm], dp0[:, :, perm])
    obs2, rew2, _, _ = sim.step(a[:, :, perm])
    # (to rounding: an agent in contact with several others sums their spring forces in index order)
    assert torch.allclose(sim.p, p1[:, :, perm], atol=1e-12) and torch.allclose(obs2[:, :4], obs[:, :4][:, :, perm], atol=1e-12)
    assert torch.allclose(rew2, rew[:, :, perm], atol=1e-12)
    # a sub-batch gives the same rows
    sub = BatchedPredatorPreySim(8, n_p, n_e, out_dtype=torch.float64, is_periodic=True)
    sub.set_state(p0[40:48], dp0[40:48]); sub.observe()
    o3, r3, _, _ = sub.step(a[40:48])
    assert torch.equal(o3, obs[40:48]) and torch.equal(r3, rew[40:48])


def test_bad_arguments_are_errors():
    from marl_llm_b200 import _lib
    from marl_llm_b200.predator_prey import BatchedPredatorPreySim
    with pytest.raises(ValueError):
        BatchedPredatorPreySim(2, 100, 100)
    sim = BatchedPredatorPreySim(2, 2, 2)
    sim.cfg.strategy_p = 9
    with pytest.raises(_lib.SwarmError):
        sim.observe()
