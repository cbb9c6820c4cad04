"""Device-resident replay storage (C ABI group 3 + marl_llm_b200/rollout.py) against the oracle restatement of the
reference's ReplayBufferAgent: ring contents, cursor arithmetic (step-back + wrap) and seeded sample() batches."""
import numpy as np
import pytest
import torch

from marl_llm_b200.rollout import ReplayBufferAgent
from oracle.replay_oracle import ReplayOracle

pytestmark = pytest.mark.gpu

ARRAYS = ("obs_buffs", "ac_buffs", "ac_prior_buffs", "rew_buffs", "next_obs_buffs", "done_buffs", "log_pi_buffs")


def assert_same(dev, orc):
    for name in ARRAYS:
        assert np.array_equal(getattr(dev, name).cpu().numpy(), getattr(orc, name).astype(np.float32)), name
    assert (dev.curr_i, dev.filled_i, len(dev)) == (orc.curr_i, orc.filled_i, len(orc))


@pytest.mark.parametrize("n_a,D", [(30, 192), (7, 188), (100, 192), (1, 16)])
def test_reference_shaped_host_pushes_rollover_and_seeded_samples(n_a, D):
    A = 2
    max_steps = 300200 // n_a + 3
    idx = slice(0, n_a)
    dev, orc = ReplayBufferAgent(max_steps, n_a, idx, D, A), ReplayOracle(max_steps, n_a, idx, D, A)
    rng = np.random.RandomState(n_a)
    for b in (dev, orc):
        b.curr_i = b.filled_i = b.total_length - 2 * n_a - min(3, n_a - 1) - 1
    for t in range(5):
        obs, nxt = rng.randn(D, n_a), rng.randn(D, n_a)                          # fp64, like the reference env
        act = rng.uniform(-1, 1, (A, n_a)).astype(np.float32)                     # TRAIN:99 column_stack of fp32
        prior, rew, done = rng.uniform(-1, 1, (A, n_a)), rng.rand(1, n_a), rng.rand(1, n_a) > 0.5
        logpi = rng.randn(1, n_a).astype(np.float32) if t % 2 else None
        dev.push(obs, act, rew, nxt, done, idx, prior, logpi)
        orc.push(obs, act, rew, nxt, done, idx, prior, logpi)
        assert_same(dev, orc)
    for seed in (1, 2):
        np.random.seed(seed); got = dev.sample(128, to_gpu=True, is_prior=True, is_log_pi=True)
        np.random.seed(seed); want, _ = orc.sample(128, is_prior=True, is_log_pi=True)
        for a, b in zip(got, want):
            assert a.is_cuda and np.array_equal(a.cpu().numpy(), b)
    np.random.seed(3); host = dev.sample(16)
    assert host[5] is None and host[6] is None and not host[0].is_cuda and host[0].dtype == torch.float32


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.float64])
def test_batched_device_push_from_simulator_layout(out_dtype):
    E, n_a, D, A = 37, 30, 192, 2
    idx = slice(0, n_a)
    dev, orc = ReplayBufferAgent(40, E * n_a, idx, D, A), ReplayOracle(40, E * n_a, idx, D, A)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(45):                                                           # > max_steps: wraps to row 0
        obs = torch.randn(E, D, n_a, device="cuda", generator=g).to(out_dtype)
        nxt = torch.randn(E, D, n_a, device="cuda", generator=g).to(out_dtype)
        act = torch.rand(E, A, n_a, device="cuda", generator=g) * 2 - 1
        prior = (torch.rand(E, A, n_a, device="cuda", generator=g) * 2 - 1).to(out_dtype)
        rew = torch.rand(E, 1, n_a, device="cuda", generator=g).to(out_dtype)
        done = torch.rand(E, 1, n_a, device="cuda", generator=g) > 0.7
        dev.push(obs, act, rew, nxt, done, idx, prior)
        orc.push(*(x.cpu().numpy() for x in (obs, act, rew, nxt, done)), idx, prior.cpu().numpy())
    assert_same(dev, orc)
    rows = np.random.RandomState(0).randint(0, dev.total_length, 300)
    got = dev.gather(rows, is_prior=True)
    assert np.array_equal(got[0].cpu().numpy(), orc.obs_buffs[rows].astype(np.float32))
    assert np.array_equal(got[3].cpu().numpy(), orc.next_obs_buffs[rows].astype(np.float32))
    assert np.array_equal(got[5].cpu().numpy(), orc.ac_prior_buffs[rows].astype(np.float32))


def test_agent_slice_and_errors():
    n_a, D, A = 12, 20, 2
    dev, orc = ReplayBufferAgent(50, 5, slice(3, 8), D, A), ReplayOracle(50, 5, slice(3, 8), D, A)
    rng = np.random.RandomState(2)
    for t in range(4):
        obs, nxt, act = rng.randn(D, n_a), rng.randn(D, n_a), rng.randn(A, n_a).astype(np.float32)
        rew, done = rng.rand(1, n_a), rng.rand(1, n_a) > 0.5
        dev.push(obs, act, rew, nxt, done, slice(3, 8)); orc.push(obs, act, rew, nxt, done, slice(3, 8))
    assert_same(dev, orc)
    with pytest.raises(IndexError):
        dev.gather([dev.total_length])


def test_device_rollout_loop_fills_the_buffer_consistently():
    """policy -> step -> push entirely on the device (marl_llm_b200/rollout_loop.py, TRAIN:91-111): the rows pushed at
    step t are the transposes of what the simulator / policy held at that step, and next_obs of step t is obs of t+1."""
    import torch.nn as nn
    from marl_llm_b200.batched import BatchedAssemblySim, r_avoid_for
    from marl_llm_b200.policy import DevicePolicy
    from marl_llm_b200.rollout_loop import rollout
    from tests.helpers import load_shapes
    shapes = load_shapes()
    E, n_a, T = 16, 30, 6
    sim = BatchedAssemblySim(E, n_a, int(shapes["n_g"].max()), r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"]))
    sim.set_shapes(shapes["grid_origin"], shapes["l_cell"])
    sim.reset(seed=3)
    torch.manual_seed(0)
    sd = {}
    for name, (o, i) in (("fc1", (180, 192)), ("fc2", (180, 180)), ("fc3", (180, 180)), ("fc4", (2, 180))):
        l = nn.Linear(i, o); sd[name + ".weight"] = l.weight; sd[name + ".bias"] = l.bias
    pol = DevicePolicy(192, 2, 180, noise_scale=0.3, seed=5).load_state_dict(sd)
    buf = ReplayBufferAgent(T, E * n_a, slice(0, n_a), 192, 2)
    obs0 = sim.obs.clone()
    mean_rew = rollout(sim, pol, buf, T)
    n = E * n_a
    assert len(buf) == T * n and buf.curr_i == 0 and mean_rew.shape == (T,)
    rows = lambda t: t.permute(0, 2, 1).reshape(n, -1)     # noqa: E731
    assert torch.equal(buf.obs_buffs[:n], rows(obs0))
    for t in range(T - 1):
        assert torch.equal(buf.next_obs_buffs[t * n:(t + 1) * n], buf.obs_buffs[(t + 1) * n:(t + 2) * n]), t
    assert torch.equal(buf.next_obs_buffs[(T - 1) * n:], rows(sim.obs))
    assert torch.equal(buf.rew_buffs[(T - 1) * n:], rows(sim.reward)) and torch.equal(buf.ac_prior_buffs[(T - 1) * n:], rows(sim.a_prior))
    assert float(buf.ac_buffs.abs().max()) <= 1.0 and float(buf.log_pi_buffs.abs().sum()) > 0


@pytest.mark.parametrize("prec", ["fp32", "f16_tc", "f16x3_tc"])
def test_time_indexed_ring_equals_the_conventional_buffer(prec):
    """EpisodeRing (observations stored once, rows written by the policy kernel) against ReplayBufferAgent (obs and next_obs
    transposed by the push) for the same seeded rollout: identical transitions, for every policy kernel."""
    import torch.nn as nn
    from marl_llm_b200.batched import BatchedAssemblySim, r_avoid_for
    from marl_llm_b200.episode_ring import EpisodeRing
    from marl_llm_b200.policy import DevicePolicy
    from marl_llm_b200.rollout_loop import rollout, rollout_ring
    from tests.helpers import load_shapes
    shapes = load_shapes()
    E, n_a, T = 21, 30, 5
    n = E * n_a
    torch.manual_seed(0)
    sd = {}
    for name, (o, i) in (("fc1", (180, 192)), ("fc2", (180, 180)), ("fc3", (180, 180)), ("fc4", (2, 180))):
        l = nn.Linear(i, o); sd[name + ".weight"] = l.weight; sd[name + ".bias"] = l.bias
    outs = []
    for kind in ("buffer", "ring", "direct"):
        sim = BatchedAssemblySim(E, n_a, int(shapes["n_g"].max()), r_avoid_for(n_a, shapes["n_g"], shapes["l_cell"]),
                                 obs_layout="agent_major" if kind == "direct" else "reference")
        sim.set_shapes(shapes["grid_origin"], shapes["l_cell"])
        sim.reset(seed=8)
        pol = DevicePolicy(192, 2, 180, noise_scale=0.3, seed=5, precision=prec).load_state_dict(sd)
        if kind == "buffer":
            store = ReplayBufferAgent(T, n, slice(0, n_a), 192, 2)
            rollout(sim, pol, store, T)
            outs.append(store.gather(np.arange(T * n), is_prior=True, is_log_pi=True))
        elif kind == "direct":
            # agent-major simulator: it writes every observation straight into the ring slot the policy reads next
            from marl_llm_b200.rollout_loop import rollout_ring_direct
            ring = EpisodeRing(T, E, n_a, 192, 2)
            ring.begin_direct(sim)
            rollout_ring_direct(sim, pol, ring, T)
            assert len(ring) == T * n and ring.closed
            outs.append(ring.gather(np.arange(T * n), is_prior=True, is_log_pi=True))
        else:
            ring = EpisodeRing(T, E, n_a, 192, 2)
            assert len(ring) == 0
            rollout_ring(sim, pol, ring, T)
            assert len(ring) == T * n and ring.closed
            outs.append(ring.gather(np.arange(T * n), is_prior=True, is_log_pi=True))
            smp = ring.sample(64, is_prior=True)
            assert smp[0].shape == (64, 192) and smp[5].shape == (64, 2) and smp[6] is None
            with pytest.raises(IndexError):
                ring.gather([T * n])
    for a, b, c in zip(*outs):
        assert torch.equal(a, b) and torch.equal(a, c)
