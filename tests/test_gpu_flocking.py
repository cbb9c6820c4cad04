"""FlockingSwarm variant (VARIANTS.md §3).  The reference ships no source for it, so there is no oracle: these are
SELF-CONSISTENCY tests — (1) everything shared with the assembly env is bit-identical to the assembly simulator (itself pinned to
the reference), (2) the new reward matches its written specification, (3) equivariance / determinism invariants."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.helpers import load_shapes, reset_like_reference

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_a,E,periodic,self_state", [(30, 48, False, True), (30, 16, True, True), (64, 8, False, False), (7, 32, False, True)])
def test_shared_core_is_bit_identical_to_the_assembly_simulator(n_a, E, periodic, self_state):
    """Same state + actions into the assembly simulator and the flocking simulator: positions, velocities, neighbour lists and
    the observation head rows (CPP:102-126) must agree bit for bit at every step — the two run the same first-half kernel."""
    from marl_llm_b200.batched import BatchedAssemblySim
    from marl_llm_b200.flocking import BatchedFlockingSim
    shapes = load_shapes()
    rng = np.random.RandomState(n_a)
    r_avoid = 0.26
    grids, P, DP, lc = [], [], [], []
    for e in range(E):
        k, grid, p, dp = reset_like_reference(rng, n_a, shapes)
        grids.append(grid); P.append(p); DP.append(dp); lc.append(float(shapes["l_cell"][k]))
    ngm = int(shapes["n_g"].max())
    asm = BatchedAssemblySim(E, n_a, ngm, r_avoid, out_dtype=torch.float64, is_periodic=periodic, is_con_self_state=self_state)
    blocks, n_g = asm.pack_grids(grids, ngm)
    asm.set_grid(blocks, n_g, lc); asm.set_state(np.stack(P), np.stack(DP)); asm.observe()
    flk = BatchedFlockingSim(E, n_a, r_avoid, out_dtype=torch.float64, is_periodic=periodic, is_con_self_state=self_state)
    flk.set_state(np.stack(P), np.stack(DP)); flk.observe()
    head = flk.obs_dim
    assert head == 4 * (6 + int(self_state))
    assert torch.equal(flk.obs, asm.obs[:, :head]) and torch.equal(flk.neighbor_index, asm.neighbor_index)
    for t in range(40):
        a = torch.from_numpy(rng.uniform(-1, 1, (E, 2, n_a)).astype(np.float32)).cuda()
        asm.step(a); obs, rew, done, _ = flk.step(a)
        assert torch.equal(flk.p, asm.p) and torch.equal(flk.dp, asm.dp), t
        assert torch.equal(obs, asm.obs[:, :head]) and torch.equal(flk.neighbor_index, asm.neighbor_index), t
        assert not done.any() and bool((rew <= 0).all())


def test_reward_matches_its_specification_and_is_permutation_equivariant():
    from marl_llm_b200.flocking import BatchedFlockingSim, reward_reference
    E, n_a, r_avoid = 24, 30, 0.26
    sim = BatchedFlockingSim(E, n_a, r_avoid, out_dtype=torch.float64)
    sim.reset(seed=3)
    sim.p.mul_(0.35)                                      # a denser swarm: collisions and full neighbour lists occur
    sim.observe()
    g = torch.Generator(device="cuda").manual_seed(1)
    worst = 0.0
    for t in range(30):
        a = torch.rand(E, 2, n_a, device="cuda", generator=g) * 2 - 1
        _, rew, _, _ = sim.step(a)
        p, dp, nbr = sim.p.cpu().numpy(), sim.dp.cpu().numpy(), sim.neighbor_index.cpu().numpy()
        for e in range(0, E, 5):
            want = reward_reference(p[e], dp[e], nbr[e], r_avoid)
            worst = max(worst, float(np.abs(rew[e, 0].cpu().numpy() - want).max()))
    assert worst < 1e-12, worst
    assert float(rew.min()) < -1.0                        # some collisions happened
    # relabelling the agents permutes rewards and observations the same way (neighbour ids are relabelled too)
    perm = torch.randperm(n_a, device="cuda", generator=g)
    a = torch.rand(E, 2, n_a, device="cuda", generator=g) * 2 - 1
    twin = BatchedFlockingSim(E, n_a, r_avoid, out_dtype=torch.float64)
    twin.set_state(sim.p[:, :, perm], sim.dp[:, :, perm]); twin.observe()
    _, r1, _, _ = sim.step(a)
    _, r2, _, _ = twin.step(a[:, :, perm].contiguous())
    assert torch.equal(twin.p, sim.p[:, :, perm])
    assert torch.allclose(r2, r1[:, :, perm], rtol=0, atol=1e-12)      # ties in the neighbour order could reorder a sum


def test_determinism_and_fp32_outputs_are_rounded_fp64():
    from marl_llm_b200.flocking import BatchedFlockingSim
    E, n_a = 4096, 30
    a64, b64, c32 = (BatchedFlockingSim(E, n_a, out_dtype=dt) for dt in (torch.float64, torch.float64, torch.float32))
    for s_ in (a64, b64, c32):
        s_.reset(seed=9)
    g = torch.Generator(device="cuda").manual_seed(2)
    for t in range(20):
        a = torch.rand(E, 2, n_a, device="cuda", generator=g) * 2 - 1
        for s_ in (a64, b64, c32):
            s_.step(a)
    assert torch.equal(a64.p, b64.p) and torch.equal(a64.obs, b64.obs) and torch.equal(a64.reward, b64.reward)
    assert torch.equal(c32.p, a64.p) and torch.equal(c32.obs, a64.obs.float()) and torch.equal(c32.reward, a64.reward.float())
