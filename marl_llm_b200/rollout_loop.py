"""The rollout phase of the reference's training loop (marl_llm/train/train_assembly.py:91-111) with every stage on the
device: policy (DevicePolicy) -> env.step (BatchedAssemblySim) -> buffer.push (ReplayBufferAgent).  Nothing crosses PCIe
inside the loop; per-step statistics stay on the device until the caller asks for them.

    for et in range(episode_length):                       TRAIN:91
        actions = maddpg.step(obs, explore=True)           TRAIN:97-99   -> policy.step(obs_prev, explore)
        next_obs, rew, done, _, prior = env.step(actions)  TRAIN:102     -> sim.step(actions)
        buffer.push(obs, actions, rew, next_obs, done, index, prior)     TRAIN:105-106
        obs = next_obs                                      TRAIN:107

The previous observation (policy input, replay `obs`) must outlive the step that produces the next one, so the simulator's
observation output alternates between two device buffers (`sim.set_obs_buffer`): no copy."""
import torch


def rollout(sim, policy, buffer, steps, explore=True, index=None):
    """Runs `steps` environment steps from the simulator's current observation.  Returns the per-step mean reward
    ([steps] float64 CUDA tensor; TRAIN:110 accumulates np.mean(rewards) per step).  On return `sim.obs` is the last
    observation (it may be either of the two buffers)."""
    index = index if index is not None else slice(0, sim.n_a)
    obs_prev, spare = sim.obs, torch.empty_like(sim.obs)
    act = torch.empty(sim.E, policy.act_dim, sim.n_a, dtype=torch.float32, device=sim.device)
    mean_rew = torch.zeros(steps, dtype=torch.float64, device=sim.device)
    for t in range(steps):
        _, log_pi = policy.step(obs_prev, explore=explore, out=act)
        sim.set_obs_buffer(spare)
        next_obs, rew, done, _, prior = sim.step(act)
        buffer.push(obs_prev, act, rew, next_obs, done, index, prior, log_pi)
        mean_rew[t] = rew.double().mean()
        obs_prev, spare = next_obs, obs_prev
    return mean_rew


def rollout_ring(sim, policy, ring, steps, explore=True):
    """The same loop on a time-indexed ring (marl_llm_b200/episode_ring.py): the policy kernel writes each step's
    observations as replay rows while it reads them, env.step follows, and only the small per-agent arrays are pushed —
    no observation transposes in the loop.  Returns the per-step mean reward ([steps] float64 CUDA tensor)."""
    assert steps <= ring.T and sim.out_dtype == torch.float32
    ring.begin()
    obs_prev, spare = sim.obs, torch.empty_like(sim.obs)
    act = torch.empty(sim.E, policy.act_dim, sim.n_a, dtype=torch.float32, device=sim.device)
    mean_rew = torch.zeros(steps, dtype=torch.float64, device=sim.device)
    for t in range(steps):
        _, log_pi = policy.step(obs_prev, explore=explore, out=act, rows_out=ring.slot(t))
        sim.set_obs_buffer(spare)
        next_obs, rew, done, _, prior = sim.step(act)
        ring.record(t, act, rew, done, prior, log_pi)
        mean_rew[t] = rew.double().mean()
        obs_prev, spare = next_obs, obs_prev
    ring.close(obs_prev)
    return mean_rew


def rollout_ring_direct(sim, policy, ring, steps, explore=True):
    """The ring loop with an agent-major simulator (`obs_layout='agent_major'`): the simulator writes the observation of step
    t + 1 straight into ring slot t + 1 (`set_obs_buffer`), the policy reads slot t as contiguous 768-byte rows — no
    observation is copied or transposed anywhere in the loop.  The simulator's current observation must already be in slot 0
    (`ring.begin_direct(sim)` does that).  Returns the per-step mean reward ([steps] float64 CUDA tensor)."""
    assert steps <= ring.T and sim.out_dtype == torch.float32 and sim.obs_layout == "agent_major"
    act = torch.empty(sim.E, policy.act_dim, sim.n_a, dtype=torch.float32, device=sim.device)
    mean_rew = torch.zeros(steps, dtype=torch.float64, device=sim.device)
    for t in range(steps):
        _, log_pi = policy.step(ring.slot_env(t), explore=explore, out=act, agent_major=True)
        sim.set_obs_buffer(ring.slot_env(t + 1))
        _, rew, done, _, prior = sim.step(act)
        ring.record(t, act, rew, done, prior, log_pi)
        mean_rew[t] = rew.double().mean()
    ring.closed = True                     # slot `steps` already holds the last observation
    return mean_rew
