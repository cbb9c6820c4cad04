"""TEST INFRASTRUCTURE — CPU restatement of the reference's replay buffer (marl_llm/algorithm/utils/buffer_agent.py = BUF).
Only tests/ may import this; the product (marl_llm_b200/rollout.py) never does.

Pinned: tests/test_replay_oracle.py drives this class and the real `ReplayBufferAgent` (imported from /root/reference)
with the same seeded pushes / samples and demands identical arrays.  Extension over the reference: `push` accepts a
leading env axis ([E, dim, n_a]) and appends the envs one after the other (the reference has a single env)."""
import numpy as np


class ReplayOracle:
    def __init__(self, max_steps, num_agents, start_stop_index, state_dim, action_dim):
        self.max_steps, self.num_agents = max_steps, num_agents                   # BUF:33-34
        self.total_length = max_steps * num_agents                                # BUF:46
        z = lambda d: np.zeros((self.total_length, d))                            # noqa: E731   BUF:49-55 (fp64)
        self.obs_buffs, self.ac_buffs, self.ac_prior_buffs, self.log_pi_buffs = z(state_dim), z(action_dim), z(action_dim), z(1)
        self.rew_buffs, self.next_obs_buffs, self.done_buffs = z(1), z(state_dim), z(1)
        self.filled_i = 0; self.curr_i = 0                                        # BUF:58-59
        self.agent_index = start_stop_index

    def __len__(self):
        return self.filled_i

    def _push_one(self, obs, act, rew, nxt, done, index, prior, logpi):
        span = range(index.start, index.stop); n = len(span)                      # BUF:86-90
        if self.curr_i + n > self.total_length:                                   # BUF:96-99: steps back, does not wrap
            self.curr_i -= n - (self.total_length - self.curr_i)
        s = slice(self.curr_i, self.curr_i + n)
        self.obs_buffs[s] = obs[:, index].T; self.ac_buffs[s] = act[:, index].T   # BUF:102-106
        self.rew_buffs[s] = rew[:, index].T; self.next_obs_buffs[s] = nxt[:, index].T; self.done_buffs[s] = done[:, index].T
        if prior is not None: self.ac_prior_buffs[s] = prior[:, index].T          # BUF:109-112
        if logpi is not None: self.log_pi_buffs[s] = logpi[:, index].T
        self.curr_i += n                                                          # BUF:116
        if self.filled_i < self.total_length: self.filled_i += n                  # BUF:119-120
        if self.curr_i == self.total_length: self.curr_i = 0                      # BUF:123-124

    def push(self, obs, act, rew, nxt, done, index, prior=None, logpi=None):
        obs = np.asarray(obs)
        if obs.ndim == 2:
            return self._push_one(obs, np.asarray(act), np.asarray(rew), np.asarray(nxt), np.asarray(done), index, prior, logpi)
        # batched extension: one logical push of E*n rows (cursor arithmetic applied once, like a single push of that length)
        E = obs.shape[0]; n = len(range(index.start, index.stop)) * E
        if self.curr_i + n > self.total_length:
            self.curr_i -= n - (self.total_length - self.curr_i)
        s = slice(self.curr_i, self.curr_i + n)
        t = lambda a: np.asarray(a)[:, :, index].transpose(0, 2, 1).reshape(n, -1)   # noqa: E731
        self.obs_buffs[s] = t(obs); self.ac_buffs[s] = t(act); self.rew_buffs[s] = t(rew)
        self.next_obs_buffs[s] = t(nxt); self.done_buffs[s] = t(done)
        if prior is not None: self.ac_prior_buffs[s] = t(prior)
        if logpi is not None: self.log_pi_buffs[s] = t(logpi)
        self.curr_i += n
        if self.filled_i < self.total_length: self.filled_i += n
        if self.curr_i == self.total_length: self.curr_i = 0

    def sample(self, N, is_prior=False, is_log_pi=False):
        """BUF:130-177 up to the tensor cast: fp32 arrays (torch.Tensor(x) rounds fp64 to fp32)."""
        begin_index_range = 3e5                                                   # BUF:147
        begin_index = np.random.randint(0, begin_index_range)
        inds = np.random.choice(np.arange(begin_index, self.total_length - begin_index_range + begin_index, dtype=np.int32),
                                size=N, replace=False)                            # BUF:151-157
        c = lambda a: a[inds, :].astype(np.float32)                               # noqa: E731
        return (c(self.obs_buffs), c(self.ac_buffs), c(self.rew_buffs), c(self.next_obs_buffs), c(self.done_buffs),
                c(self.ac_prior_buffs) if is_prior else None, c(self.log_pi_buffs) if is_log_pi else None), inds
