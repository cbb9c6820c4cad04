#!/usr/bin/env python
"""Record golden trajectories from the UNMODIFIED reference env (build container only).

TEST/FIXTURE INFRASTRUCTURE.  Drives the real `AssemblySwarmEnv` (imported from /root/reference via
oracle/live_reference.py, bound to oracle/_ref/libAssemblyEnv.so = the reference's own C++) exactly like
marl_llm/train/train_assembly.py:48-50,81,102 does, and stores everything a step returns or leaves in
the env.  The fixtures travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_goldens.py        # rewrites tests/golden/traj_*.npz

Per trajectory: shape/grid parameters, initial state, the float32 actions fed in, then for EVERY step
p, dp, reward, a_prior, neighbor_index, in_flags and SHA-1 digests of obs / sensed_index / occupied_index,
and the full obs / sensed_index / occupied_index arrays on every `full_every`-th step and the last.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import live_reference as lr          # noqa: E402
from tests.helpers import goal_seeking_action    # noqa: E402

CASES = [
    # name,            n_a, seed, mode,     steps, full_every
    ("a30_random_s226", 30, 226, "random", 200, 10),
    ("a30_goal_s3",     30, 3,   "goal",   200, 10),
    ("a10_goal_s15",    10, 15,  "goal",   100, 10),
    ("a64_goal_s75",    64, 75,  "goal",    60, 10),
    ("a30_periodic_s7", 30, 7,   "drift",  150, 10),   # is_boundary=False: periodic wrap (assembly.py:99-103, 651-652)
]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def record(n_a, seed, mode, steps, full_every):
    periodic = mode == "drift"
    env = lr.make_env(n_a, is_boundary=not periodic)
    np.random.seed(seed)
    obs0 = env.reset()
    e = env.env
    out = dict(n_a=n_a, seed=seed, steps=steps, is_periodic=int(periodic), n_g=e.n_g, l_cell=e.l_cell, r_avoid=e.r_avoid, d_sen=e.d_sen,
               grid_center=e.grid_center.copy(), boundary_pos=e.boundary_pos.copy(),
               p0=e.p.copy(), dp0=e.dp.copy(), obs0=obs0.copy(), nbr0=e.neighbor_index.copy(),
               in_flags0=e.in_flags.copy(), sensed0=e.sensed_index.copy(), occupied0=e.occupied_index.copy())
    rng = np.random.RandomState(seed + 1)
    acts, P, DP, R, PR, NB, INF, h_obs, h_sen, h_occ = [], [], [], [], [], [], [], [], [], []
    full_steps, F_obs, F_sen, F_occ = [], [], [], []
    for t in range(steps):
        if mode == "random":
            a = rng.uniform(-1, 1, (2, n_a)).astype(np.float32)
        elif mode == "drift":      # outward drift: agents cross the box edges and wrap around
            a = np.clip(0.8 * np.sign(e.p) + rng.normal(0, 0.5, (2, n_a)), -1, 1).astype(np.float32)
        else:
            a = goal_seeking_action(e.obs, e.dp, rng)
        obs, rew, done, info, prior = env.step(a)
        assert not done.any() and done.shape == (1, n_a) and done.dtype == bool
        acts.append(a); P.append(e.p.copy()); DP.append(e.dp.copy()); R.append(rew.copy()); PR.append(prior.copy())
        NB.append(e.neighbor_index.copy()); INF.append(e.in_flags.astype(np.int8))
        h_obs.append(sha(obs)); h_sen.append(sha(e.sensed_index)); h_occ.append(sha(e.occupied_index))
        if t % full_every == 0 or t == steps - 1:
            full_steps.append(t); F_obs.append(obs.copy())
            F_sen.append(e.sensed_index.astype(np.int16)); F_occ.append(e.occupied_index.astype(np.int16))
    out.update(act=np.stack(acts), p=np.stack(P), dp=np.stack(DP), reward=np.stack(R), a_prior=np.stack(PR),
               neighbor_index=np.stack(NB).astype(np.int8), in_flags=np.stack(INF),
               sha_obs=np.array(h_obs), sha_sensed=np.array(h_sen), sha_occupied=np.array(h_occ),
               full_steps=np.array(full_steps), full_obs=np.stack(F_obs), full_sensed=np.stack(F_sen),
               full_occupied=np.stack(F_occ))
    return out


if __name__ == "__main__":
    for name, n_a, seed, mode, steps, fe in CASES:
        d = record(n_a, seed, mode, steps, fe)
        path = os.path.join(HERE, f"traj_{name}.npz")
        np.savez_compressed(path, **d)
        print(name, "in_shape agent-steps:", int(d["in_flags"].sum()), "reward:", float(d["reward"].sum()),
              "size KB:", os.path.getsize(path) // 1024)
