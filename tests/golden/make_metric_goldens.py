#!/usr/bin/env python
"""Record the three wrapper metrics from the UNMODIFIED reference (build container only) -> tests/golden/metrics.npz.

TEST/FIXTURE INFRASTRUCTURE.  Drives the real AssemblySwarmWrapper (assembly_wrapper.py:48-129) over goal-seeking rollouts
(so cells get covered and the Voronoi counts spread) and stores, at every 10th step, the state and the metric values the
reference computed.  The GPU-side test compares swarm_metrics / k_metrics with these numbers."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import live_reference as lr          # noqa: E402
from tests.helpers import goal_seeking_action    # noqa: E402

CASES = [(30, 11, 200), (30, 12, 200), (10, 13, 120), (64, 14, 80)]

if __name__ == "__main__":
    out = {}
    for k, (n_a, seed, steps) in enumerate(CASES):
        env = lr.make_env(n_a)
        np.random.seed(seed)
        env.reset()
        e = env.env
        rng = np.random.RandomState(seed)
        P, M = [], []
        for t in range(steps):
            env.step(goal_seeking_action(e.obs, e.dp, rng))
            if t % 10 == 9:
                P.append(e.p.copy())
                M.append([env.coverage_rate(), env.distribution_uniformity(), env.voronoi_based_uniformity()])
        out[f"c{k}_p"], out[f"c{k}_metrics"] = np.stack(P), np.array(M)
        out[f"c{k}_grid"], out[f"c{k}_r_avoid"], out[f"c{k}_l_cell"] = e.grid_center.copy(), e.r_avoid, e.l_cell
        print(n_a, seed, "coverage up to", np.max(np.array(M)[:, 0]))
    out["n_cases"] = len(CASES)
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
