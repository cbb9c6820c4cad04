"""NumPy restatement of the reference wrapper's three evaluation metrics (TEST INFRASTRUCTURE ONLY — see oracle/oracle.py).

WRAP = cus_gym/gym/wrappers/customized_envs/assembly_wrapper.py.  Pinned to the real AssemblySwarmWrapper by
tests/test_metrics_oracle.py (live, build container) and through tests/golden/metrics.npz (recorded from the real wrapper
by tests/golden/make_metric_goldens.py)."""
import numpy as np


def coverage_rate(p, grid_center, r_avoid):
    """WRAP:48-72: fraction of cells with an agent strictly within r_avoid/2."""
    n_g = grid_center.shape[1]
    occupied = 0
    for c in range(n_g):
        rel = p - grid_center[:, [c]]
        if (np.sqrt(rel[0] * rel[0] + rel[1] * rel[1]) < r_avoid / 2).any():      # np.linalg.norm(axis=0)
            occupied += 1
    return occupied / n_g


def distribution_uniformity(p):
    """WRAP:74-101: (var(m) - min(m)) / (max(m) - min(m)), m_i = distance from agent i to its nearest non-coincident agent."""
    n_a = p.shape[1]
    mins = []
    for i in range(n_a):
        rel = p - p[:, [i]]
        d = np.sqrt(rel[0] * rel[0] + rel[1] * rel[1])
        mins.append(np.min(d[d != 0]))
    return (np.var(mins) - np.min(mins)) / (np.max(mins) - np.min(mins))


def voronoi_based_uniformity(p, grid_center):
    """WRAP:103-129: same statistic on the number of cells whose nearest agent (first minimum, np.argmin) is i."""
    cnt = np.zeros(p.shape[1])
    for c in range(grid_center.shape[1]):
        rel = p - grid_center[:, [c]]
        cnt[np.argmin(np.sqrt(rel[0] * rel[0] + rel[1] * rel[1]))] += 1
    return (np.var(cnt) - np.min(cnt)) / (np.max(cnt) - np.min(cnt))


def all_metrics(p, grid_center, r_avoid):
    return np.array([coverage_rate(p, grid_center, r_avoid), distribution_uniformity(p), voronoi_based_uniformity(p, grid_center)])
