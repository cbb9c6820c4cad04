"""(f2) eval metrics: the NumPy restatement in oracle/metrics_oracle.py against the values recorded from the real
AssemblySwarmWrapper (tests/golden/metrics.npz, any box) and against the live wrapper (build container only)."""
import os

import numpy as np
import pytest

from oracle import live_reference as lr
from oracle import metrics_oracle as mo
from tests.helpers import GOLDEN, goal_seeking_action


def test_restatement_matches_recorded_reference_metrics():
    z = np.load(os.path.join(GOLDEN, "metrics.npz"))
    covered = 0.0
    for k in range(int(z["n_cases"])):
        grid, r = z[f"c{k}_grid"], float(z[f"c{k}_r_avoid"])
        for p, want in zip(z[f"c{k}_p"], z[f"c{k}_metrics"]):
            got = mo.all_metrics(p, grid, r)
            assert got[0] == want[0]                                     # a count ratio: exact
            assert np.allclose(got[1:], want[1:], rtol=1e-13, atol=0)     # np.var's summation order may differ across NumPy builds
            covered = max(covered, want[0])
    assert covered > 0.3                                                  # the rollouts really filled the shapes


@pytest.mark.reference
@pytest.mark.skipif(not lr.available(), reason="reference checkout / oracle/_ref not available")
def test_restatement_matches_the_live_wrapper():
    env = lr.make_env(30)
    np.random.seed(5)
    env.reset()
    e = env.env
    rng = np.random.RandomState(5)
    for t in range(60):
        env.step(goal_seeking_action(e.obs, e.dp, rng))
        if t % 6 == 5:
            want = np.array([env.coverage_rate(), env.distribution_uniformity(), env.voronoi_based_uniformity()])
            assert np.array_equal(mo.all_metrics(e.p, e.grid_center, e.r_avoid), want)
