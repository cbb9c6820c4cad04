"""Acceptance tests of the drop-in boundary (north star: "train_assembly.py and eval_assembly.py run
unchanged"): the reference's OWN files, staged byte for byte under baseline/_ref (tools/stage_reference.py, MANIFEST.json),
run on top of the B200 library.

  1. the real assembly.py (+ its Gym fork, wrapper and c_lib.py) bound to libswarm_b200.so replays the golden trajectories
     recorded from the reference on its own C++ — bit for bit (INTEGRATION.md §A under test: the five legacy symbols in the
     order and with the buffers the reference's step() uses, assembly.py:234-255, 357-380, 495-504, 613-624);
  2. train_assembly.py (2 episodes) then eval_assembly.py (its 300 steps incl. render(), the shape swap and the three
     metrics) exit 0, unmodified, in both arrangements: 'legacy' (reference Gym fork + real env on the legacy ABI) and
     'dropin' (marl_llm_b200/compat's gym -> drop-in class on the batched ABI).
"""
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from tests import ref_scripts as rs
from tests.helpers import GOLDEN_CASES, REPO

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rs.staged(), reason="baseline/_ref not staged (tools/stage_reference.py)")]

REPLAY = r'''
import sys, os, numpy as np
sys.path.insert(0, {repo!r}); sys.path.insert(0, {shims!r}); sys.path.append({cus_gym!r})
import gym
from gym.envs.customized_envs import assembly
import gym.wrappers
from oracle import live_reference as lr           # only default_args / write_results_pkl (shape pickle rebuilt from shapes.npz)
from tests.helpers import load_golden, replay_golden
assert os.path.realpath(gym.__file__).startswith(os.path.realpath({cus_gym!r})), gym.__file__
loaded = os.path.realpath(assembly._LIB._name)
assert loaded == os.path.realpath({lib!r}), loaded
for case in {cases!r}:
    g = load_golden(case)
    n_a = int(g["n_a"])
    env = gym.wrappers.AssemblySwarmWrapper(gym.make("AssemblySwarm-v0").unwrapped,
                                            lr.default_args(n_a=n_a, is_boundary=not bool(g.get("is_periodic", 0))))
    e = env.env
    def snap(obs, rew=None, prior=None):
        return dict(p=e.p, dp=e.dp, obs=obs, reward=rew, a_prior=prior, nbr=e.neighbor_index, in_flags=e.in_flags,
                    sensed=e.sensed_index, occupied=e.occupied_index)
    def reset_fn(g):
        np.random.seed(int(g["seed"]))
        obs = env.reset()
        assert np.array_equal(e.grid_center, g["grid_center"]) and np.array_equal(e.p, g["p0"])
        return snap(obs)
    def step_fn(a):
        obs, rew, done, info, prior = env.step(a)
        assert not done.any()
        return snap(obs, rew, prior)
    replay_golden(g, reset_fn, step_fn)
    print("REPLAYED", case, int(g["steps"]), "steps")
print("ALL_OK")
'''


def test_real_assembly_py_on_libswarm_b200_replays_goldens():
    from marl_llm_b200 import _lib
    rs.verify_manifest(["cus_gym/gym/envs/customized_envs/assembly.py", "cus_gym/gym/envs/customized_envs/envs_cplus/c_lib.py",
                        "cus_gym/gym/wrappers/customized_envs/assembly_wrapper.py"])
    lib = rs.install_library(_lib.LIB_PATH)
    code = REPLAY.format(repo=REPO, shims=rs.SHIMS, cus_gym=os.path.join(rs.REF, "cus_gym"), lib=lib, cases=GOLDEN_CASES)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=REPO)
    assert r.returncode == 0 and "ALL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("REPLAYED") == len(GOLDEN_CASES)


@pytest.mark.parametrize("mode", ["legacy", "dropin"])
def test_train_and_eval_scripts_run_unchanged(mode, tmp_path):
    from marl_llm_b200 import _lib
    rs.verify_manifest(["marl_llm/train/train_assembly.py", "marl_llm/eval/eval_assembly.py", "marl_llm/cfg/assembly_cfg.py",
                        "marl_llm/algorithm/algorithms/maddpg.py", "cus_gym/gym/envs/customized_envs/assembly.py"])
    if mode == "legacy":
        rs.install_library(_lib.LIB_PATH)
    cwd = rs.make_workdir(str(tmp_path))
    # train_assembly.py:75-170: 2 episodes x 200 steps, render() at episode 0, 20 update rounds per episode, checkpoints
    r = rs.run_script("marl_llm/train/train_assembly.py", mode, cwd, ["--n_episodes", "2"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "Training Starts..." in r.stdout and "Episodes 0 of 2" in r.stdout
    runs = glob.glob(os.path.join(cwd, "models", "assembly", "*"))
    assert len(runs) == 1 and os.path.isfile(os.path.join(runs[0], "model.pt")) and os.path.isfile(os.path.join(runs[0], "logs", "summary.json"))
    # eval_assembly.py:81 hard-codes the run name it evaluates
    shutil.move(runs[0], os.path.join(cwd, "models", "assembly", "your_run_name"))
    import json
    summ = json.load(open(os.path.join(cwd, "models", "assembly", "your_run_name", "logs", "summary.json")))
    # eval_assembly.py:213-218 looks its series up by the ORIGINAL log dir string; keep them reachable under the new name
    fixed = {k.replace(os.path.basename(runs[0]), "your_run_name"): v for k, v in summ.items()}
    json.dump(fixed, open(os.path.join(cwd, "models", "assembly", "your_run_name", "logs", "summary.json"), "w"))
    r = rs.run_script("marl_llm/eval/eval_assembly.py", mode, cwd)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("Coverage:") == 300 and "Episode 1/" in r.stdout
    res = os.path.join(cwd, "models", "assembly", "your_run_name", "results")
    z = np.load(os.path.join(res, "state_data.npz"))
    assert z["pos"].shape == (2, 30, 300) and np.isfinite(z["pos"]).all() and os.path.isfile(os.path.join(res, "metrics.pkl"))
    # the policy is untrained after 2 episodes, but the physics ran: agents moved and stayed inside the soft walls
    assert np.abs(z["pos"][:, :, -1] - z["pos"][:, :, 0]).max() > 0.05 and np.abs(z["pos"]).max() < 2.6
