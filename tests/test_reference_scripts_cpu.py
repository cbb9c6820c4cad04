"""The acceptance harness itself (tests/ref_scripts.py + tests/shims), validated HERE on the reference's own C++ library:
the byte-identical train_assembly.py / eval_assembly.py run to completion in the arranged working directory.  The GPU-side
twin (tests/test_gpu_reference_scripts.py) runs the same harness with libswarm_b200.so underneath."""
import glob
import json
import os
import shutil

import numpy as np
import pytest

from tests import ref_scripts as rs
from tests.helpers import REPO

REF_SO = os.path.join(REPO, "oracle", "_ref", "libAssemblyEnv.so")
pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not (rs.staged() and os.path.isfile(REF_SO)), reason="baseline/_ref or oracle/_ref not available")]


def test_staged_files_are_the_reference_bytes():
    man = json.load(open(os.path.join(rs.REF, "MANIFEST.json")))
    assert "marl_llm/train/train_assembly.py" in man and "cus_gym/gym/envs/customized_envs/assembly.py" in man
    rs.verify_manifest(list(man))


def test_scripts_run_unchanged_on_the_reference_library(tmp_path):
    rs.install_library(REF_SO)
    cwd = rs.make_workdir(str(tmp_path))
    r = rs.run_script("marl_llm/train/train_assembly.py", "legacy", cwd, ["--n_episodes", "1", "--episode_length", "40"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    runs = glob.glob(os.path.join(cwd, "models", "assembly", "*"))
    assert len(runs) == 1
    dst = os.path.join(cwd, "models", "assembly", "your_run_name")
    shutil.move(runs[0], dst)
    sj = os.path.join(dst, "logs", "summary.json")
    summ = json.load(open(sj))
    json.dump({k.replace(os.path.basename(runs[0]), "your_run_name"): v for k, v in summ.items()}, open(sj, "w"))
    r = rs.run_script("marl_llm/eval/eval_assembly.py", "legacy", cwd)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("Coverage:") == 300
    z = np.load(os.path.join(dst, "results", "state_data.npz"))
    assert z["pos"].shape == (2, 30, 300) and np.isfinite(z["pos"]).all()
