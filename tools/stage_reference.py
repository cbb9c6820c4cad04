"""Stage an UNMODIFIED copy of the reference files the acceptance tests execute into baseline/_ref/ (git-ignored, travels to
the GPU box with the gpurun snapshot; /root/reference itself does not exist there).

    baseline/_ref/cus_gym/      the reference's Gym fork incl. the real assembly.py / assembly_wrapper.py / c_lib.py
    baseline/_ref/marl_llm/     train/, eval/, cfg/, algorithm/ (the scripts that must run unchanged)
    baseline/_ref/fig/          the seven target-shape bitmaps cfg/assembly_cfg.py preprocesses at import
    baseline/_ref/MANIFEST.json sha256 of every staged file, so the GPU-side tests can prove they ran the reference's bytes

Test infrastructure: nothing in the product imports baseline/_ref.  Called by __graft_entry__.build() in the build container."""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SWARM_REF_ROOT", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
PARTS = ["cus_gym", "marl_llm/train", "marl_llm/eval", "marl_llm/cfg", "marl_llm/algorithm", "fig"]


def stage(force=False):
    if not os.path.isdir(os.path.join(REF, "cus_gym")):
        return False
    manifest_path = os.path.join(DST, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return True
    manifest = {}
    for part in PARTS:
        src, dst = os.path.join(REF, part), os.path.join(DST, part)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git*"))
        for root, dirs, files in os.walk(dst):
            os.chmod(root, 0o755)
            for f in files:
                path = os.path.join(root, f)
                os.chmod(path, 0o644)
                manifest[os.path.relpath(path, DST)] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    # the reference's own files are compared byte for byte with their source
    for rel, digest in manifest.items():
        assert hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() == digest, rel
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print("staged" if ok else f"{REF} not mounted: nothing staged", DST)
