"""TEST SHIM (not product code): the two SummaryWriter features the reference's scripts use — add_scalars
(train_assembly.py:157-161, maddpg.py:185-190) and export_scalars_to_json (train_assembly.py:173), in tensorboardX's own
JSON layout ({logdir/main_tag/tag: [[wall_time, step, value], ...]}), which eval_assembly.py:209-218 reads back."""
import json
import os
import time


class SummaryWriter:
    def __init__(self, logdir=None, *a, **k):
        self.logdir = str(logdir) if logdir is not None else "runs"
        os.makedirs(self.logdir, exist_ok=True)
        self.scalar_dict = {}

    def add_scalar(self, tag, value, global_step=None, walltime=None):
        self.scalar_dict.setdefault(tag, []).append([walltime or time.time(), global_step, float(value)])

    def add_scalars(self, main_tag, tag_scalar_dict, global_step=None, walltime=None):
        for tag, value in tag_scalar_dict.items():
            key = self.logdir + "/" + main_tag + "/" + tag
            self.scalar_dict.setdefault(key, []).append([walltime or time.time(), global_step, float(value)])

    def export_scalars_to_json(self, path):
        with open(path, "w") as f:
            json.dump(self.scalar_dict, f)

    def flush(self):
        pass

    def close(self):
        pass
