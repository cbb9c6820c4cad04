"""Times k_rollout_push / k_rollout_gather at the BASELINE config-3 shape (65536 envs x 30 agents x 192 features, fp32)
with CUDA events; prints achieved GB/s against MEASURED_PEAKS.json (profiling aid for DESIGN.md §4)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_llm_b200.rollout import ReplayBufferAgent

E, n_a, D, A = 65536, 30, 192, 2
steps = 6
buf = ReplayBufferAgent(steps, E * n_a, slice(0, n_a), D, A)
obs = [torch.randn(E, D, n_a, device="cuda") for _ in range(3)]
act = torch.rand(E, A, n_a, device="cuda"); prior = torch.rand(E, A, n_a, device="cuda")
rew = torch.rand(E, 1, n_a, device="cuda"); done = torch.zeros(E, 1, n_a, dtype=torch.bool, device="cuda")
idx = slice(0, n_a)
for k in range(3):
    buf.push(obs[k % 3], act, rew, obs[(k + 1) % 3], done, idx, prior)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
K = 12
ev[0].record()
for k in range(K):
    buf.push(obs[k % 3], act, rew, obs[(k + 1) % 3], done, idx, prior)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / K
rows = E * n_a
bytes_push = rows * (2 * D * 4 * 2 + 2 * A * 4 * 2 + 4 * 2 + 1 + 4)      # read + write of obs, next_obs, act, prior, rew; done
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
print(json.dumps({"kernel": "k_rollout_push", "rows": rows, "ms": ms, "GBps": bytes_push / ms / 1e6, "frac_of_hbm_peak": bytes_push / ms / 1e6 / peak,
                  "agent_rows_per_s": rows / ms * 1e3}))
N = 1 << 20
inds = torch.randint(0, buf.total_length, (N,), device="cuda")
buf.gather(inds, is_prior=True); torch.cuda.synchronize()
ev[0].record()
for k in range(5):
    buf.gather(inds, is_prior=True)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
bytes_g = N * (2 * D * 4 * 2 + 2 * A * 4 * 2 + 16 + 8)
print(json.dumps({"kernel": "k_rollout_gather", "rows": N, "ms": ms, "GBps": bytes_g / ms / 1e6, "frac_of_hbm_peak": bytes_g / ms / 1e6 / peak}))
