"""Sanity of the tcgen05 actor far beyond the test sizes: 1 000 003 environments (7813 tiles, ~53 per SM, a partial last tile) against the
mma.sync kernel - same noise bit for bit, means within the summation-order difference; and repeated launches are deterministic."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from marinevehiclereinforcementlearning_b200 import MlpGaussianPolicy
n = 1000003
ld = (n + 31) // 32 * 32
obs = torch.zeros((9, ld), device="cuda"); obs[:, :n] = torch.rand((9, n), device="cuda") * 2 - 1
out = []
for flag in ("0", "1"):
    os.environ["MVRL_POLICY_MMA_SYNC"] = flag
    p = MlpGaussianPolicy(9, 6, device="cuda", seed=3)
    for b in p.biases: b.uniform_(-0.3, 0.3, generator=torch.Generator().manual_seed(5))
    p.sync_weights()
    act, mean, eps = (torch.full((6, ld), 7.0, device="cuda") for _ in range(3)); logp = torch.full((ld,), 7.0, device="cuda")
    p.act_into(obs, act, n, logp=logp, mean=mean, eps=eps, env_id0=5, step=9)
    act2 = torch.full((6, ld), 7.0, device="cuda")
    p.act_into(obs, act2, n, env_id0=5, step=9)
    assert torch.equal(act, act2), "not deterministic"
    out.append((act, mean, eps, logp))
(a5, m5, e5, l5), (a1, m1, e1, l1) = out
assert torch.equal(e5, e1)
print("max |mean diff|", float((m5[:, :n] - m1[:, :n]).abs().max()), "mean", float((m5[:, :n] - m1[:, :n]).abs().mean()))
assert float((m5[:, :n] - m1[:, :n]).abs().max()) < 3e-3 and torch.allclose(l5[:n], l1[:n], atol=1e-6)
assert bool((a5[:, n:] == 7.0).all()) and bool(torch.isfinite(a5[:, :n]).all())
torch.cuda.synchronize(); print("actor ok at", n, "environments")
