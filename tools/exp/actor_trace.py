import torch, sys
sys.path.insert(0, '/root/repo')
from marinevehiclereinforcementlearning_b200 import MlpGaussianPolicy
pol = MlpGaussianPolicy(9, 6, device="cuda", seed=1)
n = 131072
obs = torch.rand((9, n), device="cuda") * 2 - 1
act = torch.zeros((6, n), device="cuda")
for k in range(5):
    pol.act_into(obs, act, n, step=k)
torch.cuda.synchronize()
