import sys, copy
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/generative-turbulence_b200"); sys.path.insert(0, "/root/repo/tests")
import torch
from turbdiff_b200.optim import FusedRAdam
from util import rel_l2
torch.manual_seed(0)
ps = [torch.nn.Parameter(torch.randn(257, 33, device="cuda")), torch.nn.Parameter(torch.randn(1000, device="cuda"))]
ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
a, b = FusedRAdam(ps, lr=1e-2), torch.optim.RAdam(ref, lr=1e-2)
def step(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for p, q in zip(ps, ref):
        p.grad = torch.randn(p.shape, device="cuda", generator=g)
        q.grad = p.grad.clone()
    a.step(); b.step()
    print(seed, [rel_l2(p, q) for p, q in zip(ps, ref)], [rel_l2(a.state[p]["exp_avg_sq"], b.state[q]["exp_avg_sq"]) for p, q in zip(ps, ref)], [float(a.state[p]["step"]) for p in ps], [float(b.state[q]["step"]) for q in ref])
for i in range(7): step(i)
a.load_state_dict(copy.deepcopy(b.state_dict()))
print("loaded", a.param_groups[0].keys())
for i in range(7, 10): step(i)
