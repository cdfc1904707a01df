import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from rl4rs_b200.policy import MaskedPolicy
from rl4rs_b200.trainer import KernelOps
dev = torch.device("cuda"); A = 284; n = 36864; mb = 256
pol = MaskedPolicy(A, dev); ops = KernelOps(A, dev, pol.n_params)
g = torch.Generator().manual_seed(0)
obs = torch.randn(n, 256, generator=g).to(dev); mask = torch.ones(n, A, dtype=torch.uint8, device=dev)
act = torch.randint(0, A, (n,), generator=g).to(dev); lg = torch.randn(n, A, generator=g).to(dev)
lp = torch.log_softmax(lg, -1).gather(1, act[:, None]).squeeze(1).contiguous(); v = torch.randn(n, generator=g).to(dev)
adv = torch.randn(n, generator=g).to(dev); tg = torch.randn(n, generator=g).to(dev)
data = (obs, mask, act, lp, lg, v, adv, tg); perm = torch.randperm(n, generator=g).to(dev)
hp = {"clip": 0.3, "vf_clip": 500.0, "vf_coeff": 0.5, "kl_coeff": 0.2, "ent_coeff": 0.0}
flat = pol.flat.detach().clone()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
print("ppo_epoch (144 steps): %.2f ms" % timeit(lambda: ops.ppo_epoch(flat, data, perm, n, mb, hp, 1e-4, None)))
def only_grad():
    for s in range(0, n, mb): ops.policy_grad(0, flat, data, perm, s, mb, hp, 1 / mb, 1 / mb)
def only_adam():
    for s in range(0, n, mb): ops.adam(flat, 1e-4, 1.0, None)
print("144 x policy_grad(+reduce): %.2f ms" % timeit(only_grad))
print("144 x adam: %.2f ms" % timeit(only_adam))
