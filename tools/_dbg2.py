import sys, traceback
sys.path.insert(0, "/root/repo")
import torch
import torchopt_b200 as tb
from bench import RESNET18_LEAVES, EPS_ROOT
mode = sys.argv[1]
dev = torch.device("cuda")
torch.manual_seed(4321)
L = RESNET18_LEAVES[:8]
params = [(torch.randn(k, device=dev) * 0.1).requires_grad_(True) for k in L]
ws = [torch.rand(k, device=dev) + 0.5 for k in L]
lay = tb.FlatLayout(params)
w_flat = lay.pack(ws)

def once():
    opt = tb.FuncAdam(1e-3, eps_root=EPS_ROOT, moment_requires_grad=False)
    cur = list(params)
    for _ in range(2):
        cur = opt.apply_gradients([w * p for w, p in zip(ws, cur)], cur)
    outer = torch.stack([p.sum() for p in cur]).sum()
    return torch.autograd.grad(outer, params)

def once_flat():
    opt = tb.FuncAdam(1e-3, eps_root=EPS_ROOT, moment_requires_grad=False)
    cur = lay.pack(params)
    for _ in range(2):
        (cur,) = opt.apply_gradients([w_flat * cur], [cur])
    return torch.autograd.grad(cur.sum(), params)

def cap(fn, name):
    try:
        s = tb.graphs.capture(fn); s(); torch.cuda.synchronize(); print(mode, name, "capture OK")
        return s
    except Exception as e:
        print(mode, name, "capture FAILED:", str(e).splitlines()[0]); sys.exit(0)

if mode == "A":
    cap(once_flat, "flat")
elif mode == "B":
    once_flat(); cap(once_flat, "flat")
elif mode == "C":
    s = cap(once, "leaf"); del s; cap(once_flat, "flat")
elif mode == "D":
    s = cap(once, "leaf"); cap(once_flat, "flat (first capture alive)")
elif mode == "E":
    once(); s = cap(once, "leaf"); got = s(); del s; once_flat(); cap(once_flat, "flat")
elif mode == "F":
    s = cap(once_flat, "flat"); s2 = cap(once_flat, "flat again")
