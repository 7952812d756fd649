import sys, gc
sys.path.insert(0, "/root/repo")
import torch
import torchopt_b200 as tb
mode = sys.argv[1]
dev = torch.device("cuda")
L = [9408, 64, 64, 36864]
params = [(torch.randn(k, device=dev) * 0.1).requires_grad_(True) for k in L]
ws = [torch.rand(k, device=dev) + 0.5 for k in L]
lay = tb.FlatLayout(params)
w_flat = lay.pack(ws)

def f_pack_only():
    cur = lay.pack(params)
    return torch.autograd.grad((cur * w_flat).sum(), params)

def f_cat_only():
    cur = torch.cat([p.reshape(-1) for p in params])
    return torch.autograd.grad((cur * cur).sum(), params)

def f_cat_mul():
    cur = torch.cat([p * 1.0 for p in params])
    return torch.autograd.grad((cur * cur).sum(), params)

def f_opt():
    opt = tb.FuncAdam(1e-3, eps_root=1e-8, moment_requires_grad=False)
    cur = lay.pack(params)
    (cur,) = opt.apply_gradients([w_flat * cur], [cur])
    return torch.autograd.grad(cur.sum(), params)

def f_leaf_sum():
    return torch.autograd.grad(torch.stack([(p * p).sum() for p in params]).sum(), params)

fn = {"pack": f_pack_only, "cat": f_cat_only, "catmul": f_cat_mul, "opt": f_opt, "leaf": f_leaf_sum}[mode]
r = fn()
if len(sys.argv) > 2 and sys.argv[2] == "del":
    del r
    gc.collect()
try:
    s = tb.graphs.capture(fn); s(); torch.cuda.synchronize(); print(mode, sys.argv[2:], "OK")
except Exception as e:
    print(mode, sys.argv[2:], "FAILED:", str(e).splitlines()[0])
