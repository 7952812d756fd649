import sys, gc, os, warnings
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

def f_opt():
    opt = tb.FuncAdam(1e-3, eps_root=1e-8, moment_requires_grad=False)
    cur = lay.pack(params)
    (cur,) = opt.apply_gradients([w_flat * cur], [cur])
    return torch.autograd.grad(cur.sum(), params)

def f_chain():
    opt = tb.adam(1e-3, eps_root=1e-8)
    cur = lay.pack(params)
    st = opt.init([cur])
    (u,), st = opt.update([w_flat * cur], st, params=[cur], inplace=False)
    cur = cur + u
    return torch.autograd.grad(cur.sum(), params)

def f_passthrough():
    cur = lay.pack(params)
    cur = cur + (w_flat * cur) * 0.1
    return torch.autograd.grad(cur.sum(), params)

fn = {"opt": f_opt, "chain": f_chain, "pass": f_passthrough}[mode.split("-")[0]]
r = fn()
if "gc" in mode:
    gc.collect()
if "clone" in mode:
    r = [x.clone() for x in r]
try:
    s = tb.graphs.capture(fn); s(); torch.cuda.synchronize(); print(mode, "OK")
except Exception as e:
    print(mode, "FAILED:", str(e).splitlines()[0])
