import sys; sys.path.insert(0,'/root/repo')
import torch
from torchopt_b200 import abi
for dt, tdt, name, bf, bb in ((abi.F64, torch.float64, 'f64', 48, 72), (abi.F32, torch.float32, 'f32', 24, 36)):
    n = 1 << 27
    t = [torch.rand(n, device='cuda', dtype=tdt) * 1e-2 + 1e-4 for _ in range(6)]
    o = [torch.empty(n, device='cuda', dtype=tdt) for _ in range(6)]
    P = [[x.data_ptr()] for x in t + o]
    prep = abi.PreparedFused(dt, [P[0], P[1], P[2], P[6], P[7], P[8]], [P[3], P[8], P[6], P[0], P[4], P[5], P[9], P[10], P[11]], [n], [3], 0.9, 0.999, 1e-8, 1e-8)
    s = torch.cuda.current_stream().cuda_stream
    for nm, fn, nb in (('fwd', prep.forward, bf*n), ('bwd', prep.backward, bb*n)):
        for _ in range(3): assert fn(s) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn(s)
        e1.record(); torch.cuda.synchronize()
        print(name, nm, '%.0f GB/s' % (nb / (e0.elapsed_time(e1)/10*1e-3) / 1e9))
