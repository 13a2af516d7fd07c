import sys, torch
sys.path.insert(0, '.')
import diso_b200
from diso_b200 import _lib, synthetic as syn
n = 512
sdf = syn.random_sdf(n, "flexi", 0).cuda().requires_grad_(True)
deform = syn.random_deform(n, 1000).cuda().requires_grad_(True)
m = diso_b200.DiffDMC()
for rq in (True, False):
    for _ in range(2):
        v, f = m(sdf, deform, return_quads=rq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with _lib.kernel_profile() as prof:
        e0.record()
        for _ in range(5):
            v, f = m(sdf, deform, return_quads=rq)
        e1.record(); torch.cuda.synchronize()
    print("return_quads=%s forward only: %.3f ms, faces %s" % (rq, e0.elapsed_time(e1) / 5, tuple(f.shape)))
    for k, x in sorted(prof.times.items(), key=lambda kv: -sum(kv[1])):
        print("   %-20s %.3f ms" % (k, sum(x) / 5))
