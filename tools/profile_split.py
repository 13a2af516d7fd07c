#!/usr/bin/env python3
"""ncu driver for the DMC quad -> triangle split (default return_quads=False path) on a random-init grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diso_b200
from diso_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sdf = syn.random_sdf(n, "flexi", 0).cuda()
deform = syn.random_deform(n, 1000).cuda()
m = diso_b200.DiffDMC()
for _ in range(2):
    v, f = m(sdf, deform)
torch.cuda.synchronize()
print("done", v.shape, f.shape)
