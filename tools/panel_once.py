#!/usr/bin/env python
"""One recursive LU panel factorisation (M x w) for a kernel-by-kernel launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/panel_once.py 8192 512"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import math_atlas_b200 as ab
from math_atlas_b200 import dist as abd
M = int(sys.argv[1]); w = int(sys.argv[2])
dev = torch.device("cuda", 0)
ops = abd.CudaOps("d")
W = torch.rand((w, M), device=dev, dtype=torch.float64) - 0.5
piv = torch.zeros(w, dtype=torch.int32, device=dev); info = torch.zeros(1, dtype=torch.int32, device=dev)
ops.getrf(M, w, W.data_ptr(), M, piv, info)
torch.cuda.synchronize()
print("launches", ab.launch_count())
