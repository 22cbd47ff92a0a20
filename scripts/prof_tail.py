"""Development probe (GPU box): the dense-head tail of bench.py once, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, r3det_b200 as R
bench._time = lambda torch, fn, iters, warm=3: (fn(), fn(), 0.001)[2]        # two plain calls per timed section
print(bench.bench_dense_tail(torch, R, torch.device('cuda:0')))
