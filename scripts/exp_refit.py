"""Refit vs rebuild of the 1M-triangle map, and trace speed on the refitted tree after a deformation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth
V, F = synth.building(1_000_000)
t0 = time.perf_counter(); m = rmcl_b200.Map(V, F); t_build = time.perf_counter() - t0
rng = np.random.default_rng(0)
V2 = (V + rng.normal(0, 0.01, V.shape)).astype(np.float32)
m.refit(V2)
t0 = time.perf_counter(); m.refit(V2); t_refit = time.perf_counter() - t0
sensor = synth.c2_sensor()
def find_us(mp):
    h = rmcl_b200.RCCB200Spherical(mp); h.setTsb(synth.scenario_tsb()); h.setModel(sensor)
    T = synth.compose(synth.building_gt_pose(), synth.scenario_pose_offset())
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); h.find(T); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))
print("build %.1f ms (incl. 24 MB upload) | refit %.2f ms (incl. 12 MB vertex upload) | find on refitted tree %.1f us | find on a fresh build of the moved mesh %.1f us"
      % (t_build * 1e3, t_refit * 1e3, find_us(m), find_us(rmcl_b200.Map(V2, F))))
