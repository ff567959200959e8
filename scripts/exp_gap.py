"""Where does the time between the end of k_rcc_find and the end of k_icp_loop go?  %globaltimer stamps from both kernels (profiling aid)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth

V, F = synth.building(1_000_000)
gmap = rmcl_b200.Map(V, F)
m = synth.c2_sensor()
Tsb, Tgt, I = synth.scenario_tsb(), synth.building_gt_pose(), synth.make_transform()
h = rmcl_b200.RCCB200Spherical(gmap)
h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
h.find(Tgt)
h.setRanges(synth.noisy_ranges(h.modelView()["ranges"], m.range_max))
Tom = synth.compose(Tgt, synth.scenario_pose_offset())
lib = rmcl_b200.load_library()
nw = (m.size + 31) // 32
buf = torch.zeros(2 * nw, dtype=torch.int64, device="cuda")
for _ in range(5):
    h.correctOnce(Tom, I, 5, 0.0)
lib.b2_rcc_debug_find_warp_times(h._h, C.c_void_p(buf.data_ptr()))
for k in range(8):
    h.enableTiming(k >= 4)
    t0 = time.perf_counter()
    h.correctOnce(Tom, I, 5, 0.0)
    wall = (time.perf_counter() - t0) * 1e6
    out = (C.c_ulonglong * 16)()
    lib.b2_rcc_debug_clocks(h._h, out)
    t = buf.cpu().numpy().reshape(-1, 2)
    f0, f1 = int(t[:, 0].min()), int(t[:, 1].max())
    print("find: first warp start 0, last warp end %.1f us | icp loop block 0: start %.1f, end %.1f us | host wall clock of the call %.1f us | events (find, loop) ms %s"
          % ((f1 - f0) / 1e3, (int(out[6]) - f0) / 1e3, (int(out[7]) - f0) / 1e3, wall, h.lastTiming() if k >= 4 else "off"))
lib.b2_rcc_debug_find_warp_times(h._h, None)
