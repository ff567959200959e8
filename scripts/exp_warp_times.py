"""Per-warp start/end times of k_rcc_find on the C2 scan (profiling aid): how much of the kernel is tail?"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth

V, F = synth.building(1_000_000)
gmap = rmcl_b200.Map(V, F)
m = synth.c2_sensor()
Tsb, Tgt = synth.scenario_tsb(), synth.building_gt_pose()
h = rmcl_b200.RCCB200Spherical(gmap)
h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
Tbm = synth.compose(Tgt, synth.scenario_pose_offset())
lib = rmcl_b200.load_library()
nw = (m.size + 31) // 32
buf = torch.zeros(2 * nw, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    h.find(Tbm)
torch.cuda.synchronize()
lib.b2_rcc_debug_find_warp_times(h._h, C.c_void_p(buf.data_ptr()))
for cold in (True, False):
    if cold:
        flush.fill_(1)
    torch.cuda.synchronize()
    h.find(Tbm)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(-1, 2).astype(np.float64)
    t0 = t[:, 0].min()
    st, en = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    dur = en - st
    print(("cold" if cold else "warm"), "kernel span %.1f us | warp start: median %.1f max %.1f | warp duration: median %.1f p90 %.1f p99 %.1f max %.1f | end: median %.1f p90 %.1f p99 %.1f"
          % (en.max(), np.median(st), st.max(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), np.median(en), np.percentile(en, 90), np.percentile(en, 99)))
    alive = [(en > x).sum() for x in np.linspace(0, en.max(), 11)]
    print("  warps still running at 0,10,..,100 % of the span:", alive)
lib.b2_rcc_debug_find_warp_times(h._h, None)
