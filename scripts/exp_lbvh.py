"""Device LBVH vs host SAH: build time, traversal counters and trace speed on the C2 scan."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth
V, F = synth.building(1_000_000)
m = synth.c2_sensor()
Tsb, Tgt = synth.scenario_tsb(), synth.building_gt_pose()
Tom = synth.compose(Tgt, synth.scenario_pose_offset())
for mode, name in ((0, "host SAH"), (1, "device LBVH")):
    t0 = time.perf_counter(); mp = rmcl_b200.Map(V, F, build_mode=mode); t1 = time.perf_counter()
    mp2 = rmcl_b200.Map(V, F, build_mode=mode); t2 = time.perf_counter()
    info = mp.info()
    h = rmcl_b200.RCCB200Spherical(mp)
    stream = torch.cuda.current_stream(); h.setStream(stream.cuda_stream)
    h.setTsb(Tsb); h.setModel(m)
    ts = []
    for i in range(25):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); h.find(Tom); b.record(stream); torch.cuda.synchronize()
        if i >= 5: ts.append(a.elapsed_time(b) * 1e3)
    print(f"{name}: build {1e3*(t1-t0):.1f} ms (2nd {1e3*(t2-t1):.1f} ms), nodes {info['n_nodes']}, depth {info['max_depth']}, bvh {info['bvh_bytes']/1e6:.1f} MB, find warm {np.median(ts):.1f} us")
