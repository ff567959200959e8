"""Experiment: find() kernel time, cold (L2 flushed) vs warm, via CUDA events. env B2_FIND_PREFETCH selects the prefetch mode."""
import os, sys
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
stream = torch.cuda.current_stream()
h.setStream(stream.cuda_stream)
h.setTsb(Tsb); h.setModel(m)
Tom = synth.compose(Tgt, synth.scenario_pose_offset())
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
flush2 = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def run(cold, n=30, readflush=False):
    ts = []
    for i in range(n + 5):
        if cold:
            flush.fill_(i & 255)
            if readflush:
                flush2.copy_(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); h.find(Tom); b.record(stream)
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    return np.median(ts), np.min(ts)

print("prefetch mode", os.environ.get("B2_FIND_PREFETCH", "default(1)"))
print("cold (write flush)      : median %.1f us  min %.1f us" % run(True))
print("cold (write+copy flush) : median %.1f us  min %.1f us" % run(True, readflush=True))
print("warm                    : median %.1f us  min %.1f us" % run(False))
