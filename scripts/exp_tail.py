import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for k in range(3):
    h.correctOnce(Tom, I, 5, 0.0)
    out = (C.c_ulonglong * 16)()
    lib.b2_rcc_debug_clocks(h._h, out)
    print("cycles (coop loop, iteration 1, block 0): pass+blockreduce %d | grid.sync %d | partial sum %d | tail %d || prologue (state + pair loads) %d | whole kernel in block 0: %d cycles = %d ns (globaltimer)" % (out[0], out[1], out[2], out[3], out[4], out[5], out[6]))
h.enableTiming(True)
for k in range(5):
    h.correctOnce(Tom, I, 5, 0.0)
print("event times (find_ms, loop_ms):", h.lastTiming())
