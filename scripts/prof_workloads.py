"""Short run of every kernel for ncu (few launches each).  python scripts/prof_workloads.py [faces]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth

faces = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
V, F = synth.building(faces)
gmap = rmcl_b200.Map(V, F)
m = synth.c2_sensor()
Tsb, Tgt, I = synth.scenario_tsb(), synth.building_gt_pose(), synth.make_transform()
h = rmcl_b200.RCCB200Spherical(gmap)
h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
h.find(Tgt)
ranges = synth.noisy_ranges(h.modelView()["ranges"], m.range_max)
h.setRanges(ranges)
Tom = synth.compose(Tgt, synth.scenario_pose_offset())
for _ in range(steps):
    h.correctOnce(Tom, I, 5, 0.0)
pts = h.modelView()["points"]
beams = synth.pf_beams(pts, 180)
P, A = synth.pf_particles(100_000)
up = rmcl_b200.PCDSensorUpdaterB200(gmap)
Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
for _ in range(2):
    up.update(Pd, Ad, Tsb, beams)
torch.cuda.synchronize()
hv = rmcl_b200.SphereCorrectorB200(gmap)
hv.setTsb(Tsb)
mv = synth.vlp16_900(); mv.range_min = 0.0
hv.setModel(mv); hv.find(Tgt); hv.setInputData(hv.modelView()["ranges"])
T = synth.transforms(1000); T[:] = Tom
T["t"] += np.random.default_rng(0).uniform(-0.05, 0.05, (1000, 3)).astype(np.float32)
Td = torch.from_numpy(T.view(np.float32).reshape(-1, 8).copy()).cuda()
for _ in range(2):
    hv.correct(Td)
torch.cuda.synchronize()
# closest-point correspondences on the same scan; 8(f2): the rest of the PF cycle
ds = h.datasetView()
hc = rmcl_b200.CPCB200(gmap)
hc.setTsb(Tsb); hc.setParams(1.0, 0.15); hc.setDataset(ds["points"], ds["mask"])
for _ in range(2):
    hc.correctOnce(Tom, I, 5, 0.0)
up.update(Pd, Ad, Tsb, beams, rmcl_b200.PFParams.defaults(0, 1))
up.motionUpdate(Pd, Ad, synth.make_transform((0.02, 0, 0), (0, 0, 0.01)), 0.01)
up.likelihoodStats(Ad)
Pn, An = torch.empty_like(Pd), torch.empty_like(Ad)
up.resample(Pd, Ad, Pn, An)
torch.cuda.synchronize()
# C4 (BASELINE config 4): pinhole 640 x 480 on the 500k-triangle indoor mesh -- k_rcc_find#4800 and the ICP loop with pairs in shared memory
V4, F4 = synth.indoor(500_000)
map4 = rmcl_b200.Map(V4, F4)
m4 = synth.c4_sensor()
h4 = rmcl_b200.RCCB200Pinhole(map4)
h4.setTsb(Tsb); h4.setModel(m4); h4.setParams(1.0, 0.15)
T4 = synth.indoor_gt_pose()
h4.find(T4)
h4.setRanges(synth.noisy_ranges(h4.modelView()["ranges"], m4.range_max))
for _ in range(2):
    h4.correctOnce(synth.compose(T4, synth.scenario_pose_offset()), I, 5, 0.0)
torch.cuda.synchronize()
print("done", rmcl_b200.kernel_launch_count())
