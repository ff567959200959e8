"""Small run of every entry point for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_workload.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth

V, F = synth.building(20000)
for mode in (1, 0):
    gmap = rmcl_b200.Map(V, F, build_mode=mode)
    m = synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 15, 16, -np.pi, 2 * np.pi / 128, 128, 0.5, 120.0)
    Tsb, Tgt, I = synth.scenario_tsb(), synth.building_gt_pose(), synth.make_transform()
    h = rmcl_b200.RCCB200Spherical(gmap)
    h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
    h.find(Tgt)
    ranges = synth.noisy_ranges(h.modelView()["ranges"], m.range_max)
    h.setRanges(ranges)
    Tom = synth.compose(Tgt, synth.scenario_pose_offset())
    h.correctOnce(Tom, I, 5, 0.0)
    h.correctOnce(Tom, I, 5, 0.3, ranges=ranges)
    h.correctOnce(Tom, I, 5, 0.0, ranges=torch.from_numpy(ranges.copy()).pin_memory())      # zero-copy scan + programmatic launch
    h.find(Tom); h.computeCrossStatistics(I, 0.0); h.segment(0.15, 0.15)
    T = synth.transforms(7); T[:] = Tom
    h.correct(T)
    ds = h.datasetView()
    hc = rmcl_b200.CPCB200(gmap); hc.setTsb(Tsb); hc.setParams(1.0, 0.15); hc.setDataset(ds["points"], ds["mask"])
    hc.find(Tom); hc.correctOnce(Tom, I, 3, 0.0)
    gmap.intersect(np.zeros((33, 3), np.float32) + [30, 20, 1], np.random.default_rng(0).normal(size=(33, 3)))
    beams = synth.pf_beams(h.modelView()["points"], 20)
    P, A = synth.pf_particles(333)
    up = rmcl_b200.PCDSensorUpdaterB200(gmap)
    A1 = up.update(P, A, Tsb, beams)
    up.update(P, A, Tsb, beams, rmcl_b200.PFParams.defaults(1, 1))
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A1.view(np.float32).reshape(-1, 9).copy()).cuda()
    up.motionUpdate(Pd, Ad, synth.make_transform((0.02, 0, 0), (0, 0, 0.01)), 0.01)
    up.motionUpdate(Pd, Ad, synth.make_transform((0.5, 0, 0), (0, 0, 0.01)), 0.01, check_collision=True)
    up.update(Pd, Ad, Tsb, beams)
    for mode in (1, 2, 0, 3):                                  # both ray mappings of the PF kernel, with and without the device-side particle sort
        up.setMapping(mode)
        up.update(Pd, Ad, Tsb, beams)
    P2, A2 = synth.pf_particles(40000)                        # the chunked two-stream host path
    up.update(P2, A2, Tsb, beams)
    up.likelihoodStats(Ad)
    Pn, An = torch.empty_like(Pd), torch.empty_like(Ad)
    up.resample(Pd, Ad, Pn, An)
    rmcl_b200.umeyama_transform(h.computeCrossStatistics(I, 0.0).reshape(1))
    torch.cuda.synchronize()
    # round 2: exec modes, queued asynchronous calls, a scan with pairs in shared memory and beyond (streamed), simulate switches, bound caller
    # buffers, two sensors in one loop, v1 benchmark stage split, peer-memory resampling on a one-device world
    for em in (1, 0, 2):
        h.setExecMode(em); h.correctOnce(Tom, I, 5, 0.0)
    for k in range(4):
        h.correctOnceAsync(Tom, I, 5, 0.0)
    for k in range(4):
        h.correctOnceWait()
    mb = synth.SphericalModel(np.radians(-30.0), np.radians(60.0) / 383, 384, -np.pi, 2 * np.pi / 1024, 1024, 0.5, 120.0)      # 393 216 rays: registers + shared memory
    hb = rmcl_b200.RCCB200Spherical(gmap)
    hb.setTsb(Tsb); hb.setModel(mb); hb.setParams(1.0, 0.15)
    hb.find(Tgt); rb = synth.noisy_ranges(hb.modelView()["ranges"], mb.range_max)
    hb.setRanges(rb); hb.correctOnce(Tom, I, 3, 0.0); hb.correctOnce(Tom, I, 3, 0.0, ranges=torch.from_numpy(rb.copy()).pin_memory())
    mh = synth.SphericalModel(np.radians(-30.0), np.radians(60.0) / 1023, 1024, -np.pi, 2 * np.pi / 1024, 1024, 0.5, 120.0)    # 1 048 576 rays: the tail of each thread's list streams from L2
    hb.setModel(mh); hb.find(Tgt); hb.setRanges(synth.noisy_ranges(hb.modelView()["ranges"], mh.range_max)); hb.correctOnce(Tom, I, 2, 0.0)
    h.setSimOptions(1, 1, 1); h.find(Tom); h.setSimOptions(0, 0, 0)
    dsd = torch.from_numpy(ds["points"]).cuda(); dmd = torch.from_numpy(ds["mask"]).cuda()
    bp, bn, bh = torch.empty((m.size, 3), device="cuda"), torch.empty((m.size, 3), device="cuda"), torch.empty(m.size, dtype=torch.uint8, device="cuda")
    h.bindDataset(dsd, dmd); h.bindModelBuffers(bp, bn, bh); h.find(Tom); h.computeCrossStatistics(I, 0.0); h.correctOnce(Tom, I, 3, 0.0)
    h.bindModelBuffers(None, None, None); h.setRanges(ranges)
    mp = synth.PinholeModel(64, 48, 52.5, 52.5, 31.5, 23.5, 0.3, 30.0)
    hp = rmcl_b200.RCCB200Pinhole(gmap); hp.setTsb(Tsb); hp.setModel(mp); hp.setParams(1.0, 0.15)
    hp.find(Tgt); hp.setRanges(hp.modelView()["ranges"])
    rmcl_b200.micp_correct_once([h, hp], np.stack([I, I]), Tom, 5, 0.0, merge_weights=[1.0, 0.4])
    h.benchmark(T, 2)
    Pall = torch.cat([Pd, Pd]); Aall = torch.cat([Ad, Ad])
    up.resampleP2PLocalWorld(Pall, Aall, 2, 1)
    torch.cuda.synchronize()
    del hb, hp
    g2 = rmcl_b200.Map.from_blob(gmap.export_blob())
    g2.intersect(np.zeros((5, 3), np.float32) + [30, 20, 1], np.eye(3, dtype=np.float32)[[0, 1, 2, 0, 1]])
    if mode == 1:
        gmap.refit((V + 0.01).astype(np.float32))
        h.find(Tom)
    del h, hc, up, gmap, g2
print("sanitize workload done", rmcl_b200.kernel_launch_count())
