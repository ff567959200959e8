"""Experiment: the particle filter's ray mapping.  Shipped: consecutive lanes = consecutive BEAMS of one particle.  Experiment build
(-DB2_PF_LANE_PARTICLE, 32 particles per block; select it with B2_LIB_PATH): consecutive lanes = the 32 PARTICLES of a block, same beam --
coherent only if neighbouring particles are close in pose, hence the particle orders: random, sorted by (heading bin, Morton cell).
Uniform particles = global localisation (the C3 workload); converged cloud = tracking."""
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
h.find(Tgt)
beams = synth.pf_beams(h.modelView()["points"], 180)
N = 100_000
P, A = synth.pf_particles(N)
up = rmcl_b200.PCDSensorUpdaterB200(gmap)
stream = torch.cuda.current_stream()
up.setStream(stream.cuda_stream)
prm = rmcl_b200.PFParams.defaults()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(Px, Ax, reps=4):
    Pd = torch.from_numpy(Px.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(Ax.view(np.float32).reshape(-1, 9).copy()).cuda()
    ts, out = [], None
    for i in range(reps + 1):
        flush.fill_(i)
        Ai = Ad.clone()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); up.update(Pd, Ai, Tsb, beams, prm); b.record(stream)
        torch.cuda.synchronize()
        if i:
            ts.append(a.elapsed_time(b))
        out = Ai
    return float(np.median(ts)), out.cpu().numpy()


def morton2(x, y):
    def spread(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


def order_sorted(Px, yaw_bins):
    yaw = 2.0 * np.arctan2(Px["R"][:, 2], Px["R"][:, 3])
    yb = np.floor((yaw + np.pi) / (2 * np.pi) * yaw_bins).astype(np.int64) % yaw_bins
    cell = morton2(np.clip(Px["t"][:, 0] / 60.0 * 1023, 0, 1023).astype(np.int64), np.clip(Px["t"][:, 1] / 40.0 * 1023, 0, 1023).astype(np.int64))
    return np.lexsort((cell, yb))


print("library:", os.environ.get("B2_LIB_PATH", "shipped"))
t0, ref = timed(P, A)
print("uniform particles, random order:                              %.3f ms = %.2f G rays/s" % (t0, N * 180 / t0 / 1e6))
for yb in (1, 16, 64, 256):
    o = order_sorted(P, yb)
    t, out = timed(P[o], A[o])
    inv = np.empty_like(o); inv[o] = np.arange(N)
    same = out[inv].tobytes() == ref.tobytes()
    print("uniform particles, sorted by (%3d heading bins, Morton cell):  %.3f ms = %.2f G rays/s   results equal: %s" % (yb, t, N * 180 / t / 1e6, same))
rng = np.random.default_rng(7)
Pc = P.copy()
Pc["t"][:, 0] = Tgt["t"][0] + rng.normal(0, 0.3, N); Pc["t"][:, 1] = Tgt["t"][1] + rng.normal(0, 0.3, N); Pc["t"][:, 2] = Tgt["t"][2]
yawc = rng.normal(0.0, np.radians(5.0), N)
Pc["R"][:, 2] = np.sin(yawc / 2); Pc["R"][:, 3] = np.cos(yawc / 2)
t, _ = timed(Pc, A)
print("converged cloud (sigma 0.3 m, 5 deg), random order:           %.3f ms = %.2f G rays/s" % (t, N * 180 / t / 1e6))
oc = order_sorted(Pc, 64)
t, _ = timed(Pc[oc], A[oc])
print("converged cloud, sorted:                                      %.3f ms = %.2f G rays/s" % (t, N * 180 / t / 1e6))
