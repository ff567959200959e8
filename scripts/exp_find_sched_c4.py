"""Experiment: would the tile schedule pay on C4 (640x480 depth image, 9600 tiles = 2.3 waves of k_rcc_find)?  The tiles of the pinhole launch
are replayed as an O1Dn model in raster order / slowest first / fastest first (directions = normalised hit points of the pinhole launch)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth

V, F = synth.indoor(500_000)
gmap = rmcl_b200.Map(V, F)
m = synth.c4_sensor()
Tsb, Tgt = synth.scenario_tsb(), synth.indoor_gt_pose()
Tbm = synth.compose(Tgt, synth.scenario_pose_offset())
h = rmcl_b200.RCCB200Pinhole(gmap)
h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
lib = rmcl_b200.load_library()
W, H = m.width, m.height
nw = W * H // 32
buf = torch.zeros(2 * nw, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for _ in range(3):
    h.find(Tbm)
torch.cuda.synchronize()
lib.b2_rcc_debug_find_warp_times(h._h, C.c_void_p(buf.data_ptr()))
h.find(Tbm); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 2).astype(np.float64)
dur = (t[:, 1] - t[:, 0]) / 1e3
print("pinhole launch (warm): span %.1f us, warp duration median %.1f p90 %.1f p99 %.1f max %.1f; last warp starts at %.1f us"
      % ((t[:, 1].max() - t[:, 0].min()) / 1e3, np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), (t[:, 0].max() - t[:, 0].min()) / 1e3))
lib.b2_rcc_debug_find_warp_times(h._h, None)
mv = h.modelView()
pts = mv["points"].reshape(-1, 3).astype(np.float64)
ok = np.isfinite(pts).all(1)
dirs = np.zeros_like(pts); dirs[:, 0] = 1.0
dirs[ok] = pts[ok] / np.linalg.norm(pts[ok], axis=1, keepdims=True)
dirs = dirs.astype(np.float32)
print("rays with a hit: %d of %d" % (ok.sum(), len(ok)))
tpr = W // 8


def tile_rays(tile):
    within = np.arange(32)
    return ((tile // tpr) * 4 + (within >> 3)) * W + (tile % tpr) * 8 + (within & 7)


def timed_find(tiles, label):
    ids = np.concatenate([tile_rays(tl) for tl in tiles])
    mo = synth.O1DnModel(8, 4 * len(tiles), np.zeros(3, np.float32), dirs[ids].copy(), m.range_min, m.range_max)
    ho = rmcl_b200.RCCB200O1Dn(gmap)
    ho.setTsb(Tsb); ho.setModel(mo); ho.setParams(1.0, 0.15)
    ho.setStream(stream.cuda_stream)
    out = []
    for cold in (True, False):
        ts = []
        for i in range(12):
            if cold:
                flush.fill_(i & 255)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ho.find(Tbm); b.record(stream)
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b) * 1e3)
        out.append(np.median(ts))
    print("%-44s %5d tiles: cold %.1f us, warm %.1f us" % (label, len(tiles), out[0], out[1]))


order = np.argsort(-dur)
timed_find(np.arange(nw), "all tiles, raster order")
timed_find(order, "all tiles, slowest first")
timed_find(order[::-1], "all tiles, fastest first")
cls = np.minimum(15, (dur * 16.0 / (dur.max() * 1.0001)).astype(int))
timed_find(np.argsort(-cls, kind="stable"), "all tiles, 16 duration classes, slowest class first")
