"""Experiment: how long do the heaviest tiles of the C2 scan take when they run alone (uncrowded SMs), and the rest without them?  Input for a
cost-aware tile schedule of k_rcc_find.  Tiles = the kernel's 8x4 raster tiles (one per warp); cost = the warp's duration in the normal launch."""
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
Tbm = synth.compose(Tgt, synth.scenario_pose_offset())
h = rmcl_b200.RCCB200Spherical(gmap)
h.setTsb(Tsb); h.setModel(m); h.setParams(1.0, 0.15)
lib = rmcl_b200.load_library()
W, H = m.width, m.height
nw = W * H // 32
buf = torch.zeros(2 * nw, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
for _ in range(4):
    h.find(Tbm)
torch.cuda.synchronize()
lib.b2_rcc_debug_find_warp_times(h._h, C.c_void_p(buf.data_ptr()))
h.find(Tbm); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 2).astype(np.float64)
dur = (t[:, 1] - t[:, 0]) / 1e3
print("normal launch (warm): span %.1f us, warp duration median %.1f p90 %.1f p99 %.1f max %.1f" % ((t[:, 1].max() - t[:, 0].min()) / 1e3, np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max()))
lib.b2_rcc_debug_find_warp_times(h._h, None)

# directions of the spherical model in buffer order (vid * W + hid), rmagine's getDirection
vid, hid = np.divmod(np.arange(W * H), W)
phi, theta = m.phi_min + vid * m.phi_inc, m.theta_min + hid * m.theta_inc
dirs = np.stack([np.cos(phi) * np.cos(theta), np.cos(phi) * np.sin(theta), np.sin(phi)], 1).astype(np.float32)
tpr = W // 8


def tile_rays(tile):            # buffer ids of the 32 rays of a tile, in lane order
    within = np.arange(32)
    return ((tile // tpr) * 4 + (within >> 3)) * W + (tile % tpr) * 8 + (within & 7)


def timed_find(tiles, label):
    ids = np.concatenate([tile_rays(tl) for tl in tiles])
    # new raster: width 8, height 4 * len(tiles); tile k of the new model = rows 4k..4k+3 = the 32 rays of tiles[k] in lane order
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
timed_find(np.arange(nw), "all tiles, raster order (= the spherical launch)")
timed_find(order, "all tiles, heaviest first")
timed_find(order[::-1], "all tiles, lightest first")
for K in (16, 74, 148, 296, 592, 1184):
    timed_find(order[:K], "heaviest %d alone" % K)
    timed_find(order[K:], "all but the heaviest %d" % K)
