"""One C2 correctOnce under the microscope: SM-clock cost of the phases of an ICP-loop iteration, %globaltimer timeline of find and loop, and the
host wall clock of the call for the three ways the scan can arrive (resident, pageable host, pinned host)."""
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
ranges = synth.noisy_ranges(h.modelView()["ranges"], m.range_max)
pinned = torch.from_numpy(ranges.copy()).pin_memory()
h.setRanges(ranges)
Tom = synth.compose(Tgt, synth.scenario_pose_offset())
lib = rmcl_b200.load_library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
nw = (m.size + 31) // 32
buf = torch.zeros(2 * nw, dtype=torch.int64, device="cuda")
for _ in range(5):
    h.correctOnce(Tom, I, 5, 0.0)
out = (C.c_ulonglong * 16)()
lib.b2_rcc_debug_clocks(h._h, out)
print("SM cycles, iteration 1, block 0: pass+block reduce+publish %d | collect all blocks' partials %d | per-sensor statistics (15 lanes) %d | rest of the tail %d || pair loads %d | whole kernel %d" % tuple(out[i] for i in range(6)))
if any(out[8 + q] for q in range(8)):        # library built with `make PROFILE=1`
    print("tail stamps (cycles after the statistics): merge %d | Newton polar %d | FP64 polish %d | quat+t %d | T update %d | Tpre %d"
          % tuple(out[8 + q] for q in range(2, 8)))
if hasattr(lib, "b2_rcc_debug_blocks") and any(out[8 + q] for q in range(8)):
    blk = (C.c_ulonglong * 640)()
    lib.b2_rcc_debug_blocks(h._h, blk)
    b = np.array(list(blk), dtype=np.int64).reshape(160, 4)[:148]
    pub, col = b[:, 0] - b[:, 0].min(), b[:, 1] - b[:, 0].min()
    print("iteration 1, all blocks: publish at %d..%d ns (median %d) | collected at %d..%d ns (median %d) | cycles reduce->publish median %d | publish->collected median %d, min %d, max %d"
          % (pub.min(), pub.max(), np.median(pub), col.min(), col.max(), np.median(col), np.median(b[:, 2]), np.median(b[:, 3] - b[:, 2]), (b[:, 3] - b[:, 2]).min(), (b[:, 3] - b[:, 2]).max()))
    print("   last publisher: block %d; blocks sorted by publish time (ns): %s" % (int(pub.argmax()), np.sort(pub)[::12].tolist()))
lib.b2_rcc_debug_find_warp_times(h._h, C.c_void_p(buf.data_ptr()))
for name, src in (("resident", None), ("pageable", ranges), ("pinned", pinned)):
    for k in range(6):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.correctOnce(Tom, I, 5, 0.0, ranges=src)
        wall = (time.perf_counter() - t0) * 1e6
        lib.b2_rcc_debug_clocks(h._h, out)
        t = buf.cpu().numpy().reshape(-1, 2)
        f0, f1 = int(t[:, 0].min()), int(t[:, 1].max())
        if k >= 3:
            print("%-9s find: last warp end %.1f us | loop block 0: start %.1f, end %.1f us (after find's end: %.1f) | host wall clock %.1f us"
                  % (name, (f1 - f0) / 1e3, (int(out[6]) - f0) / 1e3, (int(out[7]) - f0) / 1e3, (int(out[7]) - f1) / 1e3, wall))
    print("   %-9s cycles: pass %d | collect %d | statistics %d | rest of tail %d || prologue %d | whole kernel %d" % ((name,) + tuple(out[i] for i in range(6))))
    if hasattr(lib, "b2_rcc_debug_blocks"):
        blk = (C.c_ulonglong * 640)()
        lib.b2_rcc_debug_blocks(h._h, blk)
        print("   block 0, iteration 0 (first phase of the tile sort on warps 1..15), cycles from the block reduce to the end barrier: warp 0 %d, slowest other warp (max over calls) %d"
              % (blk[625], blk[626]))
lib.b2_rcc_debug_find_warp_times(h._h, None)
for mode in (2, 1, 0):
    h.setExecMode(mode)
    ts = []
    for k in range(30):
        flush.fill_(1); torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.correctOnce(Tom, I, 5, 0.0)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("exec mode %d: host wall clock of the resident call, median %.1f us" % (mode, float(np.median(ts[5:]))))
