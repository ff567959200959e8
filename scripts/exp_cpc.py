"""Timing of the closest-point correspondence path on the C2 scan: CUDA-event times of k_cpc_find and of the ICP loop."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import synth


def main():
    V, F = synth.building(1_000_000)
    mp = rmcl_b200.Map(V, F)
    m = synth.c2_sensor()
    Tgt, Tsb = synth.building_gt_pose(), synth.scenario_tsb()
    rcc = rmcl_b200.RCCB200Spherical(mp); rcc.setTsb(Tsb); rcc.setModel(m); rcc.setParams(1.0, 0.15)
    rcc.find(Tgt)
    clean = rcc.modelView()["ranges"]
    ranges = synth.noisy_ranges(clean, m.range_max)
    rcc.setRanges(ranges)
    ds = rcc.datasetView()
    h = rmcl_b200.CPCB200(mp); h.setTsb(Tsb); h.setParams(1.0, 0.15); h.setDataset(ds["points"], ds["mask"])
    Tbo = synth.make_transform((0.05, 0.02, 0.0), (0, 0, 0.1))
    Tom = synth.compose(synth.compose(Tgt, synth.scenario_pose_offset()), synth.inverse(Tbo))
    h.enableTiming(True); rcc.enableTiming(True)
    out = {}
    for tag, obj in (("cpc", h), ("rcc", rcc)):
        ts = []
        for i in range(30):
            obj.correctOnce(Tom, Tbo, 5, 0.0)
            ts.append(obj.lastTiming())
        ts = np.array(ts[5:])
        out[tag] = dict(find_us=float(np.median(ts[:, 0]) * 1e3), loop_us=float(np.median(ts[:, 1]) * 1e3), n=int(len(ds["mask"])))
    out["cpc"]["M_queries_per_s"] = out["cpc"]["n"] / out["cpc"]["find_us"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
