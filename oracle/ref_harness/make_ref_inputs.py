"""Scenario inputs for gen_ref_golden (the reference-side generator) AND for the oracle: meshes as ascii PLY, poses / models / scans in
.b2ref containers.  Deterministic (rmcl_b200.synth, fixed seeds).  python oracle/ref_harness/make_ref_inputs.py <out dir>

The scans are produced by the ORACLE at the ground-truth pose (+ noise); the reference-side generator re-simulates at the same pose with
Embree ("<tag>.gt.*"), which is the first thing tests/test_ref_golden.py compares."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import b2ref  # noqa: E402
from rmcl_b200 import synth  # noqa: E402


def write_ply(path, V, F):
    with open(path, "w") as f:
        f.write(f"ply\nformat ascii 1.0\nelement vertex {len(V)}\nproperty float x\nproperty float y\nproperty float z\nelement face {len(F)}\n"
                "property list uchar int vertex_indices\nend_header\n")
        for v in V:
            f.write(f"{v[0]!r} {v[1]!r} {v[2]!r}\n")
        for t in F:
            f.write(f"3 {t[0]} {t[1]} {t[2]}\n")


def tf8(T):
    return np.concatenate([np.asarray(T["R"], np.float32), np.asarray(T["t"], np.float32), [0.0]]).astype(np.float32)


def scenarios():
    """name -> (mesh name, (V, F), model, Tgt, dict of extras)"""
    from oracle import pyoracle as po
    out = {}
    # C1 (BASELINE.json configs[0])
    V, F = synth.cube(29)
    m = synth.c1_sensor()
    Tsb = synth.scenario_tsb()
    Tgt = synth.make_transform((0.5, -0.3, 0.2), (0.02, -0.01, 0.3))
    out["c1"] = ("cube29", (V, F), m, Tgt, Tsb, 7)
    # pinhole in the small building
    V2, F2 = synth.building(60000)
    m2 = synth.PinholeModel(160, 120, 131.25, 131.25, 79.5, 59.5, 0.3, 30.0)
    out["pin"] = ("building60k", (V2, F2), m2, synth.building_gt_pose(), synth.make_transform((0.3, 0.1, 0.4), (0.0, 0.1, -0.4)), 8)
    return out, po


def build(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    sc, po = scenarios()
    made = {}
    for tag, (mesh_name, (V, F), m, Tgt, Tsb, seed) in sc.items():
        write_ply(os.path.join(out_dir, mesh_name + ".ply"), np.asarray(V, np.float32).tolist(), np.asarray(F).tolist())
        osc = po.Scene(V, F)
        o, d = po.model_rays(m)
        ranges = synth.noisy_ranges(osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"], m.range_max, seed=seed)
        Tbo = synth.make_transform((0.05, 0.02, 0.0), (0, 0, 0.1))
        Tom = synth.compose(synth.compose(Tgt, synth.scenario_pose_offset()), synth.inverse(Tbo))
        if type(m).__name__ == "SphericalModel":
            model = [m.phi_min, m.phi_inc, m.phi_size, m.theta_min, m.theta_inc, m.theta_size, m.range_min, m.range_max]
        else:
            model = [m.width, m.height, m.fx, m.fy, m.cx, m.cy, m.range_min, m.range_max]
        rec = {f"{tag}.model": np.float32(model), f"{tag}.Tsb": tf8(Tsb), f"{tag}.Tbo": tf8(Tbo), f"{tag}.Tom": tf8(Tom), f"{tag}.Tgt": tf8(Tgt),
               f"{tag}.ranges": ranges.astype(np.float32), f"{tag}.params": np.float32([1.0, 0.15, 0.0, 5])}
        if tag == "c1":
            rng = np.random.default_rng(11)
            ro = rng.uniform(-9.5, 9.5, (4096, 3)).astype(np.float32)
            rd = rng.normal(size=(4096, 3))
            rd = (rd / np.linalg.norm(rd, axis=1, keepdims=True)).astype(np.float32)
            rec["c1rays.origs"], rec["c1rays.dirs"] = ro.reshape(-1), rd.reshape(-1)
        b2ref.write(os.path.join(out_dir, tag + ".b2ref"), rec)
        made[tag] = (V, F, m, Tgt, Tsb, Tbo, Tom, ranges, rec)
    return made


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "inputs"))
    print("inputs written")
