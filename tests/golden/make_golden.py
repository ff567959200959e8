"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.c) plus closed-form values.

The reference ships no tests, fixtures or known-answer vectors for this path and cannot be built here, so these
vectors pin OUR restatement: (a) closed-form geometry that any correct closest-hit must reproduce (axis-aligned cube from inside),
(b) regression values of the oracle itself so later edits cannot silently change the parity target.
Run:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po          # noqa: E402
from rmcl_b200 import synth                # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    # ---- C1: 32x32 spherical sensor inside the 10 092-triangle cube (plumbing config) ----
    V, F = synth.cube(29)
    sc = po.Scene(V, F)
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    Tsb = synth.scenario_tsb()
    Tgt = synth.make_transform((0.5, -0.3, 0.2), (0.02, -0.01, 0.3))
    sim = sc.simulate(Tgt, Tsb, o, d, m.range_max)
    # closed form: sensor origin in the map frame + rotated directions against the 6 planes |x|,|y|,|z| = 10
    Tsm = synth.compose(Tgt, Tsb)
    q = np.asarray(Tsm["R"], np.float64)
    dm = synth._qrot(q[None, :], d.astype(np.float64))
    om = np.asarray(Tsm["t"], np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        tpos = (10.0 - om[None, :]) / dm
        tneg = (-10.0 - om[None, :]) / dm
    tt = np.where(dm > 0, tpos, tneg)
    analytic = tt.min(1)
    axis = tt.argmin(1)
    normal_m = np.zeros_like(dm)
    normal_m[np.arange(len(dm)), axis] = -np.sign(dm[np.arange(len(dm)), axis])
    normal_s = synth._qrot((q * np.array([-1, -1, -1, 1.0]))[None, :], normal_m)
    rng = synth.noisy_ranges(sim["ranges"], m.range_max, seed=7)
    dp, dmask, nv = po.dataset_from_ranges(o, d, rng, m.range_min, m.range_max)
    Tguess = synth.compose(Tgt, synth.scenario_pose_offset())
    model_at_guess = sc.simulate(Tguess, Tsb, o, d, m.range_max)
    I = synth.make_transform()
    st32 = po.statistics_p2l(I, dp, dmask, model_at_guess["points"], model_at_guess["normals"], model_at_guess["hits"], 1.0, f64=False)
    st64 = po.statistics_p2l(I, dp, dmask, model_at_guess["points"], model_at_guess["normals"], model_at_guess["hits"], 1.0, f64=True)
    Tn, Td, Cm = sc.micp_correct_once(o, d, m.range_max, dp, dmask, Tguess, I, Tsb, iterations=5, max_dist=1.0, adaptive_max_dist_min=0.15, f64_accum=True)
    np.savez_compressed(os.path.join(OUT, "c1_cube.npz"), Tgt=Tgt, Tsb=Tsb, Tguess=Tguess, dirs=d,
                        ranges=sim["ranges"], hits=sim["hits"], face_ids=sim["face_ids"], points=sim["points"], normals=sim["normals"],
                        analytic_ranges=analytic.astype(np.float64), analytic_normals=normal_s.astype(np.float64),
                        noisy_ranges=rng, dataset_points=dp, dataset_mask=dmask,
                        guess_face_ids=model_at_guess["face_ids"], guess_ranges=model_at_guess["ranges"],
                        stats_f32=st32, stats_f64=st64, umeyama_f64=po.umeyama(st64), Tom_new=Tn, T_onew_oold=Td, Cmerged=Cm)

    # ---- Umeyama known answers: stats built from an exact rigid motion (incl. a near-planar / reflection-prone set) ----
    rs = np.random.default_rng(11)
    cases = []
    for k in range(6):
        P = rs.normal(size=(200, 3)) * np.array([3.0, 2.0, 0.01 if k >= 4 else 1.5])
        rpy = rs.uniform(-0.5, 0.5, 3)
        t = rs.uniform(-2, 2, 3)
        qd = synth.quat_from_rpy(*rpy)
        Q = synth._qrot(qd[None, :], P) + t
        dmn, mmn = P.mean(0), Q.mean(0)
        Cc = (Q - mmn).T @ (P - dmn) / len(P)            # C[r,c] = mean (m-mbar)_r (d-dbar)_c
        s = np.zeros((), po.CROSS_STATS)
        s["dataset_mean"], s["model_mean"], s["n_meas"] = dmn, mmn, len(P)
        s["covariance"] = Cc.T.reshape(-1)               # column-major
        T = po.umeyama(s)
        cases.append((s, T, np.concatenate([qd, t])))
    np.savez_compressed(os.path.join(OUT, "umeyama.npz"), stats=np.array([c[0] for c in cases]), T=np.array([c[1] for c in cases]),
                        truth=np.array([c[2] for c in cases]))

    # ---- particle filter: 64 particles x 24 beams inside the cube ----
    beams = synth.pf_beams(sim["points"], 24, seed=3)
    P, A = synth.pf_particles(64, footprint=(16.0, 16.0), z=0.0, seed=5, margin=0.0)
    P["t"][:, :2] -= 8.0
    out = {}
    for ng_mode in (0, 1):
        out[f"attrs_ng{ng_mode}"] = sc.pf_update(P, A, Tsb, beams, po.PFParams.defaults(ng_mode))
    np.savez_compressed(os.path.join(OUT, "pf_cube.npz"), poses=P, attrs0=A, beams=beams, Tsb=Tsb, **out)
    widened(sc, sim, o, d, m, Tsb, Tgt, P, A)
    print("golden vectors written to", OUT)


def widened(sc, sim, o, d, m, Tsb, Tgt, P, A):
    """Regression vectors for the SURVEY 8(f) rows: closest-point find, PF with the closest-point error, motion update with the wall check,
    Gladiator resampling (Philox draws), scan-vs-map segmentation.  Same cube, small sizes."""
    rs = np.random.default_rng(17)
    q = rs.uniform(-11, 11, (256, 3)).astype(np.float32)
    cpc = sc.cpc_find(Tgt, Tsb, q, 0.8)
    beams = synth.pf_beams(sim["points"], 24, seed=3)
    A1 = sc.pf_update(P, A, Tsb, beams, po.PFParams.defaults(0, 1))
    A1["likelihood"]["n_meas"] = rs.integers(1, 9000, len(A1)).astype(np.uint32)
    Tmo = synth.make_transform((1.5, 0.0, 0.0), (0, 0, 0.05))
    Pm, Am = po.pf_motion_update(P, A1, Tmo, 0.03, scene=sc)
    cfg = po.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    raw, nrm = po.pf_gladiator_randoms(1234, 3, 0, len(Pm))
    Pr, Ar = po.pf_gladiator_resample(Pm, Am, 0, len(Pm), raw, nrm, cfg)
    real = sim["ranges"].copy()
    k = rs.permutation(len(real))
    real[k[:40]] *= 0.6; real[k[40:80]] *= 1.3; real[k[80:90]] = m.range_max + 1
    seg_scan, seg_map, seg_lab = po.segment(o, d, m.range_min, m.range_max, real, sim["ranges"], sim["normals"], 0.15, 0.1)
    np.savez_compressed(os.path.join(OUT, "f_rows.npz"), Tgt=Tgt, Tsb=Tsb, queries=q, cpc_points=cpc["points"], cpc_normals=cpc["normals"], cpc_hits=cpc["hits"],
                        cpc_faces=cpc["face_ids"], cpc_dists=cpc["dists"], poses=P, attrs0=A, beams=beams, attrs_cpc=A1, T_motion=Tmo, poses_moved=Pm,
                        attrs_moved=Am, glad_raw=raw, glad_normals=nrm, poses_resampled=Pr, attrs_resampled=Ar, real_ranges=real,
                        seg_scan=seg_scan, seg_map=seg_map, seg_labels=seg_lab)


if __name__ == "__main__":
    main()
