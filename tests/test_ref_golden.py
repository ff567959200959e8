"""Oracle (and CUDA path) against outputs of the REAL reference (rmcl + rmagine + Embree), produced by oracle/ref_harness on a machine that
has them and committed as tests/golden/ref_*.npz.  Those fixtures do not exist yet (the authoring image has neither rmagine nor Embree): the
comparisons then SKIP with "parity unpinned" -- they never pass vacuously.  The pipeline itself (inputs, container format, conversion,
comparison code) is exercised on every run with the oracle standing in for the generator."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_harness"))

UNPINNED = "parity unpinned: tests/golden/ref_{}.npz not generated (needs rmagine + Embree, see oracle/ref_harness/README.md)"


def _tf(v):
    from rmcl_b200 import synth
    T = np.zeros((), synth.TRANSFORM_DTYPE)
    T["R"], T["t"] = v[:4], v[4:7]
    return T


def _model(tag, v):
    from rmcl_b200 import synth
    if tag == "c1":
        return synth.SphericalModel(float(v[0]), float(v[1]), int(v[2]), float(v[3]), float(v[4]), int(v[5]), float(v[6]), float(v[7]))
    return synth.PinholeModel(int(v[0]), int(v[1]), float(v[2]), float(v[3]), float(v[4]), float(v[5]), float(v[6]), float(v[7]))


def _mesh(tag):
    from rmcl_b200 import synth
    return synth.cube(29) if tag == "c1" else synth.building(60000)


def compare_simulate(ref_pts, ref_nrm, ref_hits, got, tol=1e-4):
    """hit flags equal except near silhouettes / range limits; points and normals of common hits within tol"""
    both = (ref_hits > 0) & (got["hits"] > 0)
    disagree = (ref_hits > 0) != (got["hits"] > 0)
    dp = np.abs(ref_pts.reshape(-1, 3)[both] - got["points"][both]).max() if both.any() else 0.0
    dn = np.abs(ref_nrm.reshape(-1, 3)[both] - got["normals"][both]).max() if both.any() else 0.0
    return float(disagree.mean()), float(dp), float(dn)


def check_against(rec, simulate, correct_once, statistics, umeyama, tag):
    """rec: the npz records; the callables run OUR side (oracle or CUDA path) on the recorded inputs"""
    m = _model(tag, rec[f"in.{tag}.model"])
    Tsb, Tbo, Tom, Tgt = (_tf(rec[f"in.{tag}.{k}"]) for k in ("Tsb", "Tbo", "Tom", "Tgt"))
    prm = rec[f"in.{tag}.params"]
    # 1. simulate at T_gt: find the A.3 setting that reproduces the reference (all three are switches on our side)
    best = None
    for opts in range(8):
        got = simulate(m, Tgt, Tsb, opts)
        miss = got["hits"] == 0
        fill_ok = True
        if miss.any():
            ref_fill = rec[f"ref.{tag}.gt.points"].reshape(-1, 3)[miss]
            fill_ok = bool(np.array_equal(np.isnan(ref_fill), np.isnan(got["points"][miss])))
        frac, dp, dn = compare_simulate(rec[f"ref.{tag}.gt.points"], rec[f"ref.{tag}.gt.normals"], rec[f"ref.{tag}.gt.hits"], got)
        score = (frac, not fill_ok, dp)
        if best is None or score < best[0]:
            best = (score, opts, frac, dp, dn, fill_ok)
    _, opts, frac, dp, dn, fill_ok = best
    assert frac <= 2e-3, f"{tag}: hit flags differ on {frac:.4%} of the rays under the best A.3 setting {opts}"
    assert dp <= 1e-4 and dn <= 1e-4 and fill_ok, (tag, opts, dp, dn, fill_ok)
    # 2. dataset
    from oracle import pyoracle as po
    o, d = po.model_rays(m)
    dpts, dmask, _ = po.dataset_from_ranges(o, d, rec[f"in.{tag}.ranges"], m.range_min, m.range_max)
    assert np.array_equal(dmask, rec[f"ref.{tag}.dataset.mask"])
    assert np.abs(dpts - rec[f"ref.{tag}.dataset.points"].reshape(-1, 3)).max() <= 1e-6
    # 3. one reduction + Umeyama on the reference's OWN model buffers (isolates statistics_p2l / umeyama from the tracer)
    s_ref = rec[f"ref.{tag}.stats0"]
    st = statistics(dpts, dmask, rec[f"ref.{tag}.guess.points"].reshape(-1, 3), rec[f"ref.{tag}.guess.normals"].reshape(-1, 3), rec[f"ref.{tag}.guess.hits"], float(prm[0]))
    n_ref = int(s_ref[15:16].view(np.uint32)[0])
    assert int(st["n_meas"]) == n_ref
    assert np.abs(st["dataset_mean"] - s_ref[0:3]).max() <= 5e-5 and np.abs(st["model_mean"] - s_ref[3:6]).max() <= 5e-5
    assert np.abs(st["covariance"] - s_ref[6:15]).max() <= 5e-3
    u = umeyama(st)
    u_ref = rec[f"ref.{tag}.umeyama0"]
    assert np.abs(u["t"] - u_ref[4:7]).max() <= 1e-4 and min(np.abs(u["R"] - u_ref[:4]).max(), np.abs(u["R"] + u_ref[:4]).max()) <= 1e-5
    # 4. the whole correctOnce: north-star tolerance 1e-5 on the pose
    Tn = correct_once(m, dpts, dmask, Tom, Tbo, Tsb, int(prm[3]), float(prm[0]), float(prm[1]), float(prm[2]), opts)
    T_ref = rec[f"ref.{tag}.Tom_new"]
    assert np.abs(Tn["t"] - T_ref[4:7]).max() <= 1e-5 and min(np.abs(Tn["R"] - T_ref[:4]).max(), np.abs(Tn["R"] + T_ref[:4]).max()) <= 1e-5
    return opts


def _oracle_side(tag):
    from oracle import pyoracle as po
    V, F = _mesh(tag)
    osc = po.Scene(V, F)

    def simulate(m, T, Tsb, opts):
        o, d = po.model_rays(m)
        return osc.simulate(T, Tsb, o, d, m.range_max, m.range_min, tfar_mode=opts & 1, min_mode=(opts >> 1) & 1, miss_fill=(opts >> 2) & 1)

    def correct_once(m, dp, dm, Tom, Tbo, Tsb, it, md, amin, cp, opts):
        o, d = po.model_rays(m)
        return osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, it, md, amin, cp, f64_accum=False)[0]

    def statistics(dp, dm, mp, mn, mh, md):
        from rmcl_b200 import synth
        return po.statistics_p2l(synth.make_transform(), dp, dm, mp, mn, mh, md, f64=False)

    return simulate, correct_once, statistics, po.umeyama


@pytest.mark.parametrize("tag", ["c1", "pin"])
def test_oracle_against_reference_outputs(po, tag):
    path = os.path.join(GOLD, f"ref_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip(UNPINNED.format(tag))
    rec = dict(np.load(path, allow_pickle=False))
    check_against(rec, *_oracle_side(tag), tag)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c1", "pin"])
def test_cuda_path_against_reference_outputs(po, tag):
    path = os.path.join(GOLD, f"ref_{tag}.npz")
    if not os.path.exists(path):
        pytest.skip(UNPINNED.format(tag))
    import rmcl_b200
    rec = dict(np.load(path, allow_pickle=False))
    V, F = _mesh(tag)
    gmap = rmcl_b200.Map(V, F)
    cls = rmcl_b200.RCCB200Spherical if tag == "c1" else rmcl_b200.RCCB200Pinhole

    def handle(m, Tsb, opts):
        h = cls(gmap)
        h.setTsb(Tsb); h.setModel(m); h.setSimOptions(opts & 1, (opts >> 1) & 1, (opts >> 2) & 1)
        return h

    def simulate(m, T, Tsb, opts):
        h = handle(m, Tsb, opts)
        h.find(T)
        return h.modelView()

    def correct_once(m, dp, dm, Tom, Tbo, Tsb, it, md, amin, cp, opts):
        h = handle(m, Tsb, opts)
        h.setParams(md, amin); h.setDataset(dp, dm)
        return h.correctOnce(Tom, Tbo, it, cp)[0]

    def statistics(dp, dm, mp, mn, mh, md):
        return _oracle_side(tag)[2](dp, dm, mp, mn, mh, md)          # the device reduction is compared with the oracle elsewhere (test_gpu_parity)

    check_against(rec, simulate, correct_once, statistics, lambda s: rmcl_b200.umeyama_transform(s[None])[0], tag)


def test_harness_pipeline_selfcheck(po, tmp_path):
    """make_ref_inputs -> (oracle standing in for gen_ref_golden) -> ref_to_npz -> check_against: the container format, both converters and the
    comparison code work end to end.  This pins nothing: it only guarantees that a future real reference run meets a working pipeline."""
    import b2ref
    import make_ref_inputs
    import ref_to_npz
    from rmcl_b200 import synth
    in_dir, out_dir, gold = str(tmp_path / "in"), str(tmp_path / "out"), str(tmp_path / "gold")
    made = make_ref_inputs.build(in_dir)
    os.makedirs(out_dir)
    assert os.path.exists(os.path.join(in_dir, "cube29.ply")) and os.path.exists(os.path.join(in_dir, "building60k.ply"))
    I = synth.make_transform()
    for tag, (V, F, m, Tgt, Tsb, Tbo, Tom, ranges, rec_in) in made.items():
        assert b2ref.read(os.path.join(in_dir, tag + ".b2ref")).keys() == rec_in.keys()
        osc = po.Scene(V, F)
        o, d = po.model_rays(m)
        gt = osc.simulate(Tgt, Tsb, o, d, m.range_max, m.range_min)
        dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
        guess = osc.simulate(synth.compose(Tom, Tbo), Tsb, o, d, m.range_max, m.range_min)
        st = po.statistics_p2l(I, dp, dm, guess["points"], guess["normals"], guess["hits"], 1.0, f64=False)
        Tn, Td, Cm = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, 0.0, f64_accum=False)

        def tf8(T):
            return np.concatenate([T["R"], T["t"], [0.0]]).astype(np.float32)

        def st16(s):
            return np.concatenate([s["dataset_mean"], s["model_mean"], s["covariance"], np.array([s["n_meas"]], np.uint32).view(np.float32)]).astype(np.float32)

        out = {f"{tag}.gt.points": gt["points"].reshape(-1), f"{tag}.gt.normals": gt["normals"].reshape(-1), f"{tag}.gt.hits": gt["hits"],
               f"{tag}.dataset.points": dp.reshape(-1), f"{tag}.dataset.mask": dm,
               f"{tag}.guess.points": guess["points"].reshape(-1), f"{tag}.guess.normals": guess["normals"].reshape(-1), f"{tag}.guess.hits": guess["hits"],
               f"{tag}.stats0": st16(st), f"{tag}.umeyama0": tf8(po.umeyama(st)), f"{tag}.Tom_new": tf8(Tn), f"{tag}.T_onew_oold": tf8(Td), f"{tag}.Cmerged_o": st16(Cm)}
        b2ref.write(os.path.join(out_dir, tag + ".b2ref"), out)
    assert ref_to_npz.convert(in_dir, out_dir, gold, "oracle standing in (self-check)") == ["c1", "pin"]
    for tag in ("c1", "pin"):
        rec = dict(np.load(os.path.join(gold, f"ref_{tag}.npz"), allow_pickle=False))
        assert check_against(rec, *_oracle_side(tag), tag) == 0          # the stand-in used the default A.3 setting: it must be the one recovered
