"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI (librmcl_b200.so) via rmcl_b200.api;
the oracle (oracle/) is only the checker.  Bar: bit-exact for hit flags / face ids / n_meas (and, by construction of the shared
hit definition, for ranges / points / normals); stated float tolerances for reduced quantities."""
import os

import numpy as np
import pytest

from conftest import gpu_map, mesh, oracle_scene, quat_close, random_rays

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

TOL_DT = 1e-5          # north-star tolerance on the pose update (metres / quaternion components)
TOL_STATS = 2e-6       # CrossStatistics means (FP64 sum form on both sides, different summation order)
TOL_LIK = 2e-6         # PF likelihood (device exp() vs glibc exp() can differ in the last double bit before rounding to float)


def _rcc(synth, name, model, Tsb=None, cls=None):
    import rmcl_b200
    h = (cls or rmcl_b200.RCCB200Spherical)(gpu_map(name))
    h.setTsb(synth.scenario_tsb() if Tsb is None else Tsb)
    h.setModel(model)
    h.setParams(1.0, 0.15)
    return h


def test_library_is_the_cuda_one():
    import rmcl_b200
    lib = rmcl_b200.load_library()
    assert lib.b2_version() >= 100
    before = rmcl_b200.kernel_launch_count()
    gpu_map("cube29").intersect([[0, 0, 0]], [[1, 0, 0]])
    assert rmcl_b200.kernel_launch_count() > before


@pytest.mark.parametrize("name,lo,hi,n", [("cube29", -9.5, 9.5, 50000), ("uvsphere:40:60", -6.0, 6.0, 50000),
                                          ("building:200000", 1.0, 2.9, 200000), ("indoor:20000", 0.2, 2.8, 50000)])
def test_closest_hit_bit_exact(po, name, lo, hi, n):
    o, d = random_rays(n, lo, hi, seed=4)
    t1, f1, n1, h1 = oracle_scene(name).intersect(o, d)
    t2, f2, n2, h2 = gpu_map(name).intersect(o, d)
    assert np.array_equal(h1, h2) and np.array_equal(f1, f2)
    assert np.array_equal(t1, t2) and np.array_equal(n1, n2)
    # axis-aligned rays and finite tfar
    rng = np.random.default_rng(5)
    ax = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 5000)] * rng.choice([-1.0, 1.0], (5000, 1)).astype(np.float32)
    a = oracle_scene(name).intersect(o[:5000], ax, tfar=3.0)
    b = gpu_map(name).intersect(o[:5000], ax, tfar=3.0)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


@pytest.mark.parametrize("name,lo,hi", [("cube29", -9.5, 9.5), ("building:200000", 1.0, 2.9)])
def test_host_sah_build_bit_exact(po, name, lo, hi):
    """the alternative host SAH builder (build_mode 0) gives the same answers as the default device build and the oracle"""
    import rmcl_b200
    V, F = mesh(name)
    sah = rmcl_b200.Map(V, F, device=0, build_mode=rmcl_b200.api.B2_BUILD_HOST_SAH)
    o, d = random_rays(50000, lo, hi, seed=11)
    t1, f1, n1, h1 = oracle_scene(name).intersect(o, d)
    t2, f2, n2, h2 = sah.intersect(o, d)
    t3, f3, n3, h3 = gpu_map(name).intersect(o, d)
    assert np.array_equal(f1, f2) and np.array_equal(t1, t2) and np.array_equal(n1, n2) and np.array_equal(h1, h2)
    assert np.array_equal(f1, f3) and np.array_equal(t1, t3)


def test_c1_find_golden(po, synth):
    """C1: 32x32 spherical on the 10 092-triangle cube, against the committed golden vectors."""
    g = np.load(os.path.join(GOLD, "c1_cube.npz"))
    h = _rcc(synth, "cube29", synth.c1_sensor(), Tsb=g["Tsb"])
    h.find(g["Tgt"])
    mv = h.modelView()
    for k in ("ranges", "hits", "face_ids", "points", "normals"):
        assert np.array_equal(mv[k], g[k], equal_nan=True), k
    hit = mv["hits"] > 0
    assert np.abs(mv["ranges"][hit] - g["analytic_ranges"][hit]).max() < 2e-5
    # P2L + Umeyama against golden statistics
    h.setDataset(g["dataset_points"], g["dataset_mask"])
    h.find(g["Tguess"])
    st = h.computeCrossStatistics(synth.make_transform())
    assert st["n_meas"] == g["stats_f64"]["n_meas"]
    assert np.abs(st["dataset_mean"] - g["stats_f64"]["dataset_mean"]).max() <= TOL_STATS
    assert np.abs(st["model_mean"] - g["stats_f64"]["model_mean"]).max() <= TOL_STATS
    assert np.abs(st["covariance"] - g["stats_f64"]["covariance"]).max() <= 2e-5
    import rmcl_b200
    T = rmcl_b200.umeyama_transform(g["stats_f64"][None])[0]
    assert np.abs(T["t"] - g["umeyama_f64"]["t"]).max() <= 1e-6 and quat_close(T["R"], g["umeyama_f64"]["R"], 1e-6)
    # whole correctOnce
    Tn, Td, Cm = h.correctOnce(g["Tguess"], synth.make_transform(), 5, 0.0)
    assert Cm["n_meas"] == g["Cmerged"]["n_meas"]
    assert np.abs(Tn["t"] - g["Tom_new"]["t"]).max() <= TOL_DT and quat_close(Tn["R"], g["Tom_new"]["R"], TOL_DT)


def test_umeyama_golden(synth):
    import rmcl_b200
    g = np.load(os.path.join(GOLD, "umeyama.npz"))
    T = rmcl_b200.umeyama_transform(g["stats"])
    for out, ref, truth in zip(T, g["T"], g["truth"]):
        assert quat_close(out["R"], ref["R"], 1e-6) and np.abs(out["t"] - ref["t"]).max() <= 1e-6
        assert quat_close(out["R"], truth[:4], 2e-6)
    z = np.zeros(1, synth.CROSS_STATS_DTYPE)
    I = rmcl_b200.umeyama_transform(z)[0]
    assert np.allclose(I["R"], [0, 0, 0, 1]) and np.allclose(I["t"], 0)


@pytest.mark.parametrize("case", ["C2", "C4", "o1dn", "ondn"])
def test_find_all_models(po, synth, case):
    import rmcl_b200
    rng = np.random.default_rng(1)
    if case == "C2":       # 128 x 1024 spherical on the 1M-triangle building
        name, m, Tbm, cls = "building:1000000", synth.c2_sensor(), synth.building_gt_pose(), rmcl_b200.RCCB200Spherical
    elif case == "C4":     # 640 x 480 pinhole on the 500k-triangle indoor scene
        name, m, Tbm, cls = "indoor:500000", synth.c4_sensor(), synth.indoor_gt_pose(), rmcl_b200.RCCB200Pinhole
    else:
        dirs = rng.normal(size=(4000, 3)).astype(np.float32)
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        name, Tbm = "indoor:20000", synth.indoor_gt_pose()
        if case == "o1dn":
            m, cls = synth.O1DnModel(400, 10, np.array([0.1, 0.0, 0.05], np.float32), dirs, 0.1, 20.0), rmcl_b200.RCCB200O1Dn
        else:
            m, cls = synth.OnDnModel(400, 10, rng.uniform(-0.2, 0.2, (4000, 3)).astype(np.float32), dirs, 0.1, 20.0), rmcl_b200.RCCB200OnDn
    o, d = po.model_rays(m)
    ref = oracle_scene(name).simulate(Tbm, synth.scenario_tsb(), o, d, m.range_max)
    h = _rcc(synth, name, m, cls=cls)
    h.find(Tbm)
    mv = h.modelView()
    assert np.array_equal(mv["hits"], ref["hits"]) and np.array_equal(mv["face_ids"], ref["face_ids"])       # bit-exact
    for k in ("ranges", "points", "normals"):
        assert np.array_equal(mv[k], ref[k], equal_nan=True), k
    assert ref["hits"].mean() > 0.3
    # idempotence: a second find at the same pose rewrites identical buffers
    h.find(Tbm)
    mv2 = h.modelView()
    assert all(np.array_equal(mv[k], mv2[k], equal_nan=True) for k in mv)


def test_c2_micp_correct_once(po, synth):
    """C2 end to end: scan at T_gt (+noise, 2 % dropped), pose guess offset, one correctOnce (5 inner iterations)."""
    name, m = "building:1000000", synth.c2_sensor()
    osc = oracle_scene(name)
    o, d = po.model_rays(m)
    Tgt, Tsb = synth.building_gt_pose(), synth.scenario_tsb()
    ranges = synth.noisy_ranges(osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"], m.range_max)
    dp, dm, nvalid = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    Tbo = synth.make_transform((0.05, 0.02, 0.0), (0, 0, 0.1))
    Tom = synth.compose(synth.compose(Tgt, synth.scenario_pose_offset()), synth.inverse(Tbo))
    h = _rcc(synth, name, m)
    h.setRanges(ranges)                                           # unpackMessage on the device
    ds = h.datasetView()
    assert np.array_equal(ds["points"], dp) and np.array_equal(ds["mask"], dm) and int(dm.sum()) == nvalid
    for cp in (0.0, 0.6):
        ref64 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, cp, f64_accum=True)
        ref32 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, cp, f64_accum=False)
        import torch
        pinned = torch.from_numpy(ranges.copy()).pin_memory()          # pinned host scan: read by the kernel directly (zero copy), unpacked on the device
        for via_ranges in (False, True, "pinned"):
            if via_ranges == "pinned":
                h.setRanges(np.full_like(ranges, 3.0))                    # scramble the resident dataset: the call must rebuild it from the pinned scan
            Tn, Td, Cm = h.correctOnce(Tom, Tbo, 5, cp, ranges=None if via_ranges is False else (pinned if via_ranges == "pinned" else ranges))
            if via_ranges == "pinned":
                ds2 = h.datasetView()
                assert np.array_equal(ds2["points"], dp) and np.array_equal(ds2["mask"], dm)
            assert abs(int(Cm["n_meas"]) - int(ref64[2]["n_meas"])) <= 2                      # gate decisions (exact unless an ulp flips one)
            assert np.abs(Tn["t"] - ref64[0]["t"]).max() <= TOL_DT and quat_close(Tn["R"], ref64[0]["R"], TOL_DT)
            assert np.abs(Td["t"] - ref64[1]["t"]).max() <= TOL_DT and quat_close(Td["R"], ref64[1]["R"], TOL_DT)
            # against the reference's own FP32 sequential accumulation the gap is that arithmetic's noise floor (documented in DESIGN.md)
            assert np.abs(Tn["t"] - ref32[0]["t"]).max() <= 2e-4 and quat_close(Tn["R"], ref32[0]["R"], 5e-5)
    # single reduction: n_meas exact
    h.find(synth.compose(Tom, Tbo))
    I = synth.make_transform()
    st = h.computeCrossStatistics(I, 0.0)
    mv = h.modelView()
    ref = po.statistics_p2l(I, dp, dm, mv["points"], mv["normals"], mv["hits"], 1.0, f64=True)
    assert st["n_meas"] == ref["n_meas"]
    assert np.abs(st["dataset_mean"] - ref["dataset_mean"]).max() <= TOL_STATS and np.abs(st["covariance"] - ref["covariance"]).max() <= 5e-5
    # round trip: noise-free scan taken at the pose itself -> identity update
    clean = osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"]
    h.setRanges(clean)
    Tn, Td, Cm = h.correctOnce(Tgt, I, 5, 0.0)
    assert np.abs(Td["t"]).max() < 1e-5 and quat_close(Td["R"], [0, 0, 0, 1], 1e-6) and Cm["n_meas"] > 100000


def test_correct_batch_v1(po, synth):
    """v1 correct(Tbm[N]) (lidar_corrector_embree_benchmark.cpp:86-133) on the sphere map with vlp16_900, range.min = 0."""
    name = "uvsphere:40:60"
    osc = oracle_scene(name)
    m = synth.vlp16_900()
    m.range_min = 0.0
    o, d = po.model_rays(m)
    I = synth.make_transform()
    ranges = osc.simulate(I, I, o, d, m.range_max)["ranges"]
    rng = np.random.default_rng(2)
    T = synth.transforms(48)
    T["t"] = rng.uniform(-0.3, 0.3, (48, 3)).astype(np.float32)
    T["t"][0] = (0, 0, 0.2)
    yaw = rng.uniform(-0.1, 0.1, 48)
    T["R"][:, 2], T["R"][:, 3] = np.sin(yaw / 2), np.cos(yaw / 2)
    Tsb = synth.make_transform((0.01, 0, 0.02), (0, 0, 0.05))
    ref = osc.correct_batch(T, Tsb, o, d, m.range_min, m.range_max, ranges, 1.0, f64_accum=True)
    h = _rcc(synth, name, m, Tsb=Tsb)
    h.setInputData(ranges)
    Td, nc, st = h.correct(T)
    assert np.array_equal(nc, ref[1])                                       # Ncorr bit-exact
    assert np.abs(Td["t"] - ref[0]["t"]).max() <= TOL_DT
    assert all(quat_close(a, b, TOL_DT) for a, b in zip(Td["R"], ref[0]["R"]))
    assert np.abs(st["dataset_mean"] - ref[2]["dataset_mean"]).max() <= 5e-6
    # ten benchmark-style iterations move pose 0 back towards the origin
    Tc = T.copy()
    for _ in range(10):
        Td, nc, _ = h.correct(Tc)
        Tc = synth.compose(Tc, Td)
    assert abs(Tc["t"][0, 2]) < 0.2


@pytest.mark.parametrize("ng_mode,corr", [(0, 0), (1, 0), (0, 1)])
def test_pf_sensor_update(po, synth, ng_mode, corr):
    """corr 0: evaluate_rcc (PCDSensorUpdaterEmbree.cpp:18-86); corr 1: evaluate_cpc (:88-95), the closest-point error."""
    import rmcl_b200
    name = "building:200000"
    osc = oracle_scene(name)
    m = synth.c2_sensor()
    o, d = po.model_rays(m)
    Tsb = synth.scenario_tsb()
    pts = osc.simulate(synth.building_gt_pose(), Tsb, o, d, 80.0)["points"]
    beams = synth.pf_beams(pts, 180)
    beams["range"][:3] = (0.01, 200.0, 90.0)                     # real misses (out of sensor range) exercise the penalty branches
    P, A = synth.pf_particles(5000)
    A["likelihood"]["n_meas"][:10] = 9990                        # clamp at MAX_N_MEAS
    prm = po.PFParams.defaults(ng_mode, corr)
    ref = osc.pf_update(P, A, Tsb, beams, prm)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map(name))
    out = up.update(P, A, Tsb, beams, rmcl_b200.PFParams.defaults(ng_mode, corr))
    assert np.array_equal(out["likelihood"]["n_meas"], ref["likelihood"]["n_meas"])
    assert np.array_equal(out["state_sigma"], ref["state_sigma"])
    assert np.abs(out["likelihood"]["mean"] - ref["likelihood"]["mean"]).max() <= TOL_LIK
    assert np.abs(out["likelihood"]["sigma"] - ref["likelihood"]["sigma"]).max() <= TOL_LIK
    frac_exact = np.mean(out["likelihood"]["mean"] == ref["likelihood"]["mean"])
    assert frac_exact > 0.99
    # sharding invariance: the two halves computed separately equal the whole, bit for bit
    a = up.update(P[:2500], A[:2500], Tsb, beams, rmcl_b200.PFParams.defaults(ng_mode, corr))
    b = up.update(P[2500:], A[2500:], Tsb, beams, rmcl_b200.PFParams.defaults(ng_mode, corr))
    assert np.concatenate([a, b]).tobytes() == out.tobytes()
    # device-resident variant (ParticleUpdater<VRAM_CUDA>)
    import torch
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
    up.update(Pd, Ad, Tsb, beams, rmcl_b200.PFParams.defaults(ng_mode, corr))
    torch.cuda.synchronize()
    assert Ad.cpu().numpy().view(synth.PARTICLE_ATTR_DTYPE).reshape(-1).tobytes() == out.tobytes()


def test_pf_golden(po, synth):
    import rmcl_b200
    g = np.load(os.path.join(GOLD, "pf_cube.npz"))
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map("cube29"))
    for ng in (0, 1):
        out = up.update(g["poses"], g["attrs0"], g["Tsb"], g["beams"], rmcl_b200.PFParams.defaults(ng))
        ref = g[f"attrs_ng{ng}"]
        assert np.array_equal(out["likelihood"]["n_meas"], ref["likelihood"]["n_meas"])
        assert np.abs(out["likelihood"]["mean"] - ref["likelihood"]["mean"]).max() <= TOL_LIK


def test_edge_cases(po, synth):
    import rmcl_b200
    # empty map -> B2_ERR_NO_MAP ("EMPTY MAP", PCDSensorUpdaterOptix.cpp:187-192)
    with pytest.raises(rmcl_b200.B2Error) as e:
        rmcl_b200.Map(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    assert e.value.code == -3
    # face index out of range
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map(np.zeros((3, 3), np.float32), np.array([[0, 1, 7]], np.uint32))
    # single-triangle map; rays that miss everything are NaN-encoded
    one = rmcl_b200.Map(np.array([[5, -1, -1], [5, 1, -1], [5, 0, 1]], np.float32), np.array([[0, 1, 2]], np.uint32))
    t, f, ng, hit = one.intersect([[0, 0, 0], [0, 0, 0]], [[1, 0, 0], [0, 1, 0]])
    assert hit.tolist() == [1, 0] and f[0] == 0 and f[1] == 0xFFFFFFFF and abs(t[0] - 5) < 1e-6
    h = rmcl_b200.RCCB200Spherical(one)
    m = synth.SphericalModel(0.0, 0.0, 1, -np.pi, 2 * np.pi / 8, 8, 0.1, 50.0)
    h.setTsb(synth.make_transform())
    h.setModel(m)
    h.find(synth.make_transform())
    mv = h.modelView()
    assert mv["hits"].sum() == 1 and np.isnan(mv["points"][mv["hits"] == 0]).all()
    # dataset size mismatch / call order errors
    with pytest.raises(rmcl_b200.B2Error):
        h.setRanges(np.ones(5, np.float32))
    h2 = rmcl_b200.RCCB200Spherical(one)
    with pytest.raises(rmcl_b200.B2Error):
        h2.find(synth.make_transform())                              # find before setModel
    # zero-size model: find silently returns (RCCOptix.cpp:30-34)
    h2.setModel(synth.SphericalModel(0, 0, 0, 0, 0, 0, 0.1, 1.0))
    h2.find(synth.make_transform())
    # fully masked dataset -> n_meas = 0 -> identity update, Tom unchanged (micp_localization.cpp:974)
    h.setDataset(np.zeros((8, 3), np.float32), np.zeros(8, np.uint8))
    Tom = synth.make_transform((0.1, 0.2, 0.3), (0, 0, 0.4))
    Tn, Td, Cm = h.correctOnce(Tom, synth.make_transform(), 5, 0.0)
    assert Cm["n_meas"] == 0 and np.array_equal(Tn["t"], Tom["t"]) and np.array_equal(Tn["R"], Tom["R"])
    # PF with zero particles / zero beams is a no-op
    up = rmcl_b200.PCDSensorUpdaterB200(one)
    P, A = synth.pf_particles(4)
    out = up.update(P, A, synth.make_transform(), np.zeros(0, synth.RANGE_MEAS_DTYPE))
    assert out.tobytes() == A.tobytes()


def test_full_size_properties_c3(po, synth):
    """BASELINE sizes where the oracle is too slow for a full compare: C3 (100k particles x 180 beams, 1M triangles) checked through
    size-independent properties: oracle parity on a sample, sharding invariance, merge-count, bounded likelihood."""
    import rmcl_b200
    name = "building:1000000"
    osc = oracle_scene(name)
    m = synth.c2_sensor()
    o, d = po.model_rays(m)
    Tsb = synth.scenario_tsb()
    pts = osc.simulate(synth.building_gt_pose(), Tsb, o, d, 80.0)["points"]
    beams = synth.pf_beams(pts, 180)
    P, A = synth.pf_particles(100_000)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map(name))
    prm = rmcl_b200.PFParams.defaults()
    out = up.update(P, A, Tsb, beams, prm)
    assert (out["likelihood"]["n_meas"] == 180).all()
    assert (out["likelihood"]["mean"] >= 0).all() and (out["likelihood"]["mean"] <= 0.19947115).all()
    idx = np.random.default_rng(0).choice(100_000, 3000, replace=False)
    ref = osc.pf_update(P[idx], A[idx], Tsb, beams, po.PFParams.defaults())
    assert np.abs(out["likelihood"]["mean"][idx] - ref["likelihood"]["mean"]).max() <= TOL_LIK
    parts = [up.update(P[a:b], A[a:b], Tsb, beams, prm) for a, b in ((0, 12500), (12500, 50000), (50000, 100000))]
    assert np.concatenate(parts).tobytes() == out.tobytes()
    # ray mappings (lanes = beams / particles / particles sorted by pose / chosen by timing): a schedule, bit-identical results
    import torch
    Pd0 = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    for mode, reps in ((0, 1), (1, 1), (2, 1), (3, 4)):
        up.setMapping(mode)
        for _ in range(reps):
            Am = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
            up.update(Pd0, Am, Tsb, beams, prm)
            assert Am.cpu().numpy().tobytes() == out.tobytes(), mode
    assert up.mapping()[0] == 3 and up.mapping()[1] in (0, 2)
    # host arrays of this size go through the chunked two-stream path; it must equal the device-resident single launch, pinned or not, in place or not
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
    up.update(Pd, Ad, Tsb, beams, prm)
    assert Ad.cpu().numpy().tobytes() == out.tobytes()
    Ph = torch.from_numpy(P.view(np.uint8).copy()).pin_memory().numpy().view(P.dtype).reshape(-1)
    Ah = torch.from_numpy(A.view(np.uint8).copy()).pin_memory().numpy().view(A.dtype).reshape(-1)
    assert up.update(Ph, Ah, Tsb, beams, prm, inplace=True) is Ah and Ah.tobytes() == out.tobytes()
    # second update accumulates: n_meas 360, and equals updating with the beams concatenated twice
    out2 = up.update(P[:2000], out[:2000], Tsb, beams, prm)
    both = up.update(P[:2000], A[:2000], Tsb, np.concatenate([beams, beams]), prm)
    assert (out2["likelihood"]["n_meas"] == 360).all() and out2.tobytes() == both.tobytes()


@pytest.mark.parametrize("name,lo,hi,n", [("cube29", -9.5, 9.5, 30000), ("building:200000", 1.0, 2.9, 100000), ("uvsphere:40:60", -6.0, 6.0, 30000)])
def test_device_lbvh_build_bit_exact(po, synth, name, lo, hi, n):
    """Map built entirely on the device (Morton + radix sort + Karras hierarchy + wide collapse): same answers as the oracle (and hence as
    the host SAH build) -- the hit definition does not depend on the tree."""
    import rmcl_b200
    V, F = mesh(name)
    lb = rmcl_b200.Map(V, F, device=0, build_mode=rmcl_b200.api.B2_BUILD_DEVICE_LBVH)
    info = lb.info()
    assert info["n_leaf_tris"] == len(F) and info["build_mode"] == 1 and 0 < info["max_depth"] < 36
    o, d = random_rays(n, lo, hi, seed=9)
    t1, f1, n1, h1 = oracle_scene(name).intersect(o, d)
    t2, f2, n2, h2 = lb.intersect(o, d)
    assert np.array_equal(h1, h2) and np.array_equal(f1, f2) and np.array_equal(t1, t2) and np.array_equal(n1, n2)
    # a full sensor path on the device-built map
    if name.startswith("building"):
        m = synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 31, 32, -np.pi, 2 * np.pi / 256, 256, 0.5, 120.0)
        oo, dd = po.model_rays(m)
        ref = oracle_scene(name).simulate(synth.building_gt_pose(), synth.scenario_tsb(), oo, dd, m.range_max)
        h = rmcl_b200.RCCB200Spherical(lb)
        h.setTsb(synth.scenario_tsb())
        h.setModel(m)
        h.find(synth.building_gt_pose())
        mv = h.modelView()
        assert all(np.array_equal(mv[k], ref[k], equal_nan=True) for k in ref)
    # degenerate inputs: one triangle, coincident triangles
    one = rmcl_b200.Map(np.array([[5, -1, -1], [5, 1, -1], [5, 0, 1]], np.float32), np.array([[0, 1, 2]], np.uint32), build_mode=1)
    t, f, ng, hit = one.intersect([[0, 0, 0]], [[1, 0, 0]])
    assert hit[0] == 1 and f[0] == 0
    Vd = np.tile(np.array([[5, -1, -1], [5, 1, -1], [5, 0, 1]], np.float32), (5, 1))
    Fd = np.arange(15, dtype=np.uint32).reshape(5, 3)[::-1].copy()
    dup = rmcl_b200.Map(Vd, Fd, build_mode=1)
    t, f, ng, hit = dup.intersect([[0, 0, 0]], [[1, 0, 0]])
    assert hit[0] == 1 and f[0] == 0                                       # tie -> smallest face id


def test_pf_motion_and_stats_gpu(po, synth):
    """SURVEY 8f2 on the device: motion update bit-exact (same quaternion op order), statistics max exact / sum to FP32 noise."""
    import torch
    import rmcl_b200
    P, A = synth.pf_particles(100_003)
    rng = np.random.default_rng(1)
    A["likelihood"]["mean"] = rng.uniform(0, 0.2, len(A)).astype(np.float32)
    A["likelihood"]["n_meas"] = rng.integers(0, 10001, len(A)).astype(np.uint32)
    T = synth.make_transform((0.1, -0.02, 0.0), (0, 0, 0.05))
    Pr, Ar = po.pf_motion_update(P, A, T, 0.03)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map("cube29"))
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
    up.motionUpdate(Pd, Ad, T, 0.03)
    torch.cuda.synchronize()
    Pg = Pd.cpu().numpy().view(synth.TRANSFORM_DTYPE).reshape(-1)
    Ag = Ad.cpu().numpy().view(synth.PARTICLE_ATTR_DTYPE).reshape(-1)
    assert np.array_equal(Pg["R"], Pr["R"]) and np.array_equal(Pg["t"], Pr["t"])
    assert Ag.tobytes() == Ar.tobytes()
    s, m = up.likelihoodStats(Ad)
    sr, mr = po.pf_likelihood_stats(Ar)
    assert m == mr and abs(s - sr) <= 2e-6 * abs(sr) + 1e-6
    s0, m0 = up.likelihoodStats(Ad[:0])
    assert s0 == 0.0 and m0 == 0.0


@pytest.mark.parametrize("name,lo,hi", [("cube29", [-12] * 3, [12] * 3), ("building:1000000", [-1, -1, -0.5], [61, 41, 3.5])])
def test_cpc_find_bit_exact(po, synth, name, lo, hi):
    """CPCEmbree::find (CPCEmbree.cpp:17-43) through the C ABI == oracle, bit for bit (points, normals, hits, faces, distances)."""
    import rmcl_b200
    osc = oracle_scene(name)
    rng = np.random.default_rng(11)
    q = rng.uniform(lo, hi, (100000, 3)).astype(np.float32)
    q[::1009] = np.nan                                               # masked-out dataset entries may hold anything; the mask is not consulted
    Tbm, Tsb = synth.make_transform([0.3, -0.2, 0.1], [0.02, 0.01, 0.7]), synth.scenario_tsb()
    h = rmcl_b200.CPCB200(gpu_map(name))
    h.setTsb(Tsb); h.setParams(0.8, 0.15)
    h.setDataset(q)
    h.find(Tbm)
    mv, ref = h.modelView(), osc.cpc_find(Tbm, Tsb, q, 0.8)
    assert len(mv["hits"]) == len(q)
    assert np.array_equal(mv["hits"], ref["hits"]) and np.array_equal(mv["face_ids"], ref["face_ids"])
    assert np.array_equal(mv["ranges"], ref["dists"]) and np.array_equal(mv["points"], ref["points"], equal_nan=True)
    assert np.array_equal(mv["normals"], ref["normals"], equal_nan=True)
    assert 0.05 < ref["hits"].mean() < 0.95
    # cross statistics on the closest-point pairs
    I = synth.make_transform()
    dm = np.ones(len(q), np.uint8)
    st = h.computeCrossStatistics(I, 0.0)
    r = po.statistics_p2l(I, q, dm, ref["points"], ref["normals"], ref["hits"], 0.8, f64=True)
    assert st["n_meas"] == r["n_meas"] and np.abs(st["dataset_mean"] - r["dataset_mean"]).max() <= 5e-6
    with pytest.raises(rmcl_b200.B2Error):
        h.setModel(synth.c1_sensor())
    with pytest.raises(rmcl_b200.B2Error):
        h.correctOnce(Tbm, I, 5, 0.0, ranges=np.ones(8, np.float32))
    # empty dataset
    h.setDataset(np.zeros((0, 3), np.float32))
    h.find(Tbm)
    assert len(h.modelView()["hits"]) == 0


def test_cpc_correct_once(po, synth):
    """MICP-L correctOnce with closest-point correspondences: find + 5 inner iterations on the device vs the oracle chain."""
    import rmcl_b200
    name, m = "building:1000000", synth.c2_sensor()
    osc = oracle_scene(name)
    o, d = po.model_rays(m)
    Tgt, Tsb = synth.building_gt_pose(), synth.scenario_tsb()
    ranges = synth.noisy_ranges(osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"], m.range_max)
    dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    Tbo = synth.make_transform((0.05, 0.02, 0.0), (0, 0, 0.1))
    Tom = synth.compose(synth.compose(Tgt, synth.scenario_pose_offset()), synth.inverse(Tbo))
    h = rmcl_b200.CPCB200(gpu_map(name))
    h.setTsb(Tsb); h.setParams(1.0, 0.15)
    h.setDataset(dp, dm)
    for cp in (0.0, 0.6):
        ref = osc.micp_correct_once(None, None, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, cp, f64_accum=True)
        Tn, Td, Cm = h.correctOnce(Tom, Tbo, 5, cp)
        assert abs(int(Cm["n_meas"]) - int(ref[2]["n_meas"])) <= 2
        assert np.abs(Tn["t"] - ref[0]["t"]).max() <= TOL_DT and quat_close(Tn["R"], ref[0]["R"], TOL_DT)
        assert np.abs(Td["t"] - ref[1]["t"]).max() <= TOL_DT and quat_close(Td["R"], ref[1]["R"], TOL_DT)
    # skip_masked: the masked-out points (dropped beams, far outside the map) are not queried; nothing the statistics use changes
    base = h.correctOnce(Tom, Tbo, 5, 0.0)
    mv_all = h.modelView()
    h.setOptions(skip_masked=True)
    skip = h.correctOnce(Tom, Tbo, 5, 0.0)
    mv_skip = h.modelView()
    assert skip[0].tobytes() == base[0].tobytes() and skip[2].tobytes() == base[2].tobytes()
    keep = dm > 0
    assert np.array_equal(mv_skip["points"][keep], mv_all["points"][keep]) and (mv_skip["hits"][~keep] == 0).all() and np.isnan(mv_skip["points"][~keep]).all()
    h.setOptions(skip_masked=False)
    # a scan already on the map: every point has distance ~0 -> (near-)identity update
    clean = osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"]
    dp2, dm2, _ = po.dataset_from_ranges(o, d, clean, m.range_min, m.range_max)
    h.setDataset(dp2, dm2)
    Tn, Td, Cm = h.correctOnce(Tgt, synth.make_transform(), 5, 0.0)
    assert np.abs(Td["t"]).max() < 1e-4 and quat_close(Td["R"], [0, 0, 0, 1], 1e-5) and Cm["n_meas"] > 100000


def test_gladiator_resample_gpu(po, synth):
    """gladiator_resample (resampling.cu:108-221) on the device.  Integer work (Philox words, opponent choice, win/lose decision,
    copied fields) is bit-exact; the perturbed pose goes through device sinf/cosf/atan2f/logf (<= 2 ulp from glibc): tolerance 2e-6."""
    import torch
    import rmcl_b200
    from test_oracle import _glad_particles
    n = 200000
    P, A = _glad_particles(synth, n)
    cfg_o = po.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    cfg = rmcl_b200.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map("cube29"))
    raw_o, nrm_o = po.pf_gladiator_randoms(1234, 3, 0, n)
    raw_d, nrm_d = up.gladiatorRandoms(1234, 3, 0, n)
    torch.cuda.synchronize()
    assert np.array_equal(raw_d.cpu().numpy().view(np.uint32), raw_o)                               # Philox4x32-10: bit-exact
    assert np.abs(nrm_d.cpu().numpy() - nrm_o).max() <= 4e-6                                        # Box-Muller through device logf/sinf/cosf
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()

    def to_np(Pt, At):
        torch.cuda.synchronize()
        return Pt.cpu().numpy().view(P.dtype).reshape(-1), At.cpu().numpy().view(A.dtype).reshape(-1)

    ref_P, ref_A = po.pf_gladiator_resample(P, A, 0, n, raw_o, nrm_o, cfg_o)
    for external in (True, False):
        Pn, An = torch.empty_like(Pd), torch.empty_like(Ad)
        if external:   # the oracle's draws uploaded: isolates the resampling arithmetic from the generator
            up.resample(Pd, Ad, Pn, An, cfg, raw=torch.from_numpy(raw_o.view(np.int32)).cuda(), normals=torch.from_numpy(nrm_o).cuda())
        else:
            up.resample(Pd, Ad, Pn, An, cfg, seed=1234, step=3)
        gP, gA = to_np(Pn, An)
        assert np.array_equal(gA["likelihood"]["mean"], ref_A["likelihood"]["mean"]) and np.array_equal(gA["state_sigma"], ref_A["state_sigma"])
        assert np.array_equal(gP["stamp"], ref_P["stamp"])                                          # who won where: exact
        if external:
            assert np.array_equal(gP["t"], ref_P["t"])                                              # t + N*noise: two individually rounded ops, bit-exact
        else:
            assert np.abs(gP["t"] - ref_P["t"]).max() <= 8e-6                                       # device normals differ by ulps -> at most 1 ulp of |t| <= 64
        assert np.abs(gP["R"] - ref_P["R"]).max() <= 2e-6
        dn = np.abs(gA["likelihood"]["n_meas"].astype(np.int64) - ref_A["likelihood"]["n_meas"].astype(np.int64))
        assert dn.max() <= 1 and (dn == 0).mean() > 0.999                                           # uint *= float truncation next to an ulp of pow()
    whole_P, whole_A = gP, gA
    # sharding invariance on the device: champions [0, h) and [h, n) against all n opponents == the whole, bit for bit
    h = n // 2
    Pa, Aa, Pb, Ab = torch.empty_like(Pd[:h]), torch.empty_like(Ad[:h]), torch.empty_like(Pd[h:]), torch.empty_like(Ad[h:])
    up.resample(Pd, Ad, Pa, Aa, cfg, seed=1234, step=3, first=0)
    up.resample(Pd, Ad, Pb, Ab, cfg, seed=1234, step=3, first=h)
    a, b = to_np(Pa, Aa), to_np(Pb, Ab)
    assert np.concatenate([a[0], b[0]]).tobytes() == whole_P.tobytes() and np.concatenate([a[1], b[1]]).tobytes() == whole_A.tobytes()
    with pytest.raises(rmcl_b200.B2Error):
        up.resample(Pd, Ad, Pd, Ad, cfg)                                                            # in place is refused


@pytest.mark.parametrize("name", ["cube29", "building:200000"])
def test_segmentation_gpu(po, synth, name):
    """scan-vs-map segmentation through the C ABI == oracle: labels and both outlier clouds (raster order), bit for bit."""
    import rmcl_b200
    from test_oracle import _segmentation_case
    osc = oracle_scene(name)
    m = synth.c2_sensor() if name != "cube29" else synth.c1_sensor()
    o, d = po.model_rays(m)
    T = synth.building_gt_pose() if name != "cube29" else synth.make_transform([0.5, -0.3, 0.2], [0, 0, 0.4])
    Tsb = synth.scenario_tsb()
    sim = osc.simulate(T, Tsb, o, d, m.range_max)
    rng = np.random.default_rng(4)
    real = sim["ranges"] + rng.normal(0, 0.01, len(o if len(o) > 1 else d)).astype(np.float32)
    k = rng.permutation(len(real))
    real[k[:300]] *= 0.6; real[k[300:600]] *= 1.3; real[k[600:650]] = m.range_max + 1; real[k[650:670]] = 0.0
    h = _rcc(synth, name, m, Tsb)
    h.setRanges(real)
    h.find(T)
    a, b, lab = h.segment(0.15, 0.1)
    ra, rb, rl = po.segment(o, d, m.range_min, m.range_max, real, sim["ranges"], sim["normals"], 0.15, 0.1)
    assert np.array_equal(lab, rl) and np.array_equal(a, ra) and np.array_equal(b, rb)
    assert len(a) >= 300 and len(b) >= 300
    h2 = _rcc(synth, name, m, Tsb)
    with pytest.raises(rmcl_b200.B2Error):
        h2.segment()                                                     # before find / without ranges


def test_map_from_file_gpu(tmp_path, synth):
    """Map.from_file (rm::import_embree_map twin): the imported mesh traces exactly like the in-memory one."""
    import rmcl_b200
    from test_abi import _write_ply
    V, F = mesh("cube29")
    p = str(tmp_path / "cube.ply")
    _write_ply(p, [tuple(float(x) for x in v) for v in V], [tuple(int(i) for i in t) for t in F], True)
    o, d = random_rays(20000, -9.5, 9.5, seed=3)
    a = rmcl_b200.Map.from_file(p).intersect(o, d)
    b = gpu_map("cube29").intersect(o, d)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    (tmp_path / "bad.dae").write_text("<COLLADA/>")
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map.from_file(str(tmp_path / "bad.dae"))


def test_destroy_order_gpu(synth):
    """The map may be destroyed before the handles created on it (shared ownership, like rm::EmbreeMapPtr): no dangling access, no pending error."""
    import rmcl_b200
    V, F = mesh("cube29")
    m = rmcl_b200.Map(V, F)
    h = rmcl_b200.RCCB200Spherical(m)
    up = rmcl_b200.PCDSensorUpdaterB200(m)
    h.setModel(synth.c1_sensor())
    m.close()                                       # creator's reference gone; the handles keep the BVH alive
    h.find(synth.make_transform())
    assert h.modelView()["hits"].sum() > 0
    h.close(); up.close()
    assert rmcl_b200.load_library().b2_peek_cuda_error() == b""


def test_bvh_blob_roundtrip_gpu(synth):
    """build once, ship the BVH blob; the re-created map traces bit-identically; corrupt blobs are refused."""
    import rmcl_b200
    m = gpu_map("cube29")
    blob = m.export_blob()
    assert blob[:7].tobytes() == b"B2BVH8F" and len(blob) == 64 + m.info()["bvh_bytes"]
    m2 = rmcl_b200.Map.from_blob(blob)
    assert {k: v for k, v in m2.info().items() if k != "build_ms"} == {k: v for k, v in m.info().items() if k != "build_ms"}
    o, d = random_rays(20000, -9.5, 9.5, seed=5)
    assert all(np.array_equal(x, y) for x, y in zip(m.intersect(o, d), m2.intersect(o, d)))
    bad = blob.copy(); bad[64 + 192:64 + 196] = 0xff                          # child_base of the root node far outside the node array
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map.from_blob(bad)
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map.from_blob(blob[: len(blob) // 2])
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map.from_blob(np.zeros(100, np.uint8))


def test_motion_update_collision_gpu(po, synth):
    """motion update with the wall check (collision ray through our tracer), bit-exact vs the oracle."""
    import torch
    import rmcl_b200
    from test_oracle import _motion_case
    osc = oracle_scene("cube29")
    P, A, T = _motion_case(synth, 50000)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map("cube29"))
    for collide in (False, True):
        ref_P, ref_A = po.pf_motion_update(P, A, T, 0.03, scene=osc if collide else None)
        Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
        Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
        up.motionUpdate(Pd, Ad, T, 0.03, check_collision=collide)
        torch.cuda.synchronize()
        gP = Pd.cpu().numpy().view(P.dtype).reshape(-1)
        gA = Ad.cpu().numpy().view(A.dtype).reshape(-1)
        assert np.array_equal(gP["R"], ref_P["R"]) and np.array_equal(gP["t"], ref_P["t"])
        assert gA.tobytes() == ref_A.tobytes()
    assert (gA["likelihood"]["n_meas"] == 10000).mean() > 0.02


def test_refit_dynamic_map_gpu(po, synth):
    """vertices move, faces stay -> refit instead of rebuild; every result equals the oracle on the moved mesh (the hit definition
    does not depend on the tree), also through a handle that was created before the refit."""
    import rmcl_b200
    V, F = mesh("building:200000")
    m = rmcl_b200.Map(V, F)
    sensor = synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 31, 32, -np.pi, 2 * np.pi / 256, 256, 0.5, 120.0)
    h = rmcl_b200.RCCB200Spherical(m)
    h.setTsb(synth.scenario_tsb()); h.setModel(sensor)
    rng = np.random.default_rng(12)
    V2 = (V + rng.normal(0, 0.02, V.shape)).astype(np.float32)               # every vertex moved by centimetres
    V2[:, 0] += (0.3 * np.sin(V[:, 1] * 0.2)).astype(np.float32)             # plus a smooth bend of the whole building
    m.refit(V2)
    osc2 = po.Scene(V2, F)
    o, d = random_rays(50000, 1.0, 2.9, seed=13)
    o[:, 0] *= 20; o[:, 1] *= 13
    t1, f1, n1, h1 = osc2.intersect(o, d)
    t2, f2, n2, h2 = m.intersect(o, d)
    assert np.array_equal(h1, h2) and np.array_equal(f1, f2) and np.array_equal(t1, t2) and np.array_equal(n1, n2)
    oo, dd = po.model_rays(sensor)
    ref = osc2.simulate(synth.building_gt_pose(), synth.scenario_tsb(), oo, dd, sensor.range_max)
    h.find(synth.building_gt_pose())
    mv = h.modelView()
    assert all(np.array_equal(mv[k], ref[k], equal_nan=True) for k in ref)
    q = rng.uniform([1, 1, 0.2], [59, 39, 2.8], (20000, 3)).astype(np.float32)
    hc = rmcl_b200.CPCB200(m)
    hc.setParams(1.0, 0.15); hc.setDataset(q); hc.find(synth.make_transform())
    a, b = hc.modelView(), osc2.cpc_find(synth.make_transform(), synth.make_transform(), q, 1.0)
    assert np.array_equal(a["face_ids"], b["face_ids"]) and np.array_equal(a["ranges"], b["dists"])
    with pytest.raises(rmcl_b200.B2Error):
        m.refit(V2[:-1])                                                     # topology must stay
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map(V, F, build_mode=0).refit(V2)                          # host-built maps keep no refit data


def test_widened_rows_golden_gpu(po, synth):
    """The SURVEY 8(f) rows through the C ABI against the committed fixtures (tests/golden/f_rows.npz)."""
    import torch
    import rmcl_b200
    g = np.load(os.path.join(GOLD, "f_rows.npz"))
    gm = gpu_map("cube29")
    hc = rmcl_b200.CPCB200(gm)
    hc.setTsb(g["Tsb"]); hc.setParams(0.8, 0.15); hc.setDataset(g["queries"]); hc.find(g["Tgt"])
    mv = hc.modelView()
    assert np.array_equal(mv["face_ids"], g["cpc_faces"]) and np.array_equal(mv["ranges"], g["cpc_dists"]) and np.array_equal(mv["hits"], g["cpc_hits"])
    assert np.array_equal(mv["points"], g["cpc_points"], equal_nan=True) and np.array_equal(mv["normals"], g["cpc_normals"], equal_nan=True)
    up = rmcl_b200.PCDSensorUpdaterB200(gm)
    A1 = up.update(g["poses"], g["attrs0"], g["Tsb"], g["beams"], rmcl_b200.PFParams.defaults(0, 1))
    assert np.abs(A1["likelihood"]["mean"] - g["attrs_cpc"]["likelihood"]["mean"]).max() <= TOL_LIK
    Pd = torch.from_numpy(g["poses"].view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(g["attrs_cpc"].view(np.float32).reshape(-1, 9).copy()).cuda()
    up.motionUpdate(Pd, Ad, g["T_motion"], 0.03, check_collision=True)
    torch.cuda.synchronize()
    assert Pd.cpu().numpy().tobytes() == g["poses_moved"].view(np.float32).tobytes() and Ad.cpu().numpy().tobytes() == g["attrs_moved"].view(np.float32).tobytes()
    raw, nrm = up.gladiatorRandoms(1234, 3, 0, len(g["poses"]))
    torch.cuda.synchronize()
    assert np.array_equal(raw.cpu().numpy().view(np.uint32), g["glad_raw"]) and np.abs(nrm.cpu().numpy() - g["glad_normals"]).max() <= 4e-6
    m = synth.c1_sensor()
    h = _rcc(synth, "cube29", m, g["Tsb"])
    h.setRanges(g["real_ranges"]); h.find(g["Tgt"])
    a, b, lab = h.segment(0.15, 0.1)
    assert np.array_equal(lab, g["seg_labels"]) and np.array_equal(a, g["seg_scan"]) and np.array_equal(b, g["seg_map"])


# =====================================================================================================================
# round 2: the configurations and branches round 1 left untested on hardware (C4 end to end, > 151 552 pairs,
# exec modes 0 / 1, v1 batch on the other sensor models, concurrent handles, multi-sensor correctOnce, async entry)
# =====================================================================================================================
TOL_DT_F32 = 5e-3      # against the oracle's FP32 one-element-at-a-time merges: that chain's own rounding (C2, 131k pairs: 2.3e-5 m; C4, 307k pairs: 1.2e-3 m
                       # measured) -- the worst case of FP32 accumulation; the reference's OpenMP/TBB reduction merges per-thread partials and sits in between


def _micp_case(po, synth, name, m, Tgt, seed=42):
    osc = oracle_scene(name)
    o, d = po.model_rays(m)
    Tsb = synth.scenario_tsb()
    ranges = synth.noisy_ranges(osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"], m.range_max, seed=seed)
    dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    Tbo = synth.make_transform((0.05, 0.02, 0.0), (0, 0, 0.1))
    Tom = synth.compose(synth.compose(Tgt, synth.scenario_pose_offset()), synth.inverse(Tbo))
    return osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom


def _assert_micp(out, ref, tol=TOL_DT, dn=2):
    Tn, Td, Cm = out
    assert abs(int(Cm["n_meas"]) - int(ref[2]["n_meas"])) <= dn, (int(Cm["n_meas"]), int(ref[2]["n_meas"]))
    assert np.abs(Tn["t"] - ref[0]["t"]).max() <= tol and quat_close(Tn["R"], ref[0]["R"], tol)
    assert np.abs(Td["t"] - ref[1]["t"]).max() <= tol and quat_close(Td["R"], ref[1]["R"], tol)


def test_c4_pinhole_correct_once(po, synth):
    """C4 (BASELINE config 4): PinholeCorrector, 640 x 480 depth camera on the 500k-triangle indoor mesh -- 307 200 pairs, i.e. more than two per
    thread of the ICP loop (pairs beyond the registers live in shared memory).  RCCEmbreePinhole::find (RCCEmbree.cpp:58-68) + correctOnce
    (micp_localization.cpp:899-984) against the oracle; resident dataset, pageable host scan (side-stream upload) and pinned host scan (zero copy)."""
    import torch
    import rmcl_b200
    name, m = "indoor:500000", synth.c4_sensor()
    osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.indoor_gt_pose())
    h = _rcc(synth, name, m, cls=rmcl_b200.RCCB200Pinhole)
    h.setRanges(ranges)
    for cp in (0.0, 0.5):
        ref64 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, cp, f64_accum=True)
        ref32 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, cp, f64_accum=False)
        assert ref64[2]["n_meas"] > 200000
        pinned = torch.from_numpy(ranges.copy()).pin_memory()
        for src in (None, ranges, pinned):
            if src is pinned:
                h.setRanges(np.full_like(ranges, 2.0))                # scramble the resident dataset: the call must rebuild it from the pinned scan
            out = h.correctOnce(Tom, Tbo, 5, cp, ranges=src)
            _assert_micp(out, ref64)
            _assert_micp(out, ref32, tol=TOL_DT_F32, dn=50)           # the FP32-sequential chain drifts by its own rounding; stated, not hidden
            if src is pinned:
                ds = h.datasetView()
                assert np.array_equal(ds["points"], dp) and np.array_equal(ds["mask"], dm)
        mv = h.modelView()
        sim = osc.simulate(synth.compose(Tom, Tbo), Tsb, o, d, m.range_max)
        assert np.array_equal(mv["face_ids"], sim["face_ids"]) and np.array_equal(mv["points"], sim["points"], equal_nan=True)
    # noise-free scan at the pose itself -> identity update
    h.setRanges(osc.simulate(synth.indoor_gt_pose(), Tsb, o, d, m.range_max)["ranges"])
    Tn, Td, Cm = h.correctOnce(synth.indoor_gt_pose(), synth.make_transform(), 5, 0.0)
    assert np.abs(Td["t"]).max() < 1e-5 and quat_close(Td["R"], [0, 0, 0, 1], 1e-6) and Cm["n_meas"] > 250000
    # the tile schedule covers this scan too (9600 tiles, 2.3 waves of the find kernel): the order stays a permutation
    import ctypes as C
    n_tiles = m.size // 32
    perm, nt = np.zeros(n_tiles, np.uint16), C.c_uint32(0)
    assert rmcl_b200.load_library().b2_rcc_debug_tile_perm(h._h, C.c_void_p(perm.ctypes.data), C.c_uint32(n_tiles), C.byref(nt)) == 0
    assert nt.value == n_tiles and np.array_equal(np.sort(perm), np.arange(n_tiles)) and not np.array_equal(perm, np.arange(n_tiles))


@pytest.mark.parametrize("rows,cols", [(256, 1024), (1024, 1024)])
def test_large_scan_correct_once(po, synth, rows, cols):
    """Spherical models beyond 2 pairs per thread of the ICP loop: 262 144 rays (registers + shared memory) and 1 048 576 rays (more than the
    loop can keep on chip: the tail of every thread's pair list is streamed from L2 in each inner iteration)."""
    name = "building:200000"
    m = synth.SphericalModel(np.radians(-30.0), np.radians(60.0) / (rows - 1), rows, -np.pi, 2 * np.pi / cols, cols, 0.5, 120.0)
    osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.building_gt_pose(), seed=5)
    ref64 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, 0.0, f64_accum=True)
    h = _rcc(synth, name, m)
    h.setRanges(ranges)
    out = h.correctOnce(Tom, Tbo, 5, 0.0)
    _assert_micp(out, ref64, dn=4)
    out2 = h.correctOnce(Tom, Tbo, 5, 0.0, ranges=ranges)
    assert out2[0].tobytes() == out[0].tobytes() and out2[2].tobytes() == out[2].tobytes()       # same pairs, same order of sums: bit-identical
    import torch
    out3 = h.correctOnce(Tom, Tbo, 5, 0.0, ranges=torch.from_numpy(ranges.copy()).pin_memory())
    assert out3[0].tobytes() == out[0].tobytes()


def test_exec_modes_agree(po, synth):
    """b2_rcc_set_exec_mode: 2 = software grid barrier + programmatic launch (default), 1 = cooperative launch of the same kernel, 0 = one
    k_p2l_reduce launch per inner iteration with the reference's own frame-algebra order (icp_step).  All against the oracle; 1 and 2 run the
    same iteration code and differ only in how the block sums cross the grid (FP64 slots behind a grid sync vs 64-bit fixed-point atomics)."""
    name, m = "building:200000", synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 63, 64, -np.pi, 2 * np.pi / 512, 512, 0.5, 120.0)
    osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.building_gt_pose(), seed=9)
    ref64 = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 1.0, 0.15, 0.3, f64_accum=True)
    h = _rcc(synth, name, m)
    h.setRanges(ranges)
    outs = {}
    for mode in (2, 1, 0, 2):
        h.setExecMode(mode)
        outs[mode] = h.correctOnce(Tom, Tbo, 5, 0.3)
        _assert_micp(outs[mode], ref64)
        outr = h.correctOnce(Tom, Tbo, 5, 0.3, ranges=ranges)
        _assert_micp(outr, ref64)
    assert np.abs(outs[1][0]["t"] - outs[2][0]["t"]).max() <= 1e-6 and outs[1][2]["n_meas"] == outs[2][2]["n_meas"]
    # mode 0 reproduces the oracle's chain more closely still: the frame algebra is the reference's, only the sums differ in order
    assert np.abs(outs[0][0]["t"] - ref64[0]["t"]).max() <= 2e-6
    with pytest.raises(Exception):
        h.setExecMode(7)
    # zero iterations: find only, pose untouched (micp_localization.cpp:915 loop not entered)
    h.setExecMode(2)
    Tn, Td, Cm = h.correctOnce(Tom, Tbo, 0, 0.0)
    assert Tn["t"].tobytes() == Tom["t"].tobytes() and Cm["n_meas"] == 0 and quat_close(Td["R"], [0, 0, 0, 1], 0)


def test_tile_schedule_invisible(po, synth):
    """k_rcc_find traces the tiles that were slowest in the previous launch first (order computed inside the ICP loop kernel).  The order must
    always be a permutation of the tiles, and the results must not depend on it."""
    import ctypes as C
    import rmcl_b200
    name, m = "building:200000", synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 63, 64, -np.pi, 2 * np.pi / 512, 512, 0.5, 120.0)
    osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.building_gt_pose(), seed=3)
    h = _rcc(synth, name, m)
    h.setRanges(ranges)
    lib = rmcl_b200.load_library()
    n_tiles = m.size // 32

    def order():
        out, n = np.zeros(n_tiles, np.uint16), C.c_uint32(0)
        assert lib.b2_rcc_debug_tile_perm(h._h, C.c_void_p(out.ctypes.data), C.c_uint32(n_tiles), C.byref(n)) == 0
        return out, n.value

    Tbm = synth.compose(Tom, Tbo)
    h.find(Tbm)
    first = {k: v.copy() for k, v in h.modelView().items()}
    p0, n0 = order()
    assert n0 == n_tiles and np.array_equal(p0, np.arange(n_tiles))           # no durations yet: identity
    ref = h.correctOnce(Tom, Tbo, 5, 0.3)                                        # the loop kernel sorts the durations of this call's find
    p1, _ = order()
    assert np.array_equal(np.sort(p1), np.arange(n_tiles)) and not np.array_equal(p1, p0)
    for mode in (2, 1):
        h.setExecMode(mode)
        for _ in range(3):
            out = h.correctOnce(Tom, Tbo, 5, 0.3)
            assert out[0].tobytes() == ref[0].tobytes() or mode == 1             # same exec mode: bit-identical whatever the order
            assert np.abs(out[0]["t"] - ref[0]["t"]).max() <= 1e-6 and out[2]["n_meas"] == ref[2]["n_meas"]
            pk, _ = order()
            assert np.array_equal(np.sort(pk), np.arange(n_tiles))
    h.setExecMode(2)
    h.find(Tbm)
    again = h.modelView()
    for k in first:
        assert first[k].tobytes() == again[k].tobytes(), k                     # find's outputs bit-identical under a different tile order


def test_exchange_range_fallback(po, synth):
    """The loop's default exchange carries the block sums as 64-bit fixed point (|block partial| < 2^46); sums beyond that must not fail or
    wrap: the call runs again through the cooperative variant (FP64 exchange).  Scene: the cube scaled to 10 000 km, ranges of 5e6 m."""
    import ctypes as C
    import rmcl_b200
    from oracle import pyoracle
    V, F = synth.cube(29)
    V = (V * np.float32(5e5)).astype(np.float32)
    osc, gm = pyoracle.Scene(V, F), rmcl_b200.Map(V, F, device=0)
    m = synth.SphericalModel(np.radians(-60.0), np.radians(120.0) / 31, 32, -np.pi, 2 * np.pi / 64, 64, 1.0, 1e8)
    o, d = po.model_rays(m)
    Tsb = synth.make_transform((0, 0, 0), (0, 0, 0))
    Tgt = synth.make_transform((1e3, -2e3, 5e2), (0, 0, 0.2))
    ranges = osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"]
    dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    Tbo = synth.make_transform((0, 0, 0), (0, 0, 0))
    Tom = synth.compose(Tgt, synth.make_transform((30.0, -20.0, 10.0), (0, 0, 1e-5)))
    ref = osc.micp_correct_once(o, d, m.range_max, dp, dm, Tom, Tbo, Tsb, 5, 200.0, 200.0, 0.0, f64_accum=True)
    assert ref[2]["n_meas"] > 1500
    h = rmcl_b200.RCCB200Spherical(gm)
    h.setTsb(Tsb); h.setModel(m); h.setParams(200.0, 200.0); h.setRanges(ranges)
    lib, cnt = rmcl_b200.load_library(), C.c_ulonglong(0)
    out1 = h.correctOnce(Tom, Tbo, 5, 0.0)
    lib.b2_rcc_debug_reruns(h._h, C.byref(cnt))
    assert cnt.value == 1                                                # the fixed-point exchange declined, the cooperative variant answered
    out2 = h.correctOnce(Tom, Tbo, 5, 0.0)
    lib.b2_rcc_debug_reruns(h._h, C.byref(cnt))
    assert cnt.value == 2 and out1[0].tobytes() == out2[0].tobytes()      # and the handle is consistent afterwards
    for out in (out1, out2):
        assert abs(int(out[2]["n_meas"]) - int(ref[2]["n_meas"])) <= 8
        assert np.abs(out[0]["t"] - ref[0]["t"]).max() <= 5.0 and quat_close(out[0]["R"], ref[0]["R"], 1e-5)      # FP32 coordinates of 5e6 m: ulp 0.5 m


@pytest.mark.parametrize("case", ["pinhole", "o1dn", "ondn"])
def test_correct_batch_other_models(po, synth, case):
    """v1 {Pinhole,O1Dn,OnDn}Corrector::correct(Tbm[N]) (shape: lidar_corrector_embree_benchmark.cpp:86-133) on the indoor scene."""
    import rmcl_b200
    name = "indoor:20000"
    osc = oracle_scene(name)
    rng = np.random.default_rng(12)
    if case == "pinhole":
        m, cls = synth.PinholeModel(160, 120, 131.25, 131.25, 79.5, 59.5, 0.0, 12.0), rmcl_b200.PinholeCorrectorB200
    else:
        dirs = rng.normal(size=(6000, 3)).astype(np.float32)
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        if case == "o1dn":
            m, cls = synth.O1DnModel(600, 10, np.array([0.1, 0.0, 0.05], np.float32), dirs, 0.0, 20.0), rmcl_b200.O1DnCorrectorB200
        else:
            m, cls = synth.OnDnModel(600, 10, rng.uniform(-0.2, 0.2, (6000, 3)).astype(np.float32), dirs, 0.0, 20.0), rmcl_b200.OnDnCorrectorB200
    o, d = po.model_rays(m)
    Tsb = synth.make_transform((0.01, 0, 0.02), (0, 0, 0.05))
    Tgt = synth.indoor_gt_pose()
    ranges = osc.simulate(Tgt, Tsb, o, d, m.range_max)["ranges"]
    T = synth.transforms(24)
    T[:] = Tgt
    T["t"] += rng.uniform(-0.15, 0.15, (24, 3)).astype(np.float32)
    ref = osc.correct_batch(T, Tsb, o, d, m.range_min, m.range_max, ranges, 1.0, f64_accum=True)
    h = _rcc(synth, name, m, Tsb=Tsb, cls=cls)
    h.setInputData(ranges)
    Td, nc, st = h.correct(T)
    assert np.array_equal(nc, ref[1]) and nc.min() > 1000                   # Ncorr bit-exact
    assert np.abs(Td["t"] - ref[0]["t"]).max() <= TOL_DT
    assert all(quat_close(a, b, TOL_DT) for a, b in zip(Td["R"], ref[0]["R"]))
    # the correction pulls every pose towards the pose the scan was taken at
    Tc = synth.compose(T, Td)
    assert np.linalg.norm(Tc["t"] - Tgt["t"], axis=1).mean() < 0.8 * np.linalg.norm(T["t"] - Tgt["t"], axis=1).mean()


def test_two_handles_two_threads(po, synth):
    """The normal multi-sensor configuration of the reference: one Correspondences object per sensor, driven from different threads on
    different streams (MICPSensor.hpp:65-73).  Two k_icp_loop grids with the software barrier must never hold half of the SMs each: the
    library chains such launches per device.  200 concurrent steps per thread, results bit-identical to the sequential ones, no stall."""
    import threading
    import time
    import torch
    name = "building:200000"
    m1 = synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 63, 64, -np.pi, 2 * np.pi / 1024, 1024, 0.5, 120.0)
    m2 = synth.SphericalModel(np.radians(-15.0), np.radians(30.0) / 31, 32, -np.pi, 2 * np.pi / 512, 512, 0.5, 80.0)
    cases = []
    for m, seed in ((m1, 1), (m2, 2)):
        osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.building_gt_pose(), seed=seed)
        h = _rcc(synth, name, m)
        st = torch.cuda.Stream()
        h.setStream(st.cuda_stream)
        h.setRanges(ranges)
        pinned = torch.from_numpy(ranges.copy()).pin_memory()
        cases.append((h, Tom, Tbo, pinned, st))
    seq = [c[0].correctOnce(c[1], c[2], 5, 0.0, ranges=c[3]) for c in cases]
    errs, outs = [], [None, None]

    def work(i):
        h, Tom, Tbo, pinned, _ = cases[i]
        try:
            for k in range(200):
                outs[i] = h.correctOnce(Tom, Tbo, 5, 0.0, ranges=pinned if k % 2 else None)
                if outs[i][0].tobytes() != seq[i][0].tobytes():
                    raise AssertionError(f"thread {i} step {k}: result differs from the sequential one")
        except Exception as e:                       # noqa: BLE001
            errs.append(e)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    assert not errs, errs
    assert dt < 5.0, f"400 concurrent steps took {dt:.2f} s: the grid barrier stalled"      # one barrier time-out alone is 2 s
    torch.cuda.synchronize()


def test_multi_sensor_correct_once(po, synth):
    """micp_localization.cpp:899-984 over two sensors (spherical LiDAR + pinhole camera, own Tsb / Tbo / merge weight) in one launch sequence:
    finds, then one k_icp_loop whose blocks are split between the sensors; against the oracle's statement-by-statement restatement."""
    import rmcl_b200
    from test_emul_parity import _two_sensor_case
    name = "building:200000"
    osc = oracle_scene(name)
    sensors, Tom = _two_sensor_case(po, synth, osc, scale=4)
    hs = []
    for sd in sensors:
        cls = rmcl_b200.RCCB200Spherical if type(sd["model"]).__name__ == "SphericalModel" else rmcl_b200.RCCB200Pinhole
        h = cls(gpu_map(name))
        h.setTsb(sd["Tsb"]); h.setModel(sd["model"]); h.setParams(1.0, 0.15)
        h.setRanges(sd["ranges"])
        hs.append(h)
    Tbo = np.stack([sd["Tbo"] for sd in sensors])
    for weights in ([1.0, 0.35], [1.0, 3.0], None):
        for k, sd in enumerate(sensors):
            sd["weight"] = 1.0 if weights is None else weights[k]
        for cp in (0.0, 0.4):
            ref = osc.micp_correct_once_multi(sensors, Tom, 5, cp, f64_accum=True)
            out = rmcl_b200.micp_correct_once(hs, Tbo, Tom, 5, cp, merge_weights=weights)
            _assert_micp(out, ref, dn=3)
            out2 = rmcl_b200.micp_correct_once(hs, Tbo, Tom, 5, cp, merge_weights=weights, ranges=[sd["ranges"] for sd in sensors])
            _assert_micp(out2, ref, dn=3)
    assert ref[2]["n_meas"] > 20000
    # one sensor through the multi entry == the single-sensor entry, bit for bit
    a = rmcl_b200.micp_correct_once(hs[:1], Tbo[:1], Tom, 5, 0.0)
    b = hs[0].correctOnce(Tom, Tbo[0], 5, 0.0)
    assert a[0].tobytes() == b[0].tobytes() and a[2].tobytes() == b[2].tobytes()
    # a CPC sensor next to a ray-casting one
    hc = rmcl_b200.CPCB200(gpu_map(name))
    hc.setTsb(sensors[1]["Tsb"]); hc.setParams(1.0, 0.15)
    hc.setDataset(sensors[1]["dataset_points"], sensors[1]["dataset_mask"])
    mixed = [sensors[0], dict(sensors[1], dirs=None, weight=1.0)]
    mixed[0]["weight"] = 1.0
    ref = osc.micp_correct_once_multi(mixed, Tom, 5, 0.0, f64_accum=True)
    out = rmcl_b200.micp_correct_once([hs[0], hc], Tbo, Tom, 5, 0.0)
    _assert_micp(out, ref, dn=3)
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.micp_correct_once([hs[0], hs[0]], Tbo, Tom, 5, 0.0)


def test_correct_once_async(po, synth):
    """b2_rcc_correct_once_async / _wait: enqueue, do something else, collect; one call pending per handle."""
    import torch
    import rmcl_b200
    name, m = "building:200000", synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 63, 64, -np.pi, 2 * np.pi / 512, 512, 0.5, 120.0)
    osc, o, d, Tsb, ranges, dp, dm, Tbo, Tom = _micp_case(po, synth, name, m, synth.building_gt_pose(), seed=9)
    h = _rcc(synth, name, m)
    h.setRanges(ranges)
    sync = h.correctOnce(Tom, Tbo, 5, 0.0)
    for mode in (2, 1, 0):
        h.setExecMode(mode)
        want = h.correctOnce(Tom, Tbo, 5, 0.0)
        h.correctOnceAsync(Tom, Tbo, 5, 0.0)
        with pytest.raises(rmcl_b200.B2Error):
            h.correctOnce(Tom, Tbo, 5, 0.0)                            # synchronous entry refuses while a call is in flight
        x = torch.ones(1 << 20, device="cuda").sum().item()              # unrelated work while the step runs
        got = h.correctOnceWait()
        assert x == float(1 << 20) and got[0].tobytes() == want[0].tobytes() and got[2].tobytes() == want[2].tobytes()
        with pytest.raises(rmcl_b200.B2Error):
            h.correctOnceWait()                                        # nothing pending
    assert np.abs(sync[0]["t"] - want[0]["t"]).max() <= 2e-6
    # a queue of calls with different poses: results come back in order; the 9th in flight is refused
    h.setExecMode(2)
    Toms = [synth.compose(Tom, synth.make_transform((0.01 * k, 0, 0), (0, 0, 0.001 * k))) for k in range(8)]
    wants = [h.correctOnce(T, Tbo, 5, 0.0) for T in Toms]
    for T in Toms:
        h.correctOnceAsync(T, Tbo, 5, 0.0)
    with pytest.raises(rmcl_b200.B2Error):
        h.correctOnceAsync(Tom, Tbo, 5, 0.0)
    for w in wants:
        g = h.correctOnceWait()
        assert g[0].tobytes() == w[0].tobytes() and g[2].tobytes() == w[2].tobytes()


def test_sim_options_gpu(po, synth):
    """b2_rcc_set_sim_options: the three open rmagine simulate() semantics on the device against the oracle, all 8 combinations,
    find and correctOnce; the v1 batch entry follows the same switches."""
    name = "building:200000"
    osc = oracle_scene(name)
    m = synth.SphericalModel(np.radians(-25.0), np.radians(40.0) / 31, 32, -np.pi, 2 * np.pi / 512, 512, 2.0, 6.0)
    o, d = po.model_rays(m)
    Tbm, Tsb = synth.building_gt_pose(), synth.scenario_tsb()
    h = _rcc(synth, name, m)
    for opts in range(8):
        tf, mn, fill = opts & 1, (opts >> 1) & 1, (opts >> 2) & 1
        h.setSimOptions(tf, mn, fill)
        h.find(Tbm)
        mv = h.modelView()
        ref = osc.simulate(Tbm, Tsb, o, d, m.range_max, m.range_min, tfar_mode=tf, min_mode=mn, miss_fill=fill)
        for k in ref:
            assert np.array_equal(mv[k], ref[k], equal_nan=True), (opts, k)
    with pytest.raises(Exception):
        h.setSimOptions(2, 0, 0)
    # v1 batch: Ncorr follows the switches (tfar = inf admits pairs whose model point lies beyond range.max)
    h.setSimOptions(0, 0, 0)
    ranges = osc.simulate(Tbm, Tsb, o, d, 50.0)["ranges"]
    h.setInputData(np.minimum(ranges, 5.9).astype(np.float32))
    T = synth.transforms(3); T[:] = Tbm
    n0 = h.correct(T)[1]
    h.setSimOptions(1, 0, 0)
    n1 = h.correct(T)[1]
    assert (n1 >= n0).all() and n1.sum() > n0.sum()


def test_gladiator_p2p_local_world(po, synth):
    """Sharded Gladiator resampling over peer memory (b2_pf_resample_gladiator_p2p) with four shards that all live on this one GPU: the
    concatenated champions equal the single-GPU resampling bit for bit, and the byte counter equals what the draws imply (4 B per remote
    opponent likelihood + 68 B per remote winner) -- the real multi-GPU run (bench.py --gpus N, scripts/check_multigpu.py) uses the same kernel
    with CUDA-IPC peer pointers."""
    import torch
    import rmcl_b200
    from test_oracle import _glad_particles
    world, n = 4, 40000
    P, A = _glad_particles(synth, world * n)
    cfg = rmcl_b200.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    up = rmcl_b200.PCDSensorUpdaterB200(gpu_map("cube29"))
    Pd = torch.from_numpy(P.view(np.float32).reshape(-1, 8).copy()).cuda()
    Ad = torch.from_numpy(A.view(np.float32).reshape(-1, 9).copy()).cuda()
    Pn, An = torch.empty_like(Pd), torch.empty_like(Ad)
    up.resample(Pd, Ad, Pn, An, cfg, seed=77, step=3)
    torch.cuda.synchronize()
    raw, _ = up.gladiatorRandoms(77, 3, 0, world * n)
    raw = raw.cpu().numpy().view(np.uint32).astype(np.int64)
    L = A["likelihood"]["mean"]
    enemy = raw % (world * n)
    total = 0
    for r in range(world):
        pr, ar, traffic = up.resampleP2PLocalWorld(Pd, Ad, world, r, cfg, seed=77, step=3)
        torch.cuda.synchronize()
        assert torch.equal(pr, Pn[r * n:(r + 1) * n]) and torch.equal(ar, An[r * n:(r + 1) * n])
        e = enemy[r * n:(r + 1) * n]
        remote = (e // n) != r
        wins = L[e] > L[r * n:(r + 1) * n]
        assert traffic == 4 * int(remote.sum()) + 68 * int((remote & wins).sum())
        total += traffic
    allgather = world * (world - 1) * n * 68          # what the all-gather variant moves (every rank receives the other ranks' shards)
    assert total < 0.6 * allgather
