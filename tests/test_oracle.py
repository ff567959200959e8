"""CPU tests: the oracle against closed-form geometry, numpy restatements and the committed golden vectors."""
import math
import os

import numpy as np
import pytest

from conftest import mesh, oracle_scene, quat_close, random_rays

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_layouts(po, synth):
    assert po.TRANSFORM.itemsize == 32 and po.CROSS_STATS.itemsize == 64 and po.PARTICLE_ATTR.itemsize == 36 and po.RANGE_MEAS.itemsize == 64
    assert synth.TRANSFORM_DTYPE == po.TRANSFORM and synth.CROSS_STATS_DTYPE == po.CROSS_STATS


def test_spherical_dirs_formula(po, synth):
    m = synth.c1_sensor()
    d = po.spherical_dirs(m)
    phi = (np.float32(m.phi_min) + np.arange(m.phi_size, dtype=np.float32) * np.float32(m.phi_inc)).astype(np.float64)
    th = (np.float32(m.theta_min) + np.arange(m.theta_size, dtype=np.float32) * np.float32(m.theta_inc)).astype(np.float64)
    ref = np.stack([np.cos(phi)[:, None] * np.cos(th)[None, :], np.cos(phi)[:, None] * np.sin(th)[None, :], np.repeat(np.sin(phi)[:, None], len(th), 1)], -1)
    assert np.abs(d.reshape(m.phi_size, m.theta_size, 3) - ref).max() < 2e-7          # buffer id = vid*W + hid
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-6


def test_pinhole_dirs_formula(po, synth):
    m = synth.PinholeModel(64, 48, 52.5, 52.5, 31.5, 23.5, 0.3, 10.0)
    d = po.pinhole_dirs(m).reshape(48, 64, 3)
    v, h = np.meshgrid(np.arange(48), np.arange(64), indexing="ij")
    opt = np.stack([(h - m.cx) / m.fx, (v - m.cy) / m.fy, np.ones_like(h, float)], -1)
    opt /= np.linalg.norm(opt, axis=-1, keepdims=True)
    ref = np.stack([opt[..., 2], -opt[..., 0], -opt[..., 1]], -1)
    assert np.abs(d - ref).max() < 2e-7


def test_cube_closed_form_and_golden(po, synth):
    g = np.load(os.path.join(GOLD, "c1_cube.npz"))
    sc = oracle_scene("cube29")
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    assert np.array_equal(d, g["dirs"])
    sim = sc.simulate(g["Tgt"], g["Tsb"], o, d, m.range_max)
    # regression against the committed vectors: bit-exact
    for k in ("ranges", "hits", "face_ids", "points", "normals"):
        assert np.array_equal(sim[k], g[k], equal_nan=True), k
    hit = sim["hits"] > 0
    assert hit.mean() > 0.99
    # closed form
    assert np.abs(sim["ranges"][hit] - g["analytic_ranges"][hit]).max() < 2e-5
    assert np.abs(sim["normals"][hit] - g["analytic_normals"][hit]).max() < 1e-6
    assert np.abs(sim["points"][hit] - d[hit] * sim["ranges"][hit, None]).max() == 0.0      # point = dir*t (+0)
    # misses are NaN-encoded
    if (~hit).any():
        assert np.isnan(sim["points"][~hit]).all() and np.isnan(sim["normals"][~hit]).all()
        assert (sim["face_ids"][~hit] == 0xFFFFFFFF).all()


@pytest.mark.parametrize("name,lo,hi", [("cube29", -9.5, 9.5), ("uvsphere:40:60", -6.0, 6.0)])
def test_bvh_equals_brute_force(po, name, lo, hi):
    sc = oracle_scene(name)
    o, d = random_rays(4000, lo, hi, seed=1)
    t1, f1, n1, h1 = sc.intersect(o, d, brute=False)
    t2, f2, n2, h2 = sc.intersect(o, d, brute=True)
    assert np.array_equal(t1, t2) and np.array_equal(f1, f2) and np.array_equal(h1, h2) and np.array_equal(n1, n2)
    # axis-aligned and grid-vertex-aimed rays (ties on shared edges / vertices)
    V, _ = mesh(name)
    rng = np.random.default_rng(3)
    gv = V[rng.integers(0, len(V), 1500)]
    o0 = np.zeros_like(gv)
    dv = gv / np.linalg.norm(gv, axis=1, keepdims=True)
    ax = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 500)] * rng.choice([-1.0, 1.0], (500, 1)).astype(np.float32)
    o1 = rng.uniform(lo, hi, (500, 3)).astype(np.float32)
    for oo, dd in ((o0, dv), (o1, ax)):
        a = sc.intersect(oo, dd, brute=False)
        b = sc.intersect(oo, dd, brute=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_tfar_and_tie_rule(po):
    # two coincident triangles: the smaller face id wins; tfar cuts hits
    V = np.array([[0, -1, -1], [0, 1, -1], [0, 0, 1]], np.float32) + np.array([5, 0, 0], np.float32)
    V2 = np.concatenate([V, V])
    F = np.array([[3, 4, 5], [0, 1, 2]], np.uint32)
    sc = po.Scene(V2, F)
    t, f, ng, h = sc.intersect([[0, 0, 0]], [[1, 0, 0]])
    assert h[0] == 1 and f[0] == 0 and abs(t[0] - 5.0) < 1e-6
    assert np.allclose(ng[0], [4.0, 0, 0])                       # raw Ng = (v1-v0)x(v2-v0), |Ng| = 2*area
    t, f, ng, h = sc.intersect([[0, 0, 0]], [[1, 0, 0]], tfar=4.9)
    assert h[0] == 0
    t, f, ng, h = sc.intersect([[0, 0, 0]], [[-1, 0, 0]])
    assert h[0] == 0                                             # t > 0 only


def test_p2l_matches_numpy_and_golden(po, synth):
    g = np.load(os.path.join(GOLD, "c1_cube.npz"))
    sc = oracle_scene("cube29")
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    model = sc.simulate(g["Tguess"], g["Tsb"], o, d, m.range_max)
    I = synth.make_transform()
    s32 = po.statistics_p2l(I, g["dataset_points"], g["dataset_mask"], model["points"], model["normals"], model["hits"], 1.0)
    s64 = po.statistics_p2l(I, g["dataset_points"], g["dataset_mask"], model["points"], model["normals"], model["hits"], 1.0, f64=True)
    assert s32.tobytes() == g["stats_f32"].tobytes() and s64.tobytes() == g["stats_f64"].tobytes()
    # numpy restatement of the formula witnessed at rmcl_ros/src/micpl/MICPSensorCPU.cpp:71-98
    D = g["dataset_points"].astype(np.float64)
    Ii, Ni = model["points"].astype(np.float64), model["normals"].astype(np.float64)
    ok = (g["dataset_mask"] > 0) & (model["hits"] > 0)
    sd = np.where(ok, ((Ii - D) * Ni).sum(1), np.inf)
    acc = np.abs(sd) < 1.0
    M = D[acc] + Ni[acc] * sd[acc, None]
    Dm, Mm = D[acc].mean(0), M.mean(0)
    Cov = (M - Mm).T @ (D[acc] - Dm) / acc.sum()
    assert int(s64["n_meas"]) == int(acc.sum()) == int(s32["n_meas"])
    assert np.abs(s64["dataset_mean"] - Dm).max() < 1e-6 and np.abs(s64["model_mean"] - Mm).max() < 1e-6
    assert np.abs(s64["covariance"].reshape(3, 3).T - Cov).max() < 2e-5
    # FP32 sequential merges (reference arithmetic) agree with the FP64 sum form to FP32 noise
    assert np.abs(s32["dataset_mean"] - s64["dataset_mean"]).max() < 1e-4
    assert np.abs(s32["covariance"] - s64["covariance"]).max() < 2e-3


def test_cross_stats_merge_and_transform(po, synth):
    rng = np.random.default_rng(0)
    P = rng.normal(size=(300, 3)) * 3
    Q = P + rng.normal(size=(300, 3)) * 0.1

    def stats(p, q):
        s = np.zeros((), po.CROSS_STATS)
        s["dataset_mean"], s["model_mean"], s["n_meas"] = p.mean(0), q.mean(0), len(p)
        s["covariance"] = ((q - q.mean(0)).T @ (p - p.mean(0)) / len(p)).T.reshape(-1)
        return s
    a, b, full = stats(P[:100], Q[:100]), stats(P[100:], Q[100:]), stats(P, Q)
    mrg = po.cross_stats_merge(a, b)
    assert mrg["n_meas"] == 300
    assert np.abs(mrg["dataset_mean"] - full["dataset_mean"]).max() < 1e-5
    assert np.abs(mrg["covariance"] - full["covariance"]).max() < 1e-4
    ident = np.zeros((), po.CROSS_STATS)
    assert po.cross_stats_merge(ident, a).tobytes() == a.tobytes()          # Identity is neutral (micp_localization.cpp:918)
    T = synth.make_transform((1, 2, 3), (0.1, -0.2, 0.7))
    tr = po.cross_stats_transform(T, full)
    q = np.asarray(T["R"], np.float64)
    Pt, Qt = synth._qrot(q[None], P) + T["t"], synth._qrot(q[None], Q) + T["t"]
    ref = stats(Pt, Qt)
    assert np.abs(tr["dataset_mean"] - ref["dataset_mean"]).max() < 1e-5 and np.abs(tr["covariance"] - ref["covariance"]).max() < 1e-4


def test_umeyama_known_answers(po):
    g = np.load(os.path.join(GOLD, "umeyama.npz"))
    for s, T, truth in zip(g["stats"], g["T"], g["truth"]):
        out = po.umeyama(s)
        assert out.tobytes() == T.tobytes()                                   # regression
        assert quat_close(out["R"], truth[:4], 2e-6) and np.abs(out["t"] - truth[4:]).max() < 2e-5   # exact rigid motion recovered
    z = np.zeros((), po.CROSS_STATS)
    I = po.umeyama(z)
    assert np.allclose(I["R"], [0, 0, 0, 1]) and np.allclose(I["t"], 0)       # n_meas == 0 -> identity


def test_umeyama_reflection_case(po):
    # covariance with det < 0 (noise-dominated planar set): result must still be a proper rotation
    s = np.zeros((), po.CROSS_STATS)
    C = np.diag([2.0, 1.0, -0.01])
    s["covariance"] = C.T.reshape(-1)
    s["n_meas"] = 10
    T = po.umeyama(s)
    q = np.asarray(T["R"], np.float64)
    assert abs(np.linalg.norm(q) - 1) < 1e-6
    assert quat_close(q, [0, 0, 0, 1], 1e-6)


def test_adaptive_max_dist(po):
    assert po.adaptive_max_dist(1.0, 0.15, 0.0) == pytest.approx(1.0)
    assert po.adaptive_max_dist(1.0, 0.15, 1.0) == pytest.approx(0.15)
    assert po.adaptive_max_dist(1.0, 0.15, 0.5) == pytest.approx(0.575)


def test_micp_correct_once_converges_and_golden(po, synth):
    g = np.load(os.path.join(GOLD, "c1_cube.npz"))
    sc = oracle_scene("cube29")
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    I = synth.make_transform()
    Tn, Td, Cm = sc.micp_correct_once(o, d, m.range_max, g["dataset_points"], g["dataset_mask"], g["Tguess"], I, g["Tsb"], f64_accum=True)
    assert Tn.tobytes() == g["Tom_new"].tobytes() and Cm.tobytes() == g["Cmerged"].tobytes()
    Tom = g["Tguess"]
    for _ in range(25):
        Tom, _, _ = sc.micp_correct_once(o, d, m.range_max, g["dataset_points"], g["dataset_mask"], Tom, I, g["Tsb"], f64_accum=True)
    assert np.abs(Tom["t"] - g["Tgt"]["t"]).max() < 0.01 and quat_close(Tom["R"], g["Tgt"]["R"], 2e-3)


def test_legacy_benchmark_scenario(po, synth):
    """lidar_corrector_embree_benchmark.cpp:84-135: sphere map, vlp16_900 with range.min=0, T_curr = I with z+0.2, 10 x correct()."""
    V, F = mesh("uvsphere:40:60")
    sc = po.Scene(V, F)
    m = synth.vlp16_900()
    m.range_min = 0.0
    o, d = po.model_rays(m)
    I = synth.make_transform()
    ranges = sc.simulate(I, I, o, d, m.range_max)["ranges"]
    T = synth.transforms(3)
    T["t"][:, 2] = 0.2
    z_prev = 0.2
    for _ in range(10):
        Td, nc, _ = sc.correct_batch(T, I, o, d, m.range_min, m.range_max, ranges, max_dist=1.0, f64_accum=True)
        assert (nc > 14000).all()
        T = np.array([po.transform_mul(T[i], Td[i]) for i in range(len(T))])
        z = float(np.abs(T["t"][:, 2]).max())
        assert z < z_prev                                            # point-to-plane slides along the equator band: slow but monotone
        z_prev = z
    assert z_prev < 0.17 and np.abs(T["t"][:, :2]).max() < 1e-3


def test_gaussian_and_pf_penalties(po, synth):
    g = np.load(os.path.join(GOLD, "pf_cube.npz"))
    sc = oracle_scene("cube29")
    for ng_mode in (0, 1):
        out = sc.pf_update(g["poses"], g["attrs0"], g["Tsb"], g["beams"], po.PFParams.defaults(ng_mode))
        assert out.tobytes() == g[f"attrs_ng{ng_mode}"].tobytes()
        assert (out["likelihood"]["n_meas"] == 24).all()
        assert (out["likelihood"]["mean"] <= 1.0 / math.sqrt(2 * 4.0 * math.pi) + 1e-7).all()
    # a beam that is out of sensor range in reality but hits in simulation -> penalty 100 m -> eval underflows to 0
    b = g["beams"][:1].copy()
    b["range"] = 500.0
    out = sc.pf_update(g["poses"][:4], g["attrs0"][:4], g["Tsb"], b, po.PFParams.defaults())
    assert (out["likelihood"]["mean"] == 0.0).all() and (out["likelihood"]["n_meas"] == 1).all()
    # n_meas clamps at MAX_N_MEAS = 10000 (ParticleAttributes.hpp:34)
    a = g["attrs0"][:2].copy()
    a["likelihood"]["n_meas"] = 9999
    out = sc.pf_update(g["poses"][:2], a, g["Tsb"], g["beams"][:5], po.PFParams.defaults())
    assert (out["likelihood"]["n_meas"] == 10000).all()


def test_dataset_from_ranges_mask(po, synth):
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    r = np.full(m.size, 5.0, np.float32)
    r[0], r[1], r[2] = 0.01, 1000.0, m.range_max
    pts, mask, nv = po.dataset_from_ranges(o, d, r, m.range_min, m.range_max)
    assert mask[0] == 0 and mask[1] == 0 and mask[2] == 1 and nv == m.size - 2
    assert np.array_equal(pts[5], d[5] * np.float32(5.0))


def test_pf_motion_and_stats(po, synth):
    """particle_move_and_forget (particle_motion.cu:11-34) and compute_stats (resampling.cu:41-92) restated."""
    P, A = synth.pf_particles(3000)
    rng = np.random.default_rng(1)
    A["likelihood"]["mean"] = rng.uniform(0, 0.2, len(A)).astype(np.float32)
    A["likelihood"]["n_meas"] = rng.integers(0, 10001, len(A)).astype(np.uint32)
    T = synth.make_transform((0.1, -0.02, 0.0), (0, 0, 0.05))
    P2, A2 = po.pf_motion_update(P, A, T, 0.03)
    ref = synth.compose(P, np.broadcast_to(T, P.shape))
    assert np.abs(P2["t"] - ref["t"]).max() < 1e-5 and np.abs(np.abs((P2["R"] * ref["R"]).sum(1)) - 1).max() < 1e-6
    nm = A["likelihood"]["n_meas"].astype(np.float64)
    assert np.array_equal(A2["likelihood"]["n_meas"], (nm - 0.03 * nm).astype(np.uint32))       # uint -= double truncates (quirk D3)
    assert np.array_equal(A2["likelihood"]["mean"], A["likelihood"]["mean"])
    s, m = po.pf_likelihood_stats(A)
    assert m == A["likelihood"]["mean"].max() and abs(s - A["likelihood"]["mean"].astype(np.float64).sum()) < 1e-3
    s0, m0 = po.pf_likelihood_stats(A[:0])
    assert s0 == 0.0 and m0 == 0.0                                                              # max starts at 0 (resampling.cu:54)


def test_closest_point_oracle(po, synth):
    """closest-point correspondences (CPCEmbree.cpp:17-43) restated; BVH walk == brute force, closed forms on the cube."""
    I = synth.make_transform()
    rng = np.random.default_rng(3)
    for name, lo, hi in (("cube29", [-12] * 3, [12] * 3), ("building:60000", [-2, -2, -1], [62, 42, 4])):
        sc = oracle_scene(name)
        q = rng.uniform(lo, hi, (3000, 3)).astype(np.float32)
        a, b = sc.cpc_find(I, I, q, 1.0, brute=False), sc.cpc_find(I, I, q, 1.0, brute=True)
        assert all(np.array_equal(a[k], b[k]) for k in a)
    sc = oracle_scene("cube29")
    q = rng.uniform(-9.9, 9.9, (2000, 3)).astype(np.float32)
    r = sc.cpc_find(I, I, q, 1.0)
    d_true = 10.0 - np.abs(q).max(1)                                     # inside the cube (half extent 10): distance to the nearest face
    assert np.abs(r["dists"] - d_true).max() <= 2e-6
    assert np.array_equal(r["hits"], (r["dists"] <= 1.0).astype(np.uint8))
    ax = np.abs(q).argmax(1)
    sign = np.sign(q[np.arange(len(q)), ax])
    assert np.allclose(np.abs(r["normals"][np.arange(len(q)), ax]), 1.0, atol=1e-6)             # the face normal, axis aligned
    assert np.allclose(r["points"][np.arange(len(q)), ax], 10.0 * sign, atol=1e-5)
    # sensor-frame output: a rigid change of frame must not change distances, and points come back in the sensor frame
    Tbm, Tsb = synth.make_transform([0.4, -0.3, 0.2], [0.1, -0.05, 0.6]), synth.make_transform([0.1, 0, 0.3], [0, 0.02, 0])
    Tsm = synth.compose(Tbm, Tsb)
    qs = synth.transform_points(synth.inverse(Tsm), q)                  # the same map-frame queries expressed in the sensor frame
    r2 = sc.cpc_find(Tbm, Tsb, qs, 1.0)
    assert np.abs(r2["dists"] - d_true).max() <= 2e-5
    assert np.abs(synth.transform_points(Tsm, r2["points"]) - r["points"]).max() <= 5e-5
    # tie rule (d2, face id), surface point, far point, non-finite query
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    F = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    s2 = po.Scene(V, F)
    t = s2.cpc_find(I, I, np.array([[0.5, 0.5, 0.25], [0.25, 0.5, 0.0], [5, 5, 5], [np.nan, 0, 0]], np.float32), 1.0)
    assert t["face_ids"][0] == 0 and abs(t["dists"][0] - 0.25) < 1e-7 and t["hits"][0] == 1    # on the shared diagonal: lower face id wins
    assert t["face_ids"][1] == 1 and t["dists"][1] == 0.0 and t["hits"][1] == 1
    assert t["hits"][2] == 0 and t["face_ids"][2] == 0 and abs(t["dists"][2] - np.sqrt(16 + 16 + 25)) < 1e-5
    assert t["hits"][3] == 0 and t["face_ids"][3] == 0xFFFFFFFF and np.isnan(t["points"][3]).all()


def test_cpc_correct_once_converges(po, synth):
    """ICP with closest-point correspondences pulls a perturbed scan back onto the map (same inner loop as MICP-L, CPC find)."""
    sc = oracle_scene("cube29")
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    Tgt, Tsb = synth.make_transform([0.3, -0.2, 0.1], [0, 0, 0.3]), synth.make_transform([0.0, 0, 0.2], [0, 0, 0])
    sim = sc.simulate(Tgt, Tsb, o, d, m.range_max)
    dp, dm, _ = po.dataset_from_ranges(o, d, sim["ranges"], m.range_min, m.range_max)
    I = synth.make_transform()
    Tom = synth.compose(Tgt, synth.make_transform([0.15, -0.1, 0.05], [0.01, -0.01, 0.04]))
    e0 = np.abs(Tom["t"] - Tgt["t"]).max()
    for _ in range(6):
        Tom, Td, Cm = sc.micp_correct_once(None, None, m.range_max, dp, dm, Tom, I, Tsb, 5, 1.0, 0.15, 0.0, f64_accum=True)
    assert Cm["n_meas"] > 0.9 * dm.sum()
    assert np.abs(Tom["t"] - Tgt["t"]).max() < 0.2 * e0


def _glad_particles(synth, n, seed=0):
    P, A = synth.pf_particles(n)
    rng = np.random.default_rng(seed)
    A["likelihood"]["mean"] = rng.uniform(0, 0.2, n).astype(np.float32)
    A["likelihood"]["n_meas"] = rng.integers(0, 10001, n).astype(np.uint32)
    A["state_sigma"] = rng.uniform(0, 1, (n, 6)).astype(np.float32)
    P["R"] = np.stack([synth.quat_from_rpy(*r) for r in rng.uniform(-0.3, 0.3, (n, 3))]).astype(np.float32)
    P["stamp"] = np.arange(n)
    return P, A


def test_philox_known_answers(po):
    """Philox4x32-10 against the published Random123 known-answer vectors (kat_vectors: zeros, all ones, digits of pi)."""
    assert [hex(x) for x in po.philox4x32_10([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in po.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in po.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    raw, nrm = po.pf_gladiator_randoms(1234, 0, 0, 200000)
    assert np.abs(nrm.mean(0)).max() < 0.01 and np.abs(nrm.std(0) - 1).max() < 0.01 and np.isfinite(nrm).all()
    assert np.abs(np.corrcoef(nrm.T) - np.eye(6)).max() < 0.01
    assert abs((raw % 7 == 0).mean() - 1 / 7) < 0.005
    # keyed by the GLOBAL index: a shard's draws are a slice of the whole
    r2, n2 = po.pf_gladiator_randoms(1234, 0, 5000, 1000)
    assert np.array_equal(r2, raw[5000:6000]) and np.array_equal(n2, nrm[5000:6000])
    assert not np.array_equal(po.pf_gladiator_randoms(1234, 1, 0, 100)[0], raw[:100])


def test_gladiator_resample_oracle(po, synth):
    """gladiator_resample_kernel (resampling.cu:108-199) restated as a pure function of (particles, draws)."""
    n = 20000
    P, A = _glad_particles(synth, n)
    cfg = po.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    raw, nrm = po.pf_gladiator_randoms(1234, 3, 0, n)
    Pn, An = po.pf_gladiator_resample(P, A, 0, n, raw, nrm, cfg)
    enemy = raw % n
    won = A["likelihood"]["mean"][enemy] > A["likelihood"]["mean"]
    assert 0.4 < won.mean() < 0.6
    assert Pn[~won].tobytes() == P[~won].tobytes() and An[~won].tobytes() == A[~won].tobytes()          # champion stays champion (:193-196)
    assert np.array_equal(An["likelihood"]["mean"][won], A["likelihood"]["mean"][enemy][won])           # the likelihood is kept (:177)
    assert np.array_equal(An["state_sigma"][won], A["state_sigma"][enemy][won]) and np.array_equal(Pn["stamp"][won], P["stamp"][enemy][won])
    dt = Pn["t"][won] - P["t"][enemy][won]
    assert np.abs(dt - nrm[won][:, :3] * np.array([0.03, 0.03, 0.01], np.float32)).max() < 1e-5            # :166-168
    assert np.abs(np.linalg.norm(Pn["R"][won].astype(np.float64), axis=1) - 1).max() < 1e-6
    # forget rule (:178-187): rot_dist is the quaternion 4-norm (~1), so forget_rate >= likelihood_forget_per_radian
    tr = np.linalg.norm(dt.astype(np.float64), axis=1)
    forget = np.maximum(1 - 0.7 ** tr, 0.2)
    expect = A["likelihood"]["n_meas"][enemy][won] * (1 - forget)
    assert np.abs(An["likelihood"]["n_meas"][won] - expect).max() <= 1.01
    # sharding invariance: two halves with global indices == the whole
    h = n // 2
    a = po.pf_gladiator_resample(P, A, 0, h, raw[:h], nrm[:h], cfg)
    b = po.pf_gladiator_resample(P, A, h, n - h, raw[h:], nrm[h:], cfg)
    assert np.concatenate([a[0], b[0]]).tobytes() == Pn.tobytes() and np.concatenate([a[1], b[1]]).tobytes() == An.tobytes()


def _segmentation_case(po, synth, name="cube29"):
    """A scan of the map with an extra obstacle in front of a wall (scan outliers), a range pushed behind the wall (map outliers),
    dropped returns and out-of-range values."""
    sc = oracle_scene(name)
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    T, Tsb = synth.make_transform([0.5, -0.3, 0.2], [0, 0, 0.4]), synth.make_transform()
    sim = sc.simulate(T, Tsb, o, d, m.range_max)
    real = sim["ranges"].copy()
    rng = np.random.default_rng(2)
    real += rng.normal(0, 0.01, len(real)).astype(np.float32)
    k = rng.permutation(len(real))
    real[k[:100]] *= 0.6                  # something in front of the surface
    real[k[100:200]] *= 1.3               # the ray cut the surface
    real[k[200:230]] = m.range_max + 1    # no return
    real[k[230:240]] = 0.0                # below range.min
    return sc, m, o, d, T, Tsb, sim, real


def test_segmentation_oracle(po, synth):
    """scan_map_segmentation_embree.cpp:110-187 restated; labels against an independent numpy evaluation."""
    sc, m, o, d, T, Tsb, sim, real = _segmentation_case(po, synth)
    a, b, lab = po.segment(o, d, m.range_min, m.range_max, real, sim["ranges"], sim["normals"], 0.15, 0.15)
    rv = (real >= m.range_min) & (real <= m.range_max)
    sv = (sim["ranges"] >= m.range_min) & (sim["ranges"] <= m.range_max)
    n = sim["normals"].astype(np.float64)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    preal, pint = d.astype(np.float64) * real[:, None] + o, d.astype(np.float64) * sim["ranges"][:, None]
    pd = np.abs(((preal - pint) * n).sum(1))
    expect = np.zeros(len(real), np.uint8)
    both = rv & sv
    margin = np.abs(pd - 0.15) > 1e-4                                   # away from the threshold the float and double evaluations agree
    expect[both & (real < sim["ranges"]) & (pd > 0.15)] = 1
    expect[both & ~(real < sim["ranges"]) & (pd > 0.15)] = 2
    expect[rv & ~sv] = 1
    expect[~rv & sv] = 2
    assert np.array_equal(lab[margin | ~both], expect[margin | ~both])
    assert (lab == 1).sum() == len(a) >= 100 and (lab == 2).sum() == len(b) >= 100
    assert np.allclose(a, preal[lab == 1], atol=1e-5)                     # raster order
    far = (lab == 2) & rv
    assert np.allclose(b[(lab == 2).nonzero()[0].searchsorted(far.nonzero()[0])], pint[far], atol=1e-5)


def _motion_case(synth, n=4000):
    """Particles inside the 20 m cube; a 1.5 m forward step pushes those facing a nearby wall through it."""
    rng = np.random.default_rng(8)
    P, A = synth.pf_particles(n, footprint=(19.0, 19.0), z=0.0, margin=0.0)
    P["t"][:, :2] -= 9.5
    A["likelihood"]["mean"] = rng.uniform(0.01, 0.2, n).astype(np.float32)
    A["likelihood"]["sigma"] = 0.5
    A["likelihood"]["n_meas"] = rng.integers(1, 10001, n).astype(np.uint32)
    return P, A, synth.make_transform((1.5, 0.0, 0.0), (0, 0, 0.05))


def test_motion_update_collision_oracle(po, synth):
    """TFMotionUpdaterCPU wall check (TFMotionUpdaterCPU.cpp:17-50,205-216): crossing a cube face <=> the new position is outside."""
    sc = oracle_scene("cube29")
    P, A, T = _motion_case(synth)
    P2, A2 = po.pf_motion_update(P, A, T, 0.03, scene=sc)
    P0, A0 = po.pf_motion_update(P, A, T, 0.03)
    assert P2.tobytes() == P0.tobytes()
    outside = np.abs(P2["t"][:, :2]).max(1) > 10.0
    assert 0.02 < outside.mean() < 0.5
    hit = (A2["likelihood"]["mean"] == 0) & (A2["likelihood"]["sigma"] == 0) & (A2["likelihood"]["n_meas"] == 10000)
    assert np.array_equal(hit, outside)
    assert A2[~hit].tobytes() == A0[~hit].tobytes()
    # no motion -> no ray (length < 1e-5)
    P3, A3 = po.pf_motion_update(P, A, synth.make_transform(), 0.0, scene=sc)
    assert A3.tobytes() == A.tobytes()


def test_widened_rows_golden(po, synth):
    """Regression fixtures of the SURVEY 8(f) rows (tests/golden/f_rows.npz, generator tests/golden/make_golden.py:widened)."""
    g = np.load(os.path.join(GOLD, "f_rows.npz"))
    sc = oracle_scene("cube29")
    m = synth.c1_sensor()
    o, d = po.model_rays(m)
    cpc = sc.cpc_find(g["Tgt"], g["Tsb"], g["queries"], 0.8)
    assert np.array_equal(cpc["face_ids"], g["cpc_faces"]) and np.array_equal(cpc["dists"], g["cpc_dists"]) and np.array_equal(cpc["hits"], g["cpc_hits"])
    assert np.array_equal(cpc["points"], g["cpc_points"], equal_nan=True) and np.array_equal(cpc["normals"], g["cpc_normals"], equal_nan=True)
    A1 = sc.pf_update(g["poses"], g["attrs0"], g["Tsb"], g["beams"], po.PFParams.defaults(0, 1))
    assert np.array_equal(A1["likelihood"]["mean"], g["attrs_cpc"]["likelihood"]["mean"])
    Pm, Am = po.pf_motion_update(g["poses"], g["attrs_cpc"], g["T_motion"], 0.03, scene=sc)
    assert Pm.tobytes() == g["poses_moved"].tobytes() and Am.tobytes() == g["attrs_moved"].tobytes()
    raw, nrm = po.pf_gladiator_randoms(1234, 3, 0, len(Pm))
    assert np.array_equal(raw, g["glad_raw"]) and np.array_equal(nrm, g["glad_normals"])
    cfg = po.GladiatorConfig(0.03, 0.03, 0.01, 0.002, 0.002, 0.01, 0.3, 0.2)
    Pr, Ar = po.pf_gladiator_resample(Pm, Am, 0, len(Pm), raw, nrm, cfg)
    assert Pr.tobytes() == g["poses_resampled"].tobytes() and Ar.tobytes() == g["attrs_resampled"].tobytes()
    sim = sc.simulate(g["Tgt"], g["Tsb"], o, d, m.range_max)
    a, b, lab = po.segment(o, d, m.range_min, m.range_max, g["real_ranges"], sim["ranges"], sim["normals"], 0.15, 0.1)
    assert np.array_equal(lab, g["seg_labels"]) and np.array_equal(a, g["seg_scan"]) and np.array_equal(b, g["seg_map"])
