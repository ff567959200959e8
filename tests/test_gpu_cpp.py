"""GPU test of the C++ drop-in classes (include/rmcl_b200/rcc_b200.hpp): the legacy benchmark scenario compiled from
examples/cpp_dropin.cpp must reproduce the Python path and the oracle."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import mesh, oracle_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin_matches_oracle(po, synth):
    exe = os.path.join(ROOT, "examples", "cpp_dropin")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    V, F = mesh("uvsphere:40:60")
    with tempfile.NamedTemporaryFile(suffix=".mesh", delete=False) as f:
        np.array([len(V), len(F)], np.uint32).tofile(f)
        V.astype(np.float32).tofile(f)
        F.astype(np.uint32).tofile(f)
        path = f.name
    try:
        out = subprocess.run([exe, "40", "60", "16", path], capture_output=True, text=True, timeout=120)
    finally:
        os.unlink(path)
    assert out.returncode == 0, out.stderr
    lines = {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines() if l and not l.startswith("PF")}
    pf = [l.split() for l in out.stdout.splitlines() if l.startswith("PF")]
    # oracle for the same scenario
    osc = oracle_scene("uvsphere:40:60")
    # the same float32 model parameters the C++ example computes (-15.0f * (float)M_PI / 180.0f, ...)
    f, pi32 = np.float32, np.float32(np.pi)
    m = synth.SphericalModel(float(f(-15.0) * pi32 / f(180.0)), float(f(2.0) * pi32 / f(180.0)), 16, float(-pi32), float(f(2.0) * pi32 / f(900.0)), 900, 0.0, 130.0)
    o, d = po.model_rays(m)
    I = synth.make_transform()
    ranges = osc.simulate(I, I, o, d, m.range_max)["ranges"]
    T = synth.transforms(1)
    T["t"][:, 2] = 0.2
    Td, nc, _ = osc.correct_batch(T, I, o, d, m.range_min, m.range_max, ranges, 1.0, f64_accum=True)
    assert int(lines["RUN0"][1]) == int(nc[0])
    assert abs(float(lines["RUN0"][3]) - float(Td["t"][0, 2])) <= 1e-6
    z_final = float(lines["FINAL"][1])
    assert 0.0 < z_final < 0.2
    # v2 round
    dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    model = osc.simulate(T[0], I, o, d, m.range_max)
    st = po.statistics_p2l(I, dp, dm, model["points"], model["normals"], model["hits"], 1.0, f64=True)
    assert int(lines["V2"][1]) == int(st["n_meas"])
    assert abs(float(lines["V2"][3]) - float(st["covariance"][[0, 4, 8]].sum())) <= 1e-4
    assert abs(float(lines["V2"][5]) - float(po.umeyama(st)["t"][2])) <= 1e-6
    # the same round fed from rmcl_msgs-shaped Scan and O1Dn messages (rmcl_msgs_adapters.hpp): identical statistics
    assert lines["MSG"][1] == lines["V2"][1] and lines["MSG"][3] == lines["V2"][3]
    assert lines["MSG"][5] == lines["V2"][1] and lines["MSG"][7] == lines["V2"][3] and lines["MSG"][9] == lines["V2"][1]
    # particle update
    beams = np.zeros(8, synth.RANGE_MEAS_DTYPE)
    th = np.float32(0.7) * np.arange(8, dtype=np.float32)
    beams["dir"] = np.stack([np.cos(th), np.sin(th), np.zeros(8, np.float32)], 1)
    beams["range"] = 10.0
    P = synth.transforms(4)
    P["t"][1, 0], P["t"][2, 1], P["t"][3, 2] = 1.0, -2.0, 0.5
    A = np.zeros(4, synth.PARTICLE_ATTR_DTYPE)
    A["likelihood"]["mean"] = 1.0
    A["state_sigma"] = 0.1
    ref = osc.pf_update(P, A, I, beams, po.PFParams.defaults())
    assert len(pf) == 4
    for row, r in zip(pf, ref):
        assert int(row[7]) == int(r["likelihood"]["n_meas"])
        assert abs(float(row[3]) - float(r["likelihood"]["mean"])) <= 2e-6


def test_cpp_classes_behind_the_reference_interface(po, synth):
    """examples/cpp_dropin_rmagine: RCCB200Spherical used through std::shared_ptr<rmcl::Correspondences_<rm::VRAM_CUDA>> (the reference's
    unmodified header), dataset written through the public member, model buffers read through the inherited modelView(); PCDSensorUpdaterB200
    through rmcl::SensorUpdater<rm::VRAM_CUDA>.  Built in this repo's container (needs /root/reference); the binary travels to the GPU box."""
    exe = os.path.join(ROOT, "examples", "cpp_dropin_rmagine")
    if not os.path.exists(exe):
        pytest.skip("examples/cpp_dropin_rmagine not built (needs the reference headers: __graft_entry__.build() in the authoring container)")
    V, F = mesh("uvsphere:40:60")
    with tempfile.NamedTemporaryFile(suffix=".mesh", delete=False) as f:
        np.array([len(V), len(F)], np.uint32).tofile(f)
        V.astype(np.float32).tofile(f)
        F.astype(np.uint32).tofile(f)
        path = f.name
    try:
        out = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
    finally:
        os.unlink(path)
    assert out.returncode == 0, out.stderr
    rows = {l.split()[0]: l.split()[1:] for l in out.stdout.splitlines() if l and not l.startswith("PF")}
    pf = [l.split() for l in out.stdout.splitlines() if l.startswith("PF")]
    osc = oracle_scene("uvsphere:40:60")
    f32, pi32 = np.float32, np.float32(np.pi)
    m = synth.SphericalModel(float(f32(-15.0) * pi32 / f32(180.0)), float(f32(2.0) * pi32 / f32(180.0)), 16, float(-pi32), float(f32(2.0) * pi32 / f32(900.0)), 900, 0.0, 130.0)
    o, d = po.model_rays(m)
    I = synth.make_transform()
    ranges = osc.simulate(I, I, o, d, m.range_max)["ranges"]
    T = synth.transforms(1)
    T["t"][:, 2] = 0.2
    dp, dm, _ = po.dataset_from_ranges(o, d, ranges, m.range_min, m.range_max)
    model = osc.simulate(T[0], I, o, d, m.range_max)
    st = po.statistics_p2l(I, dp, dm, model["points"], model["normals"], model["hits"], 1.0, f64=True)
    v = rows["V2R"]
    assert int(v[1]) == int(st["n_meas"]) and int(v[7]) == m.size and int(v[9]) == int(model["hits"].sum()) and int(v[11]) == m.size
    assert abs(float(v[3]) - float(st["covariance"][[0, 4, 8]].sum())) <= 1e-4
    assert abs(float(v[5]) - float(po.umeyama(st)["t"][2])) <= 1e-6
    # member-dataset path == setter path of the same class, bit for bit
    assert rows["V2P"][1] == v[1] and rows["V2P"][3] == v[3]
    assert all(float(x) > 0 for x in (rows["BENCH"][1], rows["BENCH"][3], rows["BENCH"][5]))
    beams = np.zeros(8, synth.RANGE_MEAS_DTYPE)
    th = np.float32(0.7) * np.arange(8, dtype=np.float32)
    beams["dir"] = np.stack([np.cos(th), np.sin(th), np.zeros(8, np.float32)], 1)
    beams["range"] = 10.0
    P = synth.transforms(4)
    P["t"][1, 0], P["t"][2, 1], P["t"][3, 2] = 1.0, -2.0, 0.5
    A = np.zeros(4, synth.PARTICLE_ATTR_DTYPE)
    A["likelihood"]["mean"] = 1.0
    A["state_sigma"] = 0.1
    ref = osc.pf_update(P, A, I, beams, po.PFParams.defaults())
    assert len(pf) == 4
    for row, r in zip(pf, ref):
        assert int(row[7]) == int(r["likelihood"]["n_meas"]) and abs(float(row[3]) - float(r["likelihood"]["mean"])) <= 2e-6
