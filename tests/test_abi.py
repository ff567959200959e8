"""CPU tests: the C-ABI library loads and exports every symbol include/rmcl_b200.h declares; failure without a GPU is loud."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rmcl_b200.h")).read()
    return sorted(set(re.findall(r"B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("b2_mesh_create", "b2_rcc_find", "b2_rcc_cross_statistics", "b2_rcc_correct_once", "b2_rcc_correct_batch", "b2_umeyama_batch",
                 "b2_pf_sensor_update", "b2_pf_sensor_update_host", "b2_last_error"):
        assert must in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol():
    import rmcl_b200
    path = rmcl_b200.lib_path()
    assert os.path.exists(path), "librmcl_b200.so not built (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    assert sorted(rmcl_b200.api.EXPORTS) == _declared()


def test_pod_sizes_match_header():
    import rmcl_b200
    from rmcl_b200 import api, synth
    assert ctypes.sizeof(api.PFParams) == 32 and ctypes.sizeof(api.GladiatorConfig) == 32 and ctypes.sizeof(api._SphericalModel) == 32 and ctypes.sizeof(api._PinholeModel) == 32
    assert synth.TRANSFORM_DTYPE.itemsize == 32 and synth.CROSS_STATS_DTYPE.itemsize == 64


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import rmcl_b200
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.Map(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.uint32))


def test_product_does_not_reference_oracle():
    """The product tree must never import / include / link the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rmcl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "oracle.h" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower()


def _write_ply(path, V, F, binary):
    import struct
    with open(path, "wb") as f:
        f.write(b"ply\nformat " + (b"binary_little_endian" if binary else b"ascii") + b" 1.0\ncomment test\n")
        f.write(f"element vertex {len(V)}\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\n".encode())
        f.write(f"element face {len(F)}\nproperty list uchar int vertex_indices\nend_header\n".encode())
        for v in V:
            f.write(struct.pack("<fffB", *v, 7) if binary else f"{v[0]!r} {v[1]!r} {v[2]!r} 7\n".encode())
        for t in F:
            f.write(struct.pack("<B" + "i" * len(t), len(t), *t) if binary else (" ".join(map(str, [len(t), *t])) + "\n").encode())


def test_mesh_file_import(tmp_path):
    """rm::import_embree_map stand-in -- PLY (ascii, binary) and OBJ readers, polygons fan-triangulated, errors reported."""
    import rmcl_b200
    from rmcl_b200 import synth
    V, F = synth.cube(3)
    V = V.astype(np.float32)
    for binary in (False, True):
        p = str(tmp_path / f"m{int(binary)}.ply")
        _write_ply(p, [tuple(float(x) for x in v) for v in V], [tuple(int(i) for i in t) for t in F], binary)
        V2, F2 = rmcl_b200.read_mesh_file(p)
        assert np.array_equal(V2, V) and np.array_equal(F2, F)
    q = str(tmp_path / "quad.ply")
    _write_ply(q, [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0.5, 2, 0)], [(0, 1, 2, 3), (3, 2, 4)], False)
    Vq, Fq = rmcl_b200.read_mesh_file(q)
    assert Fq.tolist() == [[0, 1, 2], [0, 2, 3], [3, 2, 4]]
    o = str(tmp_path / "m.obj")
    with open(o, "w") as f:
        f.write("# test\n" + "".join(f"v {v[0]!r} {v[1]!r} {v[2]!r}\n" for v in V.tolist()) + "vn 0 0 1\n")
        f.write("".join(f"f {t[0] + 1}//1 {t[1] + 1}//1 {t[2] + 1}//1\n" for t in F.tolist()[:-1]))
        t = F.tolist()[-1]
        f.write(f"f {t[0] - len(V)} {t[1] - len(V)} {t[2] - len(V)}\n")               # negative (relative) indices
    Vo, Fo = rmcl_b200.read_mesh_file(o)
    assert np.array_equal(Vo, V) and np.array_equal(Fo, F)
    for bad, txt in (("x.dae", "<COLLADA/>"), ("x.stl", "solid"), ("empty.ply", "ply\nformat ascii 1.0\nelement vertex 0\nproperty float x\nproperty float y\nproperty float z\nend_header\n"),
                     ("oob.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")):
        bp = tmp_path / bad
        bp.write_text(txt)
        with pytest.raises(rmcl_b200.B2Error):
            rmcl_b200.read_mesh_file(str(bp))
    with pytest.raises(rmcl_b200.B2Error):
        rmcl_b200.read_mesh_file(str(tmp_path / "missing.ply"))


def _dae(geoms, scene_nodes, up="Z_UP"):
    """COLLADA 1.4.1 document: geoms = {id: (positions (n,3), primitive xml)}, scene_nodes = xml of the <visual_scene> body"""
    lib = ""
    for gid, (P, prim) in geoms.items():
        fa = " ".join(repr(float(x)) for x in np.asarray(P, np.float64).reshape(-1))
        lib += (f'<geometry id="{gid}" name="{gid}"><mesh><source id="{gid}-pos"><float_array id="{gid}-arr" count="{len(P) * 3}">{fa}</float_array>'
                f'<technique_common><accessor source="#{gid}-arr" count="{len(P)}" stride="3"><param name="X" type="float"/><param name="Y" type="float"/>'
                f'<param name="Z" type="float"/></accessor></technique_common></source>'
                f'<source id="{gid}-nrm"><float_array id="{gid}-narr" count="3">0 0 1</float_array></source>'
                f'<vertices id="{gid}-vtx"><input semantic="POSITION" source="#{gid}-pos"/></vertices>{prim.format(g=gid)}</mesh></geometry>')
    return (f'<?xml version="1.0" encoding="utf-8"?>\n<!-- test map -->\n<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">'
            f'<asset><unit name="meter" meter="1"/><up_axis>{up}</up_axis></asset><library_geometries>{lib}</library_geometries>'
            f'<library_visual_scenes><visual_scene id="Scene">{scene_nodes}</visual_scene></library_visual_scenes>'
            f'<scene><instance_visual_scene url="#Scene"/></scene></COLLADA>')


def test_collada_import(tmp_path):
    """the reference's example maps are COLLADA (docs/MICPL.md:46-49, loaded through assimp by rm::import_embree_map,
    micp_localization.cpp:188).  <triangles> / <polylist> / <polygons> with interleaved inputs, node transforms (matrix, translate, rotate,
    scale, nesting), two instances of one geometry, up_axis left alone."""
    import rmcl_b200
    from rmcl_b200 import synth
    V, F = synth.cube(2)
    tri = '<triangles count="%d"><input semantic="VERTEX" source="#{g}-vtx" offset="0"/><input semantic="NORMAL" source="#{g}-nrm" offset="1"/><p>%s</p></triangles>' % (
        len(F), " ".join(f"{i} 0" for i in F.reshape(-1)))
    quad = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0.5, 2, 0)]
    plist = '<polylist count="2"><input semantic="VERTEX" source="#{g}-vtx" offset="0"/><vcount>4 3</vcount><p>0 1 2 3 3 2 4</p></polylist>'
    pgons = '<polygons count="1"><input semantic="VERTEX" source="#{g}-vtx" offset="0"/><p>0 1 2 3</p></polygons>'
    nodes = ('<node id="a"><matrix>1 0 0 5  0 1 0 -2  0 0 1 0.5  0 0 0 1</matrix><instance_geometry url="#cube"/></node>'
             '<node id="b"><translate>10 0 0</translate><rotate>0 0 1 90</rotate><scale>2 2 2</scale><instance_geometry url="#quad"/>'
             '<node id="c"><translate>0 0 1</translate><instance_geometry url="#pg"/></node></node>'
             '<node id="d"><instance_geometry url="#cube"/></node>')
    p = tmp_path / "map.dae"
    p.write_text(_dae({"cube": (V, tri), "quad": (quad, plist), "pg": (quad[:4], pgons)}, nodes))
    V2, F2 = rmcl_b200.read_mesh_file(str(p))
    nv, nf = len(V), len(F)
    assert len(V2) == 2 * nv + 5 + 4 and len(F2) == 2 * nf + 3 + 2
    assert np.allclose(V2[:nv], V + np.float32([5, -2, 0.5])) and np.array_equal(F2[:nf], F)                    # <matrix>
    q = np.float32(quad)
    want = np.stack([-2 * q[:, 1] + 10, 2 * q[:, 0], 2 * q[:, 2]], 1)                                                # T * Rz(90) * S(2)
    assert np.allclose(V2[nv:nv + 5], want, atol=1e-5) and F2[nf:nf + 3].tolist() == [[nv, nv + 1, nv + 2], [nv, nv + 2, nv + 3], [nv + 3, nv + 2, nv + 4]]
    qq = q[:4] + np.float32([0, 0, 1])                                                                              # nested: parent * translate
    want = np.stack([-2 * qq[:, 1] + 10, 2 * qq[:, 0], 2 * qq[:, 2]], 1)
    assert np.allclose(V2[nv + 5:nv + 9], want, atol=1e-5)
    assert np.allclose(V2[nv + 9:], V) and np.array_equal(F2[nf + 5:] - (nv + 9), F)                               # second instance, identity
    # no scene graph: geometries are taken as they are
    p2 = tmp_path / "bare.dae"
    p2.write_text(_dae({"cube": (V, tri)}, ""))
    V3, F3 = rmcl_b200.read_mesh_file(str(p2))
    assert np.allclose(V3, V) and np.array_equal(F3, F)
    for bad in ('<COLLADA><library_geometries><geometry id="g"><mesh><triangles count="1"><p>0 1 2</p></triangles></mesh></geometry></library_geometries></COLLADA>',
                "<COLLADA><asset>", "<html/>"):
        bp = tmp_path / "bad.dae"
        bp.write_text(bad)
        with pytest.raises(rmcl_b200.B2Error):
            rmcl_b200.read_mesh_file(str(bp))


def test_ply_header_cannot_force_huge_allocation(tmp_path):
    """a header announcing billions of vertices in a tiny file is refused before anything is reserved"""
    import rmcl_b200
    p = tmp_path / "huge.ply"
    p.write_text("ply\nformat binary_little_endian 1.0\nelement vertex 4000000000\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
                 "property list uchar int vertex_indices\nend_header\n")
    with pytest.raises(rmcl_b200.B2Error) as e:
        rmcl_b200.read_mesh_file(str(p))
    assert "more" in str(e.value)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints exactly one JSON line with the contract's keys; runs without a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--faces", "20000", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "ray-correspondences/sec" and d["unit"] == "rays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
