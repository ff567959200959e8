import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rmcl_b200
from rmcl_b200 import synth
V, F = synth.cube(29)
def step(name, fn):
    try:
        r = fn(); print("OK  ", name); return r
    except Exception as e:
        print("FAIL", name, e); return None
mp = step("map lbvh", lambda: rmcl_b200.Map(V, F))
step("intersect", lambda: mp.intersect([[0,0,0]],[[1,0,0]]))
step("stats", lambda: mp.traversal_stats(np.zeros((100,3),np.float32), np.tile([[1,0,0]],(100,1)).astype(np.float32)))
h = step("rcc", lambda: rmcl_b200.RCCB200Spherical(mp))
m = synth.c1_sensor(); I = synth.make_transform()
step("setmodel", lambda: (h.setTsb(I), h.setModel(m)))
step("find", lambda: h.find(I))
mv = step("modelview", lambda: h.modelView())
step("setranges", lambda: h.setRanges(mv["ranges"]))
step("correctOnce", lambda: h.correctOnce(I, I, 5, 0.0))
step("correct batch", lambda: h.correct(synth.transforms(4)))
up = step("pf create", lambda: rmcl_b200.PCDSensorUpdaterB200(mp))
P, A = synth.pf_particles(100, footprint=(16.0,16.0), z=0.0, margin=0.0); P["t"][:, :2] -= 8
beams = synth.pf_beams(mv["points"], 24)
step("pf update", lambda: up.update(P, A, I, beams))
step("intersect again", lambda: mp.intersect([[0,0,0]],[[1,0,0]]))
