"""v1 batched correct(): 1000 poses x vlp16_900 on the 1M-triangle building; CUDA-event time per call (B2_BATCH_MINB selects the register cap)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rmcl_b200
from rmcl_b200 import synth
V, F = synth.building(1_000_000)
gmap = rmcl_b200.Map(V, F)
Tsb, Tgt = synth.scenario_tsb(), synth.building_gt_pose()
hv = rmcl_b200.SphereCorrectorB200(gmap)
hv.setTsb(Tsb)
mv = synth.vlp16_900(); mv.range_min = 0.0
hv.setModel(mv); hv.setParams(1.0, 0.15); hv.find(Tgt); hv.setInputData(hv.modelView()["ranges"])
T = synth.transforms(1000); T[:] = synth.compose(Tgt, synth.scenario_pose_offset())
T["t"] += np.random.default_rng(0).uniform(-0.05, 0.05, (1000, 3)).astype(np.float32)
Td = torch.from_numpy(T.view(np.float32).reshape(-1, 8).copy()).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = hv.correct(Td); b.record(); torch.cuda.synchronize()
    if i >= 3: ts.append(a.elapsed_time(b))
print("B2_BATCH_MINB=%s  ms per correct(): median %.3f  -> %.2f G rays/s" % (os.environ.get("B2_BATCH_MINB", "4"), np.median(ts), 1000 * 14400 / np.median(ts) / 1e6))
