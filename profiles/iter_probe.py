import sys, os, numpy as np
sys.path.insert(0, '.')
import torch, elimaloc_b200 as E
from elimaloc_b200 import synth
raw = synth.map_u(10_000_000, 100.0)
gm = E.VoxelHashMap(1.0, 30, device=0); gm.AddPoints(raw)
scan = synth.scan_u(131072, 40.0)
T0 = synth.se3([50, 50, 50], np.deg2rad([1.0, -2.0, 30.0]))
reg = E.Registration(device=0)
prev = T0
for it in range(1, 8):
    cfg = E.RegistrationConfig(icp_method=0, max_iteration=it, **synth.timing_knobs())
    T = reg.RunRegister(scan, gm, T0, cfg)[0]
    d = np.linalg.inv(prev) @ T
    print(it, "step translation %.4f m rotation %.5f rad" % (np.linalg.norm(d[:3, 3]), np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))), flush=True)
    prev = T
reg.set_stats(True)
cfg = E.RegistrationConfig(icp_method=0, max_iteration=20, **synth.timing_knobs())
reg.RunRegister(scan, gm, T0, cfg)
print(reg.stats())
