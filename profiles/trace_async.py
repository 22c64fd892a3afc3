#!/usr/bin/env python
"""Developer trace of ONE warm iteration in the concurrent-refresh mode (needs a -DELM_TRACE build of the library:
profiles/build_variant.sh trace "-DELM_TRACE"; ELIMALOC_B200_LIB=elimaloc_b200/lib_trace.so python profiles/trace_async.py).
Prints the wall-clock marks (%globaltimer, ns) the two kernels of the last iteration leave in the stats slots, relative to the start of the
reuse kernel: when the refresh blocks became resident, got their first work list, left the tile loop, passed griddepcontrol.wait, reached
the ticket, finished the solve."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import elimaloc_b200 as E
from elimaloc_b200 import synth

gm = E.VoxelHashMap(1.0, 30, device=0)
gm.AddPoints(synth.map_u(10_000_000, 100.0))
T0 = synth.se3([50, 50, 50], np.deg2rad([1.0, -2.0, 30.0]))
reg = E.Registration(device=0)
d = torch.from_numpy(synth.scan_u(131072, 40.0)).cuda()
for iters in (3, 4, 6, 10):
    cfg = E.RegistrationConfig(icp_method=0, max_iteration=iters, **synth.timing_knobs())
    reg.enqueue(d.data_ptr(), 131072, gm, T0, cfg); reg.fetch()          # warm-up of everything
    reg.set_stats(True)
    reg.enqueue(d.data_ptr(), 131072, gm, T0, cfg); reg.fetch()
    s = reg.stats_raw()
    reg.set_stats(False)
    ref = s[2]
    names = {2: "reuse: first block past griddepcontrol.wait (FIRST over all iterations of the call)", 3: "reuse: last block done (LAST)",
             4: "refresh: first block resident (FIRST)", 5: "refresh: first chunk of flags seen (FIRST)", 6: "refresh: last block left the chunk loop (LAST)",
             7: "refresh: last block past griddepcontrol.wait (LAST)", 9: "refresh: last block at the ticket (LAST)", 10: "refresh: solve done (LAST)"}
    print(f"--- call of {iters} iterations ({iters - 2} in the concurrent mode); FIRST marks belong to the first such iteration, LAST marks to the last")
    for k in (2, 4, 5, 3, 6, 7, 9, 10):
        print(f"   slot {k:2d} {(s[k] - ref) / 1e3:9.2f} us   {names[k]}")
