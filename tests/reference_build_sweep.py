#!/usr/bin/env python
"""One-off (build container, ~1 minute): the randomised oracle-vs-reference-sources tests over many more seeds than the test suite
runs — 300 registration worlds, 200 EKF event streams, 300 deskew streams.

    python tests/reference_build_sweep.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_reference_build as TR  # noqa: E402
import test_reference_build_ekf as TE  # noqa: E402
import test_reference_build_node as TN  # noqa: E402
from elimaloc_b200 import synth  # noqa: E402


def sweep(name, fn, seeds):
    bad, t0 = [], time.time()
    for seed in seeds:
        try:
            fn(seed)
        except AssertionError as e:
            bad.append((seed, str(e)[:200]))
    print(f"{name}: seeds {seeds[0]}..{seeds[-1]}, mismatches {len(bad)} {bad[:5]}, {time.time() - t0:.0f} s", flush=True)
    return len(bad)


raw = synth.map_u(100, 5.0)
n = sweep("registration worlds (map build, searches, RunRegister, all four methods)", TR.test_randomised_worlds, list(range(6, 306)))
n += sweep("EKF event streams", TE.test_randomised_event_streams, list(range(8, 208)))
n += sweep("deskew streams", lambda s: TN.test_randomised_deskew_streams(raw, s), list(range(14, 314)))
sys.exit(1 if n else 0)
