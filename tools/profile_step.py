#!/usr/bin/env python
"""One device-resident pass of the bench workload (for ncu): python tools/profile_step.py EVENTS [STEPS]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sofa_b200.engine import CudaEngine

events = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = CudaEngine(0)
wl = bench.Workload(eng, events, bench.SEED)
wl.generate()
wl.fetch_host_text()
wl.make_side()
kinds = [int(x) for x in os.environ.get('SOFA_KINDS', '0,1,2').split(',')]
analyze = os.environ.get('SOFA_ANALYZE', '1') == '1' and 1 in kinds
for _ in range(steps):
    wl.step(False, kinds=kinds, analyze=analyze)
eng.close()
print("done", events, steps)
