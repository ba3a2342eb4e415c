#!/usr/bin/env python
"""Golden vectors for the AISI per-iteration profile: the UNMODIFIED reference `iter_profile` (bin/sofa_aisi.py:21-59)
called on the per-iteration slices of :429-453 (gpu rows of deviceId 1 with times[i-1] <= timestamp < times[i]) of a
gputrace.csv written by the unmodified reference preprocess.  The iteration boundaries are given (the reference finds
them with a suffix tree + fuzzy matching + KMeans, out of scope); cpu / strace / mpstat frames are passed empty (the
reference slices them with a boolean mask that belongs to the gpu frame, which current pandas refuses).
Build container only.

    python tests/golden/make_golden_aisi.py
"""
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from sofa_b200 import synth  # noqa: E402

CASES = {"aisi_iters": (dict(seed=611, n_gpu=6000), 9), "aisi_units": (dict(seed=612, n_gpu=2500, units="ms_us_KB", p2p_cols=True), 5)}


def main():
    if not os.path.isdir(mg.REF_BIN):
        sys.exit("needs /root/reference (build container only)")
    mg.install_stubs()
    sys.modules['fuzzywuzzy'].fuzz = types.SimpleNamespace(ratio=lambda a, b: 0)
    sys.path.insert(0, mg.REF_BIN)
    shimdir = tempfile.mkdtemp(prefix="sofa_shims_")
    mg.install_shims(shimdir)
    import numpy as np
    import pandas as pd
    import sofa_preprocess as ref_pre
    import sofa_aisi as ref_aisi
    for name, (mk, nb) in CASES.items():
        work = tempfile.mkdtemp(prefix="sofa_golden_aisi_")
        logdir = os.path.join(work, "sofalog") + "/"
        synth.make_logdir(logdir, **mk)
        shutil.copy(logdir + "gputrace.tmp", logdir + "gputrace.src")
        cfg = mg.make_ref_cfg(logdir, {})
        with contextlib.redirect_stdout(io.StringIO()):
            ref_pre.sofa_preprocess(cfg)
        df_gpu = pd.read_csv(logdir + "gputrace.csv")
        x1 = df_gpu[df_gpu.deviceId == 1]
        lo, hi = float(x1['timestamp'].min()), float(x1['timestamp'].max())
        times = [float(t) for t in np.linspace(lo - 1e-4, hi + 1e-4, nb)]
        times[3] = float(x1['timestamp'].iloc[len(x1) // 3])          # a boundary exactly on a row's timestamp
        times = sorted(times)
        empty = pd.DataFrame([], columns=df_gpu.columns)
        cfg.aisi_via_strace = False
        rows = []
        for i in range(1, len(times)):
            sl = x1[(x1['timestamp'] >= times[i - 1]) & (x1['timestamp'] < times[i])]
            if len(sl) == 0:
                rows.append(None)
                continue
            with contextlib.redirect_stdout(io.StringIO()):
                d = ref_aisi.iter_profile(cfg, [], empty, sl, empty, empty)
            rows.append({k: float(v) for k, v in d.items()})
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        for fn in ("gputrace.tmp", "misc.txt", "sofa_time.txt", "perf_events_used.txt", "perf_timebase.txt"):
            shutil.copy(logdir + fn, os.path.join(out, fn))
        with open(os.path.join(out, "case.json"), "w") as f:
            json.dump({"make_logdir": mk, "times": times, "iterations": rows}, f, indent=1)
        print(name, len(rows), sum(r is not None for r in rows))
        shutil.rmtree(work, ignore_errors=True)
    shutil.rmtree(shimdir, ignore_errors=True)


if __name__ == "__main__":
    main()
