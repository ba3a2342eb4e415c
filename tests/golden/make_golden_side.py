#!/usr/bin/env python
"""Golden fixtures for concurrency_breakdown (a14): the UNMODIFIED reference preprocess turns
synthetic mpstat / nvsmi_query / netstat raw logs into mpstat.csv, nvsmi_trace.csv, netstat.csv
(+ cputrace.csv, usr_sys.csv, netbandwidth.csv and the side series of report.js, the N1 preprocess
fixtures; raw inputs are stored next to them, time zone pinned to UTC because the reference reads
the naive nvidia-smi stamps as local time); the UNMODIFIED reference concurrency_breakdown (called directly: the analyze
driver line :981 fails on pandas 3) produces performance.csv and the feature rows.  Build
container only (needs /root/reference)."""
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from sofa_b200 import synth  # noqa: E402

CASES = {
    "side_small": (dict(seed=91, n_samples=60, n_cores=4, n_gpus=2), dict()),
    "side_thresh": (dict(seed=92, n_samples=45, n_cores=2, n_gpus=1), dict(is_idle_threshold=25, sys_mon_rate=5)),
    # + vmstat.txt and nvsmi.txt (1 Hz), every optional series switched on
    "side_full": (dict(seed=93, n_samples=50, n_cores=3, n_gpus=2, hz1=dict(n_seconds=30, n_gpus=2)),
                  dict(enable_vmstat=True, enable_encode_decode=True)),
    # diskstat.txt (+ the files of side_small)
    "disk_small": (dict(seed=95, n_samples=30, n_cores=2, n_gpus=1, disk=dict(n_samples=40)), dict()),
    # strace.txt + blktrace.txt + pystacks.txt (opt-in recorder outputs; vmstat / nvsmi give `cpuid` its value)
    "oracle_strace_blk": (dict(seed=96, n_samples=30, n_cores=2, n_gpus=2, hz1=dict(n_seconds=10, n_gpus=2),
                               strace=dict(n_strace=300, n_blk=200)), dict(plot_ratio=2)),
    "side_abs": (dict(seed=94, n_samples=40, n_cores=2, n_gpus=3, hz1=dict(n_seconds=12, n_gpus=3)),
                 dict(absolute_timestamp=True, enable_vmstat=True, enable_encode_decode=True)),
}


def main():
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    mg.install_stubs()
    sys.path.insert(0, mg.REF_BIN)
    shimdir = tempfile.mkdtemp(prefix="sofa_shims_")
    mg.install_shims(shimdir)
    import pandas as pd
    for name, (mk, ov) in CASES.items():
        work = tempfile.mkdtemp(prefix="sofa_golden_")
        logdir = os.path.join(work, "sofalog") + "/"
        synth.make_logdir(logdir, seed=mk["seed"], n_perf=120)
        mk = dict(mk)
        hz1 = mk.pop("hz1", None)
        disk = mk.pop("disk", None)
        strace = mk.pop("strace", None)
        synth.write_side_raw(logdir, **mk)
        if hz1:
            synth.write_side_raw_1hz(logdir, mk["seed"], **hz1)
        if disk:
            synth.write_diskstat(logdir, mk["seed"], **disk)
        if strace:
            synth.write_strace_blktrace(logdir, mk["seed"], **strace)
            synth.write_pystacks(logdir, mk["seed"])
        cfg = mg.make_ref_cfg(logdir, ov)
        import sofa_preprocess as ref_pre
        with contextlib.redirect_stdout(io.StringIO()):
            ref_pre.sofa_preprocess(cfg)
        import sofa_analyze as ref_an
        df_mpstat = pd.read_csv(logdir + "mpstat.csv")
        df_cpu = pd.read_csv(logdir + "cputrace.csv")
        df_nvsmi = pd.read_csv(logdir + "nvsmi_trace.csv")
        df_bw = pd.read_csv(logdir + "netstat.csv")
        feats = pd.DataFrame({'name': ['elapsed_time'], 'value': [12.5]}, columns=['name', 'value'])
        with contextlib.redirect_stdout(io.StringIO()):
            feats = ref_an.concurrency_breakdown(logdir, cfg, df_mpstat, df_cpu, pd.DataFrame([], columns=cfg.columns),
                                                 df_nvsmi, df_bw, feats)
        # a17 nvsmi_profile (sofa_analyze.py:259-341), called directly; verbose output kept as a fixture
        nv_feats = pd.DataFrame({'name': ['elapsed_time'], 'value': [12.5]}, columns=['name', 'value'])
        cfg.verbose = True
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            nv_feats = ref_an.nvsmi_profile(logdir, cfg, df_nvsmi, nv_feats)
        cfg.verbose = False
        with open(logdir + "nvsmi_profile_verbose.txt", "w") as f:
            f.write(buf.getvalue())
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        if hz1:
            for fn in ("vmstat.txt", "nvsmi.txt", "vmstat.csv"):
                shutil.copy(logdir + fn, os.path.join(out, fn))
        if disk:
            for fn in ("diskstat.txt", "diskstat.csv", "diskstat_vector.csv", "diskstat_vector_ui.csv"):
                shutil.copy(logdir + fn, os.path.join(out, fn))
        if strace:
            for fn in ("strace.txt", "blktrace.txt", "strace.csv", "blktrace.csv", "pystacks.txt", "pystacks.csv"):
                shutil.copy(logdir + fn, os.path.join(out, fn))
        for fn in ("mpstat.csv", "nvsmi_trace.csv", "netstat.csv", "cputrace.csv", "performance.csv", "misc.txt",
                   "usr_sys.csv", "netbandwidth.csv", "report.js",
                   "mpstat.txt", "nvsmi_query.txt", "netstat.txt", "perf.script", "sofa_time.txt",
                   "perf_events_used.txt", "perf_timebase.txt", "nvsmi_profile_verbose.txt"):
            shutil.copy(logdir + fn, os.path.join(out, fn))
        meta = {"make": dict(mk, **({"hz1": hz1} if hz1 else {}), **({"disk": disk} if disk else {}),
                             **({"strace": strace} if strace else {})), "cfg": ov,
                "nvsmi_features": [[str(n), float(v)] for n, v in zip(nv_feats['name'], nv_feats['value'])],
                "features": [[str(n), (None if v != v else float(v))] for n, v in zip(feats['name'], feats['value'])]}
        with open(os.path.join(out, "case.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        print(name, sorted(os.listdir(out)), meta["features"])
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
