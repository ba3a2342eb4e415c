#!/usr/bin/env python
"""Golden fixtures of the UNMODIFIED reference driver `sofa_analyze(cfg)` (bin/sofa_analyze.py:793-1055): its whole
stdout (feature table, verbose tables of every profile function) and the files it writes, on synthetic logdirs that
went through the unmodified `sofa_preprocess(cfg)` first.  Build container only (needs /root/reference).

The driver and three of its profile functions use pandas APIs that pandas >= 2 removed; they are restored at RUN
TIME (the reference source is untouched, nothing of it is copied):
  * MultiIndex.set_levels(inplace=True)      net_profile  :458      (make_golden.install_pandas_shim)
  * DataFrame.append                          driver       :981      (result unused there)
  * iterating `Series.str` (tuple unpacking)  mpstat_profile :740, vmstat_profile :713
  * `df_summary.iloc[i]['USR'] = x`           mpstat_profile :752-756: chained assignment, a silent no-op under pandas 3
    copy-on-write (cpu_util prints 0 and --verbose dies on int(nan)).  The function is re-executed from its own
    source with those five statements spelled `df_summary.loc[i, 'USR'] = x` - what every pandas the reference
    was written for did.
  * Index.format(), `series[-1]` on a string index (positional fallback), read_table(delim_whitespace=True)
                                              diskstat_profile :676, :685; blktrace_latency_profile :607 (driver_disk)
matplotlib is absent: `plt` is a stub whose figures swallow savefig (no PDF is written).

    python tests/golden/make_golden_driver.py
"""
import contextlib
import inspect
import io
import json
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from sofa_b200 import synth  # noqa: E402

OUT_FILES = ("netrank.csv", "performance.csv", "comm.csv")
IN_FILES = ("misc.txt", "sofa_time.txt", "perf_events_used.txt", "perf_timebase.txt", "perf.script", "gputrace.tmp", "net.tmp",
            "mpstat.txt", "nvsmi_query.txt", "netstat.txt", "vmstat.txt", "nvsmi.txt", "cpuinfo.txt",
            "diskstat.txt", "strace.txt", "blktrace.txt", "pystacks.txt", "btt.txt", "offset_all.txt")

CASES = {
    # name: (make_logdir kwargs, side kwargs, 1 Hz side files?, spotlight pattern?, cfg overrides)
    "driver_full": (dict(seed=501, n_perf=900, n_gpu=600, n_net=300), dict(n_samples=60, n_cores=4, n_gpus=2), True, False,
                    dict(cpu_filters="idle:black,tensorflow:orange")),
    "driver_spotlight": (dict(seed=502, n_perf=500, n_gpu=300, n_net=0), dict(n_samples=80, n_cores=2, n_gpus=2), False, True,
                         dict(spotlight_gpu=True)),
    "driver_perf_only": (dict(seed=503, n_perf=400), dict(n_samples=30, n_cores=8, n_gpus=1), False, False, {}),
    # + diskstat.txt, strace.txt, blktrace.txt (with the recorder's btt.txt / offset_all.txt), pystacks.txt: strace_profile,
    # blktrace_latency_profile, diskstat_profile.  Device names that exist on no machine (no /sys/block lookups hit).
    "driver_disk": (dict(seed=504, n_perf=300, n_gpu=200), dict(n_samples=50, n_cores=2, n_gpus=1), True, False, {},
                    dict(disk=dict(n_samples=45, devs=("xda", "xda1", "fake0n1", "loopx", "dm-x")), strace=dict(n_strace=200, n_blk=300),
                         btt=True)),
}


def install_driver_shims():
    import pandas as pd
    import matplotlib.pyplot as plt
    mg.install_pandas_shim()

    class _Fig:
        def savefig(self, *a, **k):
            pass
    plt.figure = lambda *a, **k: _Fig()
    for nm in ('plot', 'legend', 'title', 'xlabel', 'ylabel', 'close'):
        setattr(plt, nm, lambda *a, **k: None)
    if not hasattr(pd.DataFrame, 'append'):
        pd.DataFrame.append = lambda self, other, **kw: pd.concat(
            [self, other.to_frame().T if isinstance(other, pd.Series) else other])
    if not hasattr(pd.Index, 'format'):
        pd.Index.format = lambda self: [str(x) for x in self]
    _getitem = pd.Series.__getitem__

    def _getitem_positional(self, key):     # pandas < 3: an int on a non-integer index is a position
        try:
            return _getitem(self, key)
        except KeyError:
            if isinstance(key, int):
                return self.iloc[key]
            raise
    pd.Series.__getitem__ = _getitem_positional
    _read_table = pd.read_table

    def _read_table_ws(*a, **k):
        if k.pop('delim_whitespace', False):
            k['sep'] = r'\s+'
        return _read_table(*a, **k)
    pd.read_table = _read_table_ws
    from pandas.core.strings.accessor import StringMethods

    def _iter(self):                       # pandas < 2: yields the i-th element of every row until a column is all-NaN
        i = 0
        g = self.get(i)
        while g.notna().any():
            yield g
            i += 1
            g = self.get(i)
    StringMethods.__iter__ = _iter


def patch_mpstat_profile(ref_an):
    src = inspect.getsource(ref_an.mpstat_profile)
    new, n = re.subn(r"df_summary\.iloc\[index\]\['(\w+)'\] =", r"df_summary.loc[index, '\1'] =", src)
    assert n == 5, n
    ns = ref_an.__dict__
    exec(compile(new, "<mpstat_profile with single-step assignment>", "exec"), ns)


def spotlight_nvsmi(logdir, n_samples, n_gpus):
    """GPU 0: idle, a hot region (sm >= 50 for 30 samples), idle again: the ROI state machine (:872-895) fires."""
    import datetime
    t0 = synth.TIME_BASE + 0.5
    with open(os.path.join(logdir, "nvsmi_query.txt"), "w") as f:
        f.write("timestamp, name, index, utilization.gpu [%], utilization.memory [%]\n")
        for s in range(n_samples):
            t = t0 + 0.013 + s * 0.1
            ts = datetime.datetime.fromtimestamp(t).strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]
            for g in range(n_gpus):
                sm = (60 + (s * 7 + g) % 35) if 12 <= s < 42 else (s * 3 + g) % 9
                f.write("%s, Tesla V100-SXM2-16GB, %d, %d %%, %d %%\n" % (ts, g, sm, (s * 11 + g) % 50))


def write_btt(logdir, seed):
    """btt.txt / offset_all.txt as the recorder's `btt -i blktrace.out -B offset` leaves them (sofa_preprocess.py:684-686)"""
    with open(os.path.join(logdir, "btt.txt"), "w") as f:
        f.write("==================== All Devices ====================\n\n")
        f.write("            ALL           MIN           AVG           MAX           N\n")
        f.write("--------------- ------------- ------------- ------------- -----------\n\n")
        f.write("Q2Q               0.000001254   0.003412771   0.120034411         299\n")
        f.write("D2C               0.000211341   0.004807392   0.031209876         143\n\n")
        f.write("==================== Device Merge Information ====================\n\n")
        f.write("       DEV |       #Q       #D   Ratio |   BLKmin   BLKavg   BLKmax    Total\n")
    with open(os.path.join(logdir, "offset_all.txt"), "w") as f:
        for i in range(40):
            b = 691248304 + 1024 * ((i * 7 + seed) % 97)
            f.write("%.9f %d %d\n" % (0.001 * i, b, b + 1024))


def run_case(name, mk, side, hz1, spot, ov, extra=None):
    import sofa_preprocess as ref_pre
    import sofa_analyze as ref_an
    work = tempfile.mkdtemp(prefix="sofa_golden_drv_")
    logdir = os.path.join(work, "sofalog") + "/"
    synth.make_logdir(logdir, **mk)
    synth.write_side_raw(logdir, mk["seed"], **side)
    if hz1:
        synth.write_side_raw_1hz(logdir, mk["seed"], n_seconds=12, n_gpus=side["n_gpus"])
    if spot:
        spotlight_nvsmi(logdir, side["n_samples"], side["n_gpus"])
    if extra:
        synth.write_diskstat(logdir, mk["seed"], **extra["disk"])
        synth.write_strace_blktrace(logdir, mk["seed"], **extra["strace"])
        synth.write_pystacks(logdir, mk["seed"])
        if extra.get("btt"):
            write_btt(logdir, mk["seed"])
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for fn in IN_FILES:
        if os.path.isfile(logdir + fn):
            shutil.copy(logdir + fn, os.path.join(out, fn))
    for src, dst in (("gputrace.tmp", "gputrace.src"), ("net.tmp", "net.src")):
        if os.path.isfile(logdir + src):
            shutil.copy(logdir + src, logdir + dst)
    cfg = mg.make_ref_cfg(logdir, ov)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_pre.sofa_preprocess(cfg)
    meta = {"make_logdir": mk, "side": side, "hz1": hz1, "spotlight_pattern": spot, "cfg": ov, "variants": {}}
    if extra:
        meta["extra"] = extra
    for verbose in (False, True):
        cfg = mg.make_ref_cfg(logdir, dict(ov, verbose=verbose))
        cfg.roi_begin = cfg.roi_end = 0
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_an.sofa_analyze(cfg)
        text = buf.getvalue().replace(logdir, "LOGDIR/").replace(logdir.rstrip("/"), "LOGDIR")
        tag = "verbose" if verbose else "plain"
        with open(os.path.join(out, "analyze_%s.txt" % tag), "w") as f:
            f.write(text)
        meta["variants"][tag] = {"roi_begin": float(cfg.roi_begin), "roi_end": float(cfg.roi_end)}
    for fn in OUT_FILES:
        if os.path.isfile(logdir + fn):
            shutil.copy(logdir + fn, os.path.join(out, fn))
    with open(os.path.join(out, "case.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(name, sorted(os.listdir(out)), meta["variants"])
    shutil.rmtree(work, ignore_errors=True)


def main():
    if not os.path.isdir(mg.REF_BIN):
        sys.exit("needs /root/reference (build container only)")
    mg.install_stubs()
    sys.path.insert(0, mg.REF_BIN)
    shimdir = tempfile.mkdtemp(prefix="sofa_shims_")
    mg.install_shims(shimdir)
    install_driver_shims()
    import sofa_analyze as ref_an
    patch_mpstat_profile(ref_an)
    only = sys.argv[1:]
    for name, args in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *args)
    shutil.rmtree(shimdir, ignore_errors=True)


if __name__ == "__main__":
    main()
