#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference on synthetic logdirs.

Runs only in the build container (needs /root/reference; it does not exist on the
GPU box).  The reference is imported from /root/reference/bin with four import
stubs (cxxfilt, sqlalchemy, fuzzywuzzy, matplotlib - absent here)
and two PATH shims (`nvprof`, `tcpdump`) that cat the pre-dumped text, so the
unmodified `sofa_preprocess(cfg)` runs end to end.  Outputs (cputrace.csv,
gputrace.csv, nettrace.csv, report.js) and the inputs that produced them are
stored under tests/golden/<case>/ ; analyze feature tables as JSON.

    python tests/golden/make_golden.py            # regenerate all cases
"""
import ctypes
import io
import json
import os
import shutil
import stat
import subprocess
import sys
import tempfile
import types
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF_BIN = "/root/reference/bin"
sys.path.insert(0, REPO)

from sofa_b200 import synth  # noqa: E402

CASES = {
    # name: (make_logdir kwargs, cfg overrides)
    "perf_small": (dict(seed=20260921 + 1, n_perf=400), {}),
    "perf_cpucol_ratio3": (dict(seed=77, n_perf=300, cpu_col=True),
                           dict(plot_ratio=3, cpu_filters="idle:black,tensorflow:orange")),
    "perf_cpuinfo": (dict(seed=78, n_perf=200,
                          cpuinfo=[(0.5, 2100.0), (0.7505, 2800.125), (0.752, 3500.5), (9.0, 1200.0)]), {}),
    "perf_nocycles": (dict(seed=79, n_perf=150, perf_events="instructions"), {}),
    "gpu_small": (dict(seed=20260921 + 2, n_gpu=400), {}),
    "gpu_units_p2p": (dict(seed=80, n_gpu=300, units="ms_us_KB", p2p_cols=True), dict(plot_ratio=2)),
    "gpu_us": (dict(seed=81, n_gpu=200, units="us_us_B"), {}),
    "net_small": (dict(seed=82, n_net=300, non_ip_pct=0), {}),
    "net_nan": (dict(seed=83, n_net=300, non_ip_pct=3), dict(plot_ratio=2)),
    "mixed": (dict(seed=20260921 + 3, n_perf=600, n_gpu=350, n_net=100),
              dict(cpu_filters="idle:black,tensorflow:orange")),
    "api_corid": (dict(seed=84, n_gpu=200, n_api=250), dict(cuda_api_tracing=True, plot_ratio=2)),
    "api_nocorid": (dict(seed=85, n_gpu=150, n_api=200, api_corid=False), dict(cuda_api_tracing=True)),
}


def demangle(name):
    """cxxfilt.demangle equivalent: libstdc++ __cxa_demangle (what cxxfilt wraps)."""
    if not name.startswith("_Z"):
        return name
    lib = ctypes.CDLL("libstdc++.so.6")
    libc = ctypes.CDLL("libc.so.6")
    lib.__cxa_demangle.restype = ctypes.c_void_p
    status = ctypes.c_int(0)
    p = lib.__cxa_demangle(name.encode(), None, None, ctypes.byref(status))
    if status.value != 0 or not p:
        raise ValueError("InvalidName: %s" % name)
    s = ctypes.string_at(p).decode()
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(p)
    return s


def install_stubs():
    m = types.ModuleType("cxxfilt"); m.demangle = demangle; sys.modules["cxxfilt"] = m
    m = types.ModuleType("sqlalchemy"); m.create_engine = lambda *a, **k: None; sys.modules["sqlalchemy"] = m
    m = types.ModuleType("fuzzywuzzy"); m.fuzz = types.SimpleNamespace(); m.process = types.SimpleNamespace()
    sys.modules["fuzzywuzzy"] = m
    sys.modules["fuzzywuzzy.fuzz"] = types.ModuleType("fuzzywuzzy.fuzz")
    sys.modules["fuzzywuzzy.process"] = types.ModuleType("fuzzywuzzy.process")
    mpl = types.ModuleType("matplotlib"); mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    os.environ["PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION"] = "python"


def install_shims(shimdir):
    os.makedirs(shimdir, exist_ok=True)
    nv = os.path.join(shimdir, "nvprof")
    with open(nv, "w") as f:
        f.write('#!/bin/sh\n# args: --csv --print-gpu-trace|--print-api-trace -i <file.nvvp>\n'
                'src=gputrace.src\nfor a; do last="$a"; [ "$a" = "--print-api-trace" ] && src=cuda_api_trace.src; done\n'
                'cat "$(dirname "$last")/$src" >&2\n')
    td = os.path.join(shimdir, "tcpdump")
    with open(td, "w") as f:
        f.write('#!/bin/sh\n# args: -q -n -tt -r <file.pcap>\n'
                'for a; do last="$a"; done\ncat "$(dirname "$last")/net.src"\n')
    for p in (nv, td):
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC | stat.S_IXGRP | stat.S_IXOTH)
    os.environ["PATH"] = shimdir + os.pathsep + os.environ["PATH"]


def make_ref_cfg(logdir, overrides):
    """SOFA_Config exactly as bin/sofa:159-312 leaves it for `sofa report`."""
    from sofa_config import SOFA_Config, Filter
    SOFA_Config.cpu_filters = []
    SOFA_Config.gpu_filters = []
    SOFA_Config.net_filters = []
    cfg = SOFA_Config()
    cfg.logdir = logdir if logdir.endswith("/") else logdir + "/"
    cfg.enable_encode_decode = False
    cfg.enable_tcpdump = False
    cfg.cpu_time_offset = 0.0
    cfg.strace_min_time = 1e-6
    ov = dict(overrides)
    cpu_filters = ov.pop("cpu_filters", None)
    if cpu_filters:
        for pair in cpu_filters.split(","):
            cfg.cpu_filters.append(Filter(pair.split(":")[0], pair.split(":")[1]))
    else:
        cfg.cpu_filters.append(Filter("idle", "black"))
    cfg.gpu_filters.append(Filter("CUDA_COPY_H2D", "Red"))
    cfg.gpu_filters.append(Filter("CUDA_COPY_D2H", "Peru"))
    cfg.gpu_filters.append(Filter("CUDA_COPY_P2P", "Purple"))
    if ov.get("cuda_api_tracing"):
        cfg.gpu_filters.append(Filter("CUDA_API", "DarkSlateGray"))          # bin/sofa:276-277
    cfg.gpu_filters.append(Filter("_fw_", "royalblue"))
    cfg.gpu_filters.append(Filter("_bw_", "darkviolet"))
    cfg.gpu_filters.append(Filter("AllReduceKernel", "indigo"))
    for k, v in ov.items():
        setattr(cfg, k, v)
    return cfg


def install_pandas_shim():
    """net_profile (sofa_analyze.py:458) calls MultiIndex.set_levels(..., inplace=True), a keyword pandas 2 removed.
    The shim accepts it again (the frame it touches is only printed with --verbose), so that the UNMODIFIED
    function runs to its outputs (netrank.csv, net_time) on this container's pandas."""
    import pandas as pd
    if getattr(pd.MultiIndex.set_levels, "_sofa_shim", False):
        return
    orig = pd.MultiIndex.set_levels

    def set_levels(self, levels, *, level=None, inplace=None, verify_integrity=True):
        res = orig(self, levels, level=level, verify_integrity=verify_integrity)
        return None if inplace else res
    set_levels._sofa_shim = True
    pd.MultiIndex.set_levels = set_levels


def run_reference(logdir, overrides, analyze=True):
    """Run reference sofa_preprocess (+ the analyze functions that work on this
    pandas) in `logdir`.  Returns dict of feature name -> value (or None)."""
    import sofa_preprocess as ref_pre
    for src, dst in (("gputrace.tmp", "gputrace.src"), ("net.tmp", "net.src"),
                     ("cuda_api_trace.tmp", "cuda_api_trace.src")):
        if os.path.isfile(os.path.join(logdir, src)):
            shutil.copy(os.path.join(logdir, src), os.path.join(logdir, dst))
    cfg = make_ref_cfg(logdir, overrides)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_pre.sofa_preprocess(cfg)
    feats = None
    if analyze and os.path.isfile(os.path.join(logdir, "gputrace.csv")):
        import pandas as pd
        import sofa_analyze as ref_an
        df_gpu = pd.read_csv(os.path.join(logdir, "gputrace.csv"))
        features = pd.DataFrame({'name': ['elapsed_time'], 'value': [12.5]}, columns=['name', 'value'])
        with contextlib.redirect_stdout(io.StringIO()):
            features = ref_an.gpu_profile(logdir, cfg, df_gpu, features)
        feats = {str(n): float(v) for n, v in zip(features['name'], features['value'])}
        with contextlib.redirect_stdout(io.StringIO()):
            feats['__top10_names__'] = ref_an.get_top_k_events(cfg, df_gpu, 10)
    if analyze and os.path.isfile(os.path.join(logdir, "nettrace.csv")):
        import pandas as pd
        import sofa_analyze as ref_an
        install_pandas_shim()
        df_net = pd.read_csv(os.path.join(logdir, "nettrace.csv"))
        nf = pd.DataFrame({'name': ['elapsed_time'], 'value': [12.5]}, columns=['name', 'value'])
        with contextlib.redirect_stdout(io.StringIO()):
            nf = ref_an.net_profile(logdir, cfg, df_net, nf)
        feats = feats if feats is not None else {}
        feats['__net_features__'] = [[str(n), float(v)] for n, v in zip(nf['name'], nf['value'])]
    return feats


def main():
    if not os.path.isdir(REF_BIN):
        sys.exit("needs /root/reference (build container only)")
    install_stubs()
    sys.path.insert(0, REF_BIN)
    shimdir = tempfile.mkdtemp(prefix="sofa_shims_")
    install_shims(shimdir)
    only = sys.argv[1:]
    for name, (mk, ov) in CASES.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp(prefix="sofa_golden_")
        logdir = os.path.join(work, "sofalog") + "/"
        synth.make_logdir(logdir, **mk)
        feats = run_reference(logdir, ov)
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        meta = {"make_logdir": mk, "cfg": ov, "features": feats}
        with open(os.path.join(out, "case.json"), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "comm.csv", "cuda_api_trace.csv", "netrank.csv"):
            p = os.path.join(logdir, fn)
            if os.path.isfile(p):
                shutil.copy(p, os.path.join(out, fn))
        print(name, sorted(os.listdir(out)), feats and len(feats))
        shutil.rmtree(work, ignore_errors=True)
    shutil.rmtree(shimdir, ignore_errors=True)


if __name__ == "__main__":
    main()
