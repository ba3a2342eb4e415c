"""Product host logic of concurrency_breakdown (sofa_b200.sofa_analyze) checked against the
unmodified reference's performance.csv and feature rows (tests/golden/side_*), with a numpy
stand-in for the device window reduction; the GPU test repeats it on the real kernels."""
import json
import os
import shutil

import pytest

from helpers import GOLDEN
from hostsim_engine import HostSimEngine
from sofa_b200.sofa_config import default_report_config
from sofa_b200 import sofa_analyze as an

SIDE = sorted(d for d in os.listdir(GOLDEN) if d.startswith("side_"))


def run_case(name, tmp_path, engine):
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    logdir = str(tmp_path) + "/"
    for fn in ("mpstat.csv", "nvsmi_trace.csv", "netstat.csv", "cputrace.csv", "misc.txt"):
        shutil.copy(os.path.join(d, fn), logdir + fn)
    cfg = default_report_config(logdir)
    for k, v in case["cfg"].items():
        if k in ("is_idle_threshold", "sys_mon_rate"):
            setattr(cfg, k, v)
    feats = an.concurrency_breakdown(logdir, cfg, engine, [('elapsed_time', 12.5)])
    want = [(n, v) for n, v in case["features"]]
    assert [(n, float(v)) for n, v in feats] == want
    assert open(logdir + "performance.csv", "rb").read() == open(os.path.join(d, "performance.csv"), "rb").read()


@pytest.mark.parametrize("name", SIDE)
def test_concurrency_breakdown_host_logic(name, tmp_path):
    run_case(name, tmp_path, HostSimEngine())


def run_nvsmi_case(name, tmp_path, engine, capsys):
    """nvsmi_profile (a17) host logic: features and the verbose tables against the reference's."""
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    logdir = str(tmp_path) + "/"
    shutil.copy(os.path.join(d, "nvsmi_trace.csv"), logdir + "nvsmi_trace.csv")
    cfg = default_report_config(logdir)
    cfg.verbose = True
    cfg.nvsmi_data = os.path.isfile(os.path.join(d, "nvsmi.txt"))     # set by sofa_preprocess when enc/dec rows exist
    capsys.readouterr()
    feats = an.nvsmi_profile(logdir, cfg, engine, [('elapsed_time', 12.5)])
    out = capsys.readouterr().out
    assert [(n, float(v)) for n, v in feats] == [tuple(x) for x in case["nvsmi_features"]]
    assert out == open(os.path.join(d, "nvsmi_profile_verbose.txt")).read()


@pytest.mark.parametrize("name", SIDE)
def test_nvsmi_profile_host_logic(name, tmp_path, capsys):
    run_nvsmi_case(name, tmp_path, HostSimEngine(), capsys)


def test_storage_profiles_host_logic(tmp_path, capsys):
    """strace / blktrace / diskstat sections of the analyze driver (sofa_analyze.py:379-382, 596-700) on the inputs of
    tests/golden/driver_disk: the verbose tables and the features against the unmodified reference's stdout.  The
    device parts (order statistics, Kahan means) are plain-Python stand-ins here; tests/test_gpu_parity.py runs the
    same case through the CUDA library."""
    import re
    import shutil
    from hostsim_engine import HostSimEngine
    from sofa_b200.sofa_config import default_report_config
    from sofa_b200.sofa_preprocess import sofa_preprocess
    import sofa_b200.sofa_preprocess as sp
    sp.SYS_BLOCK = "/nonexistent_sofa_b200/"
    gdir = os.path.join(GOLDEN, "driver_disk")
    a = os.path.join(str(tmp_path), "sofalog") + "/"
    os.makedirs(a)
    for fn in os.listdir(gdir):
        if fn.endswith(".txt") and not fn.startswith("analyze_") or fn.endswith((".tmp", ".script")):
            shutil.copy(os.path.join(gdir, fn), a + fn)
    open(a + "gputrace1.nvvp", "w").close()
    cfg = default_report_config(a)
    eng = HostSimEngine()
    sofa_preprocess(cfg, engine=eng)
    for fn in ("diskstat.csv", "diskstat_vector.csv", "blktrace.csv", "strace.csv", "pystacks.csv"):
        assert os.path.getsize(a + fn) > 200, fn
    capsys.readouterr()
    cfg.verbose = True
    feats = an.blktrace_latency_profile(a, cfg, eng, [])
    feats = an.diskstat_profile(a, cfg, eng, feats)
    got = capsys.readouterr().out.replace(a, "LOGDIR/")
    want = open(os.path.join(gdir, "analyze_verbose.txt")).read()
    i, j = want.index("Offset of Device Report"), want.index("VMSTAT Profiling:")
    j = want.rindex("\n\n\n", 0, j)
    assert got.rstrip("\n") == want[i:j].rstrip("\n")
    table = dict(re.findall(r"^\d+\s+(\w+)\s+(-?[\d.]+|nan)\s*$", want, flags=re.M))
    assert [n for n, _ in feats] == ["blktrace_latency_q1", "blktrace_latency_q2", "blktrace_latency_q3", "diskstat_q1", "diskstat_q2", "diskstat_q3"]
    for n, v in feats:
        assert "%.3lf" % v == table[n], (n, v, table[n])
    os.remove(a + "btt.txt")                       # without the recorder's btt output the reference reports the CSV itself as missing
    with pytest.raises(IOError):
        an.blktrace_latency_profile(a, cfg, eng, [])
