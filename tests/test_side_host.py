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
