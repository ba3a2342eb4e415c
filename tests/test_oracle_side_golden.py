"""Pin the oracle's concurrency_breakdown against the unmodified reference (tests/golden/side_*)."""
import io
import json
import os

import pandas as pd
import pytest

from helpers import GOLDEN
from oracle import sofa_oracle as orc

SIDE = sorted(d for d in os.listdir(GOLDEN) if d.startswith("side_"))


@pytest.mark.parametrize("name", SIDE)
def test_oracle_concurrency_breakdown_matches_reference(name):
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    feats, perf = orc.concurrency_breakdown(pd.read_csv(d + "/mpstat.csv"), pd.read_csv(d + "/cputrace.csv"),
                                            pd.read_csv(d + "/nvsmi_trace.csv"), pd.read_csv(d + "/netstat.csv"),
                                            **{k: v for k, v in case["cfg"].items() if k in ("is_idle_threshold", "sys_mon_rate")})
    want = [tuple(x) for x in case["features"][1:]]
    assert [(n, float(v)) for n, v in feats] == [(n, v) for n, v in want]
    buf = io.StringIO()
    perf.to_csv(buf)
    assert buf.getvalue().encode() == open(d + "/performance.csv", "rb").read()


@pytest.mark.parametrize("name", SIDE)
def test_oracle_side_preprocess_matches_reference(name, tmp_path):
    """N1: mpstat.txt / nvsmi_query.txt / netstat.txt -> CSVs + report.js series, byte for byte."""
    from helpers import build_side_logdir, SIDE_OUTPUTS, golden_bytes, demangle, preprocess_overrides
    case, logdir = build_side_logdir(name, tmp_path)
    orc.preprocess(orc.OracleConfig(logdir, nprocs=1, **preprocess_overrides(case)), demangle=demangle)
    for fn in SIDE_OUTPUTS:
        want = golden_bytes(name, fn)
        if want is None:
            assert not os.path.isfile(logdir + fn), fn
        else:
            assert open(logdir + fn, "rb").read() == want, fn


@pytest.mark.parametrize("name", SIDE)
def test_oracle_nvsmi_profile_matches_reference(name):
    d = os.path.join(GOLDEN, name)
    case = json.load(open(os.path.join(d, "case.json")))
    feats = orc.nvsmi_profile(pd.read_csv(d + "/nvsmi_trace.csv"), nvsmi_data=os.path.isfile(d + "/nvsmi.txt"))
    assert [(n, float(v)) for n, v in feats] == [tuple(x) for x in case["nvsmi_features"][1:]]


def test_oracle_diskstat_matches_reference(tmp_path):
    """diskstat.txt (N1): diskstat.csv, both vector tables and the report.js series."""
    import shutil
    from helpers import demangle
    name = "disk_small"
    d = os.path.join(GOLDEN, name)
    logdir = str(tmp_path) + "/"
    os.environ['TZ'] = 'UTC'
    import time
    time.tzset()
    for fn in ("mpstat.txt", "nvsmi_query.txt", "netstat.txt", "diskstat.txt", "perf.script", "sofa_time.txt",
               "perf_events_used.txt", "perf_timebase.txt", "misc.txt"):
        shutil.copy(os.path.join(d, fn), logdir + fn)
    orc.preprocess(orc.OracleConfig(logdir, nprocs=1), demangle=demangle)
    for fn in ("diskstat.csv", "diskstat_vector.csv", "diskstat_vector_ui.csv", "report.js"):
        assert open(logdir + fn, "rb").read() == open(os.path.join(d, fn), "rb").read(), fn


def test_oracle_strace_blktrace_match_reference(tmp_path):
    """strace.txt / blktrace.txt / pystacks.txt (N1): their CSVs and report.js series."""
    import shutil
    import time
    from helpers import demangle
    d = os.path.join(GOLDEN, "oracle_strace_blk")
    case = json.load(open(os.path.join(d, "case.json")))
    logdir = str(tmp_path) + "/"
    os.environ['TZ'] = 'UTC'
    time.tzset()
    for fn in ("mpstat.txt", "nvsmi_query.txt", "netstat.txt", "vmstat.txt", "nvsmi.txt", "strace.txt", "blktrace.txt", "pystacks.txt", "perf.script",
               "sofa_time.txt", "perf_events_used.txt", "perf_timebase.txt", "misc.txt"):
        shutil.copy(os.path.join(d, fn), logdir + fn)
    orc.preprocess(orc.OracleConfig(logdir, nprocs=1, **case["cfg"]), demangle=demangle)
    for fn in ("strace.csv", "blktrace.csv", "pystacks.csv", "vmstat.csv", "nvsmi_trace.csv", "report.js"):
        assert open(logdir + fn, "rb").read() == open(os.path.join(d, fn), "rb").read(), fn
