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


def test_oracle_equals_unmodified_reference_on_storage_logdir(tmp_path):
    """The logdir the GPU test feeds to the device (diskstat / blktrace / strace / pystacks at several thousand lines, with
    the odd lines the reference tolerates) through the UNMODIFIED reference (oracle/_ref, child interpreter) and through
    the oracle: every file identical.  Together with tests/test_gpu_parity.py::test_storage_sources_device_match_oracle_larger
    this ties the device output to the reference itself beyond the small golden fixtures."""
    import shutil
    import pytest
    from helpers import make_storage_logdir, STORAGE_OUTPUTS, SIDE_OUTPUTS, demangle
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    for k, over in enumerate((dict(plot_ratio=1), dict(plot_ratio=3, absolute_timestamp=True, diskstat_filters="xda1"))):
        a = str(tmp_path / ("ref%d" % k)) + "/"
        make_storage_logdir(a, n_disk=300, n_strace=2500, n_blk=3000, n_py=500)
        b = str(tmp_path / ("orc%d" % k)) + "/"
        shutil.copytree(a, b)
        ref_runner.run_in_subprocess(a, None, overrides=over, analyze=False)
        orc.preprocess(orc.OracleConfig(b, nprocs=1, **over), demangle=demangle)
        for fn in sorted(set(STORAGE_OUTPUTS) | set(SIDE_OUTPUTS)):
            ea, eb = os.path.isfile(a + fn), os.path.isfile(b + fn)
            assert ea == eb, fn
            if ea:
                assert open(a + fn, "rb").read() == open(b + fn, "rb").read(), (over, fn)


def test_oracle_equals_unmodified_reference_on_the_larger_side_logdir(tmp_path):
    """The side-file logdir of tests/test_gpu_parity.py::test_side_sources_device_match_oracle_larger_and_edge_cases (620 samples
    x 64 cores x 8 GPUs, 3000 s of 1 Hz files, repeated nvidia-smi blocks, malformed mpstat lines) through the UNMODIFIED
    reference and through the oracle."""
    import shutil
    import time
    import pytest
    from sofa_b200 import synth
    from helpers import demangle
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    os.environ['TZ'] = 'UTC'
    time.tzset()
    a = str(tmp_path / "ref") + "/"
    synth.make_logdir(a, seed=11, n_perf=500)
    synth.write_side_raw(a, seed=12, n_samples=620, n_cores=64, n_gpus=8)
    synth.write_side_raw_1hz(a, seed=13, n_seconds=3000, n_gpus=8)
    with open(a + "nvsmi_query.txt", "a") as f:
        body = open(a + "nvsmi_query.txt").read().split("\n")[1:-1]
        for rep in range(3):
            f.write("\n".join(body) + "\n")
    lines = open(a + "mpstat.txt").read().split("\n")
    lines.insert(1000, "1600000004.1,2,77")
    lines.insert(2000, "")
    lines[3000] += ",1"
    open(a + "mpstat.txt", "w").write("\n".join(lines))
    b = str(tmp_path / "orc") + "/"
    shutil.copytree(a, b)
    over = dict(nvsmi_time_zone=-5, enable_vmstat=True, enable_encode_decode=True)
    ref_runner.run_in_subprocess(a, None, overrides=over, analyze=False)
    orc.preprocess(orc.OracleConfig(b, nprocs=1, **over), demangle=demangle)
    for fn in ("mpstat.csv", "usr_sys.csv", "nvsmi_trace.csv", "netstat.csv", "netbandwidth.csv", "vmstat.csv", "report.js"):
        assert open(a + fn, "rb").read() == open(b + fn, "rb").read(), fn
