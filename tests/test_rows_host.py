"""Product host logic (sofa_b200.sofa_preprocess) + CPU build of the device row codecs,
checked byte for byte against the unmodified reference's outputs (tests/golden)."""
import os

import pytest

from helpers import golden_cases, golden_bytes, build_case_logdir
from hostsim_engine import HostSimEngine
from sofa_b200.sofa_config import default_report_config
from sofa_b200.sofa_preprocess import sofa_preprocess


@pytest.mark.parametrize("name", golden_cases())
def test_row_codecs_reproduce_reference_files(name, tmp_path):
    case, logdir = build_case_logdir(name, tmp_path)
    ov = dict(case["cfg"])
    cfg = default_report_config(logdir, cpu_filters=ov.pop("cpu_filters", None),
                                cuda_api_tracing=ov.pop("cuda_api_tracing", False))
    for k, v in ov.items():
        setattr(cfg, k, v)
    sofa_preprocess(cfg, engine=HostSimEngine())
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "cuda_api_trace.csv"):
        want = golden_bytes(name, fn)
        if want is None:
            assert not os.path.isfile(logdir + fn)
            continue
        with open(logdir + fn, "rb") as f:
            got = f.read()
        if got != want:
            i = next(k for k in range(min(len(got), len(want))) if got[k] != want[k]) if got[:len(want)] != want[:len(got)] else min(len(got), len(want))
            raise AssertionError("%s/%s differs at byte %d: got %r want %r" % (name, fn, i, got[max(0, i - 80):i + 40], want[max(0, i - 80):i + 40]))


@pytest.mark.parametrize("name", [n for n in golden_cases() if golden_bytes(n, "netrank.csv")])
def test_net_profile_host_logic(name, tmp_path):
    """a15 host logic (ranking, address text, netrank.csv) against the reference's file."""
    from sofa_b200 import sofa_analyze as an
    case, logdir = build_case_logdir(name, tmp_path)
    ov = dict(case["cfg"])
    cfg = default_report_config(logdir, cpu_filters=ov.pop("cpu_filters", None))
    for k, v in ov.items():
        setattr(cfg, k, v)
    eng = HostSimEngine()
    sofa_preprocess(cfg, engine=eng)
    feats = an.net_profile(logdir, cfg, eng, [])
    assert float(dict(feats)["net_time"]) == case["features"]["__net_features__"][1][1]
    assert open(logdir + "netrank.csv", "rb").read() == golden_bytes(name, "netrank.csv")


def test_unconvertible_perf_sample_costs_the_cpu_trace_not_the_run(tmp_path):
    """The reference wraps its perf section in a bare `except:` (sofa_preprocess.py:1783-1837): a blank or malformed
    perf.script line drops the CPU trace ("no perf traces.", no cputrace.csv, no cpu series) and the run goes on - while
    a malformed nvprof or tcpdump line raises.  Product (host logic + CPU build of the row codecs) against the oracle,
    which is pinned on exactly this against the unmodified reference by the fuzz of tests/test_fuzz_ref_cpu.py."""
    import os
    import shutil
    import pytest
    from helpers import demangle
    from hostsim_engine import HostSimEngine, HostSimError
    from oracle import sofa_oracle as orc
    from sofa_b200 import synth
    from sofa_b200.sofa_config import default_report_config
    from sofa_b200.sofa_preprocess import sofa_preprocess
    for k, bad in enumerate(("\n", " 4242/4242     5000.000404: 12x4 cycles:      7fbb9bbb4eaa memcpy+0x7cb (/lib/libc.so)\n", "4242\n")):
        a, b = str(tmp_path / ("a%d" % k)) + "/", str(tmp_path / ("b%d" % k)) + "/"
        synth.make_logdir(a, seed=77 + k, n_perf=50, n_gpu=30, n_net=20)
        lines = open(a + "perf.script").read().splitlines(True)
        lines.insert(20, bad)
        open(a + "perf.script", "w").write("".join(lines))
        shutil.copytree(a, b)
        sofa_preprocess(default_report_config(a), engine=HostSimEngine())
        orc.preprocess(orc.OracleConfig(b, nprocs=1), demangle=demangle)
        assert not os.path.isfile(a + "cputrace.csv") and not os.path.isfile(b + "cputrace.csv")
        for fn in ("gputrace.csv", "nettrace.csv", "report.js"):
            assert open(a + fn, "rb").read() == open(b + fn, "rb").read(), fn
        assert b"cpu_trace" not in open(a + "report.js", "rb").read()
    # the same damage in the nvprof text is an error in the reference, and here
    a = str(tmp_path / "g") + "/"
    synth.make_logdir(a, seed=80, n_perf=20, n_gpu=30)
    lines = open(a + "gputrace.tmp").read().splitlines(True)
    lines.insert(10, "garbage\n")
    open(a + "gputrace.tmp", "w").write("".join(lines))
    with pytest.raises(HostSimError):
        sofa_preprocess(default_report_config(a), engine=HostSimEngine())
