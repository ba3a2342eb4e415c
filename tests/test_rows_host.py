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
