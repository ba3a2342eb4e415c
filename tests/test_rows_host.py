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
