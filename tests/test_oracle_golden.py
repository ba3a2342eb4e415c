"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden)."""
import os

import pandas as pd
import pytest

from helpers import golden_cases, golden_bytes, build_case_logdir, demangle
from oracle import sofa_oracle as orc


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_reference_files(name, tmp_path):
    case, logdir = build_case_logdir(name, tmp_path)
    cfg = orc.OracleConfig(logdir, nprocs=1, **case["cfg"])
    orc.preprocess(cfg, demangle=demangle)
    checked = 0
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "cuda_api_trace.csv"):
        want = golden_bytes(name, fn)
        if want is None:
            assert not os.path.isfile(logdir + fn)
            continue
        with open(logdir + fn, "rb") as f:
            got = f.read()
        assert got == want, "%s/%s differs from the reference's output" % (name, fn)
        checked += 1
    assert checked >= 2


@pytest.mark.parametrize("name", [n for n in golden_cases() if golden_bytes(n, "gputrace.csv")])
def test_oracle_gpu_profile_matches_reference(name, tmp_path):
    case, logdir = build_case_logdir(name, tmp_path)
    cfg = orc.OracleConfig(logdir, nprocs=1, **case["cfg"])
    orc.preprocess(cfg, demangle=demangle, emit_json=False)
    df = pd.read_csv(logdir + "gputrace.csv")
    feats = dict(orc.gpu_profile(df, logdir))
    want = case["features"]
    for k, v in want.items():
        if k in ("elapsed_time", "__net_features__"):
            continue
        if k == "__top10_names__":
            top = orc.top_k_events(df, 10)
            vals = list(top.values)
            for i, nm in enumerate(v):                 # tie order is unpinned (unstable sort)
                if vals.count(vals[i]) == 1:
                    assert top.index[i] == nm
            continue
        assert float(feats[k]) == v, k
    with open(logdir + "comm.csv", "rb") as f:
        assert f.read() == golden_bytes(name, "comm.csv")


@pytest.mark.parametrize("name", [n for n in golden_cases() if golden_bytes(n, "netrank.csv")])
def test_oracle_net_profile_matches_reference(name, tmp_path):
    """a15: net_time and netrank.csv against the unmodified reference (run with the pandas shim of make_golden.py)."""
    case, logdir = build_case_logdir(name, tmp_path)
    cfg = orc.OracleConfig(logdir, nprocs=1, **case["cfg"])
    orc.preprocess(cfg, demangle=demangle, emit_json=False)
    feats = orc.net_profile(pd.read_csv(logdir + "nettrace.csv"), logdir)
    assert [(n, float(v)) for n, v in feats] == [tuple(x) for x in case["features"]["__net_features__"][1:]]
    with open(logdir + "netrank.csv", "rb") as f:
        assert f.read() == golden_bytes(name, "netrank.csv")


@pytest.mark.parametrize("kw,ratio", [
    (dict(seed=101, n_perf=30000, n_gpu=20000, n_net=10000, non_ip_pct=1), 7),
    (dict(seed=102, n_perf=25000, cpu_col=True, n_gpu=15000, units="ms_us_KB", p2p_cols=True, n_net=5000, non_ip_pct=0), 1),
])
def test_oracle_equals_unmodified_reference_on_the_larger_gpu_test_logs(kw, ratio, tmp_path):
    """The logdirs of tests/test_gpu_parity.py::test_cuda_path_matches_oracle_on_larger_logs through the UNMODIFIED reference
    (oracle/_ref, child interpreter) and through the oracle: the device is compared with the oracle there, the oracle with
    the reference here - at sizes far beyond the golden fixtures."""
    import shutil
    from sofa_b200 import synth
    from helpers import demangle
    from oracle import ref_runner
    from oracle import sofa_oracle as orc
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    a, b = str(tmp_path / "ref") + "/", str(tmp_path / "orc") + "/"
    synth.make_logdir(a, **kw)
    shutil.copytree(a, b)
    ref_runner.run_in_subprocess(a, "idle:black,tensorflow:orange", overrides=dict(plot_ratio=ratio), analyze=False)
    orc.preprocess(orc.OracleConfig(b, plot_ratio=ratio, cpu_filters="idle:black,tensorflow:orange"), demangle=demangle)
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
        assert open(a + fn, "rb").read() == open(b + fn, "rb").read(), fn
