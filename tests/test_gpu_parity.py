"""GPU parity: the product path (sofa_preprocess -> ctypes -> libsofa_b200.so -> sm_100a
kernels) against (1) the reference's golden files, (2) the CPU oracle on larger seeded
logdirs, (3) column-level bit equality."""
import os

import numpy as np
import pytest

from helpers import golden_cases, golden_bytes, build_case_logdir
from sofa_b200 import synth
from sofa_b200.sofa_config import default_report_config
from sofa_b200.sofa_preprocess import sofa_preprocess, demangle
from oracle import sofa_oracle as orc

pytestmark = pytest.mark.gpu


def _diff(name, fn, got, want):
    if got == want:
        return
    n = min(len(got), len(want))
    i = next((k for k in range(n) if got[k] != want[k]), n)
    raise AssertionError("%s/%s differs at byte %d (len %d vs %d): got %r want %r" % (
        name, fn, i, len(got), len(want), got[max(0, i - 80):i + 40], want[max(0, i - 80):i + 40]))


@pytest.mark.parametrize("name", golden_cases())
def test_cuda_path_reproduces_reference_files(name, tmp_path):
    case, logdir = build_case_logdir(name, tmp_path)
    ov = dict(case["cfg"])
    cfg = default_report_config(logdir, cpu_filters=ov.pop("cpu_filters", None),
                                cuda_api_tracing=ov.pop("cuda_api_tracing", False))
    for k, v in ov.items():
        setattr(cfg, k, v)
    sofa_preprocess(cfg)
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "cuda_api_trace.csv"):
        want = golden_bytes(name, fn)
        if want is None:
            assert not os.path.isfile(logdir + fn)
            continue
        with open(logdir + fn, "rb") as f:
            _diff(name, fn, f.read(), want)


@pytest.mark.parametrize("kw", [
    dict(seed=101, n_perf=30000, n_gpu=20000, n_net=10000, non_ip_pct=1),
    dict(seed=102, n_perf=25000, cpu_col=True, n_gpu=15000, units="ms_us_KB", p2p_cols=True, n_net=5000, non_ip_pct=0),
])
@pytest.mark.parametrize("ratio", [1, 7])
def test_cuda_path_matches_oracle_on_larger_logs(kw, ratio, tmp_path):
    a = os.path.join(str(tmp_path), "gpu") + "/"
    b = os.path.join(str(tmp_path), "cpu") + "/"
    for d in (a, b):
        synth.make_logdir(d, **kw)
    cfg = default_report_config(a, cpu_filters="idle:black,tensorflow:orange")
    cfg.plot_ratio = ratio
    sofa_preprocess(cfg)
    ocfg = orc.OracleConfig(b, plot_ratio=ratio, cpu_filters="idle:black,tensorflow:orange")
    orc.preprocess(ocfg, demangle=demangle)
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("seed%d" % kw["seed"], fn, f.read(), g.read())


def test_columns_bit_exact(tmp_path):
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import hostlogic as hl
    logdir = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(logdir, seed=9, n_perf=5000)
    eng = CudaEngine(0)
    try:
        p = hl.base_params(hl.SRC_PERF, 1600000000.0, False, hl.read_perf_timebase(logdir))
        with open(logdir + "perf.script", "rb") as f:
            text = f.read()
        n, _ = eng.load(hl.SRC_PERF, text, p)
        assert n == 5000
        rows = [orc.parse_perf_sample(l, False, 1600000000.0, p['t_offset'], [0.0], [3000.0], 'cycles')
                for l in text.decode().splitlines()]
        for col in (0, 1, 2):
            want = np.array([r[col] for r in rows], dtype=np.float64)
            got = eng.column(hl.SRC_PERF, col, n)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), col
        assert np.array_equal(eng.column(hl.SRC_PERF, 10, n), np.array([r[10] for r in rows], dtype=np.int32))
    finally:
        eng.close()


def test_malformed_line_fails_loudly(tmp_path):
    """nvprof / tcpdump text: a record the reference dies on (IndexError / ValueError inside Pool.map) is an error that names
    its row.  perf.script: the reference swallows every exception of its perf section (sofa_preprocess.py:1783-1837) - the
    CPU trace is dropped with a warning and the run goes on; so does this (oracle pinned on it by tests/test_fuzz_ref_cpu.py)."""
    from sofa_b200._lib import SofaB200Error
    logdir = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(logdir, seed=9, n_perf=100, n_gpu=100)
    with open(logdir + "perf.script", "a") as f:
        f.write("garbage line\n")
    sofa_preprocess(default_report_config(logdir))
    assert not os.path.isfile(logdir + "cputrace.csv") and os.path.isfile(logdir + "gputrace.csv")
    with open(logdir + "report.js") as f:
        js = f.read()
    assert "cpu_trace = " not in js and "gpu_kernel_trace = " in js
    with open(logdir + "gputrace.tmp", "a") as f:
        f.write("garbage line\n")
    with pytest.raises(SofaB200Error) as ei:
        sofa_preprocess(default_report_config(logdir))
    assert ei.value.row in (100, 103) and ei.value.row_code in (1, 2, 3)      # row 100 of the data rows (103 counting the header rows)


def test_empty_and_unterminated_inputs(tmp_path):
    logdir = os.path.join(str(tmp_path), "l") + "/"
    b = os.path.join(str(tmp_path), "o") + "/"
    for d in (logdir, b):
        synth.make_logdir(d, seed=3, n_perf=50, n_net=40)
        for fn in ("perf.script", "net.tmp"):         # drop the final newline
            with open(d + fn, "rb") as f:
                t = f.read()
            with open(d + fn, "wb") as f:
                f.write(t[:-1])
    cfg = default_report_config(logdir)
    sofa_preprocess(cfg)
    orc.preprocess(orc.OracleConfig(b, nprocs=1), demangle=demangle)
    for fn in ("cputrace.csv", "nettrace.csv", "report.js"):
        with open(logdir + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("unterminated", fn, f.read(), g.read())


@pytest.mark.parametrize("name", [n for n in golden_cases() if golden_bytes(n, "gputrace.csv")])
def test_gpu_profile_features_match_reference(name, tmp_path, capsys):
    from helpers import load_case
    from sofa_b200.sofa_analyze import sofa_analyze
    case, logdir = build_case_logdir(name, tmp_path)
    ov = dict(case["cfg"])
    cfg = default_report_config(logdir, cpu_filters=ov.pop("cpu_filters", None),
                                cuda_api_tracing=ov.pop("cuda_api_tracing", False))
    for k, v in ov.items():
        setattr(cfg, k, v)
    sofa_preprocess(cfg)
    feats = dict(sofa_analyze(cfg))
    for k, v in case["features"].items():
        if k == "__net_features__":
            assert float(feats["net_time"]) == v[1][1]
            with open(logdir + "netrank.csv", "rb") as f:
                _diff(name, "netrank.csv", f.read(), golden_bytes(name, "netrank.csv"))
        elif k not in ("elapsed_time", "__top10_names__"):
            assert float(feats[k]) == v, (k, feats[k], v)
    assert "Complete!!" in capsys.readouterr().out
    with open(logdir + "comm.csv", "rb") as f:
        _diff(name, "comm.csv", f.read(), golden_bytes(name, "comm.csv"))


def test_gpu_profile_matches_oracle_on_larger_log(tmp_path):
    import pandas as pd
    from sofa_b200.sofa_analyze import sofa_analyze
    a = os.path.join(str(tmp_path), "gpu") + "/"
    synth.make_logdir(a, seed=202, n_gpu=150000, units="ms_us_KB", p2p_cols=True)
    cfg = default_report_config(a)
    sofa_preprocess(cfg)
    feats = dict(sofa_analyze(cfg))
    df = pd.read_csv(a + "gputrace.csv")
    want = dict(orc.gpu_profile(df))
    for k, v in want.items():
        assert float(feats[k]) == float(v), (k, feats[k], v)


def test_device_synth_matches_python_generator():
    from sofa_b200.engine import CudaEngine
    eng = CudaEngine(0)
    try:
        for kind, variant, ref in [(0, 0, lambda n: synth.perf_text(77, n)),
                                   (0, 1, lambda n: synth.perf_text(77, n, True)),
                                   (1, 0, lambda n: synth.gpu_text(77, n)),
                                   (1, 1, lambda n: synth.gpu_text(77, n, "ms_us_KB", True)),
                                   (2, 1, lambda n: synth.net_text(77, n, 1))]:
            n = 20011
            nb = eng.synth(kind, 77, n, variant)
            got = eng.text_fetch(kind, nb).tobytes()
            assert got == ref(n).encode(), (kind, variant)
    finally:
        eng.close()


def test_top_k_and_cpu_profile_match_oracle(tmp_path):
    """Device radix sort + segmented pairwise sums vs pandas groupby / sort_values."""
    import pandas as pd
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import hostlogic as hl
    from sofa_b200.sofa_analyze import load_gpu_source, load_perf_source
    a = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(a, seed=303, n_perf=60000, n_gpu=90000, perf_events="instructions")
    cfg = default_report_config(a)
    sofa_preprocess(cfg)
    eng = CudaEngine(0)
    try:
        assert load_gpu_source(cfg, eng) and load_perf_source(cfg, eng)
        df = pd.read_csv(a + "gputrace.csv")
        want = orc.top_k_events(df, 10)
        got = eng.top_k(hl.SRC_GPU, 10)
        assert [g[1] for g in got] == [float(v) for v in want.values]
        # names: compare where sums are not tied with a neighbour
        wv = list(want.values)
        for i, (nm, v) in enumerate(got):
            if wv.count(v) == 1:
                assert nm == want.index[i]
        dfc = pd.read_csv(a + "cputrace.csv")
        total, top = orc.cpu_profile(dfc)
        gtotal, rows = eng.cpu_profile(20)
        assert gtotal == float(total)
        assert [r[2] for r in rows] == [float(x) for x in top['duration'].values]
        wd = list(top['duration'].values)
        for i, (row, ts, dur, name) in enumerate(rows):
            assert dur == float(dfc['duration'].iloc[row]) and ts == float(dfc['timestamp'].iloc[row])
            assert name == dfc['name'].iloc[row]
            if wd.count(dur) == 1:
                assert row == top.index[i]
    finally:
        eng.close()


def test_name_group_hash_collisions_fail_loudly(tmp_path):
    """get_top_k_events / nccl_time group rows by a 64-bit hash of the name; the verification pass over the sorted
    rows must refuse to merge two different names that share a hash.  Collisions are forced by narrowing the hash
    (test hook), the same run with the full hash passes."""
    import ctypes
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import hostlogic as hl, _lib as L
    from sofa_b200.sofa_analyze import load_gpu_source, load_perf_source
    a = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(a, seed=77, n_perf=5000, n_gpu=20000)
    cfg = default_report_config(a)
    sofa_preprocess(cfg)
    eng = CudaEngine(0)
    try:
        assert load_gpu_source(cfg, eng) and load_perf_source(cfg, eng)
        ok = eng.top_k(hl.SRC_GPU, 5)
        gp = eng.gpu_profile()
        eng._check(eng.lib.sofa_b200_set_debug(eng.ctx, L.SOFA_DEBUG_NAME_HASH_MASK, ctypes.c_uint64(0xF)))
        for call in (lambda: eng.top_k(hl.SRC_GPU, 5), lambda: eng.top_k(hl.SRC_PERF, 5), eng.gpu_profile):
            with pytest.raises(L.SofaB200Error) as ei:
                call()
            assert "share one 64-bit hash" in str(ei.value)
        eng._check(eng.lib.sofa_b200_set_debug(eng.ctx, L.SOFA_DEBUG_NAME_HASH_MASK, ctypes.c_uint64(0xFFFFFFFFFFFFFFFF)))
        assert eng.top_k(hl.SRC_GPU, 5) == ok
        assert eng.gpu_profile().nccl_time == gp.nccl_time
    finally:
        eng.close()


def test_verbose_analyze_runs(tmp_path, capsys):
    from sofa_b200.sofa_analyze import sofa_analyze
    a = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(a, seed=304, n_perf=3000, n_gpu=5000, units="ms_us_KB", p2p_cols=True)
    cfg = default_report_config(a)
    cfg.verbose = True
    sofa_preprocess(cfg)
    sofa_analyze(cfg)
    out = capsys.readouterr().out
    assert "Data-copy payload matrix" in out and "Top 10 Events" in out and "total execution time" in out


def test_async_transfers_give_identical_outputs(tmp_path):
    """upload_async / fetch_async (copy streams, double-buffered outputs) vs the synchronous calls."""
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import hostlogic as hl
    eng = CudaEngine(0)
    try:
        text = synth.perf_text(12, 40000).encode()
        p = hl.base_params(hl.SRC_PERF, 1600000000.0, False, 1599995000.75)
        pin = eng.pinned(len(text))
        pin.array[:] = np.frombuffer(text, dtype=np.uint8)
        eng.load(hl.SRC_PERF, text, p)
        want_csv = eng.csv(hl.SRC_PERF, 1).copy()
        eng.set_demangled([(m, demangle(m)) for m in eng.viz_mangled(1)])
        want_js, _ = eng.json(hl.SRC_PERF, 1, -1)
        want_js = want_js.copy()
        outs = [eng.pinned(len(want_js) + 64) for _ in range(4)]
        for rep in range(4):
            if rep == 2:
                eng.set_output_arena(3 * (len(want_js) + len(want_csv)))      # small: forces a wrap
            eng.upload_async(hl.SRC_PERF, pin.array, 0)
            assert eng.parse(hl.SRC_PERF, p) == 40000
            n0 = eng.csv_device(hl.SRC_PERF, 1)
            t0 = eng.fetch_async(0, n0, outs[0].array)
            eng.set_demangled([(m, demangle(m)) for m in eng.viz_mangled(1)])
            n1, _ = eng.json_device(hl.SRC_PERF, 1, -1)
            t1 = eng.fetch_async(1, n1, outs[1].array)
            n2 = eng.csv_device(hl.SRC_PERF, 1)          # second emit into the other half
            t2 = eng.fetch_async(0, n2, outs[2].array)
            n3, _ = eng.json_device(hl.SRC_PERF, 1, -1)
            t3 = eng.fetch_async(1, n3, outs[3].array)
            eng.sync()
            assert n0 == n2 == len(want_csv) and n1 == n3 == len(want_js)
            for o, w in ((outs[0], want_csv), (outs[2], want_csv), (outs[1], want_js), (outs[3], want_js)):
                assert np.array_equal(o.array[:len(w)], w)
    finally:
        eng.close()


def _side_csvs(logdir, seed, n_samples, n_cores, n_gpus, rate=10):
    """Side CSVs in the format the reference's preprocess writes (see tests/golden/side_small)."""
    import random
    rng = random.Random(seed)
    hdr = "timestamp,event,duration,deviceId,copyKind,payload,bandwidth,pkt_src,pkt_dst,pid,tid,name,category\n"
    dt = 1.0 / rate
    with open(logdir + "mpstat.csv", "w") as f:
        f.write(hdr)
        for s in range(n_samples):
            for c in range(n_cores):
                r = [rng.randint(0, 60), rng.randint(0, 25), rng.randint(0, 99), rng.randint(0, 9), rng.randint(0, 3)]
                f.write("%.6f,-1,%.6f,%.6f,-1,-1,-1,-1,-1,-1,-1,mpstat_core%d (usr|sys|idl|iow|irq): |%3d|%3d|%3d|%3d|%3d|,0\n"
                        % (0.6 + s * dt, rng.random(), c, c, *r))
    with open(logdir + "nvsmi_trace.csv", "w") as f:
        f.write(hdr)
        for ev in (0, 1):
            for s in range(n_samples):
                for g in range(n_gpus):
                    v = rng.randint(0, 100)
                    f.write("%.6f,%d,%d,%d,-1,-1,-1,-1,-1,-1,-1,GPUID_%s=%d_%d,%d\n"
                            % (0.513 + s * dt, ev, v, g, "sm" if ev == 0 else "mem", g, v, g))
    with open(logdir + "netstat.csv", "w") as f:
        f.write(hdr)
        for ev in (0, 1):
            for s in range(n_samples):
                bw = rng.random() * 1e7
                f.write("%.6f,%d,-1,-1,-1,-1,%.6f,-1,-1,-1,-1,network_bandwidth_%s(bytes):%d,0\n"
                        % (0.531 + s * dt * (1 + 0.3 * (s % 3 == 0)), ev, bw, "tx" if ev == 0 else "rx", int(bw)))
    with open(logdir + "cputrace.csv", "w") as f:
        f.write(hdr + "0.750014,14.624498,0.000178,-1,-1,0,0,-1,-1,4242,4255,x,0\n")
    with open(logdir + "misc.txt", "w") as f:
        f.write("elapsed_time 12.5\ncores 4\nvcores 4\npid 1\n")


@pytest.mark.parametrize("name", sorted(d for d in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if d.startswith("side_")))
def test_concurrency_breakdown_device_matches_reference(name, tmp_path):
    from test_side_host import run_case
    from sofa_b200.engine import CudaEngine
    eng = CudaEngine(0)
    try:
        run_case(name, tmp_path, eng)
    finally:
        eng.close()


def test_concurrency_breakdown_device_matches_oracle_larger(tmp_path):
    import pandas as pd
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import sofa_analyze as an
    logdir = str(tmp_path) + "/"
    _side_csvs(logdir, seed=5, n_samples=4000, n_cores=16, n_gpus=8)
    cfg = default_report_config(logdir)
    cfg.is_idle_threshold = 30
    eng = CudaEngine(0)
    try:
        feats = an.concurrency_breakdown(logdir, cfg, eng, [])
    finally:
        eng.close()
    got_perf = open(logdir + "performance.csv", "rb").read()
    ofeats, operf = orc.concurrency_breakdown(pd.read_csv(logdir + "mpstat.csv"), pd.read_csv(logdir + "cputrace.csv"),
                                              pd.read_csv(logdir + "nvsmi_trace.csv"), pd.read_csv(logdir + "netstat.csv"),
                                              is_idle_threshold=30)
    import io
    buf = io.StringIO()
    operf.to_csv(buf)
    assert got_perf == buf.getvalue().encode()
    assert [(n, float(v)) for n, v in feats] == [(n, float(v)) for n, v in ofeats]


def test_tile_boundaries_with_ragged_line_lengths(tmp_path):
    """Lines from 20 bytes to ~6 KB (long kernel / symbol names) so that tile edges, the overlap
    window, unstaged emit blocks and multi-word spans are all exercised; byte-compare with the oracle."""
    import random
    rng = random.Random(99)
    a = os.path.join(str(tmp_path), "gpu") + "/"
    b = os.path.join(str(tmp_path), "cpu") + "/"
    for d in (a, b):
        synth.write_side_files(d)
    hdr = synth.gpu_header()
    glines, plines = [], []
    for i in range(6000):
        base = synth.gpu_line(7, i)
        if i % 3 == 0 and '"[' not in base:
            L = rng.choice([1, 7, 8, 9, 63, 64, 65, 500, 3000, 6000])
            nm = "".join(rng.choice("abcdefgh_<>():,") for _ in range(L))
            base = base[:base.rindex(',"') + 2] + nm + '"\n'
        glines.append(base)
        pl = synth.perf_line(7, i)
        if i % 5 == 0:
            L = rng.choice([1, 15, 16, 17, 200, 2500])
            tok = pl.split()
            tok[5] = "S" * L + "+0x1f"
            pl = " ".join(tok) + "\n"
        plines.append(pl)
    for d in (a, b):
        with open(d + "gputrace.tmp", "w") as f:
            f.write(hdr + "".join(glines))
        open(d + "gputrace1.nvvp", "w").close()
        with open(d + "perf.script", "w") as f:
            f.write("".join(plines))
    cfg = default_report_config(a)
    cfg.plot_ratio = 3
    sofa_preprocess(cfg)
    orc.preprocess(orc.OracleConfig(b, plot_ratio=3, nprocs=1), demangle=demangle)
    for fn in ("cputrace.csv", "gputrace.csv", "report.js"):
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("ragged", fn, f.read(), g.read())


def test_line_longer_than_overlap_window_fails_loudly(tmp_path):
    from sofa_b200._lib import SofaB200Error
    logdir = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(logdir, seed=4, n_perf=3000)
    with open(logdir + "perf.script") as f:
        lines = f.readlines()
    tok = lines[1500].split()
    tok[5] = "S" * 20000 + "+0x1"
    lines[1500] = " ".join(tok) + "\n"
    with open(logdir + "perf.script", "w") as f:
        f.write("".join(lines))
    with pytest.raises(SofaB200Error) as ei:
        sofa_preprocess(default_report_config(logdir))
    assert ei.value.row_code in (6,) and ei.value.row == 1500


# ---------------------------------------------------------------- N1: recorder side files on the device
def test_side_sources_device_reproduce_reference_files(tmp_path):
    from helpers import side_cases, build_side_logdir, SIDE_OUTPUTS, preprocess_overrides
    for name in side_cases():
        case, logdir = build_side_logdir(name, tmp_path / name)
        cfg = default_report_config(logdir)
        for k, v in preprocess_overrides(case).items():
            setattr(cfg, k, v)
        sofa_preprocess(cfg)
        for fn in SIDE_OUTPUTS:
            want = golden_bytes(name, fn)
            if want is None:
                assert not os.path.isfile(logdir + fn), fn
                continue
            with open(logdir + fn, "rb") as f:
                _diff(name, fn, f.read(), want)


def test_side_sources_device_match_oracle_larger_and_edge_cases(tmp_path):
    """40 k mpstat lines (64 cores), 20 k nvidia-smi lines, 5 k netstat lines - several tiles and emit
    blocks per file - plus malformed mpstat lines and a time-zone shift; byte-compare with the oracle."""
    import shutil
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    a = os.path.join(str(tmp_path), "gpu") + "/"
    synth.make_logdir(a, seed=11, n_perf=500)
    synth.write_side_raw(a, seed=12, n_samples=620, n_cores=64, n_gpus=8)
    synth.write_side_raw_1hz(a, seed=13, n_seconds=3000, n_gpus=8)
    with open(a + "nvsmi_query.txt", "a") as f:             # more nvidia-smi lines than samples x gpus
        body = open(a + "nvsmi_query.txt").read().split("\n")[1:-1]
        for rep in range(3):
            f.write("\n".join(body) + "\n")
    lines = open(a + "mpstat.txt").read().split("\n")
    lines.insert(1000, "1600000004.1,2,77")
    lines.insert(2000, "")
    lines[3000] += ",1"
    open(a + "mpstat.txt", "w").write("\n".join(lines))
    b = os.path.join(str(tmp_path), "cpu") + "/"
    shutil.copytree(a, b)
    cfg = default_report_config(a)
    cfg.nvsmi_time_zone = -5
    cfg.enable_vmstat = cfg.enable_encode_decode = True
    sofa_preprocess(cfg)
    orc.preprocess(orc.OracleConfig(b, nprocs=1, nvsmi_time_zone=-5, enable_vmstat=True, enable_encode_decode=True), demangle=demangle)
    for fn in ("mpstat.csv", "usr_sys.csv", "nvsmi_trace.csv", "netstat.csv", "netbandwidth.csv", "vmstat.csv", "report.js"):
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("side_large", fn, f.read(), g.read())


@pytest.mark.parametrize("ratio,absolute", [(1, False), (3, True)])
def test_storage_sources_device_match_oracle_larger(ratio, absolute, tmp_path):
    """diskstat.txt / blktrace.txt / strace.txt / pystacks.txt (N1) on the device at sizes that span many blocks (scan
    partitions, the radix sort behind the 'D' -> 'C' links, the device-name table) vs the oracle, byte for byte."""
    import shutil
    from helpers import make_storage_logdir, STORAGE_OUTPUTS
    a = os.path.join(str(tmp_path), "gpu") + "/"
    make_storage_logdir(a)
    b = os.path.join(str(tmp_path), "cpu") + "/"
    shutil.copytree(a, b)
    cfg = default_report_config(a)
    cfg.plot_ratio, cfg.absolute_timestamp, cfg.diskstat_filters = ratio, absolute, "xda1"
    sofa_preprocess(cfg)
    orc.preprocess(orc.OracleConfig(b, nprocs=1, plot_ratio=ratio, absolute_timestamp=absolute, diskstat_filters="xda1"), demangle=demangle)
    for fn in STORAGE_OUTPUTS:
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("storage_large", fn, f.read(), g.read())


def test_storage_sources_device_errors(tmp_path):
    """what the reference dies on must fail loudly here too: a diskstat line with too few fields, an unpaired pystacks stamp"""
    from helpers import build_side_logdir
    from sofa_b200._lib import SofaB200Error
    _, d = build_side_logdir("disk_small", tmp_path / "short")
    lines = open(d + "diskstat.txt").read().split("\n")
    lines[7] = "1600000000.7,sda,1,2"
    open(d + "diskstat.txt", "w").write("\n".join(lines))
    with pytest.raises(SofaB200Error):
        sofa_preprocess(default_report_config(d))
    _, d = build_side_logdir("oracle_strace_blk", tmp_path / "odd")
    with open(d + "pystacks.txt", "a") as f:
        f.write("1600000009000000\n")
    with pytest.raises(SofaB200Error):
        sofa_preprocess(default_report_config(d))


def test_side_sources_device_errors_and_no_data(tmp_path):
    from helpers import build_side_logdir
    from sofa_b200._lib import SofaB200Error
    _, d = build_side_logdir("side_small", tmp_path / "dup")
    lines = open(d + "netstat.txt").read().split("\n")
    lines[3] = lines[2]                                      # ZeroDivisionError in the reference
    open(d + "netstat.txt", "w").write("\n".join(lines))
    with pytest.raises(SofaB200Error):
        sofa_preprocess(default_report_config(d))
    _, d = build_side_logdir("side_small", tmp_path / "failed")
    with open(d + "nvsmi_query.txt", "a") as f:
        f.write("Unable to determine the device handle for GPU 0000:3B:00.0: Unknown Error (Failed)\n")
    sofa_preprocess(default_report_config(d))
    assert not os.path.isfile(d + "nvsmi_trace.csv")
    js = open(d + "report.js").read()
    assert "nvsmi_sm =" not in js and "mpstat_usr =" in js and "tx_bandwidth =" in js


@pytest.mark.parametrize("name", sorted(d for d in os.listdir(os.path.join(os.path.dirname(__file__), "golden")) if d.startswith("side_")))
def test_nvsmi_profile_device_matches_reference(name, tmp_path, capsys):
    from test_side_host import run_nvsmi_case
    from sofa_b200.engine import CudaEngine
    eng = CudaEngine(0)
    try:
        run_nvsmi_case(name, tmp_path, eng, capsys)
    finally:
        eng.close()


def test_nvsmi_profile_device_matches_oracle_larger_and_roi(tmp_path):
    """200 k rows, 8 devices whose first appearance is not in id order, negative and large durations,
    with and without the spotlight region of interest."""
    import random
    import pandas as pd
    from sofa_b200.engine import CudaEngine
    from sofa_b200 import sofa_analyze as an
    rng = random.Random(7)
    logdir = str(tmp_path) + "/"
    devs = [5, 1, 9, 0, 12, 3, 8, 2]
    with open(logdir + "nvsmi_trace.csv", "w") as f:
        f.write("timestamp,event,duration,deviceId,copyKind,payload,bandwidth,pkt_src,pkt_dst,pid,tid,name,category\n")
        for ev in (0, 1):
            for s_ in range(12500):
                for d in devs:
                    v = rng.choice([rng.randrange(0, 101), rng.randrange(-5, 3), rng.randrange(0, 1000000)])
                    f.write("%.6f,%d,%d,%d,-1,-1,-1,-1,-1,-1,-1,GPUID_x=%d_%d,%d\n" % (0.5 + s_ * 0.1, ev, v, d, d, v, d))
    df = pd.read_csv(logdir + "nvsmi_trace.csv")
    eng = CudaEngine(0)
    try:
        for roi in (None, (100.0, 733.3)):
            cfg = default_report_config(logdir)
            if roi:
                cfg.spotlight_gpu, cfg.roi_begin, cfg.roi_end = True, roi[0], roi[1]
            got = an.nvsmi_profile(logdir, cfg, eng, [])
            want = orc.nvsmi_profile(df, spotlight_gpu=bool(roi), roi_begin=roi[0] if roi else 0, roi_end=roi[1] if roi else 0)
            assert [(n, float(v)) for n, v in got] == [(n, float(v)) for n, v in want]
    finally:
        eng.close()


@pytest.mark.parametrize("name", ["net_small", "net_nan"])
def test_net_profile_device_matches_reference(name, tmp_path):
    from helpers import load_case
    from sofa_b200.sofa_analyze import sofa_analyze
    case, logdir = build_case_logdir(name, tmp_path)
    cfg = default_report_config(logdir)
    for k, v in case["cfg"].items():
        setattr(cfg, k, v)
    sofa_preprocess(cfg)
    feats = dict(sofa_analyze(cfg))
    assert float(feats["net_time"]) == case["features"]["__net_features__"][1][1]
    with open(logdir + "netrank.csv", "rb") as f:
        _diff(name, "netrank.csv", f.read(), golden_bytes(name, "netrank.csv"))


def test_net_profile_device_matches_oracle_larger(tmp_path):
    import pandas as pd
    from sofa_b200.sofa_analyze import sofa_analyze
    for ratio, pct in ((1, 0), (3, 2)):
        a = os.path.join(str(tmp_path), "r%d" % ratio) + "/"
        synth.make_logdir(a, seed=300 + ratio, n_net=120000, non_ip_pct=pct)
        cfg = default_report_config(a)
        cfg.plot_ratio = ratio
        sofa_preprocess(cfg)
        feats = dict(sofa_analyze(cfg))
        got = open(a + "netrank.csv", "rb").read()
        want = orc.net_profile(pd.read_csv(a + "nettrace.csv"), a)
        assert float(feats["net_time"]) == float(dict(want)["net_time"])
        assert got == open(a + "netrank.csv", "rb").read()


def test_net_profile_name_order_with_mixed_address_lengths(tmp_path):
    """net_time accumulates the group sums in sorted-NAME order; the device sorts by string-order keys of
    (src, dst, payload).  Addresses and payloads of different digit counts make numeric and string order differ."""
    import random
    import pandas as pd
    from sofa_b200.sofa_analyze import sofa_analyze
    rng = random.Random(5)
    a = str(tmp_path) + "/"
    synth.make_logdir(a, seed=9, n_net=10)
    hosts = ["1.2.3.4", "10.0.0.1", "9.255.1.77", "100.1.2.3", "2.0.0.200", "192.168.1.1", "19.216.81.1", "1.0.0.9"]
    t = 1600000000.5
    with open(a + "net.tmp", "w") as f:
        for i in range(60000):
            t += 0.00001
            s_, d_ = rng.choice(hosts), rng.choice(hosts)
            f.write("%.6f IP %s.%d > %s.%d: tcp %d\n" % (t, s_, rng.randrange(1, 65535), d_, rng.randrange(1, 65535),
                                                        rng.choice([0, 7, 64, 640, 1448, 14480, 65000, rng.randrange(1, 99999)])))
    cfg = default_report_config(a)
    sofa_preprocess(cfg)
    feats = dict(sofa_analyze(cfg))
    got = open(a + "netrank.csv", "rb").read()
    want = dict(orc.net_profile(pd.read_csv(a + "nettrace.csv"), a))
    assert float(feats["net_time"]) == float(want["net_time"])
    assert got == open(a + "netrank.csv", "rb").read()


def test_cli_report_then_standalone_analyze(tmp_path):
    """`bin/sofa report` on a logdir with the trace sources and all five side files, then `bin/sofa analyze` on its
    own (re-ingesting the dumped text): same feature table both times, files identical to the oracle's."""
    import subprocess
    import sys
    import time
    import pandas as pd
    os.environ['TZ'] = 'UTC'
    time.tzset()
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = os.path.join(str(tmp_path), "gpu") + "/"
    synth.make_logdir(a, seed=41, n_perf=3000, n_gpu=2000, n_net=1500, non_ip_pct=1)
    synth.write_side_raw(a, seed=42, n_samples=80, n_cores=4, n_gpus=2)
    synth.write_side_raw_1hz(a, seed=43, n_seconds=20, n_gpus=2)
    b = os.path.join(str(tmp_path), "cpu") + "/"
    import shutil
    shutil.copytree(a, b)
    env = dict(os.environ, TZ='UTC')
    out1 = subprocess.run([sys.executable, os.path.join(repo, "bin", "sofa"), "report", "--logdir", a, "--enable_vmstat",
                           "--enable_encode_decode"], capture_output=True, text=True, env=env, timeout=600)
    assert out1.returncode == 0, out1.stderr[-2000:]
    assert "Complete!!" in out1.stdout
    out2 = subprocess.run([sys.executable, os.path.join(repo, "bin", "sofa"), "analyze", "--logdir", a], capture_output=True,
                          text=True, env=env, timeout=600)
    assert out2.returncode == 0, out2.stderr[-2000:]
    table = lambda s_: [ln for ln in s_.split("\n") if ln[:1].isdigit()]
    assert table(out1.stdout) == table(out2.stdout) and len(table(out1.stdout)) > 20
    orc.preprocess(orc.OracleConfig(b, nprocs=1, enable_vmstat=True, enable_encode_decode=True), demangle=demangle)
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "mpstat.csv", "vmstat.csv", "nvsmi_trace.csv", "netstat.csv",
               "usr_sys.csv", "netbandwidth.csv", "report.js"):
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            _diff("cli", fn, f.read(), g.read())
    want = dict(orc.gpu_profile(pd.read_csv(b + "gputrace.csv"), b))
    want.update(dict(orc.net_profile(pd.read_csv(b + "nettrace.csv"), b)))
    want.update(dict(orc.nvsmi_profile(pd.read_csv(b + "nvsmi_trace.csv"), nvsmi_data=True)))
    got = {}
    for ln in table(out1.stdout):
        f = ln.split()
        got[f[1]] = float(f[2])
    for k, v in want.items():
        assert abs(got[k] - float(v)) <= 0.0005 + 1e-9 * abs(float(v)), (k, got[k], v)       # the table prints %.3lf


# ---------------------------------------------------------------------------- BASELINE config sizes (slow)
def _full_pass_parity(tmp_path, **kw):
    """One logdir through the oracle (all host cores) and through sofa_preprocess + sofa_analyze; every output file the
    reference would write must be byte-identical (bench.py's parity_sample helpers)."""
    import shutil
    import multiprocessing as mp
    import bench
    a = os.path.join(str(tmp_path), "oracle") + "/"
    b = os.path.join(str(tmp_path), "cuda") + "/"
    os.makedirs(a)
    synth.make_logdir(a, **kw)
    synth.write_side_raw(a, kw["seed"], n_samples=200, n_cores=16, n_gpus=4)
    shutil.copytree(a, b)
    bench.oracle_pass(a, mp.cpu_count())
    bench.product_pass(b, 0)
    files = bench.parity_sample(a, b)
    assert "report.js" in files
    return files


def test_config1_one_million_perf_events(tmp_path):
    """BASELINE configs[0]: 1 M-event perf-only trace, full preprocess + analyze, byte-exact."""
    files = _full_pass_parity(tmp_path, seed=20260921 + 1, n_perf=1000000)
    assert "cputrace.csv" in files and "performance.csv" in files


def test_config2_slice_two_million_gpu_events(tmp_path):
    """A 2 M-row nvprof slice of BASELINE configs[1] (+ 100 k perf and packets so every analyze function runs)."""
    files = _full_pass_parity(tmp_path, seed=20260921 + 2, n_perf=100000, n_gpu=2000000, n_net=100000)
    for fn in ("gputrace.csv", "comm.csv", "nettrace.csv", "netrank.csv"):
        assert fn in files


# ---------------------------------------------------------------------------- the whole analyze driver vs the reference
DRIVER_CASES = sorted(d for d in os.listdir(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")) if d.startswith("driver_"))


@pytest.mark.parametrize("name", DRIVER_CASES)
@pytest.mark.parametrize("variant", ["plain", "verbose", "verbose_from_csv", "plain_resident"])
def test_analyze_driver_stdout_matches_reference(name, variant, tmp_path, capsys):
    """sofa_preprocess + sofa_analyze on the inputs of tests/golden/driver_*: the whole stdout of the analyze driver
    (feature table; with --verbose every profile function's tables) and netrank.csv / performance.csv / comm.csv equal
    the unmodified reference driver's (tests/golden/make_golden_driver.py), byte for byte."""
    import json
    import shutil
    from sofa_b200.sofa_analyze import sofa_analyze
    import sofa_b200.sofa_preprocess as sp
    sp.SYS_BLOCK = "/nonexistent_sofa_b200/"      # diskstat sector sizes: the fixture's device names exist nowhere
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    meta = json.load(open(os.path.join(gdir, "case.json")))
    a = os.path.join(str(tmp_path), "sofalog") + "/"
    os.makedirs(a)
    for fn in os.listdir(gdir):
        if fn.endswith(".txt") and not fn.startswith("analyze_") or fn.endswith((".tmp", ".script")):
            shutil.copy(os.path.join(gdir, fn), a + fn)
    if os.path.isfile(a + "gputrace.tmp"):
        open(a + "gputrace1.nvvp", "w").close()
    if os.path.isfile(a + "net.tmp"):
        open(a + "sofa.pcap", "w").close()
    ov = dict(meta["cfg"])
    cfg = default_report_config(a, cpu_filters=ov.pop("cpu_filters", None))
    for k, v in ov.items():
        setattr(cfg, k, v)
    if variant == "plain_resident":
        cfg.keep_engine = True        # `sofa report`: analyze the rows preprocess left on the device
    sofa_preprocess(cfg)
    capsys.readouterr()
    cfg.verbose = variant.startswith("verbose")
    cfg.nvsmi_data = False            # the golden driver run starts from a fresh cfg (as a stand-alone `sofa analyze` does)
    if variant == "verbose_from_csv":
        # the CSV resume contract (`sofa report --skip_preprocess`, bin/sofa:90,369-371): nothing but the CSVs is left
        for fn in ("perf.script", "gputrace.tmp", "net.tmp"):
            if os.path.isfile(a + fn):
                os.remove(a + fn)
    variant = "verbose" if variant.startswith("verbose") else "plain"
    sofa_analyze(cfg)
    got = capsys.readouterr().out.replace(a, "LOGDIR/").replace(a.rstrip("/"), "LOGDIR")
    want = open(os.path.join(gdir, "analyze_%s.txt" % variant)).read()
    assert cfg.roi_begin == meta["variants"][variant]["roi_begin"] and cfg.roi_end == meta["variants"][variant]["roi_end"]
    def canon(text):
        """cpu_profile's table is `sort_values('duration')` with numpy's unstable sort: rows of equal duration come in a
        machine-dependent order in the reference itself (SURVEY.md §7 "sort ties"), and which of them make the top 20
        sets the width of the name column.  Keep what is pinned: the (ratio, duration) columns."""
        out, in_tab = [], False
        for ln in text.split("\n"):
            if ln.lstrip().startswith("timestamp") and "ratio(%)" in ln:
                in_tab = True
                out.append("timestamp ratio(%) duration name")
                continue
            if in_tab:
                tok = ln.split()
                if len(tok) >= 4 and not ln.startswith("\x1b"):
                    out.append(" ".join(tok[1:3]))
                    continue
                in_tab = False
            out.append(ln)
        return "\n".join(out)
    got, want = canon(got), canon(want)
    if got != want:
        gl, wl = got.split("\n"), want.split("\n")
        for i, (x, y) in enumerate(zip(gl, wl)):
            assert x == y, "line %d:\n got %r\nwant %r" % (i + 1, x[:300], y[:300])
        assert len(gl) == len(wl), (len(gl), len(wl), gl[-3:], wl[-3:])
    for fn in ("netrank.csv", "performance.csv", "comm.csv"):
        if os.path.isfile(os.path.join(gdir, fn)):
            _diff(name, fn, open(a + fn, "rb").read(), open(os.path.join(gdir, fn), "rb").read())


def test_network_filters_series(tmp_path):
    """--network_filters (sofa_preprocess.py:1214-1231): to_/from_ series of the node of interest vs the oracle's frame
    selection; the colours are random in the reference and are not compared."""
    import json
    import re
    import pandas as pd
    a = os.path.join(str(tmp_path), "l") + "/"
    synth.make_logdir(a, seed=909, n_net=4000, non_ip_pct=2)
    cfg = default_report_config(a)
    cfg.plot_ratio = 2
    df0 = None
    cfg.net_filters = []
    sofa_preprocess(cfg)
    df = pd.read_csv(a + "nettrace.csv")
    top = df.groupby(['pkt_src', 'pkt_dst']).size().sort_values(ascending=False).index[:6].tolist()
    node = int(top[0][0])
    peers = [int(d) for s, d in top if int(s) == node][:2] + [int(s) for s, d in top if int(d) == node][:2] + [123456789012]
    peers = list(dict.fromkeys(peers))
    cfg.net_filters = [str(node)] + [str(p) for p in peers]
    sofa_preprocess(cfg)
    js = open(a + "report.js").read()
    names = re.findall(r"^(\w+) = ", js, flags=re.M)
    for direction in ("to_", "from_"):
        for p in peers:
            nm = "%s%d" % (direction, p)
            sel = df[(df['payload'] > 0) & ((df['pkt_src'] == float(node)) & (df['pkt_dst'] == float(p)) if direction == "to_"
                                            else (df['pkt_dst'] == float(node)) & (df['pkt_src'] == float(p)))]
            if len(sel) == 0:
                assert nm not in names
                continue
            m = re.search(r"^%s = (\{.*\})$" % nm, js, flags=re.M)
            assert m, nm
            data = json.loads(m.group(1))
            assert data["name"] == "[keyword]" + nm
            # (nettrace.csv holds the "%.6f" text of the values the series carries at full precision: compare the keys
            #  that survive the round trip exactly, and the floats at the CSV's precision)
            assert [r["name"] for r in data["data"]] == sel["name"].tolist(), nm
            assert [r["payload"] for r in data["data"]] == sel["payload"].tolist(), nm
            assert [round(r["x"], 6) for r in data["data"]] == sel["timestamp"].tolist(), nm
    assert js.rstrip().endswith("]") and ("to_%d," % peers[0]) in js.split("sofa_traces = [ ")[1]


def test_sharded_trace_two_gpus(tmp_path):
    """One logdir preprocessed by two ranks on two GPUs (sofa_b200.sharded: byte-range shards, NCCL all-reduce of the dtype
    flags, ordered concat) and analysed on rank 0 after the parsed gpu / packet shards were gathered there over NCCL:
    every file equals the single-GPU run's, and so do the analyze features.  Needs 2 GPUs (skipped otherwise)."""
    import subprocess
    import sys
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import shutil
    a = os.path.join(str(tmp_path), "single") + "/"
    b = os.path.join(str(tmp_path), "sharded") + "/"
    synth.make_logdir(a, seed=606, n_perf=30000, n_gpu=24000, n_net=9000, non_ip_pct=1)
    synth.write_side_raw(a, 606, n_samples=40, n_cores=4, n_gpus=2)
    shutil.copytree(a, b)
    from sofa_b200.sofa_analyze import sofa_analyze
    cfg = default_report_config(a, cpu_filters="idle:black,tensorflow:orange")
    sofa_preprocess(cfg)
    feats = sofa_analyze(cfg)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(str(tmp_path), "feats.json")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), os.path.join(repo, "tests", "sharded_worker.py"), b, out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "comm.csv", "netrank.csv", "performance.csv",
               "mpstat.csv", "nvsmi_trace.csv", "netstat.csv"):
        _diff("sharded", fn, open(b + fn, "rb").read(), open(a + fn, "rb").read())
    got = json.load(open(out))
    assert [n for n, _ in feats] == [n for n, _ in got]
    for (n, v), (_, w) in zip(feats, got):
        assert float(v) == float(w) or (v != v and w != w), (n, v, w)


def test_segmented_preprocess_on_device(tmp_path, monkeypatch):
    """SOFA_B200_SEGMENT_BYTES smaller than the sources: sofa_preprocess walks them in segments (bounded device memory)
    and writes the same bytes as the one-piece run."""
    import shutil
    a = os.path.join(str(tmp_path), "whole") + "/"
    b = os.path.join(str(tmp_path), "segmented") + "/"
    synth.make_logdir(a, seed=717, n_perf=40000, n_gpu=30000, n_net=12000, non_ip_pct=1)
    synth.write_side_raw(a, 717, n_samples=30, n_cores=2, n_gpus=2)
    shutil.copytree(a, b)
    sofa_preprocess(default_report_config(a, cpu_filters="idle:black,tensorflow:orange"))
    monkeypatch.setenv("SOFA_B200_SEGMENT_BYTES", str(700000))
    sofa_preprocess(default_report_config(b, cpu_filters="idle:black,tensorflow:orange"))
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "mpstat.csv", "nvsmi_trace.csv", "netstat.csv"):
        _diff("segmented", fn, open(b + fn, "rb").read(), open(a + fn, "rb").read())
    assert not [f for f in os.listdir(b) if ".shard" in f]


@pytest.mark.parametrize("name", ["aisi_iters", "aisi_units"])
def test_iter_profile_matches_reference(name, tmp_path):
    """AISI per-iteration sums (sofa_aisi.py:21-59) on the device vs the unmodified reference's iter_profile on the same
    slices (tests/golden/make_golden_aisi.py): every field equal as floats; from the raw rows and from gputrace.csv."""
    import json
    import shutil
    from sofa_b200.engine import CudaEngine
    from sofa_b200.sofa_analyze import load_gpu_source
    from sofa_b200.sofa_aisi import iter_profile
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    meta = json.load(open(os.path.join(gdir, "case.json")))
    a = os.path.join(str(tmp_path), "l") + "/"
    os.makedirs(a)
    for fn in os.listdir(gdir):
        if fn != "case.json":
            shutil.copy(os.path.join(gdir, fn), a + fn)
    open(a + "gputrace1.nvvp", "w").close()
    cfg = default_report_config(a)
    sofa_preprocess(cfg)
    for from_csv in (True, False):
        if not from_csv:
            os.remove(a + "gputrace.csv")
        eng = CudaEngine(0)
        try:
            assert load_gpu_source(cfg, eng)
            got = iter_profile(cfg, eng, meta["times"])
        finally:
            eng.close()
        assert len(got) == len(meta["iterations"])
        for i, (g, w) in enumerate(zip(got, meta["iterations"])):
            assert (g is None) == (w is None), i
            if w is not None:
                for k in w:
                    assert g[k] == w[k], (from_csv, i, k, g[k], w[k])
