"""Fuzz of the ORACLE against the UNMODIFIED reference (oracle/_ref, one child interpreter per batch): one mutated line per
logdir; the oracle must accept / raise where the reference does and, when both accept, write identical files.  This is
what pins the oracle's behaviour on malformed input - e.g. that the reference swallows every exception of its perf section
("no perf traces.") but dies on a bad nvprof or tcpdump line, and converts diskstat fields only where a row uses them.
The product is fuzzed against the oracle in test_fuzz_rows_host.py / test_fuzz_side_host.py.
SOFA_FUZZ_TRIALS / SOFA_FUZZ_SEED widen the run (4000 side-file and 1700 trace-line mutations ran clean)."""
import os
import random
import shutil
import time

import pytest

import test_fuzz_rows_host as R
import test_fuzz_side_host as F
from helpers import GOLDEN, demangle
from oracle import ref_runner
from oracle import sofa_oracle as orc
from sofa_b200 import synth

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref is not built (needs /root/reference at build time)")
TRIALS = int(os.environ.get('SOFA_FUZZ_TRIALS', '6'))
SEED = int(os.environ.get('SOFA_FUZZ_SEED', '0'))


def compare(items, meta, outputs):
    res = ref_runner.run_batch_in_subprocess(items)
    agree_ok = 0
    for (d, over), what, r in zip(items, meta, res):
        b = d.replace("_ref/", "_orc/")
        try:
            orc.preprocess(orc.OracleConfig(b, nprocs=1, **over), demangle=demangle)
            o = {"ok": True}
        except BaseException as e:
            o = {"raised": type(e).__name__}
        assert ("ok" in r) == ("ok" in o), (what, r, o)
        if "ok" in r:
            for f in outputs:
                ea, eb = os.path.isfile(d + f), os.path.isfile(b + f)
                assert ea == eb, (what, f)
                if ea:
                    assert open(d + f, "rb").read() == open(b + f, "rb").read(), (what, f)
            agree_ok += 1
    return agree_ok


def test_oracle_vs_reference_side_file_fuzz(tmp_path):
    os.environ['TZ'] = 'UTC'
    time.tzset()
    cases = [("side_full", "mpstat.txt", ",", (), {}), ("side_full", "netstat.txt", ",", (), {}),
             ("side_full", "nvsmi_query.txt", ", ", (), {"nvsmi_time_zone": 3}),
             ("side_full", "vmstat.txt", None, (), {"enable_vmstat": True}),
             ("side_full", "nvsmi.txt", None, (), {"enable_encode_decode": True, "absolute_timestamp": True}),
             ("disk_small", "diskstat.txt", ",", (), {}), ("oracle_strace_blk", "blktrace.txt", None, (), {}),
             ("oracle_strace_blk", "strace.txt", None, ("vmstat.txt",), {"plot_ratio": 2}),
             ("oracle_strace_blk", "pystacks.txt", None, ("vmstat.txt",), {})]
    rng = random.Random(SEED)
    items, meta = [], []
    for ci, (case, fn, sep, extra, over) in enumerate(cases):
        base = os.path.join(GOLDEN, case)
        lines = open(os.path.join(base, fn)).read().splitlines(True)
        for t in range(TRIALS):
            ls = list(lines)
            i = rng.randrange(1 if fn in ("mpstat.txt", "nvsmi_query.txt") else 0, len(ls))
            ls[i] = F.mutate2(rng, ls[i], sep)
            for who in ("ref", "orc"):
                d = str(tmp_path / ("c%d_t%d_%s" % (ci, t, who))) + "/"
                os.makedirs(d)
                for m in F.MANDATORY + extra:
                    shutil.copy(os.path.join(base, m), d + m)
                open(d + fn, "w").write("".join(ls))
            items.append((str(tmp_path / ("c%d_t%d_ref" % (ci, t))) + "/", over))
            meta.append((fn, ls[i]))
    assert compare(items, meta, tuple(set(F.OUTPUTS) | set(F.OUTPUTS2))) >= len(items) // 4


def test_oracle_vs_reference_trace_line_fuzz(tmp_path):
    base = str(tmp_path / "base") + "/"
    synth.make_logdir(base, seed=400 + SEED, n_perf=40, n_gpu=40, n_net=40, non_ip_pct=5)
    rng = random.Random(SEED)
    items, meta = [], []
    for ci, (fn, first) in enumerate((("perf.script", 0), ("gputrace.tmp", 3), ("net.tmp", 0))):
        lines = open(base + fn).read().splitlines(True)
        for t in range(TRIALS + 2):
            ls = list(lines)
            i = rng.randrange(first, len(ls))
            ls[i] = F.mutate(rng, ls[i], ",") if fn == "gputrace.tmp" and rng.random() < 0.6 else R._mutate(rng, ls[i])
            over = dict(plot_ratio=rng.choice([1, 1, 2]), cpu_filters="idle:black,tensorflow:orange")
            for who in ("ref", "orc"):
                d = str(tmp_path / ("r%d_t%d_%s" % (ci, t, who))) + "/"
                shutil.copytree(base, d)
                open(d + fn, "w").write("".join(ls))
            items.append((str(tmp_path / ("r%d_t%d_ref" % (ci, t))) + "/", over))
            meta.append((fn, ls[i]))
    assert compare(items, meta, ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js")) >= 3


def test_oracle_vs_reference_over_report_options(tmp_path):
    """random combinations of the report options (plot_ratio, absolute_timestamp, cpu_time_offset, nvsmi_time_zone,
    enable_vmstat, enable_encode_decode, strace_min_time, diskstat_filters, cpu_filters) on a logdir holding every source:
    the oracle writes what the unmodified reference writes (270 combinations ran clean)"""
    import time
    from helpers import make_full_logdir, random_overrides, ALL_PREPROCESS_OUTPUTS
    os.environ['TZ'] = 'UTC'
    time.tzset()
    base = str(tmp_path / "base") + "/"
    make_full_logdir(base, SEED)
    rng = random.Random(SEED + 7)
    items, meta = [], []
    for t in range(TRIALS):
        over = random_overrides(rng)
        for who in ("ref", "orc"):
            shutil.copytree(base, str(tmp_path / ("o%d_%s" % (t, who))) + "/")
        items.append((str(tmp_path / ("o%d_ref" % t)) + "/", over))
        meta.append(over)
    assert compare(items, meta, ALL_PREPROCESS_OUTPUTS) == len(items)
