"""Fuzz of the recorder side-file parsers on the CPU build of the device row functions (side.cuh): one mutated line
per trial; the product (host logic + hostsim engine) must accept / reject like the reference-path oracle and, when
both accept, write identical files.  Loud refusals of spellings the device codecs do not carry are allowed."""
import os
import random
import shutil

import pytest

from helpers import GOLDEN, demangle
from hostsim_engine import HostSimEngine, HostSimError
from oracle import sofa_oracle as orc
from sofa_b200.hostlogic import PreprocessError
from sofa_b200.sofa_config import default_report_config
from sofa_b200.sofa_preprocess import sofa_preprocess

BASE = os.path.join(GOLDEN, "side_full")
MANDATORY = ("misc.txt", "sofa_time.txt", "perf_events_used.txt", "perf_timebase.txt")
OUTPUTS = ("mpstat.csv", "usr_sys.csv", "vmstat.csv", "nvsmi_trace.csv", "netstat.csv", "netbandwidth.csv", "report.js")
UNSUPPORTED = {4, 6, 7, 10}
TOKENS = ["", "x", "1e3", "+5", "007", "1.5", "-3", "12345678901234567890", "0x1f", "1_0", "nan", "inf", "1e400", ".", "-",
          "9" * 25, " 7 ", "2020/13/01 00:00:00.000", "2021/02/29 10:00:00.5", "99 %", "N/A", "[N/A]", "procs", "gpu", "#"]


def mutate(rng, line, sep):
    body = line.rstrip("\n")
    f = body.split(sep) if sep else body.split()
    k = rng.random()
    if k < 0.35 and f:
        f[rng.randrange(len(f))] = rng.choice(TOKENS)
    elif k < 0.5 and len(f) > 1:
        del f[rng.randrange(len(f))]
    elif k < 0.6:
        f.insert(rng.randrange(len(f) + 1), rng.choice(TOKENS))
    elif k < 0.7:
        return "\n"
    elif k < 0.8:
        i = rng.randrange(max(1, len(body)))
        return body[:i] + rng.choice([",", " ", "%", ":", "/", ".", "-", "e", "\t"]) + body[i:] + "\n"
    elif k < 0.85:
        return body                                   # the newline goes: two records run together
    else:
        return line
    return (sep if sep else " ").join(f) + "\n"


def run_pair(tmp, fn, text, **over):
    out = {}
    for who in ("product", "oracle"):
        d = os.path.join(tmp, who) + "/"
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        for m in MANDATORY:
            shutil.copy(os.path.join(BASE, m), d + m)
        with open(d + fn, "w") as f:
            f.write(text)
        try:
            if who == "product":
                cfg = default_report_config(d)
                for k, v in over.items():
                    setattr(cfg, k, v)
                sofa_preprocess(cfg, engine=HostSimEngine())
            else:
                orc.preprocess(orc.OracleConfig(d, nprocs=1, **over), demangle=demangle)
            out[who] = ("ok", {o: open(d + o, "rb").read() for o in OUTPUTS if os.path.isfile(d + o)})
        except (HostSimError, PreprocessError) as e:
            code = int(str(e).rsplit(" ", 1)[1]) if isinstance(e, HostSimError) else -1
            out[who] = ("refused", code)
        except Exception as e:                          # the oracle raises what the reference raises
            out[who] = ("raised", type(e).__name__)
    return out


@pytest.mark.parametrize("fn,sep,over", [
    ("mpstat.txt", ",", {}),
    ("netstat.txt", ",", {}),
    ("nvsmi_query.txt", ", ", {"nvsmi_time_zone": 3}),
    ("vmstat.txt", None, {"enable_vmstat": True}),
    ("nvsmi.txt", None, {"enable_encode_decode": True, "absolute_timestamp": True}),
])
def test_side_file_line_fuzz(fn, sep, over, tmp_path):
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    rng = random.Random(sum(fn.encode()) + int(os.environ.get('SOFA_FUZZ_SEED', '0')))
    lines = open(os.path.join(BASE, fn)).read().splitlines(True)
    agree_ok = 0
    for trial in range(60):
        ls = list(lines)
        i = rng.randrange(1 if fn in ("mpstat.txt", "nvsmi_query.txt") else 0, len(ls))
        ls[i] = mutate(rng, ls[i], sep)
        text = "".join(ls)
        r = run_pair(str(tmp_path), fn, text, **over)
        p, o = r["product"], r["oracle"]
        if o[0] == "ok":
            if p[0] == "refused":
                assert p[1] in UNSUPPORTED, (fn, trial, ls[i], p)
                continue
            assert p[0] == "ok", (fn, trial, ls[i], p)
            assert p[1].keys() == o[1].keys(), (fn, trial, ls[i], sorted(p[1]), sorted(o[1]))
            for k in o[1]:
                assert p[1][k] == o[1][k], (fn, trial, ls[i], k)
            agree_ok += 1
        else:
            assert p[0] == "refused", (fn, trial, ls[i], p, o)
    assert agree_ok >= 5


# ---------------------------------------------------------------- second group: diskstat / blktrace / strace / pystacks (side2.cuh)
OUTPUTS2 = ("diskstat.csv", "diskstat_vector.csv", "diskstat_vector_ui.csv", "blktrace.csv", "strace.csv", "pystacks.csv", "vmstat.csv",
            "report.js")
TOKENS2 = TOKENS + ["D", "C", "[pid", "8372]", "<0.000012>", "<unfinished", "...>", "idle", "a;b;c", "Read", "CPU", "sda", "sd,a", 'q"x',
                    "1600000000.5", "1600000001500000", "= 0", "<>", "<abc>", "<1e-3>", "691248304", "-5"]


def mutate2(rng, line, sep):
    global TOKENS
    keep = TOKENS
    TOKENS = TOKENS2
    try:
        return mutate(rng, line, sep)
    finally:
        TOKENS = keep


def run_pair2(tmp, base, fn, extra, text, **over):
    import sofa_b200.sofa_preprocess as sp
    sp.SYS_BLOCK = "/nonexistent_sofa_b200/"
    out = {}
    for who in ("product", "oracle"):
        d = os.path.join(tmp, who) + "/"
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        for m in MANDATORY + extra:
            shutil.copy(os.path.join(base, m), d + m)
        with open(d + fn, "w") as f:
            f.write(text)
        try:
            if who == "product":
                cfg = default_report_config(d)
                for k, v in over.items():
                    setattr(cfg, k, v)
                sofa_preprocess(cfg, engine=HostSimEngine())
            else:
                orc.preprocess(orc.OracleConfig(d, nprocs=1, **over), demangle=demangle)
            out[who] = ("ok", {o: open(d + o, "rb").read() for o in OUTPUTS2 if os.path.isfile(d + o)})
        except (HostSimError, PreprocessError) as e:
            code = int(str(e).rsplit(" ", 1)[1]) if isinstance(e, HostSimError) else -1
            out[who] = ("refused", code)
        except Exception as e:
            out[who] = ("raised", type(e).__name__)
    return out


@pytest.mark.parametrize("case,fn,sep,extra,over", [
    ("disk_small", "diskstat.txt", ",", (), {}),
    ("disk_small", "diskstat.txt", ",", (), {"absolute_timestamp": True, "diskstat_filters": "sda1"}),
    ("oracle_strace_blk", "blktrace.txt", None, (), {}),
    ("oracle_strace_blk", "strace.txt", None, ("vmstat.txt",), {"plot_ratio": 2}),
    ("oracle_strace_blk", "strace.txt", None, ("blktrace.txt",), {"strace_min_time": 0.0}),
    ("oracle_strace_blk", "pystacks.txt", None, ("vmstat.txt",), {}),
])
def test_storage_file_line_fuzz(case, fn, sep, extra, over, tmp_path):
    import time
    os.environ['TZ'] = 'UTC'
    time.tzset()
    base = os.path.join(GOLDEN, case)
    rng = random.Random(sum(fn.encode()) + len(over) + int(os.environ.get('SOFA_FUZZ_SEED', '0')))
    lines = open(os.path.join(base, fn)).read().splitlines(True)
    agree_ok = 0
    for trial in range(int(os.environ.get('SOFA_FUZZ_TRIALS', '50'))):
        ls = list(lines)
        for _ in range(1 + (trial % 3 == 0)):
            i = rng.randrange(len(ls))
            ls[i] = mutate2(rng, ls[i], sep)
        text = "".join(ls)
        r = run_pair2(str(tmp_path), base, fn, extra, text, **over)
        p, o = r["product"], r["oracle"]
        if o[0] == "ok":
            if p[0] == "refused":
                assert p[1] in UNSUPPORTED, (fn, trial, ls[i], p)
                continue
            assert p[0] == "ok", (fn, trial, ls[i], p)
            assert p[1].keys() == o[1].keys(), (fn, trial, ls[i], sorted(p[1]), sorted(o[1]))
            for k in o[1]:
                assert p[1][k] == o[1][k], (fn, trial, ls[i], k)
            agree_ok += 1
        else:
            assert p[0] == "refused", (fn, trial, ls[i], p, o)
    assert agree_ok >= 5
