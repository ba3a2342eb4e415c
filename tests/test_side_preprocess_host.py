"""N1: mpstat.txt / nvsmi_query.txt / netstat.txt through the product host logic
(sofa_b200.sofa_preprocess) and the CPU build of the device row functions (side.cuh), byte for byte
against the unmodified reference's files (tests/golden/side_*); edge cases against the oracle."""
import ctypes
import math
import os
import random
import struct

import pytest

from helpers import side_cases, build_side_logdir, golden_bytes, SIDE_OUTPUTS, demangle, preprocess_overrides
from hostsim_engine import HostSimEngine, HostSimError, build_hostsim
from oracle import sofa_oracle as orc
from sofa_b200.sofa_config import default_report_config
from sofa_b200.sofa_preprocess import sofa_preprocess


def run_product(logdir, engine, **over):
    cfg = default_report_config(logdir)
    for k, v in over.items():
        setattr(cfg, k, v)
    sofa_preprocess(cfg, engine=engine)
    return cfg


def assert_same(got, want, what):
    if got != want:
        i = next((k for k in range(min(len(got), len(want))) if got[k] != want[k]), min(len(got), len(want)))
        raise AssertionError("%s differs at byte %d: got %r want %r" % (what, i, got[max(0, i - 80):i + 60], want[max(0, i - 80):i + 60]))


@pytest.mark.parametrize("name", side_cases())
def test_side_sources_reproduce_reference_files(name, tmp_path):
    case, logdir = build_side_logdir(name, tmp_path)
    run_product(logdir, HostSimEngine(), **preprocess_overrides(case))
    for fn in SIDE_OUTPUTS:
        want = golden_bytes(name, fn)
        if want is None:
            assert not os.path.isfile(logdir + fn), fn
        else:
            assert_same(open(logdir + fn, "rb").read(), want, "%s/%s" % (name, fn))


def oracle_vs_product(tmp_path, mutate, base="side_small", **over):
    """Same mutated logdir through the oracle and the product; compare every side output."""
    import shutil
    _, a = build_side_logdir(base, tmp_path / "a")
    mutate(a)
    b = str(tmp_path / "b") + "/"
    shutil.copytree(a, b)
    run_product(a, HostSimEngine(), **over)
    orc.preprocess(orc.OracleConfig(b, nprocs=1, **over), demangle=demangle)
    for fn in SIDE_OUTPUTS:
        ea, eb = os.path.isfile(a + fn), os.path.isfile(b + fn)
        assert ea == eb, fn
        if ea:
            assert_same(open(a + fn, "rb").read(), open(b + fn, "rb").read(), fn)


def test_mpstat_bad_lines_and_absolute_timestamps(tmp_path):
    def mutate(d):
        lines = open(d + "mpstat.txt").read().split("\n")
        lines.insert(40, "1600000004.1,2,77")                # wrong field count: genfromtxt drops it
        lines.insert(90, "")                                 # blank line
        lines[120] = lines[120] + ",9"                       # one field too many
        lines[-2] = lines[-2].rsplit(",", 3)[0]              # truncated last record
        open(d + "mpstat.txt", "w").write("\n".join(lines))
    oracle_vs_product(tmp_path, mutate, absolute_timestamp=True)


def test_mpstat_idle_interval_and_no_increase(tmp_path):
    def mutate(d):
        lines = open(d + "mpstat.txt").read().split("\n")
        # repeat one sample verbatim: every delta is 0 -> "No increases in mpstat values" rows are skipped
        block = lines[1 + 5 * 20: 1 + 5 * 21]
        lines[1 + 5 * 21: 1 + 5 * 22] = [b.replace(b.split(",")[0], l.split(",")[0], 1) for b, l in zip(block, lines[1 + 5 * 21: 1 + 5 * 22])]
        open(d + "mpstat.txt", "w").write("\n".join(lines))
    oracle_vs_product(tmp_path, mutate)


def test_nvsmi_time_zone_failed_and_short_fraction(tmp_path):
    def mutate(d):
        lines = open(d + "nvsmi_query.txt").read().split("\n")
        lines[7] = lines[7][:20] + "5" + lines[7][23:]                    # 1-digit fraction ".5"
        open(d + "nvsmi_query.txt", "w").write("\n".join(lines))
    oracle_vs_product(tmp_path, mutate, nvsmi_time_zone=8)

    def failed(d):
        with open(d + "nvsmi_query.txt", "a") as f:
            f.write("Unable to determine the device handle: Failed\n")
    oracle_vs_product(tmp_path / "f", failed)


def test_netstat_single_line_and_no_trailing_newline(tmp_path):
    def one(d):
        open(d + "netstat.txt", "w").write("1600000000.5,100,200\n")
    oracle_vs_product(tmp_path / "one", one)

    def nonl(d):
        t = open(d + "netstat.txt").read().rstrip("\n")
        open(d + "netstat.txt", "w").write(t)
    oracle_vs_product(tmp_path / "nonl", nonl)

    def neg(d):                                              # counter wrap: negative bandwidth, '%d' truncates toward zero
        lines = open(d + "netstat.txt").read().split("\n")
        lines[10] = lines[10].split(",")[0] + ",5,7"
        open(d + "netstat.txt", "w").write("\n".join(lines))
    oracle_vs_product(tmp_path / "neg", neg)


def test_vmstat_and_dmon_edge_cases(tmp_path):
    def short_run(d):                                        # fewer than 4 seconds: no enc/dec rows, nvsmi_data stays False
        lines = open(d + "nvsmi.txt").read().split("\n")
        open(d + "nvsmi.txt", "w").write("\n".join(lines[:8]) + "\n")
    oracle_vs_product(tmp_path / "short", short_run, base="side_full", enable_vmstat=True, enable_encode_decode=True)

    def failed(d):
        with open(d + "nvsmi.txt", "a") as f:
            f.write("\nFailed\n")
    oracle_vs_product(tmp_path / "failed", failed, base="side_full", enable_encode_decode=True)

    def headers_only(d):
        lines = [l for l in open(d + "vmstat.txt").read().split("\n") if "procs" in l or "swpd" in l]
        open(d + "vmstat.txt", "w").write("\n".join(lines) + "\n")
    oracle_vs_product(tmp_path / "hdr", headers_only, base="side_full", enable_vmstat=True)

    def decimals(d):                                         # float() accepts "1.5" and "2e3"; int() truncates after + 1e-5
        lines = open(d + "vmstat.txt").read().split("\n")
        f = lines[5].split()
        f[8], f[9], f[12] = "1.99999", "2e3", "7.5"
        lines[5] = " ".join(f)
        open(d + "vmstat.txt", "w").write("\n".join(lines))
    oracle_vs_product(tmp_path / "dec", decimals, base="side_full", enable_vmstat=True, absolute_timestamp=True)


def test_side_errors_are_loud(tmp_path):
    _, d = build_side_logdir("side_small", tmp_path)
    lines = open(d + "netstat.txt").read().split("\n")
    lines[3] = lines[2]                                       # same timestamp twice: ZeroDivisionError in the reference
    open(d + "netstat.txt", "w").write("\n".join(lines))
    with pytest.raises(HostSimError):
        run_product(d, HostSimEngine())


def test_repr_matches_python():
    lib = build_hostsim()
    lib.hs_repr.argtypes = [ctypes.c_double, ctypes.c_char_p]
    buf = ctypes.create_string_buffer(64)
    rng = random.Random(20260921)
    vals = [0.0, -0.0, 1.0, -1.5, 0.1, 0.2 + 0.1, 1e-4, 9.999e-5, 1.2345e-5, 1e-6, 123456.789, 1e15, 9007199254740993.0,
            9999999999999998.0, 0.7000000476837158, 68.14159292035399, 8157367.779472141, 5e-324 * 0 + 2.5e-5,
            4503599627370496.0, 4503599627370497.0, 0.3, 2.0 / 3.0, 1.0 / 3.0, 100.0, 1e5, 99999.99999999999]
    for _ in range(20000):
        kind = rng.random()
        if kind < 0.4:
            vals.append(struct.unpack('<d', struct.pack('<Q', rng.getrandbits(64)))[0])
        elif kind < 0.7:
            vals.append(rng.uniform(-1e7, 1e7))
        elif kind < 0.85:
            vals.append(rng.randrange(0, 10**7) / rng.choice([1, 3, 7, 100, 1000, 12345]))
        else:
            vals.append(10 ** rng.uniform(-6, 16) * rng.choice([-1, 1]))
    n_ok = 0
    for v in vals:
        if math.isnan(v) or math.isinf(v):
            continue
        r = lib.hs_repr(v, buf)
        a = abs(v)
        if v == 0 or (1e-6 <= a < 1e16):
            assert r > 0, v
            assert buf.raw[:r].decode() == repr(v), (v, buf.raw[:r], repr(v))
            n_ok += 1
        elif r > 0:                                          # outside the guaranteed range: never wrong, may refuse
            assert buf.raw[:r].decode() == repr(v), (v, buf.raw[:r], repr(v))
    assert n_ok > 10000


@pytest.mark.parametrize("ratio,absolute", [(1, False), (3, True)])
def test_storage_sources_match_oracle_larger(ratio, absolute, tmp_path):
    """diskstat / blktrace / strace / pystacks at sizes that span many blocks, with the odd lines the reference tolerates
    (CPU build of side2.cuh through the product host logic vs the oracle; tests/test_gpu_parity.py repeats it on the device)"""
    import shutil
    from helpers import make_storage_logdir, STORAGE_OUTPUTS
    a = str(tmp_path / "a") + "/"
    make_storage_logdir(a, n_disk=400, n_strace=3000, n_blk=4000, n_py=700)
    b = str(tmp_path / "b") + "/"
    shutil.copytree(a, b)
    run_product(a, HostSimEngine(), plot_ratio=ratio, absolute_timestamp=absolute, diskstat_filters="xda1")
    orc.preprocess(orc.OracleConfig(b, nprocs=1, plot_ratio=ratio, absolute_timestamp=absolute, diskstat_filters="xda1"), demangle=demangle)
    for fn in STORAGE_OUTPUTS:
        assert_same(open(a + fn, "rb").read(), open(b + fn, "rb").read(), fn)


def test_storage_sources_edge_cases(tmp_path):
    """empty / degenerate diskstat, blktrace, strace and pystacks files: the product writes (or leaves out) the same
    files as the reference path"""
    from helpers import STORAGE_OUTPUTS

    def disk_one_sample(d):                      # fewer lines than devices: header-only CSVs, no series
        open(d + "diskstat.txt", "w").write("1600000000.5,sda,1,2,3,4,5,6\n1600000000.5,sdb,1,2,3,4,5,6\n")
        open(d + "vmstat.txt", "w").write(open(os.path.join(GOLDEN, "oracle_strace_blk", "vmstat.txt")).read())

    def disk_empty(d):
        open(d + "diskstat.txt", "w").close()

    def disk_no_final_newline(d):                # the reference drops the last character of the last line
        t = open(os.path.join(GOLDEN, "disk_small", "diskstat.txt")).read()
        open(d + "diskstat.txt", "w").write(t.rstrip("\n") + "9")

    def blk_footer_only(d):                      # no accepted line: both lists empty, no file, `cpuid` stays unset
        open(d + "blktrace.txt", "w").write("CPU0 (8,0):\n Reads Queued:           0,        0KiB\t Writes Queued:         100,    51200KiB\n")

    def blk_empty_and_short_strace(d):
        open(d + "blktrace.txt", "w").close()
        open(d + "strace.txt", "w").write("4242 1600000000.402969 read(3, \"abc\", 4096) = 4 <0.121879>\n")      # one line: ignored

    def strace_all_filtered(d):
        open(d + "vmstat.txt", "w").write(open(os.path.join(GOLDEN, "oracle_strace_blk", "vmstat.txt")).read())
        open(d + "strace.txt", "w").write("1 1600000000.4 close(3) = 0 <0.1>\n2 1600000000.5 nanosleep({1}, NULL) = 0 <0.2>\n")

    def py_two_stacks(d):                        # one row left after the last is deleted, and that one is lost: warning only
        open(d + "vmstat.txt", "w").write(open(os.path.join(GOLDEN, "oracle_strace_blk", "vmstat.txt")).read())
        open(d + "pystacks.txt", "w").write("1600000000500000\na;b\n1600000000600000\nc\n")

    from helpers import GOLDEN
    outputs = tuple(o for o in SIDE_OUTPUTS)
    for mut in (disk_one_sample, disk_empty, disk_no_final_newline, blk_footer_only, blk_empty_and_short_strace, strace_all_filtered,
                py_two_stacks):
        oracle_vs_product(tmp_path / mut.__name__, mut)


def test_product_vs_oracle_over_report_options(tmp_path):
    """random combinations of the report options on a logdir holding every source: product host logic + CPU build of the row
    codecs against the oracle (540 combinations ran clean; the oracle is held to the reference on the same generator by
    tests/test_fuzz_ref_cpu.py)"""
    import random
    import shutil
    import time
    from helpers import make_full_logdir, random_overrides, ALL_PREPROCESS_OUTPUTS
    os.environ['TZ'] = 'UTC'
    time.tzset()
    base = str(tmp_path / "base") + "/"
    make_full_logdir(base)
    rng = random.Random(int(os.environ.get('SOFA_FUZZ_SEED', '0')) + 11)
    for t in range(int(os.environ.get('SOFA_FUZZ_TRIALS', '8'))):
        over = random_overrides(rng)
        a, b = str(tmp_path / ("p%d" % t)) + "/", str(tmp_path / ("o%d" % t)) + "/"
        shutil.copytree(base, a)
        shutil.copytree(base, b)
        cfg = default_report_config(a, cpu_filters=over.get('cpu_filters'))
        for k, v in over.items():
            if k != 'cpu_filters':
                setattr(cfg, k, v)
        sofa_preprocess(cfg, engine=HostSimEngine())
        orc.preprocess(orc.OracleConfig(b, nprocs=1, **over), demangle=demangle)
        for fn in ALL_PREPROCESS_OUTPUTS:
            ea, eb = os.path.isfile(a + fn), os.path.isfile(b + fn)
            assert ea == eb, (over, fn)
            if ea:
                assert_same(open(a + fn, "rb").read(), open(b + fn, "rb").read(), "%s %s" % (over, fn))
