"""Row-parser fuzz on the CPU build of the device codecs: mutated perf / nvprof / tcpdump lines
must be accepted or rejected exactly like the reference-path oracle, and accepted lines must
produce the same CSV bytes (pandas to_csv of the oracle rows)."""
import io
import random

import pytest

from hostsim_engine import HostSimEngine, HostSimError
from oracle import sofa_oracle as orc
from sofa_b200 import hostlogic as hl, synth

TIME_BASE = 1600000000.0
UNSUPPORTED = {4, 6, 7, 10}    # spellings / ranges the device codec refuses loudly (documented)


def _mutate(rng, line):
    body = line.rstrip("\n")
    tok = body.split(" ")
    k = rng.random()
    if k < 0.15:
        return body.replace(" ", rng.choice(["  ", "\t", " \t "]), rng.randint(1, 3)) + "\n"
    if k < 0.30 and tok:
        i = rng.randrange(len(tok))
        tok[i] = rng.choice(["", "x", "1e3", "+5", "007", "1.5", "-3", "12345678901234567890", "0x1f", "1_0",
                             "nan", "1e400", ".", "IP", "tcp", "UDP,", "9" * 25, "4.35e-3", "[12]", "a/b", "1/2/3"])
        return " ".join(tok) + "\n"
    if k < 0.40 and len(tok) > 2:
        del tok[rng.randrange(len(tok))]
        return " ".join(tok) + "\n"
    if k < 0.50:
        tok.insert(rng.randrange(len(tok) + 1), rng.choice(["extra", "7", "+0x1", "(lib)"]))
        return " ".join(tok) + "\n"
    if k < 0.55:
        return "\n"
    if k < 0.60:
        return body.upper() + "\n"
    if k < 0.70:
        i = rng.randrange(max(1, len(body)))
        return body[:i] + rng.choice([":", "/", "+", "@", "]", "[", ",", '"', ".", "-", "e"]) + body[i:] + "\n"
    return line


def _check(kind, lines, params, parse_one):
    eng = HostSimEngine()
    valid = []
    for ln in lines:
        try:
            row = parse_one(ln)
            ok = True
        except Exception:
            ok = False
        try:
            eng.load(kind, ln.encode(), params)
            if ok and len(row):
                eng.csv(kind, 1)              # range refusals of the text codecs surface at emit time
            dev_ok = True
            code = 0
        except HostSimError as e:
            dev_ok = False
            code = int(str(e).rsplit(" ", 1)[1])
        if ok and not dev_ok:
            assert code in UNSUPPORTED, (ln, code)
            continue
        assert ok == dev_ok, (ln, "oracle ok=%s device ok=%s code=%d" % (ok, dev_ok, code))
        if ok and ln.endswith("\n"):
            valid.append((ln, row))
        elif ok and len(row):                 # unterminated line = last line of a file: check it alone
            buf = io.StringIO()
            orc.frame([row]).to_csv(buf, header=False, index=False, float_format="%.6f")
            assert eng.csv(kind, 1).tobytes() == buf.getvalue().encode(), ln
    assert len(valid) > len(lines) // 4
    text = "".join(ln for ln, _ in valid)
    rows = [r for _, r in valid]
    if kind == hl.SRC_NET and all(len(r) == 0 for r in rows):
        return
    eng.load(kind, text.encode(), params)
    got = eng.csv(kind, 1).tobytes()
    buf = io.StringIO()
    orc.frame(rows).to_csv(buf, header=False, index=False, float_format="%.6f")
    assert got == buf.getvalue().encode()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_perf_line_fuzz(seed):
    rng = random.Random(seed)
    lines = [_mutate(rng, synth.perf_line(500 + seed, i, cpu_col=bool(i & 1))) for i in range(1500)]
    p = hl.base_params(hl.SRC_PERF, TIME_BASE, False, 1599995000.75)
    _check(hl.SRC_PERF, lines, p,
           lambda ln: orc.parse_perf_sample(ln, False, TIME_BASE, 1599995000.75, [0.0], [3000.0], "cycles"))


@pytest.mark.parametrize("seed", [1, 2])
def test_net_line_fuzz(seed):
    rng = random.Random(seed)
    lines = [_mutate(rng, synth.net_line(600 + seed, i, 3)) for i in range(1500)]
    p = hl.base_params(hl.SRC_NET, TIME_BASE, False, 0.0)
    _check(hl.SRC_NET, lines, p, lambda ln: orc.parse_net_packet(ln, False, TIME_BASE, 0))


@pytest.mark.parametrize("seed", [1, 2])
def test_gpu_line_fuzz(seed):
    rng = random.Random(seed)
    hdr = synth.gpu_header("ms_us_KB", True)
    upd, off = hl.gpu_header(hdr.encode() + synth.gpu_line(700, 0, "ms_us_KB", True).encode())
    p = hl.base_params(hl.SRC_GPU, TIME_BASE, False, -350.0)
    upd.pop("t_base")
    p.update(upd)
    indices = hdr.split("\n")[1].replace('"', '').split(',')

    def mut(line):
        body = line.rstrip("\n")
        f = body.split(",")
        k = rng.random()
        if k < 0.2:
            i = rng.randrange(len(f))
            f[i] = rng.choice(["", "x", "1e3", '"7"', "1.5", "-3", "nan", '"a,b"', "0x1", "1_0", '""', " 2 ", "12345678901234567890123"])
            return ",".join(f) + "\n"
        if k < 0.3 and len(f) > 3:
            del f[rng.randrange(len(f))]
            return ",".join(f) + "\n"
        if k < 0.35:
            return body.replace('"', '', rng.randint(1, 4)) + "\n"
        if k < 0.4:
            return body + ',"tail"\n'
        if k < 0.45:
            return body                     # no trailing newline (last line of a file)
        return line

    lines = [mut(synth.gpu_line(700 + seed, i, "ms_us_KB", True)) for i in range(1500)]
    _check(hl.SRC_GPU, lines, p,
           lambda ln: orc.parse_gpu_record(ln, False, TIME_BASE, indices, 1, 1e3, 1e6, 1024, -350.0))


def test_gpu_copy_tags_anywhere_in_the_name():
    """The reference tags a row by substring search on the quote-free Name (HtoD / DtoH / DtoD / PtoP anywhere,
    first match in that order wins); the device looks only at 't''o' pairs.  Exercise tags at every alignment,
    split by quotes, at the ends of the field, overlapping, and near-misses."""
    rng = random.Random(11)
    hdr = synth.gpu_header("ms_us_KB", True)
    upd, off = hl.gpu_header(hdr.encode() + synth.gpu_line(700, 0, "ms_us_KB", True).encode())
    p = hl.base_params(hl.SRC_GPU, TIME_BASE, False, -350.0)
    upd.pop("t_base")
    p.update(upd)
    indices = hdr.split("\n")[1].replace('"', '').split(',')
    units = hdr.split("\n")[2]
    ts_rescale, dt_rescale, payload_unit = orc.gpu_units(units)
    base = synth.gpu_line(701, 3, "ms_us_KB", True).rstrip("\n").split(",")
    pieces = ["HtoD", "DtoH", "DtoD", "PtoP", "Hto", "toD", "HtoX", "PtoD", "to", "otot", "HHtoDD", "DtoDtoH", "t", "o", "Ht\"oD",
              "H\"toD", "\"HtoD\"", "kernel_", "volta_sgemm", "operator()", "iterator", "x" * 7, "y" * 8, "z" * 9, ""]
    lines = []
    for i in range(600):
        name = "".join(rng.choice(pieces) for _ in range(rng.randint(1, 6)))
        quoted = rng.random() < 0.7
        f = list(base)
        f[-1] = ('"%s"' % name) if quoted else name
        lines.append(",".join(f) + ("\n" if rng.random() < 0.9 or i < 599 else ""))
    lines = [ln for ln in lines if "," not in ln.split(",", 18)[-1].replace('"', '') or True]

    def parse_one(ln):
        return orc.parse_gpu_record(ln, False, TIME_BASE, indices, 1, ts_rescale, dt_rescale, payload_unit, -350.0)
    _check(hl.SRC_GPU, lines, p, parse_one)
