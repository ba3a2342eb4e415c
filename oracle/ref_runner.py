"""Runs the UNMODIFIED reference (`sofa_preprocess(cfg)` + `sofa_analyze(cfg)` of cyliustack/sofa, copied to
oracle/_ref/bin by oracle/build_ref.py) on a logdir.  TEST INFRASTRUCTURE: used by bench.py's cpu_baseline /
`--impl reference` legs and by tests; the product path never imports it.

The reference needs what tests/golden/make_golden*.py give it in the build container, all at RUN time (its sources are
untouched): import stubs for cxxfilt (libstdc++ __cxa_demangle, what the package wraps), sqlalchemy, fuzzywuzzy and
matplotlib; PATH shims `nvprof` / `tcpdump` that cat the pre-dumped text; and the pandas-3 shims of
make_golden_driver.install_driver_shims / patch_mpstat_profile (APIs pandas removed since the reference was written)."""
import contextlib
import io
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_BIN = os.path.join(HERE, "_ref", "bin")
_state = {}


def available():
    return os.path.isfile(os.path.join(REF_BIN, "sofa_preprocess.py")) and os.path.isfile(os.path.join(REF_BIN, "sofa_analyze.py"))


def _setup():
    if _state:
        return _state
    gold = os.path.join(REPO, "tests", "golden")
    for p in (REPO, gold):
        if p not in sys.path:
            sys.path.insert(0, p)
    import make_golden as mg
    import make_golden_driver as md
    mg.REF_BIN = REF_BIN
    mg.install_stubs()
    sys.path.insert(0, REF_BIN)
    shimdir = tempfile.mkdtemp(prefix="sofa_ref_shims_")
    mg.install_shims(shimdir)
    md.install_driver_shims()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        import sofa_preprocess as ref_pre
        import sofa_analyze as ref_an
    md.patch_mpstat_profile(ref_an)
    _state.update(mg=mg, pre=ref_pre, an=ref_an, shimdir=shimdir)
    return _state


def reference_pass(logdir, overrides=None, analyze=True):
    """sofa_preprocess(cfg) + sofa_analyze(cfg) of the unmodified reference in `logdir` (a complete synthetic logdir with
    pre-dumped gputrace.tmp / net.tmp next to their empty .nvvp / .pcap markers).  Returns the cfg."""
    st = _setup()
    logdir = logdir if logdir.endswith("/") else logdir + "/"
    for src, dst in (("gputrace.tmp", "gputrace.src"), ("net.tmp", "net.src"), ("cuda_api_trace.tmp", "cuda_api_trace.src")):
        if os.path.isfile(logdir + src):
            shutil.copy(logdir + src, logdir + dst)
    cfg = st["mg"].make_ref_cfg(logdir, dict(overrides or {}))
    cfg.roi_begin = cfg.roi_end = 0
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        st["pre"].sofa_preprocess(cfg)
        if analyze:
            st["an"].sofa_analyze(cfg)
    for fn in ("gputrace.src", "net.src", "cuda_api_trace.src"):
        if os.path.isfile(logdir + fn):
            os.remove(logdir + fn)
    return cfg


def run_batch_in_subprocess(items, timeout=3600):
    """items: [(logdir, overrides dict)] -> [{"ok": True} | {"raised": "<exception type>"}], preprocess only, all in ONE
    child interpreter (fuzz tests: what does the reference do with this input?)"""
    import json
    import subprocess
    spec = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
    json.dump([[d, ov] for d, ov in items], spec)
    spec.close()
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--batch", spec.name], capture_output=True, text=True, timeout=timeout)
    finally:
        os.unlink(spec.name)
    if out.returncode != 0:
        raise RuntimeError("the reference batch failed:\n%s" % out.stderr[-3000:])
    return json.loads([ln for ln in out.stdout.split("\n") if ln.startswith("[")][-1])


def run_in_subprocess(logdir, cpu_filters=None, repeat=1, timeout=3600, overrides=None, analyze=True):
    """reference_pass in a child interpreter (its import stubs and pandas shims stay out of the caller's process);
    -> list of wall-clock seconds of `sofa_preprocess(cfg) + sofa_analyze(cfg)`, one per repeat (imports excluded)"""
    import json
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), logdir, "--repeat", str(repeat)]
    if cpu_filters:
        cmd += ["--cpu_filters", cpu_filters]
    if overrides:
        cmd += ["--overrides", json.dumps(overrides)]
    if not analyze:
        cmd += ["--no-analyze"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError("the reference failed on %s:\n%s" % (logdir, out.stderr[-3000:]))
    line = [ln for ln in out.stdout.split("\n") if ln.startswith("{")][-1]
    return json.loads(line)["seconds"]


if __name__ == "__main__":
    import argparse
    import json
    import time
    ap = argparse.ArgumentParser()
    ap.add_argument("logdir", nargs="?")
    ap.add_argument("--batch", help="JSON file: [[logdir, overrides], ...]; preprocess only, one result per entry")
    ap.add_argument("--cpu_filters")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--overrides", help="JSON dict of cfg attributes (plot_ratio, absolute_timestamp, ...)")
    ap.add_argument("--no-analyze", action="store_true")
    a = ap.parse_args()
    if not available():
        sys.exit("oracle/_ref/bin is not built (python oracle/build_ref.py in the build container)")
    _setup()
    if a.batch:
        res = []
        for d, ov in json.load(open(a.batch)):
            try:
                reference_pass(d, ov, analyze=False)
                res.append({"ok": True})
            except SystemExit as e:
                res.append({"raised": "SystemExit"})
            except BaseException as e:                     # what the reference raises is the answer
                res.append({"raised": type(e).__name__})
        print(json.dumps(res))
        sys.exit(0)
    secs = []
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        ov = json.loads(a.overrides) if a.overrides else {}
        if a.cpu_filters:
            ov["cpu_filters"] = a.cpu_filters
        reference_pass(a.logdir, ov, analyze=not a.no_analyze)
        secs.append(time.perf_counter() - t0)
    print(json.dumps({"seconds": secs}))
