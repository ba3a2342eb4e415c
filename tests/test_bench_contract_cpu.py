"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract
(one line, the agreed keys, rank != 0 silent under torchrun)."""
import json
import os
import subprocess
import sys

from helpers import REPO

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample", "3000"], capture_output=True, text=True, env=env, timeout=600, cwd=REPO)


def test_reference_arm_prints_one_contract_line():
    out = run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if os.path.isfile(os.path.join(REPO, "oracle", "_ref", "bin", "sofa_preprocess.py")):
        assert d["cpu_baseline"]["kind"] == "reference" and "reference_error" not in d["config"]     # the unmodified reference ran
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    out = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29599"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert not [ln for ln in out.stdout.split("\n") if ln.startswith("{")]


import pytest


@pytest.mark.parametrize("n_events", [20000, 100000])
def test_port_equals_the_unmodified_reference_on_a_bench_sample(n_events, tmp_path):
    """oracle/_ref (the reference's own modules, oracle/build_ref.py) and the oracle port on the same slice of the bench
    workload (preprocess + the whole analyze driver): all twelve output files byte-identical.  Skipped where the
    reference was never built into oracle/_ref."""
    import shutil
    sys.path.insert(0, REPO)
    import bench
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("oracle/_ref is not built (needs /root/reference at build time)")
    a, b = str(tmp_path / "ref") + "/", str(tmp_path / "port") + "/"
    os.makedirs(a)
    bench.write_sample_logdir(a, n_events, bench.SEED)
    shutil.copytree(a, b)
    secs = ref_runner.run_in_subprocess(a, bench.CPU_FILTERS)
    assert len(secs) == 1 and secs[0] > 0
    bench.oracle_pass(b, 4)
    assert set(bench.parity_sample(a, b)) == set(bench.PARITY_FILES)


def test_oracle_ref_holds_the_reference_unmodified():
    """oracle/_ref/bin is a byte-for-byte copy of the modules under /root/reference/bin (build container only) and is
    kept out of the history"""
    import glob
    import pytest
    if not os.path.isdir("/root/reference/bin"):
        pytest.skip("no /root/reference here")
    sys.path.insert(0, REPO)
    from oracle import build_ref
    assert build_ref.build_ref() >= 10
    for p in glob.glob("/root/reference/bin/*.py"):
        q = os.path.join(REPO, "oracle", "_ref", "bin", os.path.basename(p))
        assert open(p, "rb").read() == open(q, "rb").read(), q
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], capture_output=True, text=True, cwd=REPO).stdout.strip()
    assert tracked == ""
    assert "oracle/_ref/" in open(os.path.join(REPO, ".gitignore")).read()
