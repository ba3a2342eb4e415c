"""world_size-2 gloo test of the N>1 host logic (no GPU): every rank preprocesses its own
logdirs (CPU build of the device row codecs as the engine - test infrastructure), results are
checked against the oracle per rank, and the throughput aggregation (sum of events / max of
step times) is checked across ranks."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _worker(rank, world, port, workdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hostsim_engine import HostSimEngine
        from helpers import demangle
        from sofa_b200 import synth, multirank
        from sofa_b200.sofa_config import default_report_config
        from oracle import sofa_oracle as orc
        assert multirank.rank_env() == (rank, rank, world)
        ips = ["10.0.0.%d" % i for i in range(1, 5)]
        base = os.path.join(workdir, "sofalog")
        dirs = multirank.cluster_logdirs(base, ",".join(ips))
        assert dirs[2] == base + "-10.0.0.3/"
        mine = multirank.assign_logdirs(dirs, rank, world)
        assert mine == [dirs[rank], dirs[rank + 2]]
        n_events = 0
        for i, d in enumerate(mine):
            seed = 1000 + 10 * rank + i
            synth.make_logdir(d, seed=seed, n_perf=150 + 10 * rank, n_gpu=120, n_net=40)
            synth.make_logdir(d.rstrip('/') + "_ref/", seed=seed, n_perf=150 + 10 * rank, n_gpu=120, n_net=40)
            n_events += 150 + 10 * rank + 120 + 40
        cfg = default_report_config(base)
        multirank.preprocess_cluster(cfg, mine, engine_factory=HostSimEngine)
        assert cfg.logdir == base + "/"
        for d in mine:
            ref = d.rstrip('/') + "_ref/"
            orc.preprocess(orc.OracleConfig(ref, nprocs=1), demangle=demangle)
            for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
                with open(d + fn, "rb") as f, open(ref + fn, "rb") as g:
                    assert f.read() == g.read(), (rank, d, fn)
        dist.barrier()
        ms_local = 100.0 * (rank + 1)
        ev, ms, thr = multirank.aggregate(dist, n_events, ms_local)
        assert ms == 100.0 * world
        assert ev == sum(2 * (150 + 10 * r + 160) for r in range(world))
        assert abs(thr - ev / (ms * 1e-3)) < 1e-9
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world = 2
    port = 29500 + (os.getpid() % 400)
    with tempfile.TemporaryDirectory() as workdir:
        mp.spawn(_worker, args=(world, port, workdir), nprocs=world, join=True)


# ---------------------------------------------------------------------------- one logical trace over two ranks
def _shard_worker(rank, world, port, workdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hostsim_engine import HostSimEngine
        from sofa_b200.sofa_config import default_report_config
        from sofa_b200 import sharded
        import sofa_b200.sofa_preprocess as pre
        from helpers import demangle
        pre.demangle = demangle
        for case in sorted(os.listdir(workdir)):
            cfg = default_report_config(os.path.join(workdir, case, "sharded"), cpu_filters="idle:black,tensorflow:orange")
            if "_ratio" in case:
                cfg.plot_ratio = int(case.split("_ratio")[1])
            io = sharded.ShardIO(rank, world, dist)
            lo, hi = io.byte_range(cfg.logdir + "perf.script") if os.path.isfile(cfg.logdir + "perf.script") else (0, 0)
            eng = HostSimEngine()
            pre.sofa_preprocess(cfg, engine=eng, io=io)
            if os.path.isfile(cfg.logdir + "perf.script") and "_ratio" not in case:
                off, total = io.row_offsets(open(cfg.logdir + "perf.script", "rb").read()[lo:hi].count(b"\n"))
                assert total == open(cfg.logdir + "perf.script", "rb").read().count(b"\n")
                assert (off == 0) == (rank == 0)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_shards_equal_one_shard_gloo():
    """sofa_preprocess of ONE logdir by two ranks (newline-aligned byte ranges, dtype flags all-reduced, ordered concat)
    writes byte-for-byte the files a single rank writes - including when only ONE shard holds the row that decides a
    column's dtype (a Throughput value / a non-IP packet)."""
    import shutil
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    from hostsim_engine import HostSimEngine
    from helpers import demangle
    from sofa_b200 import synth
    from sofa_b200.sofa_config import default_report_config
    import sofa_b200.sofa_preprocess as pre
    pre.demangle = demangle
    world = 2
    port = 29900 + (os.getpid() % 90)
    with tempfile.TemporaryDirectory() as workdir:
        for case, kw in (("mixed", dict(seed=41, n_perf=700, n_gpu=500, n_net=300, non_ip_pct=2)),
                         ("gpu_only_kernels", dict(seed=42, n_gpu=400)),
                         ("net_nan_late", dict(seed=43, n_net=400, non_ip_pct=0)),
                         # plot_ratio > 1: shards must start at rows whose GLOBAL index is a multiple of the ratio
                         ("mixed_ratio3", dict(seed=44, n_perf=701, n_gpu=503, n_net=307, non_ip_pct=2)),
                         ("mixed_ratio7", dict(seed=45, n_perf=1000, n_gpu=333, n_net=200, non_ip_pct=4))):
            a = os.path.join(workdir, case, "single") + "/"
            synth.make_logdir(a, **kw)
            if case == "gpu_only_kernels":
                # keep exactly one memcpy row (with a Throughput) and put it into the LAST quarter: only rank 1 sees it
                lines = open(a + "gputrace.tmp").read().splitlines(keepends=True)
                head, rows = lines[:3], lines[3:]
                mem = [r for r in rows if "[CUDA memcpy" in r]
                ker = [r for r in rows if "[CUDA memcpy" not in r]
                rows = ker[:len(ker) * 3 // 4] + mem[:1] + ker[len(ker) * 3 // 4:]
                open(a + "gputrace.tmp", "w").write("".join(head + rows))
            if case == "net_nan_late":
                lines = open(a + "net.tmp").read().splitlines(keepends=True)
                lines.insert(len(lines) - 5, "1600000001.500000 ARP, Request who-has 10.0.0.1 tell 10.0.0.2, length 28\n")
                open(a + "net.tmp", "w").write("".join(lines))
            shutil.copytree(a, os.path.join(workdir, case, "sharded"))
            cfg = default_report_config(a, cpu_filters="idle:black,tensorflow:orange")
            if "_ratio" in case:
                cfg.plot_ratio = int(case.split("_ratio")[1])
            pre.sofa_preprocess(cfg, engine=HostSimEngine())
        mp.spawn(_shard_worker, args=(world, port, workdir), nprocs=world, join=True)
        for case in ("mixed", "gpu_only_kernels", "net_nan_late", "mixed_ratio3", "mixed_ratio7"):
            a, b = os.path.join(workdir, case, "single") + "/", os.path.join(workdir, case, "sharded") + "/"
            for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
                if os.path.isfile(a + fn):
                    with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
                        assert f.read() == g.read(), (case, fn)
            assert not [f for f in os.listdir(b) if ".shard" in f]


@pytest.mark.parametrize("ratio", [1, 4])
def test_segmented_preprocess_equals_whole(ratio, tmp_path):
    """sofa_preprocess_segmented (bounded device memory: the sources walked in K newline-aligned segments, outputs assembled
    from part files) writes the same bytes as the one-piece run."""
    import shutil
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    from hostsim_engine import HostSimEngine
    from helpers import demangle
    from sofa_b200 import synth, sharded
    from sofa_b200.sofa_config import default_report_config
    import sofa_b200.sofa_preprocess as pre
    pre.demangle = demangle
    a = os.path.join(str(tmp_path), "whole") + "/"
    b = os.path.join(str(tmp_path), "segmented") + "/"
    synth.make_logdir(a, seed=51, n_perf=900, n_gpu=700, n_net=300, non_ip_pct=1)
    synth.write_side_raw(a, 51, n_samples=20, n_cores=2, n_gpus=2)
    shutil.copytree(a, b)
    ca = default_report_config(a, cpu_filters="idle:black,tensorflow:orange")
    cb = default_report_config(b, cpu_filters="idle:black,tensorflow:orange")
    ca.plot_ratio = cb.plot_ratio = ratio            # > 1: segments start at multiples of the ratio; packet dtype flags over the kept rows
    pre.sofa_preprocess(ca, engine=HostSimEngine())
    eng = HostSimEngine()
    sharded.sofa_preprocess_segmented(cb, eng, 20000)
    for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js", "mpstat.csv", "nvsmi_trace.csv", "netstat.csv"):
        with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
            assert f.read() == g.read(), fn
    assert not [f for f in os.listdir(b) if ".shard" in f]


def test_three_shards_with_plot_ratio_gloo():
    """three ranks, plot_ratio 5: the middle shard both hands lines to its predecessor and takes lines from its successor"""
    import shutil
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    from hostsim_engine import HostSimEngine
    from helpers import demangle
    from sofa_b200 import synth
    from sofa_b200.sofa_config import default_report_config
    import sofa_b200.sofa_preprocess as pre
    pre.demangle = demangle
    world = 3
    port = 29700 + (os.getpid() % 90)
    with tempfile.TemporaryDirectory() as workdir:
        case = "three_ratio5"
        a = os.path.join(workdir, case, "single") + "/"
        synth.make_logdir(a, seed=46, n_perf=997, n_gpu=611, n_net=401, non_ip_pct=3)
        shutil.copytree(a, os.path.join(workdir, case, "sharded"))
        cfg = default_report_config(a, cpu_filters="idle:black,tensorflow:orange")
        cfg.plot_ratio = 5
        pre.sofa_preprocess(cfg, engine=HostSimEngine())
        mp.spawn(_shard_worker, args=(world, port, workdir), nprocs=world, join=True)
        b = os.path.join(workdir, case, "sharded") + "/"
        for fn in ("cputrace.csv", "gputrace.csv", "nettrace.csv", "report.js"):
            with open(a + fn, "rb") as f, open(b + fn, "rb") as g:
                assert f.read() == g.read(), fn
