"""One logical trace across N GPUs (BASELINE configs[3]).

Every source file is cut into N newline-aligned byte ranges, one per rank (one process per GPU, torchrun).  Rows are
independent in sofa_preprocess and each source stays in log order, so rank r parses and formats only its range and
the concatenation of the ranks' outputs in rank order IS the file.  What is NOT local, and is exchanged with
torch.distributed collectives (NCCL on the GPUs, gloo in the CPU tests):

  * the dtype of a CSV / report.js column is inferred by pandas over the WHOLE file ("bandwidth" is a float column
    iff some row has a Throughput; one non-IP packet turns every numeric column of nettrace.csv into floats):
    all-reduce of the per-shard row flags before anything is formatted (ShardIO.sync_flags);
  * record counts of every report.js series (a series with no records anywhere is left out of `sofa_traces`):
    all-gather in ShardIO.finish;
  * row counts per rank (ShardIO.row_offsets: the global index of a shard's first row, e.g. the index column of comm.csv).

The unique mangled viz names are demangled per rank (a name that occurs in two shards is demangled twice: no exchange
needed).  The small recorder side files (mpstat, nvidia-smi, netstat, vmstat) are handled by rank 0 alone.
sofa_analyze runs on rank 0 over the assembled logdir.
"""
import os


class ShardIO:
    """Same interface as sofa_preprocess.PlainIO."""

    def __init__(self, rank, world, dist, device=None):
        self.rank, self.world, self.dist, self.device = rank, world, dist, device
        self.csv_parts = []            # (path, header) in emit order
        self.rows = {}
        self.tm = None

    # ---- input: newline-aligned byte ranges
    def size(self, path):
        return os.path.getsize(path)

    def head(self, path, n):
        with open(path, 'rb') as f:
            return f.read(n)

    @staticmethod
    def _line_start_at_or_after(f, pos, lo, size):
        """first offset p >= pos (p >= lo) at which a line starts"""
        if pos <= lo:
            return lo
        f.seek(pos - 1)
        p = pos - 1
        while p < size:
            chunk = f.read(1 << 16)
            if not chunk:
                break
            i = chunk.find(b'\n')
            if i >= 0:
                return p + i + 1
            p += len(chunk)
        return size

    def byte_range(self, path, data_off=0):
        size = os.path.getsize(path)
        span = max(0, size - data_off)
        with open(path, 'rb') as f:
            lo = self._line_start_at_or_after(f, data_off + span * self.rank // self.world, data_off, size)
            hi = self._line_start_at_or_after(f, data_off + span * (self.rank + 1) // self.world, data_off, size) \
                if self.rank + 1 < self.world else size
        return lo, max(lo, hi)

    # ---- plot_ratio > 1: the reference keeps rows i % plot_ratio == 0 of the WHOLE file (list_downsample, :51-57); a shard
    # must therefore start at a row whose global index is a multiple of the ratio.  Line counts of the byte ranges are
    # all-gathered, and every rank hands the first (-global_start) % ratio lines of its range to its predecessor.
    ratio = 1

    @staticmethod
    def _count_lines(path, lo, hi, size):
        n, last = 0, b'\n'
        with open(path, 'rb') as f:
            f.seek(lo)
            left = hi - lo
            while left > 0:
                chunk = f.read(min(left, 1 << 24))
                if not chunk:
                    break
                n += chunk.count(b'\n')
                last = chunk[-1:]
                left -= len(chunk)
        return n + (1 if hi > lo and hi == size and last != b'\n' else 0)      # an unterminated last line is a row too

    @staticmethod
    def _skip_lines(path, lo, hi, k):
        """offset of the line k lines after the one starting at lo (never beyond hi)"""
        if k <= 0:
            return lo
        with open(path, 'rb') as f:
            f.seek(lo)
            pos = lo
            while pos < hi:
                chunk = f.read(min(hi - pos, 1 << 16))
                if not chunk:
                    break
                i = -1
                while k > 0:
                    i = chunk.find(b'\n', i + 1)
                    if i < 0:
                        break
                    k -= 1
                if k == 0:
                    return pos + i + 1
                pos += len(chunk)
        return hi

    def _all_ranges(self, path, data_off, lo, hi):
        """[[lines, lo, hi] of every rank's byte range]"""
        return self.all_gather_counts([self._count_lines(path, lo, hi, os.path.getsize(path)), lo, hi])

    def aligned_range(self, path, data_off=0):
        lo, hi = self.byte_range(path, data_off)
        if self.ratio <= 1 or self.world == 1:
            return lo, hi
        per = self._all_ranges(path, data_off, lo, hi)
        start, skip = 0, []
        for r in range(self.world):
            skip.append((-start) % self.ratio if r else 0)
            start += per[r][0]
        if any(skip[r] > per[r][0] for r in range(self.world)):
            raise NotImplementedError("%s: a shard holds fewer rows than plot_ratio; use fewer ranks for so small a trace" % path)
        new_lo = self._skip_lines(path, lo, hi, skip[self.rank])
        new_hi = self._skip_lines(path, per[self.rank + 1][1], per[self.rank + 1][2], skip[self.rank + 1]) \
            if self.rank + 1 < self.world else hi
        return new_lo, max(new_lo, new_hi)

    def read(self, path, data_off=0):
        lo, hi = self.aligned_range(path, data_off)
        with open(path, 'rb') as f:
            if self.rank == 0:                       # rank 0 keeps the header rows in front of its shard
                return f.read(hi), data_off
            f.seek(lo)
            return f.read(hi - lo), 0

    # ---- collectives
    def _tensor(self, values, dtype=None):
        import torch
        return torch.tensor(values, dtype=dtype or torch.int64, device=self.device)

    def sync_flags(self, eng, kind, stride):
        """bitwise OR over the ranks of the shard's row flags (NCCL has no BOR: MAX of the three bits)"""
        import torch
        fl = eng.row_flags(kind, stride)
        t = self._tensor([fl & 1, (fl >> 1) & 1, (fl >> 2) & 1], torch.int32)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        b = [int(x) for x in t.tolist()]
        eng.set_global_flags(kind, stride, b[0] | (b[1] << 1) | (b[2] << 2))

    def all_gather_counts(self, counts):
        import torch
        t = self._tensor(list(counts))
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    def row_offsets(self, n_rows):
        """global index of this shard's first row, total rows"""
        per = [c[0] for c in self.all_gather_counts([n_rows])]
        return sum(per[:self.rank]), sum(per)

    # ---- output: part files, assembled by rank 0 in rank order
    def write_csv(self, path, header, body):
        with open('%s.shard%d' % (path, self.rank), 'wb') as f:
            f.write(memoryview(body))
        self.csv_parts.append((path, header))

    def _write_series_parts(self, series, path):
        for i, s in enumerate(series):
            if s.nrec > 0:
                with open('%s.s%d.shard%d' % (path, i, self.rank), 'wb') as f:
                    f.write(memoryview(s.records))

    def _assemble(self, series, counts, path):
        for p, header in self.csv_parts:
            with open(p, 'wb') as out:
                out.write(header)
                for r in range(self.world):
                    part = '%s.shard%d' % (p, r)
                    with open(part, 'rb') as f:
                        while True:
                            chunk = f.read(1 << 24)
                            if not chunk:
                                break
                            out.write(chunk)
                    os.remove(part)
        if len(series):
            self._write_report(series, counts, path)

    def finish(self, series, path):
        counts = self.all_gather_counts([s.nrec for s in series])
        self._write_series_parts(series, path)
        self.dist.barrier()
        if self.rank == 0:
            self._assemble(series, counts, path)
        self.dist.barrier()

    def _write_report(self, series, counts, path):
        from .sofa_preprocess import _json_str
        total = [sum(counts[r][i] for r in range(self.world)) for i in range(len(series))]
        with open(path, 'wb') as f:
            for i, s in enumerate(series):
                if total[i] > 0:
                    f.write(("%s = {\"name\": %s, \"color\": %s, \"data\": [" %
                             (s.name, _json_str(s.title), _json_str(s.color))).encode())
                    first = True
                    for r in range(self.world):
                        if counts[r][i] == 0:
                            continue
                        part = '%s.s%d.shard%d' % (path, i, r)
                        if not first:
                            f.write(b", ")
                        first = False
                        with open(part, 'rb') as g:
                            while True:
                                chunk = g.read(1 << 24)
                                if not chunk:
                                    break
                                f.write(chunk)
                        os.remove(part)
                    f.write(b"]}")
                f.write(b"\n\n")
            f.write(b"sofa_traces = [ ")
            for i, s in enumerate(series):
                if total[i] > 0:
                    f.write((s.name + ",").encode())
            f.write(b" ]")


def sofa_preprocess_sharded(cfg, dist, engine=None, device=None):
    """sofa_preprocess(cfg) of ONE logdir by all ranks of `dist` together; every rank must call it."""
    from .sofa_preprocess import sofa_preprocess
    rank, world = dist.get_rank(), dist.get_world_size()
    if getattr(cfg, 'cuda_api_tracing', False):
        raise NotImplementedError("sharded preprocess does not cover --cuda_api_tracing")
    io = ShardIO(rank, world, dist, device)
    return sofa_preprocess(cfg, engine=engine, io=io)


# ----------------------------------------------------------------------------- gathering parsed shards on one rank
class _DevMem:
    """n bytes of device memory as an object torch.as_tensor can alias (no copy)"""
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def _view(ptr, nbytes, device):
    import torch
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def gather_source(eng, kind, dist, device, dst=0):
    """Move every rank's parsed shard of source `kind` (its text range and its SoA columns, device memory) to rank
    `dst` over NCCL (point-to-point sends straight into their final position in rank dst's buffers; NVLink between the
    GPUs of a node) and make them that rank's parsed state, as if it had parsed the whole file: sofa_analyze's group
    sums follow numpy's pairwise order over ALL rows of a column, so the analyze half runs where the whole column is.
    Every rank must call it.  Returns the bytes this rank sent or received."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    d = eng.shard_export(kind)
    own_text = d.text_bytes            # rank 0's text keeps its header rows in front: the rows' line offsets count from the buffer start
    t = torch.tensor([d.n_rows, own_text], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(sizes, t)
    rows = [int(x[0]) for x in sizes]
    texts = [int(x[1]) for x in sizes]
    ncols = d.n_cols
    moved = 0
    torch.cuda.synchronize(device)
    if rank == dst:
        g = eng.shard_reserve(kind, sum(rows), sum(texts))
        reqs = []
        row0, text0 = 0, 0
        for r in range(world):
            parts = [(g.text + text0, texts[r], d.text)]
            for q in range(ncols):
                parts.append((g.cols[q] + row0 * g.col_elem[q], rows[r] * g.col_elem[q], d.cols[q]))
            for dptr, nb, sptr in parts:
                if nb == 0:
                    continue
                if r == rank:
                    _view(dptr, nb, device).copy_(_view(sptr, nb, device))
                else:
                    reqs.append(dist.P2POp(dist.irecv, _view(dptr, nb, device), r))
                    moved += nb
            row0 += rows[r]
            text0 += texts[r]
        if reqs:
            for q in dist.batch_isend_irecv(reqs):           # one NCCL group: the transfers of all shards and columns overlap
                q.wait()
        torch.cuda.synchronize(device)
        eng.shard_commit(kind, rows, texts)
    else:
        parts = [(d.text, own_text)] + [(d.cols[q], d.n_rows * d.col_elem[q]) for q in range(ncols)]
        reqs = []
        for sptr, nb in parts:
            if nb:
                reqs.append(dist.P2POp(dist.isend, _view(sptr, nb, device), dst))
                moved += nb
        if reqs:
            for q in dist.batch_isend_irecv(reqs):
                q.wait()
        torch.cuda.synchronize(device)
    return moved


# ----------------------------------------------------------------------------- one GPU, bounded device memory
class SegmentIO(ShardIO):
    """The same byte-range splitting walked SEQUENTIALLY by one process: segment i of K plays rank i of K.  Device
    memory then holds one segment of a source at a time instead of the whole file (text + columns + outputs are ~5x
    the text: a 100 GB perf.script does not fit 180 GB of HBM in one piece), and the logdir's outputs are assembled
    from the segments' part files exactly as a sharded run assembles its ranks'.  The file-level dtype flags are
    collected by a parse-only pass over the gpu / packet segments first (state['flags'])."""

    def __init__(self, index, count, state):
        ShardIO.__init__(self, index, count, None)
        self.state = state
        self.ratio = state.get('ratio', 1)

    def _all_ranges(self, path, data_off, lo, hi):           # no peers to ask: the segments of a file are walked once and cached
        key = (path, data_off)
        cache = self.state.setdefault('ranges', {})
        if key not in cache:
            size = os.path.getsize(path)
            out = []
            for r in range(self.world):
                a, b = ShardIO(r, self.world, None).byte_range(path, data_off)
                out.append([self._count_lines(path, a, b, size), a, b])
            cache[key] = out
        return cache[key]

    def sync_flags(self, eng, kind, stride):
        eng.set_global_flags(kind, stride, self.state['flags'].get(kind, 0))

    def finish(self, series, path):
        self.state['counts'].append([s.nrec for s in series])
        self.state['csv_parts'] = self.csv_parts
        self._write_series_parts(series, path)
        if self.rank == 0:
            self.state['series0'] = series
        if self.rank + 1 == self.world:
            self._assemble(self.state['series0'], self.state['counts'], path)


def sofa_preprocess_segmented(cfg, engine, segment_bytes):
    """sofa_preprocess(cfg) with at most ~segment_bytes of any source resident on the device at a time."""
    import glob
    from . import hostlogic as hl
    from .sofa_preprocess import sofa_preprocess
    logdir = cfg.logdir
    if getattr(cfg, 'cuda_api_tracing', False):
        raise NotImplementedError("segmented preprocess does not cover --cuda_api_tracing")
    ratio = int(cfg.plot_ratio)
    files = {hl.SRC_PERF: logdir + 'perf.script', hl.SRC_GPU: logdir + 'gputrace.tmp', hl.SRC_NET: logdir + 'net.tmp'}
    sizes = [os.path.getsize(p) for p in files.values() if os.path.isfile(p)]
    K = max(1, max([-(-sz // int(segment_bytes)) for sz in sizes] or [1]))
    if K == 1:
        return sofa_preprocess(cfg, engine=engine)
    state = {'flags': {}, 'counts': [], 'ratio': ratio}          # plot_ratio > 1: segments start at multiples of it (aligned_range)
    # ---- pass 1: file-level dtype flags of the gpu and packet sources (parse only)
    time_base = hl.read_time_base(logdir, cfg.cpu_time_offset)
    for kind in (hl.SRC_NET, hl.SRC_GPU):
        path = files[kind]
        if not os.path.isfile(path) or os.path.getsize(path) == 0:
            continue
        if kind == hl.SRC_GPU and not glob.glob(logdir + "gputrace*[0-9].nvvp"):
            continue
        data_off, p = 0, hl.base_params(kind, time_base, cfg.absolute_timestamp, 0.0)
        if kind == hl.SRC_GPU:
            with open(path, 'rb') as f:
                hdr = hl.gpu_header(f.read(1 << 16))
            if hdr is None:
                continue
            upd, data_off = hdr
            upd.pop('t_base')
            p.update(upd)
        fl = 0
        for i in range(K):
            text, off = SegmentIO(i, K, state).read(path, data_off)
            engine.load(kind, text, p, off)
            fl |= engine.row_flags(kind, ratio if kind == hl.SRC_NET else 1)     # nettrace.csv holds the down-sampled rows
        state['flags'][kind] = fl
    # ---- pass 2: every segment through the ordinary preprocess
    for i in range(K):
        sofa_preprocess(cfg, engine=engine, io=SegmentIO(i, K, state))
