"""Per-iteration profile of the AISI path (reference: bin/sofa_aisi.py:21-59 `iter_profile` over the slices of :429-453).

The reference finds the iteration boundaries with a suffix tree over kernel-name tokens, fuzzy matching and KMeans
(:218-286, 414-419; out of scope) and then, for every pair of consecutive boundaries, slices the gpu
frame of deviceId 1 with two boolean masks and sums columns of the slice with pandas - O(iterations x rows).  Here the
boundaries are an input and all iterations are reduced in one device pass (libsofa_b200: iteration index + keyword
bits per row, radix sort by iteration, numpy-pairwise segment sums).
"""


def iter_profile(cfg, eng, times, device_id=1):
    """[dict | None per iteration] with the reference's field names; None where the reference would skip the iteration
    (no gpu row in it, :451).  cpu / strace / mpstat sums are 0: the reference slices those frames with a mask that
    belongs to the gpu frame (an error on current pandas), see tests/golden/make_golden_aisi.py."""
    out = []
    for st in eng.iter_profile(times, device_id):
        if st.n_rows == 0:
            out.append(None)
            continue
        first, last = float('%.6f' % st.first_ts), float('%.6f' % st.last_ts)           # the CSV values the reference reads
        copy_time = st.copy_name_time + st.allreduce_time                                # :40
        out.append({'elapsed_time': last - first, 'fw_time': st.fw_time, 'bw_time': st.bw_time,
                    'kernel_time': st.gpu_time - copy_time, 'payload': float(st.payload), 'copy_time': copy_time,
                    'allreduce_time': st.allreduce_time, 'gpu_time': st.gpu_time, 'gemm_time': st.gemm_time,
                    'streams': float(st.streams), 'syscall_time': 0.0, 'cpu_time': 0.0, 'mpstat_usr': 0.0,
                    'mpstat_sys': 0.0, 'mpstat_iow': 0.0})
    return out
