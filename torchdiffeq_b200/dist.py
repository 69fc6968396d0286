"""Batch-sharded solves: every rank integrates its own trajectories; the only
exchange is the all-reduce(SUM) of the few float64 partial sums behind each norm, so that all ranks
take the same accept/reject decision and the same next dt as the unsharded reference would
(its RMS norm is a mean over the WHOLE batch, misc.py:22-23)."""
import torch
import torch.distributed as dist


_HOST_GROUPS = {}


def _host_group(pg):
    """A gloo group over the same ranks, for host-side agreement that must not queue behind GPU work."""
    key = id(pg)
    if key not in _HOST_GROUPS:
        if dist.get_backend(pg) == "gloo":
            _HOST_GROUPS[key] = pg
        else:
            ranks = None if pg is None else dist.get_process_group_ranks(pg)
            _HOST_GROUPS[key] = dist.new_group(ranks=ranks, backend="gloo")
    return _HOST_GROUPS[key]


def make_agree(process_group):
    """agree(n) -> max over ranks of n, on the host (gloo).

    Every attempt of a sharded solve contains a collective, so all ranks must queue the SAME number of
    attempts.  Their decisions are identical, but with run-ahead each host notices the end at a slightly
    different time and may have queued a different number of trailing no-op attempts; before waiting for
    its stream each rank tops up to the maximum."""
    pg = None if process_group is True else process_group
    hg = _host_group(pg)

    def agree(n):
        t = torch.tensor([int(n)], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=hg)
        return int(t[0])

    return agree


def make_reduce(process_group, segs, device):
    """Returns (reduce_fn, n_global, seg_counts_global) for AdaptiveEngine.

    reduce_fn(buf) sums the float64 buffer [n_seg + 1] across ranks in place, on the current stream
    (capturable with the NCCL backend).  Segment element counts are summed once, here."""
    pg = None if process_group is True else process_group
    counts = torch.tensor([int(l) for _, l in segs], dtype=torch.int64, device=device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)
    counts_list = [int(c) for c in counts.tolist()]

    def reduce_fn(buf):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=pg)

    return reduce_fn, sum(counts_list), counts_list
