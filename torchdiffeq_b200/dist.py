"""Batch-sharded solves: every rank integrates its own trajectories; the only
exchange is the all-reduce(SUM) of the few float64 partial sums behind each norm, so that all ranks
take the same accept/reject decision and the same next dt as the unsharded reference would
(its RMS norm is a mean over the WHOLE batch, misc.py:22-23)."""
import torch
import torch.distributed as dist


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
