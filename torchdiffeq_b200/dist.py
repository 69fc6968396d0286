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


def make_reduce(process_group, segs, device, replicated=()):
    """Returns (reduce_fn, n_global, seg_counts_global) for AdaptiveEngine.

    reduce_fn(buf) sums the float64 buffer [n_seg + 1] across ranks in place, on the current stream
    (capturable with the NCCL backend).  Segment element counts are summed once, here.
    `replicated`: indices of segments every rank holds IDENTICALLY (the adjoint's vjp_t and parameter-gradient
    segments after their per-evaluation all-reduce): their sums arrive R times in the all-reduce, so their count is
    R x len and the mean is the local one."""
    pg = None if process_group is True else process_group
    counts = torch.tensor([int(l) for _, l in segs], dtype=torch.int64, device=device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)       # replicated: same len everywhere => R x len
    counts_list = [int(c) for c in counts.tolist()]

    def reduce_fn(buf):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=pg)

    return reduce_fn, sum(counts_list), counts_list


class PeerExchange:
    """NVLink peer-memory exchange of the norm partials, fused into tdq_controller (include/tdq.h).

    One cudaMalloc'ed exchange buffer per rank, exported with CUDA IPC, opened by every peer of the
    same node; the controller kernel then does the all-reduce itself (P2P stores + release/acquire
    flags), so an attempt contains no collective launch at all."""

    def __init__(self, process_group, device):
        import ctypes as C
        from . import _lib
        self._lib_mod = _lib
        self.lib = _lib.load()
        pg = None if process_group is True else process_group
        self.rank, self.world = dist.get_rank(pg), dist.get_world_size(pg)
        self.own, self.peers, self._opened = None, [], []
        # Every step below is collective and every rank takes part in all of them, whatever happened locally,
        # so that a failure anywhere (IPC not permitted, too many ranks) makes ALL ranks fall back together.
        err, hbytes = None, None
        try:
            if self.world > _lib.TDQ_MAX_RANKS:
                raise _lib.TdqError("peer exchange supports at most %d ranks" % _lib.TDQ_MAX_RANKS)
            own = C.c_void_p()
            h = _lib.IpcHandle()
            _lib.check(self.lib.tdq_xchg_create(C.byref(own), C.byref(h)))
            self.own = own.value
            hbytes = bytes(h.bytes)
        except Exception as e:
            err = "%s: %s" % (type(e).__name__, e)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (err, hbytes), group=pg)
        errs = [g[0] for g in gathered if g[0] is not None]
        if not errs:
            try:
                for r, (_, hb) in enumerate(gathered):
                    if r == self.rank:
                        self.peers.append(self.own)
                        continue
                    ph = _lib.IpcHandle()
                    C.memmove(ph.bytes, hb, 64)
                    ptr = C.c_void_p()
                    _lib.check(self.lib.tdq_xchg_open(C.byref(ph), C.byref(ptr)))
                    self.peers.append(ptr.value)
                    self._opened.append(ptr.value)
            except Exception as e:
                err = "%s: %s" % (type(e).__name__, e)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, err, group=pg)       # doubles as the "everybody has mapped everybody" barrier
            errs = [g for g in gathered if g is not None]
        if errs:
            self.close()
            raise _lib.TdqError("peer exchange unavailable on at least one rank: %s" % errs[0])
        self.ptrs = _lib.ptr_array(self.peers)
        self.epoch = 0

    def arm(self, ctrl_ptr, stream):
        """Call after tdq_ctrl_init of every solve (solves are collective, so epochs agree)."""
        self.epoch += 1
        self._lib_mod.check(self.lib.tdq_ctrl_set_exchange(ctrl_ptr, self.ptrs, self.rank, self.world, self.epoch,
                                                           stream))

    def close(self):
        try:
            for p in self._opened:
                self.lib.tdq_xchg_close(p)
            self._opened = []
            if self.own:
                self.lib.tdq_xchg_destroy(self.own)
                self.own = None
        except Exception:
            pass

    def __del__(self):
        self.close()
