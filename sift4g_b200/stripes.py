"""NVLink-striped database across the GPUs of one box (one process per GPU under torchrun).

Every rank keeps one stripe of the residue array resident (s4g_stripe_create), hands the stripe's file descriptor to
its peers over a Unix socket (SCM_RIGHTS) and maps all stripes into one contiguous virtual range (s4g_view_open): the
kernels of a rank then see the WHOLE database and read the peers' pages over NVLink.  A rank runs the hot path for its
own slice of the queries, so candidate lists and hits never have to be merged (include/sift4g_b200.h, csrc/view.cu).
torch.distributed is plumbing only (barriers); no collective is on the data path.
"""
import os
import socket
import threading

import numpy as np

from . import capi

TAIL_PAD = 256          # S4G_DB_TAIL_PAD
_calls = [0]


def stripe_bounds(shard_starts, total_bytes, gran):
    """Byte boundaries P[0..n] of the stripes (multiples of `gran`, every stripe at least one granule) for shards that
    start at shard_starts[s]: the boundary nearest to the shard's first byte, so a rank's own sequences are (up to one
    granule at either end) in its own HBM.  Placement is only a locality hint -- the view is one flat array."""
    n = len(shard_starts)
    P = [0]
    for s in range(1, n):
        P.append(gran * max(int(round(shard_starts[s] / gran)), P[-1] // gran + 1))
    end = -(-(int(total_bytes) + TAIL_PAD) // gran) * gran
    P.append(max(P[-1] + gran, end))
    return P


def exchange_fds(my_fd, rank, world, barrier, tag=None):
    """Every rank offers one file descriptor; returns the list of all ranks' descriptors as valid in THIS process
    (entry `rank` is my_fd itself).  Abstract Unix sockets + SCM_RIGHTS; `barrier` is a callable (dist.barrier)."""
    if world == 1:
        return [my_fd]
    _calls[0] += 1
    base = "\0s4g-%s-%s-%d-" % (os.environ.get("MASTER_PORT", "0"), tag or "fd", _calls[0])
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(base + str(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            c, _a = srv.accept()
            socket.send_fds(c, [b"x"], [my_fd])
            c.close()

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    barrier()                                   # every listener is up
    fds = [None] * world
    fds[rank] = my_fd
    for p in range(world):
        if p == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(base + str(p))
        _msg, got, _flags, _addr = socket.recv_fds(c, 16, 1)
        c.close()
        if len(got) != 1:
            raise RuntimeError("rank %d: no file descriptor received from rank %d" % (rank, p))
        fds[p] = got[0]
    th.join()
    srv.close()
    barrier()
    return fds


class StripedDatabase:
    """The whole database as one view on this rank's GPU.

    ctx            capi.Context of this rank
    shard_codes    this rank's residues (device tensor / pointer holding global bytes [offsets[lo], offsets[hi]))
    offsets        GLOBAL int64 offsets[n + 1] (numpy; every rank holds them)
    lo, hi         the sequences this rank contributes
    dist           torch.distributed (initialised) or None
    emulate        single process: cut the database into this many stripes on the one GPU (tests on a 1-GPU box;
                   shard_codes then holds the whole database)
    """

    def __init__(self, ctx, shard_codes, offsets, lo, hi, dist=None, emulate=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        total = int(offsets[-1])
        gran = int(ctx.lib.s4g_stripe_granularity(ctx.h))
        if gran <= 0:
            raise capi.S4GError("sift4g_b200: no virtual-memory allocation granularity: %s" % ctx.lib.s4g_last_error(ctx.h).decode())
        world = dist.get_world_size() if dist is not None else 1
        rank = dist.get_rank() if dist is not None else 0
        self.ctx, self.dist = ctx, dist
        self.stripes, self.view, self.db = [], None, None
        if dist is None:
            parts = max(1, int(emulate))
            starts = [int(offsets[n * s // parts]) for s in range(parts)]
            P = stripe_bounds(starts, total, gran)
            self.stripes = [capi.Stripe(ctx, P[s + 1] - P[s]) for s in range(parts)]
            members = list(self.stripes)
            if parts > 1:                       # one stripe travels as a file descriptor, like a peer's would
                fd = self.stripes[-1].export_fd()
                members[-1] = fd
            self.view = capi.View(ctx, members, [P[s + 1] - P[s] for s in range(parts)])
            if parts > 1:
                os.close(fd)
            self.view.write(int(offsets[lo]), shard_codes, int(offsets[hi] - offsets[lo]))
            ctx.sync()
        else:
            starts = [int(offsets[n * s // world]) for s in range(world)]
            P = stripe_bounds(starts, total, gran)
            mine = capi.Stripe(ctx, P[rank + 1] - P[rank])
            self.stripes = [mine]
            fd = mine.export_fd()
            fds = exchange_fds(fd, rank, world, dist.barrier, tag="stripe")
            members = [mine if p == rank else fds[p] for p in range(world)]
            self.view = capi.View(ctx, members, [P[s + 1] - P[s] for s in range(world)])
            for f in fds:
                os.close(f)
            self.view.write(int(offsets[lo]), shard_codes, int(offsets[hi] - offsets[lo]))     # peer stores where the shard overhangs
            ctx.sync()
            dist.barrier()                      # every shard is in place
        self.db = self.view.database(offsets)
        ctx.sync()
        if dist is not None:
            dist.barrier()
        self.bounds = P

    def close(self):
        if self.db is not None:
            self.db.close(); self.db = None
        if self.view is not None:
            self.ctx.sync()
            if self.dist is not None:
                self.dist.barrier()             # no peer is still reading
            self.view.close(); self.view = None
        for s in self.stripes:
            s.free()
        self.stripes = []
