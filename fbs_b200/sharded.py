"""Particle-sharded CSMC sweep: ONE chain whose particle set is block-partitioned over the ranks of a
``torch.distributed`` group (BASELINE.json configs[4]: a CelebA-HQ-shaped score network with a very large particle set
on 8 GPUs).  Everything else in this package shards by independent chains and needs no communication
(fbs_b200/parallel.py); this module is the one place with a data-path exchange.

Per time step (csmc.py:132-148), with N particles, rank r owning rows [r n, (r + 1) n), n = N / G:

1. **weights** -- every rank holds the unnormalised log-weights of its n children; ONE all-gather of N floats gives
   every rank the full vector.  Normalisation, the cumulative sum and the conditional resampling are then computed
   redundantly and deterministically on every rank by the same kernel as on one GPU, so the ancestor indices are
   *bit-identical to the unsharded sweep for any G* -- the summation-order contract of DESIGN.md needs no
   "offset of shard totals" arithmetic whose rounding would depend on G.  Cost: 4 N bytes per step (64 KB at
   N = 16384), against 4 du bytes *per moved particle* (12 KB at CelebA-64 inpainting) for step 2.
2. **particle exchange** -- child j of rank r needs parent row A[j], which lives on rank A[j] // n.  On GPUs the ancestor
   gather reads it straight from the owner's memory: every rank's particle rows live in two ping-pong buffers shared with
   the other ranks through CUDA IPC, and ``fbs_gather_rows_peer_f32`` loads each parent row over NVLink inside the gather
   kernel (no staging, no send / receive lists).  The all-gather of step 1 doubles as the ordering point: a rank's new rows
   are written before it joins the collective, so they are complete before any peer's next gather runs, and the buffer a
   rank overwrites was last read before the previous collective.  ``FBS_SHARD_EXCHANGE=nccl`` selects the alternative --
   every rank knows all of A, hence its send and receive lists without a handshake, and the rows travel as one batch of
   point-to-point sends / receives (also what the gloo CPU tests exercise).  With ``killing`` resampling surviving
   particles keep their own slot (resamplings.py:72-74), so only killed particles move.
3. **transition + weight** on the local parents: one score-network evaluation per rank on its n particles, the
   transition noise being rows [r n, (r + 1) n) of the same ``normal(key, (N, p, c))`` array as unsharded.

The result is the sweep of fbs_b200.samplers.csmc.csmc.forward_pass_nn, row-partitioned.
"""
import os
import numpy as np
import torch
import torch.distributed as dist


class ParticleShard:
    """Block partition of N particle rows over the ranks of a process group."""

    def __init__(self, N: int, rank: int, world: int):
        if N % world:
            raise ValueError(f'the particle count ({N}) must be a multiple of the number of ranks ({world})')
        self.N, self.rank, self.world = int(N), int(rank), int(world)
        self.n = self.N // self.world
        self.lo, self.hi = self.rank * self.n, (self.rank + 1) * self.n

    def owner(self, rows):
        return rows // self.n


def exchange_plan(A, shard: ParticleShard):
    """Send / receive lists of rank ``shard.rank`` for the global ancestor vector ``A`` (int64 [N], identical on all ranks).

    Returns ``(send_rows, recv_pos)``: ``send_rows[d]`` = local row indices this rank sends to rank d (in increasing
    child order), ``recv_pos[s]`` = local child positions filled by what rank s sends (same order).  ``d == rank``
    entries describe the local copies.
    """
    A = A.to(torch.int64)
    own = torch.div(A, shard.n, rounding_mode='floor')          # rank owning each child's parent
    send_rows, recv_pos = [], []
    mine = A[shard.lo:shard.hi]
    src = own[shard.lo:shard.hi]
    for r in range(shard.world):
        recv_pos.append(torch.nonzero(src == r, as_tuple=False).reshape(-1))            # my children fed by rank r
        seg = A[r * shard.n:(r + 1) * shard.n]                                          # parents of rank r's children
        send_rows.append(seg[own[r * shard.n:(r + 1) * shard.n] == shard.rank] - shard.lo)  # ... that are mine
    assert int(sum(p.numel() for p in recv_pos)) == shard.n
    return send_rows, recv_pos, mine


def exchange_rows(rows_local, A, shard: ParticleShard, group=None):
    """parents[j] = rows_global[A[lo + j]] for the n children of this rank; ``rows_local`` = rows [lo, hi) of the global
    array (any trailing shape).  One batch of point-to-point transfers; returns ``(parents_local, moved_rows)``."""
    send_rows, recv_pos, _ = exchange_plan(A, shard)
    out = torch.empty_like(rows_local)
    ops, bufs = [], []
    for r in range(shard.world):
        if r == shard.rank:
            out[recv_pos[r]] = rows_local[send_rows[r]]
            continue
        if send_rows[r].numel():
            payload = rows_local[send_rows[r]].contiguous()
            bufs.append(payload)
            ops.append(dist.P2POp(dist.isend, payload, r if group is None else dist.get_global_rank(group, r), group))
        if recv_pos[r].numel():
            buf = torch.empty((recv_pos[r].numel(),) + tuple(rows_local.shape[1:]), dtype=rows_local.dtype, device=rows_local.device)
            bufs.append((buf, recv_pos[r]))
            ops.append(dist.P2POp(dist.irecv, buf, r if group is None else dist.get_global_rank(group, r), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    moved = 0
    for b in bufs:
        if isinstance(b, tuple):
            out[b[1]] = b[0]
            moved += b[1].numel()
    return out, moved


class PeerRows:
    """This rank's particle rows as two ping-pong buffers ``[2, n, row]`` that the other ranks of the box can read directly:
    the buffers are exported with CUDA IPC, the 64-byte handles travel through ``all_gather_object``, and every rank holds a
    device table ``[2, G]`` of pointers (its own buffer + the imports) for ``fbs_gather_rows_peer_f32``."""

    def __init__(self, shard: ParticleShard, row: int, device, group=None):
        import ctypes as C
        from . import _native as nat
        self.shard, self.row, self.group = shard, int(row), group
        self.buf = torch.empty((2, shard.n, self.row), dtype=torch.float32, device=device)
        self._imports = []
        self.error = None
        bases = [None] * shard.world
        bases[shard.rank] = self.buf.data_ptr()
        if shard.world > 1:
            # every rank goes through the same collectives whatever happens locally (a rank that raised before a collective
            # would leave the others waiting in it); failures are agreed on afterwards
            mine = None
            try:
                handle = (C.c_ubyte * 64)()
                off = C.c_int64(0)
                nat.call('fbs_ipc_export', self.buf.data_ptr(), handle, C.byref(off))
                mine = (bytes(handle), int(off.value))
            except Exception as e:                                   # noqa: BLE001
                self.error = repr(e)
            everyone = [None] * shard.world
            dist.all_gather_object(everyone, mine, group=group)
            if self.error is None and all(x is not None for x in everyone):
                try:
                    for r, (hb, o) in enumerate(everyone):
                        if r == shard.rank:
                            continue
                        hbuf = (C.c_ubyte * 64).from_buffer_copy(hb)
                        out = C.c_void_p()
                        nat.call('fbs_ipc_import', hbuf, o, C.byref(out))
                        self._imports.append((out.value, o))
                        bases[r] = out.value
                except Exception as e:                               # noqa: BLE001
                    self.error = repr(e)
            elif self.error is None:
                self.error = 'a peer could not export its buffer'
            ok = torch.tensor([0 if self.error else 1], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.error = self.error or 'a peer could not import the buffers'
                self.close()
                self.table = None
                return
        slot_bytes = shard.n * self.row * 4
        self.table = torch.tensor([[b + s * slot_bytes for b in bases] for s in range(2)], dtype=torch.int64, device=device)
        self.cur = 0

    def local(self):
        """This rank's current rows ``[n, row]`` (a view of the shared buffer)."""
        return self.buf[self.cur]

    def gather(self, parents_global, out):
        """``out[j] = rows_global[parents_global[j]]``: every parent row is loaded from the GPU that owns it."""
        from . import _native as nat
        from ._tensor import ptr, stream
        nat.call('fbs_gather_rows_peer_f32', stream(), ptr(self.table[self.cur]), ptr(parents_global), out.shape[0], self.row,
                 self.shard.n, self.shard.world, ptr(out))

    def next_rows(self):
        """The OTHER buffer: where the step's kernel writes the new rows directly (no staging copy)."""
        return self.buf[self.cur ^ 1]

    def flip(self):
        """Make the buffer written through :meth:`next_rows` current.  The caller's next collective (the all-gather of the
        log-weights) is the point after which the peers read it."""
        self.cur ^= 1

    def publish(self, rows):
        """Copy rows computed elsewhere into the other buffer and make it current."""
        self.next_rows().copy_(rows.reshape(self.shard.n, self.row))
        self.flip()

    def close(self):
        from . import _native as nat
        for p_, o in self._imports:
            nat.call('fbs_ipc_release', p_, o)
        self._imports = []


_PEER_CACHE = {}


def _peer_rows(shard: ParticleShard, row: int, device, group=None) -> PeerRows:
    """One :class:`PeerRows` per (group, partition, row size): exporting / importing the IPC handles costs milliseconds, a
    sweep step microseconds.  Reuse across sweeps is safe: after a sweep's last all-gather no rank reads the buffers, and the
    next sweep starts with a collective after its initial rows are written."""
    key = (id(group), shard.N, shard.rank, shard.world, int(row), str(device))
    pr = _PEER_CACHE.get(key)
    if pr is None:
        pr = _PEER_CACHE[key] = PeerRows(shard, row, device, group)
    return pr


def close_peer_buffers():
    """Release every imported peer buffer (call on all ranks, after the last sharded sweep, before destroying the group)."""
    for pr in _PEER_CACHE.values():
        pr.close()
    _PEER_CACHE.clear()


def all_gather_rows(x_local, shard: ParticleShard, group=None):
    """Concatenation over ranks of equally sized local pieces (the N log-weights of a step)."""
    full = torch.empty((shard.N,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(full, x_local.contiguous(), group=group)
    return full


def forward_pass_sharded(key, us_star, bs_star, vs, model, init, cond_resampling, nsamples, group=None, history=False):
    """csmc.py:132-164 for one chain over a :class:`fbs_b200.nn.ScoreNetModel`, particle rows sharded over ``group``.

    Arguments as ``forward_pass_nn`` (identical on every rank).  Returns a dict with this rank's shard:
    ``us_last [n, p, c]``, the FULL ``log_ws_last [N]`` (identical on every rank), ``lo`` / ``hi``, ``moved`` = rows received
    from other ranks per step, and with ``history`` the local ``uss [K + 1, n, p, c]`` plus the full ``As [K, N]`` /
    ``log_wss [K + 1, N]``.
    """
    from .samplers.csmc.csmc import NormalInit, DegenerateInit, _scheme_of
    from ._tensor import dev, empty, ptr, stream
    from . import random as frandom, _native as nat
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    scheme = _scheme_of(cond_resampling, 'conditional')
    k = dev(key, torch.uint32).reshape(2)
    K, p, q, c = model.K, model.p, model.q, model.c
    us_star = dev(us_star, torch.float32).reshape(K + 1, p, c)
    v = dev(vs, torch.float32).reshape(K + 1, q, c)
    bs = dev(bs_star, torch.int32).reshape(K + 1)
    bs_host = [int(b) for b in bs.cpu().tolist()]
    ks = frandom.split(k, 2)
    key_init, key_scan = ks[0].contiguous(), ks[1].contiguous()
    sk = frandom.split(frandom.split(key_scan, K), 2).contiguous()
    ts = model.ts
    N = int(nsamples) + 1 if isinstance(init, NormalInit) else init.nparticles
    shard = ParticleShard(N, rank, world)
    lo, hi, n = shard.lo, shard.hi, shard.n

    def pin(us_local, row, value):
        if lo <= row < hi:
            us_local[row - lo].copy_(value)

    if isinstance(init, NormalInit):
        us = frandom.normal(key_init, (N, p, c))[lo:hi].contiguous()               # rows of the same global draw
        pin(us, bs_host[0], us_star[0])
        lw_local = model.likelihood_logpdf(v[0], us, v[1], ts[0])
        lw = all_gather_rows(lw_local, shard, group)
    elif isinstance(init, DegenerateInit):
        us = us_star[0:1].expand(n, p, c).contiguous()
        lw = torch.full((N,), init.init_log_w, dtype=torch.float32, device=us.device)
    else:
        raise TypeError('init must be a DegenerateInit or NormalInit')
    # per step: normalise + exp (one launch, redundantly on every rank), conditional resampling (ditto), the ancestor gather
    # over peer memory, image assembly + score network + Euler--Maruyama step / weights / reference pin written straight
    # into the peer-visible buffer, ONE all-gather of the new log-weights -- no eager tensor arithmetic
    from .nn import ops as nnops
    log_w, w = empty((N,), torch.float32), empty((N,), torch.float32)
    nnops.normalise_logw(lw.contiguous(), log_w, w)
    As = log_wss = uss = None
    if history:
        As = empty((K, N), torch.int32)
        log_wss = empty((K + 1, N), torch.float32)
        uss = empty((K + 1, n, p, c), torch.float32)
        log_wss[0].copy_(log_w)
        uss[0].copy_(us)
    A = empty((1, N), torch.int32)
    moved = []
    # particle exchange: 'peer' = the gather kernel loads parent rows from the owners' memory over NVLink (CUDA IPC, default on
    # GPUs); 'nccl' = one batch of point-to-point sends / receives (also what the gloo CPU tests exercise)
    mode = os.environ.get('FBS_SHARD_EXCHANGE', 'peer' if us.is_cuda else 'nccl')
    peer = None
    if mode == 'peer':
        peer = _peer_rows(shard, p * c, us.device, group)                          # IPC set-up once per (group, shape)
        if peer.table is None:
            raise RuntimeError('peer-memory particle exchange is unavailable on this box (' + str(peer.error) + '); '
                               'set FBS_SHARD_EXCHANGE=nccl for the send / receive exchange')   # raised on EVERY rank
        peer.cur = 0
        peer.buf[0].copy_(us.reshape(n, p * c))
        if world > 1:
            dist.all_reduce(torch.zeros(1, device=us.device), group=group)         # every rank's initial rows are in place
    A_all = empty((K, N), torch.int32)                                             # every step's ancestors (for `moved`)
    parents = torch.empty((n, p, c), dtype=torch.float32, device=us.device)
    lw_full = empty((N,), torch.float32)
    for kk in range(K):
        A = A_all[kk:kk + 1]
        nat.call('fbs_cond_resample_f32', stream(), scheme, ptr(sk[kk, 0]), ptr(w), ptr(bs[kk:kk + 1]), ptr(bs[kk + 1:kk + 2]), 1,
                 1, N, ptr(A))                                                     # identical on every rank
        if peer is not None:
            peer.gather(A[0, lo:hi], parents)                                      # a contiguous slice: no copy
            us, lw_local = model.step(parents, v[kk], v[kk + 1], ts[kk], sk[kk, 1], row_offset=lo, rows_total=N,
                                      pin_row=bs[kk + 1:kk + 2], pin_value=us_star[kk + 1], out=peer.next_rows())
            peer.flip()
        else:
            parents, mv = exchange_rows(us, A[0], shard, group)
            moved.append(mv)
            us, lw_local = model.step(parents, v[kk], v[kk + 1], ts[kk], sk[kk, 1], row_offset=lo, rows_total=N,
                                      pin_row=bs[kk + 1:kk + 2], pin_value=us_star[kk + 1])
        dist.all_gather_into_tensor(lw_full, lw_local, group=group)
        nnops.normalise_logw(lw_full, log_w, w)
        if history:
            As[kk].copy_(A[0])
            log_wss[kk + 1].copy_(log_w)
            uss[kk + 1].copy_(us)
    if peer is not None:
        mine = A_all[:, lo:hi]
        moved = [int(m) for m in ((mine < lo) | (mine >= hi)).sum(dim=1).cpu().tolist()]
    if peer is not None:
        us = us.reshape(n, p, c).clone()            # the peer buffers are reused by the next sweep
    return dict(N=N, lo=lo, hi=hi, us_last=us, log_ws_last=log_w, moved=moved, As=As, log_wss=log_wss, uss=uss,
                exchange=mode, ancestors=A_all)
