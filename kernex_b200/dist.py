"""Axis-0 slab decomposition of kmap stencils over several GPUs (one process per GPU).

The reference has no multi-device path beyond ``pmap`` with one *window* per device
(``kernex/_src/utils.py:28``).  Windows of a kmap are independent, so a large grid
shards naturally into row slabs: rank ``r`` owns ``rows_per_rank`` rows
of a global ``(world*rows_per_rank, ...)`` array, and applying a stencil whose window
spans ``k0`` rows needs ``(k0-1)//2`` rows from the rank above and ``k0//2`` rows from
the rank below.  Axis-0 halos of a C-contiguous array are themselves contiguous, so they
travel with plain ``isend/irecv`` (NCCL over NVLink on GPUs, gloo in the CPU tests) --
no pack kernel.  Per step:

    side stream : post irecv/isend of the halo rows
    main stream : interior launch (windows that touch no halo)     <- overlaps the exchange
    main stream : wait for the halos, two O(halo)-row strip launches

:class:`SlabLayout` is pure geometry (testable on CPU), :func:`exchange_halos` is the
only communication, :class:`SlabStencil` glues them to the CUDA engine.
"""

from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

__all__ = ["SlabLayout", "halo_pull_plan", "exchange_halos", "exchange_halo_grads", "SlabStencil", "ColumnShardedScan",
           "column_shard_plan", "column_exchange_plan"]


@dataclass(frozen=True)
class Piece:
    """One launch: stencil applied to ``ext[in0:in1]`` with axis-0 border ``border0``
    writes local output rows ``[out0, out1)``."""

    in0: int
    in1: int
    border0: tuple
    out0: int
    out1: int
    needs_halo: bool


class SlabLayout:
    """Geometry of one rank's slab for a stride-1 stencil with ``k0`` rows per window and
    global axis-0 border ``(lo, hi)``.

    The rank keeps an *extended* buffer ``ext = [top halo | own rows | bottom halo]``
    (halos absent at the global edges).  Output rows are assigned to the rank that owns
    the window's centre row; rows produced purely by global padding go to the edge ranks.
    """

    def __init__(self, rank: int, world: int, rows: int, k0: int, border0=(0, 0)):
        if rows < k0:
            raise ValueError(f"rows_per_rank={rows} must be >= the window height {k0}")
        self.rank, self.world, self.rows, self.k0 = rank, world, rows, k0
        self.lo, self.hi = int(border0[0]), int(border0[1])
        if self.lo < 0 or self.hi < 0:
            raise ValueError("slab decomposition needs a non-negative axis-0 border")
        self.first, self.last = rank == 0, rank == world - 1
        self.halo_top = 0 if self.first else (k0 - 1) // 2
        self.halo_bot = 0 if self.last else k0 // 2
        self.ext_rows = self.halo_top + rows + self.halo_bot
        # what the neighbours need from me
        self.send_up = 0 if self.first else k0 // 2          # my first rows -> bottom halo of rank-1
        self.send_down = 0 if self.last else (k0 - 1) // 2   # my last rows  -> top halo of rank+1
        pieces = []
        lo_eff = self.lo if self.first else 0
        hi_eff = self.hi if self.last else 0
        if world == 1:
            n_out = rows + self.lo + self.hi - k0 + 1
            pieces.append(Piece(0, rows, (self.lo, self.hi), 0, max(n_out, 0), False))
            self.pieces, self.out_rows = pieces, max(n_out, 0)
            return
        # top strip: every window that touches the top halo (or the global top padding)
        top_in = self.halo_top + k0 - 1
        top_cnt = max(top_in + lo_eff - k0 + 1, 0)
        # bottom strip: windows touching the bottom halo (or the global bottom padding)
        bot_in = self.halo_bot + k0 - 1
        bot_cnt = max(bot_in + hi_eff - k0 + 1, 0)
        mid_cnt = rows - k0 + 1
        o = 0
        if top_cnt:
            pieces.append(Piece(0, top_in, (lo_eff, 0), o, o + top_cnt, not self.first))
        o += top_cnt
        pieces.append(Piece(self.halo_top, self.halo_top + rows, (0, 0), o, o + mid_cnt, False))
        o += mid_cnt
        if bot_cnt:
            pieces.append(Piece(self.ext_rows - bot_in, self.ext_rows, (0, hi_eff), o, o + bot_cnt, not self.last))
        o += bot_cnt
        self.pieces, self.out_rows = pieces, o

    @property
    def own(self):
        return slice(self.halo_top, self.halo_top + self.rows)


def halo_pull_plan(rank: int, world: int, rows: int, k0: int, border0=(0, 0)):
    """One-sided form of :func:`exchange_halos`: ``[(my_ext_rows, peer_rank, peer_ext_rows), ...]`` -- which rows of
    which neighbour's ``ext`` buffer fill my halos (pure geometry; the buffers all start at row 0 of their allocation)."""
    L = SlabLayout(rank, world, rows, k0, border0)
    plan = []
    if not L.first and L.halo_top:       # the last send_down own rows of rank-1
        up = SlabLayout(rank - 1, world, rows, k0, border0)
        a = up.halo_top + rows - up.send_down
        plan.append((slice(0, L.halo_top), rank - 1, slice(a, a + up.send_down)))
    if not L.last and L.halo_bot:        # the first send_up own rows of rank+1
        dn = SlabLayout(rank + 1, world, rows, k0, border0)
        b = L.halo_top + rows
        plan.append((slice(b, b + L.halo_bot), rank + 1, slice(dn.halo_top, dn.halo_top + dn.send_up)))
    return plan


def exchange_halos(ext: torch.Tensor, layout: SlabLayout, group=None):
    """Post the halo traffic of one step; returns the list of outstanding requests."""
    ops = []
    r, ht, hb, R = layout.rank, layout.halo_top, layout.halo_bot, layout.rows
    if not layout.first:
        ops.append(dist.P2POp(dist.irecv, ext[0:ht], r - 1, group))
        ops.append(dist.P2POp(dist.isend, ext[ht:ht + layout.send_up], r - 1, group))
    if not layout.last:
        ops.append(dist.P2POp(dist.irecv, ext[ht + R:ht + R + hb], r + 1, group))
        ops.append(dist.P2POp(dist.isend, ext[ht + R - layout.send_down:ht + R], r + 1, group))
    ops = [op for op in ops if op.tensor.numel() > 0]
    return dist.batch_isend_irecv(ops) if ops else []


def exchange_halo_grads(dext: torch.Tensor, layout: SlabLayout, group=None):
    """Adjoint of :func:`exchange_halos`: the cotangent rows that landed in my halos belong to the
    neighbours' edge rows (they were copies of them), so they are sent back and ADDED there; I receive
    and add the rows my neighbours accumulated for my edge rows.  In place on ``dext``; blocks until
    the traffic is done.  ``<exchange_halos(x), y> == <x, exchange_halo_grads(y)>`` over the ranks."""
    L, r = layout, layout.rank
    ht, hb, R = L.halo_top, L.halo_bot, L.rows
    ops, adds = [], []
    if not L.first:
        buf = torch.empty_like(dext[ht:ht + L.send_up])
        ops.append(dist.P2POp(dist.irecv, buf, r - 1, group))
        ops.append(dist.P2POp(dist.isend, dext[0:ht].contiguous(), r - 1, group))
        adds.append((slice(ht, ht + L.send_up), buf))
    if not L.last:
        buf = torch.empty_like(dext[ht + R - L.send_down:ht + R])
        ops.append(dist.P2POp(dist.irecv, buf, r + 1, group))
        ops.append(dist.P2POp(dist.isend, dext[ht + R:ht + R + hb].contiguous(), r + 1, group))
        adds.append((slice(ht + R - L.send_down, ht + R), buf))
    ops = [op for op in ops if op.tensor.numel() > 0]
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    for sl, buf in adds:
        if buf.numel():
            dext[sl] += buf
    return dext


class SlabStencil:
    """A kmap applied to a row-slab-decomposed global array on CUDA.

    ``transport``:

    * ``"p2p"`` (default on several GPUs): the slab lives in a CUDA-IPC exported buffer (``peer.py``) that the two
      neighbouring ranks have mapped; one ``kx_halo_pull`` launch per step signals "my rows are final" to the
      neighbours, waits for theirs, pulls their edge rows over NVLink and signals "read" back -- no NCCL kernel, no
      rendezvous of host threads, nothing between the previous step and the strip launches but that one kernel.
    * ``"nccl"``: ``batch_isend_irecv`` of the halo rows (the round-1 transport; also what the CPU/gloo tests run).

    ``KX_DIST_TRANSPORT`` overrides the default.  If the peer mapping cannot be set up (no IPC / no peer access) the
    constructor falls back to NCCL on EVERY rank (the decision is all-reduced) and records why in ``transport_note``.
    """

    def __init__(self, func, kernel_size, relative, rows_per_rank, width=None, device=None, padding="valid",
                 trailing_shape=None, group=None, args=(), transport=None, rank=None, world=None, peers=None):
        import os

        from .engine import PreparedMap
        from .resolve import resolve_padding

        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.transport = transport or os.environ.get("KX_DIST_TRANSPORT") or ("p2p" if self.world > 1 else "nccl")
        if self.transport not in ("nccl", "p2p"):
            raise ValueError(f"transport must be 'nccl' or 'p2p', got {self.transport!r}")
        self.transport_note = ""
        kernel_size = tuple(kernel_size)
        trailing = tuple(trailing_shape) if trailing_shape is not None else (int(width),)
        border = resolve_padding(padding, kernel_size)
        self.layout = SlabLayout(self.rank, self.world, rows_per_rank, kernel_size[0], border[0])
        L = self.layout
        self._peers, self._pull, self._epoch = None, None, 0
        if self.transport == "p2p" and self.world > 1:
            self._init_peer_memory(kernel_size[0], rows_per_rank, trailing, peers)
        if self._peers is None:
            self.transport = "nccl" if self.world > 1 else self.transport
            self.ext = torch.zeros((L.ext_rows,) + trailing, dtype=torch.float32, device=self.device)
        strides = (1,) * len(kernel_size)
        self._func_map, self._relative, self._args = {func: ()}, relative, tuple(args)
        self.maps = []
        out_trailing = None
        for p in L.pieces:
            shape = (p.in1 - p.in0,) + trailing
            pm = PreparedMap({func: ()}, shape, kernel_size, strides, (p.border0,) + tuple(border[1:]), relative,
                             args=tuple(args))
            self.maps.append((p, pm))
            out_trailing = tuple(pm.out_shape[1:])
        self.out = torch.empty((L.out_rows,) + out_trailing, dtype=torch.float32, device=self.device)
        self.local_windows = int(self.out.numel())
        # high priority: the O(halo) work must not queue behind the blocks of the interior launch
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self._warm_up()

    def _warm_up(self):
        """Launch every kernel of a step once before the first real step.  CUDA loads kernels lazily, and loading one
        can wait for the running kernels of the context: a first-time load issued while this rank's pull kernel spins
        on a neighbour's flag would serialise behind it.  After this, a step only launches resident kernels."""
        for p, pm in self.maps:
            if self.out[p.out0:p.out1].numel():
                pm(self.ext[p.in0:p.in1], out=self.out[p.out0:p.out1])
        if self._peers is not None:
            lib, C = self._klib.load(), self._C
            s = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            idle = self._klib.KxHaloPull()
            idle.n_sides, idle.epoch, idle.counter = 1, 0, self._pull.counter     # one side, no bytes, no flags
            self._klib.check(lib.kx_halo_pull(C.byref(idle), s))
            self._klib.check(lib.kx_flag_wait(C.c_void_p(self._peers.flag_ptr(self.rank, 0)), C.c_uint64(0), s))
        torch.cuda.synchronize(self.device)

    # ---- peer-memory transport ------------------------------------------------------------------------------
    def _init_peer_memory(self, k0, rows, trailing, peers):
        """Allocate the slab in a peer-mapped buffer (same size on every rank) and prepare the pull descriptor."""
        import ctypes as C

        from . import _lib, peer

        L = self.layout
        row_bytes = 4
        for t in trailing:
            row_bytes *= int(t)
        rows_max = rows + (k0 - 1) // 2 + k0 // 2          # the interior ranks' ext_rows: identical everywhere
        nbytes = _lib.KX_PEER_FLAG_BYTES + rows_max * row_bytes
        why = "" if row_bytes % 16 == 0 else f"rows of {row_bytes} bytes are not 16-byte multiples"
        grp = peers
        if grp is None and not why:
            try:
                grp = peer.PeerGroup.create(nbytes, self.device, self.group)
            except Exception as e:   # no CUDA IPC / peer access on this box
                why = f"peer mapping failed: {e}"
        if peers is None and dist.is_initialized():
            ok = torch.tensor([0 if why else 1], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if not bool(ok.item()):
                if grp is not None:
                    grp.close()
                self.transport_note = why or "peer mapping failed on another rank"
                return
        elif why:
            raise ValueError(f"p2p transport: {why}")
        self._peers = grp
        self._C, self._klib, self._row_bytes = C, _lib, row_bytes
        self.ext = grp.tensor(_lib.KX_PEER_FLAG_BYTES, (L.ext_rows,) + tuple(trailing))
        plan = {pr: (mine, theirs) for mine, pr, theirs in halo_pull_plan(self.rank, self.world, rows, k0, (L.lo, L.hi))}
        self._pull, self._done_flags = self._build_pull(grp, plan, row_bytes)

    def _build_pull(self, grp, plan, row_bytes, accumulate=False):
        """``KxHaloPull`` for ``plan`` = {neighbour rank: (my ext rows, their ext rows)}; both neighbours that exist get a
        side (READY / DONE flags are exchanged with both even when nothing is pulled from one of them)."""
        from . import peer
        _lib = self._klib
        F = _lib.KX_PEER_FLAG_BYTES
        pull = _lib.KxHaloPull()
        pull.n_sides, pull.epoch, pull.accumulate = 0, 0, 1 if accumulate else 0
        pull.counter = grp.local_ptr + peer.COUNTER_BYTE
        for nb, my_ready, their_ready, their_done in ((self.rank - 1, peer.READY_FROM_UP, peer.READY_FROM_DOWN, peer.DONE_FROM_DOWN),
                                                      (self.rank + 1, peer.READY_FROM_DOWN, peer.READY_FROM_UP, peer.DONE_FROM_UP)):
            if not (0 <= nb < self.world):
                continue
            sd = pull.side[pull.n_sides]
            mine, theirs = plan.get(nb, (slice(0, 0), slice(0, 0)))
            sd.bytes = (mine.stop - mine.start) * row_bytes
            sd.dst = grp.local_ptr + F + mine.start * row_bytes
            sd.src = grp.peer_ptrs[nb] + F + theirs.start * row_bytes
            sd.peer_ready = grp.flag_ptr(nb, their_ready)     # the neighbour waits on ITS flag "ready from me"
            sd.my_ready = grp.flag_ptr(self.rank, my_ready)
            sd.peer_done = grp.flag_ptr(nb, their_done)
            pull.n_sides += 1
        done = [grp.flag_ptr(self.rank, sl) for nb, sl in ((self.rank - 1, peer.DONE_FROM_UP), (self.rank + 1, peer.DONE_FROM_DOWN))
                if 0 <= nb < self.world]
        return pull, done

    def _init_grad_peers(self, grad_peers=None):
        """Second peer-mapped buffer, for the cotangent of ``ext`` (collective: every rank makes its first ``vjp`` call
        together).  The adjoint of the halo exchange is a pull-ADD: the cotangent rows that landed in a neighbour's halo
        are pulled over NVLink and added to my edge rows by the same kernel (``accumulate``)."""
        from . import peer
        L = self.layout
        grp = grad_peers if grad_peers is not None else peer.PeerGroup.create(self._peers.nbytes, self.device, self.group)
        self._gpeers, self._gepoch = grp, 0
        self._dext = grp.tensor(self._klib.KX_PEER_FLAG_BYTES, tuple(self.ext.shape))
        plan = {}
        if not L.first and L.send_up:       # rank-1's bottom halo rows hold the cotangents of my first send_up rows
            up = SlabLayout(self.rank - 1, self.world, L.rows, L.k0, (L.lo, L.hi))
            b = up.halo_top + L.rows
            plan[self.rank - 1] = (slice(L.halo_top, L.halo_top + L.send_up), slice(b, b + up.halo_bot))
        if not L.last and L.send_down:      # rank+1's top halo rows hold the cotangents of my last send_down rows
            dn = SlabLayout(self.rank + 1, self.world, L.rows, L.k0, (L.lo, L.hi))
            e = L.halo_top + L.rows
            plan[self.rank + 1] = (slice(e - L.send_down, e), slice(0, dn.halo_top))
        self._gpull, self._gdone = self._build_pull(grp, plan, self._row_bytes, accumulate=True)

    def _pull_halos(self, stream):
        """One launch on ``stream``: READY -> wait -> pull over NVLink -> DONE (kx_peer.cu)."""
        self._epoch += 1
        self._pull.epoch = self._epoch
        self._klib.check(self._klib.load().kx_halo_pull(self._C.byref(self._pull), self._C.c_void_p(stream.cuda_stream)))

    def _wait_neighbours_done(self):
        """Before my own rows change: the neighbours must have pulled the rows of the last step (their DONE flags)."""
        if self._peers is not None and self._epoch:
            s = self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            lib = self._klib.load()
            for f in self._done_flags:
                self._klib.check(lib.kx_flag_wait(self._C.c_void_p(f), self._C.c_uint64(self._epoch), s))

    def close(self):
        if self.__dict__.get("_gpeers") is not None:
            self._gpeers.close()
            self._gpeers = None
        if self._peers is not None:
            self._peers.close()
            self._peers = None

    def fill_random(self, generator=None):
        self._wait_neighbours_done()
        self.ext[self.layout.own].normal_(generator=generator)

    def set_local(self, rows: torch.Tensor):
        self._wait_neighbours_done()
        self.ext[self.layout.own].copy_(rows)

    def step(self):
        main = torch.cuda.current_stream(self.device)
        reqs = []
        if self.world > 1:
            self.side.wait_stream(main)  # own rows written by earlier work on `main` must be visible
            if self._peers is not None:
                self._pull_halos(self.side)
            else:
                with torch.cuda.stream(self.side):
                    reqs = exchange_halos(self.ext, self.layout, self.group)
        for p, pm in self.maps:          # interior first: overlaps the exchange
            if not p.needs_halo:
                pm(self.ext[p.in0:p.in1], out=self.out[p.out0:p.out1])
        # the O(halo)-row strips run on the side stream right behind the halo rows, i.e.
        # concurrently with the tail of the interior launch; `main` joins at the end
        with torch.cuda.stream(self.side):
            for r in reqs:
                r.wait()
            for p, pm in self.maps:
                if p.needs_halo:
                    pm(self.ext[p.in0:p.in1], out=self.out[p.out0:p.out1])
        if self.world > 1:
            main.wait_stream(self.side)
        return self.out

    def step_host(self, x_host: torch.Tensor, y_host: torch.Tensor, chunk_bytes: int = 64 << 20):
        """One application with this rank's rows in PINNED HOST memory on both sides (the end-to-end path of a
        host-resident global grid): ``x_host`` = my ``rows_per_rank`` rows, ``y_host`` = my output rows.

        PCIe is full duplex and the link, not HBM, bounds this call, so the slab is streamed: the edge rows the
        neighbours need go first (their halo pull over NVLink then overlaps everything else), the remaining rows
        follow in chunks on a copy-in stream, every chunk's windows are launched as soon as its rows have landed,
        and finished output rows leave on a copy-out stream.  Synchronises before returning."""
        if self.world == 1:
            raise ValueError("step_host is the multi-GPU path; on one GPU call the kmap on the NumPy array (hostpipe.py)")
        from .engine import PreparedMap
        L, rows, k0 = self.layout, self.layout.rows, self.layout.k0
        if tuple(x_host.shape) != (rows,) + tuple(self.ext.shape[1:]) or tuple(y_host.shape) != tuple(self.out.shape):
            raise ValueError(f"step_host expects host rows {(rows,) + tuple(self.ext.shape[1:])} -> {tuple(self.out.shape)}")
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise ValueError("step_host needs pinned host tensors (asynchronous copies)")
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            self._chunk_maps = {}
        s_in, s_out = self._s_in, self._s_out
        own = self.ext[L.halo_top:L.halo_top + rows]
        self._wait_neighbours_done()
        e_up, e_dn = L.send_up, L.send_down
        if e_up:
            own[:e_up].copy_(x_host[:e_up], non_blocking=True)
        if e_dn:
            own[rows - e_dn:].copy_(x_host[rows - e_dn:], non_blocking=True)
        reqs = []
        self.side.wait_stream(main)
        if self._peers is not None:
            self._pull_halos(self.side)
        else:
            with torch.cuda.stream(self.side):
                reqs = exchange_halos(self.ext, L, self.group)
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        mid = next(p for p, _ in self.maps if p.in0 == L.halo_top and p.in1 == L.halo_top + rows)   # the interior piece
        mid_cnt = mid.out1 - mid.out0
        row_bytes = own[0].numel() * 4
        step_rows = max(int(chunk_bytes // max(row_bytes, 1)), k0)
        c0, j_done = e_up, 0
        hi = rows - e_dn
        while c0 < hi or j_done < mid_cnt:
            c1 = min(c0 + step_rows, hi)
            if c1 > c0:
                with torch.cuda.stream(s_in):
                    own[c0:c1].copy_(x_host[c0:c1], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(s_in)
                main.wait_event(ev_in)
            loaded = rows if c1 >= hi else c1                 # own rows [0, loaded) are on the device (edges went first)
            j_end = min(max(loaded - k0 + 1, j_done), mid_cnt)
            if j_end > j_done:
                n_in = j_end - j_done + k0 - 1
                pm = self._chunk_maps.get(n_in)
                if pm is None:
                    ref = self.maps[[p for p, _ in self.maps].index(mid)][1]
                    border = ((0, 0),) + tuple((int(ref.geom.lo[d]), int(ref.geom.hi[d])) for d in range(1, ref.geom.ndim))
                    ksize = tuple(int(ref.geom.ksize[d]) for d in range(ref.geom.ndim))
                    pm = PreparedMap(self._func_map, (n_in,) + tuple(self.ext.shape[1:]), ksize, (1,) * len(ksize), border,
                                     self._relative, args=self._args)
                    self._chunk_maps[n_in] = pm
                a0 = L.halo_top + j_done
                o0 = mid.out0 + j_done
                pm(self.ext[a0:a0 + n_in], out=self.out[o0:o0 + (j_end - j_done)])
                ev_k = torch.cuda.Event()
                ev_k.record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_k)
                    y_host[o0:o0 + (j_end - j_done)].copy_(self.out[o0:o0 + (j_end - j_done)], non_blocking=True)
                j_done = j_end
            c0 = c1
            if c1 >= hi and j_done >= mid_cnt:
                break
        with torch.cuda.stream(self.side):
            for r in reqs:
                r.wait()
            self.side.wait_stream(s_in)                       # the strips read own rows next to the halos
            for p, pm in self.maps:               # halo strips and, on the edge ranks, the global-padding strips
                if p is not mid and p.out1 > p.out0:
                    pm(self.ext[p.in0:p.in1], out=self.out[p.out0:p.out1])
                    y_host[p.out0:p.out1].copy_(self.out[p.out0:p.out1], non_blocking=True)
        main.wait_stream(self.side)
        main.wait_stream(s_out)
        main.synchronize()
        return y_host

    def vjp(self, g_local: torch.Tensor, want_input=True, want_coef=False, grad_peers=None):
        """Adjoint of :meth:`step`: ``g_local`` is the cotangent of this rank's output rows.

        * input: every piece's flipped-tap stencil of its cotangent rows is accumulated into an
          ext-shaped buffer; the rows that landed in the halos belong to the neighbours, so they travel
          back with the transposed halo exchange (``isend`` of my halo rows, ``irecv`` + add into my
          edge rows).  Returns the cotangent of the OWN rows (a view of a persistent buffer: valid until the next call).
        * coefficients (device weights): per-piece window correlations summed locally, then one
          ``all_reduce`` of ``n_taps`` floats.
        """
        L = self.layout
        g_local = g_local.contiguous()
        dx_own = dcoef = None
        if want_input:
            # the interior piece writes every own row; only the halo rows need zeros before the strips accumulate.
            # ONE persistent buffer (like self.out; the returned rows are valid until the next call): a fresh 1 GB
            # tensor per call is handed to NCCL's stream by the exchange, so the caching allocator could not reuse the
            # previous one and fell back to cudaMalloc every call (1.84 ms instead of 0.40 ms per call at C2 size)
            if self._peers is not None and self.__dict__.get("_gpeers") is None and (grad_peers is not None or dist.is_initialized()):
                self._init_grad_peers(grad_peers)
            gp2p = self.__dict__.get("_gpeers") is not None
            dext = self.__dict__.get("_dext")
            if dext is None or dext.shape != self.ext.shape:
                dext = self._dext = torch.empty(self.ext.shape, dtype=torch.float32, device=self.device)
            if gp2p and self._gepoch:      # the neighbours must have pulled the halo rows of my previous cotangent
                st = self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                for f in self._gdone:
                    self._klib.check(self._klib.load().kx_flag_wait(self._C.c_void_p(f), self._C.c_uint64(self._gepoch), st))
            if L.halo_top:
                dext[:L.halo_top].zero_()
            if L.halo_bot:
                dext[L.halo_top + L.rows:].zero_()
            if not any((not p.needs_halo and p.in0 == L.halo_top and p.in1 == L.halo_top + L.rows and
                        g_local[p.out0:p.out1].shape[0] > 0) for p, _ in self.maps):
                dext.zero_()
            for p, pm in self.maps:
                gp = g_local[p.out0:p.out1]
                if gp.shape[0] == 0:
                    continue
                if not p.needs_halo and p.in0 == L.halo_top and p.in1 == L.halo_top + L.rows:
                    pm.vjp_input(gp, out=dext[p.in0:p.in1])       # interior piece: written in place
            for p, pm in self.maps:
                gp = g_local[p.out0:p.out1]
                if gp.shape[0] == 0 or (not p.needs_halo and p.in0 == L.halo_top and p.in1 == L.halo_top + L.rows):
                    continue
                dext[p.in0:p.in1] += pm.vjp_input(gp)             # O(halo)-row strips
            if self.world > 1:
                if gp2p:               # transposed halo exchange as ONE pull-add kernel over NVLink
                    self._gepoch += 1
                    self._gpull.epoch = self._gepoch
                    st = self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                    self._klib.check(self._klib.load().kx_halo_pull(self._C.byref(self._gpull), st))
                else:
                    exchange_halo_grads(dext, L, self.group)
            dx_own = dext[L.own]
        if want_coef:
            for p, pm in self.maps:
                gp = g_local[p.out0:p.out1]
                if gp.shape[0] == 0:
                    continue
                part = pm.vjp_coef(self.ext[p.in0:p.in1], gp)
                dcoef = part if dcoef is None else dcoef + part
            if self.world > 1:
                dist.all_reduce(dcoef, op=dist.ReduceOp.SUM, group=self.group)
        return dx_own, dcoef


# --------------------------------------------------------------------------- #
# kscan: column (axis-1) sharding of time-marching meshes                       #
# --------------------------------------------------------------------------- #


def column_exchange_plan(nx_global: int, world: int, rank: int, t_rows: int, rl: int, rr: int):
    """Halo-EXCHANGE form: the shard holds its own columns plus ``t_rows*rl`` / ``t_rows*rr`` ghost columns
    (one temporal block of the dependency cone), refreshed from the neighbours before every block.
    Returns ``(col_base, width, keep_lo, keep_hi, ghost_l, ghost_r)`` in the style of
    :func:`column_shard_plan`; ghost widths are rounded up to 4 columns (16-byte rows)."""
    per = -(-nx_global // world)
    per += (-per) % 4
    c0, c1 = min(rank * per, nx_global), min((rank + 1) * per, nx_global)
    gl = 0 if rank == 0 else -(-t_rows * rl // 4) * 4
    gr = 0 if c1 >= nx_global else -(-t_rows * rr // 4) * 4
    # every rank must own columns, and a ghost zone must lie entirely inside the NEIGHBOUR's own columns (it is
    # refreshed from that one rank): otherwise the isend / irecv sizes of the two sides disagree (hang) or the
    # ghost covers columns nobody sends (silently wrong values)
    last_width = nx_global - (world - 1) * per
    if last_width <= 0:
        raise ValueError(f"column exchange: nx={nx_global} leaves rank(s) without columns at world={world} "
                         f"(shards are {per} columns wide)")
    need_l, need_r = -(-t_rows * rl // 4) * 4, -(-t_rows * rr // 4) * 4
    if world > 1 and (need_l > per or need_r > per):   # (the last shard may be narrower: its ghost is clipped to nx)
        raise ValueError(f"column exchange: ghost zones of {need_l} / {need_r} columns do not fit inside the neighbouring "
                         f"shards ({per} columns, last shard {last_width}); use fewer ranks or exchange=False")
    lo, hi = c0 - gl, min(nx_global, c1 + gr)
    return lo, hi - lo, c0 - lo, c1 - lo, gl, hi - c1


def column_shard_plan(nx_global: int, world: int, rank: int, nt: int, rl: int, rr: int):
    """Columns a rank must HOLD to produce its own share of a row-marching kscan without
    talking to its neighbours.

    Axis 0 of a kscan is sequential (time), so the grid shards along axis 1.  A cell at
    row ``n`` depends on columns ``[i - n*rl, i + n*rr]`` of the initial row (``rl`` / ``rr`` =
    largest column offset any function reads to the left / right), so a shard that is
    over-allocated by ``nt*rl`` columns on the left and ``nt*rr`` on the right computes its
    own columns exactly; the overlap is recomputed redundantly (4096 / 2^21 = 0.2 % at
    BASELINE config 4) instead of exchanged.  Returns ``(col_base, width, keep_lo, keep_hi)``:
    hold global columns ``[col_base, col_base+width)``, keep local ``[keep_lo, keep_hi)``.
    """
    per = -(-nx_global // world)
    c0, c1 = min(rank * per, nx_global), min((rank + 1) * per, nx_global)
    lo = max(0, c0 - nt * rl)
    lo -= lo % 4                       # keep rows 16-byte aligned for 128-bit stores
    hi = min(nx_global, c1 + nt * rr)
    return lo, hi - lo, c0 - lo, c1 - lo


class ColumnShardedScan:
    """A ``kscan`` mesh (``F[slice] = fn`` region assignments on a ``kernex_b200.kscan`` object)
    over a global ``(nt, nx_global)`` grid, with the columns sharded over the ranks.  Only
    time-marching meshes qualify (every function reads the row above): the same dependency
    analysis as the single-GPU fast path, enforced by ``kx_scan_rowmarch_shard``."""

    def __init__(self, mesh, nt: int, nx_global: int, rank=None, world=None, device=None, exchange=False,
                 group=None):
        import ctypes as C

        from . import _lib
        from .engine import _geom, _Program, _trace_all
        from .resolve import normalize_slices

        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world, self.nt, self.nx_global = rank, world, nt, nx_global
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        shape = (nt, nx_global)
        ksize, strides, border = mesh._resolved(shape)
        if mesh.use_offset or not mesh.inplace:
            raise ValueError("ColumnShardedScan needs a kscan mesh")
        normalized = normalize_slices(mesh.container, shape)
        func_map = {mesh._named(f, ksize): keys for f, keys in normalized.items()}
        specs, device_args = _trace_all(func_map, ksize, mesh.relative, (), {})
        if device_args:
            raise ValueError("device-resident arguments are not supported in a sharded kscan")
        c1 = (ksize[1] - 1) // 2
        offs = [t[1] - c1 for s in specs for t in s.taps]
        self.rl, self.rr = max([0] + [-o for o in offs]), max([0] + offs)
        self.exchange, self.group = bool(exchange) and world > 1, group
        if self.exchange:
            # north-star form: per temporal block, T*R ghost values per side travel by NCCL send/recv
            self.t_rows = int(_lib.load().kx_scan_rowmarch_rows_per_launch())
            (self.col_base, self.width, self.keep_lo, self.keep_hi, self.ghost_l, self.ghost_r) = \
                column_exchange_plan(nx_global, world, rank, self.t_rows, self.rl, self.rr)
        else:
            # default: recompute the nt*R-column dependency cone, no communication at all
            self.col_base, self.width, self.keep_lo, self.keep_hi = column_shard_plan(
                nx_global, world, rank, nt, self.rl, self.rr)
        self._prog = _Program(specs, [tuple(v) for v in func_map.values()], 2)
        local = (nt, self.width)
        self._geom = _geom(local, local, ksize, strides, border, _lib.KX_F32, _lib.KX_F32)
        self._lib, self._C = _lib, C
        self.out = torch.empty(local, dtype=torch.float32, device=self.device)
        self.local_windows = nt * (self.keep_hi - self.keep_lo)

    def run(self):
        """Launch the scan; returns this rank's columns ``(nt, nx_local)`` as a view of the
        over-allocated buffer (row pitch = held width)."""
        C, L = self._C, self._lib
        lib = L.load()
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if self.exchange:
            return self._run_exchange(lib, stream)
        L.check(lib.kx_scan_rowmarch_shard(C.byref(self._geom), C.byref(self._prog.c), C.c_int64(self.col_base),
                                           C.c_int64(self.nx_global), C.c_void_p(self.out.data_ptr()), stream))
        return self.out[:, self.keep_lo:self.keep_hi]

    def _run_exchange(self, lib, stream):
        """T rows per block; before block b > 0 the ghost columns of row b*T-1 are received from the
        neighbours (they own them) and my edge columns of that row are sent to them."""
        C, L, T = self._C, self._lib, self.t_rows
        for n0 in range(0, self.nt, T):
            if n0 > 0:
                row, ops = self.out[n0 - 1], []
                if self.ghost_l:   # left neighbour owns my left ghost columns; it needs nothing from me unless rr > 0
                    ops.append(dist.P2POp(dist.irecv, row[0:self.ghost_l], self.rank - 1, self.group))
                if self.rank + 1 < self.world and self.rl:
                    nxt = column_exchange_plan(self.nx_global, self.world, self.rank + 1, T, self.rl, self.rr)
                    if nxt[4]:
                        ops.append(dist.P2POp(dist.isend, row[self.keep_hi - nxt[4]:self.keep_hi], self.rank + 1,
                                              self.group))
                if self.ghost_r:
                    ops.append(dist.P2POp(dist.irecv, row[self.keep_hi:self.keep_hi + self.ghost_r], self.rank + 1,
                                          self.group))
                if self.rank > 0 and self.rr:
                    prv = column_exchange_plan(self.nx_global, self.world, self.rank - 1, T, self.rl, self.rr)
                    if prv[5]:
                        ops.append(dist.P2POp(dist.isend, row[self.keep_lo:self.keep_lo + prv[5]], self.rank - 1,
                                              self.group))
                for r in (dist.batch_isend_irecv(ops) if ops else []):
                    r.wait()
            L.check(lib.kx_scan_rowmarch_shard_rows(
                C.byref(self._geom), C.byref(self._prog.c), C.c_int64(self.col_base), C.c_int64(self.nx_global),
                C.c_int32(n0), C.c_int32(min(n0 + T, self.nt)), C.c_void_p(self.out.data_ptr()), stream))
        return self.out[:, self.keep_lo:self.keep_hi]
