"""Multi-GPU ``mkfullsky``: l-sharded fill / root / draw+apply, one all-to-all, channel-sharded SHT.

This is the reference's own MPI design (``cora/core/skysim.py:97-134`` over ``caput.mpiarray``):
``alm_array`` is distributed over l (``:108-110``), each rank roots and applies its local l's
(``:114-121``), ``redistribute(axis=0)`` transposes to a frequency distribution (``:128``) and
every rank transforms its own channels (``:130-134``).  Here: one process per GPU,
``torch.distributed`` (NCCL over NVLink; gloo in the CPU tests) for the single exchange step,
and no other data-path collective.

Differences from the reference, by design:

* l is dealt to ranks **interleaved** (``l mod G``) by default, not in contiguous blocks: the
  apply cost grows like ``l + 1`` so caput's block split is 1.9x imbalanced at G = 8
  (``partition="block"`` reproduces caput's split).
* the exchanged alm are triangular-packed rows ``(l, m <= l)`` -- half the bytes of the
  reference's dense ``L x L`` squares -- and the apply kernel writes them straight into the
  per-destination send slabs (``cora_b200_draw_apply_slabs``).
* Philox draws are keyed by ``(l, m, nu)``, so the maps do not depend on G (the reference's
  pipeline caller seeds every rank identically).
"""

import ctypes
import os

import numpy as np

from . import _dev, _lib, hputil, nputil


def block_partition(n, size, rank):
    """caput.mpiarray's contiguous split: the first ``n % size`` ranks get one extra item."""
    base, rem = divmod(n, size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class ShardPlan(object):
    """Who owns which l and which channel, and where every alm row sits in the exchange buffers.

    Send buffer of rank r (complex elements): one slab per destination s, slab s =
    ``[rows_r][cb_s]`` at offset ``rows_r * chan_lo[s]``; ``rows_r = sum_{l in l_list[r]} (l + 1)``,
    rows ordered by (position of l in ``l_list[r]``, m).  Receive buffer of rank s: one slab per
    source r, ``[rows_r][cb_s]`` at offset ``cb_s * sum_{r' < r} rows_r'``.
    """

    def __init__(self, lmax, nz, size, partition="interleaved"):
        self.lmax, self.nz, self.size, self.partition = int(lmax), int(nz), int(size), partition
        L = self.lmax + 1
        if partition == "interleaved":
            self.l_lists = [np.arange(r, L, size, dtype=np.int32) for r in range(size)]
        elif partition == "block":
            self.l_lists = [np.arange(*block_partition(L, size, r), dtype=np.int32) for r in range(size)]
        else:
            raise ValueError("partition must be 'interleaved' or 'block'")
        bounds = [block_partition(self.nz, size, s) for s in range(size)]
        self.chan_lo = np.array([b[0] for b in bounds], dtype=np.int64)
        self.chan_hi = np.array([b[1] for b in bounds], dtype=np.int64)
        self.cb = self.chan_hi - self.chan_lo
        self.rows = np.array([int((ll.astype(np.int64) + 1).sum()) for ll in self.l_lists], dtype=np.int64)
        self.owner = np.empty(L, dtype=np.int64)      # rank owning l
        self.row0 = np.empty(L, dtype=np.int64)       # row of (l, m=0) inside its owner's slabs
        for r, ll in enumerate(self.l_lists):
            self.owner[ll] = r
            self.row0[ll] = np.concatenate([[0], np.cumsum(ll.astype(np.int64) + 1)[:-1]]) if len(ll) else []

    # ---- sender side (rank r) -------------------------------------------------------
    def send_splits(self, r):
        """Complex elements sent to each destination."""
        return [int(self.rows[r] * c) for c in self.cb]

    def nu_tables(self, r):
        """(nu_base[nz], nu_width[nz]) for ``cora_b200_draw_apply_slabs`` on rank r."""
        base = np.empty(self.nz, dtype=np.int64)
        width = np.empty(self.nz, dtype=np.int32)
        for s in range(self.size):
            lo, hi = int(self.chan_lo[s]), int(self.chan_hi[s])
            base[lo:hi] = self.rows[r] * lo + np.arange(hi - lo)
            width[lo:hi] = hi - lo
        return base, width

    def send_row0(self, r):
        return np.ascontiguousarray(self.row0[self.l_lists[r]], dtype=np.int64)

    # ---- receiver side (rank s) -----------------------------------------------------
    def recv_splits(self, s):
        return [int(self.rows[r] * self.cb[s]) for r in range(self.size)]

    def l_offsets(self, s):
        """Complex offset of row (l, m=0) in rank s's receive buffer, for every l."""
        src_base = np.concatenate([[0], np.cumsum(self.rows)[:-1]]) * int(self.cb[s])
        return np.ascontiguousarray(src_base[self.owner] + self.row0 * int(self.cb[s]), dtype=np.int64)

    def nalm_total(self):
        return int(self.rows.sum())

    # ---- all-gather of the alm (``alms=True``, skysim.py:123-125) -------------------------------
    def gather_tables(self):
        """(nu_base[nz], nu_width[nz]) that make ``cora_b200_draw_apply_slabs`` write ONE slab ``[rows_r][nz]``
        (every channel, not per-destination slabs): what each rank contributes to the all-gather."""
        return np.arange(self.nz, dtype=np.int64), np.full(self.nz, self.nz, dtype=np.int32)

    def gather_rows(self):
        """Rows of the padded per-rank block of the gathered buffer ``[size][gather_rows][nz]``."""
        return int(self.rows.max()) if self.size else 0

    def gather_l_offsets(self):
        """Complex offset of row (l, m=0) in the gathered buffer, for every l (``cora_b200_alm_slabs_to_panel``)."""
        return np.ascontiguousarray((self.owner * self.gather_rows() + self.row0) * self.nz, dtype=np.int64)


def _dist():
    import torch.distributed as dist

    return dist


def exchange(send, plan, rank, group=None, out=None):
    """The one collective of the path: l-major slabs -> channel-major slabs
    (``alm_array.redistribute(axis=0)``, ``skysim.py:128``).  ``send``: complex128 tensor
    (CUDA under NCCL, CPU under gloo).  Returns the receive buffer."""
    import torch

    dist = _dist()
    if plan.size == 1:
        return send   # one rank: the send slab already is the receive slab
    recv = torch.empty(sum(plan.recv_splits(rank)), dtype=send.dtype, device=send.device) if out is None else out
    s_r, r_r = torch.view_as_real(send).reshape(-1), torch.view_as_real(recv).reshape(-1)
    dist.all_to_all_single(r_r, s_r, output_split_sizes=[2 * n for n in plan.recv_splits(rank)],
                           input_split_sizes=[2 * n for n in plan.send_splits(rank)], group=group)
    return recv


def fill_tile_share(ntile, rank, size):
    """(first tile, number of tiles, stride) of a rank's share of the 21cm fill's channel-pair tiles.  Tiles are dealt
    interleaved: a tile's cost grows along the enumeration (with |chi_i - chi_j|), so contiguous ranges leave the last
    rank with the expensive ones (measured at 2 GPUs: 29.3 ms each against 42.6 ms on one GPU; interleaved 21.5 ms)."""
    n = (ntile - rank + size - 1) // size if ntile > rank else 0
    return int(rank), int(n), int(size)


DRAW_WS_BYTES = int(os.environ.get("CORA_B200_DRAW_WS_MB", "2048")) << 20


def draw_workspace_bytes(nz, lmax_loc, nl):
    """Workspace of the draw + apply stage: the Philox draws of a BATCH of l's, regenerated batch after batch in the
    same buffer (stream order is the only synchronisation).  Bounded (default 2 GB, ``CORA_B200_DRAW_WS_MB``) instead
    of holding every local l's draws (19 GB per GPU at nside 1024 x 2048 channels): a batch only has to be large
    enough that its apply launch fills the GPU for many waves."""
    lib = _lib.load()
    full = lib.cora_b200_draw_apply_workspace_bytes(nz, lmax_loc, nl)
    one = lib.cora_b200_draw_apply_workspace_bytes(nz, lmax_loc, 1)
    return int(min(full, max(one, min(DRAW_WS_BYTES, _dev.free_bytes() - (4 << 30)))))


def allgather_alm(slab, plan, rank, group=None):
    """``alm_array.allgather()`` (``skysim.py:125``): every rank contributes its ``[rows_r][nz]`` slab (rows = the
    (l, m <= l) of its l's), padded to ``plan.gather_rows()`` rows, and receives ``[size][gather_rows][nz]``.
    ``slab``: complex128 tensor (CUDA under NCCL, CPU under gloo)."""
    import torch

    dist = _dist()
    n = plan.gather_rows() * plan.nz
    if plan.size == 1:
        return slab
    mine = torch.zeros(n, dtype=slab.dtype, device=slab.device)
    mine[: slab.numel()] = slab.reshape(-1)
    parts = [torch.empty(2 * n, dtype=torch.float64, device=slab.device) for _ in range(plan.size)]
    dist.all_gather(parts, torch.view_as_real(mine).reshape(-1), group=group)
    return torch.view_as_complex(torch.cat(parts).reshape(-1, 2))


class ShardedSky(object):
    """Per-rank state of the sharded generator: plan, device tables, fill inputs.

    ``step()`` runs one full pass (fill -> root -> draw/apply -> all-to-all -> SHT) and returns
    this rank's channels ``float64[cb, npix]`` as a CUDA tensor."""

    def __init__(self, model, nside, frequencies, lmax=None, zromb=3, group=None, partition="interleaved",
                 rank=None, size=None, exchange="auto", peers=None):
        """``exchange``: ``"p2p"`` -- the fill and apply kernels store straight into the consuming
        GPU's buffers over NVLink (``cora_b200.peer``; needs one process per GPU under NCCL, or
        ``peers`` = a ``LocalPeers`` view for virtual ranks); ``"collective"`` -- apply packs send
        slabs and ``torch.distributed.all_to_all_single`` moves them (also the gloo/CPU-testable
        path); ``"auto"`` -- p2p whenever there is more than one rank and peers can be set up."""
        from . import skysim

        t = _dev.torch()
        dist = _dist()
        self.group = group
        if size is None:
            size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if size > 1 else 0
        self.rank, self.size = rank, size
        self.model, self.nside = model, int(nside)
        self.freq = np.asarray(frequencies, dtype=np.float64)
        self.nz = len(self.freq)
        self.lmax = 3 * self.nside - 1 if lmax is None else int(lmax)
        self.plan = ShardPlan(self.lmax, self.nz, size, partition)
        self.l_list = self.plan.l_lists[rank]
        self.nl = len(self.l_list)
        self.cb = int(self.plan.cb[rank])
        self.npix = 12 * self.nside**2
        self.zromb = int(zromb)
        za, self.zint = skysim._sample_frequencies(self.freq, zromb, None)
        self.fill_inputs = model._b200_fill_inputs(za, skysim.romberg_weights(zromb))
        base, width = self.plan.nu_tables(rank)
        self.nu_base, self.nu_width = _dev.to_device(base, t.int64), _dev.to_device(width, t.int32)
        self.l_off = _dev.to_device(self.plan.l_offsets(rank), t.int64)
        self.row0 = self.plan.send_row0(rank)
        self._buf = {}
        self.peers = peers
        requested = exchange
        if exchange == "auto":
            exchange = "p2p" if (size > 1 and (peers is not None or (dist.is_available() and dist.is_initialized()
                                                                      and dist.get_backend(group) == "nccl"))) else "collective"
        self.exchange = exchange if size > 1 else "collective"
        self._p2p = None
        self._k = 0
        if self.exchange == "p2p" and self.peers is None:
            # set the peer group up now.  PeerGroup agrees on failure inside its own collectives and raises
            # PeerSetupError on EVERY rank together (no peer access between some pair of GPUs, ...): then
            # everybody takes the collective path, or everybody raises if p2p was asked for explicitly.
            from . import peer as _peer

            try:
                self.peers = _peer.PeerGroup(self.rank, self.size, self.group)
            except _peer.PeerSetupError as exc:
                if requested == "p2p":
                    raise _lib.CoraB200Error("exchange='p2p' requested but peer memory could not be set up on every rank: %s" % exc)
                self._p2p_error = exc
                self.exchange = "collective"

    def _persistent(self, name, make):
        """Buffers reused from step to step (no allocator traffic inside a step)."""
        if name not in self._buf:
            self._buf[name] = make()
        return self._buf[name]

    def _draw_begin(self, seed, counter0=0):
        """Start this step's Philox draws on a side stream: they do not depend on C_l, so they overlap the fill and
        the root (the draw kernel is FP64-ALU bound, the fill L2-latency bound, the Cholesky latency bound).  Needs a
        buffer for all local draws (``16 nz sum(l+1)`` bytes); if that does not fit, the apply call draws by itself."""
        t = _dev.torch()
        lib = _lib.load()
        self._predrawn = None
        if self.nl == 0 or os.environ.get("CORA_B200_DRAW_OVERLAP", "0") != "1":   # measured: no gain (the fill slows by what the draws take), off by default
            return
        need = int(lib.cora_b200_draw_bytes(_lib.ptr(self.l_list), self.nl, self.nz))
        if "draw_buf" not in self._buf:
            if need + (8 << 30) > _dev.free_bytes():
                return
            self._buf["draw_buf"] = _dev.workspace(need)
            self._draw_stream = t.cuda.Stream()
        buf = self._buf["draw_buf"]
        main = t.cuda.current_stream()
        ev = t.cuda.Event()
        ev.record(main)                       # the previous step's apply has finished reading the buffer
        self._draw_stream.wait_event(ev)
        _lib.call("cora_b200_draw", _lib.ptr(self.l_list), self.nl, self.nz, ctypes.c_ulonglong(int(seed)), int(counter0),
                  _lib.ptr(buf), int(buf.numel()), _lib.stream_ptr(self._draw_stream))
        done = t.cuda.Event()
        done.record(self._draw_stream)
        self._predrawn = (buf, done, int(seed))

    def _draw_take(self, seed):
        """(gauss pointer, gauss_ld) of the draws started by ``_draw_begin`` for this seed (the main stream then waits
        for them), or (None, 0)."""
        pd = getattr(self, "_predrawn", None)
        self._predrawn = None
        if pd is None or pd[2] != int(seed):
            return None, 0
        _dev.torch().cuda.current_stream().wait_event(pd[1])
        return _lib.ptr(pd[0]), -1

    def fill(self, out=None, lower_only=True):
        """This rank's C_l(nu, nu') rows, ``float64[nl, nz, nz]`` (``skysim.clarray`` for the local l's).  By default
        only the lower triangles are filled (``lower_only``): that is all the root stage reads."""
        t = _dev.torch()
        if out is None:
            out = self._persistent("cla", lambda: _dev.zeros((self.nl, self.nz, self.nz), t.float64))
        step = self.size if self.plan.partition == "interleaved" else 1
        self.model._b200_fill(self.fill_inputs, int(self.l_list[0]) if self.nl else 0, step, self.nl, self.nz, self.zint, out,
                              lower_only=lower_only)
        return out

    def alm_local(self, cla, seed=0, gauss=None, roots=None, slab_tables=None):
        """root + draw/apply for the local l's -> send buffer (complex128, one slab per destination).
        ``slab_tables = (nu_base, nu_width, out)`` (device tables + output tensor) replaces the per-destination slab
        layout, e.g. by the single ``[rows][nz]`` slab of the all-gather (``ShardPlan.gather_tables``)."""
        t = _dev.torch()
        lib = _lib.load()
        if roots is None:
            outb = self._persistent("root", lambda: (_dev.empty((self.nl, self.nz, self.nz), t.float64),
                                                     _dev.empty((self.nl,), t.int32), _dev.empty((self.nl,), t.int32)))
            rws = self._persistent("root_ws", lambda: nputil.root_workspace(self.nl, self.nz, max_eigh=self._max_eigh()))
            root, used, _ = nputil.root_batched_device(cla, jitter_rel=1e-14, clip_rel=1e-16, out=outb, ws=rws)
        else:
            root, used = _dev.to_device(roots, t.float64), None
        nu_base, nu_width = self.nu_base, self.nu_width
        if slab_tables is not None:
            nu_base, nu_width, send = slab_tables
        elif self.size > 1:
            send = self._persistent("send", lambda: _dev.empty((int(self.plan.rows[self.rank]) * self.nz,), t.complex128))
        lmax_loc = int(self.l_list.max())
        gptr, gld = self._draw_take(seed) if gauss is None else (None, 0)
        if gauss is None and gptr is None:
            def mk():
                return _dev.workspace(draw_workspace_bytes(self.nz, lmax_loc, self.nl))

            ws = self._persistent("draw_ws", mk)
        elif gauss is None:
            ws = self._persistent("desc_ws", lambda: _dev.workspace(64 * self.nl + 4096))
        else:
            gauss = _dev.to_device(gauss, t.complex128)
            ws = _dev.workspace(64 * self.nl + 4096)
            gptr, gld = _lib.ptr(gauss), int(gauss.shape[-1])
        nbytes = ws.numel()
        if self.size == 1:
            # one GPU: no exchange, so the apply kernel writes the PANEL layout the SHT reads directly
            nalm = (self.lmax + 1) * (self.lmax + 2) // 2
            panel = self._persistent("panel", lambda: _dev.empty((nalm, self.nz), t.complex128))
            _lib.call("cora_b200_draw_apply", _lib.ptr(root), _lib.ptr(self.l_list), _lib.ptr(used), self.nl, self.nz,
                      self.lmax, ctypes.c_ulonglong(int(seed)), gptr, gld, _lib.ptr(panel), self.nz, 0, 0, self.nz,
                      _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
            return panel
        _lib.call("cora_b200_draw_apply_slabs", _lib.ptr(root), _lib.ptr(self.l_list), _lib.ptr(used), self.nl, self.nz,
                  self.lmax, ctypes.c_ulonglong(int(seed)), gptr, gld, _lib.ptr(self.row0), _lib.ptr(nu_base),
                  _lib.ptr(nu_width), _lib.ptr(send), _lib.ptr(ws), int(nbytes), _lib.stream_ptr())
        return send

    def alm_gathered(self, cla, seed=0, gauss=None, roots=None):
        """root + draw/apply for the local l's, then the all-gather over ranks: the whole
        ``complex128[nz, 1, L, L]`` alm array on every rank (``mkfullsky(..., alms=True)``, ``skysim.py:123-125``),
        as a CUDA tensor."""
        t = _dev.torch()
        L = self.lmax + 1
        nalm = L * (L + 1) // 2
        if self.size == 1:
            panel = self.alm_local(cla, seed=seed, gauss=gauss, roots=roots)
        else:
            # one slab [rows_r][nz] (every channel) instead of per-destination slabs
            base, width = self.plan.gather_tables()
            slab = _dev.empty((int(self.plan.rows[self.rank]) * self.nz,), t.complex128)
            if self.nl:
                self.alm_local(cla, seed=seed, gauss=gauss, roots=roots,
                               slab_tables=(_dev.to_device(base, t.int64), _dev.to_device(width, t.int32), slab))
            allb = allgather_alm(slab, self.plan, self.rank, self.group)
            panel = _dev.empty((nalm, self.nz), t.complex128)
            loff = _dev.to_device(self.plan.gather_l_offsets(), t.int64)
            _lib.call("cora_b200_alm_slabs_to_panel", _lib.ptr(allb), _lib.ptr(loff), self.lmax, self.nz, _lib.ptr(panel),
                      self.nz, 0, _lib.stream_ptr())
        return hputil.panel_to_dense(panel, self.lmax, self.nz).reshape(self.nz, 1, L, L)

    def synthesize(self, recv, out=None):
        """received slabs -> PANEL -> maps of this rank's channels."""
        t = _dev.torch()
        nalm = (self.lmax + 1) * (self.lmax + 2) // 2
        if self.size == 1:
            panel = recv    # already in PANEL layout (alm_local)
        else:
            panel = self._persistent("panel", lambda: _dev.empty((nalm, self.cb), t.complex128))
            _lib.call("cora_b200_alm_slabs_to_panel", _lib.ptr(recv), _lib.ptr(self.l_off), self.lmax, self.cb, _lib.ptr(panel),
                      self.cb, 0, _lib.stream_ptr())
        plan = _dev.sht_plan(self.nside, self.lmax)
        ws = self._persistent("sht_ws", lambda: _dev.sht_workspace(plan, _lib.ALM_PANEL, self.cb, reserve=(4 << 30) + 8 * self.cb * self.npix)[0])
        return hputil.alm2map_device(panel, self.nside, self.lmax, _lib.ALM_PANEL, self.cb, self.cb, out=out, ws=ws)

    # ---- fused-exchange path (exchange == "p2p") -----------------------------------------
    def _p2p_setup(self):
        """Peer buffers (double-buffered C_l rows and PANEL) and the pointer tables the kernels use."""
        if self._p2p is not None:
            return self._p2p
        from . import peer as _peer

        t = _dev.torch()
        if self.peers is None:
            self.peers = _peer.PeerGroup(self.rank, self.size, self.group)
        pg = self.peers
        L = self.lmax + 1
        nalm = L * (L + 1) // 2
        st = {"pairs": hasattr(self.model, "_b200_fill_tiles")}
        st["cla"], st["cla_ptrs"], st["panel"], st["panel_ptrs"] = [], [], [], []
        for _ in range(2):
            if st["pairs"]:
                own, ptrs = pg.alloc(8 * max(1, self.nl) * self.nz * self.nz)
                st["cla"].append(own)
                st["cla_ptrs"].append(ptrs)
            own, ptrs = pg.alloc(16 * nalm * max(1, self.cb))
            st["panel"].append(own)
            st["panel_ptrs"].append(ptrs)
        st["l_owner"] = _dev.to_device(self.plan.owner.astype(np.int32), t.int32)
        lrow = np.empty(L, dtype=np.int32)
        for ll in self.plan.l_lists:
            lrow[ll] = np.arange(len(ll), dtype=np.int32)
        st["l_row"] = _dev.to_device(lrow, t.int32)
        width = np.empty(self.nz, dtype=np.int32)
        for s_ in range(self.size):
            width[int(self.plan.chan_lo[s_]):int(self.plan.chan_hi[s_])] = int(self.plan.cb[s_])
        st["nu_width"] = _dev.to_device(width, t.int32)
        ntile = int(_lib.load().cora_b200_cl_fill_21cm_ntiles(self.nz)) if st["pairs"] else 0
        st["pair0"], st["npairs"], _ = fill_tile_share(ntile, self.rank, self.size)
        st["tables"] = [None, None]
        self._p2p = st
        return st

    def _p2p_tables(self, k):
        """Device pointer tables of buffer set k (resolved on first use: with virtual ranks the
        other ranks' buffers exist only after every rank ran ``_p2p_setup``)."""
        from . import peer as _peer

        t = _dev.torch()
        st = self._p2p_setup()
        if st["tables"][k] is None:
            tab = {}
            if st["pairs"]:
                tab["cla_ptrs"] = _dev.to_device(np.array(_peer.resolve(st["cla_ptrs"][k]), dtype=np.uint64).view(np.int64), t.int64)
            pp = _peer.resolve(st["panel_ptrs"][k])
            nu_ptr = np.empty(self.nz, dtype=np.uint64)
            for s_ in range(self.size):
                lo, hi = int(self.plan.chan_lo[s_]), int(self.plan.chan_hi[s_])
                nu_ptr[lo:hi] = np.uint64(pp[s_]) + np.uint64(16) * np.arange(hi - lo, dtype=np.uint64)
            tab["nu_ptr"] = _dev.to_device(nu_ptr.view(np.int64), t.int64)
            st["tables"][k] = tab
        return st["tables"][k]

    def p2p_fill(self, k):
        """Phase 1: C_l rows of the l's this rank owns land in ``cla[k]``.  Models whose fill cost
        is per channel pair (21cm) are sharded over pairs and scatter rows to the owners of l."""
        t = _dev.torch()
        st = self._p2p_setup()
        if st["pairs"]:
            tab = self._p2p_tables(k)
            self.model._b200_fill_tiles(self.fill_inputs, self.lmax + 1, self.nz, self.zint, st["pair0"], st["npairs"],
                                        tab["cla_ptrs"], st["l_owner"], st["l_row"], tile_step=self.size)
            return True     # remote stores in flight: barrier before the root reads cla[k]
        self.fill()
        return False

    def p2p_alm(self, k, seed=0, gauss=None, roots=None, cla=None):
        """Phase 2: root of the local l's, draws, apply with the a_lm stored straight into the
        PANEL buffers ``panel[k]`` of the GPUs owning each channel."""
        t = _dev.torch()
        lib = _lib.load()
        st = self._p2p_setup()
        tab = self._p2p_tables(k)
        if cla is None:
            if st["pairs"]:
                cla = self._persistent("cla_view%d" % k, lambda: st["cla"][k].tensor((self.nl, self.nz, self.nz), t.float64))
            else:
                cla = self._buf["cla"]
        if roots is None:
            outb = self._persistent("root", lambda: (_dev.empty((self.nl, self.nz, self.nz), t.float64),
                                                     _dev.empty((self.nl,), t.int32), _dev.empty((self.nl,), t.int32)))
            rws = self._persistent("root_ws", lambda: nputil.root_workspace(self.nl, self.nz, max_eigh=self._max_eigh()))
            root, used, _ = nputil.root_batched_device(cla, jitter_rel=1e-14, clip_rel=1e-16, out=outb, ws=rws)
        else:
            root, used = _dev.to_device(roots, t.float64), None
        lmax_loc = int(self.l_list.max())
        gptr, gld = self._draw_take(seed) if gauss is None else (None, 0)
        if gauss is None and gptr is None:
            def mk():
                return _dev.workspace(draw_workspace_bytes(self.nz, lmax_loc, self.nl))

            ws = self._persistent("draw_ws", mk)
        elif gauss is None:
            ws = self._persistent("desc_ws", lambda: _dev.workspace(64 * self.nl + 4096))
        else:
            gauss = _dev.to_device(gauss, t.complex128)
            ws = _dev.workspace(64 * self.nl + 4096)
            gptr, gld = _lib.ptr(gauss), int(gauss.shape[-1])
        _lib.call("cora_b200_draw_apply_peers", _lib.ptr(root), _lib.ptr(self.l_list), _lib.ptr(used), self.nl, self.nz,
                  self.lmax, ctypes.c_ulonglong(int(seed)), 0, gptr, gld, _lib.ptr(tab["nu_ptr"]), _lib.ptr(st["nu_width"]),
                  _lib.ptr(ws), int(ws.numel()), _lib.stream_ptr())

    def p2p_sht(self, k, out=None):
        """Phase 3: inverse SHT of this rank's channels from ``panel[k]``."""
        t = _dev.torch()
        st = self._p2p_setup()
        L = self.lmax + 1
        panel = self._persistent("panel_view%d" % k, lambda: st["panel"][k].tensor((L * (L + 1) // 2, self.cb), t.complex128))
        plan = _dev.sht_plan(self.nside, self.lmax)
        ws = self._persistent("sht_ws", lambda: _dev.sht_workspace(plan, _lib.ALM_PANEL, self.cb, reserve=(4 << 30) + 8 * self.cb * self.npix)[0])
        return hputil.alm2map_device(panel, self.nside, self.lmax, _lib.ALM_PANEL, self.cb, self.cb, out=out, ws=ws)

    def _max_eigh(self):
        """Eigen-fallback slots of the root workspace: all local l's when that is cheap (then the
        root stage never waits on the host), else an eighth of them."""
        full = 16 * self.nz * self.nz * self.nl
        return self.nl if full <= _dev.free_bytes() // 8 else max(4, self.nl // 8)

    def _step_p2p(self, seed=0, out=None):
        k = self._k & 1
        self._k += 1
        self._draw_begin(seed)
        if self.p2p_fill(k):
            self.peers.barrier()
        self.p2p_alm(k, seed=seed)
        self.peers.barrier()
        return self.p2p_sht(k, out=out)

    def getsky(self, seed=None, nbatch=None):
        """The public end-to-end call of the sharded generator -- ``Sky3d.getsky`` (``cora/core/maps.py:227-235``)
        for one rank of a multi-GPU run: host frequency axis in, this rank's channels out as a numpy
        ``float64[cb, npix]`` array.  Every call uploads the per-sample vectors again (what ``clarray`` does per call),
        runs fill -> root -> draw/apply -> exchange, and transforms the channels in ``nbatch`` batches whose device ->
        host copies overlap the next batch.  The result lives in a pinned buffer owned by this object and is
        overwritten by the next call.  No ``torch.distributed`` collective is issued on the p2p path."""
        from . import skysim

        t = _dev.torch()
        if seed is None:
            seed = int(np.random.randint(0, 2**31 - 1))
        za, _ = skysim._sample_frequencies(self.freq, self.zromb, None)
        self.fill_inputs = self.model._b200_fill_inputs(za, skysim.romberg_weights(self.zromb))
        L = self.lmax + 1
        nalm = L * (L + 1) // 2
        self._draw_begin(seed)
        if self.exchange == "p2p":
            k = self._k & 1
            self._k += 1
            if self.p2p_fill(k):
                self.peers.barrier()
            self.p2p_alm(k, seed=seed)
            self.peers.barrier()
            st = self._p2p_setup()
            panel = self._persistent("panel_view%d" % k, lambda: st["panel"][k].tensor((nalm, self.cb), t.complex128))
        else:
            cla = self.fill()
            send = self.alm_local(cla, seed=seed)
            if self.size == 1:
                panel = send
            else:
                rbuf = self._persistent("recv", lambda: _dev.empty((sum(self.plan.recv_splits(self.rank)),), t.complex128))
                recv = exchange(send, self.plan, self.rank, self.group, out=rbuf)
                panel = self._persistent("panel", lambda: _dev.empty((nalm, self.cb), t.complex128))
                _lib.call("cora_b200_alm_slabs_to_panel", _lib.ptr(recv), _lib.ptr(self.l_off), self.lmax, self.cb,
                          _lib.ptr(panel), self.cb, 0, _lib.stream_ptr())
        host = self._persistent("host_maps", lambda: t.empty((self.cb, self.npix), dtype=t.float64, pin_memory=True))
        sky = hputil.alm2map_to_host(panel, self.nside, self.lmax, self.cb, nbatch=nbatch, host=host, cache=self._buf)
        mean = np.asarray(self.model.mean_nu(self.freq), dtype=np.float64) if hasattr(self.model, "mean_nu") else None
        if mean is not None and mean.any():
            lo = int(self.plan.chan_lo[self.rank])
            sky += mean[lo:lo + self.cb, None]          # the reference's host-side broadcast add (maps.py:235)
        return sky

    def close(self):
        """Check the barrier status and release the peer buffers (collective when a real PeerGroup is used)."""
        if self.exchange == "p2p" and self.peers is not None:
            self.peers.check()
            self._p2p = None
            self._buf = {k: v for k, v in self._buf.items() if not k.startswith(("cla_view", "panel_view"))}
            self.peers.close()
            self.peers = None

    def step(self, seed=0, out=None):
        if self.exchange == "p2p":
            return self._step_p2p(seed=seed, out=out)
        self._draw_begin(seed)
        cla = self.fill()
        send = self.alm_local(cla, seed=seed)
        del cla
        rbuf = None
        if self.size > 1:
            rbuf = self._persistent("recv", lambda: _dev.empty((sum(self.plan.recv_splits(self.rank)),), _dev.torch().complex128))
        recv = exchange(send, self.plan, self.rank, self.group, out=rbuf)
        return self.synthesize(recv, out=out)


def single_gpu_block(model, nside, frequencies, lmax, zromb, seed, chan_lo, chan_hi):
    """The single-GPU path (all l, all channels for fill / root / apply on this device) with the inverse SHT of
    channels [chan_lo, chan_hi) only: what a rank of a sharded run must reproduce for its channel block (Philox
    counters are keyed by (l, m, nu), so the maps do not depend on the number of ranks).  CUDA ``float64[n, npix]``.
    Buffers are released stage by stage and the workspaces are kept small (peak ~3 C_l tables), so that the check
    fits beside a sharded run's own buffers."""
    from . import skysim

    t = _dev.torch()
    lib = _lib.load()
    freq = np.asarray(frequencies, dtype=np.float64)
    nz, L = len(freq), int(lmax) + 1
    nalm = L * (L + 1) // 2
    npix = 12 * int(nside) ** 2
    za, zint = skysim._sample_frequencies(freq, zromb, None)
    inputs = model._b200_fill_inputs(za, skysim.romberg_weights(zromb))
    cla = _dev.zeros((L, nz, nz), t.float64)
    model._b200_fill(inputs, 0, 1, L, nz, zint, cla, lower_only=True)
    rws = nputil.root_workspace(L, nz, max_eigh=max(4, L // 32))
    root, used, _ = nputil.root_batched_device(cla, jitter_rel=1e-14, clip_rel=1e-16, ws=rws)
    del cla, rws
    panel = _dev.empty((nalm, nz), t.complex128)
    l_list = np.arange(L, dtype=np.int32)
    one = lib.cora_b200_draw_apply_workspace_bytes(nz, int(lmax), 1)
    ws = _dev.workspace(max(one, min(lib.cora_b200_draw_apply_workspace_bytes(nz, int(lmax), L), 4 << 30)))
    _lib.call("cora_b200_draw_apply", _lib.ptr(root), _lib.ptr(l_list), _lib.ptr(used), L, nz, int(lmax),
              ctypes.c_ulonglong(int(seed)), None, 0, _lib.ptr(panel), nz, 0, 0, nz, _lib.ptr(ws), int(ws.numel()),
              _lib.stream_ptr())
    t.cuda.current_stream().synchronize()
    del root, used, ws
    n = int(chan_hi) - int(chan_lo)
    out = _dev.empty((n, npix), t.float64)
    plan = _dev.sht_plan(int(nside), int(lmax))
    nbytes = int(lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, min(n, 64)))
    ws = _dev.workspace(nbytes)
    _lib.call("cora_b200_alm2map", plan, _lib.ptr_off(panel, 16 * int(chan_lo)), _lib.ALM_PANEL, nz, n, _lib.ptr(out),
              _lib.ptr(ws), nbytes, _lib.stream_ptr())
    t.cuda.current_stream().synchronize()
    return out


def mkfullsky_sharded(corr_local, nside, l_list=None, *, lmax, group=None, partition="interleaved", seed=0, gauss=None,
                      roots=None, device_out=True, alms=False):
    """``mkfullsky`` for an l-distributed ``corr`` (the MPIArray branch of ``skysim.py:97-134``).

    ``corr_local``: this rank's ``float64[nl_local, nz, nz]`` rows, for the l's of
    ``ShardPlan(lmax, nz, G, partition).l_lists[rank]``.  Returns this rank's channels
    ``float64[cb, npix]`` (the reference returns the frequency-distributed ``MPIArray``), or with ``alms`` the
    all-gathered ``complex128[nz, 1, L, L]`` alm array (``skysim.py:123-125``: the same on every rank)."""
    t = _dev.torch()
    dist = _dist()
    size = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if size > 1 else 0
    nz = corr_local.shape[1]
    if corr_local.shape[2] != nz:
        raise Exception("Correlation matrix is incorrect shape.")

    class _NoModel(object):
        def _b200_fill_inputs(self, za, w):
            return None

    sh = ShardedSky(_NoModel(), nside, np.zeros(nz), lmax=lmax, zromb=0, group=group, partition=partition, rank=rank, size=size,
                    exchange="collective" if alms else "auto")     # the all-gather needs no peer buffers
    if l_list is not None and not np.array_equal(np.asarray(l_list), sh.l_list):
        raise Exception("l_list does not match the %s partition of rank %d" % (partition, rank))
    if corr_local.shape[0] != sh.nl:
        raise Exception("Correlation matrix is incorrect shape.")
    cla = _dev.to_device(corr_local, t.float64)
    if alms:
        out = sh.alm_gathered(cla, seed=seed, gauss=gauss, roots=roots)
        return out if device_out else _dev.to_host(out)
    if sh.exchange == "p2p":
        sh.p2p_alm(0, seed=seed, gauss=gauss, roots=roots, cla=cla)
        sh.peers.barrier()
        sky = sh.p2p_sht(0).clone()
        sh.peers.check()
        sh.peers.close()
    else:
        send = sh.alm_local(cla, seed=seed, gauss=gauss, roots=roots)
        recv = exchange(send, sh.plan, rank, group)
        sky = sh.synthesize(recv)
    return sky if device_out else _dev.to_host(sky)


def mkfullsky_mpi(corr, nside, alms=False, rng=None, seed=None, group=None):
    """``skysim.mkfullsky`` for a distributed ``corr`` (``cora/core/skysim.py:97-134``; the caller is
    ``cora/signal/lss.py:450``): ``corr`` exposes ``local_array`` (this rank's ``float64[nl_local, nz, nz]`` block of
    l's, caput's contiguous split), ``global_shape`` and -- for this package's ``mpiarray.MPIArray`` -- ``comm``, the
    ``torch.distributed`` group.  Returns what the reference returns: the maps as an array distributed over
    frequency (``wrap(sky, axis=0)``), or with ``alms`` the all-gathered ``complex128[nz, 1, L, L]`` numpy array.

    ``rng``: each rank draws ``complex_std_normal((nz, l + 1), rng)`` for its local l's in ascending order, exactly
    as ``skysim.py:114-121`` does (the pipeline caller seeds every rank identically, SURVEY App. C.5: kept, because
    it is the identical-draw parity path).  Without ``rng`` the draws are the device Philox stream keyed by
    (l, m, nu): independent of the number of ranks."""
    from . import mpiarray

    if group is None:
        group = getattr(corr, "comm", None)
    dist = _dist()
    if group is not None and not (dist.is_available() and isinstance(group, dist.ProcessGroup)):
        group = None         # an mpi4py communicator cannot drive torch.distributed: use the default group
    size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if size > 1 else 0
    L = int(corr.global_shape[0])
    loc = np.asarray(corr.local_array, dtype=np.float64)
    nz = loc.shape[1]
    if loc.ndim != 3 or loc.shape[2] != nz:
        raise Exception("Correlation matrix is incorrect shape.")
    lo, hi = block_partition(L, size, rank)
    if loc.shape[0] != hi - lo:
        raise Exception("Correlation matrix is incorrect shape.")
    gauss = None
    if rng is not None:
        gauss = np.zeros((hi - lo, nz, L), dtype=np.complex128)
        for i, l in enumerate(range(lo, hi)):
            gauss[i, :, : l + 1] = nputil.complex_std_normal((nz, l + 1), rng=rng)
    elif seed is None:
        seed = int(np.random.randint(0, 2**31 - 1))
    out = mkfullsky_sharded(loc, nside, lmax=L - 1, group=group, partition="block", seed=seed or 0, gauss=gauss,
                            device_out=False, alms=alms)
    if alms:
        return out
    wrap = getattr(type(corr), "wrap", None)
    if wrap is not None and not isinstance(corr, mpiarray.MPIArray):
        return wrap(out, axis=0)                 # caput's MPIArray.wrap(array, axis)
    return mpiarray.MPIArray.wrap(out, axis=0, comm=group)


class ShardedPolSky(object):
    """``cora-makesky gaussianfg --pol full`` sharded over GPUs ("shard (pol-block, l)
    for roots, channels for SHT").

    The reference builds the dense block-diagonal covariance ``blockdiag(T, E, B, V)`` of size
    ``(4 nfreq)^2`` per l (``cora/scripts/makesky.py:368-382``; 206 GB at nside 512 x 1024
    channels) and roots it whole.  Here the blocks are kept apart: T = ``FullSkySynchrotron``,
    E = B = ``FullSkyPolarisedSynchrotron`` (one root, applied to two independent draw streams),
    V = 0.  The regulariser keeps the reference's whole-matrix semantics (``cora_b200_root_batched_multi``):
    the jitter ``1e-14 max(diag)`` is taken over the whole matrix' diagonal, a Cholesky failure of any block
    sends every block of that l to the eigen branch, and the eigenvalue clip is ``1e-16`` times the largest
    eigenvalue over all blocks -- at 1024 channels this is what leaves 8 + 52 + 52 of 4096 modes (SURVEY
    App. C.6).  The Philox counters of block b are offset by ``b nfreq`` -- the same draws the dense
    formulation consumes -- and the output layout is ``float64[freq, 4, pix]``.  Stokes V: the reference's V
    block is ``sqrt(cmax) I`` on the Cholesky branch (a map ``<= 1e-7`` of T) and 0 on the eigen branch;
    ``self.vscale`` holds that factor per l, the V map itself is written as 0 (documented deviation at the
    level of the regulariser on the few-channel Cholesky branch only).

    Exchange: apply stores straight into the PANEL buffers (T, E, B) of the GPU owning each
    channel (``cora_b200.peer``); one flag barrier per step.  With one rank the same code runs on
    local pointers."""

    def __init__(self, nside, frequencies, lmax=None, zromb=3, group=None, partition="interleaved", rank=None, size=None,
                 peers=None, models=None):
        from . import galaxy, skysim
        from . import peer as _peer

        t = _dev.torch()
        dist = _dist()
        if size is None:
            size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if size > 1 else 0
        self.rank, self.size, self.group = int(rank), int(size), group
        self.nside = int(nside)
        self.freq = np.asarray(frequencies, dtype=np.float64)
        self.nz = len(self.freq)
        self.lmax = 3 * self.nside if lmax is None else int(lmax)   # makesky.py:367
        self.plan = ShardPlan(self.lmax, self.nz, self.size, partition)
        self.l_list = self.plan.l_lists[self.rank]
        self.nl = len(self.l_list)
        self.cb = int(self.plan.cb[self.rank])
        self.npix = 12 * self.nside**2
        self.models = models if models is not None else (galaxy.FullSkySynchrotron(), galaxy.FullSkyPolarisedSynchrotron())
        za, self.zint = skysim._sample_frequencies(self.freq, zromb, None)
        w = skysim.romberg_weights(zromb)
        self.fill_inputs = [m._b200_fill_inputs(za, w) for m in self.models]
        if peers is None:
            peers = _peer.PeerGroup(self.rank, self.size, group) if self.size > 1 else _peer.LocalPeers(1).view(0)
        self.peers = peers
        L = self.lmax + 1
        self.nalm = L * (L + 1) // 2
        # PANEL buffers: [buffer set k][field T, E, B]
        self.panel, self.panel_ptrs = [], []
        for _ in range(2):
            own, ptrs = zip(*[peers.alloc(16 * self.nalm * max(1, self.cb)) for _ in range(3)])
            self.panel.append(own)
            self.panel_ptrs.append(ptrs)
        width = np.empty(self.nz, dtype=np.int32)
        for s_ in range(self.size):
            width[int(self.plan.chan_lo[s_]):int(self.plan.chan_hi[s_])] = int(self.plan.cb[s_])
        self.nu_width = _dev.to_device(width, t.int32)
        self._tables = [None, None]
        self._buf = {}
        self._k = 0

    def _persistent(self, name, make):
        if name not in self._buf:
            self._buf[name] = make()
        return self._buf[name]

    def _nu_ptr(self, k):
        from . import peer as _peer

        t = _dev.torch()
        if self._tables[k] is None:
            tabs = []
            for f in range(3):
                pp = _peer.resolve(self.panel_ptrs[k][f])
                nu_ptr = np.empty(self.nz, dtype=np.uint64)
                for s_ in range(self.size):
                    lo, hi = int(self.plan.chan_lo[s_]), int(self.plan.chan_hi[s_])
                    nu_ptr[lo:hi] = np.uint64(pp[s_]) + np.uint64(16) * np.arange(hi - lo, dtype=np.uint64)
                tabs.append(_dev.to_device(nu_ptr.view(np.int64), t.int64))
            self._tables[k] = tabs
        return self._tables[k]

    def alm_phase(self, k, seed=0):
        """fill (T, P) -> global diagonal maxima -> block roots -> draw + apply (T, E, B) with the
        a_lm stored into the owners' PANEL buffers of set k."""
        t = _dev.torch()
        lib = _lib.load()
        nl, nz = self.nl, self.nz
        if nl == 0:
            return
        l0 = int(self.l_list[0])
        step = self.size if self.plan.partition == "interleaved" else 1
        cl = self._persistent("cl", lambda: [_dev.empty((nl, nz, nz), t.float64) for _ in range(2)])
        roots = self._persistent("root", lambda: ([_dev.empty((nl, nz, nz), t.float64) for _ in range(2)],
                                                  _dev.empty((2, nl), t.int32), _dev.empty((2, nl), t.int32)))
        self.vscale = self._persistent("vscale", lambda: _dev.empty((nl,), t.float64))
        for b in range(2):
            self.models[b]._b200_fill(self.fill_inputs[b], l0, step, nl, nz, self.zint, cl[b])
        rws = self._persistent("root_ws", lambda: nputil.root_multi_workspace(
            2, nl, nz, max_eigh=nl if 32 * nz * nz * nl <= _dev.free_bytes() // 8 else max(4, nl // 8)))
        # one jitter, one Cholesky-or-eigh decision and one clip threshold per l over blockdiag(T, E, B, V)
        nputil.root_batched_multi_device(cl, 1e-14, 1e-16, out=roots, ws=rws, zero_scale=self.vscale)
        rlist, used_all, _ = roots
        lmax_loc = int(self.l_list.max())

        def mk():
            return _dev.workspace(draw_workspace_bytes(nz, lmax_loc, nl))

        ws = self._persistent("draw_ws", mk)
        tabs = self._nu_ptr(k)
        for f, b in ((0, 0), (1, 1), (2, 1)):      # field -> block: T <- T root, E and B <- the polarised root
            root, used = rlist[b], used_all[b]
            _lib.call("cora_b200_draw_apply_peers", _lib.ptr(root), _lib.ptr(self.l_list), _lib.ptr(used), nl, nz, self.lmax,
                      ctypes.c_ulonglong(int(seed)), f * nz, None, 0, _lib.ptr(tabs[f]), _lib.ptr(self.nu_width),
                      _lib.ptr(ws), int(ws.numel()), _lib.stream_ptr())

    def sht_phase(self, k, out=None):
        """T scalar, (E, B) -> (Q, U) spin-2, V = 0 -> ``float64[cb, 4, npix]``."""
        t = _dev.torch()
        lib = _lib.load()
        cb, npix = self.cb, self.npix
        if out is None:
            out = self._persistent("out", lambda: _dev.zeros((cb, 4, npix), t.float64))
        if cb == 0:
            return out
        plan = _dev.sht_plan(self.nside, self.lmax)

        def mkws():
            need = lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, cb) * 2
            per = lib.cora_b200_alm2map_workspace_bytes(plan, _lib.ALM_PANEL, 16) * 2
            return _dev.workspace(min(need, max(per, _dev.free_bytes() - (4 << 30))))

        ws = self._persistent("sht_ws", mkws)
        pT, pE, pB = (self.panel[k][f] for f in range(3))
        base = out.data_ptr()
        _lib.call("cora_b200_alm2map_strided", plan, _lib.ptr(pT), _lib.ALM_PANEL, cb, cb, ctypes.c_void_p(base), 4 * npix,
                  _lib.ptr(ws), int(ws.numel()), _lib.stream_ptr())
        _lib.call("cora_b200_alm2map_spin2_strided", plan, _lib.ptr(pE), _lib.ptr(pB), _lib.ALM_PANEL, cb, cb,
                  ctypes.c_void_p(base + 8 * npix), ctypes.c_void_p(base + 16 * npix), 4 * npix, _lib.ptr(ws), int(ws.numel()),
                  _lib.stream_ptr())
        return out

    def step(self, seed=0, out=None):
        k = self._k & 1
        self._k += 1
        self.alm_phase(k, seed=seed)
        self.peers.barrier()
        return self.sht_phase(k, out=out)
