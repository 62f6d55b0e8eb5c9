"""Slab-decomposed multi-GPU density-field -> power-spectrum pipeline (one process per GPU).

The reference is single-process ("slab-transpose comm layer: does not exist");
this module is the new component BASELINE.json's north_star asks for.  Per snapshot, on rank r of G:

    particles (any shard)                       pos_r (n_r, 3)
      -> EITHER deposit onto a FULL partial grid, (N,N,N) float32                      [local]
                reduce-scatter (sum) into x-slabs (N/G, N, N)                          [NCCL reduce_scatter, 4 N^3 B]
         OR     route particles to the rank owning their lowest x-plane                [partition kernel + grouped NCCL send/recv, 16 B/particle,
                deposit onto slab + S-1 halo planes, pass the halo to the next rank     in pieces overlapped with the windowed deposit; halo: send/recv]
      -> overdensity with the GLOBAL mean       sum in float64, all-reduced            [NCCL all_reduce, 8 B]
      -> batched 2-D R2C over (y,z)             (N/G, N, N/2+1) complex64              [cuFFT]
      -> transpose pack + all-to-all            (G, N/G, N/G, N/2+1) blocks            [pack kernel + NCCL all_to_all]
      -> 1-D C2C along x (strided, in place)    (N, N/G, N/2+1): rank r owns ky in [r N/G, (r+1) N/G)   [cuFFT]
      -> fused deconvolve + bin + Legendre      on the transposed layout, no transpose back   [ring kernel]
      -> all-reduce of the k-bins               float64 sums + int64 counts (<= 24 MB) [NCCL all_reduce]
      -> units / normalisation on the host      identical on every rank

The local operations come from an `ops` object so the host logic (partitioning, collectives, layout
arithmetic) can be exercised on CPU with the gloo backend and a numpy stand-in (tests/cpu_slab_ops.py);
the product always uses CudaOps -- there is no CPU fallback in this package.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import Pk_library as PKL
from . import MAS_library as MASL


class CudaOps(object):
    """Local (per-rank) operations on the current CUDA device, through the C ABI."""

    def __init__(self):
        self.lib = _lib.load()
        self.dev = MASL._device()

    # ---- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.dev)

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def to_device(self, a):
        return MASL._to_device(a, self.dev)

    # ---- stages --------------------------------------------------------------------------------
    def deposit(self, pos, W, grid, BoxSize, MAS):
        MASL.MA(pos, grid, BoxSize, MAS, W=W)

    def partition(self, pos, W, BoxSize, MAS, G, dims):
        """Group this rank's particles by owning x-slab: (float32 (n,4) x,y,z,w in slab order, int32 offsets[G+1])."""
        d_pos = MASL._to_device(pos, self.dev)
        d_w = MASL._to_device(W, self.dev) if W is not None else None
        n = d_pos.shape[0]
        out = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        offsets = torch.empty(G + 1, dtype=torch.int32, device=self.dev)
        s0, s1 = d_pos.stride()
        _lib.check(self.lib.pylb_partition_xslab(d_pos.data_ptr(), n, s0, s1, d_w.data_ptr() if d_w is not None else None,
                                                 d_w.stride(0) if d_w is not None else 1, dims, float(BoxSize),
                                                 MASL._MAS_ID[MAS], G, out.data_ptr(), offsets.data_ptr(), self._stream()),
                   "pylb_partition_xslab")
        return out, offsets

    def deposit_window(self, xyzw, grid, x0, BoxSize, MAS, weighted, dims):
        """Deposit packed (x,y,z,w) particles onto `grid` = planes x0 .. x0+grid.shape[0]-1 (mod dims)."""
        n, xext = xyzw.shape[0], grid.shape[0]
        mas = MASL._MAS_ID[MAS]
        wb = self.lib.pylb_ma_window_workspace_bytes(n, dims, xext, mas, 0)
        ws = torch.empty(max(int(wb), 1), dtype=torch.uint8, device=self.dev)
        wptr = xyzw.data_ptr() + 12 if weighted else None
        _lib.check(self.lib.pylb_ma_window(xyzw.data_ptr(), n, 4, 1, grid.data_ptr(), dims, int(x0), int(xext),
                                           float(BoxSize), mas, wptr, 4, 0, ws.data_ptr(), int(wb), self._stream()),
                   "pylb_ma_window")

    def add(self, dst, src):
        _lib.check(self.lib.pylb_add_f32(dst.data_ptr(), src.data_ptr(), dst.numel(), self._stream()), "pylb_add_f32")

    def grid_sum(self, slab):
        s = torch.zeros(1, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.pylb_grid_sum(slab.data_ptr(), slab.numel(), s.data_ptr(), self._stream()), "pylb_grid_sum")
        return s

    def overdensity_apply(self, slab, total, n_total):
        _lib.check(self.lib.pylb_overdensity_apply(slab.data_ptr(), slab.numel(), total.data_ptr(), int(n_total),
                                                   self._stream()), "pylb_overdensity_apply")

    def kpitch(self, dims):
        """complex elements per kz-row of this engine's k-space buffers: even, so that every row is 16-byte aligned"""
        nz = dims // 2 + 1
        return nz + (nz & 1)

    def fft_yz(self, slab, dims):
        nxl, pitch = slab.shape[0], self.kpitch(dims)
        out = torch.empty((nxl, dims, pitch), dtype=torch.complex64, device=self.dev)
        wb = self.lib.pylb_fft_slab_yz_work_bytes(dims, nxl, pitch)
        work = torch.empty(max(int(wb), 1), dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.pylb_fft_slab_yz(slab.data_ptr(), out.data_ptr(), dims, nxl, pitch, work.data_ptr(), int(wb),
                                             self._stream()), "pylb_fft_slab_yz")
        return out

    def pack(self, cplx, dims, G):
        nxl, pitch = cplx.shape[0], cplx.shape[2]
        send = torch.empty((G, nxl, dims // G, pitch), dtype=torch.complex64, device=self.dev)
        _lib.check(self.lib.pylb_slab_pack(cplx.data_ptr(), send.data_ptr(), dims, nxl, G, pitch, self._stream()), "pylb_slab_pack")
        return send

    def fft_x(self, recv, dims):
        nyl, pitch = recv.shape[1], recv.shape[2]
        wb = self.lib.pylb_fft_slab_x_work_bytes(dims, nyl, pitch)
        work = torch.empty(max(int(wb), 1), dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.pylb_fft_slab_x(recv.data_ptr(), dims, nyl, pitch, work.data_ptr(), int(wb), self._stream()),
                   "pylb_fft_slab_x")

    def bin(self, fields, dims, axis, mas_index, want_phase, y0, nyl):
        pitch = fields[0].shape[2]
        ks = _lib.KSpace(dims, 0, dims, int(y0), int(nyl), nyl * pitch, pitch)
        if len(fields) > 3 and int(axis) == 2 and not want_phase:
            return PKL.bin_modes_by_subsets(fields, dims, 2, mas_index, ks=ks)      # three fields at a time (ring kernel)
        return PKL.bin_modes(fields, dims, axis, mas_index, want_phase, False, ks=ks)

    def overdensity_mean(self, slab, mean):
        """slab /= mean; slab -= 1 with the caller's mean (Pk_snapshot.py:84-88)."""
        _lib.check(self.lib.pylb_overdensity_mean(slab.data_ptr(), slab.numel(), float(np.float32(mean)), self._stream()),
                   "pylb_overdensity_mean")

    def load_species(self, snapshot_fname, file_slice, ptype, do_RSD, axis):
        """Positions (Mpc/h, redshift space if asked) of species `ptype` from the sub-files `file_slice` selects, as ONE
        device tensor: every block goes disk -> pinned memory -> its place in the tensor (MAS_gadget.StreamedSnapshot)."""
        from .MAS_gadget import StreamedSnapshot
        snap = StreamedSnapshot(snapshot_fname)
        snap.files = snap.files[file_slice]
        BoxSize = snap.head.boxsize / 1e3
        out = torch.empty((snap.count(ptype), 3), dtype=torch.float32, device=self.dev)
        row = 0
        for pos, vel, _, m in snap.blocks(ptype, want_vel=do_RSD):
            if do_RSD:
                snap.to_redshift_space(pos, vel, BoxSize, axis)
            out[row:row + m].copy_(pos)
            row += m
        return out


class ParticleBatches(object):
    """A rank's particles as a sequence of (pos, W) batches produced on demand (a snapshot read file by file, a synthetic
    set generated in pieces): what `SlabPk.density_slab` takes when the shard does not fit in HBM next to the grid.
    make(i) -> (pos (n_i, 3) float32, W (n_i,) float32 or None); total = particles of this rank over all batches."""

    def __init__(self, make, nbatches, total):
        self.make, self.nbatches, self.total = make, int(nbatches), int(total)

    def __iter__(self):
        for i in range(self.nbatches):
            yield self.make(i)


def _group_info(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _reduce_scatter_sum(out, full, group, G):
    """out (N/G,...) <- sum over ranks of full (N,...)[rank slab].  NCCL has the primitive; gloo does not."""
    if G == 1:
        out.copy_(full[: out.shape[0]])
        return
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        r = dist.get_rank(group)
        out.copy_(full[r * out.shape[0]:(r + 1) * out.shape[0]])


class _Result(object):
    pass


class SlabPk(object):
    """Distributed MA + Pk / XPk.  Every rank calls the same methods with its own particle shard."""

    def __init__(self, dims, BoxSize, MAS="CIC", axis=2, group=None, ops=None, exchange="auto", exchange_chunks=2,
                 fft_chunks=4):
        """exchange: how per-rank deposits become x-slabs --
             "grid"      every rank deposits onto a full partial grid, then reduce-scatter (4 N^3 bytes per rank);
             "particles" particles are routed to the rank owning their lowest touched x-plane (16 B per particle),
                         deposited onto slab + S-1 halo planes, halo planes are passed to the next rank;
             "auto"      whichever moves fewer bytes for this call.
           exchange_chunks: pieces the routed particles travel in ("particles" mode), so that the deposit of one piece
             overlaps the transfer of the next; must be the same on every rank."""
        self.exchange = exchange
        self.exchange_chunks = max(1, int(exchange_chunks))
        self.fft_chunks = max(1, int(fft_chunks))         # x-plane groups of the slab FFT (transform of one overlaps the transpose of the previous)
        self._auto_mode = {}
        self.rank, self.G = _group_info(group)
        self.group = group
        if dims % self.G != 0:
            raise ValueError("dims (%d) must be divisible by the number of ranks (%d)" % (dims, self.G))
        self.dims, self.BoxSize, self.MAS, self.axis = int(dims), BoxSize, MAS, int(axis)
        self.nxl = self.nyl = dims // self.G
        self.ops = ops if ops is not None else CudaOps()

    # ---- stage 1: particles -> overdensity slab --------------------------------------------------
    def density_slab(self, pos, W=None, MAS=None, overdensity=True):
        """Deposit this rank's particles and return the x-slab this rank owns (optionally as overdensity).
        `pos` is an (n, 3) array / tensor, or a ParticleBatches whose batches are deposited one after the other."""
        ops, N, G = self.ops, self.dims, self.G
        MAS = MAS or self.MAS
        halo = {"NGP": 0, "CIC": 1, "TSC": 2, "PCS": 3}[MAS]
        batched = isinstance(pos, ParticleBatches)
        count = pos.total if batched else int(pos.shape[0])
        mode = self.exchange
        if mode == "auto":
            # The reduce-scatter of the partial grids moves 4 N^3 bytes per rank and nothing hides it; the routed particles
            # move 16 B each, but in pieces that cross NVLink while the previous piece is being deposited, and the deposit
            # then only covers (and flushes) a slab instead of the whole cube.  Measured with 1024^3 PCS particles per GPU
            # (profiles/r2_dist_stages_{2,4,8}gpu*.txt, ms per density_slab, grid / particles): G=2, 1280^3: 122.8 / 125.4;
            # G=4, 1600^3: 141.3 / 122.3; G=8, 2048^3: 184.7 / 131.9 -- i.e. particles win once 8 N^3 > 16 B x particles.
            # Every rank must take the same branch, so the particle count is agreed on (max over ranks); the choice is
            # kept per (scheme, particle-count magnitude), and a stencil wider than a rank's slab always takes "grid".
            key = (MAS, count.bit_length())
            if key not in self._auto_mode:
                npmax = count
                if G > 1:
                    t = torch.tensor([npmax], dtype=torch.int64, device=getattr(ops, "dev", torch.device("cpu")))
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
                    npmax = int(t[0].item())
                ok = G > 1 and self.nxl >= max(halo, 1)
                self._auto_mode[key] = "particles" if (ok and 8 * N ** 3 > 16 * npmax) else "grid"
            mode = self._auto_mode[key]
        if mode == "particles" and self.nxl < halo:
            raise ValueError("particle exchange needs at least %d planes per rank for %s" % (halo, MAS))
        batches = pos if batched else self._batches_of(pos, W)
        slab = self._slab_from_particles(batches, MAS, halo) if mode == "particles" else self._slab_from_grids(batches, MAS)
        if overdensity:
            total = ops.grid_sum(slab)
            if G > 1:
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
            ops.overdensity_apply(slab, total, N ** 3)
        return slab

    def _batches_of(self, pos, W):
        """One (pos, W) batch -- or, for a HOST array feeding a CUDA engine, a stream of device chunks: the H2D copy of
        chunk i+1 runs on a side stream while chunk i is partitioned, exchanged and deposited (double-buffered, pinned
        host memory makes the copies truly asynchronous).  Chunking changes nothing but the fp32 summation order."""
        dev = getattr(self.ops, "dev", None)
        on_host = not (isinstance(pos, torch.Tensor) and pos.is_cuda)
        if dev is None or not on_host:
            return [(pos, W)]
        chunk = MASL._host_chunk(self.dims, 3)
        n = int(pos.shape[0])
        if n <= chunk:
            return [(pos, W)]
        return self._stream_host_chunks(MASL._as_cpu_tensor(pos), None if W is None else MASL._as_cpu_tensor(W), n, chunk, dev)

    def _stream_host_chunks(self, h_pos, h_w, n, chunk, dev):
        cs = MASL._copy_stream(dev)
        cur = torch.cuda.current_stream(dev)
        bufs = [torch.empty((chunk, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        wbufs = [torch.empty(chunk, dtype=torch.float32, device=dev) for _ in range(2)] if h_w is not None else None
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        bounds = MASL._chunk_bounds(n, chunk)         # the last full chunk tapers: less work left after the last copy
        nchunks = len(bounds)
        cs.wait_stream(cur)

        def issue(i):
            b = i % 2
            lo, hi = bounds[i]
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(consumed[b])
                bufs[b][: hi - lo].copy_(h_pos[lo:hi], non_blocking=True)
                if wbufs is not None:
                    wbufs[b][: hi - lo].copy_(h_w[lo:hi], non_blocking=True)
                copied[b].record(cs)

        issue(0)
        for i in range(nchunks):
            b = i % 2
            m = bounds[i][1] - bounds[i][0]
            if i + 1 < nchunks:
                issue(i + 1)
            cur.wait_event(copied[b])
            yield bufs[b][:m], (wbufs[b][:m] if wbufs is not None else None)
            consumed[b].record(cur)              # runs when the consumer asks for the next chunk: chunk i is fully queued
        for t in bufs + (wbufs or []):
            t.record_stream(cur)

    def _slab_from_grids(self, batches, MAS):
        ops, N, G = self.ops, self.dims, self.G
        partial = ops.zeros((N, N, N))
        for pos, W in batches:
            ops.deposit(pos, W, partial, self.BoxSize, MAS)
        slab = ops.empty((self.nxl, N, N))
        _reduce_scatter_sum(slab, partial, self.group, G)
        return slab

    def _peer(self, g):
        return dist.get_global_rank(self.group, g) if self.group is not None else g

    def _slab_from_particles(self, batches, MAS, halo):
        """Route every particle to the rank owning its lowest touched x-plane and deposit there.

        The routed payload moves in `exchange_chunks` pieces: all pieces are queued on NCCL's stream at once (grouped
        send/recv = all-to-all), and the windowed deposit of piece c runs on the compute stream as soon as piece c has
        arrived, i.e. while piece c+1 is still crossing NVLink.  MA only ever adds into the grid, so depositing in
        pieces (and in batches) changes nothing but the fp32 summation order."""
        ops, N, G, r = self.ops, self.dims, self.G, self.rank
        if G == 1:
            grid = ops.zeros((self.nxl, N, N))                    # the window is the whole periodic cube
            for pos, W in batches:
                send, _ = ops.partition(pos, W, self.BoxSize, MAS, G, N)
                ops.deposit_window(send, grid, 0, self.BoxSize, MAS, W is not None, N)
            return grid
        grid = ops.zeros((self.nxl + halo, N, N))
        for pos, W in batches:
            self._route_and_deposit(pos, W, MAS, grid)
        if halo:
            mine = grid[self.nxl:]                                # planes that belong to the next rank
            got = torch.empty_like(mine)
            nxt, prv = self._peer((r + 1) % G), self._peer((r - 1) % G)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, nxt, group=self.group),
                                           dist.P2POp(dist.irecv, got, prv, group=self.group)])
            for q in reqs:
                q.wait()
            ops.add(grid[:halo], got)
        return grid[: self.nxl]

    def _route_and_deposit(self, pos, W, MAS, grid):
        """One batch: partition by owner, exchange in pieces, deposit every piece onto this rank's window `grid`."""
        ops, N, G, r = self.ops, self.dims, self.G, self.rank
        send, offsets = ops.partition(pos, W, self.BoxSize, MAS, G, N)
        # the split sizes must be known on the host: every rank's G counts are gathered on the device and read back with
        # ONE device-to-host copy (row s = what rank s sends to each destination)
        counts = (offsets[1:] - offsets[:-1]).to(torch.int64)
        table = torch.empty(G * G, dtype=torch.int64, device=counts.device)          # flat: gloo wants a 1-D output
        dist.all_gather_into_tensor(table, counts.contiguous(), group=self.group)
        table = table.view(G, G).to("cpu").tolist()
        send_tot = table[r]
        recv_tot = [table[s_][r] for s_ in range(G)]
        off = [0]
        for g in range(G):
            off.append(off[-1] + send_tot[g])
        K = self.exchange_chunks

        def part(c, total):                                       # rows [lo, hi) of a `total`-row range that travel in piece c
            return (c * total) // K, ((c + 1) * total) // K       # same formula on the sending and the receiving side

        pieces = []
        for c in range(K):
            spans = [part(c, recv_tot[s]) for s in range(G)]
            buf = send.new_empty((sum(hi - lo for lo, hi in spans), 4))
            p2p, row = [], 0
            for s in range(G):
                n = spans[s][1] - spans[s][0]
                if n:
                    if s == r:                                    # my own share stays on the device
                        lo, hi = part(c, send_tot[r])
                        buf[row:row + n].copy_(send[off[r] + lo:off[r] + hi])
                    else:
                        p2p.append(dist.P2POp(dist.irecv, buf[row:row + n], self._peer(s), group=self.group))
                row += n
            for g in range(G):
                lo, hi = part(c, send_tot[g])
                if hi > lo and g != r:
                    p2p.append(dist.P2POp(dist.isend, send[off[g] + lo:off[g] + hi], self._peer(g), group=self.group))
            pieces.append((buf, dist.batch_isend_irecv(p2p) if p2p else []))
        for buf, reqs in pieces:
            for q in reqs:
                q.wait()
            if buf.shape[0]:
                ops.deposit_window(buf, grid, r * self.nxl, self.BoxSize, MAS, W is not None, N)
        del pieces, send

    # ---- stage 2: x-slab (real) -> ky-slab (k-space, transposed) --------------------------------
    def fft_slab(self, slab):
        """(N/G, N, N) real x-slab -> (N, N/G, P) complex: all kx, this rank's ky, kz >= 0 (P = the row pitch of `ops`).

        The x-planes go through in `fft_chunks` groups: while the transposed blocks of one group cross NVLink (grouped
        send/recv straight into their place in the receive buffer -- a source's planes are contiguous there), the 2-D
        transforms and the pack of the next group run on the compute stream."""
        ops, N, G, r = self.ops, self.dims, self.G, self.rank
        nxl = slab.shape[0]
        C = min(self.fft_chunks, nxl) if G > 1 else 1
        if C <= 1:
            cplx = ops.fft_yz(slab, N)
            send = ops.pack(cplx, N, G)
            del cplx
            if G > 1:
                recv = torch.empty_like(send)
                dist.all_to_all_single(recv, send, group=self.group)
            else:
                recv = send
        else:
            recv, pending, keep = None, [], []
            for c in range(C):
                lo, hi = (c * nxl) // C, ((c + 1) * nxl) // C
                send = ops.pack(ops.fft_yz(slab[lo:hi], N), N, G)             # (G, hi-lo, N/G, P)
                if recv is None:
                    recv = send.new_empty((G, nxl) + tuple(send.shape[2:]))
                recv[r, lo:hi].copy_(send[r])                                  # my own block stays on the device
                p2p = []
                for g in range(G):
                    if g != r:
                        p2p.append(dist.P2POp(dist.irecv, recv[g, lo:hi], self._peer(g), group=self.group))
                        p2p.append(dist.P2POp(dist.isend, send[g], self._peer(g), group=self.group))
                pending.append(dist.batch_isend_irecv(p2p))
                keep.append(send)
            for reqs in pending:
                for q in reqs:
                    q.wait()
            del keep
        recv = recv.view(N, self.nyl, recv.shape[-1])      # block g holds x in [g N/G, (g+1) N/G); rows keep their pitch
        ops.fft_x(recv, N)
        return recv

    # ---- stage 3: binning + all-reduce -----------------------------------------------------------
    def bin(self, fields_k, mas_list, want_phase):
        ops, N, G = self.ops, self.dims, self.G
        mas_index = [PKL.MAS_function(m) for m in mas_list]
        L, sums, counts = ops.bin(fields_k, N, self.axis, mas_index, want_phase, self.rank * self.nyl, self.nyl)
        if G > 1:
            raw = getattr(sums, "_pylb_raw", None)
            if raw is not None and raw.is_cuda and raw.numel() == L.n_doubles + L.n_counts:
                # sums and counts share one buffer: ONE all-reduce.  The mode counts travel as float64 (exact below
                # 2^53) in the words they occupy as int64, and are converted back afterwards.
                tail = raw[L.n_doubles:]
                tail.copy_(counts.to(torch.float64))
                dist.all_reduce(raw, op=dist.ReduceOp.SUM, group=self.group)
                counts.copy_(tail.clone().round_().to(torch.int64))
            else:
                dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
        return PKL._Bins(L, sums, counts, (self.BoxSize / N ** 2) ** 3 if getattr(sums, "is_cuda", False) else None)

    # ---- whole pipelines -------------------------------------------------------------------------
    def pk_from_slab(self, slab, MAS=None):
        """Pk of an already slab-distributed overdensity field (each rank passes its (N/G,N,N) slab)."""
        dk = self.fft_slab(slab)
        out = _Result()
        PKL._finish(out, self.bin([dk], [MAS or self.MAS], True), self.dims, self.BoxSize, False)
        return out

    def run(self, pos, W=None):
        """particles -> Pk (same attributes as Pk_library.Pk), identical on every rank."""
        return self.pk_from_slab(self.density_slab(pos, W))

    def pk_comp(self, snapshot_fname, ptype, do_RSD=False, folder_out=None):
        """Distributed Pk_comp (Pk_library/Pk_snapshot.py:34-91) for one species of a binary Gadget snapshot: rank r
        reads sub-files r, r+G, ... (each file is read by exactly one rank, straight into that GPU), the particles are
        deposited with this engine's exchange mode, the overdensity uses the reference's mean Np/dims^3, and rank 0
        writes the reference's output file.  Returns the Pk object (same attributes as Pk_library.Pk) on every rank."""
        from . import readgadget
        from .Pk_snapshot import name_dict
        files = readgadget.subfiles(snapshot_fname)
        head = readgadget.header(files[0][0])
        BoxSize = head.boxsize / 1e3
        if abs(BoxSize - self.BoxSize) > 1e-6 * BoxSize:
            raise ValueError("snapshot BoxSize %g Mpc/h differs from the engine's %g" % (BoxSize, self.BoxSize))
        total = sum(int(sf.npart[ptype]) for _, sf in files)
        pos = self.ops.load_species(snapshot_fname, slice(self.rank, None, self.G), ptype, do_RSD, self.axis)
        slab = self.density_slab(pos, None, "CIC", overdensity=False)
        del pos
        self.ops.overdensity_mean(slab, total * 1.0 / self.dims ** 3)
        out = self.pk_from_slab(slab, "CIC")
        if folder_out is not None and self.rank == 0:
            z = "%.3f" % head.redshift
            fout = folder_out + "/Pk_" + name_dict[str(ptype)]
            fout += ("_RS_axis=" + str(self.axis) + "_z=" + z + ".dat") if do_RSD else ("_z=" + z + ".dat")
            np.savetxt(fout, np.transpose([out.k3D, out.Pk[:, 0], out.Pk[:, 1], out.Pk[:, 2], out.Nmodes3D]))
        return out

    def run_x(self, pos_list, W_list=None, MAS_list=None):
        """Several particle sets -> XPk (same attributes as Pk_library.XPk)."""
        F = len(pos_list)
        W_list = W_list or [None] * F
        MAS_list = list(MAS_list or [self.MAS] * F)
        dks = [self.fft_slab(self.density_slab(p, w, m)) for p, w, m in zip(pos_list, W_list, MAS_list)]
        out = _Result()
        PKL._finish(out, self.bin(dks, MAS_list, False), self.dims, self.BoxSize, True)
        return out
