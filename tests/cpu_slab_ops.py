"""numpy stand-in for pylians_b200.dist.CudaOps -- TEST INFRASTRUCTURE ONLY.

Lets the slab-decomposition host logic (partitioning, reduce-scatter, all-to-all layout, windowed
binning, all-reduce) run on CPU tensors with the gloo backend.  Local stages use the oracle
(oracle/pylians_oracle.py), scipy FFTs and a vectorised numpy restatement of the mode loop restricted
to a k-space window.  Never imported by the product package.
"""
import numpy as np
import scipy.fft as sf
import torch

from oracle import pylians_oracle as O
from pylians_b200 import Pk_library as PKL


class CpuOps(object):
    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype)

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype)

    def deposit(self, pos, W, grid, BoxSize, MAS):
        O.MA(np.ascontiguousarray(pos), grid.numpy(), BoxSize, MAS, W=W)

    def partition(self, pos, W, BoxSize, MAS, G, dims):
        pos = np.ascontiguousarray(pos)
        inv = np.float32(dims) / np.float32(BoxSize)
        dist = (pos[:, 0].astype(np.float32) * inv).astype(np.float32)
        d64 = dist.astype(np.float64)
        base = {"NGP": np.trunc(d64 + 0.5), "CIC": np.trunc(d64), "TSC": np.floor(d64 - 1.5) + 1,
                "PCS": np.floor(d64 - 2.0) + 1}[MAS].astype(np.int64) % dims      # lowest touched x-plane
        dest = base // (dims // G)
        order = np.argsort(dest, kind="stable")
        w = np.ones(len(pos), np.float32) if W is None else np.asarray(W, np.float32)
        out = np.concatenate([pos[order], w[order, None]], axis=1).astype(np.float32)
        offsets = np.searchsorted(dest[order], np.arange(G + 1)).astype(np.int32)
        return torch.from_numpy(out), torch.from_numpy(offsets)

    def deposit_window(self, xyzw, grid, x0, BoxSize, MAS, weighted, dims):
        a = xyzw.numpy()
        full = np.zeros((dims,) * 3, np.float32)
        O.MA(np.ascontiguousarray(a[:, :3]), full, BoxSize, MAS, W=np.ascontiguousarray(a[:, 3]) if weighted else None)
        planes = (x0 + np.arange(grid.shape[0])) % dims
        outside = np.ones(dims, bool); outside[planes] = False
        assert not full[outside].any(), "a particle was routed to the wrong slab"
        grid.numpy()[...] += full[planes]

    def add(self, dst, src):
        dst += src

    def overdensity_mean(self, slab, mean):
        a = slab.numpy()
        a /= np.float32(mean)
        a -= np.float32(1.0)

    def load_species(self, snapshot_fname, file_slice, ptype, do_RSD, axis):
        from pylians_b200 import readgadget
        head = readgadget.header(snapshot_fname)
        parts = []
        for name, sf in readgadget.subfiles(snapshot_fname)[file_slice]:
            if int(sf.npart[ptype]) == 0:
                continue
            pos = readgadget.read_field(name, "POS ", ptype) / np.float32(1e3)
            if do_RSD:
                vel = readgadget.read_field(name, "VEL ", ptype)
                O.pos_redshift_space(pos, vel, head.boxsize / 1e3, head.Hubble, head.redshift, axis)
            parts.append(pos)
        return np.concatenate(parts) if parts else np.zeros((0, 3), np.float32)

    def grid_sum(self, slab):
        return torch.tensor([float(np.sum(slab.numpy(), dtype=np.float64))], dtype=torch.float64)

    def overdensity_apply(self, slab, total, n_total):
        a = slab.numpy()
        a[...] = (a.astype(np.float64) * (float(n_total) / float(total[0]))).astype(np.float32) - np.float32(1.0)

    def fft_yz(self, slab, dims):
        return torch.from_numpy(sf.rfft2(slab.numpy(), axes=(1, 2)).astype(np.complex64))

    def pack(self, cplx, dims, G):
        nxl, nz = cplx.shape[0], dims // 2 + 1
        return cplx.view(nxl, G, dims // G, nz).permute(1, 0, 2, 3).contiguous()

    def fft_x(self, recv, dims):
        a = recv.numpy()
        a[...] = sf.fft(a, axis=0).astype(np.complex64)

    def bin(self, fields, dims, axis, mas_index, want_phase, y0, nyl):
        """Mode loop of Pk_library.pyx:314-381 / :628-737 on the window kx in [0,N), ky in [y0,y0+nyl)."""
        F = len(fields)
        L = PKL.get_layout(dims, F)
        middle = dims // 2
        sums = np.zeros(L.n_doubles); counts = np.zeros(L.n_counts, dtype=np.int64)
        kxx = np.arange(dims)[:, None, None]; kyy = (y0 + np.arange(nyl))[None, :, None]; kzz = np.arange(middle + 1)[None, None, :]
        kx = np.where(kxx > middle, kxx - dims, kxx); ky = np.where(kyy > middle, kyy - dims, kyy); kz = kzz + 0 * kxx
        kx, ky, kz = np.broadcast_arrays(kx, ky, kz)
        even = dims % 2 == 0
        special = (kz == 0) | ((kz == middle) & even)
        drop = special & ((kx < 0) | (((kx == 0) | ((kx == middle) & even)) & (ky < 0)))
        keep = ~drop
        n2 = kx * kx + ky * ky + kz * kz
        k = np.sqrt(n2.astype(np.float64)); kidx = k.astype(np.int64)
        if axis == 0:   kpar, kper = kx, np.sqrt((ky * ky + kz * kz).astype(np.float64)).astype(np.int64)
        elif axis == 1: kpar, kper = ky, np.sqrt((kx * kx + kz * kz).astype(np.float64)).astype(np.int64)
        else:           kpar, kper = kz, np.sqrt((kx * kx + ky * ky).astype(np.float64)).astype(np.int64)
        mu = np.where(k == 0, 0.0, kpar / np.where(k == 0, 1.0, k)); mu2 = mu * mu
        w2 = (3.0 * mu2 - 1.0) / 2.0; w4 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0
        kpar = np.abs(kpar); in1d = k <= middle; i2 = (L.kmax_par + 1) * kper + kpar
        def corr(kk, p):
            x = np.pi / dims * kk
            with np.errstate(invalid="ignore", divide="ignore"):
                return np.where(x == 0, 1.0, (x / np.sin(x)) ** p)
        def acc(off, idx, wts, mask, stride=1, col=0):
            np.add.at(sums, off + idx[mask] * stride + col, wts[mask])
        n3, n1 = L.kmax + 1, L.kmax_par + 1
        np.add.at(counts, L.o_n3d + kidx[keep], 1); np.add.at(counts, L.o_n2d + i2[keep], 1)
        np.add.at(counts, L.o_n1d + kpar[keep & in1d], 1)
        acc(L.o_k3d, kidx, k, keep)
        re, im = [], []
        for f, (dk, p) in enumerate(zip(fields, mas_index)):
            a = dk.numpy()
            mf = (corr(kx, p) * corr(ky, p) * corr(kz, p)).astype(np.float32)
            r = (a.real * mf).astype(np.float32).astype(np.float64); i = (a.imag * mf).astype(np.float32).astype(np.float64)
            re.append(r); im.append(i)
            d2 = r * r + i * i
            if f == 0 and want_phase:
                ph = np.arctan2(r, np.sqrt(d2)); acc(L.o_phase, kidx, ph * ph, keep)
            acc(L.o_p1d, kpar, d2, keep & in1d, F, f); acc(L.o_p2d, i2, d2, keep, F, f)
            for l, w in enumerate((np.ones_like(d2), w2, w4)):
                np.add.at(sums, L.o_p3d + (kidx[keep] * 3 + l) * F + f, (d2 * w)[keep])
        x = 0
        for a_ in range(F):
            for b_ in range(a_ + 1, F):
                dx = re[a_] * re[b_] + im[a_] * im[b_]
                acc(L.o_x1d, kpar, dx, keep & in1d, L.X, x); acc(L.o_x2d, i2, dx, keep, L.X, x)
                for l, w in enumerate((np.ones_like(dx), w2, w4)):
                    np.add.at(sums, L.o_x3d + (kidx[keep] * 3 + l) * L.X + x, (dx * w)[keep])
                x += 1
        return L, torch.from_numpy(sums), torch.from_numpy(counts)
