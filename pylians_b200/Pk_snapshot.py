"""Power spectra of a Gadget snapshot on the GPU: mirror of library/Pk_library/Pk_snapshot.py
(`Pk_comp` :34-91, `Pk_Gadget` :107-267).  Same arguments, same output files (`Pk_<species>[_RS_axis=a]_z=<z>.dat`
with columns k, P0, P2, P4, Nmodes), same normalisations.

Built for the GPU: particles are streamed sub-file by sub-file from disk through pinned memory into HBM
(MAS_gadget.StreamedSnapshot) and deposited into density grids that never leave the device; the
overdensity, the Omega-weighted total field and the spectra are computed there, and only the binned spectra
(a few KB) come back to be written with numpy.savetxt.
"""
import os

import numpy as np
import torch

from . import _lib
from . import MAS_library as MASL
from . import Pk_library as PKL
from . import units_library as UL
from .MAS_gadget import StreamedSnapshot

rho_crit = UL.units().rho_crit

# Pk_snapshot.py:18-22
name_dict = {"0": "GAS", "01": "GCDM", "02": "GNU", "04": "Gstars",
             "1": "CDM", "12": "CDMNU", "14": "CDMStars",
             "2": "NU", "24": "NUStars",
             "4": "Stars",
             "-1": "matter"}


def _say(msg):
    if PKL.VERBOSE:
        print(msg)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _overdensity_mean(delta, mean):
    """delta /= mean; delta -= 1.0 with numpy's float32 scalar (Pk_snapshot.py:88, 194)."""
    _lib.check(_lib.load().pylb_overdensity_mean(delta.data_ptr(), delta.numel(), float(np.float32(mean)), _stream()),
               "pylb_overdensity_mean")


def _species_field(snap, ptypes, dims, BoxSize, do_RSD, axis, weighted, report=False):
    """CIC grid of the species in `ptypes` (device), the number of particles and the float64 sum of their masses."""
    delta = torch.zeros((dims, dims, dims), dtype=torch.float32, device=snap.dev)
    count, mass_sum = 0, 0.0
    mass_dev = torch.zeros(1, dtype=torch.float64, device=snap.dev)
    lo = torch.full((3,), float("inf"), device=snap.dev)
    hi = torch.full((3,), float("-inf"), device=snap.dev)
    for ptype in ptypes:
        for pos, vel, mass, n in snap.blocks(ptype, want_vel=do_RSD, want_mass=weighted):
            if report:                                        # :61-63, before the redshift-space shift
                mn, mx = torch.aminmax(pos, dim=0)
                lo, hi = torch.minimum(lo, mn), torch.maximum(hi, mx)
            if do_RSD:
                snap.to_redshift_space(pos, vel, BoxSize, axis)
            if weighted:
                if not torch.is_tensor(mass):
                    mass_sum += n * float(np.float32(mass))
                    mass = torch.full((n,), mass, dtype=torch.float32, device=snap.dev)
                else:
                    mass_dev += snap.sum_f64(mass)
                MASL.MA(pos, delta, BoxSize, "CIC", W=mass)
            else:
                MASL.MA(pos, delta, BoxSize, "CIC")
            count += n
    if report and count:
        lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
        for a, name in enumerate("XYZ"):
            _say("%.3f < %s [Mpc/h] < %.3f%s" % (lo[a], name, hi[a], "\n" if a == 2 else ""))
    if weighted:
        mass_sum += float(mass_dev.item())
    return delta, count, mass_sum


def _save(fout, k, P, Nmodes):
    np.savetxt(fout, np.transpose([k, P[:, 0], P[:, 1], P[:, 2], Nmodes]))


def Pk_comp(snapshot_fname, ptype, dims, do_RSD, axis, cpus, folder_out):
    """Pk_snapshot.py:34-91: P(k) of one species (plain counts) or of all of them (ptype = -1, mass weighted)."""
    _say("Computing power spectrum...")
    snap = StreamedSnapshot(snapshot_fname)
    head = snap.head
    BoxSize = head.boxsize / 1e3                              # Mpc/h
    z = "%.3f" % head.redshift
    fout = folder_out + "/Pk_" + name_dict[str(ptype)]
    fout += ("_RS_axis=" + str(axis) + "_z=" + z + ".dat") if do_RSD else ("_z=" + z + ".dat")
    if do_RSD:
        _say("moving particles to redshift-space...")
    if ptype == -1:                                           # :70-82, masses from the header table or the MASS block
        delta, count, mass_sum = _species_field(snap, [0, 1, 2, 3, 4, 5], dims, BoxSize, do_RSD, axis, True, PKL.VERBOSE)
        mean = mass_sum / dims ** 3
    else:                                                     # :84-86
        delta, count, _ = _species_field(snap, [ptype], dims, BoxSize, do_RSD, axis, False, PKL.VERBOSE)
        mean = count * 1.0 / dims ** 3
    _overdensity_mean(delta, mean)
    Pk = PKL.Pk(delta, BoxSize, axis=axis, MAS="CIC", threads=cpus)
    del delta
    _save(fout, Pk.k3D, Pk.Pk, Pk.Nmodes3D)


def Pk_Gadget(snapshot_fname, dims, particle_type, do_RSD, axis, cpus, folder_out=None):
    """Pk_snapshot.py:107-267: auto- and cross-spectra of the species in `particle_type` and the spectrum of their
    Omega-weighted sum; a single species (or [-1] = total matter) goes through Pk_comp."""
    if folder_out is None:
        folder_out = os.getcwd()
    if len(particle_type) == 1:
        Pk_comp(snapshot_fname, particle_type[0], dims, do_RSD, axis, cpus, folder_out)
        return None

    _say("\nREADING SNAPSHOTS PROPERTIES")
    snap = StreamedSnapshot(snapshot_fname)
    head = snap.head
    BoxSize = head.boxsize / 1e3                              # Mpc/h
    Nall = [snap.count(t) for t in range(6)]
    Masses = head.massarr * 1e10                              # Msun/h
    z = "%.3f" % head.redshift
    dims3 = dims ** 3

    # Omega of each component, :131-146
    Omega_c = Masses[1] * Nall[1] / BoxSize ** 3 / rho_crit
    Omega_n = Masses[2] * Nall[2] / BoxSize ** 3 / rho_crit
    Omega_g, Omega_s = 0.0, 0.0
    if Nall[0] > 0:
        if Masses[0] > 0:
            Omega_g = Masses[0] * Nall[0] / BoxSize ** 3 / rho_crit
            Omega_s = Masses[4] * Nall[4] / BoxSize ** 3 / rho_crit
        else:
            def total_mass(t):                                # np.sum(MASS block * 1e10, dtype=float64), on the device
                acc = torch.zeros(1, dtype=torch.float64, device=snap.dev)
                for _, sf in snap.files:
                    n = int(sf.npart[t])
                    if n == 0:
                        continue
                    if sf.massarr[t] != 0:
                        acc += n * float(np.float32(sf.massarr[t] * 1e10))
                        continue
                    snap._turn += 1
                    m = snap._upload(sf, "MASS", t, (n,), "mass")
                    _lib.check(snap.lib.pylb_scale_f32(m.data_ptr(), n, 1e10, snap.stream.cuda_stream), "pylb_scale_f32")
                    acc += snap.sum_f64(m)
                return float(acc.item())
            Omega_g = total_mass(0) / BoxSize ** 3 / rho_crit
            Omega_s = total_mass(4) / BoxSize ** 3 / rho_crit
    _say("Omega_gas    =  %s" % Omega_g)
    _say("Omega_cdm    =  %s" % Omega_c)
    _say("Omega_nu     =  %s" % Omega_n)
    _say("Omega_star   =  %s" % Omega_s)
    _say("Omega_m      =  %s" % (Omega_g + Omega_c + Omega_n + Omega_s))
    _say("Omega_m snap =  %s" % head.omega_m)
    Omega_dict = {0: Omega_g, 1: Omega_c, 2: Omega_n, 4: Omega_s}

    suffix = ("_RS_axis=" + str(axis) + "_z=" + z + ".dat") if do_RSD else ("_z=" + z + ".dat")

    # overdensity of every requested species (plain counts, :176-194); the grids stay in HBM
    delta = {}
    for ptype in particle_type:
        d, count, _ = _species_field(snap, [ptype], dims, BoxSize, do_RSD, axis, False)
        _overdensity_mean(d, count * 1.0 / dims3)
        delta[ptype] = d

    # auto- and cross-spectra of every pair, :199-237
    for i, ptype1 in enumerate(particle_type):
        for ptype2 in particle_type[i + 1:]:
            fout1 = folder_out + "/Pk_" + name_dict[str(ptype1)] + suffix
            fout2 = folder_out + "/Pk_" + name_dict[str(ptype2)] + suffix
            fout12 = folder_out + "/Pk_" + name_dict[str(ptype1) + str(ptype2)] + suffix
            _say("\nComputing the auto- and cross-power spectra of types:  %s - %s" % (ptype1, ptype2))
            _say("saving results in:")
            _say("%s \n%s \n%s" % (fout1, fout2, fout12))
            data = PKL.XPk([delta[ptype1], delta[ptype2]], BoxSize, axis=axis, MAS=["CIC", "CIC"], threads=cpus)
            _save(fout12, data.k3D, data.XPk[:, :, 0], data.Nmodes3D)
            _save(fout1, data.k3D, data.Pk[:, :, 0], data.Nmodes3D)
            _save(fout2, data.k3D, data.Pk[:, :, 1], data.Nmodes3D)

    # spectrum of the Omega-weighted sum of the components, :242-267
    _say("\ncomputing P(k) of all components")
    lib = _lib.load()
    delta_tot = torch.zeros((dims, dims, dims), dtype=torch.float32, device=snap.dev)
    Omega_tot, fout = 0.0, folder_out + "/Pk_"
    for ptype in particle_type:
        _lib.check(lib.pylb_axpy_f32(delta_tot.data_ptr(), delta[ptype].data_ptr(), float(np.float32(Omega_dict[ptype])),
                                     delta_tot.numel(), _stream()), "pylb_axpy_f32")
        Omega_tot += Omega_dict[ptype]
        fout += name_dict[str(ptype)] + "+"
    _lib.check(lib.pylb_divide(delta_tot.data_ptr(), delta_tot.numel(), float(np.float32(Omega_tot)), _stream()),
               "pylb_divide")
    del delta
    fout = fout[:-1]
    data = PKL.Pk(delta_tot, BoxSize, axis=axis, MAS="CIC", threads=cpus)
    del delta_tot
    _save(fout + suffix, data.k3D, data.Pk, data.Nmodes3D)
