#!/usr/bin/env python
"""Generates tests/golden/*.npz.  Run HERE (needs /root/reference and the built oracle):

    make -C oracle && python tools/make_fixtures.py

* optics_<arm>.npz : the COSY tables of the shipped .dat files, parsed by the oracle's
  restatement of transp_init / mc_*_recon's loader (shared/transp.f:294-474).
* transport_<arm>.npz : seeded single-arm input rows and the ORACLE's outputs for them
  (the reference cannot be executed here, so these are regression vectors of the oracle, not
  reference outputs -- parity stays "unpinned").
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simc_gfortran_b200.optics import OpticsTables  # noqa: E402
from tests.oracle_lib import Oracle, transport_inputs  # noqa: E402

REF = os.environ.get("SIMC_REFERENCE", "/root/reference")
FILES = {
    1: ("hms", "hms/forward_cosy.dat", "hms/recon_cosy.dat"),
    5: ("shms", "shms/shms_forward.dat", "shms/shms_recon.dat"),
    2: ("sos", "sos/forward_cosy.dat", "sos/recon_cosy.dat"),
    3: ("hrsr", "hrsr/hrs_forward_cosy.dat", "hrsr/hrs_recon_cosy.dat"),
    4: ("hrsl", "hrsl/hrs_forward_cosy.dat", "hrsl/hrs_recon_cosy.dat"),
}


def main():
    orc = Oracle()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for arm, (name, fwd, rec) in FILES.items():
        orc.load_optics(arm, os.path.join(REF, fwd), os.path.join(REF, rec))
        t = orc.export_optics(arm)
        t.save(os.path.join(out_dir, f"optics_{name}.npz"))
        print(name, "classes", t.n_classes, "fwd terms", len(t.fwd_coeff), "rec terms", len(t.rec_coeff))
        if not orc.has_arm(arm):
            continue
        for tag, kw in (("", dict(ms=True, wcs=True, decay=False)),):
            n = 4096
            inp = transport_inputs(arm, n, seed=20240611)
            out, flags = orc.transport_batch(arm, inp, seed=20240611, ctau=0.0, **kw)
            np.savez_compressed(os.path.join(out_dir, f"transport_{name}{tag}.npz"), inp=inp, out=out, flags=flags,
                                seed=20240611)
            print("  golden rows", n, "accepted", int((flags == 0).sum()))


def make_sf():
    """benharsf_12.dat (the reference's data file for A = 12, dbase.f:600) -> simc_gfortran_b200/data/benharsf_12.npz"""
    rows = []
    with open(os.path.join(REF, "benharsf_12.dat")) as f:
        n_pm, n_em = (int(x) for x in f.readline().split())
        for line in f:
            if line.strip():
                rows.append([float(x) for x in line.split()])
    a = np.array(rows).reshape(n_pm, n_em, 6)
    np.savez_compressed(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"), pm=a[:, 0, 0], em=a[0, :, 1],
                        sf_proton=a[:, :, 2], sf_neutron=a[:, :, 3], dpm=a[:, 0, 4], dem=a[0, :, 5])
    print("benharsf_12", n_pm, n_em, "sum", a[:, :, 2].sum())


def read_cteq5_tbl(path):
    """ReadTbl (cteq5/Ctq5Pdf.f:239-281): list-directed reads between comment lines."""
    with open(path) as f:
        lines = f.read().split("\n")
    toks = " ".join(lines[2:]).replace("D", "E")
    # line 3: Dr, Fl, Al, masses; then 'NX NT NfMx' header + values; then QINI.. ; XMIN.. ; table
    it = iter(lines)
    next(it); next(it)
    dr, fl, al, *masses = (float(x) for x in next(it).split())
    next(it)
    nx, nt, nfmx = (int(x) for x in next(it).split())
    next(it)
    rest = []
    for line in it:
        rest.append(line)
    def take(block_lines, n):
        vals = []
        while len(vals) < n:
            vals += [float(x) for x in block_lines.pop(0).split()]
        assert len(vals) == n, (len(vals), n)
        return vals
    q = take(rest, 2 + nt + 1)
    rest.pop(0)
    x = take(rest, 1 + nx + 1)
    rest.pop(0)
    upd = take(rest, (nx + 1) * (nt + 1) * (nfmx + 3))
    return dict(nx=nx, nt=nt, nfmx=nfmx, lam=al, qini=q[0], qmax=q[1], qv=np.array(q[2:]), xmin=x[0],
                xv=np.array(x[1:]), upd=np.array(upd))


def make_semi():
    """cteq5/cteq5m.tbl (SetCtq5 with Iset = 1, semi_physics.f:226-229) and deut.dat (dbase.f:564) ->
    simc_gfortran_b200/data/cteq5m.npz, simc_gfortran_b200/data/pfermi_deut.npz"""
    t = read_cteq5_tbl(os.path.join(REF, "cteq5", "cteq5m.tbl"))
    np.savez_compressed(os.path.join(ROOT, "simc_gfortran_b200", "data", "cteq5m.npz"), **t)
    print("cteq5m", t["nx"], t["nt"], t["nfmx"], t["lam"], len(t["upd"]))
    rows = []
    with open(os.path.join(REF, "deut.dat")) as f:
        for line in f:
            if line.strip():
                rows.append([float(x.replace("d", "e").replace("D", "e")) for x in line.split()])
    a = np.array(rows)[:2000]
    np.savez_compressed(os.path.join(ROOT, "simc_gfortran_b200", "data", "pfermi_deut.npz"), pval=a[:, 0], mprob=a[:, 1])
    print("deut.dat", a.shape, a[-1])


def read_theory(path):
    """theory_init's file format (init.f:853-868): header, shell lines, then (Pm, rho) rows; a shell ends where
    Pm stops increasing."""
    toks = open(path).read().split()
    n_shells, absorption, e_fermi = int(toks[0]), float(toks[1]), float(toks[2])
    sh = np.array([float(x) for x in toks[3:3 + 4 * n_shells]]).reshape(n_shells, 4)
    rows = np.array([float(x) for x in toks[3 + 4 * n_shells:]]).reshape(-1, 2)
    cuts = [0] + [i for i in range(1, len(rows)) if rows[i, 0] <= rows[i - 1, 0]] + [len(rows)]
    assert len(cuts) == n_shells + 1, (len(cuts), n_shells)
    return dict(n_shells=n_shells, absorption=absorption, e_fermi=e_fermi, nprot=sh[:, 0], em=sh[:, 1], emsig=sh[:, 2],
                bs_norm=sh[:, 3], n_pm=np.diff(cuts).astype(np.int32), pm_first=rows[cuts[:-1], 0],
                pm_bin=rows[np.array(cuts[:-1]) + 1, 0] - rows[cuts[:-1], 0], rho=rows[:, 1])


def make_theory():
    """h2.theory and c12.theory (theory_init, init.f:838-851) -> tests/golden/theory_h2.npz, theory_c12.npz"""
    for name in ("h2", "c12"):
        t = read_theory(os.path.join(REF, name + ".theory"))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"theory_{name}.npz"), **t)
        print(name, t["n_shells"], t["n_pm"], t["pm_first"], t["pm_bin"], t["e_fermi"])


def make_maid():
    """maidpipn.dat / maidpimp.dat (sigmaid, physics_pion.f:611-625) -> tests/golden/maid_pipn.npz, maid_pimp.npz:
    the six angle rows and four columns sig0 reads, [25, 46, 6, 4]."""
    for name in ("pipn", "pimp"):
        tbl = np.zeros((25, 46, 6, 4))
        with open(os.path.join(REF, f"maid{name}.dat")) as f:
            for iq in range(25):
                for iw in range(46):
                    for ith in range(23):
                        line = f.readline()
                        if ith < 6:
                            tbl[iq, iw, ith] = [float(line[0:11]), float(line[11:19]), float(line[19:27]), float(line[27:35])]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"maid_{name}.npz"), tbl=tbl)
        print("maid", name, tbl.shape, tbl[..., 0].max())


def make_saghai():
    """saghai_proton.dat / saghai_sigma0.dat as dbase.f:644-679 reads them -> tests/golden/saghai.npz: proton
    [12, 19, 11, 10] and sigma0 [12, 19, 10, 20] float32 (zrff1..6 then ziff1..6; C order of the Fortran arrays
    (iread, iq2, iang), i.e. iread fastest).  '(6e12.4)' fields are cut by column like a Fortran formatted read."""
    out = {}
    for which, (name, n1, n2) in enumerate((("proton", 10, 11), ("sigma0", 20, 10))):
        tbl = np.zeros((12, 19, n2, n1), dtype=np.float32)
        with open(os.path.join(REF, f"saghai_{name}.dat")) as f:
            for ir in range(n1):
                for iq in range(n2):
                    if which == 0:
                        f.readline()
                    for ia in range(19):
                        if which == 1:
                            f.readline()
                        f.readline()
                        for half in range(2):
                            line = f.readline().rstrip("\n").ljust(72)
                            # a blank field reads as zero (the Sigma0 file starts with a fragment of a line, so
                            # every read of dbase.f:662-676 sits one line early: its "amplitude" lines are the
                            # five-number kinematic line and the first amplitude line -- reproduced as read)
                            v = [np.float32(float(line[12 * k:12 * k + 12].strip() or 0.0)) for k in range(6)]
                            for k in range(3):
                                tbl[3 * half + k, ia, iq, ir] = v[2 * k]
                                tbl[6 + 3 * half + k, ia, iq, ir] = v[2 * k + 1]
            # saghai_proton.dat goes on (14 values of s in the file, dbase.f reads the first 10)
        out[name] = tbl
        print("saghai", name, tbl.shape, float(np.abs(tbl).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "saghai.npz"), **out)


def make_fdss():
    """fdss/KANLO.GRID (fDSS with IH = 2, IO = 1: kaons at NLO, semi_physics.f:497-505) -> tests/golden/fdss_kanlo.npz"""
    rows = []
    with open(os.path.join(REF, "fdss", "KANLO.GRID")) as f:
        for line in f:
            if len(line) >= 90:
                rows.append([float(line[10 * k:10 * k + 10]) for k in range(9)])
    a = np.array(rows).reshape(34, 24, 9)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fdss_kanlo.npz"), parton=a)
    print("fdss KANLO", a.shape)


def make_he3():
    """he3.dat (dbase.f:566-567) and benharsf_3mod.dat (dbase.f:596-598) -> tests/golden/pfermi_he3.npz, benharsf_3mod.npz"""
    rows = []
    with open(os.path.join(REF, "he3.dat")) as f:
        for line in f:
            if line.strip():
                rows.append([float(x.replace("d", "e").replace("D", "e")) for x in line.split()[:2]])
    a = np.array(rows)[:2000]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pfermi_he3.npz"), pval=a[:, 0], mprob=a[:, 1])
    rows = []
    with open(os.path.join(REF, "benharsf_3mod.dat")) as f:
        n_pm, n_em = (int(x) for x in f.readline().split())
        for line in f:
            if line.strip():
                rows.append([float(x) for x in line.split()])
    b = np.array(rows).reshape(n_pm, n_em, 6)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "benharsf_3mod.npz"), pm=b[:, 0, 0], em=b[0, :, 1],
                        sf_proton=b[:, :, 2], sf_neutron=b[:, :, 3], dpm=b[:, 0, 4], dem=b[0, :, 5])
    print("he3", a.shape, n_pm, n_em)


def make_field():
    """trg_field_map.dat (trgInit, trg_track.f:305-313: 51 x 51 rows 'z r Bz Br ...', z fastest) ->
    simc_gfortran_b200/data/trg_field_map.npz: bz, br in the file's reading order"""
    a = np.loadtxt(os.path.join(REF, "trg_field_map.dat"))
    assert a.shape == (51 * 51, 7)
    np.savez_compressed(os.path.join(ROOT, "simc_gfortran_b200", "data", "trg_field_map.npz"), bz=a[:, 2], br=a[:, 3])
    print("field map", a.shape, a[:, 2].max())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "field":
        make_field()
        sys.exit(0)
    make_field()
    make_he3()
    make_fdss()
    make_maid()
    make_saghai()
    make_sf()
    make_semi()
    make_theory()
    main()
