"""ctypes access to the CPU oracle (oracle/liboracle.so) -- test infrastructure only.

Only tests/, tools/make_fixtures.py, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

EVENT_NREC = 60          # SIMC_EVENT_NREC, SIMC_NTUPLE_MAXCOL of include/simc_b200.h
NTUPLE_MAXCOL = 68

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-j8"], check=True, capture_output=True)
    return ORACLE_SO


RADC_NOUT = 26          # include/simc_b200.h: SIMC_RADC_NOUT


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# momentum (MeV/c) and mass^2 used for the synthetic single-arm rows of each spectrometer:
# the C1 kinematics of SURVEY 8(d) (electron in the HMS, proton in the SHMS)
ARM_KIN = {1: (4531.0, 0.51099906 ** 2), 5: (5122.0, 938.27231 ** 2), 2: (1300.0, 0.51099906 ** 2),
           3: (1292.58, 493.677 ** 2), 4: (1571.272, 0.51099906 ** 2)}
# half-widths of the generated box: delta %, x cm, y cm, dxdz, dydz (1.2 x the SPedge box of C1)
ARM_BOX = {1: (12.0, 0.2, 2.5, 0.12, 0.06), 5: (18.0, 0.2, 2.5, 0.06, 0.11), 2: (24.0, 0.2, 2.0, 0.09, 0.09),
           3: (6.0, 0.2, 2.0, 0.08, 0.04), 4: (6.0, 0.2, 2.0, 0.08, 0.04)}


def transport_inputs(arm, n, seed, p_spec=None, m2=None):
    """Seeded rows {dpp,x,y,z,dxdz,dydz,m2,p_spec,fry} (SoA [9,n]) for the single-arm entry points."""
    rng = np.random.default_rng(seed + 1000 * arm)
    box = ARM_BOX[arm]
    inp = np.zeros((9, n))
    inp[0] = rng.uniform(-box[0], box[0], n)
    inp[1] = rng.uniform(-box[1], box[1], n)
    inp[2] = rng.uniform(-box[2], box[2], n)
    inp[4] = rng.uniform(-box[3], box[3], n)
    inp[5] = rng.uniform(-box[4], box[4], n)
    inp[6] = ARM_KIN[arm][1] if m2 is None else m2
    inp[7] = ARM_KIN[arm][0] if p_spec is None else p_spec
    if arm in (1, 5):
        inp[8] = inp[1]      # fry = xtar_init (simc.f:1441,1463) for HMS/SHMS
    else:
        inp[8] = rng.uniform(-0.1, 0.1, n)   # fry = -rastery for SOS/HRS (simc.f:1350,1472)
    return inp


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        self.L = C.CDLL(ORACLE_SO)
        self.L.oracle_last_error.restype = C.c_char_p

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.oracle_last_error().decode())

    def has_arm(self, arm):
        return arm in (1, 2, 3, 4, 5)

    def set_sf_table(self, pm, em, sf):
        pm = np.ascontiguousarray(pm, dtype=np.float64)
        em = np.ascontiguousarray(em, dtype=np.float64)
        sf = np.ascontiguousarray(sf, dtype=np.float64)
        assert sf.shape == (len(pm), len(em))
        self._check(self.L.oracle_set_sf_table(len(pm), len(em), _p(pm), _p(em), _p(sf)))

    def sf_batch(self, em, pm):
        em = np.ascontiguousarray(em, dtype=np.float64)
        pm = np.ascontiguousarray(pm, dtype=np.float64)
        out = np.zeros(len(em))
        self._check(self.L.oracle_sf_batch(C.c_int64(len(em)), _p(em), _p(pm), _p(out)))
        return out

    def load_optics(self, arm, fwd, rec):
        self._check(self.L.oracle_load_optics(arm, fwd.encode(), rec.encode()))

    def export_optics(self, arm):
        from simc_gfortran_b200.optics import OpticsTables
        nc, nf, nr = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.oracle_optics_sizes(arm, C.byref(nc), C.byref(nf), C.byref(nr)))
        t = OpticsTables(arm=arm, class_start=np.zeros(nc.value + 1, np.int32), fwd_coeff=np.zeros((nf.value, 5)),
                         fwd_expon=np.zeros((nf.value, 5), np.int8), length_cm=np.zeros(nc.value),
                         adrift=np.zeros(nc.value, np.int8), driftdist=np.zeros(nc.value),
                         rec_coeff=np.zeros((nr.value, 4)), rec_expon=np.zeros((nr.value, 5), np.int8))
        self._check(self.L.oracle_optics_export(arm, _p(t.class_start), _p(t.fwd_coeff), _p(t.fwd_expon),
                                                _p(t.length_cm), _p(t.adrift), _p(t.driftdist), _p(t.rec_coeff),
                                                _p(t.rec_expon)))
        return t

    def set_optics(self, t):
        cs = np.ascontiguousarray(t.class_start, np.int32)
        fc = np.ascontiguousarray(t.fwd_coeff, np.float64)
        fe = np.ascontiguousarray(t.fwd_expon, np.int8)
        ln = np.ascontiguousarray(t.length_cm, np.float64)
        ad = np.ascontiguousarray(t.adrift, np.int8)
        dd = np.ascontiguousarray(t.driftdist, np.float64)
        rc_ = np.ascontiguousarray(t.rec_coeff, np.float64)
        re_ = np.ascontiguousarray(t.rec_expon, np.int8)
        self._check(self.L.oracle_set_optics(t.arm, len(cs) - 1, _p(cs), _p(fc), _p(fe), _p(ln), _p(ad), _p(dd),
                                             len(rc_), _p(rc_), _p(re_)))

    def transport_batch(self, arm, inp, seed, ms=True, wcs=True, decay=False, coll=False, ctau=0.0):
        inp = np.ascontiguousarray(inp, np.float64)
        n = inp.shape[1]
        out = np.zeros((12, n))
        flags = np.zeros(n, np.int32)
        self._check(self.L.oracle_transport_batch(arm, C.c_int64(n), _p(inp), C.c_uint64(seed), int(ms), int(wcs),
                                                  int(decay), int(coll), C.c_double(ctau), _p(out), _p(flags)))
        return out, flags

    def ranlux(self, seed, lux, n):
        out = np.zeros(n)
        self.L.oracle_ranlux(C.c_int(seed), C.c_int(lux), C.c_int64(n), _p(out))
        return out

    def philox_uniforms(self, seed, try_index, n):
        out = np.zeros(n)
        self.L.oracle_philox_uniforms(C.c_uint64(seed), C.c_uint64(try_index), C.c_int64(n), _p(out))
        return out


def _oracle_run(self, cfg, first, n, seed, threads=1, ranlux=False):
    """The loop of simc.f:169-351 on the CPU oracle; returns a simc_gfortran_b200.Accum.
    ranlux=True uses the reference's RANLUX stream (one generator per thread) instead of Philox."""
    from simc_gfortran_b200.lib import Accum
    acc = Accum()
    self._check(self.L.oracle_accum_clear(C.byref(cfg), C.byref(acc)))
    self._check(self.L.oracle_run_rng(C.byref(cfg), C.c_int64(first), C.c_int64(n), C.c_uint64(seed), int(threads),
                                      int(bool(ranlux)), C.byref(acc)))
    return acc


def _oracle_event_batch(self, cfg, first, n, seed):
    rec = np.zeros((EVENT_NREC, n))
    status = np.zeros(n, np.int32)
    self._check(self.L.oracle_event_batch(C.byref(cfg), C.c_int64(first), C.c_int64(n), C.c_uint64(seed), _p(rec),
                                          _p(status)))
    return rec, status


def _oracle_ntuple_batch(self, cfg, first, n, seed):
    rows = np.zeros((max(n, 1), NTUPLE_MAXCOL))
    tries = np.zeros(max(n, 1), np.int64)
    nc, nr = C.c_int32(0), C.c_int64(0)
    self._check(self.L.oracle_ntuple_batch(C.byref(cfg), C.c_int64(first), C.c_int64(n), C.c_uint64(seed), _p(rows),
                                           C.byref(nc), C.byref(nr), _p(tries)))
    return rows[:nr.value, :nc.value].copy(), tries[:nr.value].copy()


Oracle.ntuple_batch = _oracle_ntuple_batch
Oracle.run = _oracle_run
Oracle.event_batch = _oracle_event_batch


def _oracle_radc_batch(self, cfg, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros((RADC_NOUT, inp.shape[1]))
    self._check(self.L.oracle_radc_batch(C.byref(cfg), C.c_int64(inp.shape[1]), _p(inp), _p(out)))
    return out


Oracle.radc_batch = _oracle_radc_batch


# ---- semi-inclusive production (C4): tables and stage-level entry points ---------------------------
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cteq5_fixture():
    """simc_gfortran_b200/data/cteq5m.npz (tools/make_fixtures.py: the reference's cteq5/cteq5m.tbl)."""
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "cteq5m.npz"))
    return {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}


def load_pfermi_fixture():
    """simc_gfortran_b200/data/pfermi_deut.npz (the reference's deut.dat)."""
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "pfermi_deut.npz"))
    return z["pval"], z["mprob"]


def _oracle_set_pfermi_table(self, pval, mprob):
    pval = np.ascontiguousarray(pval, np.float64)
    mprob = np.ascontiguousarray(mprob, np.float64)
    self._check(self.L.oracle_set_pfermi_table(len(pval), _p(pval), _p(mprob)))


def _oracle_set_cteq5_table(self, t):
    xv = np.ascontiguousarray(t["xv"], np.float64)
    qv = np.ascontiguousarray(t["qv"], np.float64)
    upd = np.ascontiguousarray(t["upd"], np.float64)
    self._check(self.L.oracle_set_cteq5_table(int(t["nx"]), int(t["nt"]), int(t["nfmx"]), C.c_double(t["lam"]),
                                              C.c_double(t["qini"]), C.c_double(t["qmax"]), C.c_double(t["xmin"]),
                                              _p(xv), _p(qv), _p(upd)))


def _oracle_ctq5pdf_batch(self, iparton, x, q):
    x = np.ascontiguousarray(x, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    out = np.zeros(len(x))
    self._check(self.L.oracle_ctq5pdf_batch(int(iparton), C.c_int64(len(x)), _p(x), _p(q), _p(out)))
    return out


def _oracle_christy_batch(self, w2, q2):
    w2 = np.ascontiguousarray(w2, np.float64)
    q2 = np.ascontiguousarray(q2, np.float64)
    out = np.zeros((6, len(w2)))
    self._check(self.L.oracle_christy_batch(C.c_int64(len(w2)), _p(w2), _p(q2), _p(out)))
    return out


def _oracle_semi_batch(self, cfg, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros((16, inp.shape[1]))
    self._check(self.L.oracle_semi_batch(C.byref(cfg), C.c_int64(inp.shape[1]), _p(inp), _p(out)))
    return out


Oracle.set_pfermi_table = _oracle_set_pfermi_table
Oracle.set_cteq5_table = _oracle_set_cteq5_table
Oracle.ctq5pdf_batch = _oracle_ctq5pdf_batch
Oracle.christy_batch = _oracle_christy_batch
Oracle.semi_batch = _oracle_semi_batch


# ---- independent-particle spectral function (theory files) -------------------------------------------
def load_theory_fixture(name):
    """tests/golden/theory_h2.npz / theory_c12.npz (tools/make_fixtures.py: the reference's *.theory files)."""
    z = np.load(os.path.join(GOLDEN, f"theory_{name}.npz"))
    return {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}


def write_theory_file(t, path):
    """Writes a theory table in the reference's text format (read back by theory_init-style readers)."""
    with open(path, "w") as f:
        f.write(f"{int(t['n_shells'])}\t{float(t['absorption'])!r}\t{float(t['e_fermi'])!r}\n")
        for m in range(int(t["n_shells"])):
            f.write(f"{float(t['nprot'][m])!r}\t{float(t['em'][m])!r}\t{float(t['emsig'][m])!r}\t{float(t['bs_norm'][m])!r}\n")
        pos = 0
        for m in range(int(t["n_shells"])):
            for k in range(int(t["n_pm"][m])):
                f.write(f"  {float(t['pm_first'][m] + k * t['pm_bin'][m])!r}  {float(t['rho'][pos + k])!r}\n")
            pos += int(t["n_pm"][m])


def _oracle_set_theory_table(self, t, doing_heavy):
    a = {k: np.ascontiguousarray(t[k], np.float64) for k in ("nprot", "em", "emsig", "bs_norm", "pm_first", "pm_bin", "rho")}
    n_pm = np.ascontiguousarray(t["n_pm"], np.int32)
    self._check(self.L.oracle_set_theory_table(int(bool(doing_heavy)), int(t["n_shells"]), C.c_double(t["absorption"]),
                                               C.c_double(t["e_fermi"]), _p(a["nprot"]), _p(a["em"]), _p(a["emsig"]),
                                               _p(a["bs_norm"]), _p(n_pm), _p(a["pm_first"]), _p(a["pm_bin"]), _p(a["rho"])))


def _oracle_theory_batch(self, cfg, em, pm):
    em = np.ascontiguousarray(em, np.float64)
    pm = np.ascontiguousarray(pm, np.float64)
    out = np.zeros(len(em))
    self._check(self.L.oracle_theory_batch(C.byref(cfg), C.c_int64(len(em)), _p(em), _p(pm), _p(out)))
    return out


Oracle.set_theory_table = _oracle_set_theory_table
Oracle.theory_batch = _oracle_theory_batch


def write_pfermi_file(pval, mprob, path):
    """deut.dat-style momentum distribution in the reference's text format (Fortran d exponents included)."""
    with open(path, "w") as f:
        for a, b in zip(pval, mprob):
            f.write(f"  {float(a)!r}       {float(b):.15e}".replace("e-", "d-").replace("e+", "d+") + "\n")


# ---- MAID-2007 table (peepi below W = 2 GeV) ------------------------------------------------------------
def load_maid_fixture(ipi):
    """tests/golden/maid_pipn.npz (ipi 3) / maid_pimp.npz (ipi 4): the slice [25, 46, 6, 4] sigmaid's sig0 reads."""
    z = np.load(os.path.join(GOLDEN, "maid_pipn.npz" if ipi == 3 else "maid_pimp.npz"))
    return z["tbl"]


def write_maid_file(tbl, path):
    """The reference's layout: 25 x 46 x 23 rows of (f11.6,18f8.4); rows / columns the fixture lacks are zero."""
    with open(path, "w") as f:
        for iq in range(25):
            for iw in range(46):
                for ith in range(23):
                    v = np.zeros(18)
                    if ith < 6:
                        v[:4] = tbl[iq, iw, ith]
                    f.write(f"{v[0]:11.5f}" + "".join(f"{x:8.4f}" for x in v[1:]) + "\n")


def _oracle_set_maid_table(self, ipi, tbl):
    tbl = np.ascontiguousarray(tbl, np.float64).ravel() if tbl is not None else np.zeros(0)
    self._check(self.L.oracle_set_maid_table(int(ipi), C.c_int64(len(tbl)), _p(tbl)))


def _oracle_sigmaid_batch(self, ipi, q2, w, e0, costh, phi):
    arrs = [np.ascontiguousarray(a, np.float64) for a in (q2, w, e0, costh, phi)]
    out = np.zeros(len(arrs[0]))
    self._check(self.L.oracle_sigmaid_batch(int(ipi), C.c_int64(len(out)), *[_p(a) for a in arrs], _p(out)))
    return out


Oracle.set_maid_table = _oracle_set_maid_table
Oracle.sigmaid_batch = _oracle_sigmaid_batch


# ---- Saghai amplitude tables (ntuple column sigcm1 of kaon production) ---------------------------------
def load_saghai_fixture(which):
    """tests/golden/saghai.npz: which = 0 K+ Lambda [12, 19, 11, 10], 1 K+ Sigma0 [12, 19, 10, 20], float32, as
    dbase.f:644-679 reads saghai_proton.dat / saghai_sigma0.dat (zrff1..6 then ziff1..6; memory order = the Fortran
    arrays')."""
    z = np.load(os.path.join(GOLDEN, "saghai.npz"))
    return np.ascontiguousarray(z["proton" if which == 0 else "sigma0"], np.float32)


def _oracle_set_saghai_table(self, which, tbl):
    tbl = np.ascontiguousarray(tbl, np.float32).ravel() if tbl is not None else np.zeros(0, np.float32)
    self._check(self.L.oracle_set_saghai_table(int(which), C.c_int64(len(tbl)), _p(tbl)))


def _oracle_saghai_batch(self, lam, mrec_struck, ss, q22, angl, theta, phi, epsi):
    arrs = [np.ascontiguousarray(a, np.float64) for a in (ss, q22, angl, theta, phi, epsi)]
    out = np.zeros(len(arrs[0]))
    self._check(self.L.oracle_saghai_batch(int(bool(lam)), C.c_double(mrec_struck), C.c_int64(len(out)),
                                           *[_p(a) for a in arrs], _p(out)))
    return out


def _oracle_fint(self, arg, nent, ent, table):
    arg = np.ascontiguousarray(arg, np.float32)
    nent = np.ascontiguousarray(nent, np.int32)
    ent = np.ascontiguousarray(ent, np.float32)
    table = np.ascontiguousarray(table, np.float32)
    self.L.oracle_fint.restype = C.c_double
    return float(self.L.oracle_fint(len(arg), _p(arg), _p(nent), _p(ent), _p(table)))


Oracle.set_saghai_table = _oracle_set_saghai_table
Oracle.saghai_batch = _oracle_saghai_batch
Oracle.fint = _oracle_fint


# ---- DSS fragmentation functions (semi-inclusive kaons) ------------------------------------------------
def load_fdss_fixture():
    """tests/golden/fdss_kanlo.npz: the rows of the reference's fdss/KANLO.GRID, [34, 24, 9]."""
    return np.load(os.path.join(GOLDEN, "fdss_kanlo.npz"))["parton"]


def write_fdss_file(parton, path):
    with open(path, "w") as f:
        for row in np.asarray(parton).reshape(-1, 9):
            f.write("".join(f"{x:10.3E}" for x in row) + "\n")


def _oracle_set_fdss_table(self, parton):
    parton = np.ascontiguousarray(parton, np.float64)
    self._check(self.L.oracle_set_fdss_table(_p(parton)))


def _oracle_fdss_batch(self, ic, x, q2):
    x = np.ascontiguousarray(x, np.float64)
    q2 = np.ascontiguousarray(q2, np.float64)
    out = np.zeros((6, len(x)))
    self._check(self.L.oracle_fdss_batch(int(ic), C.c_int64(len(x)), _p(x), _p(q2), _p(out)))
    return out


Oracle.set_fdss_table = _oracle_set_fdss_table
Oracle.fdss_batch = _oracle_fdss_batch


# ---- A > 2 meson production: generate_em --------------------------------------------------------------
def load_he3_fixtures():
    """(pval, mprob) of he3.dat and the Benhar-type table of benharsf_3mod.dat (tests/golden/*.npz)."""
    z = np.load(os.path.join(GOLDEN, "pfermi_he3.npz"))
    return (z["pval"], z["mprob"]), np.load(os.path.join(GOLDEN, "benharsf_3mod.npz"))


def write_sf_file(z, path):
    """benharsf_*.dat layout: 'numPm numEm', then rows 'Pm Em S_p S_n dPm dEm' with Em running fastest."""
    with open(path, "w") as f:
        f.write(f"{len(z['pm'])} {len(z['em'])}\n")
        for i, pm in enumerate(z["pm"]):
            for j, em in enumerate(z["em"]):
                f.write(f"{float(pm)!r} {float(em)!r} {float(z['sf_proton'][i, j])!r} {float(z['sf_neutron'][i, j])!r} "
                        f"{float(z['dpm'][i])!r} {float(z['dem'][j])!r}\n")


def _oracle_set_sf_em_widths(self, dem):
    dem = np.ascontiguousarray(dem, np.float64)
    self._check(self.L.oracle_set_sf_em_widths(len(dem), _p(dem)))


def _oracle_generate_em_batch(self, seed, pm):
    pm = np.ascontiguousarray(pm, np.float64)
    out = np.zeros(len(pm))
    self._check(self.L.oracle_generate_em_batch(C.c_uint64(seed), C.c_int64(len(pm)), _p(pm), _p(out)))
    return out


Oracle.set_sf_em_widths = _oracle_set_sf_em_widths
Oracle.generate_em_batch = _oracle_generate_em_batch


# ---- end of the loop body on dumped vectors (twin of simc_b200_weight_batch) --------------------------
def _oracle_weight_inputs(self, cfg, first, n, seed):
    """What complete_recon_ev / complete_main read for tries [first, first+n) once montecarlo returned:
    (inp [44, n], valid [n])."""
    inp = np.zeros((44, n))
    valid = np.zeros(n, dtype=np.int32)
    self._check(self.L.oracle_weight_inputs(C.byref(cfg), C.c_int64(first), C.c_int64(n), C.c_uint64(seed), _p(inp), _p(valid)))
    return inp, valid.astype(bool)


def _oracle_weight_batch(self, cfg, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros((15, inp.shape[1]))
    self._check(self.L.oracle_weight_batch(C.byref(cfg), C.c_int64(inp.shape[1]), _p(inp), _p(out)))
    return out


Oracle.weight_inputs = _oracle_weight_inputs
Oracle.weight_batch = _oracle_weight_batch


def load_field_fixture():
    """simc_gfortran_b200/data/trg_field_map.npz (tools/make_fixtures.py: the reference's trg_field_map.dat): bz, br."""
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "trg_field_map.npz"))
    return z["bz"], z["br"]


def write_field_file(bz, br, path):
    """Rows 'z r Bz Br |B| ratio diff' like trg_field_map.dat (trgInit reads them list-directed)."""
    with open(path, "w") as f:
        for ir in range(51):
            for iz in range(51):
                k = ir * 51 + iz
                f.write("%8.3f%8.3f %.17g %.17g %.17g %.9f %12.1E\n" % (2.0 * iz, 2.0 * ir, bz[k], br[k], np.hypot(bz[k], br[k]), 1.0, 0.0))


def _oracle_set_field_map(self, bz, br, theta_e_deg, theta_p_deg):
    """trgInit; bz = br = None: the uniform 5 T test field."""
    if bz is None:
        self._check(self.L.oracle_set_field_map(None, None, C.c_double(theta_e_deg), C.c_double(theta_p_deg)))
        return
    bz = np.ascontiguousarray(bz, np.float64).ravel()
    br = np.ascontiguousarray(br, np.float64).ravel()
    self._check(self.L.oracle_set_field_map(_p(bz), _p(br), C.c_double(theta_e_deg), C.c_double(theta_p_deg)))


def _oracle_field_batch(self, spect, inp):
    inp = np.ascontiguousarray(inp, np.float64)
    out = np.zeros((6, inp.shape[1]))
    self._check(self.L.oracle_field_batch(int(spect), C.c_int64(inp.shape[1]), _p(inp), _p(out)))
    return out


def _oracle_field_at(self, spect, xyz):
    xyz = np.ascontiguousarray(xyz, np.float64)
    out = np.zeros((3, xyz.shape[1]))
    self._check(self.L.oracle_field_at(int(spect), C.c_int64(xyz.shape[1]), _p(xyz), _p(out)))
    return out


def _oracle_field_steps(self, spect, u0, E, dl, n_steps):
    u0 = np.ascontiguousarray(u0, np.float64)
    traj = np.zeros((n_steps, 6, u0.shape[1]))
    self._check(self.L.oracle_field_steps(int(spect), C.c_int64(u0.shape[1]), _p(u0), C.c_double(E), C.c_double(dl),
                                          int(n_steps), _p(traj)))
    return traj


Oracle.set_field_map = _oracle_set_field_map
Oracle.field_batch = _oracle_field_batch
Oracle.field_at = _oracle_field_at
Oracle.field_steps = _oracle_field_steps
