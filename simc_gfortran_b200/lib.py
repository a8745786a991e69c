"""ctypes binding of ``libsimc_b200.so`` (include/simc_b200.h).

Mirrors the C ABI one to one; the structures below repeat the header's layouts and are
checked against ``simc_b200_sizeof`` at load time.  Nothing here computes: a missing library
or a missing GPU is a loud error, never a fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ARM_HMS, ARM_SOS, ARM_HRSR, ARM_HRSL, ARM_SHMS = 1, 2, 3, 4, 5
WEIGHT_NIN, WEIGHT_NOUT = 44, 15
TRANSPORT_NIN, TRANSPORT_NOUT = 9, 12
EVENT_NREC = 60
RADC_NOUT = 26
NTUPLE_MAXCOL = 68
NHIST, H_PER_SET, NSTOP = 50, 8, 64
ABI_VERSION = 3

_HERE = os.path.dirname(os.path.abspath(__file__))


class SimcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsimc_b200 error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    # SIMC_B200_LIB: an alternative build of the same library (tools/perf_sweep.sh tries compile-time knobs)
    return os.environ.get("SIMC_B200_LIB") or os.path.join(_HERE, "libsimc_b200.so")


# ---- structures (same order as include/simc_b200.h) -------------------------------------
class Cut(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double)]


class Range(C.Structure):
    _fields_ = [("lo", C.c_double), ("hi", C.c_double)]


class ArmCuts(C.Structure):
    _fields_ = [("delta", Cut), ("yptar", Cut), ("xptar", Cut), ("z", Cut)]


class ArmLimits(C.Structure):
    _fields_ = [("delta", Cut), ("yptar", Cut), ("xptar", Cut), ("E", Cut)]


class EdgeArm(C.Structure):
    _fields_ = [("E", Cut), ("yptar", Cut), ("xptar", Cut)]


class Edge(C.Structure):
    _fields_ = [("e", EdgeArm), ("p", EdgeArm), ("Em", Cut), ("Pm", Cut), ("Mrec", Cut), ("Trec", Cut),
                ("Trec_struck", Cut)]


class GenLimits(C.Structure):
    _fields_ = [("e", ArmLimits), ("p", ArmLimits), ("sumEgen", Cut), ("Trec", Cut), ("xwid", C.c_double),
                ("ywid", C.c_double)]


class Spectrometer(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("P", "theta", "cos_th", "sin_th", "phi", "off_x", "off_y", "off_z",
                                          "off_xptar", "off_yptar")]


class Axis(C.Structure):
    _fields_ = [("min", C.c_double), ("bin", C.c_double)]


class Target(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "A", "Z", "N", "mass_amu", "M", "mrec_amu", "Mrec", "rho", "thick", "angle", "abundancy", "length",
        "zoffset", "X0", "X0_cm", "L1", "L2", "fr1", "fr2", "xoffset", "yoffset", "Coulomb_ave", "Coulomb_min",
        "Coulomb_max", "Coulomb_constant", "Mtar_struck", "Mrec_struck")] + [("fr_pattern", C.c_int32),
                                                                              ("can", C.c_int32)]


_INT_FLAGS = (
    "abi_version",
    "doing_phsp", "doing_hyd_elast", "doing_deuterium", "doing_heavy", "doing_eep",
    "doing_pion", "doing_kaon", "doing_delta", "doing_rho", "doing_semi",
    "doing_hydpi", "doing_deutpi", "doing_hepi",
    "doing_hydkaon", "doing_deutkaon", "doing_hekaon",
    "doing_hydsemi", "doing_deutsemi",
    "doing_semipi", "doing_semika", "do_fermi",
    "doing_hplus", "doing_decay",
    "which_pion", "which_kaon",
    "using_rad", "using_Eloss", "using_Coulomb", "correct_Eloss", "correct_raster",
    "mc_smear", "hard_cuts",
    "using_E_arm_montecarlo", "using_P_arm_montecarlo",
    "electron_arm", "hadron_arm",
    "using_HMScoll", "using_SHMScoll", "use_benhar_sf",
    "rad_flag", "extrad_flag", "intcor_mode", "use_expon", "use_offshell_rad",
)
_DBL_SCALARS = (
    "Mh", "Mh2", "Ebeam", "dEbeam", "Ebeam_vertex_ave",
    "dE_edge_test", "Egamma_gen_max", "ctau", "transparency", "drift_to_cal",
    "targ_Bangle", "targ_Bphi", "targ_pol", "sign_hadron",
    "etatzai", "Egamma_tot_max", "Egamma1_max", "Egamma2_max", "Egamma3_max", "Egamma_res_limit",
)


class RunConfig(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in _INT_FLAGS] +
                [("doing_tail", C.c_int32 * 3), ("hardwired_rad", C.c_int32), ("deForest_flag", C.c_int32),
                 ("doing_pizero", C.c_int32), ("pizero_ngamma", C.c_int32),
                 ("using_tgt_field", C.c_int32), ("pad_flags", C.c_int32)] +
                [(n, C.c_double) for n in _DBL_SCALARS] +
                [("gen", GenLimits), ("spec_e", Spectrometer), ("spec_p", Spectrometer),
                 ("cuts_Em", Cut), ("cuts_Pm", Cut), ("edge", Edge), ("VERTEXedge", Edge),
                 ("SPedge_e", ArmCuts), ("SPedge_p", ArmCuts),
                 ("slop_MC_e_used", C.c_double * 3), ("slop_MC_p_used", C.c_double * 3),
                 ("targ", Target), ("hist_axis", (Axis * H_PER_SET) * 3), ("w_ref", C.c_double)])

    def __init__(self, **kw):
        super().__init__(**kw)
        self.abi_version = ABI_VERSION


class Fixed128(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_int64), ("qexp", C.c_int32), ("pad", C.c_int32)]

    def value(self) -> float:
        v = (int(self.hi) << 64) + int(self.lo)
        return float(v) * 2.0 ** int(self.qexp) if abs(v) < (1 << 1000) else float("inf")

    def exact(self):
        """(integer, exponent): value = integer * 2**exponent."""
        return (int(self.hi) << 64) + int(self.lo), int(self.qexp)


class Accum(C.Structure):
    _fields_ = [("ntried", C.c_int64), ("nsuccess", C.c_int64), ("ncontribute", C.c_int64),
                ("npasscuts", C.c_int64), ("ncontribute_no_rad_proton", C.c_int64),
                ("wtcontribute", Fixed128), ("sum_sigcc", Fixed128),
                ("sumerr", Fixed128 * 8), ("sumerr2", Fixed128 * 8),
                ("hist_w", (Fixed128 * NHIST) * 6),
                ("hist_n", ((C.c_int64 * NHIST) * H_PER_SET) * 3),
                ("contrib", Range * 32), ("slop", Range * 8),
                ("stop", (C.c_int64 * NSTOP) * 2),
                ("transp_calls", (C.c_int64 * 48) * 2), ("unsupported", C.c_int64), ("nonfinite", C.c_int64)]


_lib = None


class Results(C.Structure):
    """simc_results (include/simc_b200.h): normalisation and resolutions of a finished run."""
    _fields_ = [("luminosity", C.c_double), ("genvol", C.c_double), ("normfac", C.c_double), ("yield_", C.c_double),
                ("central_sigcc_ave", C.c_double), ("nevent", C.c_int64), ("aveerr", C.c_double * 8), ("resol", C.c_double * 8)]


class ReportInfo(C.Structure):
    """simc_report_info (include/simc_b200.h): init-only values subroutine report prints."""
    _fields_ = [("ngen", C.c_int32), ("random_seed", C.c_int32), ("one_tail", C.c_int32),
                ("doing_pizero", C.c_int32), ("pizero_ngamma", C.c_int32), ("use_first_cer", C.c_int32), ("using_tgt_field", C.c_int32),
                ("doing_hyddelta", C.c_int32), ("doing_deutdelta", C.c_int32), ("doing_hedelta", C.c_int32),
                ("doing_hydrho", C.c_int32), ("doing_deutrho", C.c_int32), ("doing_herho", C.c_int32), ("pad", C.c_int32),
                ("charge_mC", C.c_double),
                ("Eloss_ave", C.c_double * 3), ("Eloss_min", C.c_double * 3), ("Eloss_max", C.c_double * 3),
                ("teff_ave", C.c_double * 3), ("teff_min", C.c_double * 3), ("teff_max", C.c_double * 3),
                ("musc_max", C.c_double * 3), ("musc_nsig_max", C.c_double), ("slop_total_Em_used", C.c_double),
                ("theory_file", C.c_char * 128)]


class Central(C.Structure):
    """simc_central: event_central as calculate_central fills it (simc.f:1143-1306)."""
    _fields_ = [("e_delta", C.c_double), ("e_xptar", C.c_double), ("e_yptar", C.c_double), ("p_delta", C.c_double),
                ("p_xptar", C.c_double), ("p_yptar", C.c_double),
                ("Q2", C.c_double), ("q", C.c_double), ("nu", C.c_double), ("Em", C.c_double), ("Pm", C.c_double),
                ("W", C.c_double), ("MM", C.c_double), ("sigcc", C.c_double),
                ("hardcorfac", C.c_double), ("etatzai", C.c_double), ("frac", C.c_double * 3), ("lambda_", C.c_double * 3),
                ("bt", C.c_double * 2), ("c_int", C.c_double * 4), ("c_ext", C.c_double * 4), ("c", C.c_double * 4),
                ("g_int", C.c_double), ("g_ext", C.c_double), ("g", C.c_double * 4)]


def load_library():
    """Loads libsimc_b200.so from the package directory; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SimcError(-100, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or `make -C simc_gfortran_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(path)
    L.simc_b200_last_error.restype = C.c_char_p
    L.simc_b200_last_error.argtypes = [C.c_void_p]
    L.simc_b200_stop_name.restype = C.c_char_p
    L.simc_b200_stop_name.argtypes = [C.c_int, C.c_int]
    L.simc_b200_stream.restype = C.c_void_p
    L.simc_b200_stream.argtypes = [C.c_void_p]
    L.simc_b200_launch_count.restype = C.c_int64
    L.simc_b200_launch_count.argtypes = [C.c_void_p]
    L.simc_b200_sizeof.restype = C.c_int64
    L.simc_b200_sizeof.argtypes = [C.c_int]
    L.simc_b200_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.simc_b200_destroy.argtypes = [C.c_void_p]
    L.simc_b200_destroy.restype = None
    L.simc_b200_set_mode.argtypes = [C.c_void_p, C.c_int]
    L.simc_b200_set_compiled_maps.argtypes = [C.c_void_p, C.c_int]
    L.simc_b200_weight_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.simc_b200_precompile_optics.argtypes = ([C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p,
                                               C.c_char_p, C.c_void_p, C.c_char_p, C.c_int])
    L.simc_b200_load_optics.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
    L.simc_b200_set_optics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p]
    L.simc_b200_optics_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    tb = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
          C.c_void_p]
    L.simc_b200_transport_batch.argtypes = tb
    L.simc_b200_transport_batch_device.argtypes = tb
    L.simc_b200_sync.argtypes = [C.c_void_p]
    L.simc_b200_set_field_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.simc_b200_load_field_file.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_field_batch.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_void_p, C.c_void_p]
    for name, args in (("simc_b200_accum_clear", [C.c_void_p, C.c_void_p]),
                       ("simc_b200_run", [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p]),
                       ("simc_b200_run_async", [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]),
                       ("simc_b200_fetch", [C.c_void_p, C.c_void_p]),
                       ("simc_b200_event_batch", [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p,
                                                  C.c_void_p])):
        if hasattr(L, name):
            getattr(L, name).argtypes = args
    L.simc_b200_config_from_deck.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p,
                                             C.c_int]
    L.simc_b200_config_from_deck_data.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_char_p, C.c_int]
    L.simc_b200_set_theory_table.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 8
    L.simc_b200_load_theory_file.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_set_maid_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.simc_b200_load_maid_file.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.simc_b200_set_saghai_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.simc_b200_load_saghai_files.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_read_saghai_file.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
    L.simc_b200_set_fdss_table.argtypes = [C.c_void_p, C.c_void_p]
    L.simc_b200_load_fdss_file.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_set_sf_em_widths.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.simc_b200_normalise.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    L.simc_b200_report_info_from_deck.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
    L.simc_b200_central_event.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.simc_b200_write_geni.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    L.simc_b200_format_real.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_char_p, C.c_int]
    L.simc_b200_write_gen.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    L.simc_b200_write_hist.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p]
    L.simc_b200_ntuple_tags.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.simc_b200_ntuple_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    L.simc_b200_ntuple_append.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.simc_b200_ntuple_close.argtypes = [C.c_void_p]
    L.simc_b200_set_batch.argtypes = [C.c_void_p, C.c_int64]
    L.simc_b200_radc_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.simc_b200_set_pfermi_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.simc_b200_load_pfermi_file.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_set_cteq5_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.simc_b200_load_cteq5_file.argtypes = [C.c_void_p, C.c_char_p]
    L.simc_b200_semi_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.simc_b200_stage_times.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.simc_b200_device_accum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.simc_b200_reduce_gathered.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.simc_b200_accum_merge.argtypes = [C.c_void_p, C.c_void_p]
    L.simc_b200_fp64_peak.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.simc_b200_log_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    if hasattr(L, "simc_b200_event_field_name"):
        L.simc_b200_event_field_name.restype = C.c_char_p
        L.simc_b200_event_field_name.argtypes = [C.c_int]
    # layout check: the ctypes mirrors must have the C sizes
    for which, typ in ((0, RunConfig), (1, Accum)):
        c_size = L.simc_b200_sizeof(which)
        if c_size != C.sizeof(typ):
            raise SimcError(-101, f"ABI mismatch: sizeof({typ.__name__}) is {C.sizeof(typ)} in Python, {c_size} in C")
    _lib = L
    return L


def config_from_deck(deck_path: str, extra_deck_dir: str | None = None, data_dir: str | None = None):
    """(RunConfig, ngen, charge_mC) from a CTP deck; host only, no GPU needed.  data_dir: where the reference's
    working-directory data files live (h2.theory, c12.theory, ... for D(e,e'p) / A(e,e'p) without use_benhar_sf)."""
    L = load_library()
    cfg = RunConfig()
    ngen = C.c_int32()
    charge = C.c_double()
    err = C.create_string_buffer(512)
    rc = L.simc_b200_config_from_deck_data(deck_path.encode(), (extra_deck_dir or os.path.dirname(deck_path)).encode(),
                                           data_dir.encode() if data_dir else None,
                                           C.byref(cfg), C.byref(ngen), C.byref(charge), err, 512)
    if rc != 0:
        raise SimcError(rc, err.value.decode())
    return cfg, ngen.value, charge.value


def read_saghai_file(path: str, which: int) -> np.ndarray:
    """The library's reader of saghai_proton.dat (which = 0) / saghai_sigma0.dat (1), host only: float32
    [12, 19, n_q2, n_s] (the Fortran arrays' memory order)."""
    L = load_library()
    n1, n2 = (10, 11) if which == 0 else (20, 10)
    tbl = np.zeros((12, 19, n2, n1), dtype=np.float32)
    msg = C.create_string_buffer(512)
    rc = L.simc_b200_read_saghai_file(path.encode(), int(which), _ptr(tbl), msg, 512)
    if rc != 0:
        raise SimcError(rc, msg.value.decode())
    return tbl


def accum_merge(into: Accum, other: Accum) -> Accum:
    """into += other, exactly (simc_b200_accum_merge): counters and 128-bit sums add, ranges widen.  Host only."""
    L = load_library()
    rc = L.simc_b200_accum_merge(C.byref(into), C.byref(other))
    if rc != 0:
        raise SimcError(rc, "simc_b200_accum_merge: fixed-point sums on different quanta (different w_ref)")
    return into


def normalise(cfg: RunConfig, acc, ngen: int, charge_mC: float) -> Results:
    """simc.f:94-101, 366-432: luminosity, generation volume, normfac, normalised yield, resolutions.  Host only.
    ngen: the deck's ngen -- negative: that many tries, every try counts as an event (nevent = ntried); positive:
    until that many successes (nevent = successes), simc.f:346-350."""
    L = load_library()
    r = Results()
    rc = L.simc_b200_normalise(C.byref(cfg), C.byref(acc), int(ngen), float(charge_mC), C.byref(r))
    if rc:
        raise SimcError(rc, "simc_b200_normalise")
    return r


def report_info_from_deck(deck_path: str, extra_deck_dir: str | None = None, data_dir: str | None = None) -> ReportInfo:
    L = load_library()
    info = ReportInfo()
    err = C.create_string_buffer(512)
    extra = (extra_deck_dir or os.path.dirname(os.path.abspath(deck_path))).encode()
    rc = L.simc_b200_report_info_from_deck(deck_path.encode(), extra, data_dir.encode() if data_dir else None, C.byref(info), err, 512)
    if rc:
        raise SimcError(rc, err.value.decode())
    return info


def central_event(cfg: RunConfig, info: ReportInfo, sim=None) -> Central:
    """calculate_central (simc.f:1143-1306).  sim: a Simc handle with the run's tables for central%sigcc (GPU); without
    it the kinematics and the radiative constants only (host)."""
    L = load_library()
    c = Central()
    rc = L.simc_b200_central_event(sim.h if sim is not None else None, C.byref(cfg), C.byref(info), C.byref(c))
    if rc:
        raise SimcError(rc, "simc_b200_central_event")
    return c


def write_reports(base: str, cfg: RunConfig, info: ReportInfo, central: Central, acc, res: Results, t1: str, t2: str):
    """<base>.geni, <base>.gen, <base>.hist in the reference's layout (simc.f:446-1139)."""
    L = load_library()
    for rc, what in ((L.simc_b200_write_geni((base + ".geni").encode(), C.byref(cfg), C.byref(acc)), "geni"),
                     (L.simc_b200_write_gen((base + ".gen").encode(), C.byref(cfg), C.byref(acc)), "gen"),
                     (L.simc_b200_write_hist((base + ".hist").encode(), C.byref(cfg), C.byref(info), C.byref(central), C.byref(acc),
                                             C.byref(res), t1.encode(), t2.encode()), "hist")):
        if rc:
            raise SimcError(rc, "simc_b200_write_" + what)


def ntuple_tags(cfg: RunConfig):
    L = load_library()
    buf = (C.c_char * 17 * 80)()
    n = L.simc_b200_ntuple_tags(C.byref(cfg), buf, 80)
    if n < 0:
        raise SimcError(n, "simc_b200_ntuple_tags")
    return [bytes(buf[i]).split(b"\0")[0].decode() for i in range(n)]


def write_ntuple_file(cfg: RunConfig, path: str, rows: np.ndarray):
    """rows[n, n_cols] -> the reference's unformatted .bin ntuple (NtupleInit.f, results_write.f:264-266).  Host only."""
    L = load_library()
    h = C.c_void_p()
    rc = L.simc_b200_ntuple_open(C.byref(cfg), path.encode(), C.byref(h))
    if rc:
        raise SimcError(rc, "simc_b200_ntuple_open")
    full = np.zeros((len(rows), NTUPLE_MAXCOL))
    full[:, :rows.shape[1]] = rows
    rc = L.simc_b200_ntuple_append(h, _ptr(full), len(full))
    rc2 = L.simc_b200_ntuple_close(h)
    if rc or rc2:
        raise SimcError(rc or rc2, "simc_b200_ntuple_append/close")


def read_ntuple_file(path: str):
    """Reads a Fortran unformatted sequential ntuple the way util/root_tree/make_root_tree.f does:
    -> (tags, rows[n, n_cols])."""
    raw = np.fromfile(path, dtype=np.uint8)
    pos = 0

    def record():
        nonlocal pos
        n = int(raw[pos:pos + 4].view(np.int32)[0])
        body = raw[pos + 4:pos + 4 + n]
        assert int(raw[pos + 4 + n:pos + 8 + n].view(np.int32)[0]) == n, "record markers disagree"
        pos += 8 + n
        return body
    size = int(record().view(np.int32)[0])
    tags = [bytes(record()).decode().rstrip() for _ in range(size)]
    rest = raw[pos:]
    assert len(rest) % (16 * size) == 0
    recs = rest.reshape(-1, 16)
    assert np.all(recs[:, :4].copy().view(np.int32) == 8) and np.all(recs[:, 12:].copy().view(np.int32) == 8)
    vals = recs[:, 4:12].copy().view(np.float64).reshape(-1, size)
    return tags, vals


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def precompile_optics(t, strict: bool = True, cache_dir: str | None = None, dump_source: str | None = None):
    """Runs the map compiler on optics tables `t` (no GPU needed) and leaves the cubin in the cache.
    Returns (stretches, source bytes, cubin bytes, was_cached)."""
    L = load_library()
    cs = np.ascontiguousarray(t.class_start, dtype=np.int32)
    fc = np.ascontiguousarray(t.fwd_coeff, dtype=np.float64)
    fe = np.ascontiguousarray(t.fwd_expon, dtype=np.int8)
    ln = np.ascontiguousarray(t.length_cm, dtype=np.float64)
    rc = np.ascontiguousarray(t.rec_coeff, dtype=np.float64)
    re_ = np.ascontiguousarray(t.rec_expon, dtype=np.int8)
    info = (C.c_int64 * 4)()
    msg = C.create_string_buffer(8192)
    r = L.simc_b200_precompile_optics(int(t.arm), len(cs) - 1, _ptr(cs), _ptr(fc), _ptr(fe), _ptr(ln), len(rc), _ptr(rc), _ptr(re_),
                                      1 if strict else 0, cache_dir.encode() if cache_dir else None,
                                      dump_source.encode() if dump_source else None, info, msg, 8192)
    if r != 0:
        raise SimcError(r, msg.value.decode())
    return tuple(int(x) for x in info)


class Simc:
    """One handle = one GPU + one run configuration (simc_b200_create)."""

    def __init__(self, cfg: RunConfig | None = None, device: int = 0, mode: str | None = None, compiled_maps: bool | None = None):
        self.L = load_library()
        self.h = C.c_void_p()
        self.cfg = cfg
        rc = self.L.simc_b200_create(C.byref(cfg) if cfg is not None else None, device, C.byref(self.h))
        if rc != 0:
            raise SimcError(rc, (self.L.simc_b200_last_error(None) or b"").decode())
        self.mode = os.environ.get("SIMC_B200_MODE", "strict")
        if mode is not None:
            self.set_mode(mode)
        if compiled_maps is not None:
            self.set_compiled_maps(compiled_maps)

    def _check(self, rc: int):
        if rc != 0:
            raise SimcError(rc, (self.L.simc_b200_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.simc_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mode(self, mode: str):
        assert mode in ("strict", "fast")
        self.mode = mode
        self._check(self.L.simc_b200_set_mode(self.h, 1 if mode == "strict" else 0))

    def weight_batch(self, inp: np.ndarray) -> np.ndarray:
        """complete_recon_ev + complete_main + pass_cuts on dumped vectors: inp [WEIGHT_NIN, n] -> [WEIGHT_NOUT, n]
        (column order: include/simc_b200.h)."""
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        assert inp.ndim == 2 and inp.shape[0] == WEIGHT_NIN
        n = inp.shape[1]
        out = np.zeros((WEIGHT_NOUT, n))
        self._check(self.L.simc_b200_weight_batch(self.h, n, _ptr(inp), _ptr(out)))
        return out

    def set_compiled_maps(self, on: bool):
        """Generated straight-line kernels for the RNG-free stretches of the arm programs (default on); off = the
        record interpreter everywhere."""
        self._check(self.L.simc_b200_set_compiled_maps(self.h, 1 if on else 0))

    # ---- optics
    def load_optics(self, arm: int, forward_path: str, recon_path: str):
        self._check(self.L.simc_b200_load_optics(self.h, arm, forward_path.encode(), recon_path.encode()))

    def set_optics(self, t):
        cs = np.ascontiguousarray(t.class_start, dtype=np.int32)
        fc = np.ascontiguousarray(t.fwd_coeff, dtype=np.float64)
        fe = np.ascontiguousarray(t.fwd_expon, dtype=np.int8)
        ln = np.ascontiguousarray(t.length_cm, dtype=np.float64)
        rc_ = np.ascontiguousarray(t.rec_coeff, dtype=np.float64)
        re_ = np.ascontiguousarray(t.rec_expon, dtype=np.int8)
        self._check(self.L.simc_b200_set_optics(self.h, t.arm, len(cs) - 1, _ptr(cs), _ptr(fc), _ptr(fe), _ptr(ln),
                                                len(rc_), _ptr(rc_), _ptr(re_)))

    def optics_info(self, arm: int) -> dict:
        info = np.zeros(8, dtype=np.int64)
        self._check(self.L.simc_b200_optics_info(self.h, arm, _ptr(info)))
        keys = ("n_classes", "fwd_terms", "fwd_nonzero", "rec_terms", "n_rec_words", "n_coef", "n_ops")
        return dict(zip(keys, (int(x) for x in info[:7])))

    # ---- Benhar spectral function, A(e,e'p)
    def set_sf_table(self, pm, em, sf):
        """pm[n_pm], em[n_em] bin centres; sf[n_pm, n_em] proton or neutron strength (un-normalised)."""
        pm = np.ascontiguousarray(pm, dtype=np.float64)
        em = np.ascontiguousarray(em, dtype=np.float64)
        sf = np.ascontiguousarray(sf, dtype=np.float64)
        assert sf.shape == (len(pm), len(em))
        self._check(self.L.simc_b200_set_sf_table(self.h, len(pm), len(em), _ptr(pm), _ptr(em), _ptr(sf)))

    def set_sf_em_widths(self, dem):
        """Widths of the table's Em bins (generate_em: pion/kaon production from A > 2)."""
        dem = np.ascontiguousarray(dem, dtype=np.float64)
        self._check(self.L.simc_b200_set_sf_em_widths(self.h, len(dem), _ptr(dem)))

    def load_sf_file(self, path: str, proton: bool = True):
        self._check(self.L.simc_b200_load_sf_file(self.h, path.encode(), 1 if proton else 0))

    # ---- DSS fragmentation functions (semi-inclusive kaons)
    def set_fdss_table(self, parton):
        """parton[34, 24, 9]: the rows of a fdss/*.GRID file in reading order."""
        parton = np.ascontiguousarray(parton, dtype=np.float64)
        assert parton.shape == (34, 24, 9)
        self._check(self.L.simc_b200_set_fdss_table(self.h, _ptr(parton)))

    def load_fdss_file(self, path: str):
        self._check(self.L.simc_b200_load_fdss_file(self.h, path.encode()))

    # ---- MAID-2007 table of peepi's low-W branch
    def set_maid_table(self, ipi: int, tbl):
        """ipi = 3 (pi+ n) or 4 (pi- p); tbl[25, 46, 6, 4] (see include/simc_b200.h)."""
        tbl = np.ascontiguousarray(tbl, dtype=np.float64)
        assert tbl.shape == (25, 46, 6, 4)
        self._check(self.L.simc_b200_set_maid_table(self.h, int(ipi), _ptr(tbl)))

    def set_saghai_table(self, which: int, tbl):
        """Saghai amplitude tables of peeK's ntuple column sigcm1: which = 0 K+ Lambda [12][19*11*10], 1 K+ Sigma0
        [12][19*10*20], float32 in the Fortran arrays' memory order (simc_b200_set_saghai_table)."""
        tbl = np.ascontiguousarray(tbl, dtype=np.float32)
        assert tbl.size == 12 * 19 * (110 if which == 0 else 200)
        self._check(self.L.simc_b200_set_saghai_table(self.h, int(which), _ptr(tbl)))

    def load_saghai_files(self, directory: str):
        self._check(self.L.simc_b200_load_saghai_files(self.h, directory.encode()))

    def load_maid_file(self, ipi: int, path: str):
        self._check(self.L.simc_b200_load_maid_file(self.h, int(ipi), path.encode()))

    # ---- independent-particle spectral function (h2.theory, c12.theory, ...): D(e,e'p), A(e,e'p) without Benhar
    def set_theory_table(self, t):
        """t: dict with n_shells, absorption, e_fermi, nprot, em, emsig, bs_norm, n_pm, pm_first, pm_bin, rho."""
        a = {k: np.ascontiguousarray(t[k], dtype=np.float64) for k in ("nprot", "em", "emsig", "bs_norm", "pm_first",
                                                                       "pm_bin", "rho")}
        n_pm = np.ascontiguousarray(t["n_pm"], dtype=np.int32)
        self._check(self.L.simc_b200_set_theory_table(self.h, int(t["n_shells"]), float(t["absorption"]),
                                                      float(t["e_fermi"]), _ptr(a["nprot"]), _ptr(a["em"]),
                                                      _ptr(a["emsig"]), _ptr(a["bs_norm"]), _ptr(n_pm),
                                                      _ptr(a["pm_first"]), _ptr(a["pm_bin"]), _ptr(a["rho"])))

    def load_theory_file(self, path: str):
        self._check(self.L.simc_b200_load_theory_file(self.h, path.encode()))

    # ---- semi-inclusive production: momentum distribution (deut.dat) and CTEQ5 parton distributions
    def set_pfermi_table(self, pval, mprob):
        pval = np.ascontiguousarray(pval, dtype=np.float64)
        mprob = np.ascontiguousarray(mprob, dtype=np.float64)
        assert pval.shape == mprob.shape and pval.ndim == 1
        self._check(self.L.simc_b200_set_pfermi_table(self.h, len(pval), _ptr(pval), _ptr(mprob)))

    def load_pfermi_file(self, path: str):
        self._check(self.L.simc_b200_load_pfermi_file(self.h, path.encode()))

    def set_cteq5_table(self, t):
        """t: dict with nx, nt, nfmx, lam, qini, qmax, xmin, xv[nx+1], qv[nt+1], upd (the fields of a cteq5*.tbl)."""
        xv = np.ascontiguousarray(t["xv"], dtype=np.float64)
        qv = np.ascontiguousarray(t["qv"], dtype=np.float64)
        upd = np.ascontiguousarray(t["upd"], dtype=np.float64)
        nx, nt, nfmx = int(t["nx"]), int(t["nt"]), int(t["nfmx"])
        assert len(xv) == nx + 1 and len(qv) == nt + 1 and len(upd) == (nx + 1) * (nt + 1) * (nfmx + 3)
        self._check(self.L.simc_b200_set_cteq5_table(self.h, nx, nt, nfmx, float(t["lam"]), float(t["qini"]),
                                                     float(t["qmax"]), float(t["xmin"]), _ptr(xv), _ptr(qv), _ptr(upd)))

    def load_cteq5_file(self, path: str):
        self._check(self.L.simc_b200_load_cteq5_file(self.h, path.encode()))

    # ---- field of the polarised target (trg_track.f)
    def set_field_map(self, bz=None, br=None):
        """bz, br: 51 x 51 nodes in the file's reading order (simc_b200_set_field_map); None: the uniform 5 T test field."""
        if bz is None:
            self._check(self.L.simc_b200_set_field_map(self.h, None, None))
            return
        bz = np.ascontiguousarray(bz, dtype=np.float64).ravel()
        br = np.ascontiguousarray(br, dtype=np.float64).ravel()
        assert bz.size == br.size == 51 * 51
        self._check(self.L.simc_b200_set_field_map(self.h, _ptr(bz), _ptr(br)))

    def load_field_file(self, path: str):
        self._check(self.L.simc_b200_load_field_file(self.h, path.encode()))

    def field_batch(self, spect: int, theta_deg: float, inp: np.ndarray) -> np.ndarray:
        """track_from_tgt on rows (x, y, z, dx, dy, mom, mass) -> (x, y, z, dx, dy, ok)."""
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        assert inp.ndim == 2 and inp.shape[0] == 7
        out = np.zeros((6, inp.shape[1]))
        self._check(self.L.simc_b200_field_batch(self.h, int(spect), float(theta_deg), inp.shape[1], _ptr(inp), _ptr(out)))
        return out

    def semi_batch(self, inp: np.ndarray) -> np.ndarray:
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        assert inp.ndim == 2 and inp.shape[0] == 16
        out = np.zeros((16, inp.shape[1]))
        self._check(self.L.simc_b200_semi_batch(self.h, inp.shape[1], _ptr(inp), _ptr(out)))
        return out

    # ---- ntuple rows (results_ntu_write)
    def ntuple_batch(self, first_try: int, n: int, seed: int):
        """-> (rows[n_rows, n_cols], try_of_row[n_rows]) for the contributing tries of the range."""
        rows = np.zeros((max(n, 1), NTUPLE_MAXCOL), dtype=np.float64)
        tries = np.zeros(max(n, 1), dtype=np.int64)
        nc, nr = C.c_int32(0), C.c_int64(0)
        self._check(self.L.simc_b200_ntuple_batch(self.h, C.c_int64(first_try), C.c_int64(n), C.c_uint64(seed), _ptr(rows),
                                                  C.byref(nc), C.byref(nr), _ptr(tries)))
        return rows[:nr.value, :nc.value].copy(), tries[:nr.value].copy()

    # ---- single-arm batch (host buffers)
    def transport_batch(self, arm: int, inp: np.ndarray, seed: int, ms_flag=True, wcs_flag=True, decay_flag=False,
                        using_coll=False):
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        assert inp.ndim == 2 and inp.shape[0] == TRANSPORT_NIN
        n = inp.shape[1]
        out = np.zeros((TRANSPORT_NOUT, n), dtype=np.float64)
        flags = np.zeros(n, dtype=np.int32)
        self._check(self.L.simc_b200_transport_batch(self.h, arm, n, _ptr(inp), seed, int(ms_flag), int(wcs_flag),
                                                     int(decay_flag), int(using_coll), _ptr(out), _ptr(flags)))
        return out, flags

    def transport_batch_device(self, arm: int, n: int, d_in: int, seed: int, d_out: int, d_flags: int, ms_flag=True,
                               wcs_flag=True, decay_flag=False, using_coll=False):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on the handle's stream."""
        self._check(self.L.simc_b200_transport_batch_device(self.h, arm, n, C.c_void_p(d_in), seed, int(ms_flag),
                                                            int(wcs_flag), int(decay_flag), int(using_coll),
                                                            C.c_void_p(d_out), C.c_void_p(d_flags)))

    def sync(self):
        self._check(self.L.simc_b200_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.L.simc_b200_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.simc_b200_launch_count(self.h))

    def stop_name(self, arm: int, code: int) -> str:
        return self.L.simc_b200_stop_name(arm, code).decode()

    # ---- the loop
    def accum_clear(self, acc: Accum | None = None) -> Accum:
        acc = acc if acc is not None else Accum()
        self._check(self.L.simc_b200_accum_clear(self.h, C.byref(acc)))
        return acc

    def run(self, first_try: int, n_tries: int, seed: int, acc: Accum) -> Accum:
        self._check(self.L.simc_b200_run(self.h, first_try, n_tries, seed, C.byref(acc)))
        return acc

    def set_batch(self, tries_per_batch: int):
        self._check(self.L.simc_b200_set_batch(self.h, tries_per_batch))

    def stage_times(self, enable: bool = True):
        """(ms[4], launches[4]) of generate / hadron arm / electron arm / finish since the last call."""
        ms = (C.c_double * 4)()
        nl = (C.c_int64 * 4)()
        self._check(self.L.simc_b200_stage_times(self.h, int(enable), ms, nl))
        return list(ms), list(nl)

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.L.simc_b200_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def log_batch(self, x: np.ndarray):
        """The device's log and log10 (fastlog.cuh) of x."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        a, b = np.zeros_like(x), np.zeros_like(x)
        self._check(self.L.simc_b200_log_batch(self.h, x.size, _ptr(x), _ptr(a), _ptr(b)))
        return a, b

    def run_async(self, first_try: int, n_tries: int, seed: int):
        self._check(self.L.simc_b200_run_async(self.h, first_try, n_tries, seed))

    def fetch(self, acc: Accum) -> Accum:
        self._check(self.L.simc_b200_fetch(self.h, C.byref(acc)))
        return acc

    def device_accum(self):
        """(device pointer, number of 64-bit words) of the handle's accumulator block (simc_b200_device_accum)."""
        ptr, n = C.c_void_p(), C.c_int64()
        self._check(self.L.simc_b200_device_accum(self.h, C.byref(ptr), C.byref(n), None, None))
        return int(ptr.value), int(n.value)

    def reduce_gathered(self, d_gathered: int, n_ranks: int):
        """Folds n_ranks accumulator blocks (device pointer, rank after rank) into the handle's own; asynchronous on
        the handle's stream (simc_b200_reduce_gathered)."""
        self._check(self.L.simc_b200_reduce_gathered(self.h, C.c_void_p(d_gathered), n_ranks))

    def event_batch(self, first_try: int, n: int, seed: int):
        rec = np.zeros((EVENT_NREC, n), dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        self._check(self.L.simc_b200_event_batch(self.h, first_try, n, seed, _ptr(rec), _ptr(status)))
        return rec, status

    def radc_batch(self, inp: np.ndarray) -> np.ndarray:
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        assert inp.ndim == 2 and inp.shape[0] == 16
        out = np.zeros((RADC_NOUT, inp.shape[1]))
        self._check(self.L.simc_b200_radc_batch(self.h, inp.shape[1], _ptr(inp), _ptr(out)))
        return out

    def event_field_names(self):
        return [self.L.simc_b200_event_field_name(k).decode() for k in range(EVENT_NREC)]
