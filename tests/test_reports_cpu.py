"""The end-of-run text files in the reference's layout (simc.f:446-1139): <base>.geni (list-directed STOP counters),
<base>.gen (histograms in 3(1x,2(e11.4,1x))) and <base>.hist (subroutine report).  The accumulators come from the CPU
oracle's loop (same simc_accum layout); the writers are the product's host code.  No reference output exists to diff
against (no Fortran compiler), so the layout is checked against the Fortran format semantics line by
line: field widths, justification, the E-format's 0.ddd mantissa, list-directed I12 integers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import (central_event, config_from_deck, load_library, load_optics_fixture, report_info_from_deck,
                                write_reports)
from simc_gfortran_b200.lib import normalise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fmt(kind, w, d, v):
    L = load_library()
    out = C.create_string_buffer(64)
    assert L.simc_b200_format_real(ord(kind), w, d, float(v), out, 64) == 0
    return out.value.decode()


def test_fortran_edit_descriptors():
    # Ew.d: mantissa 0.ddd, two-digit exponent; the optional leading zero is dropped when the field is too narrow
    assert fmt("E", 12, 6, 1.0) == "0.100000E+01"
    assert fmt("E", 12, 6, -1.0) == "-.100000E+01"
    assert fmt("E", 11, 4, 0.0) == " 0.0000E+00"
    assert fmt("E", 11, 4, 12345.678) == " 0.1235E+05"
    assert fmt("E", 11, 4, -0.00098765) == "-0.9877E-03"
    assert fmt("E", 16, 8, 18.785375382) == "  0.18785375E+02"
    assert fmt("E", 16, 6, 9.99999999e-10) == "    0.100000E-08"
    assert fmt("E", 11, 4, 1.0e100) == " 0.1000+101"            # three-digit exponents lose the E
    # Fw.d: right-justified, sign kept on values that round to zero, asterisks on overflow
    assert fmt("F", 15, 4, 8800.0) == "      8800.0000"
    assert fmt("F", 10, 3, -0.0004) == "    -0.000"
    assert fmt("F", 6, 1, 123456.0) == "******"
    assert fmt("F", 5, 3, 0.5) == "0.500"
    assert fmt("F", 4, 3, 0.5) == ".500"
    assert fmt("F", 12, 5, -1.5) == "    -1.50000"


@pytest.fixture(scope="module")
def run(oracle_with_optics, tmp_path_factory):
    deck = os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")
    cfg, ngen, charge = config_from_deck(deck)
    acc = oracle_with_optics.run(cfg, 0, 30000, 11, threads=8)
    info = report_info_from_deck(deck)
    info.random_seed = 11
    info.ngen = -30000
    central = central_event(cfg, info)                  # host part only: kinematics and radiative constants
    res = normalise(cfg, acc, -30000, charge)
    base = str(tmp_path_factory.mktemp("reports") / "c1")
    write_reports(base, cfg, info, central, acc, res, "Sat Oct 18 05:22:11 2026\n", "Sat Oct 18 05:22:12 2026\n")
    return cfg, acc, info, central, res, base


def test_geni_file(run):
    cfg, acc, info, central, res, base = run
    lines = open(base + ".geni").read().split("\n")
    # HMS is the electron arm, SHMS the hadron arm: HMS block first (simc.f:447), then SHMS (simc.f:508)
    assert lines[0] == " HMS Trials:           " + "%12d" % acc.stop[0][0]
    assert lines[6] == " Events reaching hut   " + "%12d" % acc.stop[0][2]
    assert lines[8] == " Successes             " + "%12d" % acc.stop[0][1]
    assert lines[9] == ""
    assert lines[10] == " SHMS Trials:          " + "%12d" % acc.stop[1][0]
    assert lines[11].startswith(" HB phys entrance/mag entr/mag exit/phys exit  ") and len(lines[11]) == 47 + 4 * 12
    shms = [l for l in lines if l.startswith(" Successes")]
    assert shms[1] == " Successes             " + "%12d" % acc.stop[1][1]
    q1 = [l for l in lines if l.startswith(" Q1 entrance/mid/exit  ")][0]
    assert len(q1) == 23 + 3 * 12 and [int(x) for x in q1[23:].split()] == [acc.stop[0][2 + 4], acc.stop[0][2 + 5], acc.stop[0][2 + 6]]
    assert acc.stop[0][0] >= acc.stop[0][2] >= acc.stop[0][1]


def test_gen_file(run):
    cfg, acc, info, central, res, base = run
    lines = open(base + ".gen").read().split("\n")
    titles = [" E arm Experimental Target Distributions:", " P arm Experimental Target Distributions:",
              " Distributions of Contributing E arm Events:", " Distributions of Contributing P arm Events:",
              " Original E arm Events:", " Original P arm Events:", " Original Em/Pm distributions:"]
    pos = [lines.index(t) for t in titles]
    assert pos == [k * 52 for k in range(7)]
    assert lines[1] == "       delta     EXPERIM       yptar     EXPERIM       xptar     EXPERIM"
    assert lines[2 * 52 + 1] == "       delta     CONTRIB      yuptar     CONTRIB       xptar     CONTRIB"
    assert lines[6 * 52 + 1] == "          Em      ORIGIN          Pm      ORIGIN"
    efield = re.compile(r"^[ -]0\.\d{4}E[+-]\d{2}$")
    for k in range(6):
        for i in range(50):
            l = lines[pos[k] + 2 + i]
            assert len(l) == 3 * 25, (k, i, l)
            fields = [l[25 * j + 1: 25 * j + 12] for j in range(3)] + [l[25 * j + 13: 25 * j + 24] for j in range(3)]
            assert all(efield.match(x) for x in fields), l
    # block 3 (contributing E arm events): the counts of the gen histograms, bin centres of the gen axes
    ax = cfg.hist_axis[1][0]
    for i in (0, 17, 49):
        l = lines[pos[2] + 2 + i]
        assert float(l[1:12]) == pytest.approx(ax.min + (i + 0.5) * ax.bin, rel=6e-4, abs=1e-12)
        assert float(l[13:24]) == pytest.approx(acc.hist_n[1][0][i], rel=6e-4)
    # block 1: RECON weighted sums
    for i in (10, 25, 40):
        l = lines[pos[0] + 2 + i]
        assert float(l[13:24]) == pytest.approx(acc.hist_w[0][i].value(), rel=6e-4, abs=1e-30)
    # block 5 prints the gen buffer in its delta column (simc.f:585, as written) and geni in the others
    l = lines[pos[4] + 2 + 25]
    assert float(l[13:24]) == pytest.approx(acc.hist_n[1][0][25], rel=6e-4)
    assert float(l[38:49]) == pytest.approx(acc.hist_n[2][1][25], rel=6e-4)
    # the Em/Pm block: four fields per line
    l = lines[pos[6] + 2 + 3]
    assert len(l) == 1 + 4 * 12


def test_hist_file(run):
    cfg, acc, info, central, res, base = run
    text = open(base + ".hist").read()
    lines = text.split("\n")
    assert lines[0] == "" and lines[1] == " BEGIN Time: Sat Oct 18 05:22:11 2026      " and lines[2].startswith(" END Time:   Sat Oct 18")
    assert lines[3] == " KINEMATICS:"
    assert lines[4] == "               ****--------  H(e,e'p)  --------****"
    assert lines[5] == "                Ebeam =       8800.0000         MeV"
    assert lines[6] == "           (dE/E)beam =          0.0005  (full wid)"
    assert "           fr_pattern =               1   1=square,2=circ" in lines
    assert " " * 29 + "____E arm____    ____P arm____" in lines          # (9x,18x,2(a15,2x)): a15 right-justifies 13 characters
    assert "                angle =         25.9000          22.7300      deg" in lines
    assert "             momentum =       4531.0000        5122.0000    MeV/c" in lines
    k = lines.index(' ' + '                      VALUES FOR "CENTRAL" EVENT:')
    assert lines[k + 4].startswith("                         Q2 = ") and lines[k + 4].endswith("  (GeV/c)^2")
    assert float(lines[k + 4][30:45]) == pytest.approx(central.Q2 / 1e6, abs=6e-5)
    # target block: 9911 format(2x,2(5x,a10,' = ',e12.6,1x,a5))
    assert "                A = 0.100000E+01                    Z = 0.100000E+01      " in lines
    assert ("             mass = " + fmt("E", 12, 6, cfg.targ.mass_amu) + "   amu           mass = " + fmt("E", 12, 6, cfg.targ.M) + "   MeV") in lines
    assert "                   __ave__         __lo__         __hi__" in lines
    t = [l for l in lines if l.startswith("      Eloss_beam")][0]
    assert len(t) == 1 + 15 + 45 + 2 + 6 and t.endswith("     MeV")
    assert float(t[16:31]) == pytest.approx(info.Eloss_ave[0], abs=6e-6)
    m = [l for l in lines if l.startswith("   musc_nsig_max")][0]
    assert m == "   musc_nsig_max                3.50000        "
    # flags: 5x,3(2x,a19,'=',l2) -- Aw right-justifies, or keeps the leftmost w characters of a longer name
    def A(t, w):
        return t[:w] if len(t) >= w else t.rjust(w)

    def flags(*items):
        return "     " + "".join("  " + A(n, 19) + "=" + (v if isinstance(v, str) else ("%2d" % v)) for n, v in items)
    assert flags(("doing_eep", " T"), ("doing_kaon", " F"), ("doing_pion", " F")) in lines
    assert flags(("mc_smear", " T"), ("electron_arm", 1), ("hadron_arm", 5)) in lines
    assert flags(("using_E_arm_montecarlo", " T"), ("using_P_arm_montecarlo", " T"), ("use_benhar_sf", " F")) in lines
    assert "using_E_arm_monteca= T" in text
    assert "       " + A("use_first_cer", 19) + "= T" in lines
    assert "       " + A("ctau", 11) + "=" + fmt("F", 10, 3, cfg.ctau) + "  cm" in lines
    # counters
    assert "            Ngen (request) =     -30000" in lines
    assert "            Ntried         =      30000" in lines
    assert "            Ncontribute    = %10d" % acc.ncontribute in lines
    w = [l for l in lines if l.startswith("               MeV: wtcontr= ")][0]
    assert float(w[29:]) == pytest.approx(res.yield_ / 30000, rel=6e-8)
    # radiative block: 4(x,a14,'=',l3) / 4(x,a14,'=',i3)
    def rad(fmtv, *items):
        return "".join(" " + A(n, 14) + "=" + (fmtv % v if fmtv else v) for n, v in items)
    assert rad(None, ("use_expon", "  F"), ("include_hard", "  T"), ("calc_spence", "  T")) in lines
    assert rad(None, ("using_rad", "  T"), ("use_offshell_rad", "  T")) in lines
    assert rad("%3d", ("rad_flag", 0), ("extrad_flag", 2), ("one_tail", 0), ("intcor_mode", 1)) in lines
    hc = [l for l in lines if l.startswith("         hardcorfac = ")][0]
    assert float(hc[22:33]) == pytest.approx(central.hardcorfac, abs=6e-4)
    assert any(l.startswith("         c_int(0:3) = ") and len(l) == 22 + 44 for l in lines)
    # miscellaneous: 9915 format(12x,a14,' = ',e16.6,1x,a6)
    nf = [l for l in lines if l.startswith("                   normfac = ")][0]
    assert float(nf[29:45]) == pytest.approx(res.normfac, rel=2e-6)
    assert " " * 15 + "Random Seed = " + "%10d" % 11 in lines
    # limits tables: 9917 format(1x,a18,t21,2f12.3,t50,2f10.3,2x,a5)
    e = [l for l in lines if l.startswith("       E arm  delta")][0]
    assert len(e) == 49 + 20 + 2 + 5 and e.endswith("      %")
    assert float(e[20:32]) == pytest.approx(cfg.gen.e.delta.min, abs=6e-4)
    assert float(e[49:59]) == pytest.approx(acc.contrib[0].lo, abs=6e-4)
    assert " Limiting RADIATION values CONTRIBUTING to the (Em,Pm) distributions:" in lines
    s = [l for l in lines if l.startswith(" " + A("slop.MC  ", 10) + A("E arm delta", 12))][0]
    assert float(s[24:34]) == pytest.approx(cfg.slop_MC_e_used[0], abs=6e-4)
    assert text.endswith("\n\n\n")


def test_central_event_kinematics(run):
    """calculate_central: complete_recon_ev on the spectrometer axes (event.f:1056-1359)."""
    cfg, acc, info, central, res, base = run
    Ein = cfg.Ebeam_vertex_ave - cfg.targ.Coulomb_ave
    eE, th = cfg.spec_e.P, cfg.spec_e.theta
    Q2 = 2 * Ein * eE * (1 - np.cos(th))
    assert central.Q2 == pytest.approx(Q2, rel=1e-14)
    assert central.nu == pytest.approx(Ein - eE, rel=1e-14)
    assert central.W == pytest.approx(np.sqrt(938.27231 ** 2 + 2 * 938.27231 * (Ein - eE) - Q2), rel=1e-12)
    assert central.etatzai == cfg.etatzai
    assert abs(sum(central.frac) - 1.0) < 1e-12 and 0.85 < central.hardcorfac < 1.0
    assert central.g[0] == pytest.approx(central.g[1] + central.g[2] + central.g[3], rel=1e-14)
