"""CPU tests of the end-of-run host code: normalisation (simc.f:94-101, 366-432) and the ntuple file writer
(NtupleInit.f, results_write.f:264-266) -- host-only entry points of the C ABI, no GPU needed."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck
from simc_gfortran_b200.lib import Accum, Fixed128, normalise, ntuple_tags, read_ntuple_file, write_ntuple_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def deck(name):
    return config_from_deck(os.path.join(ROOT, "decks", name))


def fixed(x, qexp=-40):
    v = int(round(x * 2.0 ** (-qexp)))
    f = Fixed128()
    f.lo = v & ((1 << 64) - 1)
    f.hi = v >> 64
    f.qexp = qexp
    return f


def test_normalisation_hydrogen_elastic():
    cfg, ngen, charge = deck("c1_eep_hydrogen_hms_shms.inp")
    acc = Accum()
    acc.ntried, acc.nsuccess, acc.npasscuts = 1000000, 200000, 150000
    acc.wtcontribute = fixed(3.5)
    acc.sum_sigcc = fixed(200000 * 0.25)
    for k in range(8):
        acc.sumerr[k] = fixed(150000 * 0.01 * (k + 1))
        acc.sumerr2[k] = fixed(150000 * ((0.01 * (k + 1)) ** 2 + 0.04))
    assert ngen < 0                      # every shipped deck asks for |ngen| tries
    r = normalise(cfg, acc, ngen, charge)
    # simc.f:94-101: EXPER%charge / (mass_amu / 3.75914e6 / abundancy * cos(angle) / (thick [mg/cm2]))
    lumi = charge / (cfg.targ.mass_amu / 3.75914e6 / (cfg.targ.abundancy / 100.) / (cfg.targ.thick * 1000.))
    assert abs(r.luminosity / lumi - 1) < 1e-14
    genvol = (cfg.gen.e.yptar.max - cfg.gen.e.yptar.min) * (cfg.gen.e.xptar.max - cfg.gen.e.xptar.min)     # 2-fold
    assert abs(r.genvol / genvol - 1) < 1e-14
    # simc.f:346-350,368: ngen < 0 -> nevent counts every try, normfac = luminosity/ntried*nevent*genvol = luminosity*genvol
    assert r.nevent == 1000000
    assert abs(r.normfac / (lumi * genvol) - 1) < 1e-14
    assert abs(r.yield_ / (3.5 * r.normfac) - 1) < 1e-11
    assert abs(r.central_sigcc_ave - 200000 * 0.25 / 1000000) < 1e-11              # simc.f:959: sum_sigcc/nevent
    # ngen > 0: the loop runs until ngen successes, nevent = successes
    r2 = normalise(cfg, acc, 200000, charge)
    assert r2.nevent == 200000
    assert abs(r2.normfac / (lumi / 1000000 * 200000 * genvol) - 1) < 1e-14
    assert abs(r2.central_sigcc_ave - 0.25) < 1e-11
    for k in range(8):
        assert abs(r.aveerr[k] - 0.01 * (k + 1)) < 1e-10 and abs(r.resol[k] - 0.2) < 1e-9


@pytest.mark.parametrize("name,fold", [("c3_eepi_hydrogen_hms_shms.inp", 5), ("c2_eep_carbon_hms_sos.inp", 6),
                                       ("c4_semi_deuterium_hms_shms.inp", 6)])
def test_generation_volume_by_reaction(name, fold):
    cfg, _, charge = deck(name)
    acc = Accum()
    acc.ntried, acc.nsuccess = 10, 5
    g = cfg.gen
    v = (g.e.yptar.max - g.e.yptar.min) * (g.e.xptar.max - g.e.xptar.min) * (g.p.yptar.max - g.p.yptar.min) * \
        (g.p.xptar.max - g.p.xptar.min) * (g.e.E.max - g.e.E.min)
    if fold == 6:
        v *= g.p.E.max - g.p.E.min            # simc.f:392-394: doing_heavy .or. doing_semi
    assert abs(normalise(cfg, acc, -10, charge).genvol / v - 1) < 1e-14


def test_ntuple_tags_follow_ntupleinit():
    assert len(ntuple_tags(deck("c1_eep_hydrogen_hms_shms.inp")[0])) == 46
    t = ntuple_tags(deck("c3_eepi_hydrogen_hms_shms.inp")[0])
    assert len(t) == 53 and t[33] == "missmass" and t[45] == "Weight" and t[52] == "phipqi"
    t = ntuple_tags(deck("c5_eek_hydrogen_hrsl_hrsr.inp")[0])
    assert len(t) == 55 and t[53:] == ["saghai", "factor"]
    t = ntuple_tags(deck("c4_semi_deuterium_hms_shms.inp")[0])
    assert len(t) == 56 and t[0] == "hsdelta" and t[12] == "ssdelta" and t[43:47] == ["z", "zi", "pt2", "pt2i"]


def test_ntuple_file_is_fortran_unformatted_sequential(tmp_path):
    cfg = deck("c1_eep_hydrogen_hms_shms.inp")[0]
    rows = np.random.default_rng(3).normal(size=(137, 46))
    path = str(tmp_path / "run.bin")
    write_ntuple_file(cfg, path, rows)
    # header: [4][int32 46][4], then 46 x [16][16 chars][16]; each value [8][float64][8]
    assert os.path.getsize(path) == 12 + 46 * 24 + 137 * 46 * 16
    raw = open(path, "rb").read()
    assert raw[:12] == (4).to_bytes(4, "little") + (46).to_bytes(4, "little") + (4).to_bytes(4, "little")
    assert raw[12:16] == (16).to_bytes(4, "little") and raw[16:32] == b"hsdelta" + b" " * 9
    tags, vals = read_ntuple_file(path)
    assert tags == ntuple_tags(cfg) and np.array_equal(vals, rows)


def test_ntuple_tags_of_the_new_layouts():
    """Tags of NtupleInit.f:101-343 for the layouts added in round 2: rho (59), pi0 -> gamma gamma (65: the reference's
    tags for the photon columns are named in NtupleInit.f:163-184; here the row layout is checked, the file header takes
    the meson tags), and the eight polarised-target tags behind phipqi."""
    import os
    from simc_gfortran_b200 import config_from_deck
    from simc_gfortran_b200.lib import ntuple_tags
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rho = config_from_deck(os.path.join(root, "decks", "r1_eerho_hydrogen_sos_hms.inp"))[0]
    t = ntuple_tags(rho)
    assert len(t) == 59 and t[33] == "missmass" and t[55] == "phipqi" and t[56:] == ["Mrho", "Thrho", "mmnuc"]
    pol = config_from_deck(os.path.join(root, "decks", "w1_poltar_eepi_hydrogen_hms_shms.inp"))[0]
    t = ntuple_tags(pol)
    assert len(t) == 61 and t[52] == "phipqi" and t[53:] == ["th_tarq", "phitarq", "beta", "phis", "phic", "betai", "phisi", "phici"]
    rho.using_tgt_field = 1
    t = ntuple_tags(rho)
    assert len(t) == 67 and t[56:64] == ["th_tarq", "phitarq", "beta", "phis", "phic", "betai", "phisi", "phici"] and t[64:] == ["Mrho", "Thrho", "mmnuc"]


def test_ntuple_tags_pizero():
    import os
    from simc_gfortran_b200 import config_from_deck
    from simc_gfortran_b200.lib import ntuple_tags
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = ntuple_tags(config_from_deck(os.path.join(root, "decks", "z1_eepi0_hydrogen_hms_nps.inp"))[0])
    assert len(t) == 65 and t[52] == "phipqi" and t[53:56] == ["xcal_gamma1", "ycal_gamma1", "Egamma1"] and t[64] == "Pgamma2z"
