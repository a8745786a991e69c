"""End-to-end test of the stand-alone driver `simc_b200` (what program simc does around its loop): deck ->
optics and tables from a working directory laid out like the reference's -> run -> the reference's .hist / .gen / .geni
text files and the .bin ntuple.  Compared with the same run made through the Python host API on the same files."""
import os
import re
import subprocess

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, central_event, config_from_deck, load_optics_fixture, report_info_from_deck, write_reports
from simc_gfortran_b200.lib import normalise, ntuple_tags, read_ntuple_file
from simc_gfortran_b200.optics import write_cosy_files

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "simc_gfortran_b200", "simc_b200")
FILES = {1: ("hms/forward_cosy.dat", "hms/recon_cosy.dat"), 5: ("shms/shms_forward.dat", "shms/shms_recon.dat")}


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("simc_work")
    for arm, (fwd, rec) in FILES.items():
        os.makedirs(d / os.path.dirname(fwd), exist_ok=True)
        write_cosy_files(load_optics_fixture(arm), str(d / fwd), str(d / rec))
    return d


def run_driver(workdir, ngen, out, extra=()):
    txt = open(os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")).read()
    txt, n = re.subn(r"ngen = -?\d+", f"ngen = {ngen}", txt, count=1)
    assert n == 1
    deck = str(workdir / f"deck_{out}.inp")
    open(deck, "w").write(txt)
    if not os.path.exists(DRIVER):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([DRIVER, deck, "--data", str(workdir), "--out", str(workdir / out), "--seed", "5", *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    return deck, parse_hist(str(workdir / out) + ".hist")


def parse_hist(path):
    """The numbers of subroutine report (simc.f:913-967) the tests look at."""
    hist = {}
    for line in open(path):
        m = re.match(r"^\s+(Ntried|Ncontribute|Ngen \(request\))\s*=\s*(-?\d+)$", line.rstrip())
        if m:
            hist[m.group(1)] = int(m.group(2))
        m = re.match(r"^\s+(CENTRAL.sigcc|AVERAGE.sigcc|normfac|genvol|charge)\s=\s+([-0-9.E+]+)", line.rstrip())
        if m:
            hist[m.group(1)] = float(m.group(2))
        m = re.match(r"^\s+MeV: wtcontr=\s+([-0-9.E+]+)$", line.rstrip())
        if m:
            hist["wtcontr"] = float(m.group(1))
    return hist


def test_fixed_number_of_tries(workdir):
    deck, hist = run_driver(workdir, -60000, "tries", ("--ntuple", "1", "--chunk", "25000"))
    cfg, ngen, charge = config_from_deck(deck)
    sim = Simc(cfg, mode="strict")
    try:
        for arm, (fwd, rec) in FILES.items():
            sim.load_optics(arm, str(workdir / fwd), str(workdir / rec))
        acc = sim.accum_clear()
        sim.run(0, 60000, 5, acc)
        rows, _ = sim.ntuple_batch(0, 60000, 5)
        res = normalise(cfg, acc, ngen, charge)
        info = report_info_from_deck(deck)
        info.random_seed = 5
        central = central_event(cfg, info, sim)
        write_reports(str(workdir / "py"), cfg, info, central, acc, res, "t1", "t2")
    finally:
        sim.close()
    assert hist["Ntried"] == 60000 and hist["Ncontribute"] == acc.ncontribute and hist["Ngen (request)"] == -60000
    assert abs(hist["wtcontr"] * 60000 / res.yield_ - 1) < 1e-7 and abs(hist["normfac"] / res.normfac - 1) < 1e-5
    # central%sigcc: sigep at the spectrometer settings (calculate_central), through the weight kernel
    assert hist["CENTRAL.sigcc"] > 0 and abs(hist["CENTRAL.sigcc"] / central.sigcc - 1) < 1e-5
    assert 0.1 < hist["CENTRAL.sigcc"] / hist["AVERAGE.sigcc"] < 100.0
    # the driver's three text files are byte for byte what the host API writes for the same run (but for the times)
    for ext in (".geni", ".gen"):
        assert open(str(workdir / "tries") + ext).read() == open(str(workdir / "py") + ext).read(), ext
    a, b = open(str(workdir / "tries.hist")).read().split("\n"), open(str(workdir / "py.hist")).read().split("\n")
    assert len(a) == len(b) and a[:1] == b[:1] and a[3:] == b[3:]
    tags, vals = read_ntuple_file(str(workdir / "tries.bin"))
    assert tags == ntuple_tags(cfg) and vals.shape == rows.shape == (acc.ncontribute, 46)
    assert np.array_equal(vals, rows)                   # same kernels, same tries: identical rows
    # the weights in the file, normalised, give the yield of the summary (events inside the cuts)
    assert res.yield_ <= vals[:, 43].sum() * res.normfac * (1 + 1e-9)


def test_until_n_successes(workdir):
    """ngen > 0 (simc.f:346-350): stop at exactly N successes; the try range is bisected, every try reproducible."""
    _, hist = run_driver(workdir, 2500, "succ", ("--chunk", "20000"))
    assert hist["Ncontribute"] == 2500
    assert 2500 / 0.35 < hist["Ntried"] < 2500 / 0.1


def test_missing_file_is_an_error(workdir, tmp_path):
    txt = open(os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")).read()
    deck = str(tmp_path / "d.inp")
    open(deck, "w").write(txt)
    r = subprocess.run([DRIVER, deck, "--data", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "load_optics" in r.stderr


def run_deck(workdir, deck_name, ngen, out, extra=()):
    txt = open(os.path.join(ROOT, "decks", deck_name)).read()
    txt, n = re.subn(r"ngen = -?\d+", f"ngen = {ngen}", txt, count=1)
    assert n == 1
    deck = str(workdir / f"deck_{out}.inp")
    open(deck, "w").write(txt)
    r = subprocess.run([DRIVER, deck, "--data", str(workdir), "--out", str(workdir / out), "--seed", "9", "--ntuple", "1", *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    return deck, parse_hist(str(workdir / out) + ".hist")


def test_polarised_target_and_calorimeter_decks(workdir):
    """The driver on the two decks that need more than maps: using_tgt_field (it reads trg_field_map.dat from the data
    directory, like trgInit) and the pi0 deck with the NPS calorimeter as the hadron arm (no maps for that arm).  The
    .bin files carry the wider rows (61 and 65 columns) under the reference's tags."""
    from tests.oracle_lib import load_field_fixture, write_field_file
    bz, br = load_field_fixture()
    write_field_file(bz, br, str(workdir / "trg_field_map.dat"))
    n = 60000
    deck, hist = run_deck(workdir, "w1_poltar_eepi_hydrogen_hms_shms.inp", -n, "pol")
    cfg = config_from_deck(deck)[0]
    sim = Simc(cfg, mode="strict")
    try:
        for arm, (fwd, rec) in FILES.items():
            sim.load_optics(arm, str(workdir / fwd), str(workdir / rec))
        sim.load_field_file(str(workdir / "trg_field_map.dat"))
        acc = sim.accum_clear()
        sim.run(0, n, 9, acc)
        rows, _ = sim.ntuple_batch(0, n, 9)
    finally:
        sim.close()
    assert hist["Ntried"] == n and hist["Ncontribute"] == acc.ncontribute > 100
    tags, vals = read_ntuple_file(str(workdir / "pol.bin"))
    assert tags == ntuple_tags(cfg) and len(tags) == 61 and tags[53] == "th_tarq"
    assert vals.shape == rows.shape and np.array_equal(vals, rows)
    deck, hist = run_deck(workdir, "z1_eepi0_hydrogen_hms_nps.inp", -n, "pi0")
    cfg = config_from_deck(deck)[0]
    tags, vals = read_ntuple_file(str(workdir / "pi0.bin"))
    assert hist["Ntried"] == n and hist["Ncontribute"] == len(vals) > 100
    assert tags == ntuple_tags(cfg) and len(tags) == 65 and tags[55] == "Egamma1"
    g = vals[:, 55:59] + vals[:, 61:65]
    assert np.allclose(np.sqrt(g[:, 0] ** 2 - (g[:, 1:] ** 2).sum(axis=1)), 134.9766, rtol=1e-9)
