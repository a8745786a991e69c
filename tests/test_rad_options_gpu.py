"""GPU parity of the radiative option branches no shipped deck selects:
the on-shell `brem` (brem.f:6-214, use_offshell_rad = 0), `schwinger` (radc.f:711, intcor_mode = 0), the Friedrich
prescription (`extrad_friedrich` radc.f:650, `extrad_phi` with extrad_flag = 3 radc.f:695-703), BASICRAD alone in the
peaked basis (rad_flag = 1) and the (Egamma1, Egamma2, Egamma3) basis of rad_flag = 2 and 3 (radc.f:198-213: one tail
drawn uniformly / all tails at once).  Stage level: `simc_b200_radc_batch` against the oracle at 1e-12 on dumped
vectors for every setting.  Loop level: edited decks, both sides run the same tries and must agree exactly on
every counter, STOP counter and count histogram and to LOOSE on the weight sum."""
import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.test_loop_gpu import LOOSE, RTOL, accum_equal_exact
from tests.test_loop_variants_gpu import ROOT, VARIANTS, edited_deck

pytestmark = pytest.mark.gpu

C1, C2, C3, D1 = ("c1_eep_hydrogen_hms_shms.inp", "c2_eep_carbon_hms_sos.inp", "c3_eepi_hydrogen_hms_shms.inp",
                  "d1_eep_deuterium_hms_sos.inp")
RAD_VARIANTS = {
    "basicrad_peaked": (C1, {"rad_flag": "1"}),
    "basicrad_peaked_phi1": (C1, {"rad_flag": "1", "extrad_flag": "1"}),
    "basicrad_peaked_friedrich": (C1, {"rad_flag": "1", "extrad_flag": "3"}),
    "one_of_three_tails": (C1, {"rad_flag": "2"}),
    "all_three_tails": (C1, {"rad_flag": "3"}),
    "all_three_tails_friedrich": (C1, {"rad_flag": "3", "extrad_flag": "3"}),
    "all_three_tails_default_extrad": (C1, {"rad_flag": "3", "extrad_flag": "0"}),
    "onshell_brem": (C1, {"use_offshell_rad": "0"}),
    "schwinger": (C1, {"intcor_mode": "0"}),
    "friedrich_peaked": (C1, {"extrad_flag": "3"}),
    "default_extrad": (C1, {"extrad_flag": "0"}),
    "use_expon": (C1, {"use_expon": "1"}),
    "all_tails_no_proton": (C1, {"rad_flag": "3", "one_tail": "-3"}),
    "carbon_all_three_tails": (C2, {"rad_flag": "3"}),
    "carbon_one_of_three": (C2, {"rad_flag": "2", "extrad_flag": "1"}),
    "pion_all_three_tails": (C3, {"rad_flag": "3"}),
    "pion_one_of_three": (C3, {"rad_flag": "2"}),
    "deuterium_all_three_tails": (D1, {"rad_flag": "3"}),
    "deuterium_one_of_three": (D1, {"rad_flag": "2"}),
}
VARIANTS.update({"rad_" + k: v for k, v in RAD_VARIANTS.items()})


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    from tests.oracle_lib import load_theory_fixture, write_theory_file
    d = tmp_path_factory.mktemp("theory_rad")
    write_theory_file(load_theory_fixture("h2"), str(d / "h2.theory"))
    return str(d)


def _tables(cfg, sim, orc):
    if cfg.doing_heavy:
        from tests.test_weight_gpu import sf_table
        orc.set_sf_table(*sf_table())
        sim.set_sf_table(*sf_table())
    if cfg.doing_deuterium:
        from tests.oracle_lib import load_theory_fixture
        t = load_theory_fixture("h2")
        orc.set_theory_table(t, cfg.doing_heavy)
        sim.set_theory_table(t)


@pytest.mark.parametrize("name", sorted(RAD_VARIANTS))
def test_loop_with_radiative_option(tmp_path, name, oracle_with_optics, data_dir):
    orc = oracle_with_optics
    cfg = config_from_deck(edited_deck(tmp_path, "rad_" + name), data_dir=data_dir)[0]
    edits = RAD_VARIANTS[name][1]
    if "rad_flag" in edits:
        assert cfg.rad_flag == int(edits["rad_flag"])
    if edits.get("extrad_flag") == "0":          # radc_init's defaults, init.f:620-626
        assert cfg.extrad_flag == (3 if cfg.rad_flag == 0 else 1)
    sim = Simc(cfg, mode="strict")
    try:
        for arm in sorted({cfg.electron_arm, cfg.hadron_arm}):
            sim.set_optics(load_optics_fixture(arm))
        _tables(cfg, sim, orc)
        n = 40000
        ref = orc.run(cfg, 0, n, 33, threads=8)
        acc = sim.accum_clear()
        sim.run(0, n, 33, acc)
        accum_equal_exact(acc, ref)
        assert ref.nsuccess > 100, ref.nsuccess
        a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
        assert abs(a - b) <= LOOSE * abs(b), (a, b)
        assert acc.nonfinite == ref.nonfinite == 0
        # the photon-energy ranges (contrib%rad%Egamma(1..3), event.f:79-86) show which tails radiated
        eg_hi = [acc.contrib[26 + k].hi for k in range(3)]
        if cfg.rad_flag == 3 and name != "all_tails_no_proton":
            assert all(x > 0 for x in eg_hi), eg_hi
        if name == "all_tails_no_proton":
            assert eg_hi[0] > 0 and eg_hi[1] > 0 and eg_hi[2] == 0
    finally:
        sim.close()


@pytest.mark.parametrize("name", ["all_three_tails", "one_of_three_tails", "onshell_brem", "carbon_all_three_tails",
                                  "pion_all_three_tails"])
def test_event_records_with_radiative_option(tmp_path, name, oracle_with_optics, data_dir):
    """Per-try records: photon energies, ntail, the generation weight with the finished radiative weight."""
    orc = oracle_with_optics
    cfg = config_from_deck(edited_deck(tmp_path, "rad_" + name), data_dir=data_dir)[0]
    sim = Simc(cfg, mode="strict")
    try:
        for arm in sorted({cfg.electron_arm, cfg.hadron_arm}):
            sim.set_optics(load_optics_fixture(arm))
        _tables(cfg, sim, orc)
        n = 20000
        rec, status = sim.event_batch(0, n, 5)
        ref, ref_status = orc.event_batch(cfg, 0, n, 5)
        assert np.array_equal(status, ref_status)
        names = sim.event_field_names()
        for col in ("n_draws", "stop_p", "stop_e", "ntail"):
            k = names.index(col)
            assert np.array_equal(rec[k], ref[k]), col
        gen_ok = status >= 1
        for col in ("Egamma_used1", "Egamma_used2", "Egamma_used3", "orig.e.E", "orig.p.E", "vertex.Ein"):
            k = names.index(col)
            err = np.abs(rec[k] - ref[k])[gen_ok] / np.maximum(np.abs(ref[k])[gen_ok], 1.0)
            assert err.max() < 5e-9, (col, err.max())
        done = status == 4
        k = names.index("gen_weight")
        err = np.abs(rec[k] - ref[k])[done] / np.abs(ref[k])[done]
        assert err.max() < (1e-9 if cfg.rad_flag >= 2 else LOOSE), err.max()
        if cfg.rad_flag == 3:
            k = names.index("ntail")
            assert np.all(rec[k][gen_ok] == 0)
    finally:
        sim.close()


FLAG_SETS = [dict(rad_flag=0, extrad_flag=2, intcor_mode=1, use_offshell_rad=1),
             dict(rad_flag=0, extrad_flag=3, intcor_mode=1, use_offshell_rad=1),
             dict(rad_flag=0, extrad_flag=2, intcor_mode=1, use_offshell_rad=0),
             dict(rad_flag=0, extrad_flag=1, intcor_mode=0, use_offshell_rad=1),
             dict(rad_flag=1, extrad_flag=2, intcor_mode=1, use_offshell_rad=1),
             dict(rad_flag=1, extrad_flag=3, intcor_mode=1, use_offshell_rad=0),
             dict(rad_flag=3, extrad_flag=3, intcor_mode=1, use_offshell_rad=1),
             dict(rad_flag=2, extrad_flag=2, intcor_mode=0, use_offshell_rad=0)]


@pytest.mark.parametrize("flags", FLAG_SETS, ids=lambda f: "-".join(f"{k[:3]}{v}" for k, v in f.items()))
def test_radc_stage_with_option(flags, oracle_with_optics):
    orc = oracle_with_optics
    """radc_init_ev, the basis constants, extrad_phi, schwinger, extrad_friedrich, brem and peaked_rad_weight on
    dumped vertex vectors, 1e-12 (the rows where glibc's pow(x, 0.5) and sqrt differ are allowed and counted)."""
    import os
    cfg = config_from_deck(os.path.join(ROOT, "decks", C1))[0]
    for k, v in flags.items():
        setattr(cfg, k, v)
    rng = np.random.default_rng(11)
    k = 20000
    Ein = rng.uniform(8000.0, 8800.0, k)
    eE = rng.uniform(4100.0, 4900.0, k)
    Mp = 938.27231
    cth = 1.0 - (Ein / eE - 1.0) * Mp / Ein
    eth = np.arccos(cth)
    phi = rng.uniform(4.6, 4.8, k)
    ue = np.stack([np.sin(eth) * np.cos(phi), np.sin(eth) * np.sin(phi), np.cos(eth)])
    nu = Ein - eE
    q = np.sqrt(2 * Ein * eE * (1 - ue[2]) + nu * nu)
    up = np.stack([-eE * ue[0] / q, -eE * ue[1] / q, (Ein - eE * ue[2]) / q])
    pE = np.sqrt(q * q + Mp * Mp)
    emax = rng.uniform(20.0, 1200.0, k)
    emin = np.where(rng.uniform(size=k) < 0.5, rng.uniform(-50.0, 0.0, k), rng.uniform(0.0, 0.9, k) * emax)
    inp = np.stack([Ein, eE, eth, ue[0], ue[1], ue[2], pE, q, up[0], up[1], up[2], rng.uniform(2e-3, 3e-2, k),
                    rng.uniform(5e-3, 5e-2, k), rng.uniform(0.0, 1.0, k) * emax, emin, emax])
    sim = Simc(cfg, mode="strict")
    try:
        ref = orc.radc_batch(cfg, inp)
        out = sim.radc_batch(inp)
    finally:
        sim.close()
    assert np.isfinite(ref).all()
    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300)
    # The on-shell `brem` rebuilds the scattering angle as 2*asin((..)**0.5) and its e-e interference term forms
    # ar1 ~ 1e-8 as 0.5 - 0.4999.. (brem.f:86-88), which amplifies one ulp of asin/cos by ~1e7: the few % of rows where
    # the CUDA and glibc libm differ in the last bit show at 1e-9.  Those columns (23, 25 always; g(4), c(4) and the
    # peaked weight when the run uses it) get the loop's LOOSE bound and a bounded outlier rate; everything else 1e-12.
    soft = [23, 25] + ([5, 7, 9] if not flags["use_offshell_rad"] and flags["intcor_mode"] == 1 else [])
    if not flags["use_offshell_rad"] and flags["rad_flag"] == 0:
        soft = sorted(set(soft + [9]))
    hard = [k for k in range(err.shape[0]) if k not in soft]
    outliers = (err[hard] > RTOL).any(axis=0)
    assert outliers.mean() < 2e-3, float(outliers.mean())
    assert err[hard][:, ~outliers].max() <= RTOL
    assert err.max() < 1e-6
    assert (err[soft] > RTOL).any(axis=0).mean() < 0.08
    assert err[soft].max() < 1e-7
