"""CPU checks of the oracle's radiative option branches (rows a10 / a11 of SURVEY 8(a)): the two soft-photon
calculations of brem.f must agree with each other on elastic kinematics (`brem` rebuilds the event from ein, eout;
`bremos` takes the four-vectors: two independent restatements of two independent reference routines), Spence series
against closed forms, extrad_phi limits, and the behaviour of generate_rad in the (Egamma1, Egamma2, Egamma3) basis."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck, load_optics_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C1 = os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")


def elastic_rows(k, seed=1):
    rng = np.random.default_rng(seed)
    Mp = 938.27231
    Ein = rng.uniform(8000.0, 8800.0, k)
    eth = rng.uniform(0.40, 0.50, k)
    eE = Ein * Mp / (Mp + Ein * (1 - np.cos(eth)))          # event.f:518
    phi = rng.uniform(4.6, 4.8, k)
    ue = np.stack([np.sin(eth) * np.cos(phi), np.sin(eth) * np.sin(phi), np.cos(eth)])
    nu = Ein - eE
    q = np.sqrt(2 * Ein * eE * (1 - ue[2]) + nu * nu)
    up = np.stack([-eE * ue[0] / q, -eE * ue[1] / q, (Ein - eE * ue[2]) / q])
    pE = np.sqrt(q * q + Mp * Mp)
    emax = rng.uniform(20.0, 1200.0, k)
    emin = np.where(rng.uniform(size=k) < 0.5, rng.uniform(-50.0, 0.0, k), rng.uniform(0.0, 0.9, k) * emax)
    return np.stack([Ein, eE, eth, ue[0], ue[1], ue[2], pE, q, up[0], up[1], up[2], rng.uniform(2e-3, 3e-2, k),
                     rng.uniform(5e-3, 5e-2, k), rng.uniform(0.0, 1.0, k) * emax, emin, emax])


def test_onshell_and_offshell_brem_agree_on_elastic_kinematics(oracle):
    cfg = config_from_deck(C1)[0]
    inp = elastic_rows(4000)
    off = oracle.radc_batch(cfg, inp)                      # radc_init_ev through bremos
    cfg.use_offshell_rad = 0
    on = oracle.radc_batch(cfg, inp)                       # ... through brem
    # hard correction: the same formula of Q2 (brem.f:189 / :553)
    assert np.allclose(on[6], off[6], rtol=1e-9, atol=0)
    # d(bsoft)/dE at 450 MeV: g(4) = -dsoft_prime*Ecutoff + bt(1) + bt(2)
    assert np.allclose(on[5], off[5], rtol=2e-6, atol=0), np.abs(on[5] / off[5] - 1).max()
    # brem's own dbsoft column is what the on-shell g(4) was made of
    g4 = -on[25] * 450.0 + on[0] + on[1]
    assert np.allclose(g4, on[5], rtol=1e-14, atol=0)
    # the peaked weight (two soft-photon integrals per row) agrees as well
    ok = off[9] > 0
    assert ok.mean() > 0.9
    assert np.allclose(on[9][ok], off[9][ok], rtol=1e-4, atol=0), np.abs(on[9][ok] / off[9][ok] - 1).max()


def test_spence_series_and_schwinger(oracle):
    cfg = config_from_deck(C1)[0]
    inp = elastic_rows(1000, seed=2)
    out = oracle.radc_batch(cfg, inp)
    Ein, eE, eth = inp[0], inp[1], inp[2]
    Q2 = 2 * Ein * eE * (1 - np.cos(eth))
    Me, alpi, amu = 0.51099906, (1 / 137.0359895) / np.pi, 931.49432
    lq = np.log(Q2 / Me ** 2) - 1.0
    s2 = np.sin(eth / 2) ** 2
    # Li2(x) from scipy: spen(x) of radc.f is the series sum x^i / i^2, cut at a relative term of 1e-4
    from scipy.special import spence as sp_spence
    li2 = sp_spence(1.0 - (1.0 - s2))                       # scipy's spence(z) = Li2(1 - z)
    dhard = -alpi * (2.166666 * lq + (li2 - 2.5893784) - np.log(Ein / eE) ** 2 / 2.0)
    assert np.allclose(out[20], dhard, rtol=0, atol=alpi * 2e-3)
    b = 1.0 + 2.0 * (Ein - eE) * s2 / (cfg.targ.A * amu)
    dsoft = alpi * lq * np.log(Ein * eE * b / 450.0 ** 2)
    assert np.allclose(out[19], dsoft, rtol=1e-13, atol=0)
    # intcor_mode = 0: the hard factor comes from schwinger, and g(4) has no internal part (dsoft_prime stays 0.0)
    cfg.intcor_mode = 0
    sch = oracle.radc_batch(cfg, inp)
    assert np.allclose(sch[6], 1.0 / (1.0 - out[20]), rtol=1e-14, atol=0)
    assert np.allclose(sch[5], sch[0] + sch[1], rtol=1e-15, atol=0)


def test_extrad_phi_and_friedrich(oracle):
    cfg = config_from_deck(C1)[0]
    inp = elastic_rows(1000, seed=3)
    eg = inp[13]
    cfg.extrad_flag = 1
    assert np.all(oracle.radc_batch(cfg, inp)[17:19] == 1.0)
    cfg.extrad_flag = 2
    o2 = oracle.radc_batch(cfg, inp)
    # radc.f:690: 1 - bt(i)/E(i)/g(i)*Egamma with g(i) = lambda(i) + bt(i)
    assert np.allclose(o2[17], 1 - o2[0] / inp[0] / (o2[2] + o2[0]) * eg, rtol=1e-14)
    assert np.allclose(o2[18], 1 - o2[1] / inp[1] / (o2[3] + o2[1]) * eg, rtol=1e-14)
    cfg.extrad_flag = 3
    o3 = oracle.radc_batch(cfg, inp)
    eta = cfg.etatzai
    x = eg / inp[0]
    t = o3[0] / eta
    from scipy.special import gamma
    want = (1 - x + x * x / eta) * np.exp(t * ((eta - 0.5) - eta * x + x * x / 2)) * gamma(1 + o3[0])
    assert np.allclose(o3[17], want, rtol=1e-4)             # radc.f gamma() is a 5th-order polynomial fit (5e-5)
    # extrad_friedrich (radc.f:650-664): closed form
    db = t * (-(eta - 0.5) - eta * np.log(x) + eta * x - 0.5 * x * x)
    assert np.allclose(o3[21], db, rtol=1e-12)
    assert np.allclose(o3[22], -t / inp[0] * (eta / x - eta + x), rtol=1e-12)
    # extrad_flag = 3 changes nothing in the peaked weight: the Friedrich terms are computed and never read
    cfg.extrad_flag = 2
    assert np.array_equal(o3[9], oracle.radc_batch(cfg, inp)[9])


def test_basis_constants(oracle):
    """c(0) of the combined tails (init.f:796-802) rebuilt from c(1..3), g(0..3)."""
    cfg = config_from_deck(C1)[0]
    inp = elastic_rows(500, seed=4)
    o = oracle.radc_batch(cfg, inp)
    from scipy.special import gamma
    g1, g2, g3 = o[2] + o[0], o[3] + o[1], o[4]
    g0 = g1 + g2 + g3
    c0 = o[11] * o[12] * g0 / g1 / g2 * o[13] / g3 * gamma(1 + g1) * gamma(1 + g2) * gamma(1 + g3) / gamma(1 + g0)
    assert np.all(g3 > 0)
    assert np.allclose(o[14], c0, rtol=5e-4)
    assert np.allclose(o[16], o[2] + o[3] + o[4], rtol=1e-15)


@pytest.mark.parametrize("rad_flag", [2, 3])
def test_generate_rad_in_the_three_photon_basis(oracle_with_optics, rad_flag):
    orc = oracle_with_optics
    cfg = config_from_deck(C1)[0]
    cfg.rad_flag = rad_flag
    cfg.extrad_flag = 1
    n = 30000
    rec, status = orc.event_batch(cfg, 0, n, 9)
    ok = status >= 1
    ntail = rec[25][ok]
    eg = rec[22:25][:, ok]
    if rad_flag == 2:
        # one tail, drawn with equal probability (radc.f:206-208); generation succeeds at different rates per tail
        assert set(np.unique(ntail)) == {1.0, 2.0, 3.0}
        for k in (1, 2, 3):
            sel = ntail == k
            assert (eg[k - 1][sel] > 0).mean() > 0.99      # x = y**(1/g) with g ~ 0.03 can underflow to zero
            assert np.all(np.delete(eg, k - 1, axis=0)[:, sel] == 0)
    else:
        assert np.all(ntail == 0)
        assert (eg > 0).mean() > 0.99                      # every tail radiates in every event (up to underflow)
    done = status == 4
    assert done.sum() > 500
    assert np.all(np.isfinite(rec[7][done])) and np.all(rec[7][done] > 0)      # gen_weight
    # the same run split in two halves gives the same records (no state carried from event to event)
    rec2, status2 = orc.event_batch(cfg, n // 2, n // 2, 9)
    assert np.array_equal(status2, status[n // 2:])
    assert np.array_equal(rec2, rec[:, n // 2:])
