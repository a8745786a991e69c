"""The product's one-time setup (simc_b200_config_from_deck: dbase_read post-processing, target_init, limits_init,
radc_init; csrc/run_init.cu) against the ORACLE'S OWN restatement of the same Fortran (oracle/init.cpp, on the oracle's
own trip_thru_target / enerloss_new), field by field of simc_run_config, for every deck under decks/.  The deck text is
split into key = value pairs here, not by the product's reader.  This makes the host init no longer common-mode
between product and oracle in the parity tests."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import RunConfig, config_from_deck
from tests.oracle_lib import (Oracle, load_he3_fixtures, load_pfermi_fixture, load_theory_fixture, write_pfermi_file,
                              write_sf_file, write_theory_file)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DECKS = sorted(glob.glob(os.path.join(ROOT, "decks", "*.inp")))
RTOL = 1e-13


def deck_pairs(path):
    """CTP `begin parm ... end parm` blocks: `key = value ; comment` lines (later assignments win)."""
    kv = {}
    for line in open(path):
        line = line.split(";")[0].strip()
        if not line or line.lower().startswith(("begin", "end")):
            continue
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_%]*)\s*=\s*(.*)$", line)
        if m:
            kv[m.group(1).lower()] = m.group(2).strip().strip("'")
    return kv


def data_extras(kv):
    """What the reference reads from data files before limits_init, computed from the fixtures as the Fortran does:
    theory_init (init.f:862-876: Pm_theory(m)%min/max = first/last momentum of a shell -/+ half a bin; E_Fermi from
    the header), pval(nump) of deut.dat / he3.dat (dbase.f:563-587), Emval(numEm) of the spectral function."""
    x = {}
    A = round(float(kv["targ%a"]))
    on = lambda k: int(float(kv.get(k, "0"))) > 0
    eep = not any(on(k) for k in ("doing_pion", "doing_kaon", "doing_delta", "doing_semi", "doing_rho"))
    if eep and (A == 2 or (A >= 3 and not on("use_benhar_sf"))):
        t = load_theory_fixture("h2" if A == 2 else "c12")
        absmax = 0.0
        for m in range(int(t["n_shells"])):
            first, b, n = float(t["pm_first"][m]), float(t["pm_bin"][m]), int(t["n_pm"][m])
            lo = first - b / 2.
            hi = (first + (n - 1) * b) + b / 2.
            absmax = max(absmax, abs(lo), abs(hi))
        x["x_pm_theory_absmax"] = repr(absmax)
        x["x_e_fermi"] = repr(float(t["e_fermi"]))
    if not eep and not on("doing_semi") and A >= 2:
        if A == 2:
            x["x_pval_last"] = repr(float(load_pfermi_fixture()[0][-1]))
        else:
            (pval, _), sf = load_he3_fixtures()
            x["x_pval_last"] = repr(float(pval[-1]))
            x["x_emval_last"] = repr(float(sf["em"][-1]))
    return x


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    """The reference's working-directory data files, written from the fixtures in the reference's text formats."""
    d = tmp_path_factory.mktemp("data")
    write_theory_file(load_theory_fixture("h2"), str(d / "h2.theory"))
    write_theory_file(load_theory_fixture("c12"), str(d / "c12.theory"))
    write_pfermi_file(*load_pfermi_fixture(), str(d / "deut.dat"))
    (pv, mp), sf = load_he3_fixtures()
    write_pfermi_file(pv, mp, str(d / "he3.dat"))
    write_sf_file(sf, str(d / "benharsf_3mod.dat"))
    return str(d)


def fields(obj, prefix=""):
    """Flattens a ctypes structure into (name, value) pairs."""
    out = []
    for name, typ in obj._fields_:
        v = getattr(obj, name)
        if isinstance(v, C.Structure):
            out += fields(v, prefix + name + ".")
        elif isinstance(v, C.Array):
            flat = np.ctypeslib.as_array(v) if not issubclass(v._type_, (C.Structure, C.Array)) else None
            if flat is not None:
                out += [(f"{prefix}{name}[{i}]", float(flat.flat[i])) for i in range(flat.size)]
            else:
                def walk(a, p):
                    for i, e in enumerate(a):
                        if isinstance(e, C.Array):
                            walk(e, f"{p}[{i}]")
                        elif isinstance(e, C.Structure):
                            out.extend(fields(e, f"{p}[{i}]."))
                        else:
                            out.append((f"{p}[{i}]", float(e)))
                walk(v, prefix + name)
        else:
            out.append((prefix + name, float(v)))
    return out


@pytest.mark.parametrize("deck", DECKS, ids=[os.path.basename(d) for d in DECKS])
def test_product_init_equals_independent_restatement(deck, data_dir):
    cfg = config_from_deck(deck, data_dir=data_dir)[0]
    kv = deck_pairs(deck)
    kv.update(data_extras(kv))
    text = "\n".join(f"{k} = {v}" for k, v in kv.items())
    orc = Oracle()
    ref = RunConfig()
    orc.L.oracle_init_from_kv.argtypes = [C.c_char_p, C.c_void_p]
    rc = orc.L.oracle_init_from_kv(text.encode(), C.byref(ref))
    assert rc == 0, orc.L.oracle_last_error().decode() if hasattr(orc.L, "oracle_last_error") else rc
    a, b = dict(fields(cfg)), dict(fields(ref))
    assert a.keys() == b.keys()
    skip = {"w_ref", "Egamma_res_limit"}           # product-side quantities with no counterpart in init.f
    bad = []
    for k in a:
        if k in skip:
            continue
        x, y = a[k], b[k]
        if x == y or (np.isnan(x) and np.isnan(y)):
            continue
        if abs(x - y) <= RTOL * max(abs(x), abs(y)):
            continue
        bad.append((k, x, y))
    assert not bad, bad[:20]
