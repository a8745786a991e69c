"""The product's own COSY file readers (csrc/optics_host.cpp, entry point simc_b200_read_optics_files) on the five
REAL forward / reconstruction pairs of the reference tree, against (1) an independent reading of the same files
written here directly from the Fortran format statements -- shared/transp.f:484 `format(1x,5g14.7,1x,6i1)` with the
time-of-flight column dropped (:377-393), hms/mc_hms_recon.f:144 `format(1x,4g16.9,1x,5i1)` -- and (2) the committed
fixtures tests/golden/optics_*.npz that the GPU tests and bench.py load (made by the ORACLE's loader,
tools/make_fixtures.py).  Bit for bit.  The reference tree only exists in the build container: skipped elsewhere."""
import ctypes as C
import os

import numpy as np
import pytest

from simc_gfortran_b200 import load_library, load_optics_fixture

REF = os.environ.get("SIMC_REFERENCE", "/root/reference")
FILES = {
    1: ("hms/forward_cosy.dat", "hms/recon_cosy.dat", 12),
    5: ("shms/shms_forward.dat", "shms/shms_recon.dat", 32),
    2: ("sos/forward_cosy.dat", "sos/recon_cosy.dat", 10),
    3: ("hrsr/hrs_forward_cosy.dat", "hrsr/hrs_recon_cosy.dat", 12),
    4: ("hrsl/hrs_forward_cosy.dat", "hrsl/hrs_recon_cosy.dat", 12),
}
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def fortran_real(field):
    """A g14.7 / g16.9 input field: blanks are ignored, D exponents allowed."""
    s = field.replace(" ", "").replace("D", "E").replace("d", "e")
    return float(s) if s else 0.0


def read_forward_python(path):
    """transp_init (shared/transp.f:294-474): leading '!' lines are a header; a class is the data lines up to a line
    starting with ' ---'; comments between classes are skipped; lines whose TOF digit (6th integer column... the 5th of
    the six i1 fields) is non-zero are dropped."""
    lines = [l.rstrip("\n") for l in open(path)]
    k = 0
    while lines[k].startswith("!"):
        k += 1
    classes, cur = [], []
    in_class = True
    while k < len(lines):
        line = lines[k]
        if in_class:
            if line.startswith(" ---"):
                classes.append(cur)
                cur = []
                in_class = False
            else:
                line = line.ljust(80)
                coef = [fortran_real(line[1 + 14 * i:15 + 14 * i]) for i in range(5)]
                digits = [int(line[72 + j]) if line[72 + j].strip() else 0 for j in range(6)]
                if digits[4] == 0:                      # idummy: time-of-flight exponent
                    cur.append((coef, digits[:4] + [digits[5]]))
        else:
            if not (line.startswith("!") or line.startswith(" ---") or not line.strip()):
                in_class = True
                continue
        k += 1
    coef = np.array([c for cl in classes for c, _ in cl], dtype=np.float64).reshape(-1, 5)
    expo = np.array([e for cl in classes for _, e in cl], dtype=np.int8).reshape(-1, 5)
    start = np.cumsum([0] + [len(cl) for cl in classes]).astype(np.int32)
    return start, coef, expo


def read_recon_python(path):
    lines = [l.rstrip("\n") for l in open(path)]
    k = 0
    while lines[k].startswith("!"):
        k += 1
    coef, expo = [], []
    while not lines[k].startswith(" ---"):
        line = lines[k].ljust(80)
        coef.append([fortran_real(line[1 + 16 * i:17 + 16 * i]) for i in range(4)])
        expo.append([int(line[66 + j]) if line[66 + j].strip() else 0 for j in range(5)])
        k += 1
    return np.array(coef, dtype=np.float64), np.array(expo, dtype=np.int8)


def product_read(fwd, rec):
    L = load_library()
    L.simc_b200_read_optics_files.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.c_char_p, C.c_int]
    cap = 45000
    cs = np.zeros(42, np.int32); fc = np.zeros((cap, 5)); fe = np.zeros((cap, 5), np.int8)
    ln = np.zeros(41); ad = np.zeros(41, np.int32); dd = np.zeros(41)
    rc = np.zeros((1000, 4)); re_ = np.zeros((1000, 5), np.int8); n = np.zeros(3, np.int32)
    msg = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    r = L.simc_b200_read_optics_files(fwd.encode(), rec.encode(), cap, 1000, p(cs), p(fc), p(fe), p(ln), p(ad), p(dd), p(rc), p(re_),
                                      p(n), msg, 512)
    assert r == 0, msg.value.decode()
    nc, nf, nr = (int(x) for x in n)
    return cs[:nc + 1], fc[:nf], fe[:nf], ln[:nc], ad[:nc], dd[:nc], rc[:nr], re_[:nr]


@pytest.mark.parametrize("arm", sorted(FILES))
def test_product_reader_on_the_reference_files(arm):
    fwd, rec, n_classes = FILES[arm]
    cs, fc, fe, ln, ad, dd, rc, re_ = product_read(os.path.join(REF, fwd), os.path.join(REF, rec))
    assert len(cs) - 1 == n_classes                 # mc_hms.f:188, mc_shms.f:344-345, mc_sos.f:158, mc_hrsl.f:151
    # (1) an independent reading of the Fortran formats
    pcs, pfc, pfe = read_forward_python(os.path.join(REF, fwd))
    prc, pre = read_recon_python(os.path.join(REF, rec))
    assert np.array_equal(cs, pcs) and np.array_equal(fe, pfe) and np.array_equal(re_, pre)
    assert np.array_equal(fc.view(np.int64), pfc.view(np.int64)), "forward coefficients differ in some bit"
    assert np.array_equal(rc.view(np.int64), prc.view(np.int64)), "reconstruction coefficients differ in some bit"
    # (2) the fixtures every GPU test and bench.py use
    t = load_optics_fixture(arm)
    assert np.array_equal(cs, np.asarray(t.class_start, np.int32))
    assert np.array_equal(fc.view(np.int64), np.ascontiguousarray(t.fwd_coeff, np.float64).view(np.int64))
    assert np.array_equal(fe, np.asarray(t.fwd_expon, np.int8))
    assert np.array_equal(rc.view(np.int64), np.ascontiguousarray(t.rec_coeff, np.float64).view(np.int64))
    assert np.array_equal(re_, np.asarray(t.rec_expon, np.int8))
    assert np.array_equal(ln, np.asarray(t.length_cm, np.float64))
    # drift classes: the extracted length equals the !LENGTH: comment where there is one (transp.f:399-454)
    for k in range(n_classes):
        if ad[k] and ln[k] > 0:
            assert abs(dd[k] - ln[k]) < 0.01, (k + 1, dd[k], ln[k])
    assert ad.sum() >= (1 if arm == 2 else 4)          # the SOS file has a single pure drift
