"""The Fortran side of the boundary (integration/simc_b200_api.f90, integration/simc_b200_shim.f) cannot be compiled
in this image (no Fortran compiler), so its bind(C) derived types are checked here instead: the file is parsed, every
type is laid out with the C rules a `bind(C)` type follows, and names, order, array extents and byte offsets must equal
the ctypes mirror of include/simc_b200.h (which is itself checked against the C sizes when the library loads).  The
shim's pack / unpack routines are checked for completeness: every field of simc_run_config is assigned, every
accumulator is read, every interface names a symbol the library exports."""
import ctypes as C
import os
import re

import pytest

from simc_gfortran_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
API = os.path.join(ROOT, "integration", "simc_b200_api.f90")
SHIM = os.path.join(ROOT, "integration", "simc_b200_shim.f")

SCALARS = {"real(c_double)": (8, 8), "integer(c_int32_t)": (4, 4), "integer(c_int64_t)": (8, 8), "integer(c_int)": (4, 4)}
MIRROR = {"simc_cut": L.Cut, "simc_range": L.Range, "simc_arm_cuts": L.ArmCuts, "simc_arm_limits": L.ArmLimits,
          "simc_edge_arm": L.EdgeArm, "simc_edge": L.Edge, "simc_gen_limits": L.GenLimits,
          "simc_spectrometer": L.Spectrometer, "simc_axis": L.Axis, "simc_target": L.Target,
          "simc_run_config": L.RunConfig, "simc_fixed128": L.Fixed128, "simc_accum": L.Accum, "simc_results": L.Results}


def free_form_statements(path):
    text = open(path).read()
    text = re.sub(r"&\s*\n\s*&?", " ", text)                   # continuation lines
    return [re.sub(r"!.*", "", ln).strip() for ln in text.splitlines()]


def parse_types(path):
    """{type name: [(field, element type, [extents in Fortran order])]}"""
    types, cur = {}, None
    for ln in free_form_statements(path):
        m = re.match(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)", ln, re.I)
        if m:
            cur = types.setdefault(m.group(1), [])
            continue
        if re.match(r"end\s+type", ln, re.I):
            cur = None
            continue
        if cur is None or "::" not in ln:
            continue
        decl, names = ln.split("::", 1)
        decl = decl.strip().lower().replace(" ", "")
        if decl.startswith("type("):
            decl = "type:" + decl[5:-1]
        for item in re.findall(r"(\w+)\s*(\(([^)]*)\))?", names):
            ext = [int(x) for x in item[2].split(",")] if item[2] else []
            cur.append((item[0], decl, ext))
    return types


def layout(types, name, cache):
    """(size, align, [(field, offset, size)]) with the C rules (natural alignment, tail padding)."""
    if name in cache:
        return cache[name]
    off, align, fields = 0, 1, []
    for fname, decl, ext in types[name]:
        if decl.startswith("type:"):
            sz, al, _ = layout(types, next(t for t in types if t.lower() == decl[5:]), cache)
        else:
            sz, al = SCALARS[decl]
        n = 1
        for e in ext:
            n *= e
        off = (off + al - 1) // al * al
        fields.append((fname, off, sz * n))
        off += sz * n
        align = max(align, al)
    size = (off + align - 1) // align * align
    cache[name] = (size, align, fields)
    return cache[name]


def ctypes_extents(typ):
    ext = []
    while hasattr(typ, "_length_"):
        ext.append(typ._length_)
        typ = typ._type_
    return ext


def test_bind_c_types_match_the_c_layout():
    types = parse_types(API)
    assert set(MIRROR) <= set(types), sorted(set(MIRROR) - set(types))
    cache = {}
    for name, ctype in MIRROR.items():
        size, _, fields = layout(types, name, cache)
        assert size == C.sizeof(ctype), (name, size, C.sizeof(ctype))
        cf = [(n.rstrip("_"), getattr(ctype, n).offset, getattr(ctype, n).size) for n, _ in ctype._fields_]
        assert [(f.lower(), o, s) for f, o, s in fields] == [(n.lower(), o, s) for n, o, s in cf], name
        # array extents: Fortran lists them in the opposite order of C
        for (fname, decl, ext), (cname, ct) in zip(types[name], ctype._fields_):
            assert ext == list(reversed(ctypes_extents(ct))), (name, fname, ext, ctypes_extents(ct))
    assert layout(types, "simc_run_config", cache)[0] == L.load_library().simc_b200_sizeof(0)
    assert layout(types, "simc_accum", cache)[0] == L.load_library().simc_b200_sizeof(1)


def test_interfaces_name_exported_symbols():
    lib = L.load_library()
    names = re.findall(r"bind\(C,\s*name='(\w+)'\)", open(API).read())
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in simc_b200_api.f90 but not exported by libsimc_b200.so"
    header = open(os.path.join(ROOT, "include", "simc_b200.h")).read()
    for n in names:
        assert re.search(r"\b" + n + r"\s*\(", header), f"{n} is not declared in include/simc_b200.h"


def leaf_paths(ctype, prefix=""):
    out = []
    for n, t in ctype._fields_:
        base = t
        while hasattr(base, "_length_"):
            base = base._type_
        if hasattr(base, "_fields_"):
            out += leaf_paths(base, prefix + n + "%")
        else:
            out.append(prefix + n)
    return out


def test_pack_run_config_assigns_every_field():
    src = open(SHIM).read().lower()
    body = src[src.index("subroutine simc_b200_pack_run_config"):src.index("subroutine simc_b200_unpack_accum")]
    missing = []
    for path in leaf_paths(L.RunConfig):
        p = path.lower()
        if p.startswith(("edge%", "vertexedge%")):
            p = "c%" + p.split("%", 1)[1]                      # filled by simc_b200_pack_edge(c, e)
        elif p.startswith("hist_axis%"):
            p = "ax(1)%" + p.split("%", 1)[1]                  # filled by simc_b200_pack_axes(ax, hs)
        else:
            p = "cfg%" + p
        if not re.search(re.escape(p) + r"(\(\w+\))?\s*=", body):
            missing.append(path)
    assert not missing, missing
    # the three sets of axes and both edges are packed
    assert body.count("call simc_b200_pack_axes") == 3 and body.count("call simc_b200_pack_edge") == 2


def test_unpack_accum_reads_every_accumulator():
    src = open(SHIM).read().lower()
    body = src[src.index("subroutine simc_b200_unpack_accum"):src.index("subroutine simc_b200_widen")]
    for f in ("ntried", "nsuccess", "ncontribute", "npasscuts", "ncontribute_no_rad_proton", "wtcontribute", "sum_sigcc"):
        assert "acc%" + f in body, f
    for k in range(1, 9):
        assert f"acc%sumerr({k})" in body and f"acc%sumerr2({k})" in body
    for k in range(1, 7):
        assert f"acc%hist_w(i,{k})" in body
    # count histograms: RECON Em / Pm, gen 1..7 (gen%Pm is never filled, simc.f:270-286), geni 1..8
    used = set(re.findall(r"acc%hist_n\(i,(\d),(\d)\)", body))
    assert used == {("7", "1"), ("8", "1")} | {(str(k), "2") for k in range(1, 8)} | {(str(k), "3") for k in range(1, 9)}
    assert sorted(int(k) for k in re.findall(r"acc%contrib\((\d+)\)", body)) == list(range(1, 31))
    assert sorted(set(int(k) for k in re.findall(r"acc%slop\((\d+)\)%lo", body))) == list(range(1, 9))
    # STOP slots: every aperture code the library names for an arm is read in that arm's branch
    lib = L.load_library()
    for arm, tag in ((1, "arm.eq.1"), (2, "arm.eq.2"), (3, "arm.eq.3"), (4, "arm.eq.4"), (5, "arm.eq.5")):
        seg = body[body.index(tag):]
        seg = seg[:seg.index("else if") if "else if" in seg else seg.index("endif")]
        slots = set(int(k) for k in re.findall(r"acc%stop\((\d+),w\)", seg))
        n_codes = 0
        while n_codes < 60 and (lib.simc_b200_stop_name(arm, n_codes + 1) or b"").decode() not in ("", "?", "unknown"):
            n_codes += 1
        assert n_codes >= 14, (arm, n_codes)
        # cal (HMS, SOS, HRS) never stops a track in the reference and has no variable in some arms: allow the tail
        want = set(range(1, 4)) | set(range(4, 4 + n_codes))
        assert slots <= want and len(want - slots) <= 1, (arm, sorted(want - slots), sorted(slots - want))
