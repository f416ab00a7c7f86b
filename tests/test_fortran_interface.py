"""fortran/mglc_iso_c.f90 cannot be compiled in this image (no Fortran compiler), so its interface blocks are checked textually
against include/mglc.h and libmglc.so: every bind(C) name is an exported entry point, the dummy-argument count equals the C
prototype's parameter count, every dummy argument is declared, and the bind(C) derived types have as many components as the C
structs they mirror."""
import ctypes as C
import os
import re

from mglc_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_prototypes():
    src = open(os.path.join(ROOT, "include", "mglc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(mglc_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",")])
    return protos, src


def fortran_functions():
    src = open(os.path.join(ROOT, "fortran", "mglc_iso_c.f90")).read()
    src = re.sub(r"&\s*\n\s*&?", " ", src)                                  # join continuation lines
    out = []
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)\s*result\((\w+)\)(.*?)end function", src, flags=re.S | re.I):
        name, args, cname, res, body = m.groups()
        out.append((name, [a.strip() for a in args.split(",") if a.strip()], cname, res, body))
    return out, src


def test_every_fortran_binding_matches_a_c_prototype():
    protos, _ = c_prototypes()
    funcs, _ = fortran_functions()
    assert len(funcs) >= 60
    lib = C.CDLL(L.LIB_PATH)
    for name, args, cname, res, body in funcs:
        assert name == cname, (name, cname)
        assert cname in protos and hasattr(lib, cname), f"{cname}: bound in Fortran but not declared / exported"
        assert len(args) == protos[cname], f"{cname}: {len(args)} dummy arguments in Fortran, {protos[cname]} parameters in C"
        declared = set()
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" in line:
                rhs = re.sub(r"\([^)]*\)", "", line.split("::", 1)[1])          # drop array specs like f(*), dims(2)
                declared |= {v.strip().lower() for v in rhs.split(",") if v.strip()}
        for a in args + [res]:
            assert a.lower() in declared, f"{cname}: dummy argument {a} is not declared"


def test_derived_types_mirror_the_c_structs():
    _, csrc = c_prototypes()
    _, fsrc = fortran_functions()

    def c_fields(struct):
        body = re.search(r"typedef struct " + struct + r"\s*\{(.*?)\}\s*" + struct + r"\s*;", csrc, flags=re.S).group(1)
        n = 0
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if stmt:
                n += len(stmt.split(","))
        return n

    def f_fields(struct):
        body = re.search(r"type,\s*bind\(C\)\s*::\s*" + struct + r"\b(.*?)end type", fsrc, flags=re.S | re.I).group(1)
        n = 0
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" in line:
                n += len(re.sub(r"\([^)]*\)", "", line.split("::", 1)[1]).split(","))
        return n
    for struct in ("mglc_lbm_desc", "mglc_p2d_desc", "mglc_l2d_desc", "mglc_t2d_desc", "mglc_aa_desc"):
        assert c_fields(struct) == f_fields(struct), (struct, c_fields(struct), f_fields(struct))


def c_constants():
    """MGLC_* integer constants of include/mglc.h: #define NAME value and enum { NAME = value, NAME, ... } (implicit values count up)"""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mglc.h")).read(), flags=re.S)
    vals = {}
    for m in re.finditer(r"#define\s+(MGLC_[A-Z0-9_]+)\s+\(?(-?\d+)\)?\s*$", src, flags=re.M):
        vals[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"enum\s*\w*\s*\{([^}]*)\}", src):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            name, _, v = (s.strip() for s in item.partition("="))
            nxt = int(v, 0) if v else nxt
            vals[name] = nxt
            nxt += 1
    return vals


def test_fortran_constants_equal_the_c_header():
    c = c_constants()
    src = open(os.path.join(ROOT, "fortran", "mglc_iso_c.f90")).read()
    src = "\n".join(l.split("!")[0] for l in src.splitlines())
    seen = 0
    for m in re.finditer(r"integer\(c_int\),\s*parameter\s*::\s*(.*)", src):
        for item in m.group(1).split(","):
            name, _, v = (s.strip() for s in item.partition("="))
            assert name in c, f"{name} is not a constant of include/mglc.h"
            assert c[name] == int(v), (name, c[name], v)
            seen += 1
    assert seen >= 15
    for must in ("MGLC_L2D_C", "MGLC_L2D_F", "MGLC_L2D_INCOMP", "MGLC_L2D_C_SRT", "MGLC_T2D_MPI", "MGLC_T2D_ACC", "MGLC_BCT_PERIODIC"):
        assert re.search(r"\b%s\b" % must, src), must
