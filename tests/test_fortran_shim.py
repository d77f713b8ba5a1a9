"""The Fortran side of the boundary is never compiled here (no Fortran compiler in the image), so its interface block is
checked mechanically against include/mohid_adt.h: every exported function is bound under its own name, with the same
number of arguments, each passed the way the C prototype expects (a C pointer parameter is either a Fortran dummy passed
by reference -- scalar or assumed-size array of the matching base type -- or a `type(c_ptr), value`), and the bind(c)
derived types list the fields of the C structs in the same order with matching types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mohid_adt.h")
SHIM = os.path.join(ROOT, "fortran", "ModuleAdvectionDiffusionB200.F90")

C_BASE = {"int": "integer(c_int)", "double": "real(c_double)", "char": "character(kind=c_char)",
          "long long": "integer(c_long_long)", "mohid_adt_size3d": "type(T_AdtSize3D)",
          "mohid_adt_options": "type(T_AdtOptions)", "mohid_adt_params": "type(T_AdtParams)"}


def c_prototypes():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+(mohid_adt_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        params = []
        for a in " ".join(m.group(2).split()).split(","):
            a = a.strip()
            name = re.search(r"(\w+)$", a).group(1)
            typ = a[: -len(name)].replace("const", " ").strip()
            depth = typ.count("*")
            base = " ".join(typ.replace("*", " ").split())
            params.append((name, base, depth))
        out[m.group(1)] = params
    return out


def fortran_interfaces():
    src = open(SHIM).read()
    src = re.sub(r"&\s*\n\s*&?", " ", src)                       # join continuation lines
    src = "\n".join(l.split("!")[0] for l in src.split("\n"))    # strip comments
    block = src[src.index("interface\n"): src.index("end interface")]
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(c,\s*name=\"(\w+)\"\)(.*?)end function", block, flags=re.S):
        fname, args, cname, body = m.group(1), m.group(2), m.group(3), m.group(4)
        args = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        for line in body.split("\n"):
            if "::" not in line or line.strip().startswith("import"):
                continue
            left, right = line.split("::")
            left = left.strip()
            base = left.split(",")[0].strip()
            if base.startswith("type("):
                base = left[: left.index(")") + 1]
            attrs = left[len(base):].lower()
            for nm in right.split(","):
                decl[nm.strip().lower()] = (base.replace(" ", ""), "value" in attrs, "dimension(*)" in attrs.replace(" ", ""))
        out[cname] = (fname, args, decl)
    return out, src


def test_every_export_is_bound_with_matching_arguments():
    protos = c_prototypes()
    ifaces, _ = fortran_interfaces()
    assert len(protos) >= 40
    missing = sorted(set(protos) - set(ifaces))
    assert not missing, f"not bound in the Fortran shim: {missing}"
    extra = sorted(set(ifaces) - set(protos))
    assert not extra, f"bound but not declared in the header: {extra}"
    for cname, params in protos.items():
        fname, args, decl = ifaces[cname]
        assert fname == cname
        assert len(args) == len(params), f"{cname}: {len(args)} Fortran dummies for {len(params)} C parameters"
        for (pname, base, depth), a in zip(params, args):
            assert a.lower() == pname.lower(), f"{cname}: dummy {a} at the position of C parameter {pname}"
            ftype, by_value, _ = decl[a.lower()]
            assert depth >= 1, f"{cname}.{pname}: the C-ABI passes everything by address"
            if by_value:
                assert ftype == "type(c_ptr)", f"{cname}.{pname}: only an address may be passed by value"
            elif depth == 2 or base == "void":
                # pointer to pointer (array of addresses / out address) or raw bytes by reference
                assert ftype in ("type(c_ptr)", "character(kind=c_char)"), f"{cname}.{pname}: {ftype} for {base}{'*' * depth}"
            else:
                want = C_BASE[base].replace(" ", "")
                assert ftype.lower() == want.lower(), f"{cname}.{pname}: {ftype} for {base} *"


def _c_struct(name):
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), h, flags=re.S).group(1)
    fields = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        typ, names = stmt.split(" ", 1)
        for nm in names.split(","):
            nm = nm.strip()
            dim = re.search(r"\[(\d+)\]", nm)
            fields.append((re.sub(r"\[.*\]", "", nm), typ, int(dim.group(1)) if dim else 0))
    return fields


def _f_type(src, name):
    body = re.search(r"type, bind\(c\) :: %s\n(.*?)end type" % name, src, flags=re.S).group(1)
    fields = []
    for line in body.split("\n"):
        if "::" not in line:
            continue
        left, right = line.split("::")
        for nm in right.split(","):
            nm = nm.strip()
            dim = re.search(r"\((\d+)\)", nm)
            fields.append((re.sub(r"\(.*\)", "", nm), left.strip(), int(dim.group(1)) if dim else 0))
    return fields


def test_bind_c_types_mirror_the_structs():
    _, src = fortran_interfaces()
    for cname, fname in (("mohid_adt_size3d", "T_AdtSize3D"), ("mohid_adt_params", "T_AdtParams"), ("mohid_adt_options", "T_AdtOptions")):
        c, f = _c_struct(cname), _f_type(src, fname)
        assert [x[0] for x in c] == [x[0] for x in f], f"{fname}: field order differs from {cname}"
        for (n, ctyp, cdim), (_, ftyp, fdim) in zip(c, f):
            assert C_BASE[ctyp].replace(" ", "") == ftyp.replace(" ", ""), f"{fname}%{n}: {ftyp} for {ctyp}"
            assert cdim == fdim, f"{fname}%{n}: array extent"


def test_wrappers_cover_the_reference_entry_points():
    """The public procedures of ModuleAdvectionDiffusion (AD:400, 978, 1040, 697, 775, 1108, 5849) have a B200_* wrapper."""
    _, src = fortran_interfaces()
    for w in ("B200_Start", "B200_Kill", "B200_SetStep", "B200_AdvectBatch", "B200_SetDischarges", "B200_UnSetDischarges",
              "B200_GetAdvFlux", "B200_GetDifFlux", "B200_CommInit", "B200_ExchangeHalos"):
        assert re.search(r"subroutine %s\b" % w, src), w
