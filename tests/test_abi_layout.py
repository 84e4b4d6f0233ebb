"""Layout of the four descriptor structs of the C ABI, three ways (`-m "not gpu"`):
  * what the C compiler lays out for include/vela_b200.h (offsetof / sizeof, compiled here),
  * what Julia lays out for the mirror structs of vela.jl_b200/julia/VelaB200.jl (isbits structs follow the C rules:
    every field aligned to its own size, NTuple{N,T} = N consecutive T, Ptr = 8 bytes) - computed from the parsed
    Julia source, since Julia itself is not in this image,
  * what the ctypes mirror of vela.jl_b200/descriptor.py uses.
A field added or reordered in one of the three but not the others fails here instead of corrupting a create() call.
"""
import ctypes as C
import os
import re
import subprocess

from conftest import ROOT

STRUCTS = ["VelaComponentDesc", "VelaGPDesc", "VelaModelDesc", "VelaPriorDesc"]
JL_SIZES = {"Int32": 4, "UInt32": 4, "Int64": 8, "UInt64": 8, "Float64": 8, "Cint": 4, "UInt8": 1}


def c_layout(tmp_path):
    hdr = open(os.path.join(ROOT, "include", "vela_b200.h")).read()
    fields = {}
    for s in STRUCTS:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (s, s), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields[s] = []
        for decl in body.split(";"):
            for part in decl.split(","):     # `double a, b;` declares two fields
                if part.strip():
                    fields[s].append(re.search(r"(\w+)(\[[^\]]*\])?\s*$", part.strip()).group(1))
    src = ['#include <stdio.h>\n#include <stddef.h>\n#include "vela_b200.h"\nint main(void) {']
    for s in STRUCTS:
        src.append(f'printf("{s} sizeof %zu\\n", sizeof({s}));')
        for f in fields[s]:
            src.append(f'printf("{s} {f} %zu\\n", offsetof({s}, {f}));')
    src.append("return 0; }")
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(cfile)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    lay = {s: {} for s in STRUCTS}
    for ln in out.splitlines():
        s, f, v = ln.split()
        lay[s][f] = int(v)
    return lay, fields


def julia_layout():
    src = open(os.path.join(ROOT, "vela.jl_b200", "julia", "VelaB200.jl")).read()
    max_par = int(re.search(r"const VELA_MAX_PAR\s*=\s*(\d+)", src).group(1))
    lay = {}
    for s in STRUCTS:
        body = re.search(r"^struct %s\n(.*?)^end" % s, src, re.S | re.M).group(1)
        off, align_max, out = 0, 1, {}
        for ln in body.splitlines():
            m = re.match(r"\s*(\w+)::([^#]+)", ln)
            if not m:
                continue
            name, ty = m.group(1), m.group(2).strip()
            nt = re.match(r"NTuple\{(\w+),\s*(\w+)\}", ty)
            if nt:
                n = max_par if nt.group(1) == "VELA_MAX_PAR" else int(nt.group(1))
                el = JL_SIZES[nt.group(2)]
                size, align = n * el, el
            elif ty.startswith("Ptr{"):
                size = align = 8
            else:
                size = align = JL_SIZES[ty]
            off = (off + align - 1) // align * align
            out[name] = off
            off += size
            align_max = max(align_max, align)
        out["sizeof"] = (off + align_max - 1) // align_max * align_max
        lay[s] = out
    return lay


def test_c_julia_and_ctypes_layouts_agree(vb, tmp_path):
    c_lay, fields = c_layout(tmp_path)
    jl = julia_layout()
    D = vb.descriptor
    py = {"VelaComponentDesc": D.VelaComponentDesc, "VelaGPDesc": D.VelaGPDesc, "VelaModelDesc": D.VelaModelDesc,
          "VelaPriorDesc": D.VelaPriorDesc}
    for s in STRUCTS:
        assert jl[s] == c_lay[s], f"{s}: Julia mirror {jl[s]} != C header {c_lay[s]}"
        assert C.sizeof(py[s]) == c_lay[s]["sizeof"]
        assert [f for f, _ in py[s]._fields_] == fields[s]
        for f in fields[s]:
            assert getattr(py[s], f).offset == c_lay[s][f], f"{s}.{f}"
