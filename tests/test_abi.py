"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol that
include/w2v2_b200.h declares, with the binding in _lib.py covering all of them (no compute calls)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "w2v2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(w2v2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from w2v2_speaker_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f"{s} declared in the header but not bound in _lib.py"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in include/w2v2_b200.h"
    assert lib.w2v2_abi_version() >= 1


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from w2v2_speaker_b200 import ops
    from w2v2_speaker_b200._lib import W2V2Error
    with pytest.raises(W2V2Error):
        ops.gemm_f16(torch.zeros(4, 64, dtype=torch.float16), torch.zeros(4, 64, dtype=torch.float16))


def _struct_fields(name):
    """Field names of `typedef struct {...} name;` in the header, in declaration order."""
    text = open(os.path.join(ROOT, "include", "w2v2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    body = re.search(r"typedef\s+struct\s+\w*\s*\{([^}]*)\}\s*" + name + r"\s*;", text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"[A-Za-z_]\w*", part)[-1])
    return fields


def test_argument_blocks_match_the_header_layout(tmp_path):
    """The ctypes / numpy mirrors of the header's structs (schedule.LayerFwdArgs / LayerBwdArgs, engine.WeightPrep.DT)
    have the size and the per-field offsets the C compiler gives the header's declarations."""
    import ctypes
    import shutil
    import subprocess
    import pytest
    from w2v2_speaker_b200 import schedule
    from w2v2_speaker_b200.engine import WeightPrep
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"w2v2_layer_fwd_args": schedule.LayerFwdArgs, "w2v2_layer_bwd_args": schedule.LayerBwdArgs,
               "w2v2_prep_job": WeightPrep.DT}
    lines = ["#include <stddef.h>", "#include <stdio.h>", '#include "w2v2_b200.h"', "int main(void) {"]
    for s in structs:
        lines.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for f in _struct_fields(s):
            lines.append(f'  printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run([cc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for s, mirror in structs.items():
        names = _struct_fields(s)
        if isinstance(mirror, type) and issubclass(mirror, ctypes.Structure):
            assert [n for n, _ in mirror._fields_] == names, s
            assert ctypes.sizeof(mirror) == int(out[s]), s
            for n in names:
                assert getattr(mirror, n).offset == int(out[f"{s}.{n}"]), f"{s}.{n}"
        else:
            assert list(mirror.names) == names, s
            assert mirror.itemsize == int(out[s]), s
            for n in names:
                assert mirror.fields[n][1] == int(out[f"{s}.{n}"]), f"{s}.{n}"


_C2CTYPES = {"int": "c_int", "int64_t": "c_int64", "uint64_t": "c_uint64", "float": "c_float", "double": "c_double",
             "unsigned": "c_uint", "uint32_t": "c_uint32", "size_t": "c_size_t"}


def test_binding_types_match_the_header():
    """Every prototype in the header has the parameter list (count AND scalar widths: int vs int64_t vs float) and the
    return type its ctypes signature in _lib.py declares; a wrong width would silently corrupt the following arguments."""
    import ctypes
    from w2v2_speaker_b200 import _lib
    text = open(os.path.join(ROOT, "include", "w2v2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b([a-z_0-9 ]+?[\s\*]+)(w2v2_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert {n for _, n, _ in protos} == set(_lib.SIGNATURES)
    for ret, name, params in protos:
        res, bound = _lib.SIGNATURES[name]
        params = [p.strip() for p in params.split(",") if p.strip() not in ("", "void")]
        assert len(params) == len(bound), f"{name}: header has {len(params)} parameters, _lib.py binds {len(bound)}"
        for i, (p, b) in enumerate(zip(params, bound)):
            if "*" in p:
                assert b in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(b, "contents") or b.__name__.startswith("LP_"), (name, i, p)
            else:
                ty = [t for t in re.findall(r"[A-Za-z_]\w*", p) if t != "const"][0]
                assert b is getattr(ctypes, _C2CTYPES[ty]), f"{name} arg {i} `{p}` bound as {b.__name__}"
        ret = ret.strip()
        want = ctypes.c_char_p if "char" in ret else getattr(ctypes, _C2CTYPES[ret])
        assert res is want, f"{name}: returns `{ret}`, bound as {res}"
