"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol that include/pfd_b200.h declares (no compute calls are made here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pfd_b200.h")).read()
    return sorted(set(re.findall(r"PFD_API\s+[\w\s\*]+?\b(pfd_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "pfd_gemm_f16" in names and "pfd_groupnorm_f16" in names and len(names) >= 15


def test_library_exports_every_declared_symbol():
    from pfd_b200 import native
    lib = native.load()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/pfd_b200.h but not exported"
    assert lib.pfd_version() == 2


def test_python_binding_lists_match_header():
    from pfd_b200 import native
    assert sorted(native.EXPORTS) == _declared()


def test_gemm_desc_layout_matches_header():
    # the ctypes mirror must have the same size as the C struct (computed from the header fields)
    from pfd_b200 import native
    d = native.GemmDesc()
    # 1+3+3 int32 (=28, pad to 32) + 3 ptr + 9 int64 + 6 int32 + ptr + int32(+pad) + 2 int64 + float + int32
    # + 4 ptr + 6 int64 + 4 int32 (ndiv, cdiv, bn_force, tap_off) + stream ptr
    assert ctypes.sizeof(d) == 32 + 24 + 72 + 24 + 8 + 8 + 16 + 8 + 40 + 48 + 16 + 8
    # and the same field order as the header declares
    import re
    src = open(os.path.join(ROOT, "include", "pfd_b200.h")).read()
    start = "typedef struct pfd_gemm_desc {"
    body = src[src.index(start) + len(start):src.index("} pfd_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"(\w+)\s*(\[\w+\])?\s*$", part.strip())
            names.append(m.group(1))
    assert names == [f[0] for f in native.GemmDesc._fields_], (names, [f[0] for f in native.GemmDesc._fields_])


def test_bad_descriptor_is_rejected_without_gpu():
    from pfd_b200 import native
    lib = native.load()
    d = native.GemmDesc()
    d.nseg = 0
    assert lib.pfd_gemm_f16(ctypes.byref(d)) != 0
    assert b"nseg" in lib.pfd_last_error()
