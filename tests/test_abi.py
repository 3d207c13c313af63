"""The C-ABI boundary: libicpb200.so loads, exports every symbol include/icpb200.h declares, the
header is plain C, POD layouts match the bindings, and nothing computes without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import icpb200

HEADER = icpb200.HEADER_PATH


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(icpb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = icpb200.load()
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.icpb_version() == 100


def test_no_undeclared_icpb_exports():
    out = subprocess.check_output(["nm", "-D", "--defined-only", icpb200.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r"\b(icpb_[a-z0-9_]+)\b", out)))
    assert set(exported) == set(_declared())


def test_header_is_plain_c_and_layouts_match(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "icpb200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu\\n\", sizeof(icpb_point), sizeof(icpb_intrinsics),"
        " sizeof(icpb_icp_params), sizeof(icpb_icp_result), offsetof(icpb_icp_params, idx_trace),"
        " offsetof(icpb_icp_result, pose_R)); return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER),
                           str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [icpb200.POINT_DTYPE.itemsize, C.sizeof(icpb200.Intrinsics), C.sizeof(icpb200.IcpParams),
            C.sizeof(icpb200.IcpResult), icpb200.IcpParams.idx_trace.offset, icpb200.IcpResult.pose_R.offset]
    assert got == want and got[0] == 16


def test_no_cpu_fallback_without_a_gpu():
    lib = icpb200.load()
    if icpb200.device_count() > 0:
        pytest.skip("a GPU is present; the no-device path cannot be exercised here")
    h = C.c_void_p()
    rc = lib.icpb_ctx_create(0, C.byref(h))
    assert rc == icpb200.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.icpb_last_error(None)
    with pytest.raises(icpb200.IcpbError):
        icpb200.Context(0)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(icpb200, "_lib", None)
    monkeypatch.setattr(icpb200, "LIB_PATH", "/nonexistent/libicpb200.so")
    with pytest.raises(icpb200.IcpbError) as e:
        icpb200.load()
    assert "no CPU fallback" in str(e.value)


def test_status_strings():
    lib = icpb200.load()
    assert lib.icpb_status_string(0) == b"ok"
    assert lib.icpb_status_string(icpb200.ERR_EMPTY) == b"empty cloud"


def test_product_sources_never_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may name it."""
    root = os.path.dirname(os.path.dirname(HEADER))
    bad = []
    for base in [os.path.join(root, "icp-slam-prototype_b200", "csrc"), os.path.join(root, "include"),
                 os.path.join(root, "icp-slam-prototype_b200", "host"),
                 os.path.join(root, "icp-slam-prototype_b200", "python")]:
        for dp, _, fns in os.walk(base):
            for fn in fns:
                if fn.endswith((".cu", ".h", ".hpp", ".cpp", ".py", ".cuh")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"\borc_[a-z]|icp_oracle|from oracle|import oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
