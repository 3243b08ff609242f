"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/stmqr_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import refapi as R
import stmqr_b200 as sq

HEADER = os.path.join(R.ROOT, "include", "stmqr_b200.h")


def declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"\b(stmqr_b200_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.skipif(not os.path.exists(sq.LIB_PATH), reason="libstmqr_b200.so not built")
def test_library_exports_every_declared_symbol():
    lib = C.CDLL(sq.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/stmqr_b200.h but not exported"
    assert set(sq.EXPORTS) == set(syms)


@pytest.mark.skipif(not os.path.exists(sq.LIB_PATH), reason="libstmqr_b200.so not built")
def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", sq.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


@pytest.mark.skipif(not os.path.exists(sq.DROPIN_PATH), reason="libstmqr_dropin.so not built")
def test_dropin_exports_reference_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", sq.DROPIN_PATH], capture_output=True, text=True).stdout
    names = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert "qr_factorize" in names                 # STMMQR/include/SparseQR.h:127
    assert "stmqr_b200_qr_factorize" in names
    und = subprocess.run(["nm", "-D", "--undefined-only", sq.DROPIN_PATH], capture_output=True, text=True).stdout
    # host code allocates through the reference's counted allocator and calls the C ABI only
    assert "SparseCore_malloc" in und and "stmqr_b200_factorize" in und
    assert "oracle" not in und.lower()


@pytest.mark.skipif(not os.path.exists(sq.LIB_PATH), reason="libstmqr_b200.so not built")
def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = sq.load_library()
    assert lib.stmqr_b200_device_count() == 0
    with pytest.raises(sq.EngineError):
        sq.Engine(0)


def test_product_does_not_reference_oracle():
    """The product tree must not import / link / call anything under oracle/."""
    pkg = R.PKG
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".c", ".h", ".py", ".sh")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "stmqr_oracle" not in txt and "libref_harness" not in txt, os.path.join(root, f)


def test_header_is_plain_c(tmp_path):
    """include/stmqr_b200.h is the FFI surface: it must compile as C99 on its own (no C++, no
    reference headers, no torch types)."""
    src = tmp_path / "t.c"
    src.write_text('#include "stmqr_b200.h"\nint main (void) { stmqr_numeric_info i ; (void) i ; return STMQR_OK ; }\n')
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                          "-I", os.path.join(R.ROOT, "include"), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def _static_link(tmp_path):
    """qrtest of the reference linked STATICALLY (its stock build: archives + driver) with the drop-in object in
    place of the one renamed symbol -- host/link_static.sh, INTEGRATION.md Option C."""
    import subprocess
    import refapi as R
    obj = os.path.join(R.REF_DIR, "obj")
    qrtest_c = "/root/reference/STMMQR/test/qrtest.c"
    if not (os.path.isdir(obj) and os.path.exists(qrtest_c)):
        pytest.skip("needs the reference sources and oracle/_ref/obj (this container)")
    blas = open(os.path.join(R.REF_DIR, "blas_path.txt")).read().strip()
    shim = os.path.join(R.ROOT, "oracle", "shim")
    env = dict(os.environ, DRIVER_CFLAGS=f"-include {shim}/tpsm_platform.h -I{shim}")
    out = subprocess.run(["bash", os.path.join(R.PKG, "host", "link_static.sh"), obj, qrtest_c, str(tmp_path), blas],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    return os.path.join(str(tmp_path), "qrtest_static")


def test_static_link_of_the_dropin(tmp_path):
    """the statically linked driver defines qr_factorize (drop-in) AND keeps the reference's own numeric phase as
    qr_factorize_cpu, chunk_getSettings and qr_larftb (SparseQR.h:137,:263; qr_panel still needs the latter)"""
    import subprocess
    exe = _static_link(tmp_path)
    syms = subprocess.run(["nm", exe], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in syms.splitlines() if len(ln.split()) == 3 and ln.split()[1] in "TtDdBbCc"}
    for s in ("qr_factorize", "qr_factorize_cpu", "stmqr_b200_qr_factorize", "chunk_getSettings", "qr_larftb", "SparseQR"):
        assert s in defined, s
    undefined = {ln.split()[-1] for ln in syms.splitlines() if len(ln.split()) == 2 and ln.split()[0] == "U"}
    assert "stmqr_b200_analyze" in undefined and "stmqr_b200_factorize_streamed" in undefined   # from libstmqr_b200.so
