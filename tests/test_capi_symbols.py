"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every symbol that
include/pg_align.h declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pg_align.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    from paragraph_b200 import capi
    lib = capi.load()
    syms = declared_symbols()
    assert set(syms) == set(capi.SYMBOLS)
    for s in syms:
        assert getattr(lib, s) is not None, s
    assert b"sm_100a" in lib.pg_version()


def test_record_layout_matches_header(built):
    from paragraph_b200 import capi
    assert capi.RECORD_DTYPE.itemsize == 20


def test_library_contains_sm100a_code(built):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "paragraph_b200", "libpgalign.so")],
                         capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_gpu_fails_loudly(built):
    """Without a usable device the product must refuse, not fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from paragraph_b200 import capi
    with pytest.raises(capi.PgError):
        capi.Context(0)


def test_product_does_not_import_oracle():
    """paragraph_b200/ (the product) must never reach into oracle/ or the emulator."""
    pkg = os.path.join(ROOT, "paragraph_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".hh", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle/", "ORACLE_DIR_MENTION") \
                    or "import" not in src or "from oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src and "refbind" not in src, f
                assert "pgemu" not in src and "pgshim" not in src and "pg_shim_names" not in src, f


def test_read_length_limit_is_the_same_everywhere(built):
    """PG_MAX_READ_LEN of the header, of the Python binding, of the device source and of the version string agree."""
    from paragraph_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "pg_align.h")).read()
    core = open(os.path.join(ROOT, "paragraph_b200", "csrc", "pg_core.cuh")).read()
    h = int(re.search(r"#define\s+PG_MAX_READ_LEN\s+(\d+)", hdr).group(1))
    c = int(re.search(r"constexpr int MAX_READ_LEN = (\d+);", core).group(1))
    assert h == c == capi.MAX_READ_LEN
    assert ("reads<=%d" % h).encode() in capi.load().pg_version()
