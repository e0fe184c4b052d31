"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/opvd.h
declares, refuses to compute without a GPU (no CPU fallback), and the CLI honours the process contract."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import opv_cxx_demod_b200 as pkg

    if not (os.path.exists(pkg.LIB_PATH) and os.path.exists(pkg.CLI_PATH)):
        pkg.build()
    return pkg


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "opvd.h")).read()
    declared = sorted(set(re.findall(r"\b(opvd_[a-z0-9_]+)\s*\(", hdr)))
    from opv_cxx_demod_b200 import capi

    assert declared == sorted(capi.EXPORTS), "capi.EXPORTS must list exactly what include/opvd.h declares"
    L = C.CDLL(built.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), f"libopvd.so does not export {sym}"


def test_struct_layouts_match_header(built, tmp_path):
    """sizeof of every ABI struct as gcc sees include/opvd.h == the ctypes mirror in capi.py"""
    from opv_cxx_demod_b200 import capi

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "opvd.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(opvd_config),sizeof(opvd_event),sizeof(opvd_frame_info),sizeof(opvd_stream_info),'
                   'sizeof(opvd_synth));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.Config), C.sizeof(capi.Event), C.sizeof(capi.FrameInfo),
                     C.sizeof(capi.StreamInfo), C.sizeof(capi.Synth)]


def test_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path cannot be observed")
    from opv_cxx_demod_b200 import DemodBank, OpvdError

    with pytest.raises(OpvdError, match="no CPU fallback"):
        DemodBank(1, max_samples=1024)
    # the drop-in CLI must fail loudly as well (exit 2, not 0/1 which mean "frames"/"no frames")
    p = subprocess.run([built.CLI_PATH, "-r", "-q"], input=b"\0" * 4000, capture_output=True)
    assert p.returncode == 2 and b"no CPU fallback" in p.stderr and p.stdout == b""


def test_bank_cli_no_cpu_fallback_and_help(built, tmp_path):
    """the multi-stream front-end (opv-demod-bank) obeys the same rules: help on -h, exit 2 without a device"""
    import torch

    assert os.path.exists(built.BANK_CLI_PATH)
    p = subprocess.run([built.BANK_CLI_PATH, "-h"], capture_output=True)
    assert p.returncode == 0 and b"INPUT" in p.stderr and b"-s" in p.stderr and b"--devices" in p.stderr
    p = subprocess.run([built.BANK_CLI_PATH], capture_output=True)
    assert p.returncode == 2 and b"no inputs" in p.stderr
    if torch.cuda.is_available():
        return
    f = tmp_path / "a.iq"
    f.write_bytes(b"\0" * 4000)
    p = subprocess.run([built.BANK_CLI_PATH, "-s", "-q", str(f)], capture_output=True)
    assert p.returncode == 2 and b"no CPU fallback" in p.stderr
    assert not os.path.exists(str(f) + ".frames") or os.path.getsize(str(f) + ".frames") == 0


def test_integration_md_stub_compiles_against_header(built, tmp_path):
    """The binding INTEGRATION.md shows a reference maintainer is real code: it compiles and links against
    include/opvd.h + libopvd.so as written, and fails loudly (no frames, non-zero exit) without a device."""
    import torch

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```cpp\n(.*?)```", doc, re.S).group(1)
    src = tmp_path / "main.cpp"
    src.write_text(code)
    exe = tmp_path / "main"
    libdir = os.path.dirname(built.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", str(src), "-I", os.path.join(ROOT, "include"),
                    "-L", libdir, "-lopvd", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    if not torch.cuda.is_available():
        p = subprocess.run([str(exe)], input=b"\0" * 4000, capture_output=True)
        assert p.returncode != 0 and p.stdout == b"" and b"no CPU fallback" in p.stderr


def test_cli_help_contract(built):
    p = subprocess.run([built.CLI_PATH, "-h"], capture_output=True)
    assert p.returncode == 0 and b"-s" in p.stderr and b"-r" in p.stderr and b"-o <hz>" in p.stderr


def test_product_never_imports_oracle():
    # the oracle is test infrastructure: nothing under the product package may reference it
    pkg = os.path.join(ROOT, "opv_cxx_demod_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
