"""CPU: the C-ABI library loads, exports every symbol include/mmdit_b200.h declares, and the
ctypes mirrors agree with the header (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from mmdit import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmdit_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b(?:int|int64_t|const char\*|unsigned long long)\s+(mmdit_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), f"{n} declared in mmdit_b200.h but not exported"
    assert L.mmdit_abi_version() == 1


def test_binding_table_matches_header():
    declared = set(_declared()) - {"mmdit_last_error", "mmdit_abi_version", "mmdit_device_check",
                                   "mmdit_launch_count"}
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, argtypes in _abi.SIGNATURES.items():
        m = re.search(rf"\b{name}\s*\((.*?)\)\s*;", src, flags=re.S)
        nargs = len([a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"])
        assert nargs == len(argtypes), (name, nargs, len(argtypes))


def test_struct_layouts_match_the_c_compiler(tmp_path):
    prog = tmp_path / "sz.cpp"
    prog.write_text('#include "mmdit_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                    'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(mmdit_gemm_args), sizeof(mmdit_attn_args),'
                    'offsetof(mmdit_gemm_args, remap_rows), offsetof(mmdit_attn_args, d_o), offsetof(mmdit_attn_args, delta));}')
    exe = tmp_path / "sz"
    subprocess.run(["g++", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(_lib.GemmArgs), ctypes.sizeof(_lib.AttnArgs), _lib.GemmArgs.remap_rows.offset,
                   _lib.AttnArgs.d_o.offset, _lib.AttnArgs.delta.offset]


def test_comm_struct_layout_matches_the_c_compiler(tmp_path):
    from mmdit.comm import CommStruct
    prog = tmp_path / "sz.cpp"
    prog.write_text('#include "mmdit_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                    'int main(){printf("%zu %zu %zu %zu\\n", sizeof(mmdit_comm), offsetof(mmdit_comm, flag),'
                    'offsetof(mmdit_comm, state), offsetof(mmdit_comm, world));}')
    exe = tmp_path / "sz"
    subprocess.run(["g++", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(CommStruct), CommStruct.flag.offset, CommStruct.state.offset,
                   CommStruct.world.offset]


def test_sass_is_blackwell_native():
    """tcgen05 / TMA / TMEM mnemonics must be in the shipped cubin (B200_PROFILING.md)."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    # UTMASTG: TMA-store epilogues; the CTA-pair GEMM shows as UTCHMMA.2CTA
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM"):
        assert mnemonic in out, mnemonic
    assert "HMMA." not in out.replace("UTCHMMA", "")  # no legacy mma.sync tensor path


def test_second_generation_row_kernels_are_in_the_cubin():
    """rowwise2.cu / qknorm2.cu: packed fp32x2 math (FFMA2 / FMUL2 / FADD2) and the thread-block-cluster
    barrier of the DSMEM column fold must be in the shipped SASS, next to the first-generation kernels."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    for kernel in ("ln_mod_fwd2_kernel", "ln_mod_bwd2_kernel", "gate_res_ln_fwd2_kernel", "gate_bwd2_kernel",
                   "qknorm_rope_fwd2_kernel", "qknorm_rope_bwd2_kernel", "ln_mod_bwd_kernel", "qknorm_rope_bwd_kernel"):
        assert kernel in out, kernel
    for mnemonic in ("FFMA2", "FMUL2", "FADD2", "UCGABAR_ARV", "UCGABAR_WAIT"):
        assert mnemonic in out, mnemonic


def test_argument_errors_are_reported_not_thrown():
    L = _lib.lib()
    a = _lib.GemmArgs()        # null pointers
    rc = L.mmdit_gemm_bf16(ctypes.byref(a), None)
    assert rc < 0 and b"null" in L.mmdit_last_error()
    with pytest.raises(RuntimeError, match="mmdit_gemm_bf16"):
        _lib.check(rc, "mmdit_gemm_bf16")
