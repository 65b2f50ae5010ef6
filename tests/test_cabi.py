"""The C-ABI library loads on a CPU-only box and exports every symbol include/jatts_b200.h declares.
No compute is attempted here (no GPU)."""
import ctypes as C
import os
import re

from jatts_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "jatts_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"JATTS_API\s+[\w\s\*]+?\b(jatts_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in jatts_b200.h but not exported"


def test_abi_version_and_launch_counter():
    assert _lib.lib.jatts_abi_version() == 1
    assert _lib.launch_count() >= 0


def test_struct_layouts_match_the_header():
    # field order/count of the ctypes mirrors vs the header's struct bodies
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)

    def fields(struct):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + struct + ";", src).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"(\w+)\s*(?:\[\d+\])*\s*$", part.strip())[0])
        return names

    assert fields("jatts_fs2_config") == [f[0] for f in _lib.Fs2Config._fields_]
    assert fields("jatts_hifigan_config") == [f[0] for f in _lib.HifiganConfig._fields_]
    assert fields("jatts_conv_gemm_args") == [f[0] for f in _lib.ConvGemmArgs._fields_]
    assert fields("jatts_tensor") == [f[0] for f in _lib.Tensor._fields_]
    assert fields("jatts_matcha_config") == [f[0] for f in _lib.MatchaConfig._fields_]
    assert fields("jatts_mrf_pair_args") == [f[0] for f in _lib.MrfPairArgs._fields_]
    assert fields("jatts_relpos_attention_args") == [f[0] for f in _lib.RelposAttentionArgs._fields_]


def test_struct_sizes_and_offsets_match_a_c_compiler(tmp_path):
    """the header compiled as plain C (it is the drop-in boundary: no C++ in the signatures): sizeof and the offset of every
    field of every struct, as gcc lays them out, equal the ctypes mirrors'"""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        import pytest

        pytest.skip("no C compiler")
    structs = {"jatts_tensor": _lib.Tensor, "jatts_fs2_config": _lib.Fs2Config, "jatts_matcha_config": _lib.MatchaConfig,
               "jatts_hifigan_config": _lib.HifiganConfig, "jatts_conv_gemm_args": _lib.ConvGemmArgs,
               "jatts_mrf_pair_args": _lib.MrfPairArgs, "jatts_relpos_attention_args": _lib.RelposAttentionArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, mirror in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    (tmp_path / "layout.c").write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(tmp_path / "layout"), str(tmp_path / "layout.c")], check=True)
    got = dict(line.split() for line in subprocess.run([str(tmp_path / "layout")], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, mirror in structs.items():
        assert int(got[cname]) == C.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(mirror, fname).offset, f"{cname}.{fname}"


def test_errors_are_reported_not_swallowed():
    rc = _lib.lib.jatts_fs2_create(None, None, 0, None)
    assert rc == -2
    assert b"null" in _lib.lib.jatts_last_error()
    try:
        _lib.check(rc, "fs2_create")
    except ValueError as e:
        assert "fs2_create" in str(e)
    else:
        raise AssertionError("check() must raise")
