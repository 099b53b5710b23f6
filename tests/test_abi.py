"""The C-ABI library loads and exports every symbol include/fs2b200.h declares (no GPU needed)."""
import ctypes
import os
import subprocess

import pytest

from fs2b200 import sub


def test_library_exports_every_declared_symbol():
    cabi = sub("_cabi")
    protos = cabi.parse_header()
    assert len(protos) >= 30
    if not os.path.exists(cabi.so_path()):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(cabi.so_path())
    for name in protos:
        assert hasattr(lib, name), "missing export: %s" % name
    lib.fs2_version.restype = ctypes.c_int
    assert lib.fs2_version() == 2
    lib.fs2_launch_count.restype = ctypes.c_int64
    assert lib.fs2_launch_count() == 0  # nothing launched without a GPU


def test_no_extra_exports_outside_header():
    cabi = sub("_cabi")
    out = subprocess.run(["nm", "-D", "--defined-only", cabi.so_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("fs2_")}
    assert exported == set(cabi.parse_header()), exported ^ set(cabi.parse_header())


def test_struct_layout_matches_header():
    """sizeof(fs2_gemm) seen by ctypes == sizeof seen by the C compiler."""
    cabi = sub("_cabi")
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu", sizeof(fs2_operand), sizeof(fs2_gemm));}' \
        % cabi.header_path()
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", c, "-o", exe], check=True)
        a, b = subprocess.run([exe], capture_output=True, text=True).stdout.split()
    assert int(a) == ctypes.sizeof(cabi.Operand) and int(b) == ctypes.sizeof(cabi.Gemm)


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    M = sub("transformer")
    from oracle import synth

    enc = M.Encoder2(synth.model_cfg())
    x = torch.zeros(1, 4, 256)
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(x, torch.zeros(1, 4, dtype=torch.bool))
