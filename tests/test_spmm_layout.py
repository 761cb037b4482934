"""CPU replay of the shared-memory layout of spmm_tma_kernel (no GPU needed): tests/spmm_layout_emulation.cpp places
a block slab and an X tile as the TMA engine does under the 128-byte swizzle, walks the consumers' fragment
addressing with the very index functions the kernel compiles (csrc/spmm_layout.h), applies the DMMA fragment
semantics and the epilogue's column mapping, and compares with a plain product — for Float32 / Float64 /
ComplexF64, every right-hand-side tile width, full and partial slabs, N-form and T-form."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_layout_emulation(tmp_path):
    exe = tmp_path / "emul"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "tests" / "spmm_layout_emulation.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout[-2000:]
    # the design claim: Float64 fragment loads of full 32 x 32 slabs are bank-conflict free
    for line in out.stdout.splitlines():
        if line.startswith("S=8 "):
            assert line.endswith("A 1, B 1"), line
