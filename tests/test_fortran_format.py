"""The bulk writers format with fortran_format.hpp append_F/E/G (one 64-bit-mantissa multiplication per value for
G15.8 / E14.6, std::to_chars next to rounding boundaries and for everything else); the log and stat files with the
printf-based F/E/G that the golden-file tests pin.  Both must produce the same characters: tests/cpp/fmt_check.cpp
compares them on edge cases, ~0.5 M random values (G15.8, E14.6, G20.12, F15.6), 4 M more values aimed at the fast
path (any bit pattern, 8-digit decimals, near ties, powers of ten and their neighbours, subnormals) and a list of known
answers taken from the reference's golden files."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_formatters_match_printf_formatters(tmp_path):
    exe = str(tmp_path / "fmt_check")
    src = os.path.join(ROOT, "tests", "cpp", "fmt_check.cpp")
    inc = os.path.join(ROOT, "netl-ap-map-flow_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, src, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout
