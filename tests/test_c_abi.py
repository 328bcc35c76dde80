"""The C ABI from plain C: include/ezpz_b200.h compiles as C99 (-pedantic -Werror) and a C program solves
test_cases/tiny through libezpz_b200.so alone (tests/c_abi/solve_tiny.c)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ezpz_b200", "_lib")
SRC = os.path.join(ROOT, "tests", "c_abi", "solve_tiny.c")


def _build(tmp_path):
    exe = str(tmp_path / "solve_tiny")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
                    "-L" + LIBDIR, "-lezpz_b200", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe], check=True)
    return exe


def test_header_is_c99_and_host_side_runs_from_c(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "--no-device"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Problem size: 4 rows, 4 vars" in out.stdout and "C ABI ok (host only)" in out.stdout


def test_header_alone_compiles_as_c_and_cxx(tmp_path):
    """The header must stand alone (no include order dependency) in both languages."""
    for compiler, std, name in (("gcc", "-std=c99", "h.c"), ("g++", "-std=c++17", "h.cpp")):
        src = tmp_path / name
        src.write_text('#include "ezpz_b200.h"\nint main(void) { return sizeof(ezpz_constraint_t) == 64 ? 0 : 1; }\n')
        subprocess.run([compiler, std, "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o",
                        str(tmp_path / (name + ".out"))], check=True)
        assert subprocess.run([str(tmp_path / (name + ".out"))]).returncode == 0


@pytest.mark.gpu
def test_c_program_solves_tiny_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "Problem size: 4 rows, 4 vars" in out.stdout
    assert "Iterations needed:" in out.stdout and out.stdout.strip().endswith("C ABI ok")
