"""FastDiv (csrc/ng_common.h): floor(n / d) as one wide multiply and a shift must be exact for every
0 <= n < 2^31 — the BatchNorm and pooling kernels index with it.  A small host program (the header is
plain C++ under -DNG_EMU) checks edge divisors (1, 2, 3, powers of two and their neighbours, 2^31 - 1)
against edge and random dividends, plus the shapes the kernels actually pass (HW/4 and C of the VGG
layers)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r"""
#include "ng_emu.h"
#include "ng_common.h"
#include <cstdio>
#include <cstdint>
#include <random>
int main() {
  std::mt19937_64 rng(1);
  const uint32_t edge_d[] = {1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 255, 256, 257, 1000,
                             1023, 1024, 1025, 4095, 4096, 65535, 65536, 65537, 1u << 20, (1u << 20) + 1, 1u << 30,
                             (1u << 30) + 1, (1u << 31) - 1};
  long long checked = 0;
  for (uint32_t d : edge_d) {
    const FastDiv f = make_fastdiv(d);
    if (f.d != d) { printf("make_fastdiv(%u) not set up\n", d); return 1; }
    const uint32_t edge_n[] = {0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1u << 31) - 1, (1u << 31) - 2, (1u << 30), 123456789u};
    for (uint32_t n : edge_n) {
      if (n >= (1u << 31)) continue;
      if (fastdiv(n, f) != n / d) { printf("fastdiv(%u, %u) = %u, want %u\n", n, d, fastdiv(n, f), n / d); return 1; }
      ++checked;
    }
    for (int i = 0; i < 200000; ++i) {
      const uint32_t n = (uint32_t)(rng() >> 33);
      if (fastdiv(n, f) != n / d) { printf("fastdiv(%u, %u) = %u, want %u\n", n, d, fastdiv(n, f), n / d); return 1; }
      ++checked;
    }
  }
  for (int i = 0; i < 2000; ++i) {   // random divisors
    const uint32_t d = (uint32_t)(rng() % ((1ull << 31) - 1)) + 1;
    const FastDiv f = make_fastdiv(d);
    for (int j = 0; j < 500; ++j) {
      const uint32_t n = (uint32_t)(rng() >> 33);
      if (fastdiv(n, f) != n / d) { printf("fastdiv(%u, %u) = %u, want %u\n", n, d, fastdiv(n, f), n / d); return 1; }
      ++checked;
    }
  }
  if (make_fastdiv(0).d != 0 || make_fastdiv(1LL << 31).d != 0) { printf("out-of-range divisors must stay unset\n"); return 1; }
  printf("ok %lld\n", checked);
  return 0;
}
"""


def test_fastdiv_is_exact_below_2_pow_31(tmp_path):
    src = tmp_path / "fastdiv_check.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "fastdiv_check"
    subprocess.run(["g++", "-std=c++20", "-O2", "-DNG_EMU", "-I", os.path.join(ROOT, "tests", "emu"),
                    "-I", os.path.join(ROOT, "nanograd_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok"), out.stdout
