"""CPU, build container only: the committed table headers are exactly what tools/gen_tables.py derives from the
reference compiled into oracle/_ref (analysis window, scalefactors, DCT matrix, Hann windows, add_db table, FHT
twiddles, critical bands, threshold tables, absolute thresholds)."""
import importlib.util
import os
import shutil

import pytest

import reftool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "odr_audioenc_b200", "csrc")
pytestmark = pytest.mark.skipif(not os.path.exists(reftool.REF_LIB), reason="oracle/_ref not built (needs /root/reference)")


def test_generated_headers_are_current(tmp_path):
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(ROOT, "tools", "gen_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.OUT = str(tmp_path / "mp2_tables.h")  # the psy-2 header is written next to it
    gen.main()
    for name in ("mp2_tables.h", "mp2_psy2_tables.h"):
        assert open(tmp_path / name).read() == open(os.path.join(CSRC, name)).read(), name


def test_psy2_start_up_tables_match_the_reference_dump():
    """mp2_psy2_init.h against the reference's own init: the SMRs of psy model 2 are bit-identical in
    tests/test_oracle_vs_ref.py, which they could not be with a wrong partition or spreading table; here only the
    cheap structural facts"""
    import ctypes as C
    import subprocess
    src = r'''
#include <stdio.h>
#include "mp2_psy2_init.h"
int main(void){ static mp2_psy2_tables T; for (int r = 0; r < 3; r++) { double fs[3] = {48000, 24000, 32000};
  if (mp2_psy2_init(&T, fs[r])) return 1; int lines = 0; for (int p = 0; p < T.n_part; p++) lines += T.numlines[p];
  printf("%d %d %d %d\n", T.n_part, lines, T.first_line[T.n_part], T.absthr_table); } return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-O2", "-I" + CSRC, "-o", os.path.join(td, "t"), os.path.join(td, "t.c"), "-lm"], check=True)
        out = subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split("\n")
    for line, table in zip(out[:3], (2, 2, 0)):
        n_part, lines, end, tab = map(int, line.split())
        assert 40 < n_part <= 64 and lines == 513 and end == 513 and tab == table
