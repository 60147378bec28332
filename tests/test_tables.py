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
needs_ref = pytest.mark.skipif(not os.path.exists(reftool.REF_LIB), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
def test_generated_headers_are_current(tmp_path):
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(ROOT, "tools", "gen_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.OUT = str(tmp_path / "mp2_tables.h")  # the psy-2 header is written next to it
    gen.main()
    for name in ("mp2_tables.h", "mp2_psy2_tables.h"):
        assert open(tmp_path / name).read() == open(os.path.join(CSRC, name)).read(), name


def test_psy2_start_up_tables_equal_the_oracle_restatement(tmp_path):
    """The product takes psy model 2's start-up tables (and psy model 0's threshold minima) frozen from the compiled
    reference (mp2_psy2_tables.h, tools/gen_tables.py through oracle/psy2_tap.c); the oracle restates psycho_2_init
    with libm (oracle/mp2_psy2_init.h).  Both must hold the same bits, for every sample rate the encoder accepts."""
    import subprocess
    src = r'''
#include <stdio.h>
#include <string.h>
#include "mp2_psy2_tables.h"
#include "mp2_psy2_init.h"
int main(void){ static mp2_psy2_tables T; int bad = 0;
  for (int r = 0; r < MP2_P2_RATES; r++) { double ath[32];
    if (mp2_psy2_init(&T, (double)MP2_P2_RATE[r])) return 1;
    mp2_psy0_init(ath, (double)MP2_P2_RATE[r]);
    bad += T.absthr_table != MP2_P2_ABSTHR_TABLE[r];
    for (int i = 0; i < 513; i++) bad += T.partition[i] != MP2_P2_PARTITION[r][i];
    for (int j = 0; j < 64; j++) { bad += T.numlines[j] != MP2_P2_NUMLINES[r][j];
      bad += memcmp(&T.tmn[j], &MP2_P2_TMN[r][j], 8) != 0; bad += memcmp(&T.rnorm[j], &MP2_P2_RNORM[r][j], 8) != 0;
      bad += memcmp(&T.bmax_of[j], &MP2_P2_BMAX_OF[r][j], 8) != 0;
      for (int k = 0; k < 64; k++) bad += memcmp(&T.s[j][k], &MP2_P2_ST[r][k][j], 8) != 0; }
    for (int j = 0; j <= 64; j++) bad += T.first_line[j] != MP2_P2_FIRST_LINE[r][j];
    for (int j = 0; j < 32; j++) bad += memcmp(&ath[j], &MP2_P0_ATH_MIN[r][j], 8) != 0;
    int lines = 0; for (int p = 0; p < T.n_part; p++) lines += T.numlines[p];
    printf("%d %d %d\n", T.n_part, lines, T.first_line[T.n_part]); }
  printf("bad %d\n", bad); return 0; }
'''
    open(tmp_path / "t.c", "w").write(src)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-I" + CSRC, "-I" + os.path.join(ROOT, "oracle"), "-o", str(tmp_path / "t"),
                    str(tmp_path / "t.c"), "-lm"], check=True)
    out = subprocess.run([str(tmp_path / "t")], capture_output=True, text=True, check=True).stdout.split("\n")
    for line in out[:4]:
        n_part, lines, end = map(int, line.split())
        assert 40 < n_part <= 64 and lines == 513 and end == 513
    assert out[4] == "bad 0", out
