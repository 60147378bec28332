/*
 * psy2_tap.c -- TEST / TABLE-GENERATION INFRASTRUCTURE ONLY (never linked into the product).
 * Compiled by oracle/Makefile together with the reference's psycho_2.c (which it #includes, unmodified, from
 * /root/reference/libtoolame-dab) so that the file-static start-up tables psycho_2_init builds (psycho_2.c:259-420)
 * can be read back.  tools/gen_tables.py freezes them into odr_audioenc_b200/csrc/mp2_psy2_tables.h.
 */
#include "psycho_2.c"

void psy2_tap_init(double sfreq_hz) { psycho_2_init(sfreq_hz); }
const int *psy2_tap_partition(void) { return partition; }   /* [HBLKSIZE = 513] */
const int *psy2_tap_numlines(void) { return numlines; }     /* [CBANDS = 64] */
const double *psy2_tap_cbval(void) { return cbval; }
const double *psy2_tap_rnorm(void) { return rnorm; }
const double *psy2_tap_tmn(void) { return tmn; }
const double *psy2_tap_s(void) { return &s[0][0]; }         /* [64][64] */
const double *psy2_tap_bmax(void) { return bmax; }          /* [27] */
const double *psy2_tap_absthr(void) { return absthr; }      /* [513], the table psycho_2_init picked */
