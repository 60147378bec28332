/*
 * ref_tap.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Builds the UNMODIFIED reference encoder driver from the sources where they
 * lie under /root/reference (see oracle/Makefile, target _ref).  This
 * translation unit textually includes the reference's libtoolame-dab/toolame.c
 * so that the file-static per-frame state declared there
 * (toolame.c:89-115: sb_sample, j_sample, subband, scalar, j_scale, smr,
 * max_sc, scfsi, bit_alloc, header, frame) can be read back after every
 * toolame_encode_frame() call.  Nothing of the reference is copied into this
 * repository: the include is resolved at compile time through -I.
 *
 * The accessors below only READ the state; the encoder's behaviour is the
 * reference's own.
 */
#include "toolame.c"

#include <stdint.h>

/* One tap record per encoded frame; layout mirrored by tests/reftool.py. */
typedef struct {
    int32_t mode, mode_ext, jsbound, sblimit, nch, tablenum, bitrate_index, dab_extension;
    uint32_t scalar[2][3][SBLIMIT];   /* after sf_transmission_pattern */
    uint32_t j_scale[3][SBLIMIT];
    uint32_t scfsi[2][SBLIMIT];
    uint32_t bit_alloc[2][SBLIMIT];
    double smr[2][SBLIMIT];
    double max_sc[2][SBLIMIT];
} ref_tap_small;

extern int tablenum; /* encode_new.c:102 */

void ref_tap_read_small(ref_tap_small *t)
{
    t->mode = header.mode;
    t->mode_ext = header.mode_ext;
    t->jsbound = frame.jsbound;
    t->sblimit = frame.sblimit;
    t->nch = frame.nch;
    t->tablenum = tablenum;
    t->bitrate_index = header.bitrate_index;
    t->dab_extension = header.dab_extension;
    memcpy(t->scalar, scalar, sizeof scalar);
    memcpy(t->j_scale, j_scale, sizeof j_scale);
    memcpy(t->scfsi, scfsi, sizeof scfsi);
    memcpy(t->bit_alloc, bit_alloc, sizeof bit_alloc);
    memcpy(t->smr, smr, sizeof smr);
    memcpy(t->max_sc, max_sc, sizeof max_sc);
}

/* [2][3][12][32] doubles */
const double *ref_tap_sb_sample(void) { return &(*sb_sample)[0][0][0][0]; }
/* [2][3][12][32] unsigned */
const unsigned int *ref_tap_subband(void) { return &(*subband)[0][0][0][0]; }
