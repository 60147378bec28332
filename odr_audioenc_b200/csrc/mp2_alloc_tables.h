// mp2_alloc_tables.h -- Layer II bit-allocation / quantiser constants (hand-written).
//
// ISO 11172-3 tables B.2a-d, B.4, C.5, C.6 and ISO 13818-3 table B.1 in the compact
// "quantiser class + allocation row" form.  Semantics match what libtoolame-dab works
// with (encode_new.c:16-62 step_index/nbal/steps/bits/group/line, :96-100 SNR,
// :448-462 a/b): in particular MP2_QC_A/B are the 9-decimal ROUNDED literals of
// table C.6, not the exact fractions, because the quantiser's truncation sees them.
#pragma once
#ifndef MP2_TABLE_QUAL
#define MP2_TABLE_QUAL static const
#endif

#define MP2_NQC 18
// quantiser class q: number of steps
MP2_TABLE_QUAL int MP2_QC_STEPS[MP2_NQC] = {0, 3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};
// bits per transmitted codeword (grouped classes 3/5/9 steps send one 5/7/10-bit codeword per triplet)
MP2_TABLE_QUAL int MP2_QC_BITS[MP2_NQC] = {0, 5, 7, 3, 10, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
// codewords per sample triplet: 1 = grouped, 3 = one per sample (the reference calls this "group")
MP2_TABLE_QUAL int MP2_QC_NCODE[MP2_NQC] = {0, 1, 1, 3, 1, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3};
// largest power of two below the step count = weight of the (inverted) sign bit
MP2_TABLE_QUAL int MP2_QC_MSB[MP2_NQC] = {0, 2, 4, 4, 8, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
// signal-to-noise ratio of each class in dB (table C.5)
MP2_TABLE_QUAL double MP2_QC_SNR[MP2_NQC] = {0.00, 7.00, 11.00, 16.00, 20.84, 25.28, 31.59, 37.75, 43.84,
                                             49.89, 55.93, 61.96, 67.98, 74.01, 80.03, 86.05, 92.01, 98.01};
// quantisation coefficients (table C.6), as printed there to 9 decimals
MP2_TABLE_QUAL double MP2_QC_A[MP2_NQC] = {0,
    0.750000000, 0.625000000, 0.875000000, 0.562500000, 0.937500000, 0.968750000, 0.984375000, 0.992187500,
    0.996093750, 0.998046875, 0.999023438, 0.999511719, 0.999755859, 0.999877930, 0.999938965, 0.999969482,
    0.999984741};
MP2_TABLE_QUAL double MP2_QC_B[MP2_NQC] = {0,
    -0.250000000, -0.375000000, -0.125000000, -0.437500000, -0.062500000, -0.031250000, -0.015625000,
    -0.007812500, -0.003906250, -0.001953125, -0.000976563, -0.000488281, -0.000244141, -0.000122070,
    -0.000061035, -0.000030518, -0.000015259};

// The 9 distinct allocation rows: bits of the allocation index, and index -> quantiser class
#define MP2_NROWS 9
MP2_TABLE_QUAL int MP2_ROW_NBAL[MP2_NROWS] = {4, 4, 3, 2, 4, 3, 4, 3, 2};
MP2_TABLE_QUAL int MP2_ROW_QC[MP2_NROWS][16] = {
    {0, 1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17},   // B.2a/b sb 0-2
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 17},     // B.2a/b sb 3-10
    {0, 1, 2, 3, 4, 5, 6, 17, 0, 0, 0, 0, 0, 0, 0, 0},          // B.2a/b sb 11-22
    {0, 1, 2, 17, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},          // B.2a/b sb 23-
    {0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16},    // B.2c/d sb 0-1
    {0, 1, 2, 4, 5, 6, 7, 8, 0, 0, 0, 0, 0, 0, 0, 0},           // B.2c/d sb 2-
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},     // LSF sb 0-3
    {0, 1, 2, 4, 5, 6, 7, 8, 0, 0, 0, 0, 0, 0, 0, 0},           // LSF sb 4-10
    {0, 1, 2, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},           // LSF sb 11-29
};

// The 5 allocation tables (0..3 = MPEG-1 B.2a..d, 4 = LSF): subband limit and row per subband (-1 above it)
#define MP2_NTABLES 5
MP2_TABLE_QUAL int MP2_TAB_SBLIMIT[MP2_NTABLES] = {27, 30, 8, 12, 30};
MP2_TABLE_QUAL signed char MP2_TAB_ROW[MP2_NTABLES][32] = {
    {0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, -1, -1, -1, -1, -1},
    {0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, -1, -1},
    {4, 4, 5, 5, 5, 5, 5, 5, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1},
    {4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1},
    {6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, -1, -1},
};

// number of scalefactors transmitted for each scfsi code
MP2_TABLE_QUAL int MP2_SCFSI_NSF[4] = {3, 2, 1, 2};
// joint-stereo bound per mode_ext (common.c:64-74)
MP2_TABLE_QUAL int MP2_JSBOUND[4] = {4, 8, 12, 16};
// kbit/s per bitrate index, [version] (0 = LSF, 1 = MPEG-1) (common.c:29-32)
MP2_TABLE_QUAL int MP2_BITRATE[2][15] = {
    {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160},
    {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384}};
