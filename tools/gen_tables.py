#!/usr/bin/env python3
"""Generate odr_audioenc_b200/csrc/mp2_tables.h  (constant tables of the MP2 encode path).

The numeric tables of ISO 11172-3 / 13818-3 that libtoolame-dab embeds (analysis
window, scalefactors, critical-band boundaries, absolute-threshold tables ...)
are read back from the COMPILED reference (oracle/_ref/libtoolame_ref.so, built
by oracle/Makefile from /root/reference) through its exported symbols and
functions, and re-emitted in this project's own layout as shortest round-trip
double literals, so every value is bit-identical to what the reference computes
with.  Derived tables the reference builds at start-up with libm (DCT matrix,
Hann window, add_db table, FHT twiddle recurrences) are evaluated here with the
same expressions in IEEE double arithmetic and likewise frozen, so the device
code needs no start-up transcendental at all.

Run in the build container (needs /root/reference compiled):  python tools/gen_tables.py
tests/test_tables.py re-checks the committed header against the reference.
"""
import ctypes as C
import math
import os
import sys
from decimal import Decimal, getcontext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libtoolame_ref.so")
OUT = os.path.join(ROOT, "odr_audioenc_b200", "csrc", "mp2_tables.h")

PI_REF = 3.14159265358979  # common.h:26 (truncated on purpose)


class GThres(C.Structure):  # encoder.h:64-69
    _fields_ = [("line", C.c_int), ("bark", C.c_double), ("hear", C.c_double), ("x", C.c_double)]


def dbl(v):
    r = repr(float(v))
    if r in ("inf", "-inf", "nan"):
        raise ValueError(r)
    return r


def emit_array(f, ctype, name, dims, values, per_line=6, fmt=dbl):
    f.write("MP2_TABLE_QUAL %s %s%s = {\n" % (ctype, name, "".join("[%d]" % d for d in dims)))
    flat = list(values)
    for i in range(0, len(flat), per_line):
        f.write("  " + ", ".join(fmt(v) for v in flat[i:i + per_line]) + ",\n")
    f.write("};\n\n")


def half_angle_tables(n):
    """cos/sin(pi/2^(k+1)), k=0..n-1, correctly rounded (fft.c:38-73 holds the same constants as 50-digit literals)."""
    getcontext().prec = 80
    cos_t, sin_t = [], []
    c = Decimal(0)  # cos(pi/2)
    for k in range(n):
        cos_t.append(float(c))
        sin_t.append(float((1 - c * c).sqrt()))
        c = ((1 + c) / 2).sqrt()
    return cos_t, sin_t


def fht_twiddles():
    """(c1,s1,c2,s2) for every i of the 4 radix-4 stages, by the sequential recurrence of fft.c:1138-1148."""
    costab, sintab = half_angle_tables(16)
    out = []  # flattened: stage k=2,4,6,8 -> kx-1 entries each
    offsets = []
    for k in (2, 4, 6, 8):
        k1 = 1 << k
        kx = k1 >> 1
        t_c, t_s = costab[k], sintab[k]
        c1, s1 = 1.0, 0.0
        offsets.append(len(out) // 4)
        for i in range(1, kx):
            t = c1
            c1 = t * t_c - s1 * t_s
            s1 = t * t_s + s1 * t_c
            c2 = c1 * c1 - s1 * s1
            s2 = 2 * (c1 * s1)
            out += [c1, s1, c2, s2]
    return out, offsets


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    lib = C.CDLL(REF)

    enwindow = list((C.c_double * 512).in_dll(lib, "enwindow"))
    multiple = list((C.c_double * 64).in_dll(lib, "multiple"))
    scalefactor = list((C.c_double * 64).in_dll(lib, "scalefactor"))
    assert multiple == scalefactor, "common.c multiple[] and encode_new.c scalefactor[] differ"

    m = ((C.c_double * 32) * 16)()
    lib.create_dct_matrix(m)  # subband.c:125-137
    dct = [m[i][k] for i in range(16) for k in range(32)]

    lib.psycho_1_init_add_db()  # psycho_1.c:170-178
    dbtable = list((C.c_double * 1000).in_dll(lib, "dbtable"))
    # same expression evaluated here must agree (sanity of the libm in use)
    for i in range(1000):
        x = i / 10.0
        assert dbtable[i] == 10 * math.log10(1 + math.pow(10.0, x / 10.0)) - x

    # Hann window of psycho_1_hann_fft_pickmax (psycho_1.c:225-233)
    sqrt_8_over_3 = math.pow(8.0 / 3.0, 0.5)
    hann = [sqrt_8_over_3 * 0.5 * (1 - math.cos(2.0 * PI_REF * i / 1024)) / 1024 for i in range(1024)]

    # 20*log10(multiple*32768)-10  (psycho_1.c:575)
    sf_db = [20 * math.log10(v * 32768) - 10 for v in multiple]

    # psycho-1 critical bands and threshold tables for the 6 legal frequency indices
    # (index = sampling_frequency for MPEG-1, +4 for LSF: psycho_1.c:42-48)
    lib.psycho_1_read_freq_band.argtypes = [C.POINTER(C.POINTER(GThres)), C.c_int, C.c_int]
    cb_n, cb = [], []
    fr_n, fr_line, fr_bark, fr_hear = [], [], [], []
    for freq in range(7):
        if freq == 3:
            cb_n.append(0); cb.append([0] * 28)
            fr_n.append(0); fr_line.append([0] * 134); fr_bark.append([0.0] * 134); fr_hear.append([0.0] * 134)
            continue
        lib.psycho_1_read_cbound(2, freq)
        n = C.c_int.in_dll(lib, "crit_band").value
        p = C.POINTER(C.c_int).in_dll(lib, "cbound")
        cb_n.append(n)
        cb.append([p[i] for i in range(n)] + [0] * (28 - n))
        ltg = C.POINTER(GThres)()
        lib.psycho_1_read_freq_band(C.byref(ltg), 2, freq)
        ss = C.c_int.in_dll(lib, "sub_size").value
        fr_n.append(ss)
        fr_line.append([ltg[i].line for i in range(ss)] + [0] * (134 - ss))
        fr_bark.append([ltg[i].bark for i in range(ss)] + [0.0] * (134 - ss))
        fr_hear.append([ltg[i].hear for i in range(ss)] + [0.0] * (134 - ss))

    tw, tw_off = fht_twiddles()

    with open(OUT, "w") as f:
        f.write("// mp2_tables.h -- GENERATED by tools/gen_tables.py; do not edit.\n"
                "// Constant tables of the MPEG Layer II (DAB) encode path, bit-identical to the values\n"
                "// libtoolame-dab works with (see the generator for the provenance of each table).\n"
                "#pragma once\n#ifndef MP2_TABLE_QUAL\n#define MP2_TABLE_QUAL static const\n#endif\n\n")
        f.write("// ISO 11172-3 analysis window C[i] (reference: enwindow.h:1-130)\n")
        emit_array(f, "double", "MP2_ENWINDOW", [512], enwindow)
        f.write("// scalefactor table, index 0..62 = 2/cuberoot(2)^n, 63 = 1e-20 (encode_new.c:65-83 == common.c:34-52)\n")
        emit_array(f, "double", "MP2_SCALEFACTOR", [64], multiple, per_line=4)
        f.write("// 20*log10(scalefactor*32768)-10 (psycho_1.c:575), evaluated with the host libm\n")
        emit_array(f, "double", "MP2_SF_DB", [64], sf_db, per_line=4)
        f.write("// 16x32 analysis matrix, round(1e9*cos((2i+1)k*PI/64))*1e-9 (subband.c:125-137)\n")
        emit_array(f, "double", "MP2_DCT", [16, 32], dct, per_line=4)
        f.write("// psy-1 Hann window incl. sqrt(8/3)/1024 normalisation (psycho_1.c:225-233)\n")
        emit_array(f, "double", "MP2_HANN", [1024], hann, per_line=4)
        f.write("// add_db correction table 10*log10(1+10^(x/10))-x, x=i/10 (psycho_1.c:170-178); entry 1000 = +0.0 stands in\n"
                "// for add_db's two early returns (psycho_1.c:189-192) in the branch-free device version\n")
        emit_array(f, "double", "MP2_DBTABLE", [1001], dbtable + [0.0], per_line=4)
        f.write("// FHT-1024 twiddles (c1,s1,c2,s2) per i for the stages k1=4,16,64,256 (fft.c:1138-1148)\n")
        f.write("#define MP2_FHT_TW_COUNT %d\n" % (len(tw) // 4))
        emit_array(f, "int", "MP2_FHT_TW_OFFSET", [4], tw_off, fmt=str)
        emit_array(f, "double", "MP2_FHT_TW", [len(tw) // 4, 4], tw, per_line=4)
        f.write("// psy-1 critical band boundaries per frequency index (critband.h via psycho_1.c:94-123)\n")
        emit_array(f, "int", "MP2_CB_COUNT", [7], cb_n, per_line=7, fmt=str)
        emit_array(f, "int", "MP2_CBOUND", [7, 28], [v for r in cb for v in r], per_line=14, fmt=str)
        f.write("// psy-1 threshold calculation partitions (freqtable.h via psycho_1.c:125-157); entry 0 is the\n"
                "// reference's synthetic {line 0, bark 0, hear 0}; MP2_SUB_SIZE counts it.\n")
        emit_array(f, "int", "MP2_SUB_SIZE", [7], fr_n, per_line=7, fmt=str)
        emit_array(f, "int", "MP2_LTG_LINE", [7, 134], [v for r in fr_line for v in r], per_line=16, fmt=str)
        emit_array(f, "double", "MP2_LTG_BARK", [7, 134], [v for r in fr_bark for v in r], per_line=8)
        emit_array(f, "double", "MP2_LTG_HEAR", [7, 134], [v for r in fr_hear for v in r], per_line=8)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")

    # ---- psychoacoustic model 2: Hann window (psycho_2.c:318-319) and absolute threshold tables (absthr.h through
    # psycho_2_read_absthr, psycho_2.c:422-438)
    out2 = os.path.join(os.path.dirname(OUT), "mp2_psy2_tables.h")
    window2 = [0.5 * (1 - math.cos(2.0 * PI_REF * (i - 0.5) / 1024)) for i in range(1024)]
    lib.psycho_2_read_absthr.argtypes = [C.POINTER(C.c_double), C.c_int]
    absthr = []
    for table in range(3):
        buf = (C.c_double * 513)()
        lib.psycho_2_read_absthr(buf, table)
        absthr += list(buf)
    with open(out2, "w") as f:
        f.write("// mp2_psy2_tables.h -- GENERATED by tools/gen_tables.py; do not edit.\n"
                "// Constant tables of psychoacoustic model 2, bit-identical to the reference's values.\n"
                "#pragma once\n#ifndef MP2_TABLE_QUAL\n#define MP2_TABLE_QUAL static const\n#endif\n\n")
        f.write("// psy-2 Hann window 0.5*(1-cos(2*PI*(i-0.5)/1024)) with the reference's truncated PI (psycho_2.c:318-319)\n")
        emit_array(f, "double", "MP2_P2_WINDOW", [1024], window2, per_line=4)
        f.write("// absolute threshold per FFT line for 32/16, 44.1/22.05 and 48/24 kHz (absthr.h)\n")
        emit_array(f, "double", "MP2_ABSTHR", [3, 513], absthr, per_line=6)
        # ---- start-up tables of psycho_2_init (psycho_2.c:259-420), read back from the reference itself after it
        # has run for each sample rate the encoder accepts (oracle/psy2_tap.c), and psy model 0's lowest absolute
        # threshold per subband (psycho_0.c:36-47) from the reference's ATH_dB (ath.c:7-49)
        tap = C.CDLL(os.path.join(os.path.dirname(REF), "libpsy2_tap.so"))
        tap.psy2_tap_init.argtypes = [C.c_double]
        for fn, ct in (("partition", C.c_int), ("numlines", C.c_int), ("cbval", C.c_double), ("rnorm", C.c_double),
                       ("tmn", C.c_double), ("s", C.c_double), ("bmax", C.c_double), ("absthr", C.c_double)):
            getattr(tap, "psy2_tap_" + fn).restype = C.POINTER(ct)
        lib.ATH_dB.restype = C.c_double
        lib.ATH_dB.argtypes = [C.c_double, C.c_double]
        rates = [48000, 24000, 32000, 16000]
        part, numl, first, tmn, rnorm, bmax_of, s_t, abs_idx, ath_min = [], [], [], [], [], [], [], [], []
        for fs in rates:
            tap.psy2_tap_init(float(fs))
            p_ = [tap.psy2_tap_partition()[i] for i in range(513)]
            n_ = [tap.psy2_tap_numlines()[i] for i in range(64)]
            cb_ = [tap.psy2_tap_cbval()[i] for i in range(64)]
            s_ = [tap.psy2_tap_s()[i] for i in range(64 * 64)]
            bm = [tap.psy2_tap_bmax()[i] for i in range(27)]
            n_part = p_[512] + 1
            assert sum(n_[:n_part]) == 513 and all(b >= a for a, b in zip(p_, p_[1:]))
            fl = [p_.index(j) for j in range(n_part)] + [513] * (65 - n_part)   # partitions are runs of consecutive lines
            picked = [tap.psy2_tap_absthr()[i] for i in range(513)]
            abs_idx.append([absthr[513 * t:513 * (t + 1)] for t in range(3)].index(picked))
            part += p_ + [0] * 7
            numl += n_
            first += fl
            tmn += [tap.psy2_tap_tmn()[i] for i in range(64)]
            rnorm += [tap.psy2_tap_rnorm()[i] for i in range(64)]
            bmax_of += [bm[int(c + 0.5)] for c in cb_]            # psycho_2.c:195-196: bmax[(unsigned int)(cbval[j] + 0.5)]
            s_t += [s_[j * 64 + k] for k in range(64) for j in range(64)]   # transposed: sT[k][j] = s[j][k]
            am = [1000.0] * 32                                      # psycho_0.c:36-47
            for i in range(512):
                v = lib.ATH_dB(i * (float(fs) / 1024.0), 0.0)
                if v < am[i >> 4]:
                    am[i >> 4] = v
            ath_min += am
        f.write("// ---- per sample rate: the start-up tables psycho_2_init builds with libm (psycho_2.c:259-420), read back from\n"
                "// the compiled reference (oracle/psy2_tap.c); sT is the spreading function transposed, sT[k][j] = s[j][k]\n")
        f.write("// (host side only: the kernels get them through Mp2Psy2Tables / Mp2PsyTables)\n#ifndef MP2_DEVICE_TABLES_ONLY\n")
        f.write("#define MP2_P2_RATES %d\n" % len(rates))
        emit_array(f, "int", "MP2_P2_RATE", [len(rates)], rates, fmt=str)
        emit_array(f, "int", "MP2_P2_ABSTHR_TABLE", [len(rates)], abs_idx, fmt=str)
        emit_array(f, "unsigned char", "MP2_P2_PARTITION", [len(rates), 520], part, per_line=26, fmt=str)
        emit_array(f, "int", "MP2_P2_NUMLINES", [len(rates), 64], numl, per_line=16, fmt=str)
        emit_array(f, "int", "MP2_P2_FIRST_LINE", [len(rates), 65], first, per_line=13, fmt=str)
        emit_array(f, "double", "MP2_P2_TMN", [len(rates), 64], tmn, per_line=4)
        emit_array(f, "double", "MP2_P2_RNORM", [len(rates), 64], rnorm, per_line=4)
        emit_array(f, "double", "MP2_P2_BMAX_OF", [len(rates), 64], bmax_of, per_line=8)
        emit_array(f, "double", "MP2_P2_ST", [len(rates), 64, 64], s_t, per_line=4)
        f.write("// psy model 0: lowest absolute threshold of hearing per subband in dB (psycho_0.c:36-47 over ath.c:7-49)\n")
        emit_array(f, "double", "MP2_P0_ATH_MIN", [len(rates), 32], ath_min, per_line=4)
        f.write("#endif\n")
    print("wrote", out2, os.path.getsize(out2), "bytes")


if __name__ == "__main__":
    main()
