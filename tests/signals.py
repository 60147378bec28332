"""Deterministic synthetic PCM (s16) used by the parity tests and the bench.

Signals S1..S7 follow SURVEY.md section 8(d).  All generators return an int16
array of shape (n_samples, nch) (interleaved, WAV order).
"""
import numpy as np


def lcg_u32(n, seed):
    """n successive states of s = s*1664525 + 1013904223 (mod 2^32), state advanced BEFORE use."""
    a = np.uint32(1664525)
    c = np.uint32(1013904223)
    with np.errstate(over="ignore"):
        A = np.cumprod(np.full(n, a, dtype=np.uint32), dtype=np.uint32)          # a^k, k=1..n
        geo = np.concatenate(([np.uint32(1)], A[:-1])).astype(np.uint32)         # a^0..a^(n-1)
        C = (np.cumsum(geo, dtype=np.uint32) * c).astype(np.uint32)              # c*(1+a+..+a^(k-1))
        return (A * np.uint32(seed) + C).astype(np.uint32)


def s1(n, nch, fs, seed=12345):
    """Two tones under a slow envelope plus LCG noise (the survey's probe signal)."""
    st = lcg_u32(n * nch, seed).reshape(n, nch)
    noise = ((st >> np.uint32(16)).astype(np.int64) - 32768) / 32768.0
    t = (np.arange(n, dtype=np.float64) / fs)[:, None]
    c = np.arange(nch, dtype=np.float64)[None, :]
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 0.37 * t)
    v = env * (0.3 * np.sin(2 * np.pi * (440 + 110 * c) * t)
               + 0.2 * np.sin(2 * np.pi * (3000 + 500 * np.sin(t)) * t)) + 0.05 * noise
    return np.rint(v * 32767 * 0.8).astype(np.int16)


def s2(n, nch, fs, seed=1):
    """Full-scale uniform white noise."""
    st = lcg_u32(n * nch, seed).reshape(n, nch)
    return ((st >> np.uint32(16)).astype(np.int64) - 32768).astype(np.int16)


def s3(n, nch, fs):
    """Digital silence."""
    return np.zeros((n, nch), dtype=np.int16)


def s4(n, nch, fs):
    """1 s silence / 1 s S1 alternating."""
    x = s1(n, nch, fs)
    sec = (np.arange(n) // int(fs)) % 2
    x[sec == 0] = 0
    return x


def s5(n, nch, fs):
    """+-32767 1 kHz square wave (clipping level)."""
    t = np.arange(n, dtype=np.float64) / fs
    sq = np.where(np.sin(2 * np.pi * 1000.0 * t) >= 0, 32767, -32767).astype(np.int16)
    return np.repeat(sq[:, None], nch, axis=1)


def s6(n, nch, fs):
    """Log sweep 20 Hz -> 0.45 fs at -6 dBFS over the whole length."""
    t = np.arange(n, dtype=np.float64) / fs
    T = n / fs
    f0, f1 = 20.0, 0.45 * fs
    k = np.log(f1 / f0)
    phase = 2 * np.pi * f0 * T / k * (np.exp(t / T * k) - 1.0)
    v = 0.5 * np.sin(phase)
    x = np.rint(v * 32767).astype(np.int16)
    out = np.repeat(x[:, None], nch, axis=1)
    if nch == 2:
        out[:, 1] = -out[:, 1] // 2
    return out


def s7(n, nch, fs):
    """Single-sample impulses every 4096 samples."""
    x = np.zeros((n, nch), dtype=np.int16)
    x[::4096, 0] = 30000
    if nch == 2:
        x[2048::4096, 1] = -30000
    return x


def s8(n, nch, fs, seed=777):
    """Decorrelated L/R: different tones per channel + noise bursts (stresses joint-stereo decisions)."""
    st = lcg_u32(n * nch, seed).reshape(n, nch)
    noise = ((st >> np.uint32(16)).astype(np.int64) - 32768) / 32768.0
    t = (np.arange(n, dtype=np.float64) / fs)[:, None]
    c = np.arange(nch, dtype=np.float64)[None, :]
    burst = ((np.arange(n) // 3000) % 3 == 0)[:, None]
    v = (0.25 * np.sin(2 * np.pi * (1000 + 2500 * c) * t) + 0.15 * np.sin(2 * np.pi * (9000 - 4000 * c) * t)
         + 0.1 * np.sin(2 * np.pi * 15000 * t * (1 + 0.1 * c)) + np.where(burst, 0.4, 0.01) * noise)
    return np.rint(np.clip(v, -1, 1) * 32767 * 0.9).astype(np.int16)


SIGNALS = {"S1": s1, "S2": s2, "S3": s3, "S4": s4, "S5": s5, "S6": s6, "S7": s7, "S8": s8}


def make(name, n_frames, nch, fs):
    return SIGNALS[name](n_frames * 1152, nch, fs)
