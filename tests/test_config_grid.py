"""Every configuration the library accepts: four sample rates x four modes x the fourteen bitrates of the rate's
MPEG version x psychoacoustic models 0, 1 and 2 = 672 configurations (ref: toolame.c:175-260, common.c:96-117: the
reference checks the bitrate against the version's table and nothing else, so per-channel rates outside ISO's
allocation-table ranges, e.g. 32 kbit/s stereo at 48 kHz or 384 kbit/s mono, are encodable and land on whatever
table encode.c:60-110 picks).

 * CPU, build container: the oracle port against the compiled reference on each of them (skipped without oracle/_ref);
 * GPU: the CUDA path against the oracle on each of them, through tlb_config_check / tlb_batch_create / _encode.
"""
import numpy as np
import pytest

import oracle
import reftool
import signals

MPEG1 = [32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384]
MPEG2 = [8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]
RATES = {48000: MPEG1, 32000: MPEG1, 24000: MPEG2, 16000: MPEG2}
GRID = [(fs, mode, br) for fs in RATES for mode in "sjdm" for br in RATES[fs]]
N_CPU, N_GPU = 40, 8


def _pcm(fs, mode, n):
    nch = 1 if mode == "m" else 2
    # S8 = the noise-like signal (every subband busy, tonal and noise lists both long); the other half speech-like
    a = signals.make("S8", n // 2, nch, fs)
    b = signals.make("S1", n - n // 2, nch, fs)
    return np.concatenate([a, b])


def _ref_bytes(job):
    fs, mode, br, psy = job
    return reftool.run_ref(_pcm(fs, mode, N_CPU), fs, mode, br, psy)["bytes"]


@pytest.mark.skipif(not reftool.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("psy", [1, 2, 0])
def test_oracle_equals_reference_on_every_configuration(psy):
    from concurrent.futures import ThreadPoolExecutor
    jobs = [(fs, mode, br, psy) for fs, mode, br in GRID]
    with ThreadPoolExecutor(8) as pool:   # the reference is one subprocess per stream; the port runs in this thread
        refs = list(pool.map(_ref_bytes, jobs))
    bad = []
    for (fs, mode, br, _), ref in zip(jobs, refs):
        out, _ = oracle.encode(oracle.configure(fs, mode, br, psy), _pcm(fs, mode, N_CPU))
        if out.size != ref.size:
            bad.append(((fs, mode, br), "size %d, reference %d" % (out.size, ref.size)))
        elif not np.array_equal(out, ref):
            bad.append(((fs, mode, br), "frames %s differ" % np.flatnonzero((out.reshape(N_CPU, -1) != ref.reshape(N_CPU, -1)).any(axis=1))[:8]))
    assert not bad, bad[:10]
    assert len(refs) == 224


def test_host_configuration_equals_the_oracle_on_every_configuration():
    """tlb_config_check (what tlb_batch_create and the drop-in's toolame_set_bitrate run; no GPU needed) derives the same
    stream constants as the oracle's restatement of toolame.c:175-260 / encode.c:60-110 for all 672"""
    import odr_audioenc_b200 as tl
    for fs, mode, br in GRID:
        for psy in (0, 1, 2):
            c = oracle.configure(fs, mode, br, psy)
            rc, info = tl.config_check(fs, mode, br, psy, 0)
            assert rc == 0, (fs, mode, br, psy)
            want = dict(nch=c.nch, lg_frame=c.lg_frame, sblimit=c.sblimit, tablenum=c.tablenum, dab_ext=c.dab_ext,
                        version=c.version, bitrate_index=c.bitrate_index, sfreq_idx=c.sfreq_idx, samples_per_frame=1152,
                        halo_samples=1632 if psy == 2 else 480)
            assert info == want, (fs, mode, br, psy, info, want)
    # and what the library refuses is refused here too: bitrates of the other MPEG version, 44.1 / 22.05 kHz, psy 3
    for fs, mode, br, psy in [(48000, "j", 144, 1), (24000, "j", 192, 1), (44100, "j", 128, 1), (22050, "m", 64, 1),
                              (48000, "j", 128, 3), (48000, "x", 128, 1), (48000, "j", 7, 1)]:
        assert tl.config_check(fs, mode, br, psy, 0)[0] < 0, (fs, mode, br, psy)
    # bitrate 0 = the version's default, table entry 10 (toolame.c:217-218): 192 kbit/s MPEG-1, 96 kbit/s LSF
    assert tl.config_check(48000, "j", 0)[1]["lg_frame"] == 576 and tl.config_check(24000, "j", 0)[1]["lg_frame"] == 576
    assert tl.config_check(24000, "j", 0)[1]["bitrate_index"] == 10


@pytest.mark.gpu
@pytest.mark.parametrize("psy", [1, 2, 0])
@pytest.mark.parametrize("fs", sorted(RATES))
def test_gpu_equals_oracle_on_every_configuration(fs, psy):
    import odr_audioenc_b200 as tl
    done = 0
    for mode in "sjdm":
        for br in RATES[fs]:
            pcm = _pcm(fs, mode, N_GPU)
            c = oracle.configure(fs, mode, br, psy)
            want, _ = oracle.encode(c, pcm)
            assert tl.config_check(fs, mode, br, psy, 0)[0] == 0
            e = tl.BatchEncoder(fs, mode, br, psy, 0, 0, 0)
            got = e.encode(pcm)
            e.close()
            bad = np.flatnonzero(got != want)
            assert bad.size == 0, "%d Hz mode %s %d kbit/s psy %d: %d bytes differ, first in frame %d" % (
                fs, mode, br, psy, bad.size, bad[0] // c.lg_frame)
            done += 1
    assert done == 56
