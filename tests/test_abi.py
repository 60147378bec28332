"""CPU: the C-ABI library loads and exports every symbol the headers declare (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return re.findall(r"^TLB_API[^;(]*?\b(\w+)\(", txt, flags=re.M)


def _lib():
    import __graft_entry__ as ge
    p = os.path.join(ROOT, "odr_audioenc_b200", "libtoolame_b200.so")
    if not os.path.exists(p):
        ge.build()
    return ctypes.CDLL(p)


def test_every_declared_symbol_is_exported():
    L = _lib()
    names = _declared("toolame_b200.h") + _declared("toolame.h") + _declared("dab_framing_b200.h")
    assert len(names) == 25 + 9 + 15
    for n in names:
        assert hasattr(L, n), n


def test_reference_symbol_list():
    # libtoolame-dab.sym:1-9
    want = {"toolame_init", "toolame_finish", "toolame_enable_byteswap", "toolame_set_channel_mode",
            "toolame_set_psy_model", "toolame_set_bitrate", "toolame_set_samplerate", "toolame_set_pad",
            "toolame_encode_frame"}
    assert set(_declared("toolame.h")) == want


def test_parameter_errors_do_not_need_a_gpu():
    import odr_audioenc_b200 as tl
    for args in ((44000, "s", 192), (48000, "s", 100), (48000, "x", 192), (48000, "s", 192, 7)):
        with pytest.raises(tl.TlbError):
            tl.BatchEncoder(*args)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import odr_audioenc_b200 as tl
    with pytest.raises(tl.TlbError, match="CUDA"):
        tl.BatchEncoder(48000, "j", 192)


def test_config_check_runs_on_the_host():
    """tlb_config_check derives the per-stream constants without touching CUDA"""
    import odr_audioenc_b200 as tl
    from odr_audioenc_b200.binding import _Config, _Info
    L = tl.lib()
    want = {(48000, "j", 192, 1): (2, 576, 27, 0, 4, 480), (48000, "j", 128, 1): (2, 384, 27, 0, 4, 480),
            (24000, "m", 64, 1): (1, 384, 30, 4, 4, 480), (48000, "s", 96, 1): (2, 288, 8, 2, 2, 480),
            (48000, "j", 256, 2): (2, 768, 27, 0, 4, 1632), (32000, "s", 192, 1): (2, 864, 30, 1, 4, 480),
            (32000, "m", 48, 1): (1, 216, 12, 3, 2, 480), (16000, "m", 32, 0): (1, 288, 30, 4, 4, 480)}
    for (fs, mode, br, psy), (nch, lg, sbl, tab, ext, halo) in want.items():
        c, i = _Config(fs, ord(mode), br, psy, 0), _Info()
        assert L.tlb_config_check(ctypes.byref(c), ctypes.byref(i)) == 0
        assert (i.nch, i.lg_frame, i.sblimit, i.tablenum, i.dab_ext, i.halo_samples) == (nch, lg, sbl, tab, ext, halo)
    for fs, mode, br, psy, rc in ((44100, "s", 192, 1, -4), (22050, "m", 64, 1, -4), (48000, "s", 100, 1, -1),
                                  (48000, "s", 192, 3, -4), (48000, "s", 192, 4, -1), (12345, "s", 192, 1, -1)):
        c = _Config(fs, ord(mode), br, psy, 0)
        assert L.tlb_config_check(ctypes.byref(c), None) == rc, (fs, mode, br, psy)


def test_dropin_setters_need_no_gpu_and_fail_like_the_reference():
    """toolame_set_* validate on the host (toolame.c:168-262): 0 = accepted, non-zero = refused; nothing touches CUDA"""
    import odr_audioenc_b200 as tl
    L = tl.lib()
    assert L.toolame_init() == 0
    assert L.toolame_set_samplerate(48000) == 0 and L.toolame_set_samplerate(11025) == -1
    assert L.toolame_set_psy_model(1) == 0 and L.toolame_set_psy_model(5) == 1 and L.toolame_set_psy_model(3) == 1
    assert L.toolame_set_channel_mode(b"j") == 0 and L.toolame_set_channel_mode(b"x") == 1
    assert L.toolame_set_bitrate(192) == 0 and L.toolame_set_bitrate(100) == 1
    assert L.toolame_set_pad(58) == 0 and L.toolame_set_pad(-1) == 1 and L.toolame_set_pad(256) == 1
    assert L.toolame_b200_status() == 0
    assert L.toolame_init() == 0


def test_dropin_without_gpu_latches_an_error_instead_of_dropping_frames():
    """no device: the frames up to the first flush are buffered, the flush-due call cannot encode, the stream stops
    (status != 0, every later call returns 0 bytes, finish returns 0) until toolame_init -- never a gap in a stream"""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import odr_audioenc_b200 as tl
    s = tl.ToolameStream(48000, "j", 192)
    frame = np.zeros((1152, 2), dtype=np.int16)
    sizes = [s.encode_frame(frame).size for _ in range(12)]
    assert sizes == [0] * 12
    assert tl.lib().toolame_b200_status() != 0
    assert s.finish().size == 0
    assert tl.lib().toolame_b200_status() == 0   # finish resets the stream like toolame_init
