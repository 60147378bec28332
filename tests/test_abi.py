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
    names = _declared("toolame_b200.h") + _declared("toolame.h")
    assert len(names) == 22 + 9
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
