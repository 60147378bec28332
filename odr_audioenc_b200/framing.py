"""ctypes binding of the framing / PAD entry points (include/dab_framing_b200.h): host-only, no GPU needed."""
import ctypes as C

import numpy as np

from .binding import TlbError, lib

ZMQ_HEADER_SIZE = 12


class _EdiConfig(C.Structure):
    _fields_ = [("tist", C.c_int32), ("delay_ms", C.c_uint32), ("tagpacket_alignment", C.c_uint32),
                ("tai_utc_offset", C.c_int32), ("start_time", C.c_int64), ("version_tag", C.c_char_p)]


class _PftConfig(C.Structure):
    _fields_ = [("fec", C.c_uint32), ("chunk_len", C.c_uint32)]


def _check(rc):
    if rc < 0:
        raise TlbError("tlb error %d: %s" % (rc, lib().tlb_last_error().decode()))
    return rc


def _setup(L):
    if getattr(L, "_framing_ready", False):
        return L
    vp, sz = C.c_void_p, C.c_size_t
    L.tlb_zmq_messages.argtypes = [vp, sz, sz, vp, vp]
    L.tlb_zmq_messages.restype = C.c_long
    L.tlb_edi_create.argtypes = [C.POINTER(vp), C.POINTER(_EdiConfig)]
    L.tlb_edi_destroy.argtypes = [vp]
    L.tlb_edi_destroy.restype = None
    L.tlb_edi_packet_bound.argtypes = [vp, sz]
    L.tlb_edi_packet_bound.restype = sz
    L.tlb_edi_packets.argtypes = [vp, vp, sz, sz, vp, vp, sz, vp]
    L.tlb_edi_packets.restype = C.c_long
    L.tlb_pft_create.argtypes = [C.POINTER(vp), C.POINTER(_PftConfig)]
    L.tlb_pft_destroy.argtypes = [vp]
    L.tlb_pft_destroy.restype = None
    L.tlb_pft_bound.argtypes = [vp, sz, C.POINTER(sz)]
    L.tlb_pft_bound.restype = sz
    L.tlb_pft_fragments.argtypes = [vp, vp, sz, vp, sz, vp, sz]
    L.tlb_pft_fragments.restype = C.c_long
    L.tlb_pad_open.argtypes = [C.POINTER(vp), C.c_char_p]
    L.tlb_pad_close.argtypes = [vp]
    L.tlb_pad_close.restype = None
    L.tlb_pad_request.argtypes = [vp, C.c_int, vp]
    L.tlb_pad_fill.argtypes = [vp, C.c_int, sz, vp]
    L.tlb_pad_fill.restype = C.c_long
    L._framing_ready = True
    return L


def zmq_messages(frames, frame_len, peaks=None):
    """frames: uint8 array of n * frame_len bytes (a batch output); returns uint8 (n, 12 + frame_len)."""
    L = _setup(lib())
    frames = np.ascontiguousarray(frames, dtype=np.uint8).ravel()
    n = frames.size // frame_len
    pk = np.ascontiguousarray(peaks, dtype=np.int16) if peaks is not None else None
    out = np.empty((n, ZMQ_HEADER_SIZE + frame_len), dtype=np.uint8)
    _check(L.tlb_zmq_messages(frames.ctypes.data, n, frame_len, pk.ctypes.data if pk is not None else None, out.ctypes.data))
    return out


class EdiPacketiser:
    """One EDI stream (tlb_edi_*): stateful like Output::EDI (frame counter, AF sequence, time stamp)."""

    def __init__(self, tist=False, delay_ms=0, tagpacket_alignment=0, tai_utc_offset=37, start_time=0, version_tag=""):
        self._L = _setup(lib())
        self._h = C.c_void_p()
        self._tag = version_tag.encode()
        cfg = _EdiConfig(int(tist), delay_ms, tagpacket_alignment, tai_utc_offset, start_time, self._tag)
        _check(self._L.tlb_edi_create(C.byref(self._h), C.byref(cfg)))

    def packets(self, frames, frame_len, peaks=None):
        """Returns the list of AF packets (bytes) for a batch of frames."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8).ravel()
        n = frames.size // frame_len
        pk = np.ascontiguousarray(peaks, dtype=np.int16) if peaks is not None else None
        cap = n * self._L.tlb_edi_packet_bound(self._h, frame_len)
        out = np.empty(cap, dtype=np.uint8)
        sizes = np.zeros(n, dtype=np.uint32)
        total = _check(self._L.tlb_edi_packets(self._h, frames.ctypes.data, n, frame_len, pk.ctypes.data if pk is not None else None,
                                               out.ctypes.data, cap, sizes.ctypes.data))
        assert total == int(sizes.sum())
        ends = np.cumsum(sizes)
        return [out[e - s:e].tobytes() for s, e in zip(sizes, ends)]

    def close(self):
        if self._h:
            self._L.tlb_edi_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


class PftFragmenter:
    """The PFT layer of one EDI stream (tlb_pft_*): AF packet -> PF fragments; the PF sequence number counts packets."""

    def __init__(self, fec=0, chunk_len=0):
        self._L = _setup(lib())
        self._h = C.c_void_p()
        cfg = _PftConfig(fec, chunk_len)
        _check(self._L.tlb_pft_create(C.byref(self._h), C.byref(cfg)))

    def fragments(self, af_packet):
        af = np.frombuffer(af_packet, dtype=np.uint8)
        nmax = C.c_size_t()
        cap = self._L.tlb_pft_bound(self._h, af.size, C.byref(nmax))
        out = np.empty(cap, dtype=np.uint8)
        sizes = np.zeros(nmax.value, dtype=np.uint32)
        n = _check(self._L.tlb_pft_fragments(self._h, af.ctypes.data, af.size, out.ctypes.data, cap, sizes.ctypes.data, nmax.value))
        ends = np.cumsum(sizes[:n])
        return [out[e - s:e].tobytes() for s, e in zip(sizes[:n], ends)]

    def close(self):
        if self._h:
            self._L.tlb_pft_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close


class PadSocket:
    """The ODR-PadEnc client side (tlb_pad_*): records come back in the layout BatchEncoder.encode(xpad=...) takes."""

    def __init__(self, ident):
        self._L = _setup(lib())
        self._h = C.c_void_p()
        _check(self._L.tlb_pad_open(C.byref(self._h), ident.encode()))

    def request(self, pad_len):
        rec = np.zeros(pad_len + 1, dtype=np.uint8)
        used = _check(self._L.tlb_pad_request(self._h, pad_len, rec.ctypes.data))
        return used, rec

    def fill(self, pad_len, n_frames):
        recs = np.zeros((n_frames, pad_len + 1), dtype=np.uint8)
        n = _check(self._L.tlb_pad_fill(self._h, pad_len, n_frames, recs.ctypes.data))
        return n, recs

    def close(self):
        if self._h:
            self._L.tlb_pad_close(self._h)
            self._h = C.c_void_p()

    __del__ = close
