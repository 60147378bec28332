"""ctypes binding of libtoolame_b200.so (C ABI: include/toolame_b200.h and include/toolame.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TAP_SB_SAMPLE, TAP_SCALAR_PRE, TAP_J_SCALE, TAP_SMR, TAP_SIDE = range(5)

SIDE_DTYPE = np.dtype([
    ("bit_alloc", "u1", (2, 32)), ("scfsi", "u1", (2, 32)), ("scalar", "u1", (2, 3, 32)),
    ("scfcrc_own", "u1", (4,)), ("mode", "u1"), ("mode_ext", "u1"), ("jsbound", "u1"), ("xpad_len", "u1"),
    ("adb_left", "<i4"), ("crc16", "<u4")], align=True)


class TlbError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("sample_rate", "channel_mode", "bitrate", "psy_model", "pad_len")]


class _Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nch", "lg_frame", "sblimit", "tablenum", "dab_ext", "version",
                                         "bitrate_index", "sfreq_idx", "samples_per_frame", "halo_samples")]


def lib_path():
    return os.path.join(_HERE, "libtoolame_b200.so")


_lib = None


def lib():
    """Load the CUDA library; fails loudly when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise TlbError("%s is missing: run `make` (or __graft_entry__.build()) first" % p)
        L = C.CDLL(p)
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.tlb_batch_create.argtypes = [C.POINTER(vp), C.POINTER(_Config), i32, sz]
        L.tlb_batch_destroy.argtypes = [vp]
        L.tlb_batch_destroy.restype = None
        L.tlb_batch_info.argtypes = [vp, C.POINTER(_Info)]
        L.tlb_last_error.restype = C.c_char_p
        L.tlb_batch_encode.argtypes = [vp, vp, sz, sz, i32, vp, vp]
        L.tlb_batch_encode_device.argtypes = [vp, vp, sz, sz, i32, vp, vp]
        L.tlb_batch_sync.argtypes = [vp]
        L.tlb_batch_stream.argtypes = [vp]
        L.tlb_batch_stream.restype = vp
        L.tlb_batch_launches.argtypes = [vp]
        L.tlb_batch_launches.restype = C.c_uint64
        L.tlb_host_alloc.argtypes = [sz]
        L.tlb_host_alloc.restype = vp
        L.tlb_host_free.argtypes = [vp]
        L.tlb_host_free.restype = None
        L.tlb_batch_tap.argtypes = [vp, i32, vp, sz]
        L.tlb_batch_tap.restype = C.c_long
        L.toolame_set_channel_mode.argtypes = [C.c_char]
        L.toolame_set_samplerate.argtypes = [C.c_long]
        L.toolame_encode_frame.argtypes = [vp, vp, sz, vp, sz]
        L.toolame_finish.argtypes = [vp, sz]
        L.tlb_config_check.argtypes = [C.POINTER(_Config), C.POINTER(_Info)]
        L.tlb_selftest_log10.argtypes = [i32, C.c_ulonglong, C.POINTER(C.c_double)]
        L.tlb_selftest_log10.restype = C.c_longlong
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise TlbError("tlb error %d: %s" % (rc, lib().tlb_last_error().decode()))
    return rc


def config_check(sample_rate, mode, bitrate, psy=1, pad_len=0):
    """tlb_config_check: the host-side validation tlb_batch_create and toolame_set_bitrate run (no GPU is touched).
    Returns (rc, info dict or None)."""
    cfg, info = _Config(sample_rate, ord(mode), bitrate, psy, pad_len), _Info()
    rc = lib().tlb_config_check(C.byref(cfg), C.byref(info))
    return rc, ({n: getattr(info, n) for n, _ in _Info._fields_} if rc == 0 else None)


def selftest_log10(n=1 << 28, device=0):
    """tlb_selftest_log10: values (of n) on which the spectrum kernel's log10 differs from CUDA's log10, and one of them"""
    first = C.c_double(0.0)
    bad = lib().tlb_selftest_log10(device, n, C.byref(first))
    if bad < 0:
        raise TlbError("tlb error %d: %s" % (bad, lib().tlb_last_error().decode()))
    return bad, first.value


class _Service(C.Structure):
    _fields_ = [("cfg", _Config), ("pcm", C.c_void_p), ("n_frames", C.c_size_t), ("xpad", C.c_void_p), ("out", C.c_void_p),
                ("history_samples", C.c_size_t), ("has_next", C.c_int32)]


def encode_services(services, device=0, chunk_frames=0):
    """services: list of dicts(sample_rate, mode, bitrate, psy=1, pad_len=0, pcm=int16 array (samples, nch), xpad=None,
    history=0, has_next=False): whole streams, or time pieces whose pcm starts `history` samples before the piece's
    first frame and holds one more frame after it when has_next.  Returns the list of encoded streams / pieces
    (uint8 arrays).  tlb_encode_services: one call for a whole ensemble."""
    n = len(services)
    arr = (_Service * n)()
    keep, outs = [], []
    for i, sv in enumerate(services):
        nch = 1 if sv["mode"] == "m" else 2
        pcm = np.ascontiguousarray(sv["pcm"], dtype=np.int16).reshape(-1, nch)
        hist, nxt = int(sv.get("history", 0)), bool(sv.get("has_next", False))
        nf = (pcm.shape[0] - hist) // 1152 - (1 if nxt else 0)
        cfg = _Config(sv["sample_rate"], ord(sv["mode"]), sv["bitrate"], sv.get("psy", 1), sv.get("pad_len", 0))
        probe = BatchEncoder(sv["sample_rate"], sv["mode"], sv["bitrate"], sv.get("psy", 1), sv.get("pad_len", 0), device, 1)
        out = np.empty(nf * probe.lg_frame, dtype=np.uint8)
        probe.close()
        xp = sv.get("xpad")
        xp = np.ascontiguousarray(xp, dtype=np.uint8) if xp is not None else None
        keep += [pcm, xp]
        outs.append(out)
        arr[i] = _Service(cfg, pcm.ctypes.data + hist * nch * 2, nf, xp.ctypes.data if xp is not None else None, out.ctypes.data,
                          hist, int(nxt))
    L = lib()
    L.tlb_encode_services.argtypes = [C.POINTER(_Service), C.c_size_t, C.c_int, C.c_size_t]
    _check(L.tlb_encode_services(arr, n, device, chunk_frames))
    return outs


class BatchEncoder:
    """One stream configuration on one GPU (tlb_batch_*)."""

    def __init__(self, sample_rate, mode, bitrate, psy=1, pad_len=0, device=0, chunk_frames=0):
        self._h = C.c_void_p()
        cfg = _Config(sample_rate, ord(mode), bitrate, psy, pad_len)
        _check(lib().tlb_batch_create(C.byref(self._h), C.byref(cfg), device, chunk_frames))
        info = _Info()
        _check(lib().tlb_batch_info(self._h, C.byref(info)))
        for n, _ in _Info._fields_:
            setattr(self, n, getattr(info, n))
        self.pad_len = pad_len

    def close(self):
        if self._h:
            lib().tlb_batch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def encode(self, pcm, n_frames=None, history=0, has_next=False, xpad=None, out=None):
        """pcm: int16 array (samples, nch) whose row `history` is the first sample to encode."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, self.nch)
        if n_frames is None:
            n_frames = (pcm.shape[0] - history) // 1152 - (1 if has_next else 0)
        assert pcm.shape[0] >= history + (n_frames + (1 if has_next else 0)) * 1152
        if out is None:
            out = np.empty(n_frames * self.lg_frame, dtype=np.uint8)
        xp = None
        if xpad is not None:
            xp = np.ascontiguousarray(xpad, dtype=np.uint8)
            assert xp.size >= (n_frames + (1 if has_next else 0)) * (self.pad_len + 1)
        _check(lib().tlb_batch_encode(self._h, pcm.ctypes.data + history * self.nch * 2, n_frames, history,
                                      int(has_next), xp.ctypes.data if xp is not None else None, out.ctypes.data))
        return out

    def encode_async(self, pcm, n_frames, history, has_next, out, xpad=None):
        """tlb_batch_encode_async on caller-owned (ideally pinned) arrays: returns once the work is queued; sync() waits.
        pcm row `history` = first sample to encode."""
        L = lib()
        L.tlb_batch_encode_async.argtypes = L.tlb_batch_encode.argtypes
        _check(L.tlb_batch_encode_async(self._h, pcm.ctypes.data + history * self.nch * 2, n_frames, history, int(has_next),
                                        xpad.ctypes.data if xpad is not None else None, out.ctypes.data))

    def encode_device(self, d_pcm, n_frames, history, has_next, d_xpad, d_out):
        """Raw device pointers (ints); asynchronous on self.stream."""
        _check(lib().tlb_batch_encode_device(self._h, d_pcm, n_frames, history, int(has_next), d_xpad, d_out))

    def sync(self):
        _check(lib().tlb_batch_sync(self._h))

    @property
    def stream(self):
        return lib().tlb_batch_stream(self._h)

    @property
    def launches(self):
        return int(lib().tlb_batch_launches(self._h))

    def tap(self, what, n_frames):
        shapes = {TAP_SB_SAMPLE: (np.float64, (n_frames, self.nch, 36, 32)), TAP_SCALAR_PRE: (np.uint8, (n_frames, 2, 3, 32)),
                  TAP_J_SCALE: (np.uint8, (n_frames, 3, 32)), TAP_SMR: (np.float64, (n_frames, 2, 32)),
                  TAP_SIDE: (SIDE_DTYPE, (n_frames,))}
        dt, shp = shapes[what]
        a = np.zeros(shp, dtype=dt)
        got = _check(lib().tlb_batch_tap(self._h, what, a.ctypes.data, a.nbytes))
        if got != a.nbytes:
            raise TlbError("tap %d: %d of %d bytes available" % (what, got, a.nbytes))
        return a


class ToolameStream:
    """The libtoolame-dab call sequence of odr-audioenc (src/odr-audioenc.cpp:687-721,1139-1161) on the drop-in
    symbols.  Process-global like the reference: one instance at a time."""

    def __init__(self, sample_rate, mode, bitrate, psy=1, pad_len=0):
        L = lib()
        self.nch = 1 if mode == "m" else 2
        self.pad_len = pad_len
        if L.toolame_init():
            raise TlbError("toolame_init")
        for name, rc in (("samplerate", L.toolame_set_samplerate(sample_rate)), ("psy", L.toolame_set_psy_model(psy)),
                         ("mode", L.toolame_set_channel_mode(mode.encode())), ("bitrate", L.toolame_set_bitrate(bitrate)),
                         ("pad", L.toolame_set_pad(pad_len))):
            if rc:
                raise TlbError("toolame_set_%s -> %d" % (name, rc))
        self._out = np.zeros(4092, dtype=np.uint8)  # src/odr-audioenc.cpp:796-800

    def encode_frame(self, frame_pcm, xpad_rec=None):
        """frame_pcm: int16 (1152, nch).  Returns the bytes this call handed back (often empty)."""
        planar = np.zeros((2, 1152), dtype=np.int16)
        planar[:self.nch] = np.asarray(frame_pcm, dtype=np.int16).reshape(1152, self.nch).T
        if xpad_rec is not None:
            rec = np.ascontiguousarray(xpad_rec, dtype=np.uint8)
            n = lib().toolame_encode_frame(planar.ctypes.data, rec.ctypes.data, int(rec[self.pad_len]),
                                           self._out.ctypes.data, self._out.size)
        else:
            n = lib().toolame_encode_frame(planar.ctypes.data, None, 0, self._out.ctypes.data, self._out.size)
        return self._out[:n].copy()

    def finish(self):
        n = lib().toolame_finish(self._out.ctypes.data, self._out.size)
        return self._out[:n].copy()
