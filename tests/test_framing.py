"""CPU: the framing and PAD entry points of include/dab_framing_b200.h -- host-only code.

* ZeroMQ message header against the reference's packed struct (src/Outputs.h:76-92, src/Outputs.cpp:110-127);
* EDI AF packets against fixtures produced by the reference's own packetiser (tools/make_edi_golden.py ->
  oracle/_ref/edi_ref_driver = contrib/edioutput/{TagItems,TagPacket,AFPacket}.cpp compiled unmodified), and live
  against that driver where it is built; plus an independent structural parse (lengths, CRC, counters);
* the ODR-PadEnc socket protocol (src/PadInterface.cpp:37-150) against a stand-in PadEnc on UNIX datagram sockets."""
import os
import socket
import struct
import sys
import threading

import numpy as np
import pytest

import edi_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_zmq_messages_match_the_reference_header_layout():
    from odr_audioenc_b200 import framing
    rng = np.random.RandomState(3)
    for lg in (72, 384, 576, 1728):
        n = 9
        frames = rng.randint(0, 256, size=(n, lg)).astype(np.uint8)
        peaks = rng.randint(-32768, 32768, size=(n, 2)).astype(np.int16)
        msgs = framing.zmq_messages(frames, lg, peaks)
        assert msgs.shape == (n, framing.ZMQ_HEADER_SIZE + lg)
        for f in range(n):
            # struct zmq_frame_header_t: u16 version = 1, u16 encoder = ZMQ_ENCODER_MPEG_L2 (2), u32 datasize,
            # i16 audiolevel_left, i16 audiolevel_right; packed, host (little-endian) byte order; data follows
            want = struct.pack("<HHIhh", 1, 2, lg, int(peaks[f, 0]), int(peaks[f, 1])) + frames[f].tobytes()
            assert msgs[f].tobytes() == want
    assert framing.zmq_messages(frames, lg)[0, 8:12].tobytes() == b"\0\0\0\0"   # no peaks given


def _crc16_ccitt(data):
    crc = 0xFFFF
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc ^ 0xFFFF


def _parse_af(pkt):
    """ETSI TS 102 821 6.1 / 5.1: returns (seq, [(tag name, value bytes)])"""
    assert pkt[:2] == b"AF"
    length, seq, ar, pt = struct.unpack(">IHBc", pkt[2:10])
    assert ar == 0x90 and pt == b"T" and len(pkt) == 10 + length + 2
    assert struct.unpack(">H", pkt[-2:])[0] == _crc16_ccitt(pkt[:-2])
    items, at, body = [], 0, pkt[10:10 + length]
    while at + 8 <= len(body):
        name, bits = body[at:at + 4], struct.unpack(">I", body[at + 4:at + 8])[0]
        assert bits % 8 == 0
        items.append((name, body[at + 8:at + 8 + bits // 8]))
        at += 8 + bits // 8
    assert all(b == 0 for b in body[at:]) and len(body) - at < 8   # alignment padding only
    return seq, items


@pytest.mark.parametrize("name", sorted(edi_cases.CASES))
def test_edi_packets_equal_the_reference_packetiser(name):
    from odr_audioenc_b200 import framing
    case = edi_cases.CASES[name]
    frames, peaks = edi_cases.inputs(case)
    g = np.load(os.path.join(GOLD, "edi_%s.npz" % name))
    ends = np.cumsum(g["sizes"].astype(np.int64))
    want = [g["data"][e - s:e].tobytes() for s, e in zip(g["sizes"], ends)]
    e = framing.EdiPacketiser(case["tist"], case["delay_ms"], case["alignment"], case["tai"], case["start"], case["tag"])
    got = e.packets(frames[:40], case["frame_len"], peaks[:40]) + e.packets(frames[40:], case["frame_len"], peaks[40:])
    assert len(got) == len(want) == case["n"]
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert not bad, "packets differing from the reference packetiser: %s" % bad[:8]
    # independent structure check: sequence numbers, dsti frame counter modulo 5000, payload, levels
    for i in (0, 1, case["n"] // 2, case["n"] - 1):
        seq, items = _parse_af(got[i])
        names = [n for n, _ in items]
        assert seq == i % 65536 and names[:4] == [b"*ptr", b"dsti", b"ss\x00\x01", b"ODRa"]
        assert items[0][1] == b"DSTI\0\0\0\0"
        hdr = struct.unpack(">H", items[1][1][:2])[0]
        assert (hdr & 0xFF) + 250 * ((hdr >> 8) & 0x1F) == i % 5000 and bool(hdr & 0x4000) == case["tist"]
        assert items[2][1] == b"\0\0\0" + frames[i].tobytes()
        assert items[3][1] == struct.pack(">hh", int(peaks[i, 0]), int(peaks[i, 1]))
    assert sum(1 for p in got if b"ODRv" in p[-(len(case["tag"]) + 40):]) >= (case["n"] * 24 // 1000) // 10


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "edi_ref_driver")), reason="oracle/_ref/edi_ref_driver not built")
def test_edi_live_against_the_reference_packetiser():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_edi_golden
    from odr_audioenc_b200 import framing
    rng = np.random.RandomState(11)
    for trial in range(6):
        case = dict(tist=bool(trial & 1), delay_ms=int(rng.randint(0, 5000)), alignment=[0, 8, 8, 24, 0, 12][trial], tai=37,
                    start=int(rng.randint(10**9, 2 * 10**9)), tag="t%d" % trial, n=int(rng.randint(50, 700)),
                    frame_len=int(rng.choice([48, 144, 384, 576, 864])))
        frames, peaks = edi_cases.inputs(case, seed=trial)
        want = make_edi_golden.run_ref(case, frames, peaks)
        e = framing.EdiPacketiser(case["tist"], case["delay_ms"], case["alignment"], case["tai"], case["start"], case["tag"])
        assert e.packets(frames, case["frame_len"], peaks) == want, case


def test_edi_refuses_bad_arguments():
    from odr_audioenc_b200 import TlbError, framing
    with pytest.raises(TlbError):
        framing.EdiPacketiser(tagpacket_alignment=4)
    with pytest.raises(TlbError):
        framing.PftFragmenter(fec=1, chunk_len=208)


def _case_packets(name, count):
    from odr_audioenc_b200 import framing
    case = edi_cases.CASES[name]
    frames, peaks = edi_cases.inputs(case)
    e = framing.EdiPacketiser(case["tist"], case["delay_ms"], case["alignment"], case["tai"], case["start"], case["tag"])
    return e.packets(frames[:count], case["frame_len"], peaks[:count])


@pytest.mark.parametrize("name,fec,chunk_len,count", edi_cases.PFT_CASES, ids=["%s-m%d-k%d" % c[:3] for c in edi_cases.PFT_CASES])
def test_pft_fragments_equal_the_reference_pft_layer(name, fec, chunk_len, count):
    """PF fragments (TS 102 821 section 7) against fixtures from the reference's own PFT class + Reed-Solomon code
    (contrib/edioutput/PFT.cpp, contrib/ReedSolomon.cpp, contrib/fec: oracle/_ref/edi_ref_driver), and an independent
    check of the structure: header CRC, counters, de-interleaving back to the AF packet, RS parity = zero syndromes"""
    from odr_audioenc_b200 import framing
    g = np.load(os.path.join(GOLD, "pft_%s_m%d_k%d.npz" % (name, fec, chunk_len)))
    ends = np.cumsum(g["sizes"].astype(np.int64))
    want = [g["data"][e - s:e].tobytes() for s, e in zip(g["sizes"], ends)]
    packets = _case_packets(name, count)
    p = framing.PftFragmenter(fec, chunk_len)
    got, per_packet = [], []
    for af in packets:
        fl = p.fragments(af)
        per_packet.append(len(fl))
        got += fl
    assert per_packet == list(g["per_packet"])
    bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert len(got) == len(want) and not bad, "fragments differing from the reference PFT: %s" % bad[:8]
    # structure of the first packet's fragments
    fl, af = got[:per_packet[0]], packets[0]
    rs = fec > 0
    payloads = []
    for i, f in enumerate(fl):
        hdr = 14 if rs else 12
        assert f[:2] == b"PF" and struct.unpack(">H", f[2:4])[0] == 0
        assert int.from_bytes(f[4:7], "big") == i and int.from_bytes(f[7:10], "big") == len(fl)
        plen = struct.unpack(">H", f[10:12])[0]
        assert bool(plen & 0x8000) == rs and not plen & 0x4000 and (plen & 0x3FFF) == len(f) - hdr - 2
        assert struct.unpack(">H", f[hdr:hdr + 2])[0] == _crc16_ccitt(f[:hdr])
        payloads.append(f[hdr + 2:])
    if not rs:
        assert b"".join(payloads) == af and max(len(x) for x in payloads) <= 1400
    else:
        k, z = fl[0][12], fl[0][13]
        c = (len(af) + z) // k
        assert c * k == len(af) + z and k <= (chunk_len or 207)
        n, size = len(fl), len(payloads[0])
        block = bytes(payloads[j % n][j // n] for j in range(n * size))[:c * (k + 48)]   # de-interleave
        chunks = [block[i * (k + 48):(i + 1) * (k + 48)] for i in range(c)]
        assert (b"".join(ch[:k] for ch in chunks))[:len(af)] == af
        # every code word (data, zero fill up to 207, parity) has zero syndromes at alpha^1 .. alpha^48
        exp, v = [], 1
        for _ in range(255):
            exp.append(v)
            v <<= 1
            if v & 0x100:
                v ^= 0x11d
        log = {e: i for i, e in enumerate(exp)}
        for ch in chunks[:3]:
            word = ch[:k] + bytes(207 - k) + ch[k:]
            for root in (1, 2, 17, 48):
                acc = 0
                for b in word:   # Horner, highest power first
                    acc = (exp[(log[acc] + root) % 255] if acc else 0) ^ b
                assert acc == 0
    assert max(len(f) for f in got) <= 1400 + 14 or not rs


def _gf():
    exp, v = [], 1
    for _ in range(255):
        exp.append(v)
        v <<= 1
        if v & 0x100:
            v ^= 0x11d
    return exp * 2, {e: i for i, e in enumerate(exp)}


def _rs_fill_erasures(word, erased):
    """word: 255 symbols (highest power first) of RS(255,207) with roots alpha^1..alpha^48, `erased` = indices whose
    value is unknown (set to 0 in word).  Solves S_j = sum_l e_l X_l^j, j = 1..E (Gaussian elimination over GF(256))."""
    exp, log = _gf()
    mul = lambda a, b: exp[log[a] + log[b]] if a and b else 0
    inv = lambda a: exp[255 - log[a]]
    E = len(erased)
    assert E <= 48
    X = [exp[(254 - i) % 255] for i in erased]
    rows = []
    for j in range(1, E + 1):
        syn = 0
        for b in word:   # Horner at alpha^j
            syn = (exp[log[syn] + j] if syn else 0) ^ b
        rows.append([exp[(log[x] * j) % 255] for x in X] + [syn])
    for c in range(E):
        piv = next(r for r in range(c, E) if rows[r][c])
        rows[c], rows[piv] = rows[piv], rows[c]
        iv = inv(rows[c][c])
        rows[c] = [mul(v, iv) for v in rows[c]]
        for r in range(E):
            if r != c and rows[r][c]:
                f = rows[r][c]
                rows[r] = [a ^ mul(f, b) for a, b in zip(rows[r], rows[c])]
    out = list(word)
    for l, i in enumerate(erased):
        out[i] = rows[l][E]
    return out


@pytest.mark.parametrize("name,fec,chunk_len,count", [c for c in edi_cases.PFT_CASES if c[1] > 0],
                         ids=["%s-m%d-k%d" % c[:3] for c in edi_cases.PFT_CASES if c[1] > 0])
def test_pft_survives_the_loss_of_m_fragments(name, fec, chunk_len, count):
    """what the protection is for (TS 102 821 section 7.3): with fec = m, ANY m fragments of a packet may be lost and a
    receiver still recovers the AF packet -- encode, erase, decode with an independent erasure decoder"""
    from odr_audioenc_b200 import framing
    rng = np.random.RandomState(fec * 131 + count)
    p = framing.PftFragmenter(fec, chunk_len)
    for af in _case_packets(name, min(count, 3)):
        fl = p.fragments(af)
        n = len(fl)
        k, z = fl[0][12], fl[0][13]
        c = (len(af) + z) // k
        payloads = [f[16:] for f in fl]
        size = len(payloads[0])
        for lost in (list(range(fec)), list(range(n - fec, n)), sorted(rng.choice(n, size=fec, replace=False).tolist())):
            block, gone = bytearray(n * size), []
            for j in range(n * size):
                if j % n in lost:
                    gone.append(j)
                else:
                    block[j] = payloads[j % n][j // n]
            data = b""
            for ci in range(c):
                lo = ci * (k + 48)
                word = list(block[lo:lo + k]) + [0] * (207 - k) + list(block[lo + k:lo + k + 48])
                idx = [(j - lo) if j - lo < k else (j - lo - k + 207) for j in gone if lo <= j < lo + k + 48]
                assert len(idx) <= 48, "more erasures in a code word than the code corrects"
                data += bytes(_rs_fill_erasures(word, idx)[:k])
            assert data[:len(af)] == af, "lost fragments %s not recovered" % lost


class _FakePadEnc(threading.Thread):
    """stands in for ODR-PadEnc: answers each request [1, padlen] on /tmp/<ident>.padenc with [2] + record"""

    def __init__(self, ident, records):
        super().__init__(daemon=True)
        self.path = "/tmp/%s.padenc" % ident
        self.reply_to = "/tmp/%s.audioenc" % ident
        if os.path.exists(self.path):
            os.unlink(self.path)
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
        self.sock.bind(self.path)
        self.sock.settimeout(5.0)
        self.records, self.requests = list(records), []

    def run(self):
        try:
            while self.records:
                req = self.sock.recv(16)
                self.requests.append(req)
                rec = self.records.pop(0)
                if rec is not None:
                    self.sock.sendto(b"\x02" + rec, self.reply_to)
        except socket.timeout:
            pass
        finally:
            self.sock.close()
            os.unlink(self.path)


def test_pad_socket_protocol():
    import time
    from odr_audioenc_b200 import TlbError, framing
    ident = "tlbtest%d" % os.getpid()
    pad_len = 23
    with pytest.raises(TlbError):   # a name that does not fit a socket path is refused, not cut short
        framing.PadSocket("x" * 120)
    rec_a = bytes(range(pad_len - 8)) + b"\xAA" * 8 + bytes([8])          # 8 bytes used
    rec_b = bytes(pad_len - 2) + b"\x40\x00" + bytes([2])                   # F-PAD only
    p = framing.PadSocket(ident)
    used, rec = p.request(pad_len)          # ODR-PadEnc not running: no PAD, zero record (src/PadInterface.cpp:88-99)
    assert used == 0 and not rec.any()
    fake = _FakePadEnc(ident, [rec_a, None, rec_b, rec_a[:-3], bytes(pad_len) + b"\x01"])
    fake.start()

    def request_until_answer():
        # the reply to request n is normally picked up by request n (same host), at the latest by a later one
        for _ in range(200):
            used, rec = p.request(pad_len)
            if used:
                return used, rec
            time.sleep(0.005)
        raise AssertionError("no PAD reply")

    used, rec = request_until_answer()
    assert used == 8 and rec.tobytes() == rec_a
    used, rec = request_until_answer()      # (the request that got no reply returned 0 and was retried)
    assert used == 2 and rec.tobytes() == rec_b
    with pytest.raises(TlbError, match="Incorrect PAD length"):
        request_until_answer()
    with pytest.raises(TlbError, match="Invalid X-PAD length"):
        request_until_answer()
    fake.join(timeout=6)
    assert fake.requests and all(r == bytes([1, pad_len]) for r in fake.requests)
    p.close()
    assert not os.path.exists("/tmp/%s.audioenc" % ident)
