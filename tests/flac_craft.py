"""A tiny FLAC stream builder written from the format (RFC 9639) -- TEST INFRASTRUCTURE.  It produces streams that use
features the encoders in this repository never emit (LPC orders above 12, Rice partition orders above 8, explicit
8- and 16-bit block-size fields, 5-bit parameters with escapes in the middle of a subframe, wasted bits on LPC
subframes, ...), together with the samples a decoder must return, so that the device decoder and the independent CPU
decoder can both be checked against a third, much simpler statement of the format.
"""
import numpy as np


def _crc8(data):
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def _crc16(data):
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


class Bits:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        for k in range(n - 1, -1, -1):
            self.bits.append((value >> k) & 1)

    def put_signed(self, value, n):
        self.put(value & ((1 << n) - 1), n)

    def unary(self, q):
        self.bits.extend([0] * q)
        self.bits.append(1)

    def pad(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        assert len(self.bits) % 8 == 0
        a = np.array(self.bits, dtype=np.uint8).reshape(-1, 8)
        return bytes(np.packbits(a, axis=1).reshape(-1))


def _utf8(n):
    if n < 0x80:
        return bytes([n])
    out = []
    for extra, lead in ((1, 0xC0), (2, 0xE0), (3, 0xF0), (4, 0xF8), (5, 0xFC), (6, 0xFE)):
        if n < (1 << (6 * extra + (6 - extra))):
            for k in range(extra):
                out.append(0x80 | ((n >> (6 * k)) & 0x3F))
            out.append(lead | (n >> (6 * extra)))
            return bytes(reversed(out))
    raise ValueError(n)


def _zigzag(r):
    return (r << 1) if r >= 0 else ((-r) << 1) - 1


def residual(w, res, block, order, part_order, method, params):
    """params[p]: Rice parameter, or ('esc', width) for an escaped partition."""
    w.put(method, 2)
    w.put(part_order, 4)
    plen, esc = (5, 31) if method else (4, 15)
    psize = block >> part_order
    i = 0
    for p in range(1 << part_order):
        n = psize - (order if p == 0 else 0)
        par = params[p]
        if isinstance(par, tuple):
            w.put(esc, plen)
            w.put(par[1], 5)
            for r in res[i:i + n]:
                if par[1]:
                    w.put_signed(int(r), par[1])
        else:
            w.put(par, plen)
            for r in res[i:i + n]:
                z = _zigzag(int(r))
                w.unary(z >> par)
                if par:
                    w.put(z & ((1 << par) - 1), par)
        i += n
    assert i == len(res)


def subframe_lpc(w, samples, bps, order, coefs, precision, shift, part_order, method, params, wasted=0):
    """samples: the decoded values (multiples of 2^wasted); the residual is worked out here."""
    x = [int(v) >> wasted for v in samples]
    w.put(0, 1)
    w.put(31 + order, 6)
    if wasted:
        w.put(1, 1)
        w.unary(wasted - 1)
    else:
        w.put(0, 1)
    b = bps - wasted
    for v in x[:order]:
        w.put_signed(v, b)
    w.put(precision - 1, 4)
    w.put_signed(shift, 5)
    for c in coefs:
        w.put_signed(int(c), precision)
    res = []
    for i in range(order, len(x)):
        pred = sum(int(coefs[j]) * x[i - 1 - j] for j in range(order)) >> shift
        res.append(x[i] - pred)
    residual(w, res, len(x), order, part_order, method, params)


def subframe_fixed(w, samples, bps, order, part_order, method, params, wasted=0):
    x = [int(v) >> wasted for v in samples]
    w.put(0, 1)
    w.put(8 + order, 6)
    if wasted:
        w.put(1, 1)
        w.unary(wasted - 1)
    else:
        w.put(0, 1)
    for v in x[:order]:
        w.put_signed(v, bps - wasted)
    d = np.array(x, dtype=object)
    for _ in range(order):
        d = np.concatenate([d[:1], d[1:] - d[:-1]])
    # after `order` passes d[i] (i >= order) is the order-th difference
    res = [int(v) for v in d[order:]]
    residual(w, res, len(x), order, part_order, method, params)


def frame(number, block, rate_code, ch_code, depth_code, subframes, explicit_block=None, rate_trailer=b""):
    """subframes: callables that write one subframe into the bit writer.  explicit_block: 8 or 16 forces the explicit field;
    rate_trailer: the bytes behind the header for rate codes 12 (one byte), 13 and 14 (two)."""
    hdr = bytearray([0xFF, 0xF8])
    if explicit_block == 8:
        bs_code = 6
    elif explicit_block == 16:
        bs_code = 7
    else:
        bs_code = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14,
                   32768: 15}[block]
    hdr.append((bs_code << 4) | rate_code)
    hdr.append((ch_code << 4) | (depth_code << 1))
    hdr += _utf8(number)
    if bs_code == 6:
        hdr.append(block - 1)
    elif bs_code == 7:
        hdr += bytes([(block - 1) >> 8, (block - 1) & 0xFF])
    hdr += rate_trailer
    hdr.append(_crc8(hdr))
    w = Bits()
    for s in subframes:
        s(w)
    w.pad()
    body = bytes(hdr) + w.bytes()
    c = _crc16(body)
    return body + bytes([c >> 8, c & 0xFF])


def stream(frames, channels, bits, rate, min_block, max_block, total_samples, extra_metadata=()):
    si = bytearray()
    si += bytes([min_block >> 8, min_block & 0xFF, max_block >> 8, max_block & 0xFF])
    si += bytes(6)
    v = (rate << 44) | ((channels - 1) << 41) | ((bits - 1) << 36) | total_samples
    si += v.to_bytes(8, "big")
    si += bytes(16)
    out = bytearray(b"fLaC")
    blocks = [(0, bytes(si))] + list(extra_metadata)
    for k, (typ, body) in enumerate(blocks):
        last = 0x80 if k == len(blocks) - 1 else 0
        out += bytes([last | typ]) + len(body).to_bytes(3, "big") + body
    for f in frames:
        out += f
    return bytes(out)
