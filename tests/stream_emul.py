"""stream_emul.py -- TEST INFRASTRUCTURE: numpy emulation of the PRMT / DP4A inner loop of
gemv_kernel (t-mac_b200/csrc/tmac_kernels.cuh, Quad<PB,SYM>::run) over the stream layout
(tmac_layout.h).  Lets the CPU suite pin the layout encoder and the lookup algebra without a GPU.
The emulation is deliberately written against the PTX semantics of prmt.b32 / dp4a, not against
the oracle."""
import numpy as np


def prmt(a, b, s):
    """prmt.b32 default mode: selector nibble bits[2:0] pick a byte of {b,a}; bit 3 replicates its sign."""
    a = np.asarray(a, np.uint64); b = np.asarray(b, np.uint64); s = np.asarray(s, np.uint64)
    pool = a | (b << np.uint64(32))
    out = np.zeros(np.broadcast(a, b, s).shape, np.uint64)
    for i in range(4):
        sel = (s >> np.uint64(4 * i)) & np.uint64(0xF)
        byte = (pool >> ((sel & np.uint64(7)) * np.uint64(8))) & np.uint64(0xFF)
        msb = np.where((byte & np.uint64(0x80)) != 0, np.uint64(0xFF), np.uint64(0))
        byte = np.where((sel & np.uint64(8)) != 0, msb, byte)
        out |= byte << np.uint64(8 * i)
    return out.astype(np.uint32)


def dp4a(a, b, c):
    a = np.asarray(a, np.uint32); b = np.asarray(b, np.uint32)
    acc = np.asarray(c, np.int64).copy()
    for i in range(4):
        ai = ((a >> np.uint32(8 * i)) & np.uint32(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64)
        bi = ((b >> np.uint32(8 * i)) & np.uint32(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64)
        acc = acc + ai * bi
    return acc


def plane_weight_regs(bits):
    if bits >= 3:
        w = [1, 2, 4, 8 if bits == 4 else 0]
    elif bits == 2:
        w = [1, 2, 1, 2]
    else:
        w = [1, 1, 1, 1]
    p = sum((x & 0xFF) << (8 * i) for i, x in enumerate(w))
    n = sum(((-x) & 0xFF) << (8 * i) for i, x in enumerate(w))
    return np.uint32(p), np.uint32(n)


def _tables(qlut_g, sym):
    """qlut_g int8 [..., 16] -> list of uint32 register arrays (2 if sym else 4)."""
    q = np.ascontiguousarray(qlut_g).view(np.uint8).astype(np.uint32)
    def word(bs):
        return bs[..., 0] | (bs[..., 1] << 8) | (bs[..., 2] << 16) | (bs[..., 3] << 24)
    lo = [word(q[..., 0:4]), word(q[..., 4:8])]
    if sym:
        return lo
    hi = q[..., 8:16][..., ::-1]  # G16[8+j] = L[15-j]
    return lo + [word(hi[..., 0:4]), word(hi[..., 4:8])]


def quad(pb, sym, w, tabs, acc, wtx, wty):
    """w: uint32 [4][lanes]; tabs: per word k list of register arrays; acc: int64 [RW][lanes]."""
    M7, M4, M8, C = np.uint32(0x77777777), np.uint32(0x44444444), np.uint32(0x88888888), np.uint32(0x32103210)
    sh16 = np.uint32(16)
    w2 = np.uint32(int(wtx) << 1)

    def sign_sel(x):       # tmac_kernels.cuh sign_sel<SYM>
        return ((x & M8) | C) if sym else (((x >> np.uint32(1)) & M4) | C)

    def dp4a_signed(v, sel_, acc_):   # tmac_kernels.cuh dp4a_signed: v*2w(1-neg) - v*w
        return dp4a(v, prmt(w2, w2, sel_), dp4a(v, wty, acc_))
    if pb == 4:
        for k in range(4):
            wj = w[k] & M7; ws = sign_sel(w[k])
            t = tabs[k]
            if sym:
                acc[0] = dp4a_signed(prmt(t[0], t[1], wj), ws, acc[0])
                acc[1] = dp4a_signed(prmt(t[0], t[1], wj >> sh16), ws >> sh16, acc[1])
            else:
                acc[0] = dp4a(prmt(prmt(t[0], t[1], wj), prmt(t[2], t[3], wj), ws), wtx, acc[0])
                acc[1] = dp4a(prmt(prmt(t[0], t[1], wj >> sh16), prmt(t[2], t[3], wj >> sh16), ws >> sh16), wtx, acc[1])
    elif pb == 2:
        for pr in range(2):
            we, wo = w[2 * pr], w[2 * pr + 1]
            je, jo = we & M7, wo & M7
            se, so = sign_sel(we), sign_sel(wo)
            te, to = tabs[2 * pr], tabs[2 * pr + 1]
            if sym:
                v0a, v0b = prmt(te[0], te[1], je), prmt(te[0], te[1], je >> sh16)
                v1a, v1b = prmt(to[0], to[1], jo), prmt(to[0], to[1], jo >> sh16)
                acc[0] = dp4a_signed(prmt(v0a, v1a, 0x5410), se, acc[0])
                acc[1] = dp4a_signed(prmt(v0a, v1a, 0x7632), se >> sh16, acc[1])
                acc[2] = dp4a_signed(prmt(v0b, v1b, 0x5410), so, acc[2])
                acc[3] = dp4a_signed(prmt(v0b, v1b, 0x7632), so >> sh16, acc[3])
            else:
                l0a, l0b = prmt(te[0], te[1], je), prmt(te[0], te[1], je >> sh16)
                h0a, h0b = prmt(te[2], te[3], je), prmt(te[2], te[3], je >> sh16)
                l1a, l1b = prmt(to[0], to[1], jo), prmt(to[0], to[1], jo >> sh16)
                h1a, h1b = prmt(to[2], to[3], jo), prmt(to[2], to[3], jo >> sh16)
                acc[0] = dp4a(prmt(prmt(l0a, l1a, 0x5410), prmt(h0a, h1a, 0x5410), se), wtx, acc[0])
                acc[1] = dp4a(prmt(prmt(l0a, l1a, 0x7632), prmt(h0a, h1a, 0x7632), se >> sh16), wtx, acc[1])
                acc[2] = dp4a(prmt(prmt(l0b, l1b, 0x5410), prmt(h0b, h1b, 0x5410), so), wtx, acc[2])
                acc[3] = dp4a(prmt(prmt(l0b, l1b, 0x7632), prmt(h0b, h1b, 0x7632), so >> sh16), wtx, acc[3])
    else:
        def transpose4(v):
            t01, t23 = prmt(v[0], v[1], 0x5140), prmt(v[2], v[3], 0x5140)
            u01, u23 = prmt(v[0], v[1], 0x7362), prmt(v[2], v[3], 0x7362)
            return [prmt(t01, t23, 0x5410), prmt(t01, t23, 0x7632), prmt(u01, u23, 0x5410), prmt(u01, u23, 0x7632)]
        s = [sign_sel(w[k]) for k in range(4)]
        j = [w[k] & M7 for k in range(4)]
        def sel(r, base):
            x = s[base + (r >> 1)]
            return (x >> sh16) if (r & 1) else x
        if sym:
            xa = transpose4([prmt(tabs[k][0], tabs[k][1], j[k]) for k in range(4)])
            xb = transpose4([prmt(tabs[k][0], tabs[k][1], j[k] >> sh16) for k in range(4)])
            for r in range(4):
                acc[r] = dp4a_signed(xa[r], sel(r, 0), acc[r])
                acc[4 + r] = dp4a_signed(xb[r], sel(r, 2), acc[4 + r])
        else:
            xla = transpose4([prmt(tabs[k][0], tabs[k][1], j[k]) for k in range(4)])
            xha = transpose4([prmt(tabs[k][2], tabs[k][3], j[k]) for k in range(4)])
            xlb = transpose4([prmt(tabs[k][0], tabs[k][1], j[k] >> sh16) for k in range(4)])
            xhb = transpose4([prmt(tabs[k][2], tabs[k][3], j[k] >> sh16) for k in range(4)])
            for r in range(4):
                acc[r] = dp4a(prmt(xla[r], xha[r], sel(r, 0)), wtx, acc[r])
                acc[4 + r] = dp4a(prmt(xlb[r], xhb[r], sel(r, 2)), wtx, acc[4 + r])
    return acc


def emulate_int_sums(stream: np.ndarray, lay, qlut: np.ndarray, Mout: int, bits: int, sym: bool):
    """stream: uint8 bytes of the stream layout; lay: layout_out[12] of tmac_b200_debug_encode;
    qlut int8 [K/4][16].  Returns int64 [Mout] = sum over K of sum_b 2*alpha_b*sgn*LUT lookups."""
    pb, rw, rsb, nrsb, ck, qch, nchunk, sd, zp, one_scale, blk, wbytes = [int(x) for x in lay]
    wtx, wty = plane_weight_regs(bits)
    out = np.zeros(nrsb * rsb, np.int64)
    rsb_stride = blk * nchunk
    for r in range(nrsb):
        acc = [np.zeros(32, np.int64) for _ in range(rw)]
        for c in range(nchunk):
            base = r * rsb_stride + c * blk
            words = stream[base:base + wbytes].view(np.uint32).reshape(qch, 32, 4)
            for q in range(qch):
                g0 = (c * qch + q) * 4
                tabs = [_tables(qlut[g0 + k], sym) for k in range(4)]
                w = [words[q, :, k].copy() for k in range(4)]
                acc = quad(pb, sym, w, tabs, acc, wtx, wty)
        for i in range(rw):
            out[r * rsb + np.arange(32) * rw + i] = acc[i]
    return out[:Mout]


def decode_scales(stream: np.ndarray, lay, Mout: int):
    """Returns (scales, zeros) [Mout][nchunk] float32 as stored in the stream."""
    pb, rw, rsb, nrsb, ck, qch, nchunk, sd, zp, one_scale, blk, wbytes = [int(x) for x in lay]
    sc = np.zeros((nrsb * rsb, nchunk), np.float32)
    zr = np.zeros((nrsb * rsb, nchunk), np.float32)
    dt = np.float16 if sd == 2 else np.float32
    for r in range(nrsb):
        for c in range(nchunk):
            base = r * blk * nchunk + c * blk + wbytes
            sc[r * rsb:(r + 1) * rsb, c] = stream[base:base + rsb * sd].view(dt).astype(np.float32)
            if zp:
                zr[r * rsb:(r + 1) * rsb, c] = stream[base + rsb * sd:base + 2 * rsb * sd].view(dt).astype(np.float32)
    return sc[:Mout], zr[:Mout]
