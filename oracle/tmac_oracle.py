"""tmac_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the two CPU checkers (oracle/liboracle.so = scalar C restatement,
oracle/_ref/libtmac_ref.so = the reference's own intrinsics compiled in place) plus numpy
restatements of the offline weight packer and seeded synthetic-input generators
(SURVEY.md section 8d).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product never does.

Reference citations are relative to /root/reference/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtmac_ref.so")

ALPHAS = (0.5, 1.0, 2.0, 4.0)  # python/t_mac/utils.py:6-8

_f32p = C.POINTER(C.c_float)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    """Compile the checkers (make -C oracle). Building the checker is not using it."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.isdir("/root/reference/python/t_mac/intrins") and not os.path.exists(REF_SO)
    ):
        subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class _Lib:
    """Uniform wrapper; prefix is 'tmo_' (oracle) or 'tmr_' (reference build)."""

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = "port" if prefix == "tmo_" else "reference"

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def lut_ctor(self, b: np.ndarray, lut_scale: float) -> Tuple[np.ndarray, float, float]:
        b = np.ascontiguousarray(b, np.float32)
        act_k = b.size
        qlut = np.zeros((act_k // 4, 16), np.int8)
        ls = np.array([lut_scale], np.float32)
        lb = np.zeros(1, np.float32)
        rc = self._f("lut_ctor_g4_int8")(act_k, _ptr(qlut, _i8p), _ptr(b, _f32p), _ptr(ls, _f32p), _ptr(lb, _f32p))
        assert rc == 0
        return qlut, float(ls[0]), float(lb[0])

    def partial_max(self, b32: np.ndarray, start: float = 0.0) -> float:
        b32 = np.ascontiguousarray(b32, np.float32)
        ls = np.array([start], np.float32)
        self._f("partial_max_g4_int8_k8")(_ptr(ls, _f32p), _ptr(b32, _f32p))
        return float(ls[0])

    def preprocessor(self, B: np.ndarray, ags: int):
        B = np.ascontiguousarray(B, np.float32)
        if B.ndim == 1:
            B = B[None]
        N, K = B.shape
        ls = np.zeros((N, K // ags), np.float32)
        lb = np.zeros((N, K // ags), np.float32)
        qlut = np.zeros((N, K // 4, 16), np.int8)
        rc = self._f("preprocessor")(K, N, ags, _ptr(B, _f32p), _ptr(ls, _f32p), _ptr(lb, _f32p), _ptr(qlut, _i8p))
        if rc != 0:
            raise ValueError("preprocessor rejected shape")
        return qlut, ls, lb

    def qgemm(self, cfg: "Config", A: np.ndarray, scales: np.ndarray, qlut, ls, lb) -> np.ndarray:
        N = qlut.shape[0]
        out = np.zeros((N, cfg.Mout), np.float32)
        A = np.ascontiguousarray(A, np.uint8)
        scales = np.ascontiguousarray(scales, np.float32)
        rc = self._f("qgemm")(
            cfg.Mout, cfg.K, N, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size,
            int(cfg.zero_point), int(cfg.one_scale), _ptr(A, _u8p), _ptr(qlut, _i8p), _ptr(scales, _f32p),
            _ptr(ls, _f32p), _ptr(lb, _f32p), _ptr(out, _f32p))
        if rc != 0:
            raise ValueError("qgemm rejected config %r" % (cfg,))
        return out

    def cbits(self, cfg: "Config", A: np.ndarray, qlut) -> np.ndarray:
        N = qlut.shape[0]
        out = np.zeros((N, cfg.Mout * cfg.bits), np.int32)
        A = np.ascontiguousarray(A, np.uint8)
        rc = self._f("qgemm_cbits")(cfg.Mout, cfg.K, N, cfg.bits, cfg.bm, cfg.kfactor, _ptr(A, _u8p),
                                    _ptr(qlut, _i8p), _ptr(out, _i32p))
        if rc != 0:
            raise ValueError("cbits rejected config")
        return out

    # reference build only -------------------------------------------------
    def set_threads(self, n: int) -> int:
        return self.lib.tmr_set_threads(int(n))

    def gemv_mt(self, cfg: "Config", A, scales, B, work=None):
        B = np.ascontiguousarray(B, np.float32)
        if B.ndim == 1:
            B = B[None]
        N, K = B.shape
        if work is None:
            work = (np.zeros((N, K // 4, 16), np.int8), np.zeros((N, K // cfg.act_group_size), np.float32),
                    np.zeros((N, K // cfg.act_group_size), np.float32), np.zeros((N, cfg.Mout), np.float32))
        qlut, ls, lb, out = work
        rc = self.lib.tmr_gemv_mt(
            cfg.Mout, cfg.K, N, cfg.bits, cfg.bm, cfg.kfactor, cfg.group_size, cfg.act_group_size,
            int(cfg.zero_point), int(cfg.one_scale), _ptr(A, _u8p), _ptr(scales, _f32p), _ptr(B, _f32p),
            _ptr(qlut, _i8p), _ptr(ls, _f32p), _ptr(lb, _f32p), _ptr(out, _f32p))
        assert rc == 0
        return out


_cache = {}


def load_oracle() -> _Lib:
    if "o" not in _cache:
        if not os.path.exists(ORACLE_SO):
            build()
        _cache["o"] = _Lib(ORACLE_SO, "tmo_")
    return _cache["o"]


def load_ref() -> Optional[_Lib]:
    """The reference's own code, when oracle/_ref was built (here, from /root/reference)."""
    if "r" not in _cache:
        _cache["r"] = _Lib(REF_SO, "tmr_") if os.path.exists(REF_SO) else None
    return _cache["r"]


# ---------------------------------------------------------------------------------------
# Problem description (mirrors TMACGeMMConfig + the compile-time options of the reference)
# ---------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Config:
    Mout: int
    K: int
    bits: int
    bm: int = 0
    kfactor: int = 16
    group_size: int = 128
    act_group_size: int = 64
    zero_point: bool = False
    one_scale: bool = False  # BitNet-like: m_groups=1, act_group_size==K, int32 accumulation
    simd_n_in: int = 16
    simd_n_out: int = 8

    def resolved(self) -> "Config":
        bm = self.bm or default_bm(self.Mout * self.bits, self.bits)
        ags = self.K if self.one_scale else self.act_group_size
        return Config(self.Mout, self.K, self.bits, bm, self.kfactor, self.group_size, ags,
                      self.zero_point and not self.one_scale, self.one_scale, self.simd_n_in, self.simd_n_out)

    @property
    def scales_size(self) -> int:
        if self.one_scale:
            return 1
        return self.Mout * (self.K // self.group_size) * (2 if self.zero_point else 1)

    @property
    def n_tile_num(self) -> int:
        return self.Mout * self.bits // self.bm


def default_bm(M: int, bits: int) -> int:
    """First valid tile of the reference's tuning knob list, python/t_mac/ops/qgemm.py:99-103."""
    cands = [192, 384, 576, 768] if bits == 3 else [256, 128, 512, 1024, 320, 640]
    for bm in cands:
        if M % bm == 0 and bm % bits == 0:
            return bm
    raise ValueError("no valid bm for M=%d bits=%d" % (M, bits))


# ---------------------------------------------------------------------------------------
# Offline weight packer -- restatement of python/t_mac/weights.py:5-88 by explicit index
# algebra (no reshape chain).  Reference byte address of bit-plane row p, K-group kg:
#   tile = p // bm, slab = (p % bm) // 32, s = p % 16, half = (p % 32) // 16   (weights.py:65-70)
#   byte = ((tile*(KG/kf) + kg//kf) * (bm/32) + slab) * kf*16 + (kg % kf)*16 + s   (weights.py:69-73)
#   nibble `half` of that byte holds idx = sum_j bit(w[row, 4kg+j], b) << j           (weights.py:57-60)
# with p = (row//8)*8*bits + b*8 + row%8                                              (weights.py:65)
# ---------------------------------------------------------------------------------------
def plane_indices(w: np.ndarray, bits: int) -> np.ndarray:
    """w [Mout][K] uint8 in [0, 2^bits) -> idx [Mout][bits][K/4] 4-bit LUT indices."""
    Mout, K = w.shape
    w4 = w.reshape(Mout, K // 4, 4).astype(np.uint8)
    out = np.zeros((Mout, bits, K // 4), np.uint8)
    for b in range(bits):
        bitv = (w4 >> b) & 1
        out[:, b, :] = bitv[:, :, 0] | (bitv[:, :, 1] << 1) | (bitv[:, :, 2] << 2) | (bitv[:, :, 3] << 3)
    return out


def pack_reference_layout(w: np.ndarray, scales: np.ndarray, zeros: Optional[np.ndarray], cfg: Config):
    """Returns (A uint8 [M/bm][K/4][bm/2], Scales float32 flat) in the reference's run-time layout."""
    cfg = cfg.resolved()
    Mout, K, bits, bm, kf = cfg.Mout, cfg.K, cfg.bits, cfg.bm, cfg.kfactor
    assert w.shape == (Mout, K) and w.dtype == np.uint8
    KG = K // 4
    idx = plane_indices(w, bits)  # [Mout][bits][KG]
    row = np.arange(Mout)[:, None]
    b = np.arange(bits)[None, :]
    p = (row // 8) * 8 * bits + b * 8 + row % 8  # [Mout][bits] plane-row ids
    tile, slab, s, half = p // bm, (p % bm) // 32, p % 16, (p % 32) // 16
    kg = np.arange(KG)
    base = ((tile[..., None] * (KG // kf) + kg // kf) * (bm // 32) + slab[..., None]) * kf * 16 + (kg % kf) * 16 + s[..., None]
    A = np.zeros(Mout * bits * KG // 2, np.uint8)
    shift = (half[..., None] * 4).astype(np.uint8)
    np.bitwise_or.at(A, base.ravel(), (idx << shift).ravel())
    A = A.reshape(Mout * bits // bm, KG, bm // 2)

    if cfg.one_scale:
        S = np.asarray(scales, np.float32).reshape(-1)[:1].copy()
    else:
        gs = cfg.group_size
        rows = bm // bits
        sc = np.asarray(scales, np.float32).reshape(Mout // rows, rows, K // gs).transpose(0, 2, 1)  # weights.py:77
        sc = sc.reshape(Mout // rows, K // gs, rows // 8, 8)
        if cfg.zero_point:
            zc = np.asarray(zeros, np.float32).reshape(Mout // rows, rows, K // gs).transpose(0, 2, 1)
            zc = zc.reshape(Mout // rows, K // gs, rows // 8, 8)
            sc = np.stack([sc, zc], axis=-2)  # [..][rows/8][2][8], weights.py:82
        S = np.ascontiguousarray(sc, np.float32).reshape(-1)
    return A, S


def dense_reference(w, scales, zeros, x, cfg: Config) -> np.ndarray:
    """fp64 dense dequant matmul, semantics of tests/test_e2e.py:69-77: W = (w - 2^(bits-1))*s - z."""
    cfg = cfg.resolved()
    wf = w.astype(np.float64) - (1 << (cfg.bits - 1))
    if cfg.one_scale:
        W = wf * float(np.asarray(scales).reshape(-1)[0])
    else:
        gs = cfg.group_size
        W = wf.reshape(cfg.Mout, cfg.K // gs, gs) * np.asarray(scales, np.float64)[:, :, None]
        if cfg.zero_point:
            W = W - np.asarray(zeros, np.float64)[:, :, None]
        W = W.reshape(cfg.Mout, cfg.K)
    x = np.asarray(x, np.float64)
    if x.ndim == 1:
        x = x[None]
    return x @ W.T


# ---------------------------------------------------------------------------------------
# Seeded synthetic inputs: PCG64, fp16-representable scales/zeros/activations.
# ---------------------------------------------------------------------------------------
def make_problem(cfg: Config, seed: int = 0, N: int = 1):
    cfg = cfg.resolved()
    rng = np.random.default_rng(seed)
    if cfg.one_scale:
        w = (rng.integers(-1, 2, size=(cfg.Mout, cfg.K)) + 2).astype(np.uint8)  # ternary, convert_hf_to_gguf.py:1909-1917
        scales = np.array([0.037], np.float16).astype(np.float32)
        zeros = None
    else:
        w = rng.integers(0, 1 << cfg.bits, size=(cfg.Mout, cfg.K), dtype=np.uint8)
        scales = (np.abs(rng.standard_normal((cfg.Mout, cfg.K // cfg.group_size))) * 0.01 + 1e-4).astype(np.float16).astype(np.float32)
        zeros = (rng.standard_normal((cfg.Mout, cfg.K // cfg.group_size)) * 0.01).astype(np.float16).astype(np.float32) if cfg.zero_point else None
    x = rng.standard_normal((N, cfg.K)).astype(np.float16).astype(np.float32)
    return w, scales, zeros, x


def nmse(ref: np.ndarray, out: np.ndarray) -> float:
    ref = ref.astype(np.float64)
    out = out.astype(np.float64)
    return float(np.mean((ref - out) ** 2) / max(np.mean(ref ** 2), 1e-30))
