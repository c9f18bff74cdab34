"""decode_harness.py -- stand-in for the reference's end-to-end measurement (`llama-bench` through tools/bench_e2e.py:19-29 and
tools/run_pipeline.py:279-331): a Llama-style decoder whose quantised linears (attn q/k/v/o, ffn gate/up/down -- the set the
reference's converter quantises, 3rdparty/llama.cpp/convert_hf_to_gguf.py:353-362,1897-1905) go through the C ABI of
libtmac_b200 and whose other operators (RMSNorm, RoPE, attention over a KV cache, SiLU*mul, residuals) are plain fp16/fp32
torch ops, in the real per-layer order, with synthetic weights at the real shapes.  The token step is captured in one CUDA
graph; `tokens_per_s()` reports the full step, bench.py reports the matmul-only chain next to it.

Not a model implementation: no tokenizer, no sampling, no checkpoint loading (SURVEY 8 f1 scope).  torch is used for device
memory, streams and the non-matmul ops only; every linear is a library launch.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

import tmac_b200 as tb


class QLinear:
    """One quantised linear resident in HBM: y = gemv(W, x) through tmac_b200_gemv (LUT build fused for the fp path)."""

    def __init__(self, mout: int, k: int, bits: int, group_size: int, zero_point: bool, seed: int, one_scale: bool = False, share: Optional["QLinear"] = None):
        self.mout, self.k, self.bits, self.gs, self.zp, self.one_scale = mout, k, bits, group_size, zero_point, one_scale
        if share is not None:                       # another layer with the same synthetic tensor: its own HBM copy, one host encode
            self.w, self.sc, self.z, self.wt = share.w, share.sc, share.z, tb.clone(share.wt)
            return
        rng = np.random.default_rng(seed)
        if one_scale:
            self.w = (rng.integers(-1, 2, size=(mout, k)) + 2).astype(np.uint8)
            self.sc = np.array([0.037], np.float16).astype(np.float32); self.z = None
        else:
            self.w = rng.integers(0, 1 << bits, size=(mout, k), dtype=np.uint8)
            self.sc = (np.abs(rng.standard_normal((mout, k // group_size))) * 0.01 + 1e-4).astype(np.float16).astype(np.float32)
            self.z = (rng.standard_normal((mout, k // group_size)) * 0.01).astype(np.float16).astype(np.float32) if zero_point else None
        bm = next(b for b in ((192, 384, 576, 768) if bits == 3 else (256, 128, 512, 1024, 320, 640)) if (mout * bits) % b == 0)
        cfg = tb.make_kcfg(mout, k, bits, bm, 16, group_size, k if one_scale else 64, zero_point, one_scale)
        self.wt = tb.upload_plain(cfg, self.w, self.sc, self.z)

    def __call__(self, x, out):
        tb.gemv(self.wt, 1, x, out)
        return out

    def dense(self, torch):
        """Dense dequantised fp32 weight W_real = (w - 2^(bits-1)) * s - z (tests/test_e2e.py:69-77 of the reference)."""
        w = torch.from_numpy(self.w.astype(np.float32)) - float(1 << (self.bits - 1))
        if self.one_scale:
            return (w * float(self.sc[0])).cuda()
        s = torch.from_numpy(np.repeat(self.sc, self.gs, axis=1))
        d = w * s
        if self.z is not None:
            d = d - torch.from_numpy(np.repeat(self.z, self.gs, axis=1))
        return d.cuda()

    def free(self):
        self.wt.free()


class DecodeLayer:
    def __init__(self, hidden: int, ffn: int, heads: int, bits: int, zero_point: bool, seed: int, share: Optional["DecodeLayer"] = None, group_size: int = 128):
        import torch
        self.torch, self.hidden, self.ffn, self.heads, self.hd = torch, hidden, ffn, heads, hidden // heads
        names = [("q", hidden, hidden), ("k", hidden, hidden), ("v", hidden, hidden), ("o", hidden, hidden), ("gate", ffn, hidden), ("up", ffn, hidden), ("down", hidden, ffn)]
        self.lin = {n: QLinear(m, k, bits, group_size, zero_point, seed + i, share=share.lin[n] if share else None) for i, (n, m, k) in enumerate(names)}
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.norm1 = (1.0 + 0.1 * torch.randn(hidden, generator=g)).cuda()
        self.norm2 = (1.0 + 0.1 * torch.randn(hidden, generator=g)).cuda()
        f = lambda n: torch.zeros((1, n), device="cuda")
        self.buf = {"h1": f(hidden), "q": f(hidden), "k": f(hidden), "v": f(hidden), "att": f(hidden), "o": f(hidden), "h2": f(hidden),
                    "gate": f(ffn), "up": f(ffn), "act": f(ffn), "down": f(hidden)}

    @staticmethod
    def rmsnorm(torch, x, w, eps=1e-5):
        return x * torch.rsqrt((x * x).mean(dim=-1, keepdim=True) + eps) * w

    def rope(self, x, pos_cos, pos_sin):
        t = x.view(self.heads, self.hd // 2, 2)
        a, b = t[..., 0], t[..., 1]
        return self.torch.stack((a * pos_cos - b * pos_sin, a * pos_sin + b * pos_cos), dim=-1).reshape(1, -1)

    def forward(self, x, kcache, vcache, pos: int, pos_cos, pos_sin, linear=None):
        """x [1, hidden] fp32 -> [1, hidden].  kcache / vcache [heads, ctx, hd]; `linear(name, inp, out)` overrides the library call
        (the dense fp32 reference layer in the tests)."""
        torch, b = self.torch, self.buf
        lin = linear or (lambda n, inp, out: self.lin[n](inp, out))
        b["h1"].copy_(self.rmsnorm(torch, x, self.norm1))
        lin("q", b["h1"], b["q"]); lin("k", b["h1"], b["k"]); lin("v", b["h1"], b["v"])
        q = self.rope(b["q"], pos_cos, pos_sin).view(self.heads, 1, self.hd)
        kcache[:, pos] = self.rope(b["k"], pos_cos, pos_sin).view(self.heads, self.hd)
        vcache[:, pos] = b["v"].view(self.heads, self.hd)
        att = torch.softmax((q @ kcache[:, :pos + 1].transpose(1, 2)) / math.sqrt(self.hd), dim=-1) @ vcache[:, :pos + 1]
        b["att"].copy_(att.reshape(1, -1))
        lin("o", b["att"], b["o"])
        h = x + b["o"]
        b["h2"].copy_(self.rmsnorm(torch, h, self.norm2))
        lin("gate", b["h2"], b["gate"]); lin("up", b["h2"], b["up"])
        b["act"].copy_(torch.nn.functional.silu(b["gate"]) * b["up"])
        lin("down", b["act"], b["down"])
        return h + b["down"]

    def free(self):
        for l in self.lin.values():
            l.free()


class DecodeModel:
    """`layers` decoder layers of one shape (synthetic; layers > 0 share one host tensor set but own their HBM copies, so every
    weight byte of a token streams from HBM), a KV cache of `ctx` positions, one token step per call."""

    def __init__(self, layers: int, hidden: int, ffn: int, heads: int, bits: int, zero_point: bool, ctx: int = 512, seed: int = 0):
        import torch
        self.torch, self.hidden, self.ctx, self.heads, self.hd = torch, hidden, ctx, heads, hidden // heads
        first = DecodeLayer(hidden, ffn, heads, bits, zero_point, seed)
        self.layers: List[DecodeLayer] = [first] + [DecodeLayer(hidden, ffn, heads, bits, zero_point, seed, share=first) for _ in range(layers - 1)]
        self.kc = [torch.randn((heads, ctx, self.hd), device="cuda") * 0.1 for _ in range(layers)]
        self.vc = [torch.randn((heads, ctx, self.hd), device="cuda") * 0.1 for _ in range(layers)]
        inv = 1.0 / (10000.0 ** (torch.arange(0, self.hd, 2, device="cuda").float() / self.hd))
        self.pos = ctx - 1                          # decode at the last cache position: attention over the whole context
        ang = self.pos * inv
        self.cos, self.sin = torch.cos(ang)[None, :], torch.sin(ang)[None, :]
        self.x = torch.randn((1, hidden), device="cuda")
        self.y = torch.zeros((1, hidden), device="cuda")
        self.graph = None

    def step(self):
        h = self.x
        for i, layer in enumerate(self.layers):
            h = layer.forward(h, self.kc[i], self.vc[i], self.pos, self.cos, self.sin)
        self.y.copy_(h)

    def capture(self, stream):
        torch = self.torch
        with torch.cuda.stream(stream):
            self.step()                             # eager warm-up: library workspaces, torch allocator
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=stream):
            self.step()
        return self

    def tokens_per_s(self, stream, n: int = 10) -> float:
        torch = self.torch
        run = self.graph.replay if self.graph is not None else self.step
        with torch.cuda.stream(stream):
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                run()
            e1.record(stream)
        torch.cuda.synchronize()
        return n / (e0.elapsed_time(e1) * 1e-3)

    def free(self):
        self.graph = None
        for l in self.layers:
            l.free()
