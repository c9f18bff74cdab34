// tmac_gemv4.cuh -- the lone-launch GEMV (N = 1, one tensor per launch): a "stream-K" grid with exactly one CTA
// per SM per launch (two resident per SM, so that the next launch of a PDL chain co-resides and streams its
// weights while this one computes).
//
// Why (profiles/r1_trace_notes.md §3): a chain of dependent GEMVs (real decode: layer k+1 needs layer k) is
// latency bound -- the math of launch k+1 cannot start before launch k has completed.  gemv3's clusters were placed
// by the block scheduler wherever slots freed up, so the busiest SM held 2-3x the mean work and set the time.  Here
// the work split is arithmetic, not scheduler policy:
//
//   * blocks (row super-block, K chunk) are numbered super-block-major; CTA i takes [T*i/G, T*(i+1)/G) -- every
//     SM gets T/G +- 1 blocks whatever the shape (11008x4096 W2: 18 or 19 of 2752);
//   * the CTA's whole share (<= ~100 KB) is requested into shared memory by one `cp.async.bulk` per block,
//     completing on a warp-private mbarrier, BEFORE `griddepcontrol.wait`: the HBM stream of launch k+1 runs under
//     the math of launch k;
//   * warp w computes blocks w, w+16, ... of the share (same PRMT + DP4A body and the same fused LUT construction
//     as gemv3), partial sums per row super-block go through shared memory in fixed warp order;
//   * a super-block cut by a CTA boundary is finished by the LAST CTA that touches it: the earlier ones publish
//     their 32*RW partial sums through an 8-byte {value, flag} exchange slot in global memory (one store, one
//     polling load per row -- no fence, no atomic), summed in ascending CTA order, so results are deterministic.
//     The consumer clears the flag; the next launch's producers run after this grid has completed (stream order or
//     griddepcontrol.wait), so a slot is never rewritten before it was consumed.  Producers publish before they
//     consume, and consumers only wait on lower-numbered CTAs of a grid that is fully resident (G <= 2 * SMs).
#pragma once
#include "tmac_kernels.cuh"

namespace tmac_b200 {

constexpr int kG4Warps = 16;
constexpr int kG4MaxRsb = 256;            // rows per super-block, PB = 1

struct Gemv4Params {
    const unsigned char *W;               // first block of the launch's first row super-block
    const unsigned char *Wnext;           // tensor used next (L2 prefetch) or null
    const int8_t *qlut;                   // [K/4][16]
    const float *lut_scales, *lut_biases; // [K/ags]
    void *C;
    const void *act;                      // FUSED: activation row (f32 / f16)
    int act_f16;
    int K, row_begin, row_end, c_row0;
    int nrsb, rsb0, nchunk, ags;
    int zp, one_scale, sd, out_f16, blk_bytes;
    int total;                            // nrsb * nchunk
    int per_max, nseg_max;                // max blocks / row super-blocks per CTA
    int stage_bytes;                      // bytes reserved for the block stages (multiple of 128)
    int ntab;                             // LUT slices kept in shared memory = min(per_max, nchunk)
    size_t rsb_stride;
    float scale0;
    uint2 *xchg;                          // [grid][kG4MaxRsb] exchange slots {bits, flag}
    long long *trace;
};

__device__ __forceinline__ uint32_t g4_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g4_mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(g4_s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void g4_publish(uint2 *slot, uint32_t bits) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(bits), "r"(1u) : "memory");
}
__device__ __forceinline__ uint32_t g4_consume(uint2 *slot) {
    uint32_t v, f;
    int spins = 0;
    do {   // bounded (~1 s): a lost partner yields a wrong result that the caller's checks catch, never a hung GPU
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(slot) : "memory");
    } while (f == 0u && ++spins < (1 << 21));
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %1};" ::"l"(slot), "r"(0u) : "memory");
    return v;
}

template <int PB, bool SYM, int QCH, int AGQ, bool FUSED>
__global__ void __launch_bounds__(kG4Warps * 32, 2) gemv4_kernel(const Gemv4Params p, const uint32_t wtx, const uint32_t wty) {
    constexpr int RW = 8 / PB;
    constexpr int RSB = 32 * RW;
    constexpr int TB = SYM ? 8 : 16;
    constexpr bool INT_PATH = (AGQ == 0);
    constexpr int NAG = INT_PATH ? 1 : QCH / AGQ;
    constexpr int NW = kG4Warps;
    constexpr int TAB = QCH * 4 * TB;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int b0 = (int)(((long long)p.total * cta) / G), b1 = (int)(((long long)p.total * (cta + 1)) / G);
    const int nb = b1 - b0;
    const int sb_first = b0 / p.nchunk, sb_last = (b1 - 1) / p.nchunk;
    const int nseg = sb_last - sb_first + 1;
    // shared memory: stage [per_max][blk_bytes] | red [nseg_max][NW][RSB] f32 | tabs [ntab][TAB] | mbar [NW] | lsb [ntab][NAG+1] f32
    // tabs / lsb hold the LUT slice (table, LUT scales, LUT-bias sum) of every K chunk this CTA touches, slot = (c - c0) mod nchunk.
    unsigned char *stage0 = smem;
    float *red = reinterpret_cast<float *>(smem + p.stage_bytes);
    unsigned char *tabs = smem + p.stage_bytes + (size_t)p.nseg_max * NW * RSB * 4;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(tabs + (size_t)p.ntab * TAB) + warp;
    float *lsb = reinterpret_cast<float *>(tabs + (size_t)p.ntab * TAB + NW * 8);
    const int c0 = b0 % p.nchunk, nck = min(nb, p.nchunk);

    if (tid == 0) { TMAC_TRACE(0); }
    pdl_launch_dependents();                      // the next launch may take the other slot of this SM and stream its weights

    // ---- my share: HBM -> shared, requested before the dependency wait ---------------------------
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g4_s32(mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        int cnt = 0;
        for (int j = warp; j < nb; j += NW) ++cnt;
        if (cnt) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g4_s32(mbar)), "r"((uint32_t)(cnt * p.blk_bytes)) : "memory");
            for (int j = warp; j < nb; j += NW) {
                const int b = b0 + j, sb = b / p.nchunk, c = b - sb * p.nchunk;
                const size_t off = (size_t)sb * p.rsb_stride + (size_t)c * p.blk_bytes;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(g4_s32(stage0 + (size_t)j * p.blk_bytes)), "l"(p.W + off), "r"((uint32_t)p.blk_bytes), "r"(g4_s32(mbar)) : "memory");
                if (p.Wnext) l2_prefetch_bulk(p.Wnext + off, (uint32_t)p.blk_bytes);
            }
        }
    }
    // zero my rows of the reduction buffer (a warp only overwrites the super-blocks it touches)
    for (int s = 0; s < nseg; ++s) {
        float *r = red + ((size_t)s * NW + warp) * RSB + lane * RW;
#pragma unroll
        for (int i = 0; i < RW; ++i) r[i] = 0.f;
    }
    __syncwarp();
    if (tid == 0) TMAC_TRACE(1);
    pdl_wait();                                   // activations / LUT come from the previous kernel
    if (tid == 0) TMAC_TRACE(2);

    // ---- LUT slice of every chunk of my share -> shared memory, once (warp = chunk, lane = K-group) ----------
    const uint4 *qrow = reinterpret_cast<const uint4 *>(p.qlut);
    for (int ci = warp; ci < nck; ci += NW) {
        int c = c0 + ci; if (c >= p.nchunk) c -= p.nchunk;
        unsigned char *tab = tabs + (size_t)ci * TAB;
        float *lsd = lsb + (size_t)ci * (NAG + 1);
        if (FUSED) {
            // same operation order as preprocessor_kernel / gemv3 (lut_ctor.cc:119-215, :242-256): bit-identical LUT
            constexpr int NG = QCH * 4, W = (AGQ ? AGQ : 1) * 4;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
            if (lane < NG) {
                const size_t k0 = ((size_t)c * NG + lane) * 4;
                if (p.act_f16) {
                    const uint2 h = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(p.act) + k0));
                    const float2 f01 = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), f23 = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
                    x0 = f01.x; x1 = f01.y; x2 = f23.x; x3 = f23.y;
                } else {
                    const float4 f = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.act) + k0));
                    x0 = f.x; x1 = f.y; x2 = f.z; x3 = f.w;
                }
            }
            float m = __fadd_rn(__fadd_rn(fabsf(x0), fabsf(x1)), __fadd_rn(fabsf(x2), fabsf(x3)));
#pragma unroll
            for (int o = W / 2; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float scale = __fdiv_rn(m, 127.0f);
            const float ts = (scale != 0.0f) ? __fdiv_rn(1.0f, scale) : 0.0f;
            float od[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int idx = 2 * e + 1;
                float v = x0;
                v = (idx & 2) ? __fadd_rn(v, x1) : __fsub_rn(v, x1);
                v = (idx & 4) ? __fadd_rn(v, x2) : __fsub_rn(v, x2);
                v = (idx & 8) ? __fadd_rn(v, x3) : __fsub_rn(v, x3);
                od[e] = v;
            }
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float lv = (e & 1) ? od[e >> 1] : -od[(15 - e) >> 1];
                int q = __float2int_rn(__fmul_rn(lv, ts));
                q = max(-128, min(127, q));
                if (e < 4) lo |= (uint32_t)(q & 0xff) << (8 * e); else hi |= (uint32_t)(q & 0xff) << (8 * (e - 4));
            }
            if (lane < NG) reinterpret_cast<uint2 *>(tab)[lane] = make_uint2(lo, hi);
            float v = -od[7];
            v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
            v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
            v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
            float lbsum = 0.f;
#pragma unroll
            for (int a = 0; a < NAG; ++a) {
                const float sa = __shfl_sync(0xffffffffu, scale, a * W);
                float bias = 0.f;
#pragma unroll
                for (int k = 0; k < W / 8; ++k) bias = __fadd_rn(bias, __shfl_sync(0xffffffffu, v, a * W + 8 * k));
                lbsum += bias;
                if (lane == 0) lsd[a] = sa;
            }
            if (lane == 0) lsd[NAG] = lbsum;
        } else {
            if (lane < QCH * 4) {
                const uint4 L = __ldg(qrow + (size_t)c * QCH * 4 + lane);
                if (SYM) reinterpret_cast<uint2 *>(tab)[lane] = make_uint2(L.x, L.y);
                else reinterpret_cast<uint4 *>(tab)[lane] = make_uint4(L.x, L.y, __byte_perm(L.w, 0, 0x0123), __byte_perm(L.z, 0, 0x0123));
            }
            if (!INT_PATH && lane <= NAG) {
                float v = 0.f;
                if (lane < NAG) v = __ldg(p.lut_scales + c * NAG + lane);
                else {
#pragma unroll
                    for (int a = 0; a < NAG; ++a) v += __ldg(p.lut_biases + c * NAG + a);
                }
                lsd[lane] = v;
            }
        }
    }
    __syncthreads();
    if (tid == 0) TMAC_TRACE(3);

    float cacc[RW];
    int iacc[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) { cacc[i] = 0.f; iacc[i] = 0; }
    int cur_sb = -1;
    bool first = true;

    for (int j = warp; j < nb; j += NW) {
        const int b = b0 + j, sb = b / p.nchunk, c = b - sb * p.nchunk;
        if (sb != cur_sb) {
            if (cur_sb >= 0) {                    // flush the finished super-block of this warp
                float *r = red + ((size_t)(cur_sb - sb_first) * NW + warp) * RSB + lane * RW;
#pragma unroll
                for (int i = 0; i < RW; ++i) { r[i] = INT_PATH ? __int_as_float(iacc[i]) : cacc[i]; cacc[i] = 0.f; iacc[i] = 0; }
            }
            cur_sb = sb;
        }
        int ci = c - c0; if (ci < 0) ci += p.nchunk;
        const unsigned char *tab = tabs + (size_t)ci * TAB;
        float lsv[NAG], lbsum = 0.f;
        if (!INT_PATH) {
#pragma unroll
            for (int a = 0; a < NAG; ++a) lsv[a] = lsb[(size_t)ci * (NAG + 1) + a];
            lbsum = lsb[(size_t)ci * (NAG + 1) + NAG];
        }
        if (first) { g4_mbar_wait(mbar, 0); first = false; }   // all blocks of this warp have landed
        __syncwarp();
        const unsigned char *stage = stage0 + (size_t)j * p.blk_bytes;
        const uint4 *wp = reinterpret_cast<const uint4 *>(stage) + lane;
        float facc[RW];
#pragma unroll
        for (int i = 0; i < RW; ++i) facc[i] = 0.f;
#pragma unroll
        for (int q = 0; q < QCH; ++q) {
            const uint4 wq = wp[q * 32];
            uint32_t t[SYM ? 8 : 16];
            if (SYM) {
                const uint4 a = reinterpret_cast<const uint4 *>(tab)[2 * q], b2 = reinterpret_cast<const uint4 *>(tab)[2 * q + 1];
                t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b2.x; t[5] = b2.y; t[6] = b2.z; t[7] = b2.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 a = reinterpret_cast<const uint4 *>(tab)[4 * q + k];
                    t[4 * k] = a.x; t[4 * k + 1] = a.y; t[4 * k + 2] = a.z; t[4 * k + 3] = a.w;
                }
            }
            Quad<PB, SYM>::run(wq, t, iacc, wtx, wty);
            if (!INT_PATH && ((q + 1) % (AGQ ? AGQ : 1)) == 0) {
#pragma unroll
                for (int i = 0; i < RW; ++i) { facc[i] = fmaf(lsv[q / (AGQ ? AGQ : 1)], (float)iacc[i], facc[i]); iacc[i] = 0; }
            }
        }
        if (!INT_PATH) {
            const unsigned char *sp = stage + (size_t)QCH * 512;
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                const float s = p.one_scale ? p.scale0 : load_scale(sp, p.sd, lane * RW + i);
                float v = fmaf(0.5f * s, facc[i] + lbsum, cacc[i]);
                if (p.zp) v = fmaf(load_scale(sp + (size_t)RSB * p.sd, p.sd, lane * RW + i), lbsum, v);
                cacc[i] = v;
            }
        }
    }
    if (cur_sb >= 0) {
        float *r = red + ((size_t)(cur_sb - sb_first) * NW + warp) * RSB + lane * RW;
#pragma unroll
        for (int i = 0; i < RW; ++i) r[i] = INT_PATH ? __int_as_float(iacc[i]) : cacc[i];
    }
    if (tid == 0) TMAC_TRACE(4);
    __syncthreads();
    if (tid == 0) TMAC_TRACE(5);
    if (tid >= RSB) return;

    // ---- per row super-block: sum the warps (fixed order), then publish or finish --------------------
    const int t = tid;
    uint2 *my_slot = p.xchg + (size_t)cta * kG4MaxRsb + t;
    // pass 0: the super-block that continues in the next CTA (always my last one) is published first, so that no
    // publish ever waits behind a consume; pass 1: every super-block that ends here is finished.
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int s = (pass == 0 ? nseg - 1 : 0); s < nseg; ++s) {
            const int sb = sb_first + s;
            const bool ends_here = (long long)(sb + 1) * p.nchunk <= (long long)b1;
            if (pass == 0 ? ends_here : !ends_here) continue;
            float fsum = 0.f; int isum = 0;
            if (pass == 1 && (long long)sb * p.nchunk < (long long)b0) {
                // started in an earlier CTA: add the published partial sums in ascending CTA (= K) order
                const int bs = sb * p.nchunk;
                int fc = (int)(((long long)bs * G) / p.total);
                while ((int)(((long long)p.total * (fc + 1)) / G) <= bs) ++fc;
                while ((int)(((long long)p.total * fc) / G) > bs) --fc;
                for (int k2 = fc; k2 < cta; ++k2) {
                    const uint32_t v = g4_consume(p.xchg + (size_t)k2 * kG4MaxRsb + t);
                    if (INT_PATH) isum += (int)v; else fsum += __uint_as_float(v);
                }
            }
            for (int w = 0; w < NW; ++w) {
                const float v = red[((size_t)s * NW + w) * RSB + t];
                if (INT_PATH) isum += __float_as_int(v); else fsum += v;
            }
            if (pass == 0) {
                g4_publish(my_slot, INT_PATH ? (uint32_t)isum : __float_as_uint(fsum));
                continue;
            }
            const int row = (p.rsb0 + sb) * RSB + t;
            if (row >= p.row_begin && row < p.row_end) {
                float out;
                if (INT_PATH) {
                    // C = ((sum_b alpha_b*CBits_b) * LUT_Scales[0] + LUT_Biases[0]*alpha_0) * Scales[0] (qgemm.py:160,171-174)
                    const float cb = __fmul_rn((float)isum, 0.5f);
                    const float t1 = __fmul_rn(cb, __ldg(p.lut_scales));
                    const float t2 = __fmul_rn(__ldg(p.lut_biases), 0.5f);
                    out = __fmul_rn(__fadd_rn(t1, t2), p.scale0);
                } else
                    out = fsum;
                const size_t o = (size_t)(row - p.c_row0);
                if (p.out_f16) reinterpret_cast<__half *>(p.C)[o] = __float2half_rn(out);
                else reinterpret_cast<float *>(p.C)[o] = out;
            }
        }
        if (pass == 0 && tid == 0) TMAC_TRACE(6);
    }
    if (tid == 0) TMAC_TRACE(7);
}

typedef void (*gemv4_fn)(const Gemv4Params, const uint32_t, const uint32_t);
// Defined in tmac_gemv4.cu (its own translation unit, compiled in parallel with tmac_b200.cu); nullptr = not instantiated.
gemv4_fn pick_gemv4(int pb, bool sym, int qch, int agq, bool fused);

}  // namespace tmac_b200
