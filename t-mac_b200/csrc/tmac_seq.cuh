// tmac_seq.cuh -- the decode SEQUENCE kernel: a whole chain of dependent table-lookup GEMVs (the quantised linears of
// a token step, in model order) executed by ONE persistent launch, one CTA per SM.
//
// Why (profiles/r1_trace_notes.md §5): a chain of per-GEMV launches loses, per launch, the time
// for griddepcontrol.wait to return (0.77 us), the exposed weight stream (one launch fills the register file, so the
// next one cannot co-reside and prefetch), a cluster barrier on the slowest CTA and the final store -- 5.7 us for a
// 12.7 MB GEMV whose HBM time is 1.9 us.  The reference has the same structure on the CPU: ggml's graph loop runs one
// mul_mat node after the other on a persistent thread pool (3rdparty/llama.cpp/ggml/src/ggml.c:12562-12706 per node,
// workers parked between nodes, never re-created).  This kernel is the B200 form of that loop:
//
//   * the weight stream never stops: one producer thread per CTA walks the CTA's blocks of op 0, 1, 2, ... and
//     requests each (4.6 KB, contiguous) with one cp.async.bulk (TMA) into a ring of ~40 shared-memory slots; a
//     slot is reused when every consumer warp has published a progress counter beyond it.  HBM therefore runs
//     ~2 GEMVs ahead of the arithmetic, across the data dependencies, because weights do not depend on activations;
//   * 20 consumer warps per CTA do, per op: (A) build the LUT slices of the CTA's K chunks from the op's input
//     vector (same fp32 operation order as preprocessor_kernel / lut_ctor.cc -> bit-identical tables), (B) the
//     PRMT + DP4A lookups (Quad<>::run, as gemv3) over an arithmetic "stream-K" share of the op's blocks -- CTA c owns
//     blocks [T*c/G, T*(c+1)/G), a warp owns a contiguous run of activation-group units inside them --, (C) reduce
//     the warps through shared memory in fixed order and finish the rows;
//   * rows whose K range is split between CTAs are finished by the LAST CTA that touches them: the earlier ones
//     publish their partial sums through 8-byte {value, epoch} slots in global memory (one store, one polling load
//     per row; no fence, no atomic; ascending CTA order -> deterministic);
//   * the data dependency between ops is carried the same way: the finished row is stored as {value, epoch} into the
//     op's output vector, and the LUT build of a consumer op polls exactly the elements it needs.  An op can also
//     take an external (plain) vector as input and can write a plain copy of its output (C) for the caller.
//   Every wait is bounded (~1 s) and raises an error flag instead of hanging the GPU.
#pragma once
#include "tmac_kernels.cuh"

namespace tmac_b200 {

constexpr int kSeqWarps = 20;                         // consumer warps per CTA
constexpr int kSeqThreads = (kSeqWarps + 1) * 32;     // + 1 producer warp
constexpr int kSeqMaxRsb = 256;                       // rows per super-block, PB = 1
constexpr int kSeqSpinLimit = 1 << 20;           // x ~0.3 us per poll: every wait gives up after ~0.3 s

struct SeqOp {                            // one GEMV; read-only for the kernel
    const unsigned char *W;               // stream layout of the tensor (tmac_layout.h)
    const float *x_ext;                   // input: external fp32 vector [K] ...
    const uint2 *x_ll;                    // ... or {value, epoch} elements of an earlier op's output (already offset)
    void *C;                              // optional plain output [Mout] (f32 / f16)
    uint2 *y;                             // this op's output vector as {value, epoch} [nrsb * RSB]
    uint2 *xchg;                          // [grid][RSB] partial-sum exchange slots of this op
    unsigned long long rsb_stride;
    int K, Mout, nrsb, nchunk;
    int blk_bytes, total;                 // bytes per block; nrsb * nchunk
    int zp, one_scale, sd, out_f16;
    float scale0;
    int geff;                             // CTAs that take part in this op = min(grid, total): every one of them owns >= 1 block
};

struct SeqParams {
    const SeqOp *ops;
    int nops;
    int nslots, slot_bytes;               // weight ring
    int red_off, tab_off, lsb_off, bar_off, prog_off;   // shared-memory offsets (bytes)
    unsigned int *epochs;                 // [grid] launch counter of every CTA (incremented by that CTA at exit)
    int *err;                             // != 0: a wait expired
    long long *trace;                     // optional [nops][grid][8] globaltimer stamps
};

__device__ __forceinline__ long long seq_timer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void seq_publish(uint2 *slot, uint32_t bits, uint32_t epoch) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(bits), "r"(epoch) : "memory");
}
__device__ __forceinline__ uint32_t seq_consume(const uint2 *slot, uint32_t epoch, int *err) {
    uint32_t v, f;
    int spins = 0;
    do {
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(slot) : "memory");
    } while (f != epoch && ++spins < kSeqSpinLimit);
    if (f != epoch) { atomicExch(err, 2); v = 0; }
    return v;
}
// two adjacent {value, epoch} elements with one 16-byte load (each 8-byte element is written by one store)
__device__ __forceinline__ void seq_consume2(const uint2 *slot, uint32_t epoch, int *err, float &a, float &b) {
    uint32_t v0, f0, v1, f1;
    int spins = 0;
    do {
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(f0), "=r"(v1), "=r"(f1) : "l"(slot) : "memory");
    } while ((f0 != epoch || f1 != epoch) && ++spins < kSeqSpinLimit);
    if (f0 != epoch || f1 != epoch) { atomicExch(err, 3); v0 = v1 = 0; }
    a = __uint_as_float(v0); b = __uint_as_float(v1);
}
__device__ __forceinline__ void seq_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kSeqWarps * 32) : "memory"); }
__device__ __forceinline__ bool seq_mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    for (int spins = 0; spins < kSeqSpinLimit && !ok; ++spins)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ int ld_volatile_s32(const int *p) { int v; asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void st_volatile_s32(int *p, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }

// PB planes per word, QCH quads per chunk, AGQ quads per activation group (fp path: AGQ > 0).
template <int PB, int QCH, int AGQ>
__global__ void __launch_bounds__(kSeqThreads, 1) seq_kernel(const SeqParams p, const uint32_t wtx, const uint32_t wty) {
    constexpr int RW = 8 / PB, RSB = 32 * RW;
    constexpr int NAG = QCH / AGQ;                 // activation groups per chunk = work units per block
    constexpr int NG = QCH * 4;                    // K-groups per chunk
    constexpr int WL = AGQ * 4;                    // lanes (K-groups) per activation group
    constexpr int TAB = NG * 8;                    // table bytes per chunk (8 stored entries per group)
    constexpr int NW = kSeqWarps;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    unsigned char *ring = smem;
    float *red = reinterpret_cast<float *>(smem + p.red_off);          // [nseg][NW][RSB]
    unsigned char *tabs = smem + p.tab_off;                            // [ntab][TAB]
    float *lsb = reinterpret_cast<float *>(smem + p.lsb_off);          // [ntab][2 * NAG]: LUT scale, LUT bias per activation group
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.bar_off);   // [nslots]
    int *prog = reinterpret_cast<int *>(smem + p.prog_off);            // [NW] lowest block sequence number the warp still needs
    int *issued = prog + NW;                                           // blocks requested so far by the producer
    const uint32_t epoch = p.epochs[cta] + 1u;

    if (tid == 0) {
        for (int s = 0; s < p.nslots; ++s) mbar_init1(full + s);
        mbar_fence_init();
    }
    if (tid <= NW) prog[tid] = 0;                  // prog[NW] = issued
    __syncthreads();

    // =========================================== producer warp ===========================================
    if (warp == NW) {
        if (lane != 0) return;
        int seq = 0, slot = 0, known_min = 0;
        for (int op = 0; op < p.nops; ++op) {
            const SeqOp &o = p.ops[op];
            const int total = o.total, nchunk = o.nchunk, blk = o.blk_bytes;
            const unsigned char *W = o.W;
            const unsigned long long rs = o.rsb_stride;
            const int GE = o.geff;
            if (cta >= GE) continue;
            const int b0 = (int)(((long long)total * cta) / GE), b1 = (int)(((long long)total * (cta + 1)) / GE);
            int sb = b0 / nchunk, c = b0 - sb * nchunk;
            for (int b = b0; b < b1; ++b, ++seq) {
                if (seq >= p.nslots) {                         // slot still holds block seq - nslots: wait until nobody needs it
                    const int need = seq - p.nslots + 1;
                    int spins = 0;
                    while (known_min < need) {
                        int m = 0x7fffffff;
                        for (int w = 0; w < NW; ++w) m = min(m, ld_volatile_s32(prog + w));
                        known_min = m;
                        if (++spins > kSeqSpinLimit) { atomicExch(p.err, 1); return; }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_load(ring + (size_t)slot * p.slot_bytes, W + (size_t)sb * rs + (size_t)c * blk, (uint32_t)blk, full + slot);
                st_volatile_s32(issued, seq + 1);              // the phase parity of `full` is only meaningful once the request is out
                if (++slot == p.nslots) slot = 0;
                if (++c == nchunk) { c = 0; ++sb; }
            }
        }
        return;
    }

    // =========================================== consumer warps ===========================================
    int seq_base = 0, slot_base = 0, par_base = 0;     // sequence number / ring slot / phase parity of the op's first block
    for (int op = 0; op < p.nops; ++op) {
        const SeqOp &o = p.ops[op];
        const int total = o.total, nchunk = o.nchunk, GE = o.geff;
        if (cta >= GE) continue;                           // more CTAs than blocks: this CTA sits the op out
        const int b0 = (int)(((long long)total * cta) / GE), b1 = (int)(((long long)total * (cta + 1)) / GE);
        const int nb = b1 - b0;
        const int sb_first = b0 / nchunk, c0 = b0 - sb_first * nchunk;
        const int nseg = nb > 0 ? (b1 - 1) / nchunk - sb_first + 1 : 0;
        const int nck = min(nb, nchunk);
        // Work split inside the CTA: the share is walked in passes of at most half a ring (so that the blocks the warps
        // work on are resident together); inside a pass every warp owns a contiguous run of units.
        const int pmax = max(1, p.nslots / 2);
        const int npass = (nb + pmax - 1) / pmax, P = (nb + npass - 1) / npass;
        if (lane == 0) st_volatile_s32(prog + warp, seq_base);
        long long *tr = p.trace ? p.trace + ((size_t)op * G + cta) * 8 : nullptr;
        if (tr && tid == 0) tr[0] = seq_timer();

        // ---- (A) LUT slices of my chunks: thread = K-group.  Arithmetic = lut_ctor.cc:119-215 (AVX2 branch) and
        //      partial_max_g4_int8_k8 (:242-256) with explicit round-to-nearest operations (as preprocessor_kernel). ----
        {
            const float *xe = o.x_ext;
            const uint2 *xl = o.x_ll;
            const int ngroups = nck * NG;
            for (int t0 = warp * 32; t0 < ngroups; t0 += NW * 32) {
                const int t = t0 + lane;
                const bool valid = t < ngroups;
                const int ci = valid ? t / NG : 0, gl = t % NG;
                int c = c0 + ci; if (c >= nchunk) c -= nchunk;
                float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
                if (valid) {
                    const size_t k0 = ((size_t)c * NG + gl) * 4;
                    if (xe) {
                        const float4 f = *reinterpret_cast<const float4 *>(xe + k0);
                        x0 = f.x; x1 = f.y; x2 = f.z; x3 = f.w;
                    } else {
                        seq_consume2(xl + k0, epoch, p.err, x0, x1);
                        seq_consume2(xl + k0 + 2, epoch, p.err, x2, x3);
                    }
                }
                float m = __fadd_rn(__fadd_rn(fabsf(x0), fabsf(x1)), __fadd_rn(fabsf(x2), fabsf(x3)));
#pragma unroll
                for (int s = WL / 2; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
                const float scale = __fdiv_rn(m, 127.0f);
                const float ts = (scale != 0.0f) ? __fdiv_rn(1.0f, scale) : 0.0f;
                // odd entries 1, 3, ..., 15: ((x0 +- x1) +- x2) +- x3
                const float p01 = __fadd_rn(x0, x1), m01 = __fsub_rn(x0, x1);
                const float a0 = __fsub_rn(m01, x2), a1 = __fsub_rn(p01, x2), a2 = __fadd_rn(m01, x2), a3 = __fadd_rn(p01, x2);
                float od[8];
                od[0] = __fsub_rn(a0, x3); od[1] = __fsub_rn(a1, x3); od[2] = __fsub_rn(a2, x3); od[3] = __fsub_rn(a3, x3);
                od[4] = __fadd_rn(a0, x3); od[5] = __fadd_rn(a1, x3); od[6] = __fadd_rn(a2, x3); od[7] = __fadd_rn(a3, x3);
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {               // stored entry e: even = -LUT[15-e], odd = LUT[e]
                    const float lv = (e & 1) ? od[e >> 1] : -od[(15 - e) >> 1];
                    int q = __float2int_rn(__fmul_rn(lv, ts));
                    q = max(-128, min(127, q));
                    if (e < 4) lo |= (uint32_t)(q & 0xff) << (8 * e); else hi |= (uint32_t)(q & 0xff) << (8 * (e - 4));
                }
                if (valid) reinterpret_cast<uint2 *>(tabs + (size_t)ci * TAB)[gl] = make_uint2(lo, hi);
                // LUT bias: _mm256_addv_ps tree per 8 groups (lut_ctor.cc:24-31), serial over the blocks of a group (:157)
                float v = -od[7];
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
                float bias = 0.f;
#pragma unroll
                for (int k = 0; k < WL / 8; ++k) bias = __fadd_rn(bias, __shfl_sync(0xffffffffu, v, (lane & ~(WL - 1)) + 8 * k));
                if (valid && (gl % WL) == 0) {
                    float *d = lsb + (size_t)ci * (2 * NAG);
                    d[gl / WL] = scale;
                    d[NAG + gl / WL] = bias;
                }
            }
        }
        seq_bar(1);
        if (tr && tid == 0) tr[1] = seq_timer();

        // ---- (B) lookups over my units (unit = one activation group of one block) ----
        for (int s = 0; s < nseg; ++s) {                       // rows of super-blocks I do not touch must read as zero
            float *r = red + ((size_t)s * NW + warp) * RSB + lane * RW;
#pragma unroll
            for (int i = 0; i < RW; ++i) r[i] = 0.f;
        }
        {
            float cacc[RW];
            int iacc[RW];
#pragma unroll
            for (int i = 0; i < RW; ++i) { cacc[i] = 0.f; iacc[i] = 0; }
            int cur_seg = -1;
            for (int pass = 0; pass < npass; ++pass) {
            const int pb0 = pass * P, pb1 = min(nb, pb0 + P);
            const int nu = (pb1 - pb0) * NAG;
            const int u1 = pb0 * NAG + (int)(((long long)nu * (warp + 1)) / NW);
            int u = pb0 * NAG + (int)(((long long)nu * warp) / NW);
            int j = u / NAG;
            int sb = (b0 + j) / nchunk, c = (b0 + j) - sb * nchunk;
            int slot = slot_base + j, par = par_base;
            while (slot >= p.nslots) { slot -= p.nslots; par ^= 1; }
            if (u < u1 && lane == 0) st_volatile_s32(prog + warp, seq_base + j);
            while (u < u1) {
                const int a0 = u - j * NAG, a1 = min(NAG, a0 + (u1 - u));
                const int seg = sb - sb_first;
                if (seg != cur_seg) {
                    if (cur_seg >= 0) {
                        float *r = red + ((size_t)cur_seg * NW + warp) * RSB + lane * RW;
#pragma unroll
                        for (int i = 0; i < RW; ++i) { r[i] = cacc[i]; cacc[i] = 0.f; }
                    }
                    cur_seg = seg;
                }
                int ci = c - c0; if (ci < 0) ci += nchunk;
                const unsigned char *tab = tabs + (size_t)ci * TAB;
                const float *ls = lsb + (size_t)ci * (2 * NAG);
                {
                    int spins = 0;
                    while (ld_volatile_s32(issued) <= seq_base + j && ++spins < kSeqSpinLimit) { }
                    if (spins >= kSeqSpinLimit || !seq_mbar_wait(full + slot, (uint32_t)par)) atomicExch(p.err, 4);
                }
                const unsigned char *stage = ring + (size_t)slot * p.slot_bytes;
                const uint4 *wp = reinterpret_cast<const uint4 *>(stage) + lane;
                float facc[RW];
#pragma unroll
                for (int i = 0; i < RW; ++i) facc[i] = 0.f;
                float lbp = 0.f;
                for (int a = a0; a < a1; ++a) {
#pragma unroll
                    for (int qq = 0; qq < AGQ; ++qq) {
                        const int q = a * AGQ + qq;
                        const uint4 wq = wp[q * 32];
                        const uint4 ta = reinterpret_cast<const uint4 *>(tab)[2 * q], tb = reinterpret_cast<const uint4 *>(tab)[2 * q + 1];
                        const uint32_t t[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                        Quad<PB, true>::run(wq, t, iacc, wtx, wty);
                    }
                    const float lsv = ls[a];
#pragma unroll
                    for (int i = 0; i < RW; ++i) { facc[i] = fmaf(lsv, (float)iacc[i], facc[i]); iacc[i] = 0; }
                    lbp += ls[NAG + a];
                }
                {
                    const unsigned char *sp = stage + (size_t)QCH * 512;
#pragma unroll
                    for (int i = 0; i < RW; ++i) {
                        const float s = o.one_scale ? o.scale0 : load_scale(sp, o.sd, lane * RW + i);
                        float v = fmaf(0.5f * s, facc[i] + lbp, cacc[i]);
                        if (o.zp) v = fmaf(load_scale(sp + (size_t)RSB * o.sd, o.sd, lane * RW + i), lbp, v);
                        cacc[i] = v;
                    }
                }
                u += a1 - a0;
                if (a1 == NAG) {                                // on to the next block
                    ++j;
                    if (++c == nchunk) { c = 0; ++sb; }
                    if (++slot == p.nslots) { slot = 0; par ^= 1; }
                    if (u < u1 && lane == 0) st_volatile_s32(prog + warp, seq_base + j);
                }
            }
            __syncwarp();
            if (lane == 0) st_volatile_s32(prog + warp, seq_base + pb1);
            }
            if (cur_seg >= 0) {
                float *r = red + ((size_t)cur_seg * NW + warp) * RSB + lane * RW;
#pragma unroll
                for (int i = 0; i < RW; ++i) r[i] = cacc[i];
            }
        }
        if (tr && tid == 0) tr[2] = seq_timer();
        seq_bar(2);
        if (tr && tid == 0) tr[3] = seq_timer();

        // ---- (C) per (super-block, row): sum the warps in fixed order, then publish the partial sum or finish the row ----
        for (int item = tid; item < nseg * RSB; item += NW * 32) {
            const int s = item / RSB, t = item - s * RSB;
            const int sb = sb_first + s;
            const long long bs = (long long)sb * nchunk;
            const bool ends_here = bs + nchunk <= (long long)b1;
            float mine = 0.f;
#pragma unroll 4
            for (int w = 0; w < NW; ++w) mine += red[((size_t)s * NW + w) * RSB + t];
            if (!ends_here) {                                   // continues in the next CTA (always my last super-block)
                seq_publish(o.xchg + (size_t)cta * RSB + t, __float_as_uint(mine), epoch);
                continue;
            }
            float fsum = 0.f;
            if (bs < (long long)b0) {                           // started in earlier CTAs: their partial sums, ascending K order
                int fc = (int)((bs * GE) / total);
                while ((int)(((long long)total * (fc + 1)) / GE) <= bs) ++fc;
                while ((int)(((long long)total * fc) / GE) > bs) --fc;
                for (int k2 = fc; k2 < cta; ++k2) fsum += __uint_as_float(seq_consume(o.xchg + (size_t)k2 * RSB + t, epoch, p.err));
            }
            fsum += mine;
            const int row = sb * RSB + t;
            if (row < o.Mout) {
                if (o.C) {
                    if (o.out_f16) reinterpret_cast<__half *>(o.C)[row] = __float2half_rn(fsum);
                    else reinterpret_cast<float *>(o.C)[row] = fsum;
                }
                seq_publish(o.y + row, __float_as_uint(fsum), epoch);
            }
        }
        if (tr && tid == 0) tr[4] = seq_timer();
        seq_base += nb;
        slot_base += nb;
        while (slot_base >= p.nslots) { slot_base -= p.nslots; par_base ^= 1; }
    }
    if (tid == 0) p.epochs[cta] = epoch;
}

typedef void (*seq_fn)(const SeqParams, const uint32_t, const uint32_t);
// Defined in tmac_seq.cu (its own translation unit); nullptr = chunking not instantiated.
seq_fn pick_seq(int pb, int qch, int agq);

}  // namespace tmac_b200
