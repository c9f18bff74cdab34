// tmac_kernels.cuh -- hand-written sm_100a kernels of the T-MAC hot path.
//
//   preprocessor_kernel : activation row -> per-act-group LUT scale, LUT bias, int8 LUT
//                         (python/t_mac/intrins/lut_ctor.cc:38-260; loop nest
//                          deploy/tuned/kernels.cc:1002-1040).  Bit-exact with the x86 reference.
//   gemv3_kernel        : table-lookup GEMV over the stream layout (tmac_layout.h)
//                         (python/t_mac/intrins/tbl.cc:323-630 + generated recombine
//                          deploy/tuned/aarch64-llama-2-7b-2bit/kernels.cc:1059-1075).
//
// Lookup primitive.  The reference does 16 (NEON) / 32 (AVX2) byte lookups per `tbl`/`pshufb`
// in a 16-entry int8 table.  On B200 the register-file analogue is PRMT: 4 byte lookups in an
// 8-byte table held in two registers.  The LUT is odd-symmetric (LUT[15-i] = -LUT[i],
// lut_ctor.cc:153-155), so 8 entries + a sign bit are enough; the sign is applied by DP4A, whose
// second operand (+-plane weights 2*alpha_b = 1,2,4,8) is itself fetched with one PRMT from an
// 8-byte constant table indexed by the sign bits.  One DP4A therefore does "negate, scale by
// the bit-plane weight, add the 4 lookups" = the int16 adder tree + alpha recombination of the
// reference, exactly, in int32.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmac_b200 {

// (a & b) | c in ONE LOP3 (the compiler splits it in two when b and c are immediates)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
// x >> 16 on the FMA pipe (IMAD.HI) instead of the ALU pipe (SHF): the PRMT/LOP3 work of the lookup
// loop saturates the ALU pipe, the FMA pipe only carries the DP4As.
#ifdef TMAC_HI16_SHF
__device__ __forceinline__ uint32_t hi16(uint32_t x) { return x >> 16; }
#else
__device__ __forceinline__ uint32_t hi16(uint32_t x) { return __umulhi(x, 65536u); }
#endif
#ifndef TMAC_G3_MINB
#define TMAC_G3_MINB 4
#endif
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------
// process_quad: 4 K-groups x the lane's RW rows.  acc[r] += sum_{k<4} sum_b 2*alpha_b * sgn * T_k[j]
// ------------------------------------------------------------------------------------------
template <int PB, bool SYM> struct Quad;

// Sign handling of the symmetric (8 stored entries) path without a shift: the selector nibble of the sign PRMT is
// neg << 3 | byte position, i.e. the weight word ANDed in place.  PRMT's replicate mode (selector bit 3) turns the
// selected byte 2*w (msb clear) into 0x00, so prmt(2w, sel) = 2*w*(1 - neg); with the constant vector -w,
//   dp4a(v, 2w(1-neg)) + dp4a(v, -w) = sum v*w*(1 - 2*neg)
// exactly.  One LOP3 per word instead of SHF + LOP3 on the saturated ALU pipe; the extra DP4A rides the FMA pipe.
// The general (16 stored entries) path keeps neg << 2 | position: it selects between two looked-up candidates.
template <bool SYM> __device__ __forceinline__ uint32_t sign_sel(uint32_t w) {
    return SYM ? and_or(w, 0x88888888u, 0x32103210u) : and_or(w >> 1, 0x44444444u, 0x32103210u);
}
__device__ __forceinline__ int dp4a_signed(uint32_t v, uint32_t w2, uint32_t wneg, uint32_t sel, int acc) {
    return __dp4a((int)v, (int)prmt(w2, w2, sel), __dp4a((int)v, (int)wneg, acc));
}

// sign/plane-weight tables (bytes): index = plane-position | neg << 2
//   PB 4: {1,2,4,8 | -1,-2,-4,-8}   (bits 3: plane 3 weight 0)
//   PB 2: {1,2,1,2 | -1,-2,-1,-2}
//   PB 1: {1,1,1,1 | -1,-1,-1,-1}
template <bool SYM> struct Quad<4, SYM> {
    static __device__ __forceinline__ void run(const uint4 w, const uint32_t *tab /* 4 tables */, int *acc,
                                               uint32_t wtx, uint32_t wty) {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t wj = ww[k] & 0x77777777u;
            const uint32_t ws = sign_sel<SYM>(ww[k]);
            uint32_t v0, v1;
            if (SYM) {
                const uint32_t w2 = wtx << 1;
                v0 = prmt(tab[2 * k], tab[2 * k + 1], wj);
                v1 = prmt(tab[2 * k], tab[2 * k + 1], hi16(wj));
                acc[0] = dp4a_signed(v0, w2, wty, ws, acc[0]);
                acc[1] = dp4a_signed(v1, w2, wty, hi16(ws), acc[1]);
            } else {
                const uint32_t *g = tab + 4 * k;
                v0 = prmt(prmt(g[0], g[1], wj), prmt(g[2], g[3], wj), ws);
                v1 = prmt(prmt(g[0], g[1], hi16(wj)), prmt(g[2], g[3], hi16(wj)), hi16(ws));
                acc[0] = __dp4a((int)v0, (int)wtx, acc[0]);
                acc[1] = __dp4a((int)v1, (int)wtx, acc[1]);
            }
        }
    }
};

template <bool SYM> struct Quad<2, SYM> {
    static __device__ __forceinline__ void run(const uint4 w, const uint32_t *tab, int *acc, uint32_t wtx,
                                               uint32_t wty) {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            const uint32_t we = ww[2 * pr], wo = ww[2 * pr + 1];
            const uint32_t je = we & 0x77777777u, jo = wo & 0x77777777u;
            const uint32_t se = sign_sel<SYM>(we), so = sign_sel<SYM>(wo);
            if (SYM) {
                const uint32_t w2 = wtx << 1;
                const uint32_t *te = tab + 4 * pr, *to = tab + 4 * pr + 2;
                const uint32_t v0a = prmt(te[0], te[1], je), v0b = prmt(te[0], te[1], hi16(je));
                const uint32_t v1a = prmt(to[0], to[1], jo), v1b = prmt(to[0], to[1], hi16(jo));
                acc[0] = dp4a_signed(prmt(v0a, v1a, 0x5410), w2, wty, se, acc[0]);
                acc[1] = dp4a_signed(prmt(v0a, v1a, 0x7632), w2, wty, hi16(se), acc[1]);
                acc[2] = dp4a_signed(prmt(v0b, v1b, 0x5410), w2, wty, so, acc[2]);
                acc[3] = dp4a_signed(prmt(v0b, v1b, 0x7632), w2, wty, hi16(so), acc[3]);
            } else {
                const uint32_t *ge = tab + 8 * pr, *go = tab + 8 * pr + 4;
                const uint32_t l0a = prmt(ge[0], ge[1], je), l0b = prmt(ge[0], ge[1], hi16(je));
                const uint32_t h0a = prmt(ge[2], ge[3], je), h0b = prmt(ge[2], ge[3], hi16(je));
                const uint32_t l1a = prmt(go[0], go[1], jo), l1b = prmt(go[0], go[1], hi16(jo));
                const uint32_t h1a = prmt(go[2], go[3], jo), h1b = prmt(go[2], go[3], hi16(jo));
                // transpose lo and hi candidates, then pick by the (row-ordered) sign bits
                acc[0] = __dp4a((int)prmt(prmt(l0a, l1a, 0x5410), prmt(h0a, h1a, 0x5410), se), (int)wtx, acc[0]);
                acc[1] = __dp4a((int)prmt(prmt(l0a, l1a, 0x7632), prmt(h0a, h1a, 0x7632), hi16(se)), (int)wtx, acc[1]);
                acc[2] = __dp4a((int)prmt(prmt(l0b, l1b, 0x5410), prmt(h0b, h1b, 0x5410), so), (int)wtx, acc[2]);
                acc[3] = __dp4a((int)prmt(prmt(l0b, l1b, 0x7632), prmt(h0b, h1b, 0x7632), hi16(so)), (int)wtx, acc[3]);
            }
        }
    }
};

__device__ __forceinline__ void transpose4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t *x) {
    const uint32_t t01 = prmt(v0, v1, 0x5140), t23 = prmt(v2, v3, 0x5140);
    const uint32_t u01 = prmt(v0, v1, 0x7362), u23 = prmt(v2, v3, 0x7362);
    x[0] = prmt(t01, t23, 0x5410);
    x[1] = prmt(t01, t23, 0x7632);
    x[2] = prmt(u01, u23, 0x5410);
    x[3] = prmt(u01, u23, 0x7632);
}

template <bool SYM> struct Quad<1, SYM> {
    static __device__ __forceinline__ void run(const uint4 w, const uint32_t *tab, int *acc, uint32_t wtx,
                                               uint32_t wty) {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
        uint32_t s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = sign_sel<SYM>(ww[k]);
        uint32_t xa[4], xb[4];
        if (SYM) {
            uint32_t va[4], vb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t j = ww[k] & 0x77777777u;
                va[k] = prmt(tab[2 * k], tab[2 * k + 1], j);
                vb[k] = prmt(tab[2 * k], tab[2 * k + 1], hi16(j));
            }
            transpose4(va[0], va[1], va[2], va[3], xa);
            transpose4(vb[0], vb[1], vb[2], vb[3], xb);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t sa = (r & 1) ? (hi16(s[r >> 1])) : s[r >> 1];           // rows 0..3: words 0,0,1,1
                const uint32_t sb = (r & 1) ? (hi16(s[2 + (r >> 1)])) : s[2 + (r >> 1)]; // rows 4..7: words 2,2,3,3
                acc[r] = dp4a_signed(xa[r], wtx << 1, wty, sa, acc[r]);
                acc[4 + r] = dp4a_signed(xb[r], wtx << 1, wty, sb, acc[4 + r]);
            }
        } else {
            uint32_t la[4], lb[4], ha[4], hb[4], xla[4], xlb[4], xha[4], xhb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t j = ww[k] & 0x77777777u;
                const uint32_t *g = tab + 4 * k;
                la[k] = prmt(g[0], g[1], j); lb[k] = prmt(g[0], g[1], hi16(j));
                ha[k] = prmt(g[2], g[3], j); hb[k] = prmt(g[2], g[3], hi16(j));
            }
            transpose4(la[0], la[1], la[2], la[3], xla);
            transpose4(ha[0], ha[1], ha[2], ha[3], xha);
            transpose4(lb[0], lb[1], lb[2], lb[3], xlb);
            transpose4(hb[0], hb[1], hb[2], hb[3], xhb);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t sa = (r & 1) ? (hi16(s[r >> 1])) : s[r >> 1];
                const uint32_t sb = (r & 1) ? (hi16(s[2 + (r >> 1)])) : s[2 + (r >> 1)];
                acc[r] = __dp4a((int)prmt(xla[r], xha[r], sa), (int)wtx, acc[r]);
                acc[4 + r] = __dp4a((int)prmt(xlb[r], xhb[r], sb), (int)wtx, acc[4 + r]);
            }
        }
        (void)xa; (void)xb;
    }
};

// Host helper: the two constant registers handed to Quad<>::run.
//  SYM : (wtx, wty) = byte table {+w0,+w1,+w2,+w3 | -w0,-w1,-w2,-w3}
//  !SYM: wtx = constant plane weights (w0..w3), wty unused
__host__ __device__ inline void plane_weight_regs(int bits, bool sym, uint32_t *wtx, uint32_t *wty) {
    int w[4];
    if (bits >= 3) { w[0] = 1; w[1] = 2; w[2] = 4; w[3] = (bits == 4) ? 8 : 0; }
    else if (bits == 2) { w[0] = 1; w[1] = 2; w[2] = 1; w[3] = 2; }
    else { w[0] = w[1] = w[2] = w[3] = 1; }
    uint32_t p = 0, n = 0;
    for (int i = 0; i < 4; ++i) { p |= (uint32_t)(w[i] & 0xff) << (8 * i); n |= (uint32_t)((-w[i]) & 0xff) << (8 * i); }
    *wtx = p; *wty = n;
    (void)sym;
}

__device__ __forceinline__ float load_scale(const unsigned char *p, int sd, int i) {
    return sd == 2 ? __half2float(reinterpret_cast<const __half *>(p)[i]) : reinterpret_cast<const float *>(p)[i];
}

// ==========================================================================================
// gemv3_kernel -- the production GEMV.
//
//   decomposition : CTA = (row super-block, K slice); the K slices of one super-block form a
//                   thread-block CLUSTER (size CS <= 8) and are summed through distributed shared
//                   memory in rank order -- deterministic, no global scratch, no atomics.
//   warps         : warp-autonomous.  Warp w of CTA rank r owns the `bpw` consecutive K chunks
//                   starting at (r*WPC + w)*bpw.  Every lane owns RW rows (no cross-lane reduce).
//   memory        : each warp copies its (contiguous, 16-byte aligned) block HBM -> its private
//                   shared-memory stage with cp.async (LDGSTS, L2 evict-first) BEFORE
//                   griddepcontrol.wait: weights are static, so under programmatic dependent launch
//                   the stream of launch i+1 overlaps the math of launch i, and because the bytes
//                   wait in shared memory (not registers) both launches fit on an SM together.
//                   The activation LUT slice of the chunk (QCH*4 groups x 16 B) is fetched after the
//                   wait into a warp-private table and read back as broadcast LDS.128.
//                   One lane per warp can also issue cp.async.bulk.prefetch.L2 for the same block of
//                   the NEXT tensor (optional hint).
//   arithmetic    : Quad<PB,SYM> (PRMT lookups + DP4A), per-act-group fp32 scale, per-chunk weight
//                   scale / zero point, exactly the algebra of tbl.cc:435-529 re-associated.
//   template AGQ  : quads per activation group inside a chunk (2, 4, 8); 0 = integer path
//                   (one LUT scale for the whole row, python/t_mac/ops/qgemm.py:93-96).
// ==========================================================================================
struct Gemv3Params {
    const unsigned char *W;        // first block of the launch's first row super-block
    const unsigned char *Wnext;    // same, for the tensor that will be used next (L2 prefetch) or null
    const int8_t *qlut;            // [N][K/4][16]
    const float *lut_scales, *lut_biases;  // [N][K/ags]
    void *C;                       // [N][ldc]
    int K, ldc, row_begin, row_end, c_row0, bits;
    int nrsb, rsb0, nchunk;
    int ags;
    int zp, one_scale, sd, out_f16;
    int blk_bytes;                 // bytes per block (weights + scales)
    int cs, wpc, bpw;              // cluster size, warps per CTA, chunks per warp
    int nbuf;                      // stage buffers per warp (1 or 2)
    int pdl_late;                  // trigger dependents after the math instead of at entry
    // grouped launch: blockIdx.z selects one of `nbatch` problems with identical geometry
    int nbatch;
    const unsigned char *const *Wv;
    const int8_t *const *qlutv;
    const float *const *lsv, *const *lbv;
    void *const *Cv;
    // fused LUT construction (FUSED instantiations): activations [N][K] f32 or f16 instead of qlut / lut_scales / lut_biases
    const void *act;
    int act_f16;
    int ioff;                      // fused integer path (activation group = K): byte offset of the row-scan scratch in shared memory
    size_t rsb_stride;
    float scale0;
    long long *trace;
    // multi-GPU row sharding: the finished rows are also stored straight into the peers' output vectors (P2P over NVLink),
    // Cpeer[q] already offset to this shard's first row; N = 1 only
    int npeer;
    void *Cpeer[7];
};
#ifdef TMAC_ENABLE_TRACE
__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// slots 0..6: clock64 (SM cycles) since entry; slot 7: globaltimer (ns) at entry.
#define TMAC_TRACE(slot) do { if (p.trace) p.trace[(size_t)blockIdx.x * 8 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TMAC_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void l2_prefetch_bulk(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_plain(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Split barrier used to establish "every CTA of the cluster is running" before the first
// distributed-shared-memory access (required by the programming model), without stalling at entry.
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_f32(float *local_ptr, uint32_t rank, float v) {
    uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_ptr), raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

// One bulk (TMA) copy per block instead of a per-lane cp.async loop: 1 instruction from one lane instead of
// ~4 per 16 bytes on the ALU pipe that the lookup loop saturates; completion on a warp-private mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init1(uint64_t *b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

constexpr int kG3MaxWarps = 8;

// MINB = minimum resident CTAs per SM the register allocation is tuned for: 3 (85 registers, more ILP;
// best for a single launch per tensor) or 4 (64 registers, more CTAs in flight; best for grouped launches).
// FUSED = build the LUT slice of each chunk inside the GEMV from the activation row (same fp32 operation
// order as preprocessor_kernel / lut_ctor.cc, so QLUT bytes, LUT scales and LUT biases are bit-identical to
// the two-kernel path) instead of reading QLUT / LUT_Scales / LUT_Biases.  Requires SYM.  AGQ > 0: activation
// group inside a chunk, scale and bias per chunk by warp shuffles.  AGQ == 0 (integer path, ONE activation group
// = the whole row, BitNet): every CTA scans the row once (K * 4 bytes, L2-resident) for the row-wide abs-sum
// maximum -- cheaper than a cluster-wide exchange, which costs a rendezvous -- and the cluster leader also forms
// the row's LUT bias in the reference's order (addv tree per 32 activations, serial over the blocks).
template <int PB, bool SYM, int QCH, int AGQ, int MINB, bool FUSED = false>
__global__ void __launch_bounds__(kG3MaxWarps * 32, MINB) gemv3_kernel(const Gemv3Params p, const uint32_t wtx, const uint32_t wty) {
    constexpr int RW = 8 / PB;
    constexpr int RSB = 32 * RW;
    constexpr int TB = SYM ? 8 : 16;              // table bytes per group in shared memory
    constexpr bool INT_PATH = (AGQ == 0);
    constexpr int NAG = INT_PATH ? 1 : QCH / AGQ; // activation groups per chunk
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int WPC = p.wpc;
    const int rank = (p.cs > 1) ? (int)cluster_ctarank() : 0;
    const int rsb = blockIdx.x / p.cs, n = blockIdx.y;
    // shared memory: cl [CS][RSB] (cluster partials, leader only) | per warp: stage (blk_bytes) + table
    // (QCH*4*TB).  The CTA reduction buffer red [WPC][RSB] aliases the stages once they are consumed.
    float *cl = reinterpret_cast<float *>(smem);
    const int tab_bytes = QCH * 4 * TB;
    const int nbuf = p.nbuf;                      // 2 = double-buffered stage when a warp walks several chunks, 1 = single
    const int per_warp = nbuf * p.blk_bytes + tab_bytes;
    unsigned char *wbase = smem + (size_t)p.cs * RSB * 4;
    unsigned char *stage0 = wbase + (size_t)warp * per_warp;
    unsigned char *tab = stage0 + nbuf * p.blk_bytes;
    float *red = reinterpret_cast<float *>(wbase);
    // warp-private mbarriers (one per stage buffer) behind the region that red aliases
    const size_t wregion = max((size_t)WPC * per_warp, (size_t)WPC * RSB * 4);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbase + ((wregion + 15) & ~(size_t)15)) + warp * 2;

    if (tid == 0) { TMAC_TRACE(0); }
    if (p.cs > 1) cluster_arrive_relaxed();       // phase 1: "I am running" (waited for right before the DSMEM stores)
    if (!p.pdl_late) pdl_launch_dependents();     // the next launch may start its weight stream now

    const int c_first = (rank * WPC + warp) * p.bpw;
    const int c_end = min(p.nchunk, c_first + p.bpw);
    const unsigned char *Wb = p.W;
    const int8_t *qb = p.qlut;
    const float *lsb = p.lut_scales, *lbb = p.lut_biases;
    void *Cb = p.C;
    if (p.nbatch > 0) {
        const int z = blockIdx.z;
        Wb = p.Wv[z]; Cb = p.Cv[z];
        if (!FUSED) { qb = p.qlutv[z]; lsb = p.lsv[z]; lbb = p.lbv[z]; }   // fused: the tensors of the group share the activation rows
    }
    const unsigned char *rsb_base = Wb + (size_t)rsb * p.rsb_stride;
    const int nag = p.K / p.ags;

    // ---- first chunk: HBM -> shared, issued before the dependency wait -------------------------
    if (lane == 0) {
        mbar_init1(mbar); mbar_init1(mbar + 1);
        mbar_fence_init();
        if (c_first < c_end) {
            bulk_load(stage0, rsb_base + (size_t)c_first * p.blk_bytes, (uint32_t)p.blk_bytes, mbar);
            if (p.Wnext)                           // pull the next tensor's blocks of this warp into L2
                l2_prefetch_bulk(p.Wnext + (size_t)rsb * p.rsb_stride + (size_t)c_first * p.blk_bytes,
                                 (uint32_t)((c_end - c_first) * p.blk_bytes));
        }
    }
    __syncwarp();
    if (tid == 0) TMAC_TRACE(1);
    pdl_wait();                                   // LUT / LUT scales come from the previous kernel
    if (tid == 0) TMAC_TRACE(2);

    float cacc[RW];
    int iacc[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) { cacc[i] = 0.f; iacc[i] = 0; }
    float row_scale = 0.f, row_ts = 0.f;          // fused integer path: LUT scale of the row and its reciprocal
    float *iscr = reinterpret_cast<float *>(smem + p.ioff);   // [K/32] block sums (leader) | [8] warp maxima | [1] bias
    bool scan_split = false;                      // fused integer path: the cluster splits the row scan (else every CTA scans the row)
    if (FUSED && INT_PATH) {
        // Row-wide LUT scale and bias (lut_ctor.cc:232-260 partial_max over the whole row, :157 bias) without a preprocessor launch.
        // Short rows (one round of four 16-byte loads per thread covers the row, K = 3200): every CTA scans the whole row itself,
        // no exchange.  Long rows (K = 8640): that is 200 CTAs x 34 KB of reads on the same few L2 lines, several serial rounds;
        // the cluster splits the row instead, maxima go to every peer and the block sums of LUT[0] to the last rank through
        // distributed shared memory, one cluster barrier (measured: 4.55 / 5.31 us per 3200x3200 GEMV whole-row / split,
        // 10.7 / 9.8 us per 3200x8640).  The last rank's last warp (idle or lightest: warps beyond the last chunk have no block)
        // forms the bias serially; the leader needs it only in the epilogue.
        const int nblk = p.K >> 5, ngrp = p.K >> 2, nth = WPC * 32;
        float *wmax = iscr + nblk, *bias_out = wmax + kG3MaxWarps, *cmax = bias_out + 4;
        scan_split = p.cs > 1 && ngrp > 4 * nth;
        const bool bias_cta = rank == p.cs - 1;
        auto load_group = [&](int gi, float &b0, float &b1, float &b2, float &b3) {
            const size_t k0 = (size_t)n * p.K + (size_t)gi * 4;
            if (p.act_f16) {
                const uint2 h = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(p.act) + k0));
                const float2 f01 = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), f23 = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
                b0 = f01.x; b1 = f01.y; b2 = f23.x; b3 = f23.y;
            } else {
                const float4 f = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.act) + k0));
                b0 = f.x; b1 = f.y; b2 = f.z; b3 = f.w;
            }
        };
        // K-groups [gbeg, gend): LUT[0] = -LUT[15] of every group (:133-155), _mm256_addv_ps tree per block of 32 activations = 8
        // lanes (lut_ctor.cc:24-31, as in the fp path below), block sums -> the bias CTA; returns the abs-sum maximum (:242-256)
        auto scan = [&](int gbeg, int gend, bool want_max, bool remote) {
            float mm = 0.f;
            for (int g0 = gbeg + warp * 32; g0 < gend; g0 += nth) {   // lane = K-group (coalesced 16-byte loads)
                const int gi = g0 + lane;
                float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
                if (gi < gend) load_group(gi, b0, b1, b2, b3);
                if (want_max) mm = fmaxf(mm, __fadd_rn(__fadd_rn(fabsf(b0), fabsf(b1)), __fadd_rn(fabsf(b2), fabsf(b3))));
                float v = -__fadd_rn(__fadd_rn(__fadd_rn(b0, b1), b2), b3);
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
                if ((lane & 7) == 0 && gi < gend) {
                    if (remote) st_cluster_f32(iscr + (gi >> 3), p.cs - 1, v); else iscr[gi >> 3] = v;
                }
            }
            return mm;
        };
        float m = 0.f;
        if (scan_split) {
            const int per = ((nblk + p.cs - 1) / p.cs) * 8;      // K-groups per CTA, whole blocks
            cluster_wait();                                      // phase 1 ("every CTA of the cluster runs"): DSMEM may be written
            m = scan(rank * per, min(ngrp, rank * per + per), true, true);
        } else {
            // four independent loads per thread in one round; every CTA starts at its own offset (the L2 serves a hot line one
            // request at a time)
            const int off = (int)(((long long)blockIdx.x * 160) % ngrp);
            for (int g = tid; g < ngrp; g += 4 * nth) {
                float b[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gj = g + j * nth;
                    b[j][0] = b[j][1] = b[j][2] = b[j][3] = 0.f;
                    if (gj < ngrp) { int gi = gj + off; if (gi >= ngrp) gi -= ngrp; load_group(gi, b[j][0], b[j][1], b[j][2], b[j][3]); }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) m = fmaxf(m, __fadd_rn(__fadd_rn(fabsf(b[j][0]), fabsf(b[j][1])), __fadd_rn(fabsf(b[j][2]), fabsf(b[j][3]))));
            }
            if (bias_cta) scan(0, ngrp, false, false);           // L1 hits
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) wmax[warp] = m;
        __syncthreads();
        m = wmax[0];
        for (int w = 1; w < WPC; ++w) m = fmaxf(m, wmax[w]);
        if (scan_split) {
            if (tid < p.cs) st_cluster_f32(cmax + rank, tid, m);   // this CTA's maximum -> slot `rank` of every CTA of the cluster
            cluster_sync_all();
            m = cmax[0];
            for (int r = 1; r < p.cs; ++r) m = fmaxf(m, cmax[r]);
        }
        row_scale = __fdiv_rn(m, 127.0f);
        row_ts = (row_scale != 0.0f) ? __fdiv_rn(1.0f, row_scale) : 0.0f;
        if (bias_cta && warp == WPC - 1) {
            // serial over the blocks of the (single) activation group, from 0 (:157).  32 values per shared-memory round trip,
            // the 32 shuffles of a round issued ahead of the dependent adds (fixed trip count); the tail is padded with +0.0,
            // an exact no-op: an accumulator that starts at +0.0 is never -0.0
            float bias = 0.0f;
            for (int k0 = 0; k0 < nblk; k0 += 32) {
                const float v = (k0 + lane < nblk) ? iscr[k0 + lane] : 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k) bias = __fadd_rn(bias, __shfl_sync(0xffffffffu, v, k));
            }
            if (lane == 0) *bias_out = bias;       // forwarded to the leader with the partial sums below
        }
    }
    const uint4 *qrow = reinterpret_cast<const uint4 *>(qb + (size_t)n * p.K * 4);
    const float *lsg = lsb + (size_t)n * nag, *lbg = lbb + (size_t)n * nag;

    for (int c = c_first; c < c_end; ++c) {
        // ---- LUT slice of this chunk -> warp-private table (lane = group) ---------------------
        float lsv[NAG], lbsum = 0.f;
        if (FUSED) {
            // lane = K-group of the chunk; W = lanes per activation group.  Arithmetic = lut_ctor.cc:119-215
            // (AVX2 branch) and partial_max_g4_int8_k8 (:242-256), all with explicit round-to-nearest ops.
            constexpr int NG = QCH * 4, W = (AGQ ? AGQ : 1) * 4;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            if (lane < NG) {
                const size_t k0 = (size_t)n * p.K + ((size_t)c * NG + lane) * 4;
                if (p.act_f16) {
                    const uint2 h = __ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(p.act) + k0));
                    const float2 f01 = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), f23 = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
                    b0 = f01.x; b1 = f01.y; b2 = f23.x; b3 = f23.y;
                } else {
                    const float4 f = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.act) + k0));
                    b0 = f.x; b1 = f.y; b2 = f.z; b3 = f.w;
                }
            }
            float scale, ts;
            if (INT_PATH) { scale = row_scale; ts = row_ts; }
            else {
                float m = __fadd_rn(__fadd_rn(fabsf(b0), fabsf(b1)), __fadd_rn(fabsf(b2), fabsf(b3)));
#pragma unroll
                for (int o = W / 2; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                scale = __fdiv_rn(m, 127.0f);
                ts = (scale != 0.0f) ? __fdiv_rn(1.0f, scale) : 0.0f;
            }
            float od[8];                                   // odd entries 1,3,...,15
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int idx = 2 * e + 1;
                float v = b0;
                v = (idx & 2) ? __fadd_rn(v, b1) : __fsub_rn(v, b1);
                v = (idx & 4) ? __fadd_rn(v, b2) : __fsub_rn(v, b2);
                v = (idx & 8) ? __fadd_rn(v, b3) : __fsub_rn(v, b3);
                od[e] = v;
            }
            // stored entries 0..7: even e = -LUT[15-e] (= -od[(15-e)/2]), odd e = od[e/2]
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float lv = (e & 1) ? od[e >> 1] : -od[(15 - e) >> 1];
                int q = __float2int_rn(__fmul_rn(lv, ts));
                q = max(-128, min(127, q));
                if (e < 4) lo |= (uint32_t)(q & 0xff) << (8 * e); else hi |= (uint32_t)(q & 0xff) << (8 * (e - 4));
            }
            if (lane < NG) reinterpret_cast<uint2 *>(tab)[lane] = make_uint2(lo, hi);
            if (!INT_PATH) {
                // LUT bias: _mm256_addv_ps tree per 8 groups (lut_ctor.cc:24-31), serial over the blocks of a group (:157)
                float v = -od[7];                          // LUT[0] = -LUT[15]
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
                v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
#pragma unroll
                for (int a = 0; a < NAG; ++a) {
                    lsv[a] = __shfl_sync(0xffffffffu, scale, a * W);
                    float bias = 0.f;
#pragma unroll
                    for (int k = 0; k < W / 8; ++k) bias = __fadd_rn(bias, __shfl_sync(0xffffffffu, v, a * W + 8 * k));
                    lbsum += bias;
                }
            }
        } else {
            if (lane < QCH * 4) {
                const uint4 L = __ldg(qrow + (size_t)c * QCH * 4 + lane);
                if (SYM) reinterpret_cast<uint2 *>(tab)[lane] = make_uint2(L.x, L.y);
                else reinterpret_cast<uint4 *>(tab)[lane] = make_uint4(L.x, L.y, __byte_perm(L.w, 0, 0x0123), __byte_perm(L.z, 0, 0x0123));
            }
            if (!INT_PATH) {
#pragma unroll
                for (int a = 0; a < NAG; ++a) { lsv[a] = __ldg(lsg + c * NAG + a); lbsum += __ldg(lbg + c * NAG + a); }
            }
        }
        const int li = c - c_first, buf = li & (nbuf - 1);
        unsigned char *stage = stage0 + (size_t)buf * p.blk_bytes;
        if (nbuf == 2) {
            if (c + 1 < c_end && lane == 0) {      // request chunk c+1 into the other buffer (consumed at c-1, see the __syncwarp below)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_load(stage0 + (size_t)((li + 1) & 1) * p.blk_bytes, rsb_base + (size_t)(c + 1) * p.blk_bytes, (uint32_t)p.blk_bytes, mbar + ((li + 1) & 1));
            }
        } else if (li > 0 && lane == 0) {          // single buffer: the other resident warps cover this warp's load latency
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(stage0, rsb_base + (size_t)c * p.blk_bytes, (uint32_t)p.blk_bytes, mbar);
        }
        mbar_wait_parity(mbar + buf, nbuf == 2 ? ((li >> 1) & 1) : (li & 1));
        __syncwarp();
        if (tid == 0 && c == c_first) TMAC_TRACE(3);
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
        __syncwarp();                              // stage / table are rewritten by a later chunk
    }
    if (tid == 0) TMAC_TRACE(4);
    if (p.pdl_late) pdl_launch_dependents();

    // ---- CTA reduction (fixed warp order); red aliases the consumed stages ----------------------
    __syncthreads();
    {
        float *r = red + (size_t)warp * RSB + lane * RW;
#pragma unroll
        for (int i = 0; i < RW; ++i) r[i] = INT_PATH ? __int_as_float(iacc[i]) : cacc[i];
    }
    __syncthreads();
    if (tid == 0) TMAC_TRACE(5);
    const int nthreads = WPC * 32;
    if (p.cs > 1 && !scan_split) cluster_wait();  // every CTA of the cluster has started: its shared memory may be written
    if (FUSED && INT_PATH && p.cs > 1 && rank == p.cs - 1 && tid == (WPC - 1) * 32) {
        float *bias_out = iscr + (p.K >> 5) + kG3MaxWarps;
        st_cluster_f32(bias_out, 0, *bias_out);    // written by this thread after the row scan
    }
    for (int t = tid; t < RSB; t += nthreads) {
        float fsum = 0.f; int isum = 0;
        for (int w = 0; w < WPC; ++w) {
            const float v = red[(size_t)w * RSB + t];
            if (INT_PATH) isum += __float_as_int(v); else fsum += v;
        }
        const float mine = INT_PATH ? __int_as_float(isum) : fsum;
        if (p.cs > 1) st_cluster_f32(cl + (size_t)rank * RSB + t, 0, mine);   // into the leader's shared memory
        else cl[t] = mine;
    }
    if (p.cs > 1) cluster_sync_all(); else __syncthreads();
    if (tid == 0) TMAC_TRACE(6);
    if (rank != 0) return;
    for (int t = tid; t < RSB; t += nthreads) {
        float fsum = 0.f; int isum = 0;
        for (int k2 = 0; k2 < p.cs; ++k2) {
            const float v = cl[(size_t)k2 * RSB + t];
            if (INT_PATH) isum += __float_as_int(v); else fsum += v;
        }
        const int row = (p.rsb0 + rsb) * RSB + t;
        if (row >= p.row_begin && row < p.row_end) {
            float out;
            if (INT_PATH) {
                // C = ((sum_b alpha_b*CBits_b) * LUT_Scales[0] + LUT_Biases[0]*alpha_0) * Scales[0]
                // (python/t_mac/ops/qgemm.py:160,171-174); isum = sum_b 2*alpha_b*CBits_b exactly.
                const float cb = __fmul_rn((float)isum, 0.5f);
                const float t1 = __fmul_rn(cb, FUSED ? row_scale : __ldg(lsg));
                const float t2 = __fmul_rn(FUSED ? iscr[(p.K >> 5) + kG3MaxWarps] : __ldg(lbg), 0.5f);
                out = __fmul_rn(__fadd_rn(t1, t2), p.scale0);
            } else
                out = fsum;
            const size_t o = (size_t)n * p.ldc + (size_t)(row - p.c_row0);
            if (p.out_f16) reinterpret_cast<__half *>(Cb)[o] = __float2half_rn(out);
            else reinterpret_cast<float *>(Cb)[o] = out;
            for (int q = 0; q < p.npeer; ++q) {      // the all-gather, fused into the epilogue: one store per peer and row
                if (p.out_f16) reinterpret_cast<__half *>(p.Cpeer[q])[o] = __float2half_rn(out);
                else reinterpret_cast<float *>(p.Cpeer[q])[o] = out;
            }
        }
    }
    if (tid == 0) TMAC_TRACE(7);
}

// ------------------------------------------------------------------------------------------
// cbits_kernel (debug / parity gate G2): per-plane int32 sums in the reference plane layout.
// One thread per (row, plane); straightforward lookups in the full 16-entry table, decoding the
// stream layout nibble by nibble.  Not a performance path.
// ------------------------------------------------------------------------------------------
static __global__ void cbits_kernel(const unsigned char *W, const int8_t *qlut, int32_t *cbits, int Mout, int K, int bits,
                             int pb, int qch, int nchunk, size_t rsb_stride, size_t blk_stride, int N) {
    const int rw = 8 / pb, rsbsz = 32 * rw;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * Mout * bits;
    if (t >= total) return;
    const int n = (int)(t / ((long long)Mout * bits));
    const int rem = (int)(t % ((long long)Mout * bits));
    const int row = rem / bits, b = rem % bits;
    const int rsb = row / rsbsz, lane = (row % rsbsz) / rw, i = (row % rsbsz) % rw;
    const int8_t *lut = qlut + (size_t)n * K * 4;
    int sum = 0;
    for (int c = 0; c < nchunk; ++c) {
        const uint32_t *wq = reinterpret_cast<const uint32_t *>(W + (size_t)rsb * rsb_stride + (size_t)c * blk_stride);
        for (int q = 0; q < qch; ++q) {
            const uint32_t *w4 = wq + ((size_t)q * 32 + lane) * 4;
            for (int k = 0; k < 4; ++k) {
                const int g = (c * qch + q) * 4 + k;
                uint32_t j, neg;
                if (pb == 4) { const uint32_t nb = (w4[k] >> (4 * (4 * i + b))) & 15u; j = nb & 7u; neg = nb >> 3; }
                else if (pb == 2) {
                    j = (w4[k] >> (4 * (2 * i + b))) & 7u;
                    const int pair = k >> 1, nn = b + 2 * (k & 1);
                    neg = (w4[2 * pair + i / 2] >> (16 * (i % 2) + 4 * nn + 3)) & 1u;
                } else {
                    j = (w4[k] >> (4 * i)) & 7u;
                    neg = (w4[i / 2] >> (16 * (i % 2) + 4 * k + 3)) & 1u;
                }
                const uint32_t idx = neg ? (8u | (j ^ 7u)) : j;
                sum += lut[(size_t)g * 16 + idx];
            }
        }
    }
    const int p = (row / 8) * 8 * bits + b * 8 + row % 8;
    cbits[(size_t)n * Mout * bits + p] = sum;
}

// ------------------------------------------------------------------------------------------
// preprocessor_kernel.  grid = (ceil(nag / agb), N); block = 256.
// Each CTA handles `agb` consecutive activation groups of one activation row (agb == nag when
// act_group_size == K, i.e. a single CTA per row).  fp32 op order identical to the AVX2 branch of
// lut_ctor.cc (no FMA: only adds, one IEEE division, one IEEE reciprocal, one multiply per entry).
// ------------------------------------------------------------------------------------------
constexpr int kPreThreads = 256;

template <typename TIn>
__device__ __forceinline__ float ld_act(const TIn *p, size_t i);
template <> __device__ __forceinline__ float ld_act<float>(const float *p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_act<__half>(const __half *p, size_t i) { return __half2float(p[i]); }

template <typename TIn>
__global__ void __launch_bounds__(kPreThreads) preprocessor_kernel(const TIn *B, float *lut_scales, float *lut_biases,
                                                                    int8_t *qlut, int K, int ags, int agb) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int n = blockIdx.y;
    const int nag = K / ags;
    const int a_begin = blockIdx.x * agb;
    const int a_end = min(nag, a_begin + agb);
    const int gpa = ags / 4;                              // groups per act group
    const int g_begin = a_begin * gpa, g_end = a_end * gpa;
    const int ng = g_end - g_begin;
    int *smax = reinterpret_cast<int *>(smem);            // [agb] max abs-sum (float bits, >= 0)
    float *l0 = reinterpret_cast<float *>(smax + agb);    // [ng] LUT[0] of every group
    float *blk = l0 + ng;                                 // [ng/8] addv8 of every 32-activation block
    const TIn *b = B + (size_t)n * K;
    const int tid = threadIdx.x;

    // Programmatic dependent launch: let the consumer GEMV start streaming its (static) weights now;
    // the activations and the output buffers belong to the stream order, so wait before touching them.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int a = tid; a < a_end - a_begin; a += kPreThreads) smax[a] = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    // pass 1: abs-sum max per act group (lut_ctor.cc:242-256; association (a0+a1)+(a2+a3) :251)
    for (int g = tid; g < ng; g += kPreThreads) {
        const size_t k0 = (size_t)(g_begin + g) * 4;
        const float a0 = fabsf(ld_act(b, k0)), a1 = fabsf(ld_act(b, k0 + 1));
        const float a2 = fabsf(ld_act(b, k0 + 2)), a3 = fabsf(ld_act(b, k0 + 3));
        const float s = __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
        atomicMax(&smax[g / gpa], __float_as_int(s));
    }
    __syncthreads();
    // pass 2: tables
    for (int g = tid; g < ng; g += kPreThreads) {
        const size_t k0 = (size_t)(g_begin + g) * 4;
        const float b0 = ld_act(b, k0), b1 = ld_act(b, k0 + 1), b2 = ld_act(b, k0 + 2), b3 = ld_act(b, k0 + 3);
        const float scale = __fdiv_rn(__int_as_float(smax[g / gpa]), 127.0f);      // :256
        const float ts = (scale != 0.0f) ? __fdiv_rn(1.0f, scale) : 0.0f;          // :124
        float lut[16];
#pragma unroll
        for (int e = 1; e < 16; e += 2) {                                          // :133-151
            float v = b0;
            v = (e & 2) ? __fadd_rn(v, b1) : __fsub_rn(v, b1);
            v = (e & 4) ? __fadd_rn(v, b2) : __fsub_rn(v, b2);
            v = (e & 8) ? __fadd_rn(v, b3) : __fsub_rn(v, b3);
            lut[e] = v;
        }
#pragma unroll
        for (int e = 0; e < 16; e += 2) lut[e] = -lut[15 - e];                     // :153-155
        l0[g] = lut[0];
        uint32_t packed[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t acc = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int q = __float2int_rn(__fmul_rn(lut[4 * w + e], ts));             // :160-171 (rne)
                q = max(-128, min(127, q));                                       // packs saturation :173-176
                acc |= (uint32_t)(q & 0xff) << (8 * e);
            }
            packed[w] = acc;
        }
        reinterpret_cast<uint4 *>(qlut + (size_t)n * K * 4)[g_begin + g] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
    __syncthreads();
    // biases: per 32-activation block the _mm256_addv_ps tree (:24-31), then a serial sum over
    // the blocks of the act group (:157), starting from 0.
    for (int bi = tid; bi < ng / 8; bi += kPreThreads) {
        const float *v = l0 + bi * 8;
        const float r0 = __fadd_rn(v[4], v[0]), r1 = __fadd_rn(v[5], v[1]);
        const float r2 = __fadd_rn(v[6], v[2]), r3 = __fadd_rn(v[7], v[3]);
        blk[bi] = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
    }
    __syncthreads();
    const int bpa = ags / 32;                              // blocks per act group
    for (int a = tid; a < a_end - a_begin; a += kPreThreads) {
        float bias = 0.0f;
        for (int k = 0; k < bpa; ++k) bias = __fadd_rn(bias, blk[a * bpa + k]);
        lut_biases[(size_t)n * nag + a_begin + a] = bias;
        lut_scales[(size_t)n * nag + a_begin + a] = __fdiv_rn(__int_as_float(smax[a]), 127.0f);
    }
}


// ------------------------------------------------------------------------------------------
// peer_barrier_kernel: "every rank's launches before this point have completed" as a stream-ordered fact on every rank,
// without a collective library call.  flags[0..world) are written by the peers (slot q by rank q), flags[world] is this
// rank's own barrier count.  One block; thread q handles peer q.  Stores issued by EARLIER kernels of this stream
// (e.g. the GEMV epilogues' peer stores) are complete when this kernel starts, so publishing the count after them
// orders them before any peer's wake-up.  Bounded wait; *err != 0 if a peer never arrived.
// ------------------------------------------------------------------------------------------
static __global__ void peer_barrier_kernel(unsigned *flags, unsigned *const *peer_flags, int rank, int world, int *err) {
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) { s_epoch = flags[world] + 1u; flags[world] = s_epoch; }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    __threadfence_system();
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(peer_flags[q] + rank), "r"(epoch) : "memory");
    unsigned v;
    int spins = 0;
    do {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
    } while ((int)(v - epoch) < 0 && ++spins < (1 << 22));
    if ((int)(v - epoch) < 0) atomicExch(err, 1);
    __threadfence_system();
}

}  // namespace tmac_b200
