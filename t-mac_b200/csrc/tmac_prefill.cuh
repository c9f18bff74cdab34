// tmac_prefill.cuh -- the N>1 (prefill) tile of qgemm_lut on the 5th-generation tensor cores.
//
// Why tensor cores are legitimate here and how parity is kept (hard part 1):
// the reference rounds the *group-of-4 partial sums* (the LUT entries), so a plain int8 GEMM of
// quantised activations is different arithmetic.  The contraction is therefore taken over the LUT
// itself.  With the odd symmetry LUT[15-i] = -LUT[i] (lut_ctor.cc:153-155) the 8 stored entries T8[n][g][0..7]
// of token n, K-group g are enough, and for one weight row m
//     sum_planes 2*alpha_b * sign * LUT_g[idx_b]  =  sum_{e<8} T8[n][g][e] * S[m][g][e],
//     S[m][g][e] = sum_b 2*alpha_b * sign(m,b,g) * [j(m,b,g) == e]          (int8, |S| <= 15)
// i.e. an exact dense int8 contraction of length 8 per K-group:  I[m][n] = S[m][:] . T8[n][:]  -- a true
// int8 GEMM with int32 accumulation, bit-identical to the integer sums of the GEMV path.  It runs as
// tcgen05.mma kind::i8 (M = 128 weight rows, N = 128 tokens, K = 32 bytes per instruction), accumulator in
// TMEM.  Per activation group (64 K positions = 128 contraction bytes = 4 MMAs) the int32 tile is drained
// with tcgen05.ld and folded into fp32 registers with lut_scale[n][ag] * 0.5*scale[m][wg]; the LUT-bias /
// zero-point terms are a rank-(K/group_size) correction added at the end.
//
//   warps 0..3   producers : cp.async the (row super-block, chunk) block, expand the packed codes of "their"
//                            weight row into S (one 64-bit one-hot-signed vector per (row, group), through a
//                            256-entry shared-memory table), copy the T8 slice of 128 tokens, both in the
//                            UMMA K-major no-swizzle canonical layout; fence.proxy.async; arrive full[s].
//   warp  4      MMA issuer: one elected thread issues 4 tcgen05.mma per activation group, tcgen05.commit
//                            frees the stage and publishes the accumulator (double-buffered in TMEM).
//   warps 5..12  epilogue  : tcgen05.ld 128 lanes x 64 columns per warp pair, fp32 FMA, final store.
//
// Supported in this round: PB == 2 (W2), chunk = 128 K (QCH 8), act group 64 (AGQ 4), per-row scales
// (+ zero points), symmetric LUT.  Everything else takes the GEMV kernel once per activation row.
#pragma once
#include "tmac_kernels.cuh"

namespace tmac_b200 {

struct PrefillParams {
    const unsigned char *W;        // stream layout of the tensor (first super-block)
    const int8_t *qlut;            // [N][K/4][16]
    const float *lut_scales, *lut_biases;   // [N][K/64]
    void *C;                       // [N][ldc]
    int N, K, Mout, ldc, out_f16;
    int nchunk, zp, sd, blk_bytes;
    size_t rsb_stride;
    const unsigned char *lut_tiles; // output of lut_tile_kernel: [token tile][step][kPfRec]
    long long *dbg;                // optional trace [role][step][4] (TMAC_ENABLE_TRACE builds)
};
#ifdef TMAC_ENABLE_TRACE
#define PF_TRACE(role, step, k) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (step) < 32) p.dbg[((role) * 32 + (step)) * 4 + (k)] = clock64(); } while (0)
#else
#define PF_TRACE(role, step, k) do { } while (0)
#endif

constexpr int kPfNT = 128;                       // tokens per CTA pass (MMA N)
constexpr int kPfStageBytes = 128 * 128;         // one operand tile: 128 rows x 128 contraction bytes

__device__ __forceinline__ uint32_t pf_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t pf_desc(uint32_t saddr) {
    // K-major, no swizzle: core matrix = 8 rows x 16 B (128 B); chunk kc, row group rn at (kc*16 + rn)*128
    //   leading byte offset (between the two 16-byte K chunks of one MMA) = 2048, stride byte offset (8-row groups) = 128
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(2048u >> 4) << 16;
    d |= (uint64_t)(128u >> 4) << 32;
    d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
    return d;
}
__device__ __forceinline__ void pf_mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pf_s32(b)), "r"(c)); }
__device__ __forceinline__ void pf_mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pf_s32(b)) : "memory"); }
__device__ __forceinline__ void pf_mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(pf_s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pf_commit(uint64_t *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pf_s32(b)) : "memory");
}
#define PF_TMEM_LD16(r, taddr)                                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"             \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),      \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                       \
                 : "r"(taddr))
#define PF_TMEM_LD32(r, taddr)                                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21," \
                 "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                                   \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),      \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),       \
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                 \
                 : "r"(taddr))

// ---------------------------------------------------------------------------------------------------
// lut_tile_kernel: QLUT [N][K/4][16] + LUT_Scales/Biases [N][K/64]  ->  per (token tile of 128, activation group) one
// contiguous 17920-byte record that the GEMM stage receives with ONE bulk copy (TMA):
//     [16384 B]  the 8 stored LUT entries of the 16 groups of the step for 128 tokens, already in the UMMA K-major
//                canonical order: chunk kc (groups 2kc, 2kc+1), token t at ((kc*16 + t/8)*128 + (t%8)*16)
//     [  512 B]  lut_scale[token][step]
//     [ 1024 B]  (lut_bias[token][2c], lut_bias[token][2c+1]) of the step's chunk c = step/2
// Tokens >= N are zero.  grid = (steps, token tiles), block = 128 (thread = token).
// ---------------------------------------------------------------------------------------------------
constexpr int kPfRec = 16384 + 512 + 1024;
__global__ void __launch_bounds__(128) lut_tile_kernel(const int8_t *qlut, const float *ls, const float *lb, unsigned char *out, int N, int K) {
    const int step = blockIdx.x, tile = blockIdx.y, t = threadIdx.x, n = tile * 128 + t;
    const int nag = K / 64;
    unsigned char *rec = out + ((size_t)tile * nag + step) * kPfRec;
    const uint2 *src = reinterpret_cast<const uint2 *>(qlut) + ((size_t)n * (K / 4) + (size_t)step * 16) * 2;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
        uint2 lo = make_uint2(0, 0), hi = make_uint2(0, 0);
        if (n < N) { lo = __ldg(src + kc * 4); hi = __ldg(src + kc * 4 + 2); }
        *reinterpret_cast<uint4 *>(rec + ((size_t)(kc * 16 + (t >> 3)) * 128 + (t & 7) * 16)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    float l = 0.f; float2 b2 = make_float2(0.f, 0.f);
    if (n < N) { l = ls[(size_t)n * nag + step]; b2 = make_float2(lb[(size_t)n * nag + (step & ~1)], lb[(size_t)n * nag + (step | 1)]); }
    reinterpret_cast<float *>(rec + 16384)[t] = l;
    reinterpret_cast<float2 *>(rec + 16384 + 512)[t] = b2;
}

constexpr int kPfStages = 4;                     // stages: A tile 16 KB (expanded by the producers) + record 17.5 KB (TMA)
constexpr int kPfProdWarps = 8;
constexpr int kPfWarpMma = 8, kPfWarpTma = 9, kPfWarpEpi = 10;
#undef PF_THREADS
constexpr int kPfEpiWarps = 16;                  // 4 lane quarters x 4 column groups of 32 tokens
constexpr int kPfThreads2 = (10 + kPfEpiWarps) * 32;

__device__ __forceinline__ uint64_t pf_pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t pf_fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t pf_mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

__device__ __forceinline__ void pf_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pf_s32(dst)), "l"(src), "r"(bytes), "r"(pf_s32(bar)) : "memory");
}
__device__ __forceinline__ void pf_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pf_s32(bar)), "r"(bytes) : "memory");
}

// grid = (row super-blocks of 128 rows, ceil(N / 128)), block = 832 threads, 1 CTA per SM.
//   warps 0..7  : A producers, thread = (weight row, half of the step's 16 groups)
//   warp  8     : MMA issuer (one thread)        warp 9 : TMA issuer (one thread): one bulk copy per step
//   warps 10..25: epilogue, thread = (weight row, 32-token column group)
__global__ void __launch_bounds__(kPfThreads2, 1) prefill_w2_kernel(const PrefillParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                                              // [S][16 KB]
    unsigned char *sR = sA + kPfStages * kPfStageBytes;                    // [S][kPfRec] B tile + lut scales + lut bias pairs
    unsigned char *raw = sR + kPfStages * kPfRec;                          // [2][blk_bytes] packed block (codes + scales)
    const int rawsz = (p.blk_bytes + 127) & ~127;
    uint64_t *xtab = reinterpret_cast<uint64_t *>(raw + 2 * rawsz);        // [256] expansion table
    uint64_t *bars = xtab + 256;
    uint64_t *full = bars, *empty = bars + kPfStages, *accfull = bars + 2 * kPfStages, *accempty = accfull + 2;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rsb = blockIdx.x, tile = blockIdx.y, n0 = tile * kPfNT;
    const int ntok = min(kPfNT, p.N - n0);
    const int nag = p.K / 64;

    for (int e = tid; e < 256; e += kPfThreads2) {
        // index byte: low nibble = plane 0 (neg<<3 | j), high nibble = plane 1; value = +-1 at byte j0, +-2 at byte j1
        int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        v[e & 7] += (e & 8) ? -1 : 1;
        v[(e >> 4) & 7] += (e & 0x80) ? -2 : 2;
        uint64_t x = 0;
        for (int i = 0; i < 8; ++i) x |= (uint64_t)(uint8_t)(int8_t)v[i] << (8 * i);
        xtab[e] = x;
    }
    if (tid == 0) {
        for (int i = 0; i < kPfStages; ++i) { pf_mbar_init(full + i, kPfProdWarps + 1); pf_mbar_init(empty + i, kPfEpiWarps); }
        for (int i = 0; i < 2; ++i) { pf_mbar_init(accfull + i, 1); pf_mbar_init(accempty + i, kPfEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kPfWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(pf_s32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const unsigned char *rsb_base = p.W + (size_t)rsb * p.rsb_stride;
    const int nsteps = 2 * p.nchunk;                 // activation groups (two per chunk)

    if (warp < kPfProdWarps) {
        // ======================= A producers =======================
        const int r = tid & 127, gh = tid >> 7;      // weight row of the tile, which 8 of the step's 16 groups
        const int wl = r >> 2, wi = r & 3;           // lane / row-in-lane of the stream layout (RW = 4)
        const int n16 = p.blk_bytes >> 4;
        for (int i = tid; i < n16; i += 256) cp_async16_plain(raw + i * 16, rsb_base + i * 16);
        cp_async_commit();
        for (int c = 0; c < p.nchunk; ++c) {
            unsigned char *rb = raw + (size_t)(c & 1) * rawsz;
            cp_async_wait_all();
            asm volatile("bar.sync 1, 256;" ::: "memory");                 // block c visible to all producers; block c-1 dead
            if (c + 1 < p.nchunk) {
                unsigned char *nb = raw + (size_t)((c + 1) & 1) * rawsz;
                const unsigned char *src = rsb_base + (size_t)(c + 1) * p.blk_bytes;
                for (int i = tid; i < n16; i += 256) cp_async16_plain(nb + i * 16, src + i * 16);
                cp_async_commit();
            }
            const uint32_t *words = reinterpret_cast<const uint32_t *>(rb);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int step = 2 * c + h, s = step % kPfStages;
                if (tid == 0) PF_TRACE(0, step, 0);
                pf_mbar_wait(empty + s, ((step / kPfStages) & 1) ^ 1);     // stage released by the epilogue of step - S
                if (tid == 0) PF_TRACE(0, step, 1);
                unsigned char *a_dst = sA + (size_t)s * kPfStageBytes;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int kc = gh * 4 + kk;
                    uint64_t v2[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int gq = h * 16 + kc * 2 + u;                // group within the chunk (0..31)
                        const int q = gq >> 2, k = gq & 3;
                        const uint32_t *w4 = words + ((size_t)q * 32 + wl) * 4;
                        const uint32_t jb = (w4[k] >> (8 * wi)) & 0x77u;
                        const uint32_t ng = (w4[2 * (k >> 1) + (wi >> 1)] >> (16 * (wi & 1) + 8 * (k & 1) + 3)) & 0x11u;
                        v2[u] = xtab[jb | (ng << 3)];
                    }
                    *reinterpret_cast<uint4 *>(a_dst + ((size_t)(kc * 16 + (r >> 3)) * 128 + (r & 7) * 16)) =
                        make_uint4((uint32_t)v2[0], (uint32_t)(v2[0] >> 32), (uint32_t)v2[1], (uint32_t)(v2[1] >> 32));
                }
                if (tid == 0) PF_TRACE(0, step, 2);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) pf_mbar_arrive(full + s);
                __syncwarp();
            }
        }
    } else if (warp == kPfWarpTma) {
        // ======================= TMA issuer: one record per step =======================
        if (lane == 0) {
            const unsigned char *src = p.lut_tiles + (size_t)tile * nag * kPfRec;
            for (int step = 0; step < nsteps; ++step) {
                const int s = step % kPfStages;
                pf_mbar_wait(empty + s, ((step / kPfStages) & 1) ^ 1);
                pf_expect_tx(full + s, kPfRec);
                pf_bulk_g2s(sR + (size_t)s * kPfRec, src + (size_t)step * kPfRec, kPfRec, full + s);
            }
        }
    } else if (warp == kPfWarpMma) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPfNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int step = 0; step < nsteps; ++step) {
                const int s = step % kPfStages, b = step & 1;
                PF_TRACE(1, step, 0);
                pf_mbar_wait(full + s, (step / kPfStages) & 1);
                PF_TRACE(1, step, 1);
                pf_mbar_wait(accempty + b, ((step >> 1) & 1) ^ 1);
                PF_TRACE(1, step, 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = pf_s32(sA + (size_t)s * kPfStageBytes), b0 = pf_s32(sR + (size_t)s * kPfRec);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint64_t da = pf_desc(a0 + i * 4096), db = pf_desc(b0 + i * 4096);
                    const uint32_t acc = i > 0 ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(tmem + b * kPfNT), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
                pf_commit(accfull + b);          // accumulator complete (the stage itself is released by the epilogue)
                PF_TRACE(1, step, 3);
            }
        }
    } else {
        // ======================= epilogue: thread = (weight row, 32-token column group) =======================
        const int ew = warp - kPfWarpEpi;            // 0..15
        const int lq = warp & 3;                     // TMEM lane quarter this warp may access
        const int cg = ew >> 2;                      // column group: tokens 32*cg .. 32*cg+31
        const int r = lq * 32 + lane;                // weight row of the tile
        const int wl = r >> 2, wi = r & 3;
        uint64_t acc2[16];                           // 32 fp32 accumulators as f32x2 pairs (FFMA2 / FMUL2)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc2[j] = 0ull;
        // weight scale / zero point of (row, chunk), prefetched one chunk ahead (L2 hits, latency off the critical path)
        auto ld_s = [&](int c) { return 0.5f * load_scale(rsb_base + (size_t)c * p.blk_bytes + 4096, p.sd, wl * 4 + wi); };
        auto ld_z = [&](int c) { return p.zp ? load_scale(rsb_base + (size_t)c * p.blk_bytes + 4096 + (size_t)128 * p.sd, p.sd, wl * 4 + wi) : 0.f; };
        float hs = ld_s(0), zz = ld_z(0), hs_n = hs, zz_n = zz;
        for (int step = 0; step < nsteps; ++step) {
            const int b = step & 1, c = step >> 1, s = step % kPfStages;
            if (!(step & 1) && c + 1 < p.nchunk) { hs_n = ld_s(c + 1); zz_n = ld_z(c + 1); }
            if (warp == kPfWarpEpi && lane == 0) PF_TRACE(2, step, 0);
            pf_mbar_wait(accfull + b, (step >> 1) & 1);
            if (warp == kPfWarpEpi && lane == 0) PF_TRACE(2, step, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + b * kPfNT + cg * 32;
            const unsigned char *rec = sR + (size_t)s * kPfRec;
            const uint64_t *ls2 = reinterpret_cast<const uint64_t *>(rec + 16384) + cg * 16;     // (ls[t], ls[t+1]) pairs
            const uint64_t hs2 = pf_pack2(hs, hs);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {                 // two 16-column slices
                uint32_t v[16];
                PF_TMEM_LD16(v, taddr + hh * 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (hh == 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) pf_mbar_arrive(accempty + b);   // the accumulator buffer may be overwritten
                    __syncwarp();                                  // reconverge before the next .aligned tcgen05.ld
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint64_t w2 = pf_mul2(hs2, ls2[hh * 8 + j]);
                    acc2[hh * 8 + j] = pf_fma2(w2, pf_pack2((float)(int)v[2 * j], (float)(int)v[2 * j + 1]), acc2[hh * 8 + j]);
                }
            }
            if (step & 1) {
                // LUT-bias / zero-point term of the chunk: (0.5*s + z)[row] * (lb[2c] + lb[2c+1])[token]
                const float wz = hs + zz;
                const uint64_t wz2 = pf_pack2(wz, wz);
                const float4 *lb4 = reinterpret_cast<const float4 *>(rec + 16384 + 512) + cg * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 l = lb4[j];                 // two tokens: (lb0, lb1), (lb0, lb1)
                    acc2[j] = pf_fma2(wz2, pf_pack2(l.x + l.y, l.z + l.w), acc2[j]);
                }
                hs = hs_n; zz = zz_n;
            }
            if (warp == kPfWarpEpi && lane == 0) PF_TRACE(2, step, 2);
            __syncwarp();
            if (lane == 0) pf_mbar_arrive(empty + s);        // A tile and record of this step are dead: the stage may be refilled
            __syncwarp();
        }
        const int row = rsb * 128 + r;
        if (row < p.Mout) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float lo, hi;
                asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[j]));
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int t = cg * 32 + 2 * j + u;
                    if (t < ntok) {
                        const size_t o = (size_t)(n0 + t) * p.ldc + row;
                        const float val = u ? hi : lo;
                        if (p.out_f16) reinterpret_cast<__half *>(p.C)[o] = __float2half_rn(val);
                        else reinterpret_cast<float *>(p.C)[o] = val;
                    }
                }
            }
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kPfWarpMma) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

}  // namespace tmac_b200
