// tmac_chain.cuh -- the dependent chain of GEMVs as ONE persistent launch, second form: gemv3's decomposition kept resident.
//
// tmac_seq.cuh (stream-K shares, partial sums and LUT records exchanged through HBM words) turned out latency-bound at the
// granularity of one 12.7 MB GEMV: three CTA-wide phases and two inter-CTA hops per op.  This kernel
// keeps what measured well in the launch chain -- gemv3's decomposition: cluster of 8 CTAs x 4 warps per 128-row super-block,
// one (super-block, chunk) block per warp, LUT slice built by the warp itself, K slices summed through distributed shared
// memory in rank order -- and removes what the launch chain loses per launch:
//   * the kernel boundary (0.77 us until griddepcontrol.wait returns, plus a fresh block-scheduler placement per launch) becomes
//     one GRID BARRIER per op: a cluster barrier, one release-increment per cluster on a global counter, one polling thread per CTA;
//   * the exposed weight stream: every warp requests its block of op i+1 (one cp.async.bulk, 4.6 KB) the moment it has finished
//     reading op i's block, i.e. before the reduction, the stores and the barrier of op i -- weights do not depend on activations;
//   * CTA placement is fixed for the whole chain (688 CTAs of 128 threads, all resident: 4 or 5 per SM).
// Inputs and outputs are plain vectors: op i+1 may read op i's C (the grid barrier orders them), so a chain with true data
// dependencies needs nothing else.  Same arithmetic as gemv3's fused path: LUT bytes identical to the preprocessor, fp32 sums in
// fixed (warp, then cluster rank) order -> bit-identical to tmac_b200_gemv for every op.  Waits are bounded (error flag).
#pragma once
#include "tmac_kernels.cuh"
#include "tmac_seq.cuh"     // {value, epoch} words: seq_publish / seq_wait

namespace tmac_b200 {

constexpr int kChainCS = 8;                 // CTAs per cluster (K slices of one row super-block)
constexpr int kChainWarps = 4;              // warps per CTA
constexpr int kChainSpin = 1 << 22;

struct ChainOp {
    const unsigned char *W;                 // stream layout of the tensor
    const float *x;                         // input vector [K] (external, or an earlier op's C + offset)
    void *C;                                // output [Mout]
    unsigned long long rsb_stride;
    int K, Mout, nrsb, nchunk;
    int blk_bytes, bpw;                     // bytes per block; chunks per warp = ceil(nchunk / 32)
    int zp, one_scale, sd, out_f16;
    float scale0;
    int in_op;                              // producer op of x, or -1 (external input)
    uint2 *ll_out;                          // data-flow mode: [rows] {value bits, epoch} words written next to C
    const uint2 *ll_in;                     // data-flow mode: the producer's words, already offset (NULL: external input, plain loads)
};

struct ChainParams {
    const ChainOp *ops;
    int nops;
    int max_blk;                            // stage bytes per warp
    unsigned int *bar;                      // grid barrier counter (monotonic over launches)
    unsigned int *epochs;                   // [grid] launches seen by each CTA
    int *err;
    int flags;                              // bit 0: request op i+1's block at the top of op i (else after op i's lookups); bit 1: one poller per cluster
    long long *trace;                       // debug: [nops][grid][16] globaltimer stamps of warp 0, or NULL
};
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acq() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ long long chain_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ bool seq_chain_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok = 0;
    for (int spins = 0; spins < kChainSpin && !ok; ++spins)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

template <int PB, int QCH, int AGQ>
__global__ void __launch_bounds__(kChainWarps * 32, 5) chain_kernel(const ChainParams p, const uint32_t wtx, const uint32_t wty) {
    constexpr int RW = 8 / PB, RSB = 32 * RW;
    constexpr int NAG = QCH / AGQ, NG = QCH * 4, WL = AGQ * 4;
    constexpr int WPC = kChainWarps, CS = kChainCS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster_ctarank(), rsb = blockIdx.x / CS;
    const int nclusters = gridDim.x / CS;
    // shared memory: cl [CS][RSB] (cluster partials, leader) | red [WPC][RSB] | per warp: stage (max_blk) + table (NG * 8) | mbar [WPC]
    float *cl = reinterpret_cast<float *>(smem);
    float *red = cl + CS * RSB;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(red + WPC * RSB);
    const int per_warp = 2 * p.max_blk + NG * 8;                  // two stages: the block of op i+1 is requested while op i is computed
    unsigned char *stage0 = wbase + (size_t)warp * per_warp;
    unsigned char *tab = stage0 + 2 * p.max_blk;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbase + (size_t)WPC * per_warp) + 2 * warp;     // one per stage
    volatile int *go = reinterpret_cast<volatile int *>(wbase + (size_t)WPC * per_warp + WPC * 16);   // barrier generation, written by the cluster leader
    const unsigned epoch0 = p.epochs[blockIdx.x];                 // launches before this one
    const unsigned bar_base = epoch0 * (unsigned)p.nops * (unsigned)nclusters;

    if (lane == 0) { mbar_init1(mbar); mbar_init1(mbar + 1); mbar_fence_init(); }
    if (tid == 0) *go = 0;
    __syncwarp();
    cluster_sync_all();                                            // every CTA of the cluster runs: its shared memory may be written

    // the block of chunk c of op `op` for this warp -> stage (one bulk copy), if the warp has one
    // stage buffer b <- the block of chunk c of op `op` for this warp (one bulk copy), if the warp has one
    auto request = [&](int op, int c, int b) {
        const ChainOp &o = p.ops[op];
        if (rsb < o.nrsb && c < o.nchunk && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(stage0 + (size_t)b * p.max_blk, o.W + (size_t)rsb * o.rsb_stride + (size_t)c * o.blk_bytes, (uint32_t)o.blk_bytes, mbar + b);
        }
    };
    int nload[2] = {0, 0};                                         // bulk copies waited for per stage (mbarrier phase)
    const int c_slot = rank * WPC + warp;                          // this warp's chunk slot: chunks c_slot * bpw .. + bpw - 1
    // op i's FIRST chunk always goes to stage (i & 1); further chunks of a multi-chunk op reuse the same stage
    request(0, c_slot * p.ops[0].bpw, 0);

    for (int op = 0; op < p.nops; ++op) {
        const ChainOp &o = p.ops[op];
        const int nchunk = o.nchunk, bpw = o.bpw, zp = o.zp, sd = o.sd, one_scale = o.one_scale, blk = o.blk_bytes;
        const float scale0 = o.scale0;
        const float *x = o.x;
        const bool active = rsb < o.nrsb;
        const int c_first = c_slot * bpw, c_end = min(nchunk, c_first + bpw);
        const int sb = op & 1;                                     // this op's stage
        long long *tr = (p.trace && tid == 0) ? p.trace + ((size_t)op * gridDim.x + blockIdx.x) * 16 : nullptr;
        if (tr) tr[0] = chain_now();
        // the weight stream of the NEXT op starts now (other stage; its previous tenant, op - 1, is done): weights do not
        // depend on activations, so this crosses the barrier below
        if ((p.flags & 1) && op + 1 < p.nops) request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1);
        const bool flow = (p.flags & 32) != 0;                     // data-flow mode: no grid barrier, inputs arrive as {value, epoch} words
        const uint32_t ep_in = epoch0 * (uint32_t)p.nops + (uint32_t)o.in_op + 1u, ep_out = epoch0 * (uint32_t)p.nops + (uint32_t)op + 1u;
        if (op > 0 && !flow) {
            // ---- grid barrier: every cluster has stored op-1's rows (release-increment by its leader).  ONE thread per cluster
            //      polls the counter (688 pollers on one line cost more than the barrier); it releases the cluster's CTAs through
            //      a generation word in each CTA's shared memory (DSMEM store), on which the CTAs spin locally ----
            if (tid == 0) {
                int spins = 0;
                if (rank == 0 || !(p.flags & 2)) {
                    const unsigned target = bar_base + (unsigned)op * (unsigned)nclusters;
                    while ((int)(ld_acquire_u32(p.bar) - target) < 0 && ++spins < kChainSpin) { }
                    if (spins >= kChainSpin) atomicExch(p.err, 1);
                    if (!(p.flags & 2)) *go = op;
                    else for (int r = 0; r < CS; ++r) {
                        uint32_t laddr = (uint32_t)__cvta_generic_to_shared(const_cast<int *>(go)), raddr;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(r));
                        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(raddr), "r"(op) : "memory");
                    }
                }
                while (*go < op && ++spins < kChainSpin) { }
                if (spins >= kChainSpin) atomicExch(p.err, 4);
            }
            if (tr) tr[1] = chain_now();
            __syncthreads();
        }
        if (tr) tr[2] = chain_now();
        if ((p.flags & 4) && op + 1 < p.nops) request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1);
        float cacc[RW];
        int iacc[RW];
        bool requested = false;
#pragma unroll
        for (int i = 0; i < RW; ++i) { cacc[i] = 0.f; iacc[i] = 0; }
        if (active)
            for (int c = c_first; c < c_end; ++c) {
                // ---- LUT slice of this chunk -> warp-private table (lane = group); arithmetic = lut_ctor.cc:119-215 / :242-256 ----
                float lsv[NAG], lbsum = 0.f;
                {
                    float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
                    if (flow && o.ll_in) {                         // the producer's rows, as soon as they exist (warp-cooperative polling)
                        const uint2 *src = o.ll_in + ((size_t)c * NG + lane) * 4;
                        const uint4 v0 = seq_wait<4>(src, lane < NG, ep_in, p.err, 5), v1 = seq_wait<4>(src + 2, lane < NG, ep_in, p.err, 5);
                        if (lane < NG) { b0 = __uint_as_float(v0.x); b1 = __uint_as_float(v0.z); b2 = __uint_as_float(v1.x); b3 = __uint_as_float(v1.z); }
                    } else if (lane < NG) {
                        const float4 f = __ldcg(reinterpret_cast<const float4 *>(x + ((size_t)c * NG + lane) * 4));   // L2: written by other SMs
                        b0 = f.x; b1 = f.y; b2 = f.z; b3 = f.w;
                    }
                    float m = __fadd_rn(__fadd_rn(fabsf(b0), fabsf(b1)), __fadd_rn(fabsf(b2), fabsf(b3)));
#pragma unroll
                    for (int s = WL / 2; s >= 1; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
                    const float scale = __fdiv_rn(m, 127.0f);
                    const float ts = (scale != 0.0f) ? __fdiv_rn(1.0f, scale) : 0.0f;
                    float od[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int idx = 2 * e + 1;
                        float v = b0;
                        v = (idx & 2) ? __fadd_rn(v, b1) : __fsub_rn(v, b1);
                        v = (idx & 4) ? __fadd_rn(v, b2) : __fsub_rn(v, b2);
                        v = (idx & 8) ? __fadd_rn(v, b3) : __fsub_rn(v, b3);
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
#pragma unroll
                    for (int a = 0; a < NAG; ++a) {
                        lsv[a] = __shfl_sync(0xffffffffu, scale, a * WL);
                        float bias = 0.f;
#pragma unroll
                        for (int k = 0; k < WL / 8; ++k) bias = __fadd_rn(bias, __shfl_sync(0xffffffffu, v, a * WL + 8 * k));
                        lbsum += bias;
                    }
                }
                if (c > c_first) request(op, c, sb);               // later chunks of this op: same stage, reloaded
                if (tr) tr[3] = chain_now();
                if (!seq_chain_wait(mbar + sb, nload[sb] & 1)) atomicExch(p.err, 2);
                ++nload[sb];
                __syncwarp();
                if (tr) tr[4] = chain_now();
                if ((p.flags & 8) && c == c_first && op + 1 < p.nops) { request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1); requested = true; }
                const unsigned char *stage = stage0 + (size_t)sb * p.max_blk;
                const uint4 *wp = reinterpret_cast<const uint4 *>(stage) + lane;
                float facc[RW];
#pragma unroll
                for (int i = 0; i < RW; ++i) facc[i] = 0.f;
#pragma unroll
                for (int q = 0; q < QCH; ++q) {
                    const uint4 wq = wp[q * 32];
                    const uint4 a = reinterpret_cast<const uint4 *>(tab)[2 * q], b2 = reinterpret_cast<const uint4 *>(tab)[2 * q + 1];
                    const uint32_t t[8] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w};
                    Quad<PB, true>::run(wq, t, iacc, wtx, wty);
                    if (((q + 1) % AGQ) == 0) {
#pragma unroll
                        for (int i = 0; i < RW; ++i) { facc[i] = fmaf(lsv[q / AGQ], (float)iacc[i], facc[i]); iacc[i] = 0; }
                    }
                }
                {
                    const unsigned char *sp = stage + (size_t)QCH * 512;
#pragma unroll
                    for (int i = 0; i < RW; ++i) {
                        const float s = one_scale ? scale0 : load_scale(sp, sd, lane * RW + i);
                        float v = fmaf(0.5f * s, facc[i] + lbsum, cacc[i]);
                        if (zp) v = fmaf(load_scale(sp + (size_t)RSB * sd, sd, lane * RW + i), lbsum, v);
                        cacc[i] = v;
                    }
                }
                __syncwarp();                                      // stage / table are rewritten by the next chunk
            }
        (void)blk;
        if (tr) tr[5] = chain_now();
        if ((!(p.flags & 13) || ((p.flags & 8) && !requested)) && op + 1 < p.nops) request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1);

        // ---- CTA reduction (fixed warp order), cluster reduction through DSMEM (rank order), leader stores the rows ----
        {
            float *r = red + (size_t)warp * RSB + lane * RW;
#pragma unroll
            for (int i = 0; i < RW; ++i) r[i] = cacc[i];
        }
        __syncthreads();
        if (tr) tr[6] = chain_now();
        if (flow && op > 0) cluster_wait_acq();                        // the leader has read op-1's partials (its arrive below)
        for (int t = tid; t < RSB; t += WPC * 32) {
            float fsum = 0.f;
#pragma unroll
            for (int w = 0; w < WPC; ++w) fsum += red[(size_t)w * RSB + t];
            st_cluster_f32(cl + (size_t)rank * RSB + t, 0, fsum);
        }
        cluster_sync_all();
        if (tr) tr[7] = chain_now();
        if (rank == 0) {
            if (active)
                for (int t = tid; t < RSB; t += WPC * 32) {
                    float fsum = 0.f;
#pragma unroll
                    for (int k2 = 0; k2 < CS; ++k2) fsum += cl[(size_t)k2 * RSB + t];
                    const int row = rsb * RSB + t;
                    if (row < o.Mout) {
                        if (flow) seq_publish(o.ll_out + row, __float_as_uint(fsum), ep_out);
                        if (o.out_f16) reinterpret_cast<__half *>(o.C)[row] = __float2half_rn(fsum);
                        else reinterpret_cast<float *>(o.C)[row] = fsum;
                    }
                }
            if (!flow) {
                if (!(p.flags & 16)) __threadfence();              // (flag 16: rely on bar.sync + the cumulative release below)
                __syncthreads();
                if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.bar), "r"(1u) : "memory");
            }
            if (tr) tr[8] = chain_now();
        }
        if (flow) cluster_arrive();                                // split-phase: waited for before the next op's partials are written
    }
    // the last increments must be in before the next launch computes its base from its own epoch; every CTA passes one final
    // barrier so that no cluster of a later launch can see a counter that is still being incremented by this one
    if (tid == 0) {
        if (!(p.flags & 32)) {
            const unsigned target = bar_base + (unsigned)p.nops * (unsigned)nclusters;
            int spins = 0;
            while ((int)(ld_acquire_u32(p.bar) - target) < 0 && ++spins < kChainSpin) { }
            if (spins >= kChainSpin) atomicExch(p.err, 3);
        }
        p.epochs[blockIdx.x] = epoch0 + 1u;
    }
    if (p.flags & 32) cluster_wait_acq();                              // completes the last split-phase barrier
    cluster_sync_all();                                            // nobody leaves while a peer may still write its shared memory
}

typedef void (*chain_fn)(const ChainParams, const uint32_t, const uint32_t);
chain_fn pick_chain(int pb, int qch, int agq);

}  // namespace tmac_b200
