// tmac_chain.cuh -- the dependent chain of GEMVs as ONE persistent launch, second form: gemv3's decomposition kept resident.
//
// tmac_seq.cuh (stream-K shares, partial sums and LUT records exchanged through HBM words) turned out latency-bound at the
// granularity of one 12.7 MB GEMV: three CTA-wide phases and two inter-CTA hops per op.  This kernel
// keeps what measured well in the launch chain -- gemv3's decomposition: cluster of 8 CTAs x 4 warps per 128-row super-block,
// one (super-block, chunk) block per warp, LUT slice built by the warp itself, K slices summed in the leader's shared memory
// in rank order -- and removes what the launch chain loses per launch:
//   * the kernel boundary becomes DATA FLOW: the leader of each cluster writes its rows as {value, epoch} words next to C, and a
//     warp of the next op waits for exactly the 128 input values of its chunk (one producer cluster), not for a grid barrier;
//   * the cluster-wide rendezvous of the reduction becomes one DSMEM hop: peers send their partial sums with st.async, whose
//     completion bytes are counted on the leader's mbarrier (data and signal travel together); a split-phase cluster barrier,
//     waited for a whole op later, only guards the re-use of the buffers;
//   * the exposed weight stream: every warp requests its block of op i+1 (one cp.async.bulk, 4.6 KB, second stage buffer) when
//     it starts op i's lookups -- weights do not depend on activations -- so the stream overlaps the ALU-bound phase and is out
//     of the way of the latency-critical words (requesting it earlier, across the hand-over, measured 1.3 us per op SLOWER:
//     12.7 MB of bulk traffic queue in front of the words everybody waits for);
//   * CTA placement is fixed for the whole chain (688 CTAs of 128 threads, all resident: 4 or 5 per SM).
// Measured on the bench chain (32 x 11008x4096 W2, x[i+1] = first K outputs of op i): 5.05 us per GEMV against 6.35 for the
// launch chain and 8.0 for tmac_seq.cuh; flags bit 0 selects the grid-barrier form this kernel started as (7.0 us), kept for
// comparison.  Per-phase stamps (ChainParams::trace) are read by tools/seq_bench.py --trace.
// Same arithmetic as gemv3's fused path: LUT bytes identical to the preprocessor, fp32 sums in fixed (warp, then cluster rank)
// order -> bit-identical to tmac_b200_gemv for every op.  Waits are bounded (error flag).
#pragma once
#include "tmac_kernels.cuh"
#include "tmac_seq.cuh"     // {value, epoch} words: seq_publish

namespace tmac_b200 {

constexpr int kChainCS = 8;                 // CTAs per cluster (K slices of one row super-block)
constexpr int kChainWarps = 4;              // warps per CTA
constexpr int kChainSpin = 1 << 22;
constexpr int kChainBarrier = 1;            // ChainParams::flags bit 0: grid barrier between ops instead of data flow

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
    uint2 *ll_out;                          // [rows] {value bits, epoch} words written next to C
    const uint2 *ll_in;                     // the producer's words, already offset (NULL: external input, plain loads)
    // multi-GPU row sharding (as Gemv3Params::Cpeer): the finished rows are also stored straight into the peers' output vectors
    int npeer, pad_;
    void *Cpeer[7];
};

struct ChainParams {
    const ChainOp *ops;
    int nops;
    int max_blk;                            // stage bytes per warp
    unsigned int *bar;                      // grid barrier counter (barrier form; monotonic over launches)
    unsigned int *epochs;                   // [grid] launches seen by each CTA
    int *err;
    int flags;                              // kChainBarrier
    long long *trace;                       // debug: [nops][grid][16] globaltimer stamps of thread 0, or NULL
};

// The split-phase barrier of the data-flow form guards only the RE-USE of shared-memory buffers (red, cl): the readers have
// consumed their values before they arrive, the writers write after they have waited.  It carries no data, so it needs neither
// release on the arrive side -- which ptxas turns into MEMBAR.ALL.GPU per thread and op (ncu: 15 % of the warp stall time of the
// first version was `membar`; 5.40 -> 5.05 us per GEMV).  (barrier.cluster.wait has only the acquire form.)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acq() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
// partial sum -> the leader's shared memory, completion counted on the leader's mbarrier
__device__ __forceinline__ void st_async_f32(float *local_ptr, uint64_t *local_bar, uint32_t rank, float v) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb) : "memory");
}
// Four consecutive {value, epoch} words (32 bytes) per active lane.  Every lane loads its words once; if some are stale, the lowest
// stale lane alone polls its first word (the rest of the warp would only add L2 traffic), then the stale lanes reload both
// 16-byte halves together: data that is already there costs one L2 round trip, data that arrives while waiting two.
__device__ __forceinline__ void chain_wait_x(const uint2 *src, bool active, uint32_t epoch, int *err, float &b0, float &b1, float &b2, float &b3) {
    uint4 v0 = make_uint4(0u, epoch, 0u, epoch), v1 = v0;
    auto load = [&]() {
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w) : "l"(src) : "memory");
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "l"(src + 2) : "memory");
    };
    if (active) load();
    int spins = 0;
    for (;;) {
        const bool stale = active && (v0.y != epoch || v0.w != epoch || v1.y != epoch || v1.w != epoch);
        const unsigned m = __ballot_sync(0xffffffffu, stale);
        if (m == 0) break;
        const int leader = __ffs(m) - 1;
        if ((int)(threadIdx.x & 31) == leader) {
            uint32_t e;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(reinterpret_cast<const uint32_t *>(src) + 1) : "memory"); } while (e != epoch && ++spins < kChainSpin);
        }
        spins = __shfl_sync(0xffffffffu, spins, leader);
        if (spins >= kChainSpin) { if ((threadIdx.x & 31) == 0) atomicExch(err, 5); v0.x = v0.z = v1.x = v1.z = 0u; break; }
        if (stale) load();
    }
    b0 = __uint_as_float(v0.x); b1 = __uint_as_float(v0.z); b2 = __uint_as_float(v1.x); b3 = __uint_as_float(v1.z);
}
__device__ __forceinline__ long long chain_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool chain_mbar_wait(uint64_t *b, uint32_t parity) {
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
    const bool flow = !(p.flags & kChainBarrier);
    // shared memory: cl [CS][RSB] (cluster partials, leader) | red [WPC][RSB] | per warp: 2 stages (max_blk) + table (NG * 8) |
    //                mbar [WPC][2] (bulk copies) | rmbar (leader: the peers' partial sums have landed)
    float *cl = reinterpret_cast<float *>(smem);
    float *red = cl + CS * RSB;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(red + WPC * RSB);
    const int per_warp = 2 * p.max_blk + NG * 8;
    unsigned char *stage0 = wbase + (size_t)warp * per_warp;
    unsigned char *tab = stage0 + 2 * p.max_blk;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbase + (size_t)WPC * per_warp) + 2 * warp;
    uint64_t *rmbar = reinterpret_cast<uint64_t *>(wbase + (size_t)WPC * per_warp) + 2 * WPC;
    int nred = 0;                                                  // leader: reductions waited for (mbarrier phase)
    const unsigned epoch0 = p.epochs[blockIdx.x];                 // launches before this one
    const unsigned bar_base = epoch0 * (unsigned)p.nops * (unsigned)nclusters;

    if (lane == 0) { mbar_init1(mbar); mbar_init1(mbar + 1); if (warp == 0) mbar_init1(rmbar); mbar_fence_init(); }
    __syncwarp();
    cluster_sync_all();                                            // every CTA of the cluster runs: its shared memory may be written

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
        const int nchunk = o.nchunk, bpw = o.bpw, zp = o.zp, sd = o.sd, one_scale = o.one_scale;
        const float scale0 = o.scale0;
        const float *x = o.x;
        const bool active = rsb < o.nrsb;
        const int c_first = c_slot * bpw, c_end = min(nchunk, c_first + bpw);
        const int sb = op & 1;                                     // this op's stage
        const uint32_t ep_in = epoch0 * (uint32_t)p.nops + (uint32_t)o.in_op + 1u, ep_out = epoch0 * (uint32_t)p.nops + (uint32_t)op + 1u;
        long long *tr = (p.trace && tid == 0) ? p.trace + ((size_t)op * gridDim.x + blockIdx.x) * 16 : nullptr;
        if (tr) tr[0] = chain_now();
        if (op > 0 && !flow) {
            // ---- barrier form: every cluster has stored op-1's rows (release-increment by its leader); one polling thread per CTA ----
            if (tid == 0) {
                const unsigned target = bar_base + (unsigned)op * (unsigned)nclusters;
                int spins = 0;
                while ((int)(ld_acquire_u32(p.bar) - target) < 0 && ++spins < kChainSpin) { }
                if (spins >= kChainSpin) atomicExch(p.err, 1);
            }
            if (tr) tr[1] = chain_now();
            __syncthreads();
        }
        if (tr) tr[2] = chain_now();
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
                    if (flow && o.ll_in) {                         // the producer's rows, as soon as they exist
                        chain_wait_x(o.ll_in + ((size_t)c * NG + lane) * 4, lane < NG, ep_in, p.err, b0, b1, b2, b3);
                        if (lane >= NG) b0 = b1 = b2 = b3 = 0.f;
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
                if (!chain_mbar_wait(mbar + sb, nload[sb] & 1)) atomicExch(p.err, 2);
                ++nload[sb];
                __syncwarp();
                if (tr) tr[4] = chain_now();
                // the weight stream of the NEXT op starts with this op's lookups (other stage: its tenant, op - 1, is done)
                if (c == c_first && op + 1 < p.nops) { request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1); requested = true; }
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
        if (tr) tr[5] = chain_now();
        if (!requested && op + 1 < p.nops) request(op + 1, c_slot * p.ops[op + 1].bpw, sb ^ 1);   // warps without a block in this op

        // ---- CTA reduction (fixed warp order), cluster reduction in the leader's shared memory (rank order), leader stores the rows ----
        if (flow && op > 0) cluster_wait_acq();                    // every thread of the cluster is past op-1's reduction: red / cl may be rewritten
        {
            float *r = red + (size_t)warp * RSB + lane * RW;
#pragma unroll
            for (int i = 0; i < RW; ++i) r[i] = cacc[i];
        }
        __syncthreads();
        if (tr) tr[6] = chain_now();
        if (flow) {
            // peers send their partial with st.async (completion bytes counted on the leader's mbarrier); the leader sums in rank
            // order once (CS-1) x RSB floats have landed: one DSMEM hop instead of a cluster-wide rendezvous
            if (active) {
                if (rank == 0 && tid == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(rmbar)), "r"((uint32_t)((CS - 1) * RSB * 4)) : "memory");
                for (int t = tid; t < RSB; t += WPC * 32) {
                    float fsum = 0.f;
#pragma unroll
                    for (int w = 0; w < WPC; ++w) fsum += red[(size_t)w * RSB + t];
                    if (rank == 0) cl[t] = fsum;                   // read back by the same thread below
                    else st_async_f32(cl + (size_t)rank * RSB + t, rmbar, 0, fsum);
                }
                if (tr) tr[7] = chain_now();
                if (rank == 0) {
                    if (!chain_mbar_wait(rmbar, nred & 1)) atomicExch(p.err, 6);
                    ++nred;
                    for (int t = tid; t < RSB; t += WPC * 32) {
                        float fsum = 0.f;
#pragma unroll
                        for (int k2 = 0; k2 < CS; ++k2) fsum += cl[(size_t)k2 * RSB + t];
                        const int row = rsb * RSB + t;
                        if (row < o.Mout) {
                            seq_publish(o.ll_out + row, __float_as_uint(fsum), ep_out);
                            if (o.out_f16) reinterpret_cast<__half *>(o.C)[row] = __float2half_rn(fsum);
                            else reinterpret_cast<float *>(o.C)[row] = fsum;
                            for (int q = 0; q < o.npeer; ++q) {      // the all-gather, fused into the epilogue: one store per peer and row
                                if (o.out_f16) reinterpret_cast<__half *>(o.Cpeer[q])[row] = __float2half_rn(fsum);
                                else reinterpret_cast<float *>(o.Cpeer[q])[row] = fsum;
                            }
                        }
                    }
                    if (tr) tr[8] = chain_now();
                }
            }
            cluster_arrive();                                      // split-phase: waited for before the next op's partials are written
        } else {
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
                            if (o.out_f16) reinterpret_cast<__half *>(o.C)[row] = __float2half_rn(fsum);
                            else reinterpret_cast<float *>(o.C)[row] = fsum;
                            for (int q = 0; q < o.npeer; ++q) {
                                if (o.out_f16) reinterpret_cast<__half *>(o.Cpeer[q])[row] = __float2half_rn(fsum);
                                else reinterpret_cast<float *>(o.Cpeer[q])[row] = fsum;
                            }
                        }
                    }
                __syncthreads();                                   // the release below is cumulative over the CTA's stores
                if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.bar), "r"(1u) : "memory");
                if (tr) tr[8] = chain_now();
            }
        }
    }
    if (tid == 0) {
        if (!flow) {
            // the last increments must be in before the next launch computes its base from its own epoch: every CTA passes one final
            // barrier so that no cluster of a later launch can see a counter that is still being incremented by this one
            const unsigned target = bar_base + (unsigned)p.nops * (unsigned)nclusters;
            int spins = 0;
            while ((int)(ld_acquire_u32(p.bar) - target) < 0 && ++spins < kChainSpin) { }
            if (spins >= kChainSpin) atomicExch(p.err, 3);
        }
        p.epochs[blockIdx.x] = epoch0 + 1u;
    }
    if (flow) cluster_wait_acq();                                  // completes the last split-phase barrier
    cluster_sync_all();                                            // nobody leaves while a peer may still write its shared memory
}

typedef void (*chain_fn)(const ChainParams, const uint32_t, const uint32_t);
chain_fn pick_chain(int pb, int qch, int agq);

}  // namespace tmac_b200
