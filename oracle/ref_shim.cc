// ref_shim.cc -- TEST INFRASTRUCTURE ONLY (builds oracle/_ref/libtmac_ref.so).
//
// Compiles the REFERENCE's own SIMD intrinsics where they lie
// (/root/reference/python/t_mac/intrins/{tbl,lut_ctor}.cc, included via -I, never
// copied) and wraps them in the outer loops the reference's TVM code generator
// would emit (deploy/tuned/aarch64-llama-2-7b-2bit/kernels.cc:1059-1075,
// deploy/tuned/kernels.cc:1032-1038).  The checked-in generated kernels.cc are
// aarch64 builds (typedef _Float16 half) and cannot be compiled for x86, so the
// ~40 lines of loop nest are restated here; every arithmetic instruction executed
// is the reference's.  Exposes the same driver signatures as tmac_oracle.c
// (prefix tmr_ instead of tmo_) plus a multi-threaded GEMV driver that mimics
// ggml's tile work stealing (3rdparty/llama.cpp/ggml/src/ggml.c:12632-12703) for
// the CPU baseline in bench.py.
#include "tbl.cc"
#include "lut_ctor.cc"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define TMR_API extern "C" __attribute__((visibility("default")))

namespace {

typedef int32_t (*float_update_fn)(int32_t, float_type *, int8_t *, uint8_t *, float_type *,
                                   float_type *, float_type *);
typedef int32_t (*int32_update_fn)(int32_t, int32_t *, int8_t *, uint8_t *);

template <int KF, int BITS, int AK, bool ZP>
int32_t fupd(int32_t m, float_type *c, int8_t *lut, uint8_t *a, float_type *s, float_type *ls,
             float_type *lb) {
    return tbl_g4_int8_float_update_impl<true, KF, BITS, AK, false, ZP, false>(m, c, lut, a, s, ls, lb);
}
template <int KF, int BITS>
int32_t iupd(int32_t m, int32_t *c, int8_t *lut, uint8_t *a) {
    return tbl_g4_int8_int32_update_impl<KF, BITS>(m, c, lut, a);
}

template <int KF, int AK, bool ZP>
float_update_fn pick_bits_f(int bits) {
    switch (bits) {
        case 1: return fupd<KF, 1, AK, ZP>;
        case 2: return fupd<KF, 2, AK, ZP>;
        case 3: return fupd<KF, 3, AK, ZP>;
        case 4: return fupd<KF, 4, AK, ZP>;
    }
    return nullptr;
}
float_update_fn pick_float(int kf, int bits, int actk, bool zp) {
    if (kf == 16 && actk == 16) return zp ? pick_bits_f<16, 16, true>(bits) : pick_bits_f<16, 16, false>(bits);
    if (kf == 16 && actk == 8) return zp ? pick_bits_f<16, 8, true>(bits) : pick_bits_f<16, 8, false>(bits);
    if (kf == 8 && actk == 8) return zp ? pick_bits_f<8, 8, true>(bits) : pick_bits_f<8, 8, false>(bits);
    return nullptr;
}
int32_update_fn pick_int32(int kf, int bits) {
    if (kf == 16) switch (bits) {
            case 1: return iupd<16, 1>;
            case 2: return iupd<16, 2>;
            case 3: return iupd<16, 3>;
            case 4: return iupd<16, 4>;
        }
    if (kf == 8) switch (bits) {
            case 1: return iupd<8, 1>;
            case 2: return iupd<8, 2>;
            case 3: return iupd<8, 3>;
            case 4: return iupd<8, 4>;
        }
    return nullptr;
}

const float kAlphas[4] = {0.5f, 1.0f, 2.0f, 4.0f};

struct Problem {
    int Mout, K, bits, bm, kf, gs, ags, zp, mode;
    const uint8_t *A;
    const int8_t *LUT;
    const float *Scales, *LUT_Scales, *LUT_Biases;
    float *C;
    float_update_fn fu;
    int32_update_fn iu;
};

// One tile = bm plane rows: the body of a generated qgemm_lut_t1_int8_m{bm}_k{K}_n1_b{bits}.
int run_tile(const Problem &p, int t) {
    const int bm = p.bm, K = p.K, bits = p.bits, kf = p.kf;
    const int rows = bm / bits;
    const size_t a_tile = (size_t)K / 4 * bm / 2;
    uint8_t *A = const_cast<uint8_t *>(p.A) + t * a_tile;
    int8_t *LUT = const_cast<int8_t *>(p.LUT);
    float *C = p.C + (size_t)t * rows;
    if (p.mode == 0) {
        const int srow = rows * (p.zp ? 2 : 1);
        float *S = const_cast<float *>(p.Scales) + (size_t)t * (K / p.gs) * srow;
        alignas(32) float cbits[2048];
        if (bm > 2048) return -1;
        tbl_float_reset(bm, cbits);
        for (int ko = 0; ko < K / 4 / kf; ++ko)
            p.fu(bm, cbits, LUT + (size_t)ko * kf * 16, A + (size_t)ko * kf * bm / 2,
                 S + (size_t)(ko * kf * 4 / p.gs) * srow,
                 const_cast<float *>(p.LUT_Scales) + ko * kf * 4 / p.ags,
                 const_cast<float *>(p.LUT_Biases) + ko * kf * 4 / p.ags);
        for (int i = 0; i < rows; ++i) {
            float acc = 0.0f;
            for (int b = 0; b < bits; ++b) acc = acc + cbits[(i / 8) * 8 * bits + i % 8 + 8 * b] * kAlphas[b];
            C[i] = acc;
        }
    } else {
        alignas(32) int32_t cbits[2048];
        if (bm > 2048) return -1;
        tbl_int32_reset(bm, cbits);
        for (int ko = 0; ko < K / 4 / kf; ++ko)
            p.iu(bm, cbits, LUT + (size_t)ko * kf * 16, A + (size_t)ko * kf * bm / 2);
        for (int i = 0; i < rows; ++i) {
            float acc = 0.0f;
            for (int b = 0; b < bits; ++b)
                acc = acc + (float)cbits[(i / 8) * 8 * bits + i % 8 + 8 * b] * kAlphas[b];
            const float t1 = acc * p.LUT_Scales[0];
            const float t2 = p.LUT_Biases[0] * kAlphas[0];
            C[i] = (t1 + t2) * p.Scales[0];
        }
    }
    return 0;
}

int fill_problem(Problem &p) {
    const int M = p.Mout * p.bits;
    if (p.bm <= 0 || M % p.bm || p.bm % 32 || p.bm % p.bits || p.K % (4 * p.kf)) return -1;
    if (p.mode == 0) {
        const int actk = (p.ags / 4 < p.kf) ? p.ags / 4 : p.kf;
        p.fu = pick_float(p.kf, p.bits, actk, p.zp != 0);
        if (!p.fu) return -1;
    } else {
        p.iu = pick_int32(p.kf, p.bits);
        if (!p.iu) return -1;
    }
    return 0;
}

// Persistent worker pool: ggml-style atomic chunk stealing over weight tiles.
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int i = 1; i < n_; ++i) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return n_; }
    void run(const Problem &p) {
        prob_ = &p;
        ntile_ = p.Mout * p.bits / p.bm;
        next_.store(0, std::memory_order_relaxed);
        pending_.store(n_ - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(m_);
            ++gen_;
        }
        cv_.notify_all();
        work();
        while (pending_.load(std::memory_order_acquire) != 0) {
        }
    }

private:
    void work() {
        for (;;) {
            int t = next_.fetch_add(1, std::memory_order_relaxed);
            if (t >= ntile_) break;
            run_tile(*prob_, t);
        }
    }
    void loop() {
        unsigned seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    unsigned gen_ = 0;
    bool stop_ = false;
    const Problem *prob_ = nullptr;
    int ntile_ = 0;
    std::atomic<int> next_{0}, pending_{0};
};

Pool *g_pool = nullptr;

}  // namespace

TMR_API int tmr_partial_max_reset(float *ls) { return partial_max_reset(ls); }
TMR_API int tmr_partial_max_g4_int8_k8(float *ls, const float *b) {
    return partial_max_g4_int8_k8(ls, const_cast<float *>(b));
}
TMR_API int tmr_lut_ctor_g4_int8(int act_k, int8_t *qlut, const float *b, float *ls, float *lb) {
    // FastAggregationK / Bits template arguments do not change the arithmetic (lut_ctor.cc:36-37).
    return lut_ctor_g4_int8_impl<0, 4>(act_k, qlut, const_cast<float *>(b), ls, lb);
}

TMR_API int tmr_preprocessor(int K, int N, int ags, const float *B, float *LUT_Scales,
                             float *LUT_Biases, int8_t *QLUT) {
    if (ags <= 0 || ags % 32 || K % ags) return -1;
    const int nag = K / ags;
    for (int n = 0; n < N; ++n) {
        float *b = const_cast<float *>(B) + (size_t)n * K;
        float *ls = LUT_Scales + (size_t)n * nag, *lb = LUT_Biases + (size_t)n * nag;
        int8_t *q = QLUT + (size_t)n * K * 4;
        for (int kk = 0; kk < nag; ++kk) {
            partial_max_reset(ls + kk);
            for (int ko = 0; ko < ags / 32; ++ko) partial_max_g4_int8_k8(ls + kk, b + kk * ags + ko * 32);
        }
        for (int kk = 0; kk < nag; ++kk)
            lut_ctor_g4_int8_impl<0, 4>(ags, q + (size_t)kk * (ags / 4) * 16, b + kk * ags, ls + kk, lb + kk);
    }
    return 0;
}

TMR_API int tmr_qgemm(int Mout, int K, int N, int bits, int bm, int kf, int gs, int ags, int zero_point,
                      int mode, const uint8_t *A, const int8_t *LUT, const float *Scales,
                      const float *LUT_Scales, const float *LUT_Biases, float *C) {
    Problem p{Mout, K, bits, bm, kf, gs, ags, zero_point, mode, A, LUT, Scales, LUT_Scales, LUT_Biases, C,
              nullptr, nullptr};
    if (fill_problem(p)) return -1;
    const int nag = K / ags;
    for (int n = 0; n < N; ++n) {
        Problem q = p;
        q.LUT = LUT + (size_t)n * K * 4;
        q.LUT_Scales = LUT_Scales + (size_t)n * nag;
        q.LUT_Biases = LUT_Biases + (size_t)n * nag;
        q.C = C + (size_t)n * Mout;
        for (int t = 0; t < Mout * bits / bm; ++t)
            if (run_tile(q, t)) return -1;
    }
    return 0;
}

TMR_API int tmr_qgemm_cbits(int Mout, int K, int N, int bits, int bm, int kf, const uint8_t *A,
                            const int8_t *LUT, int32_t *cbits) {
    const int M = Mout * bits;
    int32_update_fn iu = pick_int32(kf, bits);
    if (!iu || bm <= 0 || M % bm || bm % 32 || K % (4 * kf)) return -1;
    const size_t a_tile = (size_t)K / 4 * bm / 2;
    memset(cbits, 0, (size_t)N * M * sizeof(int32_t));
    for (int n = 0; n < N; ++n)
        for (int t = 0; t < M / bm; ++t)
            for (int ko = 0; ko < K / 4 / kf; ++ko)
                iu(bm, cbits + (size_t)n * M + (size_t)t * bm,
                   const_cast<int8_t *>(LUT) + (size_t)n * K * 4 + (size_t)ko * kf * 16,
                   const_cast<uint8_t *>(A) + t * a_tile + (size_t)ko * kf * bm / 2);
    return 0;
}

// CPU baseline: preprocessor single-threaded (qgemm.py:470-474), then tiles stolen by
// `nthreads` workers; N activation rows looped like ggml.c:12680.
TMR_API int tmr_set_threads(int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (g_pool && g_pool->size() == nthreads) return nthreads;
    delete g_pool;
    g_pool = new Pool(nthreads);
    return nthreads;
}

TMR_API int tmr_gemv_mt(int Mout, int K, int N, int bits, int bm, int kf, int gs, int ags, int zero_point,
                        int mode, const uint8_t *A, const float *Scales, const float *B, int8_t *QLUT,
                        float *LUT_Scales, float *LUT_Biases, float *C) {
    if (!g_pool) tmr_set_threads(1);
    if (tmr_preprocessor(K, N, ags, B, LUT_Scales, LUT_Biases, QLUT)) return -1;
    Problem p{Mout, K, bits, bm, kf, gs, ags, zero_point, mode, A, QLUT, Scales, LUT_Scales, LUT_Biases, C,
              nullptr, nullptr};
    if (fill_problem(p)) return -1;
    const int nag = K / ags;
    for (int n = 0; n < N; ++n) {
        Problem q = p;
        q.LUT = QLUT + (size_t)n * K * 4;
        q.LUT_Scales = LUT_Scales + (size_t)n * nag;
        q.LUT_Biases = LUT_Biases + (size_t)n * nag;
        q.C = C + (size_t)n * Mout;
        g_pool->run(q);
    }
    return 0;
}
