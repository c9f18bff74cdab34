// t-mac/tmac_gemm_wrapper.h -- B200 drop-in for the reference's header-only host wrapper
// (reference: include/t-mac/tmac_gemm_wrapper.h:79-347).  Same namespace, class, method names,
// argument order and meaning; the generated kernel dispatchers it used to call
// (qgemm_lut_int8 / preprocessor_int8, deploy/tuned/<preset>/kernels.h:21-37) are now exported by
// libtmac_b200.so and launch sm_100a kernels.  Link with -ltmac_b200.
//
// Differences a maintainer must know (also in INTEGRATION.md):
//   * weights must be made resident once (ggml_tmac_b200_transform_tensor or
//     tmac_b200_upload_weights); llama_cpp_compute then accepts the same `A + tile offset` host
//     pointer the reference passes (ggml.c:12662-12691) or one whole-tensor call.
//   * T is float (the x86 reference type).  set_num_threads is accepted and ignored.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <string>

#include "tmac_b200.h"

namespace TMAC {

constexpr size_t kAllocAlignment = 64;   // tmac_gemm_wrapper.h:24

struct TMACGeMMConfig {                   // tmac_gemm_wrapper.h:26-35
    int bm;
    int simd_n_in;
    int simd_n_out;
    int kfactor;
    int group_size;
    int lut_scales_size;
    int scales_size;
    int n_tile_num;
};

inline std::string get_kcfg_file(const std::string &kcfg_file) {   // :40-56
    if (!kcfg_file.empty()) return kcfg_file;
    if (const char *e = getenv("TMAC_KCFG_FILE")) return e;
    return "";
}

template <typename T, int g = 4>
class TMACGeMMWrapper {
public:
    TMACGeMMWrapper(int n_threads, int act_group_size, const std::string &kcfg_file, const std::string & /*library_file*/)
        : _n_threads(n_threads), _act_group_size(act_group_size), _allocated(false) {
        static_assert(sizeof(T) == 4, "the B200 library mirrors the x86 reference: T = float");
        if (tmac_b200_init(-1) != 0) { std::fprintf(stderr, "TMACGeMMWrapper: %s\n", tmac_b200_last_error()); std::abort(); }
        const std::string f = get_kcfg_file(kcfg_file);
        if (!f.empty() && tmac_b200_load_kcfg_file(f.c_str()) < 0) {    // LOG(FATAL) in the reference, :49
            std::fprintf(stderr, "TMACGeMMWrapper: %s\n", tmac_b200_last_error());
            std::abort();
        }
    }
    TMACGeMMWrapper() : TMACGeMMWrapper(1, 32, "", "") {}

    void set_num_threads(int n_threads) { _n_threads = n_threads; }   // :102-112 (CPU pool size: no meaning here)

    // Activation (B): NxK.  :173-195
    void llama_cpp_init(void *B, void *qlut, void *lut_scales, void *lut_biases, int M, int K, int N, int bits) {
        const int ret = preprocessor_int8(M * bits, K, N, bits, B, lut_scales, lut_biases, qlut);
        if (ret != 0) std::fprintf(stderr, "error calling preprocessor (m=%d, k=%d, n=%d, b=%d): %s\n", M, K, N, bits, tmac_b200_last_error());
    }

    // Activation (B): NxK, Weights (A): MxK.  :200-228
    void llama_cpp_compute(void *A, void *scales, void *qlut, void *lut_scales, void *lut_biases, void *C, int M, int K, int N, int bits) {
        const int ret = qgemm_lut_int8(M * bits, K, N, bits, A, qlut, scales, lut_scales, lut_biases, C);
        if (ret != 0) std::fprintf(stderr, "error calling qgemm_lut (m=%d, k=%d, n=%d, b=%d): %s\n", M, K, N, bits, tmac_b200_last_error());
    }

    TMACGeMMConfig get_kcfg(int M, int K, int N, int bits) {          // :230-255
        tmac_b200_kcfg c;
        if (tmac_b200_find_kcfg(M * bits, K, bits, &c) != 0) return TMACGeMMConfig{0, 0, 0, 0, 0, 0, 0, 0};
        const int scales = c.one_scale ? 1 : c.M * (c.K / c.group_size) * (c.zero_point ? 2 : 1);
        return TMACGeMMConfig{c.bm, c.simd_n_in, c.simd_n_out, c.kfactor, c.group_size, N * c.K / c.act_group_size, scales, c.M * c.bits / c.bm};
    }

    // Should only be called in main thread.  :258-270 (the library owns device workspaces; the host
    // buffers are kept so that callers of the TVM-style run() path keep working)
    void set_workspace(int maxK, int maxN) {
        if (_allocated) return;
        if (posix_memalign(&_qlut, kAllocAlignment, (size_t)maxN * maxK / g * (1 << g))) _qlut = nullptr;
        if (posix_memalign(&_lut_scales, kAllocAlignment, (size_t)maxN * maxK / _act_group_size * sizeof(T))) _lut_scales = nullptr;
        if (posix_memalign(&_lut_biases, kAllocAlignment, (size_t)maxN * maxK / _act_group_size * sizeof(T))) _lut_biases = nullptr;
        _allocated = true;
    }

    ~TMACGeMMWrapper() {
        if (_allocated) { free(_qlut); free(_lut_scales); free(_lut_biases); }
    }

private:
    int _n_threads;
    int _act_group_size;
    void *_qlut = nullptr, *_lut_scales = nullptr, *_lut_biases = nullptr;
    bool _allocated;
};

}  // namespace TMAC
