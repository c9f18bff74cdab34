// examples/gguf_gemv.cc -- the reference's load-time + per-op sequence (3rdparty/llama.cpp/src/llama.cpp:5214-5217,
// ggml/src/ggml.c:12561-12707, ggml/src/ggml-tmac.cpp:267-354) written against libtmac_b200, without llama.cpp:
//   1. open the .gguf the reference pipeline produced, register the model's kcfg.ini;
//   2. upload one quantised linear (I1..I4 / Q4_0 / TQ1_0 / TQ2_0) -- what ggml_tmac_transform_tensor does at load;
//   3. per token: TMACGeMMWrapper::llama_cpp_init (activation -> LUT) and ::llama_cpp_compute (LUT GEMV) with HOST
//      buffers, exactly the calls ggml-tmac.cpp makes.
// Usage: gguf_gemv model.gguf kcfg.ini tensor-name
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "t-mac/tmac_gemm_wrapper.h"
#include "tmac_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s model.gguf kcfg.ini tensor-name\n", argv[0]); return 2; }
    const int64_t gg = tmac_b200_gguf_open(argv[1]);
    if (gg < 0) { std::fprintf(stderr, "%s\n", tmac_b200_last_error()); return 1; }
    char arch[64] = "?";
    tmac_b200_gguf_meta_string(gg, "general.architecture", arch, sizeof arch);
    std::printf("%s: %d tensors, architecture %s\n", argv[1], tmac_b200_gguf_tensor_count(gg), arch);

    const int idx = tmac_b200_gguf_find_tensor(gg, argv[3]);
    tmac_b200_gguf_tensor t;
    if (idx < 0 || tmac_b200_gguf_tensor_info(gg, idx, &t) != 0) { std::fprintf(stderr, "%s\n", tmac_b200_last_error()); return 1; }
    const int K = (int)t.ne[0], M = (int)t.ne[1], bits = ggml_tmac_get_type_bits(t.ggml_type);
    std::printf("%s: ggml type %d (%d bits), %d x %d, %llu bytes\n", t.name, t.ggml_type, bits, M, K, (unsigned long long)t.nbytes);
    if (!bits) { std::fprintf(stderr, "not a T-MAC tensor type\n"); return 1; }
    std::fflush(stdout);

    // the wrapper reads kcfg.ini like the reference's (tmac_gemm_wrapper.h:40-56, :230-255)
    TMAC::TMACGeMMWrapper<float> wrapper(1, 64, argv[2], "");
    tmac_tensor_extra_b200 extra;
    const int64_t h = tmac_b200_gguf_load_tensor(gg, idx, &extra);           // needs a B200: there is no CPU fallback
    if (h < 0) { std::fprintf(stderr, "upload failed: %s\n", tmac_b200_last_error()); tmac_b200_gguf_close(gg); return 1; }

    std::vector<float> x(K), y(M), lut_scales(extra.lut_scales_size), lut_biases(extra.lut_scales_size);
    std::vector<int8_t> qlut((size_t)K * 4);
    for (int k = 0; k < K; ++k) x[k] = (float)((k * 37) % 17 - 8) / 8.0f;
    wrapper.llama_cpp_init(x.data(), qlut.data(), lut_scales.data(), lut_biases.data(), M, K, 1, bits);
    wrapper.llama_cpp_compute(extra.qweights, extra.scales, qlut.data(), lut_scales.data(), lut_biases.data(), y.data(), M, K, 1, bits);
    std::printf("y[0..3] = %g %g %g %g\n", y[0], y[1], y[2], y[3]);
    tmac_b200_free_weights(h);
    tmac_b200_gguf_close(gg);
    return 0;
}
