// ggml-tmac.cpp -- B200 replacement for 3rdparty/llama.cpp/ggml/src/ggml-tmac.cpp of the reference.
//
// The reference file implements the ten hooks of ggml/include/ggml-tmac.h:25-38 on top of TMAC::TMACGeMMWrapper and the
// TVM-generated CPU kernels.  Six of them (init, free, mul_mat_task_init, mul_mat_task_compute, set_n_threads,
// get_type_bits) are exported by libtmac_b200.so itself with the reference's signatures; this file supplies the four that
// take `ggml_tensor *` by forwarding the tensor fields to the ggml-free entry points of the library.  It is compiled against
// the reference's own ggml.h / ggml-tmac.h (tests/test_ggml_shim.py does exactly that) and contains no arithmetic.
//
// Build: add this file instead of ggml/src/ggml-tmac.cpp, link libtmac_b200.so (CMake package TMAC, target t_mac).
#include "ggml-tmac.h"

#define TMAC_B200_NO_GGML_DECLS   // the six same-named hooks are declared by ggml-tmac.h (enum ggml_type in the prototype)
#include "tmac_b200.h"

#include <mutex>

static_assert(sizeof(tmac_tensor_extra) == sizeof(tmac_tensor_extra_b200), "tmac_tensor_extra layout (ggml-tmac.h:17-23)");
static_assert(sizeof(tmac_float_type) == sizeof(float), "the B200 library serves the x86 float build of the hook (ggml-tmac.h:10)");

#ifndef GGML_TMAC_MAX_NODES
#define GGML_TMAC_MAX_NODES 8192        // ggml-tmac.cpp:20
#endif

extern "C" {

// ggml-tmac.cpp:238-248
bool ggml_tmac_can_mul_mat(const struct ggml_tensor * src0, const struct ggml_tensor * src1, const struct ggml_tensor * dst) {
    return ggml_tmac_b200_can_mul_mat((int) src0->type, src1->type == GGML_TYPE_F32, dst->type == GGML_TYPE_F32, src0->name) != 0;
}

// ggml-tmac.cpp:250-265
size_t ggml_tmac_mul_mat_get_wsize(const struct ggml_tensor * src0, const struct ggml_tensor * src1, const struct ggml_tensor * dst) {
    (void) dst;
    return ggml_tmac_b200_mul_mat_get_wsize((int) src0->ne[1], (int) src1->ne[0], (int) src1->ne[1], ggml_tmac_get_type_bits(src0->type));
}

// ggml-tmac.cpp:277-288
size_t ggml_tmac_get_nbytes(const struct ggml_tensor * tensor) {
    return ggml_tmac_b200_get_nbytes((int) tensor->ne[0], (int) tensor->ne[1], ggml_tmac_get_type_bits(tensor->type));
}

// ggml-tmac.cpp:290-501: called once per quantised tensor at load time (src/llama.cpp:5214-5217).  The library re-permutes the
// tensor into its stream layout, uploads it and fills the extra; `extra->qweights` stays the key ggml.c adds tile offsets to.
void ggml_tmac_transform_tensor(struct ggml_tensor * tensor) {
    if (tensor->extra != nullptr || !ggml_tmac_get_type_bits(tensor->type)) return;
    static tmac_tensor_extra_b200 extras[GGML_TMAC_MAX_NODES];
    static int n_extras = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (n_extras >= GGML_TMAC_MAX_NODES) return;
    if (ggml_tmac_b200_transform_tensor_typed(tensor->data, (int) tensor->type, (int) tensor->ne[0], (int) tensor->ne[1], &extras[n_extras]) > 0)
        tensor->extra = &extras[n_extras++];
}

}  // extern "C"
