// tmac_gemv4.cu -- instantiations of gemv4_kernel (tmac_gemv4.cuh), kept in their own translation unit so that the
// two halves of the library compile in parallel.
#include "tmac_gemv4.cuh"

namespace tmac_b200 {

namespace {
template <int PB, bool SYM> gemv4_fn pick4_qa(int qch, int agq) {
    switch (qch * 16 + agq) {
        case 8 * 16 + 8: return gemv4_kernel<PB, SYM, 8, 8, false>;
        case 8 * 16 + 4: return gemv4_kernel<PB, SYM, 8, 4, false>;
        case 8 * 16 + 2: return gemv4_kernel<PB, SYM, 8, 2, false>;
        case 8 * 16 + 0: return gemv4_kernel<PB, SYM, 8, 0, false>;
        case 4 * 16 + 4: return gemv4_kernel<PB, SYM, 4, 4, false>;
        case 4 * 16 + 2: return gemv4_kernel<PB, SYM, 4, 2, false>;
        case 4 * 16 + 0: return gemv4_kernel<PB, SYM, 4, 0, false>;
    }
    return nullptr;
}
template <int PB> gemv4_fn pick4_fused(int qch, int agq) {
    switch (qch * 16 + agq) {
        case 8 * 16 + 8: return gemv4_kernel<PB, true, 8, 8, true>;
        case 8 * 16 + 4: return gemv4_kernel<PB, true, 8, 4, true>;
        case 8 * 16 + 2: return gemv4_kernel<PB, true, 8, 2, true>;
        case 4 * 16 + 4: return gemv4_kernel<PB, true, 4, 4, true>;
        case 4 * 16 + 2: return gemv4_kernel<PB, true, 4, 2, true>;
    }
    return nullptr;
}
}  // namespace

gemv4_fn pick_gemv4(int pb, bool sym, int qch, int agq, bool fused) {
    if (fused) {
        if (pb == 4) return pick4_fused<4>(qch, agq);
        if (pb == 2) return pick4_fused<2>(qch, agq);
        if (pb == 1) return pick4_fused<1>(qch, agq);
        return nullptr;
    }
    if (pb == 4) return sym ? pick4_qa<4, true>(qch, agq) : pick4_qa<4, false>(qch, agq);
    if (pb == 2) return sym ? pick4_qa<2, true>(qch, agq) : pick4_qa<2, false>(qch, agq);
    if (pb == 1) return sym ? pick4_qa<1, true>(qch, agq) : pick4_qa<1, false>(qch, agq);
    return nullptr;
}

}  // namespace tmac_b200
