// tmac_seq.cu -- instantiations of seq_kernel (tmac_seq.cuh), in their own translation unit so that the halves of the
// library compile in parallel.
#include "tmac_seq.cuh"
#include "tmac_chain.cuh"

namespace tmac_b200 {

namespace {
template <int PB> seq_fn pick_seq_qa(int qch, int agq) {
    switch (qch * 16 + agq) {
        case 8 * 16 + 8: return seq_kernel<PB, 8, 8>;
        case 8 * 16 + 4: return seq_kernel<PB, 8, 4>;
        case 8 * 16 + 2: return seq_kernel<PB, 8, 2>;
        case 4 * 16 + 4: return seq_kernel<PB, 4, 4>;
        case 4 * 16 + 2: return seq_kernel<PB, 4, 2>;
        case 2 * 16 + 2: return seq_kernel<PB, 2, 2>;
    }
    return nullptr;
}
}  // namespace

namespace {
template <int PB> chain_fn pick_chain_qa(int qch, int agq) {
    switch (qch * 16 + agq) {
        case 8 * 16 + 8: return chain_kernel<PB, 8, 8>;
        case 8 * 16 + 4: return chain_kernel<PB, 8, 4>;
        case 8 * 16 + 2: return chain_kernel<PB, 8, 2>;
        case 4 * 16 + 4: return chain_kernel<PB, 4, 4>;
        case 4 * 16 + 2: return chain_kernel<PB, 4, 2>;
        case 2 * 16 + 2: return chain_kernel<PB, 2, 2>;
    }
    return nullptr;
}
}  // namespace

chain_fn pick_chain(int pb, int qch, int agq) {
    if (pb == 4) return pick_chain_qa<4>(qch, agq);
    if (pb == 2) return pick_chain_qa<2>(qch, agq);
    if (pb == 1) return pick_chain_qa<1>(qch, agq);
    return nullptr;
}

seq_fn pick_seq(int pb, int qch, int agq) {
    if (pb == 4) return pick_seq_qa<4>(qch, agq);
    if (pb == 2) return pick_seq_qa<2>(qch, agq);
    if (pb == 1) return pick_seq_qa<1>(qch, agq);
    return nullptr;
}

}  // namespace tmac_b200
