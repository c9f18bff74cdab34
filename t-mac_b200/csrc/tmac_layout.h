// tmac_layout.h -- the B200 "stream layout" of a quantised weight tensor and the host code that
// builds it from (a) the reference run-time layout produced by python/t_mac/weights.py:5-88 /
// ggml_tmac_transform_tensor (3rdparty/llama.cpp/ggml/src/ggml-tmac.cpp:356-490) or (b) plain
// quantised weights.  Header-only, host side, no CUDA types.
//
// Why a new layout: the reference interleaves bit planes / rows for 128-bit NEON/AVX2 `tbl`
// lookups of one tile (bm plane rows x kfactor groups).  On B200 the lookup primitive is PRMT
// (4 byte lookups in an 8-byte register table per instruction) followed by DP4A, so the stream
// is organised as:
//
//   tensor   = [row super-block rsb][K chunk c] blocks, contiguous, 16-byte aligned
//   block    = weights  [quad q < QCH][lane l < 32][word k < 4]   (32-bit words, 512 B per quad)
//              scales   [lane][RW] (fp16 when every scale is fp16-representable, else fp32)
//              zeros    [lane][RW] (only with zero points)
//   word     = the 8 LUT indices of ONE K-group (4 consecutive K positions of one bit plane each)
//              for the RW = 8/PB rows owned by that lane, PB = bit planes per row in the word
//              (bits 1 -> PB 1, bits 2 -> PB 2, bits 3/4 -> PB 4; bits 3 pads a zero-weight plane)
//   nibble   = code: low 3 bits j select one of the 8 stored LUT entries, bit 3 = negate.
//              LUT[idx] = -LUT[15-idx] (lut_ctor.cc:153-155), so idx<8 -> (j=idx, neg=0) and
//              idx>=8 -> (j=(15-idx), neg=1), i.e. code = idx < 8 ? idx : idx ^ 7.
//   K chunk  = one weight-quantisation group (group_size K positions) so that a block carries the
//              scales it needs; one lane-word quad = 4 consecutive K-groups = 16 K positions.
//
// The j bits of nibble n of word k always describe (row i, plane b, group 4q+k) with
//   PB 4: n = 4i + b            (i < 2)
//   PB 2: n = 2i + b            (i < 4)
//   PB 1: n = i                 (i < 8)
// The neg bits are permuted inside a group pair (PB 2) / quad (PB 1) so that the four sign bits
// that belong to one DP4A land in one 16-bit half (see pack_quad below and the kernel).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

namespace tmac_b200 {

struct StreamLayout {
    int Mout = 0, K = 0, bits = 0;
    int pb = 0, rw = 0, rsb = 0, nrsb = 0;   // planes/word, rows/lane, rows/super-block, #super-blocks
    int ck = 0, qch = 0, nchunk = 0;         // K per chunk, quads per chunk, chunks
    int sd = 0;                              // bytes per scale (2 = fp16, 4 = fp32); 0 = no per-row scales
    int zp = 0, one_scale = 0;
    int group_size = 0, act_group_size = 0;
    size_t wbytes = 0, sbytes = 0, blk = 0, rsb_stride = 0, total = 0;
    float scale0 = 0.f;                      // the unified scale (one_scale)
};

inline int planes_per_word(int bits) { return bits == 1 ? 1 : (bits == 2 ? 2 : 4); }

// Chunking rule (see header comment).  Returns false when the shape cannot be chunked.
inline bool make_layout(int Mout, int K, int bits, int group_size, int act_group_size, int zero_point,
                        int one_scale, int scale_bytes, StreamLayout *L) {
    if (bits < 1 || bits > 4 || Mout <= 0 || K <= 0 || K % 32) return false;
    int ags = (act_group_size <= 0 || act_group_size > K) ? K : act_group_size;
    if (ags % 32 || K % ags) return false;
    int ck;
    if (!one_scale) {
        if (group_size <= 0 || K % group_size || group_size % 32) return false;
        ck = group_size;
        if (ck > 128) {                                       // keep blocks small: split big groups
            ck = 128;
            if (group_size % 128) return false;
        }
        if (ck % ags) return false;                           // act group inside weight group, qgemm.py:112-113
    } else {
        ck = 0;
        const int cands[3] = {128, 64, 32};
        for (int c : cands)
            if (K % c == 0 && (ags == K || c % ags == 0)) { ck = c; break; }
        if (!ck) return false;
    }
    L->Mout = Mout; L->K = K; L->bits = bits;
    L->pb = planes_per_word(bits);
    L->rw = 8 / L->pb;
    L->rsb = 32 * L->rw;
    L->nrsb = (Mout + L->rsb - 1) / L->rsb;
    L->ck = ck; L->qch = ck / 16; L->nchunk = K / ck;
    L->zp = (zero_point && !one_scale) ? 1 : 0;
    L->one_scale = one_scale ? 1 : 0;
    L->sd = one_scale ? 0 : scale_bytes;
    L->group_size = one_scale ? K : group_size;
    L->act_group_size = ags;
    L->wbytes = (size_t)L->qch * 512;
    L->sbytes = one_scale ? 0 : (size_t)L->rsb * L->sd * (L->zp ? 2 : 1);
    L->blk = L->wbytes + L->sbytes;
    L->rsb_stride = L->blk * L->nchunk;
    L->total = L->rsb_stride * L->nrsb;
    return true;
}

inline uint32_t idx_to_code(uint32_t idx) { return idx < 8 ? idx : (idx ^ 7u); }

// fp32 -> fp16 bits (round-to-nearest-even) and back; used to decide whether scales are
// losslessly storable as fp16.
inline uint16_t f32_to_f16_bits(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t half = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
        if (rem > mid || (rem == mid && (half & 1))) ++half;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) ++half;
    return (uint16_t)(sign | half);
}
inline float f16_bits_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (!man) x = sign;
        else {
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f; std::memcpy(&f, &x, 4);
    return f;
}
inline bool all_fp16_exact(const float *v, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (f16_bits_to_f32(f32_to_f16_bits(v[i])) != v[i]) return false;
    return true;
}

// Plain representation used between "decode reference layout" and "encode stream layout":
// idx [Mout][bits][K/4] 4-bit LUT indices (bit j = bit `b` of w[row][4g+j], weights.py:57-60),
// scales/zeros [Mout][K/group_size] fp32.
struct PlainWeights {
    int Mout = 0, K = 0, bits = 0;
    std::vector<uint8_t> idx;
    std::vector<float> scales, zeros;
    uint8_t at(int row, int b, int g) const { return idx[((size_t)row * bits + b) * (K / 4) + g]; }
};

inline void plain_from_w(const uint8_t *w, int Mout, int K, int bits, int row0, int rows, PlainWeights *P) {
    P->Mout = rows; P->K = K; P->bits = bits;
    const int KG = K / 4;
    P->idx.assign((size_t)rows * bits * KG, 0);
    for (int r = 0; r < rows; ++r) {
        const uint8_t *wr = w + (size_t)(row0 + r) * K;
        for (int g = 0; g < KG; ++g)
            for (int b = 0; b < bits; ++b) {
                uint32_t v = 0;
                for (int j = 0; j < 4; ++j) v |= ((wr[4 * g + j] >> b) & 1u) << j;
                P->idx[((size_t)r * bits + b) * KG + g] = (uint8_t)v;
            }
    }
    (void)Mout;
}

// ---- ggml block formats -> weight codes + per-block scales ---------------------------------------------
// The reference re-permutes Q4_0 / TQ1_0 / TQ2_0 tensors at load time by reading them element by element through
// "accessors" (3rdparty/llama.cpp/ggml/src/ggml-tmac.cpp:98-236, block structs ggml-common.h:145-148, :234-247):
// get_q(idx) -> code w in [0, 2^bits) with real value (w - 2^(bits-1)) * d, get_scale -> the block's fp16 d.
//   Q4_0  (type 2,  32 elems, 18 B: d | qs[16]):          w = nibble (low: elems 0-15, high: 16-31)
//   TQ1_0 (type 34, 256 elems, 54 B: qs[48] | qh[4] | d): base-3 digits, 5 per byte (4 in qh), w = trit + 1
//   TQ2_0 (type 35, 256 elems, 66 B: qs[64] | d):         2 bits per element, w = q + 1
enum { kGgmlQ4_0 = 2, kGgmlTQ1_0 = 34, kGgmlTQ2_0 = 35 };
inline int ggml_block_elems(int type) { return type == kGgmlQ4_0 ? 32 : (type == kGgmlTQ1_0 || type == kGgmlTQ2_0) ? 256 : 0; }
inline size_t ggml_block_bytes(int type) { return type == kGgmlQ4_0 ? 18 : type == kGgmlTQ1_0 ? 54 : type == kGgmlTQ2_0 ? 66 : 0; }
inline int ggml_block_bits(int type) { return type == kGgmlQ4_0 ? 4 : (type == kGgmlTQ1_0 || type == kGgmlTQ2_0) ? 2 : 0; }

inline uint8_t ggml_block_q(int type, const uint8_t *blk, int i) {
    if (type == kGgmlQ4_0) {                       // BlockQ40TypeAccessor::get_q, ggml-tmac.cpp:106-112
        const uint8_t *qs = blk + 2;
        return (uint8_t)((qs[i % 16] >> (i / 16 * 4)) & 15);
    }
    if (type == kGgmlTQ2_0) {                      // BlockTQ20TypeAccessor::get_q, :215-221
        const uint8_t *sq = blk + (i / 128) * 32;
        const int si = i % 128;
        return (uint8_t)(((sq[si % 32] >> (si / 32 * 2)) & 3) + 1);
    }
    // BlockTQ10TypeAccessor::get_q, :158-196: qs[0:32] 5 trits per byte, qs[32:48] 5 per byte, qh[0:4] 4 per byte
    static const uint8_t pow3[5] = {1, 3, 9, 27, 81};
    int byte, trit;
    if (i < 160) { byte = i % 32; trit = i / 32; }
    else if (i < 240) { byte = 32 + (i - 160) % 16; trit = (i - 160) / 16; }
    else { byte = 48 + (i - 240) % 4; trit = (i - 240) / 4; }
    const uint8_t cur = (uint8_t)(blk[byte] * pow3[trit]);
    return (uint8_t)((((uint16_t)cur * 3) >> 8) + 1);
}
inline float ggml_block_scale(int type, const uint8_t *blk) {
    const uint8_t *d = (type == kGgmlQ4_0) ? blk : (type == kGgmlTQ1_0 ? blk + 52 : blk + 64);
    return f16_bits_to_f32((uint16_t)(d[0] | (d[1] << 8)));
}
// data: [Mout][K / elems] blocks, row major (K % elems == 0).  w: [Mout][K] codes, scales: [Mout][K / elems].
inline bool decode_ggml_blocks(int type, const void *data, int Mout, int K, uint8_t *w, float *scales) {
    const int E = ggml_block_elems(type);
    const size_t B = ggml_block_bytes(type);
    if (!E || K % E) return false;
    const uint8_t *p = (const uint8_t *)data;
    for (int r = 0; r < Mout; ++r)
        for (int kb = 0; kb < K / E; ++kb) {
            const uint8_t *blk = p + ((size_t)r * (K / E) + kb) * B;
            for (int i = 0; i < E; ++i) w[(size_t)r * K + (size_t)kb * E + i] = ggml_block_q(type, blk, i);
            scales[(size_t)r * (K / E) + kb] = ggml_block_scale(type, blk);
        }
    return true;
}

// ---- GPTQ safetensors tensors -> codes + scales + biased zeros -------------------------------------------------------
// Restates unpack_gptqv2 (python/t_mac/model_utils.py:95-129), the converter's entry for GPTQ checkpoints
// (convert_hf_to_gguf.py:305-306): qweight int32 [K*bits/32][M] packs 32/bits consecutive K positions per word (LSB first),
// qzeros int32 [K/gs][M*bits/32] packs 32/bits consecutive output rows per word, scales fp16 [K/gs][M].
// Out: w [M][K] in [0, 2^bits), scales [M][K/gs], zeros [M][K/gs] = (z (+1 for AutoGPTQ v1) - 2^(bits-1)) * scale, the product
// rounded to fp16 like numpy's float16 multiply (:124-127).  bits must divide 32 (1, 2, 4).
inline bool unpack_gptq(const int32_t *qweight, const uint16_t *scales_f16, const int32_t *qzeros, int K, int M, int bits,
                        int group_size, bool gptq_v2, uint8_t *w, float *scales, float *zeros) {
    if (bits < 1 || 32 % bits || K <= 0 || M <= 0 || group_size <= 0 || K % group_size || K % (32 / bits) || M % (32 / bits)) return false;
    const int p = 32 / bits, NG = K / group_size;
    const uint32_t mask = (1u << bits) - 1u;
    for (int k = 0; k < K; ++k)
        for (int m = 0; m < M; ++m)
            w[(size_t)m * K + k] = (uint8_t)(((uint32_t)qweight[(size_t)(k / p) * M + m] >> (bits * (k % p))) & mask);
    for (int gk = 0; gk < NG; ++gk)
        for (int m = 0; m < M; ++m) {
            const float sc = f16_bits_to_f32(scales_f16[(size_t)gk * M + m]);
            int z = (int)(((uint32_t)qzeros[(size_t)gk * (M / p) + m / p] >> (bits * (m % p))) & mask);
            if (!gptq_v2) z += 1;
            const float prod = (float)(z - (1 << (bits - 1))) * sc;       // exact in fp32 (small int x 11-bit mantissa)
            scales[(size_t)m * NG + gk] = sc;
            zeros[(size_t)m * NG + gk] = f16_bits_to_f32(f32_to_f16_bits(prod));
        }
    return true;
}

// ---- converter-side quantisers (3rdparty/llama.cpp/convert_hf_to_gguf.py) ---------------------------------------------------
// BitDistiller-style asymmetric group quantiser, zero_point branch of Model._t_mac_quantize_tensor_bitdistiller (:409-452):
// per group of `group_size` columns (<= 0: the whole row) scale = max(max - min, 1e-5) / (2^bits - 1),
// zero = clamp(-round(min / scale), 0, 2^bits - 1), code = clamp(round(w / scale) + zero, 0, 2^bits - 1) (round half to even, all in
// fp32), returned zeros = (zero - 2^(bits-1)) * scale -- the T-MAC convention W = (code - 2^(bits-1)) * scale - zeros.
inline bool quantize_bitdistiller(const float *w, int rows, int cols, int bits, int group_size, uint8_t *codes, float *scales, float *zeros) {
    const int gs = group_size > 0 ? group_size : cols;
    if (rows <= 0 || cols <= 0 || bits < 1 || bits > 8 || cols % gs) return false;
    const int ng = cols / gs;
    const float max_int = (float)((1 << bits) - 1);
    for (int r = 0; r < rows; ++r)
        for (int g = 0; g < ng; ++g) {
            const float *x = w + (size_t)r * cols + (size_t)g * gs;
            float mx = x[0], mn = x[0];
            for (int i = 1; i < gs; ++i) { mx = x[i] > mx ? x[i] : mx; mn = x[i] < mn ? x[i] : mn; }
            float d = mx - mn;
            if (d < 1e-5f) d = 1e-5f;
            const float sc = d / max_int;
            float zero = -std::nearbyintf(mn / sc);
            zero = zero < 0.f ? 0.f : (zero > max_int ? max_int : zero);
            for (int i = 0; i < gs; ++i) {
                float q = std::nearbyintf(x[i] / sc) + zero;
                q = q < 0.f ? 0.f : (q > max_int ? max_int : q);
                codes[(size_t)r * cols + (size_t)g * gs + i] = (uint8_t)q;
            }
            scales[(size_t)r * ng + g] = sc;
            zeros[(size_t)r * ng + g] = (zero - (float)(1 << (bits - 1))) * sc;
        }
    return true;
}
// BitNet b1.58: BitnetModel.weight_quant (:1884-1893: s = max(mean|w|, 1e-5), t = clamp(round(w / s), -1, 1), stored as t / (1/s))
// followed by the T-MAC ternary rule (:1909-1917): scale = max|t / (1/s)|, code = round(value / scale + 2) in {1, 2, 3}; one scale
// per tensor.  The mean is accumulated in double (torch sums in fp32 with its own blocking), so the scale may differ from the
// reference by one ulp; the codes only could at an exact rounding tie.
inline void quantize_bitnet(const float *w, size_t n, uint8_t *codes, float *scale) {
    double acc = 0.0;
    for (size_t i = 0; i < n; ++i) acc += std::fabs((double)w[i]);
    float s = n ? (float)(acc / (double)n) : 0.f;
    if (s < 1e-5f) s = 1e-5f;
    const float iscale = 1.0f / s;
    float mxabs = 0.f;
    for (size_t i = 0; i < n; ++i) {
        float t = std::nearbyintf(w[i] * iscale);
        t = t < -1.f ? -1.f : (t > 1.f ? 1.f : t);
        const float v = t / iscale;
        codes[i] = (uint8_t)((int)t + 2);
        const float a = std::fabs(v);
        mxabs = a > mxabs ? a : mxabs;
    }
    *scale = mxabs;
}

// Inverse of the reference permutation (python/t_mac/weights.py:57-73): bit-plane row p of the
// tensor, K-group kg -> byte / nibble in A [M/bm][K/4][bm/2].
inline uint8_t ref_layout_idx(const uint8_t *A, int KG, int bm, int kf, int p, int kg) {
    const int tile = p / bm, pin = p % bm, slab = pin / 32, s = pin % 16, half = (pin % 32) / 16;
    const size_t byte = (((size_t)tile * (KG / kf) + kg / kf) * (bm / 32) + slab) * kf * 16 + (size_t)(kg % kf) * 16 + s;
    return (uint8_t)((A[byte] >> (4 * half)) & 15);
}

inline void plain_from_reference(const uint8_t *A, const float *S, int Mout, int K, int bits, int bm, int kf,
                                 int group_size, int zero_point, int one_scale, PlainWeights *P) {
    P->Mout = Mout; P->K = K; P->bits = bits;
    const int KG = K / 4;
    P->idx.assign((size_t)Mout * bits * KG, 0);
    for (int row = 0; row < Mout; ++row)
        for (int b = 0; b < bits; ++b) {
            const int p = (row / 8) * 8 * bits + b * 8 + row % 8;   // weights.py:65
            uint8_t *dst = &P->idx[((size_t)row * bits + b) * KG];
            for (int g = 0; g < KG; ++g) dst[g] = ref_layout_idx(A, KG, bm, kf, p, g);
        }
    if (one_scale) { P->scales.assign(1, S[0]); P->zeros.clear(); return; }
    const int NG = K / group_size, rows = bm / bits;
    P->scales.assign((size_t)Mout * NG, 0.f);
    if (zero_point) P->zeros.assign((size_t)Mout * NG, 0.f); else P->zeros.clear();
    for (int row = 0; row < Mout; ++row) {
        const int tile = row / rows, rin = row % rows;
        for (int wg = 0; wg < NG; ++wg) {                           // weights.py:75-84
            const size_t base = ((size_t)tile * NG + wg) * rows * (zero_point ? 2 : 1);
            if (zero_point) {
                P->scales[(size_t)row * NG + wg] = S[base + (size_t)(rin / 8) * 16 + rin % 8];
                P->zeros[(size_t)row * NG + wg] = S[base + (size_t)(rin / 8) * 16 + 8 + rin % 8];
            } else
                P->scales[(size_t)row * NG + wg] = S[base + rin];
        }
    }
}

// Encode the four words (groups 4q..4q+3 of the chunk) of one lane.
inline void pack_quad(const PlainWeights &P, int pb, int row_base, int g_base, uint32_t w[4]) {
    const int bits = P.bits, rw = 8 / pb;
    uint32_t code[8][4][4];  // [row i][plane b][k]
    for (int i = 0; i < rw; ++i)
        for (int b = 0; b < pb; ++b)
            for (int k = 0; k < 4; ++k) {
                const int row = row_base + i;
                uint32_t idx = (row < P.Mout && b < bits) ? P.at(row, b, g_base + k) : 0u;
                code[i][b][k] = idx_to_code(idx);
            }
    for (int k = 0; k < 4; ++k) w[k] = 0;
    if (pb == 4) {
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 2; ++i)
                for (int b = 0; b < 4; ++b) w[k] |= code[i][b][k] << (4 * (4 * i + b));
    } else if (pb == 2) {
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 4; ++i)
                for (int b = 0; b < 2; ++b) w[k] |= (code[i][b][k] & 7u) << (4 * (2 * i + b));
        for (int pair = 0; pair < 2; ++pair)          // words (2pair, 2pair+1)
            for (int r = 0; r < 4; ++r) {             // row r -> word 2pair + r/2, half r%2
                const int wi = 2 * pair + r / 2, half = r % 2;
                for (int n = 0; n < 4; ++n) {         // (ge p0, ge p1, go p0, go p1)
                    const uint32_t neg = code[r][n & 1][2 * pair + (n >> 1)] >> 3;
                    w[wi] |= neg << (16 * half + 4 * n + 3);
                }
            }
    } else {
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 8; ++i) w[k] |= (code[i][0][k] & 7u) << (4 * i);
        for (int h = 0; h < 8; ++h)                   // row h -> word h/2, half h%2, nibble k
            for (int k = 0; k < 4; ++k) w[h / 2] |= (code[h][0][k] >> 3) << (16 * (h % 2) + 4 * k + 3);
    }
}

inline void encode_stream(const PlainWeights &P, const StreamLayout &L, uint8_t *out) {
    std::memset(out, 0, L.total);
    const int NG = L.one_scale ? 1 : P.K / L.group_size;
    for (int rsb = 0; rsb < L.nrsb; ++rsb)
        for (int c = 0; c < L.nchunk; ++c) {
            uint8_t *blk = out + (size_t)rsb * L.rsb_stride + (size_t)c * L.blk;
            uint32_t *wq = reinterpret_cast<uint32_t *>(blk);
            for (int q = 0; q < L.qch; ++q)
                for (int lane = 0; lane < 32; ++lane)
                    pack_quad(P, L.pb, rsb * L.rsb + lane * L.rw, (c * L.qch + q) * 4, wq + ((size_t)q * 32 + lane) * 4);
            if (L.one_scale) continue;
            const int wg = (int)(((size_t)c * L.ck) / L.group_size);
            (void)NG;
            uint8_t *sp = blk + L.wbytes;
            for (int part = 0; part < (L.zp ? 2 : 1); ++part)
                for (int r = 0; r < L.rsb; ++r) {
                    const int row = rsb * L.rsb + r;
                    float v = 0.f;
                    if (row < P.Mout) v = part ? P.zeros[(size_t)row * (P.K / L.group_size) + wg]
                                               : P.scales[(size_t)row * (P.K / L.group_size) + wg];
                    uint8_t *dst = sp + ((size_t)part * L.rsb + r) * L.sd;
                    if (L.sd == 2) { uint16_t h = f32_to_f16_bits(v); std::memcpy(dst, &h, 2); }
                    else std::memcpy(dst, &v, 4);
                }
        }
}

}  // namespace tmac_b200
