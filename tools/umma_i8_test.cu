// Step-1 validation of tcgen05.mma kind::i8 on sm_100a: D[128][N] (int32, TMEM) = A[128][K] * B[N][K]^T, int8 operands in
// shared memory in the K-major no-swizzle canonical layout (8 rows x 16 B core matrices).  Compares with a host reference.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    return d;                          // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void umma_test(const int8_t* A, const int8_t* B, int32_t* D, int N, int K, int ncols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 128 * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kch = K / 16;
    for (int i = tid; i < 128 * kch; i += blockDim.x) {
        const int r = i / kch, kc = i % kch;
        *(uint4*)(sA + ((size_t)(kc * 16 + (r >> 3)) * 128 + (r & 7) * 16)) = *(const uint4*)(A + (size_t)r * K + kc * 16);
    }
    for (int i = tid; i < N * kch; i += blockDim.x) {
        const int r = i / kch, kc = i % kch;
        *(uint4*)(sB + ((size_t)(kc * (N / 8) + (r >> 3)) * 128 + (r & 7) * 16)) = *(const uint4*)(B + (size_t)r * K + kc * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lboA = 16 * 128, lboB = (N / 8) * 128;
        for (int k32 = 0; k32 < K / 32; ++k32) {
            const uint64_t da = make_desc(s32(sA) + k32 * 2 * lboA, lboA, 128);
            const uint64_t db = make_desc(s32(sB) + k32 * 2 * lboB, lboB, 128);
            const uint32_t acc = k32 > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(s32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N / 32; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c * 32 + j] = (int32_t)r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(ncols) : "memory");
}

int main() {
    for (int N : {64, 128, 256})
        for (int K : {32, 128, 256}) {
            std::vector<int8_t> hA(128 * K), hB((size_t)N * K);
            srand(N * 1000 + K);
            for (auto& v : hA) v = (int8_t)(rand() % 7 - 3);
            for (auto& v : hB) v = (int8_t)(rand() % 255 - 127);
            int8_t *dA, *dB; int32_t* dD;
            CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
            CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
            CK(cudaMemset(dD, 0xFF, (size_t)128 * N * 4));
            const size_t smem = (size_t)(128 + N) * K;
            CK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int ncols = 32; while (ncols < N) ncols *= 2;
            umma_test<<<1, 128, smem>>>(dA, dB, dD, N, K, ncols);
            CK(cudaDeviceSynchronize());
            std::vector<int32_t> hD((size_t)128 * N);
            CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
            long bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    int32_t s = 0;
                    for (int k = 0; k < K; ++k) s += (int)hA[m * K + k] * (int)hB[(size_t)n * K + k];
                    if (s != hD[(size_t)m * N + n]) { if (bad < 5) printf("  mismatch m=%d n=%d got %d want %d\n", m, n, hD[(size_t)m * N + n], s); ++bad; }
                }
            printf("N=%d K=%d : %s (%ld mismatches)\n", N, K, bad ? "FAIL" : "ok", bad);
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    return 0;
}
