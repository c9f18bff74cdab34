// Microbenchmark: ceiling of "read 12.7 MB per launch" on B200 (no compute), to separate the
// small-launch memory-system ceiling from the GEMV kernel's own cost.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int NLD>
__global__ void read_kernel(const uint4* __restrict__ src, size_t n16, unsigned* out, int pdl) {
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    size_t base = ((size_t)blockIdx.x * blockDim.x / 32 + threadIdx.x / 32) * (NLD * 32) + (threadIdx.x & 31);
    uint4 v[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        size_t idx = base + (size_t)i * 32;
        if (idx < n16) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w) : "l"(src + idx));
        else v[i] = make_uint4(0, 0, 0, 0);
    }
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < NLD; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
    if (acc == 0x12345678u) out[0] = acc;
}

int main() {
    const size_t bytes = 12741632 / 16 * 16;   // one W2 11008x4096 tensor incl. scales
    const int L = 32;
    std::vector<uint4*> bufs(L);
    for (int i = 0; i < L; ++i) { CK(cudaMalloc(&bufs[i], bytes)); CK(cudaMemset(bufs[i], i + 1, bytes)); }
    unsigned* out; CK(cudaMalloc(&out, 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const size_t n16 = bytes / 16;
    for (int pdl = 0; pdl < 2; ++pdl)
        for (int threads : {128, 256, 512}) {
            const int NLD = 8;
            const int warps = (int)((n16 + NLD * 32 - 1) / (NLD * 32));
            const int wpb = threads / 32;
            const int grid = (warps + wpb - 1) / wpb;
            auto launch = [&](int i) {
                cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
                cfg.attrs = at; cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, read_kernel<8>, (const uint4*)bufs[i], n16, out, pdl));
            };
            for (int i = 0; i < L; ++i) launch(i);
            CK(cudaStreamSynchronize(st));
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < L; ++i) launch(i);
            CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
            for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int reps = 20;
            CK(cudaEventRecord(e0, st));
            for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double us = ms * 1e3 / (reps * L);
            printf("pdl=%d threads=%d grid=%d : %.2f us/launch  %.0f GB/s\n", pdl, threads, grid, us, bytes / us / 1e3);
        }
    // one big launch for reference
    {
        const size_t big = (size_t)L * bytes; uint4* b; CK(cudaMalloc(&b, big)); CK(cudaMemset(b, 1, big));
        const size_t n = big / 16; const int warps = (int)((n + 255) / 256); const int grid = (warps + 7) / 8;
        for (int w = 0; w < 2; ++w) read_kernel<8><<<grid, 256, 0, st>>>(b, n, out, 0);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < 5; ++r) read_kernel<8><<<grid, 256, 0, st>>>(b, n, out, 0);
        CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("single launch of %.0f MB: %.1f us  %.0f GB/s\n", big / 1e6, ms * 1e3 / 5, big / (ms / 5 * 1e-3) / 1e9);
    }
    return 0;
}
