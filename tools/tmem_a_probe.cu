// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (D[tmem] = A[tmem] * B[smem]^T), M = 128, N = 128, K = 128, BF16.
// Establishes the packing of a K-major BF16 A tile in TMEM (which half of a 32-bit column holds the even k) before the
// sampler relies on it.  Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dvae_b200/csrc -o gpurun_out/tmem_a_probe tools/tmem_a_probe.cu && gpurun_out/tmem_a_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace dvae;
using namespace dvae::tc;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// A: [128][128] bf16 row-major (global), Bsw: [128 n][128 k] bf16 already in the SW128 K-major image (2 K blocks of 16 KB)
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* A, const unsigned char* Bsw, float* D, int swap_halves, int* status) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(base)[i] = reinterpret_cast<const uint4*>(Bsw)[i];
    const uint32_t bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) { dead_flag = 0; mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // A tile -> TMEM columns [128, 192): lane = row, column c holds (k = 2c, 2c+1)
    {
        uint32_t w[32];
        for (int part = 0; part < 2; ++part) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float a0 = __bfloat162float(A[row * 128 + 64 * part + 2 * c]);
                const float a1 = __bfloat162float(A[row * 128 + 64 * part + 2 * c + 1]);
                w[c] = swap_halves ? pack_bf16x2(a1, a0) : pack_bf16x2(a0, a1);
            }
            tmem_st32(tmem + 128 + 32 * part + ((uint32_t)(32 * warp) << 16), w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc(128);
        uint32_t acc = 0;
        for (int kb = 0; kb < 2; ++kb)
            for (int k = 0; k < 4; ++k) {
                // K = 16 per instruction = 8 TMEM columns of packed pairs
                umma_ts(tmem, tmem + 128 + 32 * kb + 8 * k, umma_desc(smem_u32(base) + kb * 16384 + 32 * k), idesc, acc);
                acc = 1;
            }
        umma_commit(bar_a);
    }
    volatile int* dead = &dead_flag;
    mbar_wait(bar_a, 0, dead, status);
    tc_fence_after();
    for (int part = 0; part < 4; ++part) {
        float v[32];
        tmem_ld32(tmem + 32 * part + ((uint32_t)(32 * warp) << 16), v);
        tmem_wait_ld();
        for (int c = 0; c < 32; ++c) D[row * 128 + 32 * part + c] = v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
    }
}

int main() {
    const int M = 128, N = 128, K = 128;
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<unsigned char> hBsw(32768, 0);
    srand(1);
    for (auto& x : hA) x = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    for (auto& x : hB) x = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const int kb = k >> 6, kk = k & 63, c = kk >> 3, e = kk & 7;          // sw128_offset(rows = 128, n, k)
            const size_t off = (size_t)kb * 128 * 128 + n * 128 + ((c ^ (n & 7)) << 4) + e * 2;
            *reinterpret_cast<__nv_bfloat16*>(&hBsw[off]) = hB[n * K + k];
        }
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)__bfloat162float(hA[m * K + k]) * (double)__bfloat162float(hB[n * K + k]);
            ref[m * N + n] = s;
        }
    __nv_bfloat16* dA; unsigned char* dB; float* dD; int* dS;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, 32768); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hBsw.data(), 32768, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int swap = 0; swap < 2; ++swap) {
        cudaMemset(dD, 0, M * N * 4); cudaMemset(dS, 0, 4);
        probe_kernel<<<1, 128, 40000>>>(dA, dB, dD, swap, dS);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> hD(M * N); int st = 0;
        cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(hD[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
        printf("swap_halves=%d: cuda=%s status=%d max|D-ref|=%.4g (max|ref|=%.4g) D[0..3]=%.4f %.4f %.4f %.4f ref=%.4f %.4f %.4f %.4f\n", swap,
               cudaGetErrorString(e), st, maxerr, maxref, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
