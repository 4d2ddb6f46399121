// Probe: sustained tcgen05.ld / tcgen05.st throughput of one SM (bytes per clock) as a function of the number of warps
// issuing, and the latency of a single load.  The sampler and the decode read 270-280 KB of accumulators per 128-row tile
// and the sampler's bias preload writes 64 KB per evaluation; DESIGN.md section 8 needs the figure these are bounded by.
// Build + run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dvae_b200/csrc -o gpurun_out/tmem_bw_probe tools/tmem_bw_probe.cu && gpurun_out/tmem_bw_probe
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace dvae;
using namespace dvae::tc;

__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* r) {
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

// mode 0: x16 loads, one wait per load (latency-exposed); 1: x16 loads, wait every 4; 2: x32 loads, wait every 4;
// 3: x32 stores, wait every 4.  out[0] = cycles of warp 0, out[1] = checksum (keeps the loads alive).
__global__ void __launch_bounds__(1024, 1) probe(int mode, int iters, long long* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;       // a warp may only touch its own lane quadrant
    const uint32_t col0 = 64 * ((warp >> 2) & 7);                       // warps of one quadrant use different columns
    float v[32];
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = 0.f; r[i] = threadIdx.x + i; }
    float sum = 0.f;
    st32(tmem + lane_off + col0, r);
    st32(tmem + lane_off + col0 + 32, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = tmem + lane_off + col0 + 16 * (it & 1);
        if (mode == 0) { tmem_ld16(a, v); tmem_wait_ld(); sum += v[0]; }
        else if (mode == 1) { tmem_ld16(a, v); if ((it & 3) == 3) { tmem_wait_ld(); sum += v[0]; } }
        else if (mode == 2) { tmem_ld32(tmem + lane_off + col0 + 32 * (it & 1), v); if ((it & 3) == 3) { tmem_wait_ld(); sum += v[0]; } }
        else { st32(tmem + lane_off + col0 + 32 * (it & 1), r); if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    }
    tmem_wait_ld();
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = t1 - t0; }
    if (sum == 123.456f) out[1] = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const int iters = 4096;
    const char* names[4] = {"ld x16, wait each", "ld x16, wait / 4", "ld x32, wait / 4", "st x32, wait / 4"};
    const int bytes[4] = {16 * 32 * 4, 16 * 32 * 4, 32 * 32 * 4, 32 * 32 * 4};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 1; warps <= 32; warps *= 2) {
            probe<<<1, 32 * warps>>>(mode, iters, out);
            long long h[2] = {0, 0};
            cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            printf("%-18s warps %2d: %7.1f clk per instruction per warp, %7.1f B/clk per SM\n", names[mode], warps,
                   (double)h[0] / iters, (double)bytes[mode] * iters * warps / (double)h[0]);
        }
    return 0;
}
