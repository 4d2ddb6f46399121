// Third-generation tcgen05 Metropolis-Hastings sampler: two tile contexts per CTA.
//
// Same math and operand images as mh_tc.cu / mh_tc2.cu.  Phase timing of the second generation
// (tools/tc_phase_clocks.py) showed ~10 k of the ~20 k cycles of an evaluation spent in serial bubbles (MMA issue and
// completion latency between the layers, accept / propose, chunk hand-over) during which the MUFU pipe -- the real
// bound of this kernel -- idles.  Two CTAs per SM would fill those bubbles for free, but 184 KB of resident weights
// allow only one.  So here the LAST layer's weights (132 KB) are no longer resident: each 128-bin chunk of W3 (32 KB)
// is streamed from L2 into a per-context slot with cp.async.bulk just before its MMA, which frees enough shared
// memory for TWO independent tile contexts inside one CTA:
//
//   shared memory   W1 | W2 | biases (shared, resident)  +  per context: activations 32 KB, W3 slot 32 KB, proposals 8 KB
//   tensor memory   per context 256 columns: layers 1-2 accumulator [0,128), layer-3 chunk [128,256)
//   threads         16 warps = 2 contexts x 8 warps (TMEM lane quadrant x column half); context-local named barriers
//
// The contexts run the simple serial schedule (write operand -> MMA -> epilogue -> ...) on different tiles and drift
// against each other, so one context's MUFU-heavy epilogue overlaps the other's MMA / load / accept latency.
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int C3_THREADS = 512;
constexpr int C3_CTX_BYTES = 32768 + 32768 + 1024 + 8192 + 8192;   // activations, W3 slot, partial sums (padded), proposals, states

struct Mh3Params {
    Dims d;
    const unsigned char* image;
    int64_t rows;
    int C;
    const float* y;
    const uint4* PVpk;            // [tile][quad][128] {P0P1, P2P3, V0V1, V2V3} in BF16
    const float* g;
    float* Z;
    float* Zs;
    const float* eps;
    const float* u;
    uint32_t* n_accept;
    float* a_trace;
    int n_burn, n_keep;
    float sd;
    int* status;
    long long* dbg;
};

static long long* g_dbg_clocks3 = nullptr;

#define DBG3(slot, cond)                                                                            \
    do {                                                                                            \
        if (p.dbg && blockIdx.x == 0 && ctx == 0 && tile == 0 && it == 4 && (cond)) p.dbg[slot] = clock64(); \
    } while (0)

__device__ __forceinline__ void ctx_bar(int ctx) { asm volatile("bar.sync %0, 256;" :: "r"(1 + ctx) : "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

template <int L>
__global__ void __launch_bounds__(C3_THREADS, 1) mh3_kernel(Mh3Params p) {
    static_assert(L == 16, "the two-context sampler keeps 128 x L proposals in 8 KB per context");
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bars[6];                       // per context: layers 1/2 ready, chunk ready, W3 slot landed
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag;

    const Dims& d = p.d;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctx = warp >> 3, w = warp & 7;
    const int q = w & 3, h = w >> 2;
    const bool owner = h == 0, lead = lane == 0;
    const int row = 32 * q + lane;
    // shared operands: W1 | W2 | biases at their image offsets (W3 of the image is NOT copied)
    const int shared_bytes = (d.off_w3 + 4 * ((d.n_hidden == 2 ? HID : 0) + NPAD) + 1023) & ~1023;
    float* biasp = reinterpret_cast<float*>(base + d.off_w3);
    unsigned char* cbase = base + shared_bytes + ctx * C3_CTX_BYTES;
    unsigned char* A = cbase;
    unsigned char* Wslot = cbase + 32768;
    float* red = reinterpret_cast<float*>(cbase + 65536);
    float* zpS = reinterpret_cast<float*>(cbase + 65536 + 1024);
    float* zS = zpS + TM * L;                          // [128][L] chain states (touched by the row's owner only)
    const uint32_t bar12 = smem_u32(&bars[3 * ctx]), bar3 = smem_u32(&bars[3 * ctx + 1]), barw = smem_u32(&bars[3 * ctx + 2]);
    uint32_t ph12 = 0, ph3 = 0, phw = 0;

    {   // W1, W2 and the biases (the biases sit behind W3 in the image)
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(base);
        for (int i = threadIdx.x; i < d.off_w3 / 16; i += C3_THREADS) dst[i] = __ldg(src + i);
        const uint4* bsrc = reinterpret_cast<const uint4*>(p.image + d.off_bias);
        uint4* bdst = reinterpret_cast<uint4*>(biasp);
        for (int i = threadIdx.x; i < (d.image_bytes - d.off_bias) / 16; i += C3_THREADS) bdst[i] = __ldg(bsrc + i);
    }
    if (threadIdx.x == 0) {
        dead_flag = 0;
        for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot + 256 * ctx;       // this context's 256 columns
    volatile int* dead = &dead_flag;

    const uint32_t a_addr = smem_u32(A), slot_addr = smem_u32(Wslot);
    const uint32_t w1_addr = smem_u32(base), w2_addr = smem_u32(base + d.off_w2);
    const float* b2 = (d.n_hidden == 2) ? biasp : nullptr;
    const unsigned char* w3g = p.image + d.off_w3;     // global W3 image: K block kb at kb*NPAD*128, row r at r*128
    const int n_iter = p.n_burn + p.n_keep;
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    const int y_dim = d.y_dim, nkb1 = d.nkb1;
    const bool two_hidden = d.n_hidden == 2;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;

    // tiles are dealt round by round over the CTAs, alternating between the two contexts: with an odd number of rounds the
    // extra tile goes to context 0 of EVERY CTA (it then runs alone) instead of a full extra pair on half of the CTAs
    for (int64_t tile = (int64_t)blockIdx.x + (int64_t)ctx * gridDim.x; tile < n_tiles; tile += 2 * (int64_t)gridDim.x) {
        const int64_t row_g = tile * TM + row;
        const bool valid = row_g < p.rows;
        const int64_t fr = valid ? row_g / p.C : 0;
        const float g_row = valid ? p.g[fr] : 1.f;
        const uint4* PVt = p.PVpk + (tile * NQ) * TM + row;

        float y0 = 0.f, y1 = 0.f, y2 = 0.f, ll_cur = 0.f, u_cur = 0.5f, prior = 0.f;
        uint32_t n_acc = 0;
        if (owner) {
#pragma unroll
            for (int l = 0; l < L / 4; ++l)
                *reinterpret_cast<float4*>(zS + row * L + 4 * l) =
                    valid ? __ldg(reinterpret_cast<const float4*>(p.Z + row_g * L) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                if (y_dim > 0) y0 = p.y[fr * y_dim];
                if (y_dim > 1) y1 = p.y[fr * y_dim + 1];
                if (y_dim > 2) y2 = p.y[fr * y_dim + 2];
            }
        }

        for (int it = -1; it < n_iter; ++it) {
            DBG3(0, threadIdx.x == 0);
            if (owner) {
                float zc[L];
#pragma unroll
                for (int l = 0; l < L / 4; ++l) {
                    const float4 t4 = *reinterpret_cast<const float4*>(zS + row * L + 4 * l);
                    zc[4 * l] = t4.x; zc[4 * l + 1] = t4.y; zc[4 * l + 2] = t4.z; zc[4 * l + 3] = t4.w;
                }
                if (it >= 0) {
                    // draws of this proposal straight from global memory: the other context hides the latency
                    float zp[L];
                    prior = 0.f;
                    if (valid) {
                        const float4* e = reinterpret_cast<const float4*>(p.eps + ((int64_t)it * p.rows + row_g) * L);
                        u_cur = __ldg(p.u + (int64_t)it * p.rows + row_g);
#pragma unroll
                        for (int l = 0; l < L / 4; ++l) {
                            const float4 e4 = __ldg(e + l);
                            zp[4 * l + 0] = __fadd_rn(zc[4 * l + 0], __fmul_rn(p.sd, e4.x));
                            zp[4 * l + 1] = __fadd_rn(zc[4 * l + 1], __fmul_rn(p.sd, e4.y));
                            zp[4 * l + 2] = __fadd_rn(zc[4 * l + 2], __fmul_rn(p.sd, e4.z));
                            zp[4 * l + 3] = __fadd_rn(zc[4 * l + 3], __fmul_rn(p.sd, e4.w));
                        }
                    } else {
#pragma unroll
                        for (int l = 0; l < L; ++l) zp[l] = 0.f;
                    }
#pragma unroll
                    for (int l = 0; l < L; ++l) prior += __fsub_rn(__fmul_rn(zc[l], zc[l]), __fmul_rn(zp[l], zp[l]));
#pragma unroll
                    for (int l = 0; l < L / 4; ++l)
                        *reinterpret_cast<float4*>(zpS + row * L + 4 * l) = make_float4(zp[4 * l], zp[4 * l + 1], zp[4 * l + 2], zp[4 * l + 3]);
                    write_a1_static<L>(y_dim, nkb1, A, row, zp, y0, y1, y2, valid);
                } else {
                    write_a1_static<L>(y_dim, nkb1, A, row, zc, y0, y1, y2, valid);
                }
            }
            // first W3 chunk of this evaluation: the slot is free (every MMA of the previous evaluation has completed)
            if (w == 3 && lead) {
                mbar_expect_tx(barw, 32768);
                bulk_g2s(slot_addr, w3g, 16384, barw);
                bulk_g2s(slot_addr + 16384, w3g + NPAD * 128, 16384, barw);
            }
            fence_async_smem();
            ctx_bar(ctx);                                                       // S1: layer-1 operand ready
            DBG3(1, threadIdx.x == 0);

            if (w == 1 && lead) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w1_addr, 16384, nkb1, tmem, HID);
                umma_commit(bar12);
            }
            mbar_wait(bar12, ph12, dead, p.status);
            ph12 ^= 1;
            tc_fence_after();
            hidden_epilogue_rows_bf(tmem, A, q, h, row, nullptr);
            fence_async_smem();
            tc_fence_before();
            ctx_bar(ctx);                                                       // S2
            DBG3(2, threadIdx.x == 0);
            if (two_hidden) {
                if (w == 2 && lead) {
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w2_addr, 16384, 2, tmem, HID);
                    umma_commit(bar12);
                }
                mbar_wait(bar12, ph12, dead, p.status);
                ph12 ^= 1;
                tc_fence_after();
                hidden_epilogue_rows_bf(tmem, A, q, h, row, b2);
                fence_async_smem();
                tc_fence_before();
                ctx_bar(ctx);                                                   // S3
                DBG3(3, threadIdx.x == 0);
            }

            // ---- layer 3: chunks of 128 bins (+ bin 512), W3 streamed chunk by chunk through the context's slot
            float acc = 0.f, accl = 0.f;
            uint4 pvA[4], pvB[4];
#define MH3_QUAD(t) (32 * ((t) >> 2) + 16 * h + 4 * ((t) & 3))
#define MH3_LOAD(t, PV)                                                                              \
    do {                                                                                             \
        _Pragma("unroll") for (int qd = 0; qd < 4; ++qd) PV[qd] = __ldg(PVt + (MH3_QUAD(t) + qd) * TM); \
    } while (0)
            MH3_LOAD(0, pvA);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                if (w == 3 && lead) {                                           // chunk j: weights landed -> MMA
                    mbar_wait(barw, phw, dead, p.status);
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, slot_addr, 16384, 2, tmem + 128, j < 4 ? 128 : 16);
                    umma_commit(bar3);
                }
                phw ^= 1;
                mbar_wait(bar3, ph3, dead, p.status);
                ph3 ^= 1;
                tc_fence_after();
                DBG3(15 + j, threadIdx.x == 0);
                if (w == 3 && lead && j < 4) {                                  // slot consumed: request the next chunk
                    const uint32_t bytes = (j + 1 < 4) ? 16384u : 2048u;
                    mbar_expect_tx(barw, 2 * bytes);
                    bulk_g2s(slot_addr, w3g + (size_t)(j + 1) * 128 * 128, bytes, barw);
                    bulk_g2s(slot_addr + 16384, w3g + NPAD * 128 + (size_t)(j + 1) * 128 * 128, bytes, barw);
                }
                if (j < 4) {
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        const int t = 4 * j + sub;
                        if (t + 1 < 16) {                                       // quads of the next sub-chunk
                            if (t & 1) MH3_LOAD(t + 1, pvA); else MH3_LOAD(t + 1, pvB);
                        }
                        float v[16];
                        tmem_ld16(tmem + 128 + lane_off + 64 * h + 16 * sub, v);
                        tmem_wait_ld();
                        if (t & 1) loglik16_pv<false>(v, pvB, g_row, acc, accl);
                        else loglik16_pv<false>(v, pvA, g_row, acc, accl);
                    }
                } else if (h == 0) {                                            // bin 512
                    float v[4];
                    tmem_ld4(tmem + 128 + lane_off, v);
                    tmem_wait_ld();
                    const uint4 pv = __ldg(PVt + 128 * TM);
                    // word 0 of the quad: P'_512 with bf16(k Vb'_512) in its low half (pack_pv_kernel); acc carries sums / k
                    const float xk = fmaf(g_row * kPairScale, ex2_approx(v[0]), __uint_as_float(pv.x << 16));
                    acc = fmaf(__uint_as_float(pv.x), rcp_approx(xk), acc);
                    accl += lg2_approx(xk);
                }
                tc_fence_before();
                if (j == 4 && h == 1) red[row] = fmaf(kLn2, accl, acc * kQuadScale);
                DBG3(4 + j, threadIdx.x == 0);
                ctx_bar(ctx);                                                   // chunk buffer drained (j = 4: S4)
            }
#undef MH3_LOAD
#undef MH3_QUAD
            DBG3(9, threadIdx.x == 0);

            if (owner) {
                const float ll_prop = fmaf(kLn2, accl, acc * kQuadScale) + red[row];
                if (it < 0) {
                    ll_cur = ll_prop;
                } else if (valid) {
                    const float a = (ll_cur - ll_prop) + 0.5f * prior;
                    if (p.a_trace) p.a_trace[(int64_t)it * p.rows + row_g] = a;
                    const bool accept = __logf(u_cur) < a;
                    if (accept) {
#pragma unroll
                        for (int l = 0; l < L / 4; ++l)
                            *reinterpret_cast<float4*>(zS + row * L + 4 * l) = *reinterpret_cast<const float4*>(zpS + row * L + 4 * l);
                        ll_cur = ll_prop;
                        ++n_acc;
                    }
                    if (it >= p.n_burn) {
                        float4* dst = reinterpret_cast<float4*>(p.Zs + (row_g * p.n_keep + (it - p.n_burn)) * L);
#pragma unroll
                        for (int l = 0; l < L / 4; ++l) dst[l] = *reinterpret_cast<const float4*>(zS + row * L + 4 * l);
                    }
                }
            }
            DBG3(23, threadIdx.x == 0);
        }
        if (owner && valid) {
#pragma unroll
            for (int l = 0; l < L / 4; ++l)
                reinterpret_cast<float4*>(p.Z + row_g * L)[l] = *reinterpret_cast<const float4*>(zS + row * L + 4 * l);
            if (p.n_accept) p.n_accept[row_g] += n_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_slot), "r"(512) : "memory");
    }
}

}  // namespace tc
}  // namespace dvae

using namespace dvae;
using namespace dvae::tc;

extern "C" int dvae_mh_chain_tc3(const DvaeMlp* dec, const void* image, const void* PVpk, const float* g,
                                 const float* y, int y_dim, float* Z, float* Zs, int64_t NT, int L, int n_chains, int n_burn,
                                 int n_keep, float var_rw, const float* eps, const float* u, uint32_t* n_accept,
                                 float* a_trace, int* status, void* stream) {
    Mh3Params p{};
    int rc = check_dims(dec, L, y_dim, "dvae_mh_chain_tc3", &p.d);
    if (rc) return rc;
    DVAE_REQUIRE(L == 16, "dvae_mh_chain_tc3: latent size must be 16 (got %d)", L);
    DVAE_REQUIRE(y_dim <= 3, "dvae_mh_chain_tc3: at most 3 label inputs");
    DVAE_REQUIRE(image && PVpk && g && Z && Zs && eps && u && status, "dvae_mh_chain_tc3: null pointer");
    DVAE_REQUIRE(y_dim == 0 || y, "dvae_mh_chain_tc3: y_dim=%d but y is null", y_dim);
    DVAE_REQUIRE(NT >= 0 && n_chains >= 1 && n_chains < 4096 && n_burn >= 0 && n_keep >= 1 && var_rw > 0.f, "dvae_mh_chain_tc3: bad sizes");
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(Zs) & 15) == 0 && (reinterpret_cast<uintptr_t>(eps) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(image) & 15) == 0 && (reinterpret_cast<uintptr_t>(Z) & 15) == 0,
                 "dvae_mh_chain_tc3: Z, Zs, eps and image must be 16-byte aligned");
    if (NT == 0) return 0;
    p.image = (const unsigned char*)image;
    p.rows = NT * n_chains; p.C = n_chains; p.y = y;
    p.PVpk = (const uint4*)PVpk; p.g = g; p.Z = Z; p.Zs = Zs;
    p.eps = eps; p.u = u;
    p.n_accept = n_accept; p.a_trace = a_trace; p.n_burn = n_burn; p.n_keep = n_keep;
    p.sd = sqrtf(var_rw);
    p.status = status;
    p.dbg = g_dbg_clocks3;
    const int shared_bytes = (p.d.off_w3 + 4 * ((p.d.n_hidden == 2 ? HID : 0) + NPAD) + 1023) & ~1023;
    const size_t smem = (size_t)shared_bytes + 2 * C3_CTX_BYTES + 1024;
    DVAE_REQUIRE(smem <= 227 * 1024, "dvae_mh_chain_tc3: shared memory budget exceeded");
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    const int grid = (int)(n_tiles < 148 ? n_tiles : 148);
    cudaFuncSetAttribute(mh3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mh3_kernel<16><<<grid, C3_THREADS, smem, (cudaStream_t)stream>>>(p);
    return check_launch("mh3_kernel");
}

extern "C" int dvae_debug_set_clock_buffer3(void* dev_buffer) {
    g_dbg_clocks3 = reinterpret_cast<long long*>(dev_buffer);
    return 0;
}
