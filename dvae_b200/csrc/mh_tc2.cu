// Second-generation tcgen05 Metropolis-Hastings sampler (the default "tc" sampler).
//
// Same math, operand images and TMEM plan as mh_tc.cu (see the notes there); what changes is the schedule inside
// the CTA, driven by the first ncu capture of mh_tc.cu (early round 1; not kept in profiles/): v1 ran at 0.7 IPC per SM because every phase was
// serialised behind CTA barriers with only two warps per scheduler and a long single-warp "owner" phase.
//
// A second capture (an intermediate 16-warp build) showed 38 % of all stalls on local-memory traffic: with 223 KB of shared
// memory carved out the L1 is tiny, so every spilled register or dynamically indexed array costs an L2 round trip.
// Hence, in this version:
//   * 8 warps only (2 per scheduler -> up to 255 registers per thread): nothing spills, nothing is indexed
//     dynamically (latent size is a template parameter, every loop over chunks is fully unrolled);
//   * the L2 latency of the P / Vb stream is hidden in software: the quads of sub-chunk t+2 (16 bins) are requested
//     while sub-chunk t is evaluated (three rotating register buffers);
//   * the layer-3 chunks are handed over with mbarriers only (chunk ready: tcgen05.commit; buffer free: 256 thread
//     arrivals), no CTA barrier inside the chunk loop;
//   * the random draws are read from global memory (either injected by the caller or produced beforehand by the
//     Philox dump kernel with the same counters as every other sampler), prefetched one iteration ahead;
//   * the tcgen05.mma groups are issued by lane 0 of warps 1 / 2 / 3 / 5 (layer 1, layer 2, layer-3 chunks 0 + 1, chunk 2);
//   * the chain state of row 32*q + lane is split over the two threads (column halves) that share the row: each carries
//     half of the latent dimensions and takes the same accept decision from the same shared-memory partial sums;
//   * every TMEM read is double-buffered, the layer-2 bias is preloaded into the layer-2 accumulator (tcgen05.st) while
//     the layer-1 GEMM runs, and the first P / Vb sub-chunks are requested two phases before the layer-3 loop
//     (phase clocks before / after in DESIGN.md section 4).
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int MH2_THREADS = 256;             // 8 warps: warp w -> TMEM lane quadrant w&3, column half w>>2

struct Mh2Params {
    Dims d;
    const unsigned char* image;
    int64_t rows;                 // NT*C chains
    int C;
    const float* y;
    const uint4* PVpk;            // [tile][quad][128] {P0P1, P2P3, V0V1, V2V3} in BF16
    const float* g;
    float* Z;
    float* Zs;
    const float* eps;             // [n_iter][rows][L] standard normals
    const float* u;               // [n_iter][rows] uniforms
    uint32_t* n_accept;
    float* a_trace;
    int n_burn, n_keep;
    float sd;
    int* status;
    long long* dbg;               // optional [64] clock stamps of CTA 0 (see dvae_debug_set_clock_buffer)
};

static long long* g_dbg_clocks = nullptr;

#define DBG_STAMP(slot, cond)                                                        \
    do {                                                                             \
        if (p.dbg && blockIdx.x == 0 && tile == 0 && it == 4 && (cond)) p.dbg[slot] = clock64(); \
    } while (0)


template <int L, int POLY>
__global__ void __launch_bounds__(MH2_THREADS, 1) mh2_kernel(Mh2Params p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bars[6];                       // layer 1/2 ready, chunk ready x3, buffer free x2
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag;

    const Dims& d = p.d;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* A = base + ((d.image_bytes + 1023) & ~1023);
    float2* red2 = reinterpret_cast<float2*>(A + A_BYTES);       // [2][128] {partial l(z'), partial prior term} of each column half
    const uint32_t bar12 = smem_u32(&bars[0]);
    const uint32_t bar3_0 = smem_u32(&bars[1]), bar3_1 = smem_u32(&bars[2]);
    const uint32_t barf_0 = smem_u32(&bars[3]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = warp >> 2;             // TMEM lane quadrant, column half
    const bool owner = h == 0;                         // warps 0-3 also write the label / constant chunk, the trace and the accept count
    // tcgen05.mma issue is spread over lane 0 of different warps so that no single warp pays for all of it:
    // layer 1: warp 1, layer 2: warp 2, layer-3 chunks 0 + 1: warp 3, chunk 2: warp 5
    const bool lead = lane == 0;
    const int row = 32 * q + lane;
    uint32_t ph12 = 0, ph3_0 = 0, ph3_1 = 0, phf_0 = 0;

    {
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(base);
        for (int i = threadIdx.x; i < d.image_bytes / 16; i += MH2_THREADS) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) {
        dead_flag = 0;
        mbar_init(bar12, 1);
        mbar_init(bar3_0, 1);
        mbar_init(bar3_1, 1);
        mbar_init(barf_0, MH2_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    volatile int* dead = &dead_flag;

    const uint32_t a_addr = smem_u32(A);
    const uint32_t w1_addr = smem_u32(base), w2_addr = smem_u32(base + d.off_w2), w3_addr = smem_u32(base + d.off_w3);
    const float* biasp = reinterpret_cast<const float*>(base + d.off_bias);
    const float* b2 = (d.n_hidden == 2) ? biasp : nullptr;
    const int n_iter = p.n_burn + p.n_keep;
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    const int y_dim = d.y_dim, nkb1 = d.nkb1;
    const bool two_hidden = d.n_hidden == 2;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;

    const long long dbg_t0 = clock64();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row_g = tile * TM + row;
        const bool valid = row_g < p.rows;
        const int64_t fr = valid ? row_g / p.C : 0;
        const float g_row = valid ? p.g[fr] : 1.f;
        const uint4* PVt = p.PVpk + (tile * NQ) * TM + row;

        // Chain state, split over the two column halves: thread (row, h) carries latent dimensions [h L/2, (h+1) L/2) of its
        // row's chain plus a copy of the scalars.  The proposal / accept code is a chain of dependent operations that runs
        // with one warp per scheduler, so halving its length matters more than the duplicated scalar work (the single-owner
        // version spent 1.0 k + 0.8 k of 15 k cycles per evaluation there: tools/tc_phase_clocks.py).  Both halves take the
        // accept decision from the same shared-memory operands in the same order, so they always agree bit for bit.
        constexpr int LH = L / 2;
        float zh[LH], zph[LH];
        float4 enn[LH / 4];                                // this half's draws of the next proposal
        float y0 = 0.f, y1 = 0.f, y2 = 0.f, ll_cur = 0.f, u_cur = 0.5f, u_nxt = 0.5f, prior_h = 0.f;
        uint32_t n_acc = 0;
#pragma unroll
        for (int l = 0; l < LH; ++l) { zh[l] = valid ? p.Z[row_g * L + h * LH + l] : 0.f; zph[l] = 0.f; }
#pragma unroll
        for (int l = 0; l < LH / 4; ++l) enn[l] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (owner && valid) {
            if (y_dim > 0) y0 = p.y[fr * y_dim];
            if (y_dim > 1) y1 = p.y[fr * y_dim + 1];
            if (y_dim > 2) y2 = p.y[fr * y_dim + 2];
        }

        // eval -1 scores the start state, eval it >= 0 scores proposal `it`
        for (int it = -1; it < n_iter; ++it) {
            DBG_STAMP(0, threadIdx.x == 0);
            if (it >= 0) {
                u_cur = u_nxt;
                prior_h = 0.f;
#pragma unroll
                for (int l = 0; l < LH / 4; ++l) {
                    zph[4 * l + 0] = __fadd_rn(zh[4 * l + 0], __fmul_rn(p.sd, enn[l].x));
                    zph[4 * l + 1] = __fadd_rn(zh[4 * l + 1], __fmul_rn(p.sd, enn[l].y));
                    zph[4 * l + 2] = __fadd_rn(zh[4 * l + 2], __fmul_rn(p.sd, enn[l].z));
                    zph[4 * l + 3] = __fadd_rn(zh[4 * l + 3], __fmul_rn(p.sd, enn[l].w));
                }
#pragma unroll
                for (int l = 0; l < LH; ++l) prior_h += __fsub_rn(__fmul_rn(zh[l], zh[l]), __fmul_rn(zph[l], zph[l]));
                write_a1_half<L>(y_dim, nkb1, A, row, h, zph, y0, y1, y2, valid);
            } else {
                write_a1_half<L>(y_dim, nkb1, A, row, h, zh, y0, y1, y2, valid);
            }
            DBG_STAMP(26, threadIdx.x == 0);
            fence_async_smem();
            __syncthreads();                                                    // S1: layer-1 operand ready
            DBG_STAMP(1, threadIdx.x == 0);

            if (warp == 1 && lead) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w1_addr, 16384, nkb1, tmem + 384, HID);
                umma_commit(bar12);
            }
            if (two_hidden) {
                // Layer-2 bias preloaded into the layer-2 accumulator, TMEM columns [0, 128), while the layer-1 GEMM runs (every
                // warp would only spin on its mbarrier here; the stores take ~0.6 k cycles).  Every thread writes the 64 columns
                // it will read back in the layer-2 epilogue.  The columns are free (the previous evaluation's reads of them
                // completed before its S4; the first layer-3 chunk is issued after S3), the layer-2 GEMM, issued after S2,
                // accumulates onto the bias, and its epilogue needs no bias loads / adds (0.5 k of its 1.7 k cycles,
                // tools/tc_phase_clocks.py).
                tc_fence_after();
                tmem_preload_bias64(tmem + lane_off + 64 * h, b2 + 64 * h);
                tmem_wait_st();
                tc_fence_before();
                DBG_STAMP(27, threadIdx.x == 0);
            }
            mbar_wait(bar12, ph12, dead, p.status);
            ph12 ^= 1;
            tc_fence_after();
            DBG_STAMP(24, threadIdx.x == 0);
            hidden_epilogue_rows_bf(tmem + 384, A, q, h, row, nullptr);              // bias rides on the constant-one column
            fence_async_smem();
            tc_fence_before();
            DBG_STAMP(20, threadIdx.x == 0);
            __syncthreads();                                                    // S2
            DBG_STAMP(2, threadIdx.x == 0);
            float acc = 0.f, accl = 0.f;
            uint4 pv0[4], pv1[4], pv2[4];
            // sub-chunk t: chunk c = t/6 (capped at 2), column inside the chunk = h*W_c + 16*(t - 6c), W = 96, 96, 80
#define MH2_CH(t) ((t) < 6 ? 0 : ((t) < 12 ? 1 : 2))
#define MH2_COL(t) (h * (MH2_CH(t) == 2 ? 80 : 96) + 16 * ((t) - 6 * MH2_CH(t)))
#define MH2_BIN(t) (192 * MH2_CH(t) + MH2_COL(t))
#define MH2_LOAD(t, PV)                                                                              \
    do {                                                                                             \
        _Pragma("unroll") for (int qd = 0; qd < 4; ++qd) PV[qd] = __ldg(PVt + ((MH2_BIN(t) >> 2) + qd) * TM); \
    } while (0)
            // the first two sub-chunks of the P / Vb stream are requested here, two phases ahead of their use (the wait for
            // them at the start of the layer-3 loop was 3.6 % of the kernel's stall samples: profiles/r01_tc_ncu_mh2.txt)
            MH2_LOAD(0, pv0);
            MH2_LOAD(1, pv1);
            if (two_hidden) {
                if (warp == 2 && lead) {
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w2_addr, 16384, 2, tmem, HID, 1);
                    umma_commit(bar12);
                }
                mbar_wait(bar12, ph12, dead, p.status);
                ph12 ^= 1;
                tc_fence_after();
                DBG_STAMP(25, threadIdx.x == 0);
                hidden_epilogue_rows_bf(tmem, A, q, h, row, nullptr);
                fence_async_smem();
                tc_fence_before();
                DBG_STAMP(21, threadIdx.x == 0);
                __syncthreads();                                                // S3
                DBG_STAMP(3, threadIdx.x == 0);
            }
            // ---- layer 3 as three chunks of 192 / 192 / 160 bins (the last one over-reads 16 rows of padding):
            // chunk 0 -> TMEM columns [0,192), chunk 1 -> [192,384), chunk 2 -> [0,160) once chunk 0 is drained
            // one thread issues chunk 0 and then chunk 1: issued from two warps at once the two GEMMs interleaved in the tensor
            // pipe and chunk 0 (the one the epilogue is waiting for) completed 0.6 k cycles later
            if (warp == 3 && lead) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w3_addr, NPAD * 128, 2, tmem, 192);
                umma_commit(bar3_0);
                issue_gemm2(a_addr, 16384, w3_addr + 192 * 128, NPAD * 128, 2, tmem + 192, 192);
                umma_commit(bar3_1);
            }

            // per thread 17 (h = 0) or 16 (h = 1) sub-chunks of 16 bins; the P / Vb quads of sub-chunk t+2 are requested
            // while sub-chunk t is evaluated (three rotating register buffers, everything statically indexed)
            // The TMEM reads are double-buffered: the 16 columns of sub-chunk t+1 are requested right after those of t have
            // arrived, so their latency runs under the arithmetic of t (with two warps per scheduler the other warp alone
            // cannot cover it: the single-buffer version spent half of the layer-3 phase with neither XU nor issue slots busy).
            float va[16], vb[16];
            mbar_wait(bar3_0, ph3_0, dead, p.status); ph3_0 ^= 1; tc_fence_after(); DBG_STAMP(15, threadIdx.x == 0);
            tmem_ld16(tmem + lane_off + MH2_COL(0), va);
#pragma unroll
            for (int t = 0; t < 17; ++t) {
                const bool live = (t < 16) || (h == 0);                         // bins 528..543 do not exist
                if (t + 2 < 17 && ((t + 2 < 16) || (h == 0))) {
                    if ((t + 2) % 3 == 0) MH2_LOAD(t + 2, pv0);
                    else if ((t + 2) % 3 == 1) MH2_LOAD(t + 2, pv1);
                    else MH2_LOAD(t + 2, pv2);
                }
                if (live) tmem_wait_ld();                                       // sub-chunk t is in registers
                if (t == 5) {                                                   // chunk 0 drained by this thread
                    DBG_STAMP(4, threadIdx.x == 0);
                    tc_fence_before();
                    mbar_arrive2(barf_0);
                }
                if (t + 1 < 17 && ((t + 1 < 16) || (h == 0))) {
                    if (t + 1 == 6) { mbar_wait(bar3_1, ph3_1, dead, p.status); ph3_1 ^= 1; tc_fence_after(); DBG_STAMP(16, threadIdx.x == 0); }
                    if (t + 1 == 12) { mbar_wait(bar3_0, ph3_0, dead, p.status); ph3_0 ^= 1; tc_fence_after(); DBG_STAMP(17, threadIdx.x == 0); }
                    if ((t + 1) % 2 == 0) tmem_ld16(tmem + 192 * (MH2_CH(t + 1) & 1) + lane_off + MH2_COL(t + 1), va);
                    else tmem_ld16(tmem + 192 * (MH2_CH(t + 1) & 1) + lane_off + MH2_COL(t + 1), vb);
                }
                if (live) {
                    if (t % 2 == 0) {
                        if (t % 3 == 0) loglik16_pv<POLY>(va, pv0, g_row, acc, accl);
                        else if (t % 3 == 1) loglik16_pv<POLY>(va, pv1, g_row, acc, accl);
                        else loglik16_pv<POLY>(va, pv2, g_row, acc, accl);
                    } else {
                        if (t % 3 == 0) loglik16_pv<POLY>(vb, pv0, g_row, acc, accl);
                        else if (t % 3 == 1) loglik16_pv<POLY>(vb, pv1, g_row, acc, accl);
                        else loglik16_pv<POLY>(vb, pv2, g_row, acc, accl);
                    }
                }
                if (t == 5) {
                    if (warp == 5) {
                        if (lead) {
                            mbar_wait(barf_0, phf_0, dead, p.status);
                            tc_fence_after();
                            issue_gemm2(a_addr, 16384, w3_addr + 384 * 128, NPAD * 128, 2, tmem, 160);
                            umma_commit(bar3_0);
                        }
                        __syncwarp();
                    }
                    phf_0 ^= 1;
                }
                if (t == 11) {
                    DBG_STAMP(5, threadIdx.x == 0);
                    // draws of the next proposal: requested here, consumed at the top of the next evaluation
                    const int nxt = it + 1;
                    if (nxt < n_iter && valid) {
                        const float4* e = reinterpret_cast<const float4*>(p.eps + ((int64_t)nxt * p.rows + row_g) * L + h * LH);
#pragma unroll
                        for (int l = 0; l < LH / 4; ++l) enn[l] = __ldg(e + l);
                        u_nxt = __ldg(p.u + (int64_t)nxt * p.rows + row_g);
                    }
                }
            }
#undef MH2_LOAD
#undef MH2_BIN
#undef MH2_COL
#undef MH2_CH
            tc_fence_before();
            const float part = fmaf(kLn2, accl, acc * kQuadScale);
            red2[h * TM + row] = make_float2(part, prior_h);
            DBG_STAMP(8, threadIdx.x == 0);
            __syncthreads();                                                    // S4: both halves of l(z') and of the prior term
            DBG_STAMP(9, threadIdx.x == 0);
            {
                const float2 r0 = red2[row], r1 = red2[TM + row];
                const float ll_prop = r0.x + r1.x;
                if (it < 0) {
                    ll_cur = ll_prop;
                } else if (valid) {
                    const float a = (ll_cur - ll_prop) + 0.5f * (r0.y + r1.y);
                    if (owner && p.a_trace) p.a_trace[(int64_t)it * p.rows + row_g] = a;
                    if (__logf(u_cur) < a) {
#pragma unroll
                        for (int l = 0; l < LH; ++l) zh[l] = zph[l];
                        ll_cur = ll_prop;
                        ++n_acc;
                    }
                    if (it >= p.n_burn) {
                        float4* dst = reinterpret_cast<float4*>(p.Zs + (row_g * p.n_keep + (it - p.n_burn)) * L + h * LH);
#pragma unroll
                        for (int l = 0; l < LH / 4; ++l) dst[l] = make_float4(zh[4 * l], zh[4 * l + 1], zh[4 * l + 2], zh[4 * l + 3]);
                    }
                }
            }
            DBG_STAMP(23, threadIdx.x == 0);
        }
        if (valid) {
#pragma unroll
            for (int l = 0; l < LH; ++l) p.Z[row_g * L + h * LH + l] = zh[l];
            if (owner && p.n_accept) p.n_accept[row_g] += n_acc;
        }
    }

    if (p.dbg && threadIdx.x == 0) {                    // per-CTA duration of the tile loop: min / max / sum over CTAs (slots 56-58)
        const long long dt = clock64() - dbg_t0;
        atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 56), (unsigned long long)dt);
        atomicMax(reinterpret_cast<unsigned long long*>(p.dbg + 57), (unsigned long long)dt);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.dbg + 58), (unsigned long long)dt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace tc
}  // namespace dvae

using namespace dvae;
using namespace dvae::tc;

extern "C" int dvae_mh_chain_tc2(const DvaeMlp* dec, const void* image, const void* PVpk, const float* g,
                                 const float* y, int y_dim, float* Z, float* Zs, int64_t NT, int L, int n_chains, int n_burn,
                                 int n_keep, float var_rw, const float* eps, const float* u, uint32_t* n_accept,
                                 float* a_trace, int flags, int* status, void* stream) {
    Mh2Params p{};
    int rc = check_dims(dec, L, y_dim, "dvae_mh_chain_tc2", &p.d);
    if (rc) return rc;
    DVAE_REQUIRE(L == 16 || L == 32, "dvae_mh_chain_tc2: latent size must be 16 or 32 (got %d)", L);
    DVAE_REQUIRE(y_dim <= 3, "dvae_mh_chain_tc2: at most 3 label inputs");
    DVAE_REQUIRE(image && PVpk && g && Z && Zs && eps && u && status, "dvae_mh_chain_tc2: null pointer");
    DVAE_REQUIRE(y_dim == 0 || y, "dvae_mh_chain_tc2: y_dim=%d but y is null", y_dim);
    DVAE_REQUIRE(NT >= 0 && n_chains >= 1 && n_chains < 4096 && n_burn >= 0 && n_keep >= 1 && var_rw > 0.f, "dvae_mh_chain_tc2: bad sizes");
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(Zs) & 15) == 0 && (reinterpret_cast<uintptr_t>(eps) & 15) == 0,
                 "dvae_mh_chain_tc2: Zs and eps must be 16-byte aligned");
    if (NT == 0) return 0;
    p.image = (const unsigned char*)image;
    p.rows = NT * n_chains; p.C = n_chains; p.y = y;
    p.PVpk = (const uint4*)PVpk; p.g = g; p.Z = Z; p.Zs = Zs;
    p.eps = eps; p.u = u;
    p.n_accept = n_accept; p.a_trace = a_trace; p.n_burn = n_burn; p.n_keep = n_keep;
    p.sd = sqrtf(var_rw);
    p.status = status;
    p.dbg = g_dbg_clocks;
    const size_t smem = smem_bytes(p.d) + (size_t)TM * L * 4;
    DVAE_REQUIRE(smem <= 227 * 1024, "dvae_mh_chain_tc2: shared memory budget exceeded");
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    const int grid = (int)(n_tiles < 148 ? n_tiles : 148);
    cudaStream_t st = (cudaStream_t)stream;
    const int poly = (flags & DVAE_TC_POLY_EX2_ALL) ? 2 : ((flags & DVAE_TC_POLY_EX2) ? 1 : 0);
#define MH2_LAUNCH(LL, PP)                                                                                   \
    do {                                                                                                     \
        cudaFuncSetAttribute(mh2_kernel<LL, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        mh2_kernel<LL, PP><<<grid, MH2_THREADS, smem, st>>>(p);                                              \
    } while (0)
    if (L == 16 && poly == 2) MH2_LAUNCH(16, 2);
    else if (L == 16 && poly == 1) MH2_LAUNCH(16, 1);
    else if (L == 16) MH2_LAUNCH(16, 0);
    else if (poly == 2) MH2_LAUNCH(32, 2);
    else if (poly == 1) MH2_LAUNCH(32, 1);
    else MH2_LAUNCH(32, 0);
#undef MH2_LAUNCH
    return check_launch("mh2_kernel");
}

// Debug aid: when a device buffer of 64 int64 is registered, CTA 0 of the sampler stamps clock64() at its phase
// boundaries during iteration 4 of its first tile (slot map in tools/tc_phase_clocks.py).  Pass NULL to disable.
extern "C" int dvae_debug_set_clock_buffer(void* dev_buffer) {
    g_dbg_clocks = reinterpret_cast<long long*>(dev_buffer);
    return 0;
}
