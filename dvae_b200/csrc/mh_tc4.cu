// Fourth-generation tcgen05 Metropolis-Hastings sampler: front / back warp specialisation over two tiles, h2 in TMEM.
//
// Same math, operand images, P / Vb stream and draws as mh_tc2.cu.  The second generation evaluates one 128-chain tile
// as a chain of dependent phases (proposal -> layer 1 -> layer 2 -> three layer-3 chunks -> accept): 17.4 k cycles per
// evaluation with every pipe below 40 % (profiles/r01_tc_ncu_mh2.txt).  A second tile in flight would fill the gaps,
// but the resident weights (187 KB) leave room for ONE activation buffer only.  Here the layer-3 A operand (h2, BF16)
// is not staged in shared memory at all: the layer-2 epilogue stores it into tensor memory with tcgen05.st and the
// layer-3 MMAs read it from there (tcgen05.mma with a TMEM A operand; lane = row, 32-bit column c = (k = 2c | 2c+1 << 16),
// probed by tools/tmem_a_probe.cu).  That frees the pipeline:
//
//   front  (warps 0-7: TMEM lane         warps 0-3 (thread = chain) finalise the chain's previous evaluation (accept /
//           quadrant x column half)      keep) and propose; all eight run the layer-1/2 epilogues through the one
//                                        shared-memory buffer and park h2 in TMEM;
//   back   (warps 8-15, same split)      log-likelihood epilogue of the four layer-3 chunks (160 / 160 / 160 / 48 bins,
//                                        two TMEM buffers; the code of mh_tc2), partial sums to shared memory; lane 0 of
//                                        back warps 0-3 issues the MMAs of chunk 0-3 (no dedicated issue warp: 16 warps
//                                        keep the 128-register budget).
//
// The CTA alternates between two tiles: while `back` scores the proposal of tile A, `front` accepts / proposes for tile
// B.  The chain state lives in global memory (Z; L2-resident, 64 bytes per chain and phase), the previous proposal is
// rebuilt from z and its draws instead of being stored, so nothing per tile has to stay in registers.
// TMEM: [0,160) [160,320) layer-3 chunk buffers, [320,448) layers 1-2 accumulator, [448,512) h2.
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int MH4_THREADS = 512;             // 8 front + 8 back warps
constexpr int MH4_TM_D12 = 320, MH4_TM_H2 = 448;

struct Mh4Params {
    Dims d;
    const unsigned char* image;
    int64_t rows;                 // NT*C chains
    int C;
    const float* y;
    const uint4* PVpk;            // see pack_pv_kernel
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
    long long* dbg;
};

static long long* g_dbg_clocks4 = nullptr;
#define DBG4(slot, cond) do { if (p.dbg && blockIdx.x == 0 && j0 == 0 && ph == 9 && (cond)) p.dbg[slot] = clock64(); } while (0)

__device__ __forceinline__ void mh4_bar_front() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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

// front epilogue of a hidden layer for the thread's row and column half hh (64 hidden units = one K block = 32 packed
// columns): D12[row][64hh .. 64hh+64) -> tanh(+bias) in BF16.
// TO_TMEM = false: into the shared-memory operand (layer 1 of a two-hidden-layer decoder);
// TO_TMEM = true : into the TMEM h2 columns (the layer-3 A operand).
template <bool TO_TMEM>
__device__ __forceinline__ void mh4_hidden_row(uint32_t tmem, unsigned char* A, int q, int hh, int row, const float* bias) {
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    uint32_t w[32];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        float v[32];
        tmem_ld32(tmem + MH4_TM_D12 + lane_off + 64 * hh + 32 * part, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            float x0 = v[2 * e], x1 = v[2 * e + 1];
            if (bias) { x0 += bias[64 * hh + 32 * part + 2 * e]; x1 += bias[64 * hh + 32 * part + 2 * e + 1]; }
            w[16 * part + e] = tanh_bf16x2(pack_bf16x2(x0, x1));
        }
    }
    if (TO_TMEM) {
        tmem_st32(tmem + MH4_TM_H2 + lane_off + 32 * hh, w);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            *reinterpret_cast<uint4*>(A + hh * 16384 + row * 128 + ((cc ^ (row & 7)) << 4)) =
                make_uint4(w[4 * cc], w[4 * cc + 1], w[4 * cc + 2], w[4 * cc + 3]);
    }
}

template <int L, bool POLY>
__global__ void __launch_bounds__(MH4_THREADS, 1) mh4_kernel(Mh4Params p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bars[9];
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag;

    const Dims& d = p.d;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* A = base + ((d.image_bytes + 1023) & ~1023);          // layer-1 operand, then h1
    float* partS = reinterpret_cast<float*>(A + A_BYTES);                 // [slot][half][128] partial l(z')
    float* priorS = partS + 4 * TM;                                       // [slot][128]
    float* llS = priorS + 2 * TM;                                         // [slot][128] l(z) of the chain
    uint32_t* naccS = reinterpret_cast<uint32_t*>(llS + 2 * TM);          // [slot][128]
    const uint32_t bar12 = smem_u32(&bars[0]);
    const uint32_t h2_full = smem_u32(&bars[1]), h2_free = smem_u32(&bars[2]);
    const uint32_t c_ready0 = smem_u32(&bars[3]), c_ready1 = smem_u32(&bars[4]);
    const uint32_t c_free0 = smem_u32(&bars[5]), c_free1 = smem_u32(&bars[6]);
    const uint32_t b_done0 = smem_u32(&bars[7]), b_done1 = smem_u32(&bars[8]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    {
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(base);
        for (int i = threadIdx.x; i < d.image_bytes / 16; i += MH4_THREADS) dst[i] = __ldg(src + i);
        uint4* az = reinterpret_cast<uint4*>(A);
        for (int i = threadIdx.x; i < A_BYTES / 16; i += MH4_THREADS) az[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) {
        dead_flag = 0;
        mbar_init(bar12, 1);
        mbar_init(h2_full, 256); mbar_init(h2_free, 1);
        mbar_init(c_ready0, 1); mbar_init(c_ready1, 1);
        mbar_init(c_free0, 256); mbar_init(c_free1, 256);
        mbar_init(b_done0, 256); mbar_init(b_done1, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
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
    // this CTA's tiles: blockIdx.x + j * gridDim.x, processed in rounds of two (a last odd tile runs alone)
    const int my_tiles = (n_tiles > blockIdx.x) ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp < 8) {
        // =============================================== front ===============================================
        const int q = warp & 3, hh = warp >> 2, row = 32 * q + lane;
        const bool owner = hh == 0;                        // warps 0-3 carry the chain logic of row 32q + lane
        uint32_t ph12 = 0, ph_free = 0, ph_done0 = 0, ph_done1 = 0;
        bool first_store = true;
        for (int j0 = 0; j0 < my_tiles; j0 += 2) {
            const int T = (my_tiles - j0 >= 2) ? 2 : 1;
            const int n_ph = T * (n_iter + 2);
            if (owner) { naccS[row] = 0u; naccS[TM + row] = 0u; }
            if (p.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 100)) p.dbg[(blockIdx.x ? 50 : 40) + (j0 >> 1)] = clock64();
            for (int ph = 0; ph < n_ph; ++ph) {
                const int s = (T == 2) ? (ph & 1) : 0;
                const int e = (T == 2) ? (ph >> 1) - 1 : ph - 1;             // evaluation index: -1 scores the start state
                const int64_t tile = blockIdx.x + (int64_t)(j0 + s) * gridDim.x;
                const int64_t row_g = tile * TM + row;
                const bool valid = (row_g < p.rows) && owner;
                const int64_t fr = valid ? row_g / p.C : 0;
                DBG4(0, threadIdx.x == 0);
                // ---- global operands of this phase, requested before any wait
                float z[L];
                float4 ep[L / 4], en[L / 4];                                  // draws of proposal e-1 (to rebuild it) and e
                float u_prev = 0.5f, y0 = 0.f, y1 = 0.f, y2 = 0.f;
#pragma unroll
                for (int l = 0; l < L / 4; ++l) { ep[l] = make_float4(0.f, 0.f, 0.f, 0.f); en[l] = ep[l]; }
                if (valid) {
                    const float4* zsrc = reinterpret_cast<const float4*>(p.Z + row_g * L);
#pragma unroll
                    for (int l = 0; l < L / 4; ++l) { const float4 t4 = __ldcg(zsrc + l); z[4 * l] = t4.x; z[4 * l + 1] = t4.y; z[4 * l + 2] = t4.z; z[4 * l + 3] = t4.w; }
                    if (e >= 1) {
                        const float4* esrc = reinterpret_cast<const float4*>(p.eps + ((int64_t)(e - 1) * p.rows + row_g) * L);
#pragma unroll
                        for (int l = 0; l < L / 4; ++l) ep[l] = __ldg(esrc + l);
                        u_prev = __ldg(p.u + (int64_t)(e - 1) * p.rows + row_g);
                    }
                    if (e >= 0 && e < n_iter) {
                        const float4* esrc = reinterpret_cast<const float4*>(p.eps + ((int64_t)e * p.rows + row_g) * L);
#pragma unroll
                        for (int l = 0; l < L / 4; ++l) en[l] = __ldg(esrc + l);
                    }
                    if (y_dim > 0) y0 = p.y[fr * y_dim];
                    if (y_dim > 1) y1 = p.y[fr * y_dim + 1];
                    if (y_dim > 2) y2 = p.y[fr * y_dim + 2];
                } else {
#pragma unroll
                    for (int l = 0; l < L; ++l) z[l] = 0.f;
                }

                // ---- finalise evaluation e-1 of this tile (needs the back's partial sums)
                if (e >= 0 && owner) {
                    if (s == 0) { mbar_wait(b_done0, ph_done0, dead, p.status); ph_done0 ^= 1; }
                    else { mbar_wait(b_done1, ph_done1, dead, p.status); ph_done1 ^= 1; }
                    const float ll_prop = partS[(2 * s) * TM + row] + partS[(2 * s + 1) * TM + row];
                    if (e == 0) {
                        llS[s * TM + row] = ll_prop;
                    } else if (valid) {
                        const int it = e - 1;
                        const float a = (llS[s * TM + row] - ll_prop) + 0.5f * priorS[s * TM + row];
                        if (p.a_trace) p.a_trace[(int64_t)it * p.rows + row_g] = a;
                        if (__logf(u_prev) < a) {
#pragma unroll
                            for (int l = 0; l < L / 4; ++l) {
                                z[4 * l + 0] = __fadd_rn(z[4 * l + 0], __fmul_rn(p.sd, ep[l].x));
                                z[4 * l + 1] = __fadd_rn(z[4 * l + 1], __fmul_rn(p.sd, ep[l].y));
                                z[4 * l + 2] = __fadd_rn(z[4 * l + 2], __fmul_rn(p.sd, ep[l].z));
                                z[4 * l + 3] = __fadd_rn(z[4 * l + 3], __fmul_rn(p.sd, ep[l].w));
                            }
                            float4* zdst = reinterpret_cast<float4*>(p.Z + row_g * L);
#pragma unroll
                            for (int l = 0; l < L / 4; ++l) __stcg(zdst + l, make_float4(z[4 * l], z[4 * l + 1], z[4 * l + 2], z[4 * l + 3]));
                            llS[s * TM + row] = ll_prop;
                            naccS[s * TM + row] += 1u;
                        }
                        if (it >= p.n_burn) {
                            float4* dst = reinterpret_cast<float4*>(p.Zs + (row_g * p.n_keep + (it - p.n_burn)) * L);
#pragma unroll
                            for (int l = 0; l < L / 4; ++l) dst[l] = make_float4(z[4 * l], z[4 * l + 1], z[4 * l + 2], z[4 * l + 3]);
                        }
                    }
                }
                if (e == n_iter) {                                            // the tile is finished
                    if (valid && p.n_accept) p.n_accept[row_g] += naccS[s * TM + row];
                    continue;
                }

                // ---- operand of evaluation e: the start state (e = -1) or proposal e
                if (e >= 0 && owner) {
                    float prior = 0.f;
#pragma unroll
                    for (int l = 0; l < L / 4; ++l) {
                        const float z0 = z[4 * l], z1 = z[4 * l + 1], z2 = z[4 * l + 2], z3 = z[4 * l + 3];
                        z[4 * l + 0] = __fadd_rn(z0, __fmul_rn(p.sd, en[l].x));
                        z[4 * l + 1] = __fadd_rn(z1, __fmul_rn(p.sd, en[l].y));
                        z[4 * l + 2] = __fadd_rn(z2, __fmul_rn(p.sd, en[l].z));
                        z[4 * l + 3] = __fadd_rn(z3, __fmul_rn(p.sd, en[l].w));
                        prior += __fsub_rn(__fmul_rn(z0, z0), __fmul_rn(z[4 * l + 0], z[4 * l + 0]));
                        prior += __fsub_rn(__fmul_rn(z1, z1), __fmul_rn(z[4 * l + 1], z[4 * l + 1]));
                        prior += __fsub_rn(__fmul_rn(z2, z2), __fmul_rn(z[4 * l + 2], z[4 * l + 2]));
                        prior += __fsub_rn(__fmul_rn(z3, z3), __fmul_rn(z[4 * l + 3], z[4 * l + 3]));
                    }
                    priorS[s * TM + row] = prior;
                }
                if (owner) write_a1_static<L>(y_dim, nkb1, A, row, z, y0, y1, y2, valid);
                fence_async_smem();
                mh4_bar_front();
                DBG4(1, threadIdx.x == 0);
                if (threadIdx.x == 0) {
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w1_addr, 16384, nkb1, tmem + MH4_TM_D12, HID);
                    umma_commit(bar12);
                }
                mbar_wait(bar12, ph12, dead, p.status);
                ph12 ^= 1;
                tc_fence_after();
                if (two_hidden) {
                    mh4_hidden_row<false>(tmem, A, q, hh, row, nullptr);          // b1 rides on the constant-one column
                    fence_async_smem();
                    tc_fence_before();
                    mh4_bar_front();
                    DBG4(2, threadIdx.x == 0);
                    if (threadIdx.x == 0) {
                        tc_fence_after();
                        issue_gemm2(a_addr, 16384, w2_addr, 16384, 2, tmem + MH4_TM_D12, HID);
                        umma_commit(bar12);
                    }
                    mbar_wait(bar12, ph12, dead, p.status);
                    ph12 ^= 1;
                    tc_fence_after();
                }
                // the layer-3 MMAs of the previous phase must have finished reading h2
                if (!first_store) { mbar_wait(h2_free, ph_free, dead, p.status); ph_free ^= 1; tc_fence_after(); }
                first_store = false;
                mh4_hidden_row<true>(tmem, A, q, hh, row, two_hidden ? b2 : nullptr);
                tc_fence_before();
                mbar_arrive2(h2_full);
                DBG4(3, threadIdx.x == 0);
                if (p.dbg && blockIdx.x == 0 && j0 == 0 && threadIdx.x == 0 && ph >= 60 && ph < 70) p.dbg[ph - 60 + 20 - 6] = clock64();
            }
        }
    } else {
        // =============================================== back ===============================================
        const int bw = warp - 8;
        const int q = bw & 3, h = bw >> 2;                 // TMEM lane quadrant, column half
        const int row = 32 * q + lane;
        const uint32_t lane_off = (uint32_t)(32 * q) << 16;
        uint32_t ph_r0 = 0, ph_r1 = 0, ph_full = 0;
        const bool lead = lane == 0;
        const uint32_t id160 = umma_idesc(160), id48 = umma_idesc(48);
        // layer-3 chunk c of the current phase: A = h2 in TMEM, B = bins [160c, 160c + N) of W3, D = chunk buffer c & 1
        auto issue_chunk = [&](int c) {
            tc_fence_after();
            const uint32_t dcol = tmem + 160 * (c & 1);
            const uint32_t idesc = (c < 3) ? id160 : id48;
            uint32_t accum = 0;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_ts(dcol, tmem + MH4_TM_H2 + 32 * kb + 8 * k,
                            umma_desc(w3_addr + (uint32_t)(160 * c) * 128 + kb * (NPAD * 128) + 32 * k), idesc, accum);
                    accum = 1;
                }
            umma_commit((c & 1) ? c_ready1 : c_ready0);
        };
        for (int j0 = 0; j0 < my_tiles; j0 += 2) {
            const int T = (my_tiles - j0 >= 2) ? 2 : 1;
            const int n_ph = T * (n_iter + 1);
            for (int ph = 0; ph < n_ph; ++ph) {
                const int s = (T == 2) ? (ph & 1) : 0;
                const int64_t tile = blockIdx.x + (int64_t)(j0 + s) * gridDim.x;
                const int64_t row_g = tile * TM + row;
                const bool valid = row_g < p.rows;
                const float g_row = valid ? __ldg(p.g + row_g / p.C) : 1.f;
                const uint4* PVt = p.PVpk + (tile * NQ) * TM + row;
                float acc = 0.f, accl = 0.f;
                uint4 pv0[4], pv1[4], pv2[4];
                // chunks 0 and 1: lane 0 of back warps 0 / 1, once h2 of this phase is in TMEM and the chunk buffer has been
                // drained (c_free completes twice per phase: the waits below are for the SECOND completion of the previous
                // phase -> parity 1; the mid-phase waits for chunks 2 / 3 use parity 0)
                if (bw < 2) {
                    if (lead) {
                        mbar_wait(h2_full, ph_full, dead, p.status);
                        mbar_wait(bw == 0 ? c_free0 : c_free1, 1u, dead, p.status);
                        issue_chunk(bw);
                    }
                    ph_full ^= 1;
                    __syncwarp();
                }
                // sub-chunk t (16 bins): chunk c = t / 5 (t < 15), else 3; chunks hold 160 / 160 / 160 / 48 bins;
                // column inside the chunk: h * 80 + 16 (t - 5c) for c < 3; chunk 3: h = 0 -> 0, 16; h = 1 -> 32
#define MH4_CH(t) ((t) < 15 ? (t) / 5 : 3)
#define MH4_COL(t) ((t) < 15 ? h * 80 + 16 * ((t) - 5 * MH4_CH(t)) : (h == 0 ? 16 * ((t) - 15) : 32))
#define MH4_BIN(t) (160 * MH4_CH(t) + MH4_COL(t))
#define MH4_LOAD(t, PV)                                                                              \
    do {                                                                                             \
        _Pragma("unroll") for (int qd = 0; qd < 4; ++qd) PV[qd] = __ldg(PVt + ((MH4_BIN(t) >> 2) + qd) * TM); \
    } while (0)
                MH4_LOAD(0, pv0);
                MH4_LOAD(1, pv1);
#pragma unroll
                for (int t = 0; t < 17; ++t) {
                    const int c = MH4_CH(t);
                    const bool live = (t < 16) || (h == 0);                     // h = 1 owns one sub-chunk of the last chunk
                    if (t + 2 < 17 && ((t + 2 < 16) || (h == 0))) {
                        if ((t + 2) % 3 == 0) MH4_LOAD(t + 2, pv0);
                        else if ((t + 2) % 3 == 1) MH4_LOAD(t + 2, pv1);
                        else MH4_LOAD(t + 2, pv2);
                    }
                    if (t == 0 || t == 10) { mbar_wait(c_ready0, ph_r0, dead, p.status); ph_r0 ^= 1; tc_fence_after(); DBG4(10 + c, threadIdx.x == 256); }
                    if (t == 5 || t == 15) { mbar_wait(c_ready1, ph_r1, dead, p.status); ph_r1 ^= 1; tc_fence_after(); DBG4(10 + c, threadIdx.x == 256); }
                    if (live) {
                        float v[16];
                        tmem_ld16(tmem + 160 * (c & 1) + lane_off + MH4_COL(t), v);
                        tmem_wait_ld();
                        if (t % 3 == 0) loglik16_pv<POLY>(v, pv0, g_row, acc, accl);
                        else if (t % 3 == 1) loglik16_pv<POLY>(v, pv1, g_row, acc, accl);
                        else loglik16_pv<POLY>(v, pv2, g_row, acc, accl);
                    }
                    if (t == 4 || t == 14) { tc_fence_before(); mbar_arrive2(c_free0); DBG4(20 + c, threadIdx.x == 256); }
                    if (t == 9 || t == 16) { tc_fence_before(); mbar_arrive2(c_free1); DBG4(20 + c, threadIdx.x == 256); }
                    if (t == 4 && bw == 2) {                                    // chunk 2 into buffer 0 once every thread drained chunk 0
                        if (lead) { mbar_wait(c_free0, 0u, dead, p.status); issue_chunk(2); }
                        __syncwarp();
                    }
                    if (t == 9 && bw == 3) {                                    // chunk 3 into buffer 1; afterwards nothing reads h2 any more
                        if (lead) { mbar_wait(c_free1, 0u, dead, p.status); issue_chunk(3); umma_commit(h2_free); }
                        __syncwarp();
                    }
                }
#undef MH4_LOAD
#undef MH4_BIN
#undef MH4_COL
#undef MH4_CH
                partS[(2 * s + h) * TM + row] = fmaf(kLn2, accl, acc * kQuadScale);
                mbar_arrive2(s ? b_done1 : b_done0);
                if (p.dbg && blockIdx.x == 0 && j0 == 0 && threadIdx.x == 256 && ph >= 60 && ph < 70) p.dbg[ph - 60 + 30] = clock64();
            }
        }
    }

    if (p.dbg && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 100)) p.dbg[(blockIdx.x ? 50 : 40) + 5] = clock64();
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

extern "C" int dvae_mh_chain_tc4(const DvaeMlp* dec, const void* image, const void* PVpk, const float* g,
                                 const float* y, int y_dim, float* Z, float* Zs, int64_t NT, int L, int n_chains, int n_burn,
                                 int n_keep, float var_rw, const float* eps, const float* u, uint32_t* n_accept,
                                 float* a_trace, int flags, int* status, void* stream) {
    Mh4Params p{};
    int rc = check_dims(dec, L, y_dim, "dvae_mh_chain_tc4", &p.d);
    if (rc) return rc;
    DVAE_REQUIRE(L == 16 || L == 32, "dvae_mh_chain_tc4: latent size must be 16 or 32 (got %d)", L);
    DVAE_REQUIRE(y_dim <= 3, "dvae_mh_chain_tc4: at most 3 label inputs");
    DVAE_REQUIRE(image && PVpk && g && Z && Zs && eps && u && status, "dvae_mh_chain_tc4: null pointer");
    DVAE_REQUIRE(y_dim == 0 || y, "dvae_mh_chain_tc4: y_dim=%d but y is null", y_dim);
    DVAE_REQUIRE(NT >= 0 && n_chains >= 1 && n_chains < 4096 && n_burn >= 0 && n_keep >= 1 && var_rw > 0.f, "dvae_mh_chain_tc4: bad sizes");
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(Zs) & 15) == 0 && (reinterpret_cast<uintptr_t>(eps) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(Z) & 15) == 0, "dvae_mh_chain_tc4: Z, Zs and eps must be 16-byte aligned");
    if (NT == 0) return 0;
    p.image = (const unsigned char*)image;
    p.rows = NT * n_chains; p.C = n_chains; p.y = y;
    p.PVpk = (const uint4*)PVpk; p.g = g; p.Z = Z; p.Zs = Zs;
    p.eps = eps; p.u = u;
    p.n_accept = n_accept; p.a_trace = a_trace; p.n_burn = n_burn; p.n_keep = n_keep;
    p.sd = sqrtf(var_rw);
    p.status = status;
    p.dbg = g_dbg_clocks4;
    const size_t smem = ((size_t)(p.d.image_bytes + 1023) & ~(size_t)1023) + A_BYTES + 10 * TM * 4 + 1024;
    DVAE_REQUIRE(smem <= 227 * 1024, "dvae_mh_chain_tc4: shared memory budget exceeded");
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    // an even number of tiles per CTA keeps both pipeline slots busy: prefer a grid that deals pairs
    int grid = (int)(n_tiles < 148 ? n_tiles : 148);
    cudaStream_t st = (cudaStream_t)stream;
    const bool poly = (flags & DVAE_TC_POLY_EX2) != 0;
#define MH4_LAUNCH(LL, PP)                                                                                   \
    do {                                                                                                     \
        cudaFuncSetAttribute(mh4_kernel<LL, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        mh4_kernel<LL, PP><<<grid, MH4_THREADS, smem, st>>>(p);                                              \
    } while (0)
    if (L == 16 && poly) MH4_LAUNCH(16, true);
    else if (L == 16) MH4_LAUNCH(16, false);
    else if (poly) MH4_LAUNCH(32, true);
    else MH4_LAUNCH(32, false);
#undef MH4_LAUNCH
    return check_launch("mh4_kernel");
}

extern "C" int dvae_debug_set_clock_buffer4(void* dev_buffer) {
    g_dbg_clocks4 = reinterpret_cast<long long*>(dev_buffer);
    return 0;
}
