// tcgen05 decode of the kept samples fused with the W-update statistics of the M-step.
//
// Replaces packages/models/mcem.py:280-290 (compute_Vs) and the reductions of mcem.py:108-110 (numerator and
// denominator of the W update) in one pass:
//
//   Vs[n][r][f]   = D([Z_r(n); y(n)])[f]                                   (written once, coalesced, FP32)
//   num[u][k][f]  = sum_{n in u} P[n][f] * (sum_r Vx^-2) * H[n][k]
//   den[u][k][f]  = sum_{n in u}           (sum_r Vx^-1) * H[n][k],        Vx = g[n] Vs[n][r][f] + Vb[n][f]
//
// so the separate W pass (one more full read of Vs from HBM) disappears.  Layers 1 and 2 run as in mh_tc.cu (TMEM
// lanes = rows).  Layer 3 is issued TRANSPOSED: A = 128 bins of W3, B = the tile's hidden activations, so TMEM lanes
// are BINS and columns are the (frame, sample) rows of the tile.  A thread therefore owns one bin: the sum over the
// R samples of a frame and the accumulation over frames stay in its registers, and a warp's stores of Vs are 128
// consecutive bytes.  Both operand images are the K-major 128B-swizzled ones of the untransposed kernels.
//
// Work item = one part of one utterance (contiguous frames), fetched from an atomic counter by persistent CTAs;
// each item writes its own partial sums (fixed order inside an item -> deterministic results).
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int WS_KT = 10;                 // NMF ranks up to 10

struct WsParams {
    Dims d;
    const unsigned char* image;
    const float* Zs;        // [NT][R][L]
    const float* y;         // [NT][y_dim] or null
    const float* P;         // [NT][ld]
    const float* Vb;        // [NT][ld]
    const float* g;         // [NT]
    const float* H;         // [NT][K]
    const int64_t* fr_off;  // [B+1]
    float* Vs;              // [NT][R][ld]
    float* wstat;           // [B*n_parts][2][WS_KT][ld]
    int* work_counter;
    int* status;
    int B, n_parts, K, ld;
    long long* dbg;         // optional clock stamps of CTA 0 (third tile of its first item)
};

static long long* g_dbg_clocks_ws = nullptr;
#define DBGW(slot, cond) do { if (p.dbg && blockIdx.x == 0 && tile_no == 2 && (cond)) p.dbg[slot] = clock64(); } while (0)

template <int R, int NCH>
struct Acc { float num[NCH][WS_KT], den[NCH][WS_KT]; };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

template <int RL> __device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v);
template <> __device__ __forceinline__ void tmem_ld_n<32>(uint32_t taddr, float* v) { tmem_ld32(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld_n<16>(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// epilogue of one transposed chunk for the thread that owns bin f: all frames of the tile.
// vb / pw: the frame's noise variance and observation at bin f, loaded by the caller BEFORE it waited for the chunk.
template <int R, int RL, int FT>
__device__ __forceinline__ void chunk_epilogue(const WsParams& p, uint32_t tbuf, int f, bool fvalid, float bias, int64_t t0,
                                               int n_fr, const float* gS, const float* HS, const float (&vb)[FT],
                                               const float (&pw)[FT], float* num, float* den) {
#pragma unroll
    for (int fi = 0; fi < FT; ++fi) {
        if (fi < n_fr) {
            float v[RL];
            tmem_ld_n<RL>(tbuf + fi * R, v);
            tmem_wait_ld();
            if (fvalid) {
                const float gg = gS[fi];
                float* dst = p.Vs + ((t0 + fi) * R) * (int64_t)p.ld + f;
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int r = 0; r + 1 < R; r += 2) {                 // two samples share one reciprocal
                    const float s0 = ex2_approx(v[r] + bias), s1 = ex2_approx(v[r + 1] + bias);
                    dst[(int64_t)r * p.ld] = s0;
                    dst[(int64_t)(r + 1) * p.ld] = s1;
                    const float x0 = fmaf(gg, s0, vb[fi]), x1 = fmaf(gg, s1, vb[fi]);
                    const float rr = rcp_approx(x0 * x1);
                    const float i0 = x1 * rr, i1 = x0 * rr;
                    a1 += i0 + i1;
                    a2 = fmaf(i0, i0, fmaf(i1, i1, a2));
                }
                if (R & 1) {
                    const float s0 = ex2_approx(v[R - 1] + bias);
                    dst[(int64_t)(R - 1) * p.ld] = s0;
                    const float i0 = rcp_approx(fmaf(gg, s0, vb[fi]));
                    a1 += i0;
                    a2 = fmaf(i0, i0, a2);
                }
                const float pa2 = pw[fi] * a2;
#pragma unroll
                for (int k = 0; k < WS_KT; ++k) {
                    const float hk = HS[fi * WS_KT + k];
                    num[k] = fmaf(pa2, hk, num[k]);
                    den[k] = fmaf(a1, hk, den[k]);
                }
            }
        }
    }
}

template <int R, int L>
__global__ void __launch_bounds__(NTHREADS, 1) decode_ws_kernel(WsParams p) {
    constexpr int RL = (R <= 16) ? 16 : 32;               // TMEM load width covering one frame's samples
    constexpr int FT = 128 / R;                           // frames per tile
    static_assert(R >= 1 && R <= 32 && (FT - 1) * R + RL <= 128, "frame columns must stay inside the 128-column tile");
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bars[5];                          // layer 1/2, chunk ready x2, buffer free x2
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag, item_slot;
    __shared__ float gS[16], HS[16 * WS_KT];

    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Dims& d = p.d;
    unsigned char* A = base + ((d.image_bytes + 1023) & ~1023);
    const uint32_t bar12 = smem_u32(&bars[0]);
    const uint32_t bar3_0 = smem_u32(&bars[1]), bar3_1 = smem_u32(&bars[2]);
    const uint32_t barf_0 = smem_u32(&bars[3]), barf_1 = smem_u32(&bars[4]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = (warp >> 2) & 1;
    const int row = 32 * q + lane;
    const bool epi = warp < 8, owner = warp < 4, ctrl = (warp == 8 && lane == 0);
    uint32_t ph12 = 0, ph3 = 0, phf_0 = 0, phf_1 = 0;     // ph3: parity of THIS group's chunk-ready barrier
    const uint32_t my_bar3 = h ? bar3_1 : bar3_0, my_barf = h ? barf_1 : barf_0;

    {
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(base);
        for (int i = threadIdx.x; i < d.image_bytes / 16; i += NTHREADS) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) {
        dead_flag = 0;
        mbar_init(bar12, 1);
        mbar_init(bar3_0, 1);
        mbar_init(bar3_1, 1);
        mbar_init(barf_0, 128);
        mbar_init(barf_1, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
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
    const float* b3 = biasp + (d.n_hidden == 2 ? HID : 0);
    const int n_items = p.B * p.n_parts;
    const int y_dim = d.y_dim, nkb1 = d.nkb1;

    for (;;) {
        if (threadIdx.x == 0) item_slot = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int item = item_slot;
        __syncthreads();
        if (item >= n_items) break;
        const int u = item / p.n_parts, part = item - u * p.n_parts;
        const int64_t ua = p.fr_off[u], ub = p.fr_off[u + 1];
        // parts are runs of whole tiles
        const int64_t tiles_u = (ub - ua + FT - 1) / FT;
        const int64_t tiles_pp = (tiles_u + p.n_parts - 1) / p.n_parts;
        const int64_t n_lo = ua + (int64_t)part * tiles_pp * FT;
        int64_t n_hi = n_lo + tiles_pp * FT;
        if (n_hi > ub) n_hi = ub;

        float num[3][WS_KT], den[3][WS_KT];               // this thread's bins: chunks h, h+2 (and 4 for h == 0)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < WS_KT; ++k) { num[c][k] = 0.f; den[c][k] = 0.f; }

        int tile_no = 0;
        for (int64_t t0 = n_lo; t0 < n_hi; t0 += FT, ++tile_no) {
            DBGW(0, threadIdx.x == 0);
            const int n_fr = (int)((n_hi - t0 < FT) ? (n_hi - t0) : FT);
            // ---- stage g, H of the tile's frames; layer-1 operand rows
            if (threadIdx.x < FT * (WS_KT + 1)) {
                const int fi = threadIdx.x / (WS_KT + 1), k = threadIdx.x % (WS_KT + 1);
                if (k == WS_KT) gS[fi] = (fi < n_fr) ? p.g[t0 + fi] : 1.f;
                else HS[fi * WS_KT + k] = (fi < n_fr && k < p.K) ? p.H[(t0 + fi) * p.K + k] : 0.f;
            }
            if (owner) {
                const int fi = row / R, r = row - fi * R;
                const bool valid = fi < n_fr;
                float z[L];
                float y0 = 0.f, y1 = 0.f, y2 = 0.f;
                if (valid) {
                    const float4* src = reinterpret_cast<const float4*>(p.Zs + ((t0 + fi) * R + r) * (int64_t)L);
#pragma unroll
                    for (int l = 0; l < L / 4; ++l) {
                        const float4 t4 = __ldg(src + l);
                        z[4 * l] = t4.x; z[4 * l + 1] = t4.y; z[4 * l + 2] = t4.z; z[4 * l + 3] = t4.w;
                    }
                    if (y_dim > 0) y0 = p.y[(t0 + fi) * y_dim];
                    if (y_dim > 1) y1 = p.y[(t0 + fi) * y_dim + 1];
                    if (y_dim > 2) y2 = p.y[(t0 + fi) * y_dim + 2];
                } else {
#pragma unroll
                    for (int l = 0; l < L; ++l) z[l] = 0.f;
                }
                write_a1_static<L>(y_dim, nkb1, A, row, z, y0, y1, y2, valid);
            }
            fence_async_smem();
            __syncthreads();                                                    // S1
            DBGW(1, threadIdx.x == 0);
            if (ctrl) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w1_addr, 16384, d.nkb1, tmem, HID);
                umma_commit(bar12);
            }
            if (epi) {
                mbar_wait(bar12, ph12, dead, p.status);
                tc_fence_after();
                hidden_epilogue_rows(tmem, A, q, h, row, nullptr);
                fence_async_smem();
                tc_fence_before();
            }
            ph12 ^= 1;
            __syncthreads();                                                    // S2
            DBGW(2, threadIdx.x == 0);
            if (d.n_hidden == 2) {
                if (ctrl) {
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w2_addr, 16384, 2, tmem, HID);
                    umma_commit(bar12);
                }
                if (epi) {
                    mbar_wait(bar12, ph12, dead, p.status);
                    tc_fence_after();
                    hidden_epilogue_rows(tmem, A, q, h, row, b2);
                    fence_async_smem();
                    tc_fence_before();
                }
                ph12 ^= 1;
                __syncthreads();                                                // S3
                DBGW(3, threadIdx.x == 0);
            }

            // ---- layer 3, transposed: D^T[bins][rows] = W3[128 bins x 128] * h2[128 rows x 128]^T
            if (ctrl) {
                tc_fence_after();
                issue_gemm2(w3_addr, NPAD * 128, a_addr, 16384, 2, tmem + 128, 128);
                umma_commit(bar3_0);
                issue_gemm2(w3_addr + 16384, NPAD * 128, a_addr, 16384, 2, tmem + 256, 128);
                umma_commit(bar3_1);
                mbar_wait(barf_0, phf_0, dead, p.status); phf_0 ^= 1;       // group 0 drained chunk 0
                tc_fence_after();
                issue_gemm2(w3_addr + 2 * 16384, NPAD * 128, a_addr, 16384, 2, tmem + 128, 128);
                umma_commit(bar3_0);
                mbar_wait(barf_1, phf_1, dead, p.status); phf_1 ^= 1;       // group 1 drained chunk 1
                tc_fence_after();
                issue_gemm2(w3_addr + 3 * 16384, NPAD * 128, a_addr, 16384, 2, tmem + 256, 128);
                umma_commit(bar3_1);
                mbar_wait(barf_0, phf_0, dead, p.status); phf_0 ^= 1;       // group 0 drained chunk 2
                tc_fence_after();
                issue_gemm2(w3_addr + 4 * 16384, NPAD * 128, a_addr, 16384, 2, tmem + 128, 128);
                umma_commit(bar3_0);
                // consume the last "free" of each group so the parities line up for the next tile
                mbar_wait(barf_1, phf_1, dead, p.status); phf_1 ^= 1;       // after chunk 3
                mbar_wait(barf_0, phf_0, dead, p.status); phf_0 ^= 1;       // after chunk 4
            }
            if (epi) {
                const uint32_t tbuf = tmem + 128 + 128 * h + ((uint32_t)(32 * q) << 16);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int j = h + 2 * c;
                    if (j < 5) {
                        const int f = 128 * j + row;
                        const bool fvalid = f < d.F;
                        const float bias = (f < NPAD) ? b3[f] : 0.f;
                        float vbr[FT], pwr[FT];                              // requested before the chunk is waited for
#pragma unroll
                        for (int fi = 0; fi < FT; ++fi) {
                            const bool ok = fvalid && fi < n_fr;
                            vbr[fi] = ok ? __ldg(p.Vb + (t0 + fi) * p.ld + f) : 1.f;
                            pwr[fi] = ok ? __ldg(p.P + (t0 + fi) * p.ld + f) : 0.f;
                        }
                        mbar_wait(my_bar3, ph3, dead, p.status);
                        ph3 ^= 1;
                        tc_fence_after();
                        DBGW(10 + j, lane == 0 && q == 0);
                        chunk_epilogue<R, RL, FT>(p, tbuf, f, fvalid, bias, t0, n_fr, gS, HS, vbr, pwr, num[c], den[c]);
                        tc_fence_before();
                        mbar_arrive(my_barf);
                        DBGW(20 + j, lane == 0 && q == 0);
                    }
                }
            }
            __syncthreads();                                                    // tile done: A, gS, HS reusable
            DBGW(4, threadIdx.x == 0);
        }

        // ---- partial sums of this item: wstat[item][0|1][k][f]
        if (epi) {
            float* ws = p.wstat + (int64_t)item * 2 * WS_KT * p.ld;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int j = h + 2 * c;
                const int f = 128 * j + row;
                if (j < 5 && f < d.F) {
#pragma unroll
                    for (int k = 0; k < WS_KT; ++k) {
                        ws[k * p.ld + f] = num[c][k];
                        ws[(WS_KT + k) * p.ld + f] = den[c][k];
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

// W <- W * sqrt(sum_parts num / sum_parts den)   (mcem.py:111), from the statistics above
__global__ void __launch_bounds__(128) w_from_stats_kernel(const float* __restrict__ wstat, int n_parts, const float* __restrict__ W,
                                                           int F, int K, int ld, float* __restrict__ Wtmp) {
    const int u = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    for (int k = 0; k < K; ++k) {
        float num = 0.f, den = 0.f;
        for (int part = 0; part < n_parts; ++part) {
            const float* ws = wstat + ((int64_t)(u * n_parts + part) * 2 * WS_KT) * ld;
            num += ws[k * ld + f];
            den += ws[(WS_KT + k) * ld + f];
        }
        const int64_t i = ((int64_t)u * K + k) * ld + f;
        Wtmp[i] = W[i] * sqrtf(num / den);
    }
}

}  // namespace tc
}  // namespace dvae

using namespace dvae;
using namespace dvae::tc;

extern "C" int64_t dvae_decode_ws_workspace_floats(int B, int n_parts, int ld) {
    if (B <= 0 || n_parts <= 0 || ld <= 0) return 0;
    return (int64_t)B * n_parts * 2 * WS_KT * ld + 4;           // statistics + the work counter
}

extern "C" int dvae_decode_ws_tc(const DvaeMlp* dec, const void* image, const float* Zs, int R, int L, const float* y, int y_dim,
                                 const float* P, const float* Vb, const float* g, const float* H, int K, const int64_t* fr_off,
                                 int B, int64_t NT, int ld, float* Vs, float* ws, int n_parts, int* status, void* stream) {
    WsParams p{};
    int rc = check_dims(dec, L, y_dim, "dvae_decode_ws_tc", &p.d);
    if (rc) return rc;
    DVAE_REQUIRE(image && Zs && P && Vb && g && H && fr_off && Vs && ws && status, "dvae_decode_ws_tc: null pointer");
    DVAE_REQUIRE(R == 10 || R == 30, "dvae_decode_ws_tc: R must be 10 or 30 (got %d)", R);
    DVAE_REQUIRE(L == 16 || L == 32, "dvae_decode_ws_tc: latent size must be 16 or 32 (got %d)", L);
    DVAE_REQUIRE(y_dim <= 3 && (reinterpret_cast<uintptr_t>(Zs) & 15) == 0, "dvae_decode_ws_tc: y_dim <= 3 and 16-byte aligned Zs required");
    DVAE_REQUIRE(K >= 1 && K <= WS_KT, "dvae_decode_ws_tc: K must be <= %d", WS_KT);
    DVAE_REQUIRE(y_dim == 0 || y, "dvae_decode_ws_tc: labels missing");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && n_parts >= 1 && n_parts <= 16 && ld >= p.d.F, "dvae_decode_ws_tc: bad sizes");
    if (NT == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    p.image = (const unsigned char*)image;
    p.Zs = Zs; p.y = y; p.P = P; p.Vb = Vb; p.g = g; p.H = H; p.fr_off = fr_off; p.Vs = Vs; p.wstat = ws;
    p.work_counter = reinterpret_cast<int*>(ws + (int64_t)B * n_parts * 2 * WS_KT * ld);
    p.status = status; p.B = B; p.n_parts = n_parts; p.K = K; p.ld = ld;
    p.dbg = g_dbg_clocks_ws;
    cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    const size_t smem = smem_bytes(p.d);
    const int n_items = B * n_parts;
    const int grid = n_items < 148 ? n_items : 148;
    if (R == 10) {
        cudaFuncSetAttribute(decode_ws_kernel<10, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(decode_ws_kernel<10, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (L == 16) decode_ws_kernel<10, 16><<<grid, NTHREADS, smem, st>>>(p); else decode_ws_kernel<10, 32><<<grid, NTHREADS, smem, st>>>(p);
    } else {
        cudaFuncSetAttribute(decode_ws_kernel<30, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(decode_ws_kernel<30, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (L == 16) decode_ws_kernel<30, 16><<<grid, NTHREADS, smem, st>>>(p); else decode_ws_kernel<30, 32><<<grid, NTHREADS, smem, st>>>(p);
    }
    return check_launch("decode_ws_kernel");
}

extern "C" int dvae_nmf_w_from_stats(const float* ws, int n_parts, const float* W, int B, int F, int K, int ld, float* Wtmp,
                                     void* stream) {
    DVAE_REQUIRE(ws && W && Wtmp && B >= 1 && n_parts >= 1 && K >= 1 && K <= WS_KT && ld >= F, "dvae_nmf_w_from_stats: bad arguments");
    w_from_stats_kernel<<<dim3((F + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(ws, n_parts, W, F, K, ld, Wtmp);
    return check_launch("w_from_stats_kernel");
}

extern "C" int dvae_debug_set_clock_buffer_ws(void* dev_buffer) {
    g_dbg_clocks_ws = reinterpret_cast<long long*>(dev_buffer);
    return 0;
}
