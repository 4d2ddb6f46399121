// NMF noise model and Wiener filter (FP32): initialisation, Vb = WH, the multiplicative M-step, cost, Wiener masks.
//
// Replaces packages/models/mcem.py:36-58 (init), 76-83 (variance helpers), 91-153 (M_step), 69-71 (cost),
// 310-329 + 176-177 (Wiener filter).
//
// All kernels here are HBM/L2-bound passes over the materialised speech variances Vs[NT][R][ld].  Per EM
// iteration (after Vs was written by the decoder): the W pass reads Vs once; the H/g/cost kernel reads each
// frame's R rows three times back to back from one CTA (first from HBM, then from L1/L2).
//
// The order of operations follows the reference exactly: the W update uses the Vb kept from the
// previous iteration, Vb is refreshed after W and after H, W/H are renormalised WITHOUT refreshing Vb, g uses the
// un-normalised product, the cost uses the new g.
#include "tc_common.cuh"

namespace dvae {

constexpr int KMAX = DVAE_MAX_K;

// ----------------------------------------------------------------------------- init (mcem.py:42-44)
__global__ void nmf_init_kernel(uint32_t seed_lo, uint32_t seed_hi, const int32_t* __restrict__ utt_ids,
                                const int64_t* __restrict__ fr_off, int B, int64_t NT, int F, int K, int ld, float eps,
                                float* __restrict__ W, float* __restrict__ H, float* __restrict__ g) {
    const int64_t nW = (int64_t)B * K * ld, nH = NT * K;
    const int64_t total = nW + nH + NT;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < nW) {
            const int f = (int)(i % ld);
            const int k = (int)((i / ld) % K);
            const int u = (int)(i / ((int64_t)ld * K));
            float v = 0.f;
            if (f < F) {
                const Philox4 r = philox4x32_10((uint32_t)utt_ids[u], (uint32_t)(f * KMAX + k), 0xFFFFFFF0u, 0u, seed_lo, seed_hi);
                v = fmaxf(u01(r.x), eps);
            }
            W[i] = v;
        } else if (i < nW + nH) {
            const int64_t j = i - nW;
            const int k = (int)(j % K);
            const int64_t n = j / K;
            int lo = 0, hi = B;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fr_off[mid] <= n) lo = mid; else hi = mid; }
            const Philox4 r = philox4x32_10((uint32_t)utt_ids[lo], (uint32_t)((n - fr_off[lo]) * KMAX + k), 0xFFFFFFF1u, 0u,
                                            seed_lo, seed_hi);
            H[j] = fmaxf(u01(r.x), eps);
        } else {
            g[i - nW - nH] = 1.0f;
        }
    }
}

// ----------------------------------------------------------------------------- Vb = W H (mcem.py:82-83)
__global__ void __launch_bounds__(128) nmf_vb_kernel(const float* __restrict__ W, const float* __restrict__ H,
                                                     const int32_t* __restrict__ frame_utt, int64_t NT, int F, int K,
                                                     int ld, float* __restrict__ Vb) {
    const int64_t n = blockIdx.x;
    if (n >= NT) return;
    const float* Wu = W + (int64_t)frame_utt[n] * K * ld;
    float h[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) h[k] = (k < K) ? H[n * K + k] : 0.f;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) v = fmaf(Wu[k * ld + f], h[k], v);
        Vb[n * ld + f] = v;
    }
}

// ----------------------------------------------------------------------------- W update (mcem.py:108-111)
// grid (ceil(F/128), B); thread = bin f, loops over the utterance's frames and samples.
__global__ void __launch_bounds__(128) nmf_w_kernel(const float* __restrict__ P, const float* __restrict__ Vs, int R,
                                                    const float* __restrict__ W, const float* __restrict__ H,
                                                    const float* __restrict__ g, const float* __restrict__ Vb,
                                                    const int64_t* __restrict__ fr_off, int F, int K, int ld,
                                                    float* __restrict__ Wtmp) {
    const int u = blockIdx.y;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float num[KMAX], den[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    for (int64_t n = n0; n < n1; ++n) {
        const float gg = g[n], vb = Vb[n * ld + f];
        const float* v = Vs + (n * R) * (int64_t)ld + f;
        float a1 = 0.f, a2 = 0.f;
        for (int r = 0; r < R; ++r) {
            const float vx = __fadd_rn(__fmul_rn(gg, v[(int64_t)r * ld]), vb);
            const float inv = __frcp_rn(vx);
            a1 += inv;
            a2 += __fmul_rn(inv, inv);
        }
        const float pa2 = P[n * ld + f] * a2;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                const float h = H[n * K + k];
                num[k] = fmaf(pa2, h, num[k]);
                den[k] = fmaf(a1, h, den[k]);
            }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < K) {
            const int64_t i = ((int64_t)u * K + k) * ld + f;
            Wtmp[i] = W[i] * sqrtf(num[k] / den[k]);
        }
}

// column L1 norms of the updated W, normalised copy into W, cost accumulator reset (mcem.py:130-132)
__global__ void __launch_bounds__(256) nmf_norm_kernel(const float* __restrict__ Wtmp, int F, int K, int ld,
                                                       float* __restrict__ W, float* __restrict__ norm) {
    __shared__ float scratch[32];
    const int u = blockIdx.x;
    for (int k = 0; k < K; ++k) {
        const float* src = Wtmp + ((int64_t)u * K + k) * ld;
        float s = 0.f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) s += fabsf(src[f]);
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) norm[u * K + k] = s;
        float* dst = W + ((int64_t)u * K + k) * ld;
        for (int f = threadIdx.x; f < F; f += blockDim.x) dst[f] = src[f] / s;
    }
}

// ----------------------------------------------------------------------------- H, g updates + cost (mcem.py:119-153, 69-71)
// grid (ceil(max_frames/FPB), B); a CTA walks FPB consecutive frames of one utterance with W_new[u] in shared memory.
constexpr int FPB = 4;

__global__ void __launch_bounds__(256) nmf_hg_kernel(const float* __restrict__ P, const float* __restrict__ Vs, int R,
                                                     const float* __restrict__ Wtmp, const float* __restrict__ norm,
                                                     float* __restrict__ H, float* __restrict__ g, float* __restrict__ Vb,
                                                     double* __restrict__ cost_part, const int64_t* __restrict__ fr_off,
                                                     int F, int K, int ld) {
    extern __shared__ float Ws[];                 // [K][ld]
    __shared__ float red[2 * KMAX][8];
    __shared__ float hnew[KMAX];
    __shared__ float scratch[32];
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    const int64_t nb = n0 + (int64_t)blockIdx.x * FPB;
    if (nb >= n1) {
        if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = 0.0;
        return;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < K * ld; i += blockDim.x) Ws[i] = Wtmp[(int64_t)u * K * ld + i];
    __syncthreads();
    const double inv_count = 1.0 / ((double)R * (double)F * (double)(n1 - n0));
    double cost_acc = 0.0;

    for (int64_t n = nb; n < n1 && n < nb + FPB; ++n) {
        const float gg = g[n];
        const float* Vn = Vs + (n * R) * (int64_t)ld;
        float h[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) h[k] = (k < K) ? H[n * K + k] : 0.f;

        // ---- H update: Vb1 = W_new H_old
        float num[KMAX], den[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) num[k] = den[k] = 0.f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            float vb = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) vb = fmaf(Ws[k * ld + f], h[k], vb);
            float a1 = 0.f, a2 = 0.f;
            for (int r = 0; r < R; ++r) {
                const float vx = __fadd_rn(__fmul_rn(gg, Vn[(int64_t)r * ld + f]), vb);
                const float inv = __frcp_rn(vx);
                a1 += inv;
                a2 += __fmul_rn(inv, inv);
            }
            const float pa2 = P[n * ld + f] * a2;
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) {
                    const float w = Ws[k * ld + f];
                    num[k] = fmaf(w, pa2, num[k]);
                    den[k] = fmaf(w, a1, den[k]);
                }
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                const float a = warp_sum(num[k]), b = warp_sum(den[k]);
                if (lane == 0) { red[2 * k][wid] = a; red[2 * k + 1][wid] = b; }
            }
        __syncthreads();
        if (threadIdx.x < K) {
            float a = 0.f, b = 0.f;
            for (int w = 0; w < 8; ++w) { a += red[2 * threadIdx.x][w]; b += red[2 * threadIdx.x + 1][w]; }
            hnew[threadIdx.x] = H[n * K + threadIdx.x] * sqrtf(a / b);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KMAX; ++k) h[k] = (k < K) ? hnew[k] : 0.f;

        // ---- g update with Vb2 = W_new H_new (kept as the model's Vb)
        float ng = 0.f, dg = 0.f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            float vb = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) vb = fmaf(Ws[k * ld + f], h[k], vb);
            Vb[n * ld + f] = vb;
            float s2 = 0.f, s1 = 0.f;
            for (int r = 0; r < R; ++r) {
                const float v = Vn[(int64_t)r * ld + f];
                const float vx = __fadd_rn(__fmul_rn(gg, v), vb);
                const float inv = __frcp_rn(vx);
                s1 = fmaf(v, inv, s1);
                s2 = fmaf(v, __fmul_rn(inv, inv), s2);
            }
            ng = fmaf(P[n * ld + f], s2, ng);
            dg += s1;
        }
        ng = block_sum(ng, scratch);
        dg = block_sum(dg, scratch);
        const float gnew = gg * sqrtf(ng / dg);

        // ---- cost with the refreshed Vx = g_new Vs + Vb2
        float c = 0.f;
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const float vb = Vb[n * ld + f];          // written by this very thread above
            const float p = P[n * ld + f];
            for (int r = 0; r < R; ++r) {
                const float vx = __fadd_rn(__fmul_rn(gnew, Vn[(int64_t)r * ld + f]), vb);
                c += logf(vx) + __fdiv_rn(p, vx);
            }
        }
        c = block_sum(c, scratch);
        cost_acc += (double)c;

        if (threadIdx.x < K) H[n * K + threadIdx.x] = hnew[threadIdx.x] * norm[u * K + threadIdx.x];
        if (threadIdx.x == 0) g[n] = gnew;
        __syncthreads();
    }
    if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = cost_acc * inv_count;
}

// ----------------------------------------------------------------------------- shared constants of the staged-frame kernels
constexpr int HG2_KT = 10;                    // NMF ranks up to 10 take the fast paths
constexpr int HG3_THREADS = 256;
constexpr int HG3_FPB = 8;                    // frames per CTA
constexpr int HG3_NV = 2 * HG2_KT;

__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ----------------------------------------------------------------------------- H, g, cost: packed-math frame tile
// Dispatched for F = 513 with FP32 variances.  ncu on its scalar predecessor (profiles/r01_tc_ncu_hg3.txt) showed that kernel
// issue-bound: 68 % of the issue slots busy, 16 k warp instructions per frame of which 9 % staged the frame and 45 % were scalar
// FP32 / LDS of the three passes.  One frame in shared memory, three CTAs per SM, and:
//   * a thread owns the ADJACENT bins 2t, 2t+1: one 8-byte shared load fetches both, and the pair is processed with the
//     sm_100 packed instructions fma/mul/add.f32x2 (two FP32 lanes per issue slot);
//   * the frame is staged by the bulk-copy engine: lanes 0..R-1 of warp 0 each issue one cp.async.bulk row copy that
//     completes on an mbarrier, instead of sixteen 16-byte cp.async per thread;
//   * five barriers per frame instead of nine, 30 shuffles instead of 100 for the 20 H sums, cost reduced once per CTA.
constexpr int HG5_LD = 520;                     // row pitch the engine uses for F = 513 (compile-time: shared addresses become immediates)
__device__ __forceinline__ f32x2 rcp2(f32x2 a) { float lo, hi; upk2(a, lo, hi); return pk2(rcp_fast(lo), rcp_fast(hi)); }
__device__ __forceinline__ f32x2 lg22(f32x2 a) { float lo, hi; upk2(a, lo, hi); return pk2(lg2_fast(lo), lg2_fast(hi)); }
__device__ __forceinline__ f32x2 lds2(const float* p) { return *reinterpret_cast<const f32x2*>(p); }
// Power of two c with c * y in [1, 2) (y > 0), and log2 c.  The cost passes multiply FOUR samples of one bin before taking one
// reciprocal and one logarithm; scaled by the c of the bin's first sample the product stays inside FP32 as long as the samples of
// a bin lie within 2^+-31 of each other, whatever the absolute level of the variances (a fixed scale of 2^8 covered variances
// from 1e-10 to 1e7 only, and the NMF part of Vx does fall below that in spectral nulls).
__device__ __forceinline__ float bin_scale(float y, float& log2c) {
    int e = (int)((__float_as_uint(y) >> 23) & 255u);
    e = e > 253 ? 253 : e;
    log2c = (float)(127 - e);
    return __uint_as_float((uint32_t)(254 - e) << 23);
}

template <int R, int ld>
__global__ void __launch_bounds__(HG3_THREADS, 3) nmf_hg5_kernel(const float* __restrict__ P, const float* __restrict__ Vs,
                                                                 const float* __restrict__ Wtmp, const float* __restrict__ norm,
                                                                 float* __restrict__ H, float* __restrict__ g,
                                                                 float* __restrict__ Vb, double* __restrict__ cost_part,
                                                                 const int64_t* __restrict__ fr_off, int K) {
    constexpr int KT = HG2_KT;
    static_assert(R % 2 == 0 && R <= 32 && ld % 4 == 0 && ld >= 513, "hg5: even R up to 32, row pitch a multiple of 16 bytes");
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                                  // [R][ld] samples of the current frame
    float* red = S + R * ld;                        // [8][HG3_NV] per-warp partials of the H sums
    __shared__ float2 red2[8];
    __shared__ double redd[8];
    __shared__ float hs[KT];
    __shared__ __align__(8) unsigned long long bars2[2];
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    const int64_t nb = n0 + (int64_t)blockIdx.x * HG3_FPB;
    if (nb >= n1) {
        if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = 0.0;
        return;
    }
    const int64_t ne = (nb + HG3_FPB < n1) ? nb + HG3_FPB : n1;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int f2 = 2 * t;                           // bins 2t, 2t+1; sample t of bin 512 lives on threads t < R
    const bool xl = t < R;
    // The frame is staged in two parts, rows [0, RA) and [RA, R), each completing on its own mbarrier: part A of the NEXT
    // frame is requested as soon as the cost pass has left it, part B at the end of the frame, so most of the copy runs
    // under arithmetic (the single-buffer version spent 24 % of its warp time waiting for the frame, profiles/r01_tc_ncu_hg5.txt)
    constexpr int RA = (R >= 30) ? 16 : R;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bars2[0]), bar_b = (unsigned)__cvta_generic_to_shared(&bars2[1]);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_a), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_b), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    f32x2 w2[KT];
    float wX[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const float* wk = Wtmp + ((int64_t)u * K + k) * ld;
        w2[k] = (k < K) ? *reinterpret_cast<const f32x2*>(wk + f2) : 0ull;
        wX[k] = (k < K) ? wk[512] : 0.f;
    }
    double cost_d = 0.0;                            // per-thread partial, reduced once per CTA
    unsigned phase_a = 0, phase_b = 0;
    constexpr unsigned row_bytes = (unsigned)ld * 4u;
    auto issue_rows = [&](int64_t n, int lo, int hi, unsigned bar) {      // warp 0: bulk copies of rows [lo, hi) of frame n
        if (wid == 0) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(row_bytes * (unsigned)(hi - lo)) : "memory");
            __syncwarp();
            if (lane >= lo && lane < hi) {
                const float* src = Vs + (n * R + lane) * (int64_t)ld;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(S + lane * ld);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
            }
        }
    };
    auto wait_rows = [&](unsigned bar, unsigned& ph) {
        unsigned done = 0, spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(ph) : "memory");
            if (!done && ++spins > (1u << 22)) __trap();                // a lost copy must not hang the device
        }
        ph ^= 1;
    };
    __syncthreads();                                                    // mbarriers initialised
    issue_rows(nb, 0, RA, bar_a);
    if (RA < R) issue_rows(nb, RA, R, bar_b);

    for (int64_t n = nb; n < ne; ++n) {
        const float gg = g[n];
        const f32x2 gg2 = pk2(gg, gg);
        const f32x2 p2 = *reinterpret_cast<const f32x2*>(P + n * ld + f2);
        const float pX = P[n * ld + 512];
        float h[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) h[k] = (k < K) ? H[n * K + k] : 0.f;
        // ---- H update (Vb1 = W_new H_old)
        f32x2 vb2 = 0ull;
        float vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on threads of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        f32x2 a1 = 0ull, a2 = 0ull;
        wait_rows(bar_a, phase_a);
#pragma unroll 4
        for (int r = 0; r < RA; r += 2) {
            const f32x2 x0 = fma2(gg2, lds2(S + r * ld + f2), vb2), x1 = fma2(gg2, lds2(S + (r + 1) * ld + f2), vb2);
            const f32x2 rr = rcp2(mul2(x0, x1));
            const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
            a1 = add2(a1, add2(i0, i1));
            a2 = fma2(i0, i0, fma2(i1, i1, a2));
        }
        if (RA < R) wait_rows(bar_b, phase_b);
#pragma unroll
        for (int r = RA; r < R; r += 2) {
            const f32x2 x0 = fma2(gg2, lds2(S + r * ld + f2), vb2), x1 = fma2(gg2, lds2(S + (r + 1) * ld + f2), vb2);
            const f32x2 rr = rcp2(mul2(x0, x1));
            const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
            a1 = add2(a1, add2(i0, i1));
            a2 = fma2(i0, i0, fma2(i1, i1, a2));
        }
        float a1X = 0.f, a2X = 0.f;                 // bin 512: this thread's single sample
        if (xl) { const float ix = rcp_fast(fmaf(gg, S[t * ld + 512], vbX)); a1X = ix; a2X = ix * ix; }
        {
            const f32x2 q2 = mul2(p2, a2);
            const float qX = pX * a2X;
            // 20 warp sums with 30 shuffles: fold over lane bit 4 (each lane keeps half of the values), then bit 3, then a
            // butterfly over the remaining 8 lanes; lanes 0, 8, 16, 24 end up with five totals each
            const bool b4 = lane & 16, b3 = lane & 8;
            float a[10];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[10];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = 5 * hf + j;
                    float nlo, nhi, dlo, dhi;
                    upk2(mul2(w2[k], q2), nlo, nhi);
                    upk2(mul2(w2[k], a1), dlo, dhi);
                    v[2 * j] = nlo + nhi;
                    v[2 * j + 1] = dlo + dhi;
                }
                if (wid == 0) {                     // the other warps hold no sample of bin 512 (qX = a1X = 0 there)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        v[2 * j] = fmaf(wX[5 * hf + j], qX, v[2 * j]);
                        v[2 * j + 1] = fmaf(wX[5 * hf + j], a1X, v[2 * j + 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float send = b4 ? v[j] : v[5 + j], keep = b4 ? v[5 + j] : v[j];
                    a[5 * hf + j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            float o5[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float send = b3 ? a[j] : a[5 + j], keep = b3 ? a[5 + j] : a[j];
                o5[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1)
#pragma unroll
                for (int j = 0; j < 5; ++j) o5[j] += __shfl_xor_sync(0xffffffffu, o5[j], o);
            if ((lane & 7) == 0) {                  // value index = 10 * bit3 + 5 * bit4 + j
                float* dst = red + wid * HG3_NV + 10 * ((lane >> 3) & 1) + 5 * (lane >> 4);
#pragma unroll
                for (int j = 0; j < 5; ++j) dst[j] = o5[j];
            }
        }
        __syncthreads();
        if (t < K) {
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { num += red[w * HG3_NV + 2 * t]; den += red[w * HG3_NV + 2 * t + 1]; }
            hs[t] = H[n * K + t] * sqrtf(num / den);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KT; ++k) h[k] = hs[k];

        // ---- g update (Vb2 = W_new H_new, kept as the model's Vb)
        vb2 = 0ull; vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on threads of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        *reinterpret_cast<f32x2*>(Vb + n * ld + f2) = vb2;
        if (t == 0) Vb[n * ld + 512] = vbX;
        f32x2 s1 = 0ull, s2 = 0ull;
#pragma unroll 5
        for (int r = 0; r < R; r += 2) {
            const f32x2 v0 = lds2(S + r * ld + f2), v1 = lds2(S + (r + 1) * ld + f2);
            const f32x2 x0 = fma2(gg2, v0, vb2), x1 = fma2(gg2, v1, vb2);
            const f32x2 rr = rcp2(mul2(x0, x1));
            const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
            const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
            s1 = add2(s1, add2(t0, t1));
            s2 = fma2(t0, i0, fma2(t1, i1, s2));
        }
        float s1X = 0.f, s2X = 0.f;
        if (xl) { const float sx = S[t * ld + 512]; const float ix = rcp_fast(fmaf(gg, sx, vbX)); s1X = sx * ix; s2X = s1X * ix; }
        {
            float ps_lo, ps_hi, s1lo, s1hi;
            upk2(mul2(p2, s2), ps_lo, ps_hi);
            upk2(s1, s1lo, s1hi);
            const float v2 = warp_sum(fmaf(pX, s2X, ps_lo + ps_hi)), v1 = warp_sum(s1lo + s1hi + s1X);
            if (lane == 0) red2[wid] = make_float2(v2, v1);
        }
        __syncthreads();
        float t2 = 0.f, t1s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float2 rr = red2[w]; t2 += rr.x; t1s += rr.y; }
        const float gnew = gg * sqrtf(t2 / t1s);
        const f32x2 gn2 = pk2(gnew, gnew);

        // ---- cost with Vx = g_new Vs + Vb2: FOUR samples share one reciprocal and one log2.  With y_i = c Vx_i (c = 2^8 keeps
        //      the four-fold products inside FP32 for variances from 1e-10 to 1e7; it rides on g_new and Vb2 for free):
        //      sum_i log2 Vx_i = log2(y0 y1 y2 y3) - 32,  sum_i 1 / Vx_i = c ((y0 + y1) y2 y3 + (y2 + y3) y0 y1) / (y0 y1 y2 y3)
        float c_lo, c_hi, lc_lo, lc_hi;                 // per-bin scale (bin_scale) from the bin's first sample
        {
            float y_lo, y_hi;
            upk2(fma2(gn2, lds2(S + f2), vb2), y_lo, y_hi);
            c_lo = bin_scale(y_lo, lc_lo);
            c_hi = bin_scale(y_hi, lc_hi);
        }
        const f32x2 c2 = pk2(c_lo, c_hi);
        const f32x2 gc2 = mul2(gn2, c2), vc2 = mul2(vb2, c2);
        f32x2 cl = 0ull, cp = 0ull;
        constexpr int R4 = R & ~3;
        constexpr int RA4 = (RA < R) ? RA : 0;          // rows [0, RA4) are released to the next frame's copy half-way
        static_assert(RA4 % 4 == 0, "part A must hold whole quads");
#pragma unroll
        for (int r = 0; r < RA4; r += 4) {
            const f32x2 y0 = fma2(gc2, lds2(S + r * ld + f2), vc2), y1 = fma2(gc2, lds2(S + (r + 1) * ld + f2), vc2);
            const f32x2 y2 = fma2(gc2, lds2(S + (r + 2) * ld + f2), vc2), y3 = fma2(gc2, lds2(S + (r + 3) * ld + f2), vc2);
            const f32x2 p01 = mul2(y0, y1), p23 = mul2(y2, y3);
            const f32x2 m = mul2(p01, p23);
            cl = add2(cl, lg22(m));
            cp = fma2(fma2(add2(y0, y1), p23, mul2(add2(y2, y3), p01)), rcp2(m), cp);
        }
        float cX = 0.f;
        if (RA4 > 0) {
            // bin 512 of rows < RA4 before part A is handed over
            if (xl && t < RA4) { const float x0 = fmaf(gnew, S[t * ld + 512], vbX); cX = fmaf(0.6931471805599453f, lg2_fast(x0), pX * rcp_fast(x0)); }
            __syncthreads();                            // every thread has left rows [0, RA)
            if (n + 1 < ne) issue_rows(n + 1, 0, RA, bar_a);
        }
#pragma unroll
        for (int r = RA4; r < R4; r += 4) {
            const f32x2 y0 = fma2(gc2, lds2(S + r * ld + f2), vc2), y1 = fma2(gc2, lds2(S + (r + 1) * ld + f2), vc2);
            const f32x2 y2 = fma2(gc2, lds2(S + (r + 2) * ld + f2), vc2), y3 = fma2(gc2, lds2(S + (r + 3) * ld + f2), vc2);
            const f32x2 p01 = mul2(y0, y1), p23 = mul2(y2, y3);
            const f32x2 m = mul2(p01, p23);
            cl = add2(cl, lg22(m));
            cp = fma2(fma2(add2(y0, y1), p23, mul2(add2(y2, y3), p01)), rcp2(m), cp);
        }
#pragma unroll
        for (int r = R4; r < R; r += 2) {              // R is even: one pair left when R % 4 == 2
            const f32x2 y0 = fma2(gc2, lds2(S + r * ld + f2), vc2), y1 = fma2(gc2, lds2(S + (r + 1) * ld + f2), vc2);
            const f32x2 pr = mul2(y0, y1);
            cl = add2(cl, lg22(pr));
            cp = fma2(add2(y0, y1), rcp2(pr), cp);
        }
        if (xl && t >= RA4) { const float x0 = fmaf(gnew, S[t * ld + 512], vbX); cX = fmaf(0.6931471805599453f, lg2_fast(x0), pX * rcp_fast(x0)); }
        {
            float cl_lo, cl_hi, pc_lo, pc_hi;
            upk2(cl, cl_lo, cl_hi);
            upk2(mul2(mul2(p2, c2), cp), pc_lo, pc_hi);
            // every sample of a bin carries the bin's log2 c; the reciprocal sums carry a factor 1 / c
            cost_d += (double)(fmaf(0.6931471805599453f, fmaf(-(float)R, lc_lo, cl_lo) + fmaf(-(float)R, lc_hi, cl_hi), pc_lo + pc_hi) + cX);
        }

        if (t < K) H[n * K + t] = hs[t] * norm[u * K + t];
        if (t == 0) g[n] = gnew;
        __syncthreads();                                // the frame is fully consumed: the rest of the buffer may be refilled
        if (n + 1 < ne) {
            if (RA4 > 0) issue_rows(n + 1, RA, R, bar_b);
            else issue_rows(n + 1, 0, RA, bar_a);
        }
    }
    cost_d = warp_sum_d(cost_d);
    if (lane == 0) redd[wid] = cost_d;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += redd[w];
        cost_part[(int64_t)u * gridDim.x + blockIdx.x] = sum / ((double)R * 513.0 * (double)(n1 - n0));
    }
}

// ----------------------------------------------------------------------------- H, g, cost: windowed frame (many samples)
// Sixth version, for sample counts that do not fit shared memory (multi-chain runs: R = chains x kept samples, e.g. 160):
// the frame is streamed in windows of RW samples (30 or 10).  Each of the three dependent passes walks all windows and
// keeps its per-bin partial sums in registers, so the frame is read three times (the first time from HBM, later ones from
// L2 when it still holds the frame); everything else is hg5: adjacent-bin pairs in packed FP32, bulk-copy staging, pair /
// quad shared reciprocals, cost reduced once per CTA.
template <int RW, int ld>
__global__ void __launch_bounds__(HG3_THREADS, 3) nmf_hg6_kernel(const float* __restrict__ P, const float* __restrict__ Vs,
                                                                 const float* __restrict__ Wtmp, const float* __restrict__ norm,
                                                                 float* __restrict__ H, float* __restrict__ g,
                                                                 float* __restrict__ Vb, double* __restrict__ cost_part,
                                                                 const int64_t* __restrict__ fr_off, int K, int R) {
    constexpr int KT = HG2_KT;
    static_assert(RW % 2 == 0 && RW <= 32 && ld % 4 == 0 && ld >= 513, "hg6: even window up to 32 samples");
    extern __shared__ __align__(16) float sm[];
    float* S = sm;                                  // [RW][ld] samples of the current window
    float* red = S + RW * ld;                       // [8][HG3_NV] per-warp partials of the H sums
    __shared__ float2 red2[8];
    __shared__ double redd[8];
    __shared__ float hs[KT];
    __shared__ __align__(8) unsigned long long bar;
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    const int64_t nb = n0 + (int64_t)blockIdx.x * HG3_FPB;
    if (nb >= n1) {
        if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = 0.0;
        return;
    }
    const int64_t ne = (nb + HG3_FPB < n1) ? nb + HG3_FPB : n1;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int f2 = 2 * t;                           // bins 2t, 2t+1; sample t of a window's bin 512 lives on threads t < RW
    const bool xl = t < RW;
    const int NW = R / RW;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_a), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    f32x2 w2[KT];
    float wX[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const float* wk = Wtmp + ((int64_t)u * K + k) * ld;
        w2[k] = (k < K) ? *reinterpret_cast<const f32x2*>(wk + f2) : 0ull;
        wX[k] = (k < K) ? wk[512] : 0.f;
    }
    double cost_d = 0.0;
    unsigned phase = 0;
    constexpr unsigned row_bytes = (unsigned)ld * 4u;

    // stage window w of frame n (RW bulk row copies) and wait for it; contains the barrier that frees the buffer
    auto stage = [&](int64_t n, int w) {
        __syncthreads();
        if (wid == 0) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(row_bytes * RW) : "memory");
            __syncwarp();
            if (lane < RW) {
                const float* src = Vs + (n * R + (int64_t)w * RW + lane) * (int64_t)ld;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(S + lane * ld);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst), "l"(src), "r"(row_bytes), "r"(bar_a) : "memory");
            }
        }
        unsigned done = 0, spins = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
            if (!done && ++spins > (1u << 22)) __trap();
        }
        phase ^= 1;
    };

    for (int64_t n = nb; n < ne; ++n) {
        const float gg = g[n];
        const f32x2 gg2 = pk2(gg, gg);
        const f32x2 p2 = *reinterpret_cast<const f32x2*>(P + n * ld + f2);
        const float pX = P[n * ld + 512];
        float h[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) h[k] = (k < K) ? H[n * K + k] : 0.f;

        // ---- H update (Vb1 = W_new H_old)
        f32x2 vb2 = 0ull;
        float vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on threads of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        f32x2 a1 = 0ull, a2 = 0ull;
        float a1X = 0.f, a2X = 0.f;
        for (int w = 0; w < NW; ++w) {
            stage(n, w);
#pragma unroll 5
            for (int r = 0; r < RW; r += 2) {
                const f32x2 x0 = fma2(gg2, lds2(S + r * ld + f2), vb2), x1 = fma2(gg2, lds2(S + (r + 1) * ld + f2), vb2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                a1 = add2(a1, add2(i0, i1));
                a2 = fma2(i0, i0, fma2(i1, i1, a2));
            }
            if (xl) { const float ix = rcp_fast(fmaf(gg, S[t * ld + 512], vbX)); a1X += ix; a2X = fmaf(ix, ix, a2X); }
        }
        {
            const f32x2 q2 = mul2(p2, a2);
            const float qX = pX * a2X;
            const bool b4 = lane & 16, b3 = lane & 8;
            float a[10];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[10];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = 5 * hf + j;
                    float nlo, nhi, dlo, dhi;
                    upk2(mul2(w2[k], q2), nlo, nhi);
                    upk2(mul2(w2[k], a1), dlo, dhi);
                    v[2 * j] = nlo + nhi;
                    v[2 * j + 1] = dlo + dhi;
                }
                if (wid == 0) {                     // the other warps hold no sample of bin 512 (qX = a1X = 0 there)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        v[2 * j] = fmaf(wX[5 * hf + j], qX, v[2 * j]);
                        v[2 * j + 1] = fmaf(wX[5 * hf + j], a1X, v[2 * j + 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float send = b4 ? v[j] : v[5 + j], keep = b4 ? v[5 + j] : v[j];
                    a[5 * hf + j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            float o5[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float send = b3 ? a[j] : a[5 + j], keep = b3 ? a[5 + j] : a[j];
                o5[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1)
#pragma unroll
                for (int j = 0; j < 5; ++j) o5[j] += __shfl_xor_sync(0xffffffffu, o5[j], o);
            if ((lane & 7) == 0) {
                float* dst = red + wid * HG3_NV + 10 * ((lane >> 3) & 1) + 5 * (lane >> 4);
#pragma unroll
                for (int j = 0; j < 5; ++j) dst[j] = o5[j];
            }
        }
        __syncthreads();
        if (t < K) {
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { num += red[w * HG3_NV + 2 * t]; den += red[w * HG3_NV + 2 * t + 1]; }
            hs[t] = H[n * K + t] * sqrtf(num / den);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KT; ++k) h[k] = hs[k];

        // ---- g update (Vb2 = W_new H_new, kept as the model's Vb)
        vb2 = 0ull; vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on threads of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        *reinterpret_cast<f32x2*>(Vb + n * ld + f2) = vb2;
        if (t == 0) Vb[n * ld + 512] = vbX;
        f32x2 s1 = 0ull, s2 = 0ull;
        float s1X = 0.f, s2X = 0.f;
        for (int w = 0; w < NW; ++w) {
            stage(n, w);
#pragma unroll 5
            for (int r = 0; r < RW; r += 2) {
                const f32x2 v0 = lds2(S + r * ld + f2), v1 = lds2(S + (r + 1) * ld + f2);
                const f32x2 x0 = fma2(gg2, v0, vb2), x1 = fma2(gg2, v1, vb2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                s1 = add2(s1, add2(t0, t1));
                s2 = fma2(t0, i0, fma2(t1, i1, s2));
            }
            if (xl) { const float sx = S[t * ld + 512]; const float ix = rcp_fast(fmaf(gg, sx, vbX)); const float tx = sx * ix; s1X += tx; s2X = fmaf(tx, ix, s2X); }
        }
        {
            float ps_lo, ps_hi, s1lo, s1hi;
            upk2(mul2(p2, s2), ps_lo, ps_hi);
            upk2(s1, s1lo, s1hi);
            const float v2 = warp_sum(fmaf(pX, s2X, ps_lo + ps_hi)), v1 = warp_sum(s1lo + s1hi + s1X);
            if (lane == 0) red2[wid] = make_float2(v2, v1);
        }
        __syncthreads();
        float t2 = 0.f, t1s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float2 rr = red2[w]; t2 += rr.x; t1s += rr.y; }
        const float gnew = gg * sqrtf(t2 / t1s);

        // ---- cost with Vx = g_new Vs + Vb2 (scaled by c = 2^8: four samples share one reciprocal and one log2)
        f32x2 c2 = 0ull, gn2c = 0ull, vc2 = 0ull;
        float lc_lo = 0.f, lc_hi = 0.f;
        f32x2 cl = 0ull, cp = 0ull;
        float cX = 0.f;
        constexpr int RW4 = RW & ~3;
        for (int w = 0; w < NW; ++w) {
            stage(n, w);
            if (w == 0) {                            // per-bin scale from the bin's first sample (bin_scale, see hg5)
                float y_lo, y_hi;
                upk2(fma2(pk2(gnew, gnew), lds2(S + f2), vb2), y_lo, y_hi);
                c2 = pk2(bin_scale(y_lo, lc_lo), bin_scale(y_hi, lc_hi));
                gn2c = mul2(pk2(gnew, gnew), c2);
                vc2 = mul2(vb2, c2);
            }
#pragma unroll
            for (int r = 0; r < RW4; r += 4) {
                const f32x2 y0 = fma2(gn2c, lds2(S + r * ld + f2), vc2), y1 = fma2(gn2c, lds2(S + (r + 1) * ld + f2), vc2);
                const f32x2 y2 = fma2(gn2c, lds2(S + (r + 2) * ld + f2), vc2), y3 = fma2(gn2c, lds2(S + (r + 3) * ld + f2), vc2);
                const f32x2 p01 = mul2(y0, y1), p23 = mul2(y2, y3);
                const f32x2 m = mul2(p01, p23);
                cl = add2(cl, lg22(m));
                cp = fma2(fma2(add2(y0, y1), p23, mul2(add2(y2, y3), p01)), rcp2(m), cp);
            }
#pragma unroll
            for (int r = RW4; r < RW; r += 2) {
                const f32x2 y0 = fma2(gn2c, lds2(S + r * ld + f2), vc2), y1 = fma2(gn2c, lds2(S + (r + 1) * ld + f2), vc2);
                const f32x2 pr = mul2(y0, y1);
                cl = add2(cl, lg22(pr));
                cp = fma2(add2(y0, y1), rcp2(pr), cp);
            }
            if (xl) { const float x0 = fmaf(gnew, S[t * ld + 512], vbX); cX += fmaf(0.6931471805599453f, lg2_fast(x0), pX * rcp_fast(x0)); }
        }
        {
            float cl_lo, cl_hi, pc_lo, pc_hi;
            upk2(cl, cl_lo, cl_hi);
            upk2(mul2(mul2(p2, c2), cp), pc_lo, pc_hi);
            cost_d += (double)(fmaf(0.6931471805599453f, fmaf(-(float)R, lc_lo, cl_lo) + fmaf(-(float)R, lc_hi, cl_hi), pc_lo + pc_hi) + cX);
        }
        __syncthreads();                            // hs and the reduction buffers are free for the next frame
        if (t < K) H[n * K + t] = hs[t] * norm[u * K + t];
        if (t == 0) g[n] = gnew;
        __syncthreads();                                // the frame is fully consumed: its buffer may be refilled
    }
    cost_d = warp_sum_d(cost_d);
    if (lane == 0) redd[wid] = cost_d;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += redd[w];
        cost_part[(int64_t)u * gridDim.x + blockIdx.x] = sum / ((double)R * 513.0 * (double)(n1 - n0));
    }
}

// ----------------------------------------------------------------------------- H, g, cost on the sampler's BF16 emission
// The M-step's frame kernel on the variances the tcgen05 sampler wrote itself (vst.cu: VsT[tile][slot][bin group][row][8 words]
// + the per-sample slot index; values without the decoder's output-layer bias E[f] = exp(b3[f])).  ncu on the first version
// (hg5's schedule reading the frame from shared memory in every pass: profiles/r02_ncu_hg7_v1.txt) showed it issue-bound by
// its own overhead: 12 % of all warp instructions unpacked BF16 pairs, 9 % computed cp.async addresses, 8 % were the three
// passes' shared-memory loads.  Here:
//   * a thread owns the adjacent bins 2t, 2t+1 = ONE 32-bit word per sample (consecutive lanes, consecutive banks);
//   * the packed FP32 lanes are two SAMPLES of one bin, not two bins of one sample: the odd bin's pair (w_r, w_r+1) is the raw
//     words themselves (adjacent registers, no instruction), the even bin's pair is two shifts;
//   * the pair trick (two variances share one reciprocal) runs between sample pairs (r, r+1) x (r+2, r+3);
//   * two shared-memory buffers of R x 1 056 bytes (together hg5's footprint: three CTAs per SM): the next frame is staged
//     with 16-byte cp.async (a warp copies whole rows: one slot lookup and one 64-bit address per row) under the current
//     frame's arithmetic.  A variant that kept the frame's words in registers for the three passes (60 shared-memory loads
//     fewer per thread and frame, but 128 registers = two CTAs per SM) was not faster: the kernel is bound by latency between
//     its five barriers per frame, i.e. by the number of resident warps (profiles/r02_ncu_v2.txt);
//   * E rides on g in the variance (Vx = (g E) v + Vb) and multiplies the g-update sums once per frame.
constexpr int HG7_ROWW = 264;                   // 32-bit words per staged row (528 bins)
__device__ __forceinline__ void hg7_cp16(unsigned smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gmem_src) : "memory");
}

template <int R>
__global__ void __launch_bounds__(HG3_THREADS, 3) nmf_hg7_kernel(const float* __restrict__ P, const uint4* __restrict__ VsT,
                                                                 const uint8_t* __restrict__ vs_idx, const float* __restrict__ bias_log2,
                                                                 const float* __restrict__ Wtmp, const float* __restrict__ norm,
                                                                 float* __restrict__ H, float* __restrict__ g,
                                                                 float* __restrict__ Vb, double* __restrict__ cost_part,
                                                                 const int64_t* __restrict__ fr_off, int K) {
    constexpr int KT = HG2_KT;
    constexpr int ld = HG5_LD;
    constexpr int NBG = 33, TMR = 128;
    constexpr int NPC = (R * 66 + HG3_THREADS - 1) / HG3_THREADS;     // 16-byte pieces per thread and frame
    static_assert(R % 2 == 0 && R <= 30, "hg7: even R up to 30");
    extern __shared__ __align__(16) uint32_t smw[];
    uint32_t* S0 = smw;                              // [2][R][HG7_ROWW] the current frame and the next one being staged
    float* red = reinterpret_cast<float*>(smw + 2 * R * HG7_ROWW);      // [8][HG3_NV] per-warp partials of the H sums
    __shared__ float2 red2[8];
    __shared__ double redd[8];
    __shared__ __align__(16) float hs[12];          // KT = 10 activations, read back as two 16-byte and one 8-byte load
    __shared__ __align__(16) float hs_old[12];
    static_assert(KT == 10, "hg7 reads its ten activations as float4 + float4 + float2");
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    // blocks of HG3_FPB = 8 frames aligned in the GLOBAL frame index (the first and last block of an utterance may be shorter): a
    // block's rows then fill whole 64-byte pairs of sectors of every (slot, bin group) of the emission - with blocks counted from
    // the utterance's first frame (185 frames per utterance: odd offsets) a block straddled five pairs instead of four
    static_assert(HG3_FPB == 8, "aligned blocks of eight frames");
    const int64_t nb0 = ((n0 >> 3) + (int64_t)blockIdx.x) << 3;
    const int64_t nb = nb0 > n0 ? nb0 : n0;
    const int64_t ne = (nb0 + HG3_FPB < n1) ? nb0 + HG3_FPB : n1;
    if (nb >= ne) {
        if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = 0.0;
        return;
    }
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int f2 = 2 * t;                           // bins 2t, 2t+1; sample t of bin 512 lives on threads t < R
    const bool xl = t < R;
    f32x2 w2[KT];
    float wX[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const float* wk = Wtmp + ((int64_t)u * K + k) * ld;
        w2[k] = (k < K) ? *reinterpret_cast<const f32x2*>(wk + f2) : 0ull;
        wX[k] = (k < K) ? wk[512] : 0.f;
    }
    const float E0 = exp2f(__ldg(bias_log2 + f2)), E1 = exp2f(__ldg(bias_log2 + f2 + 1)), EX = exp2f(__ldg(bias_log2 + 512));
    double cost_d = 0.0;                            // per-thread partial, reduced once per CTA

    // staging: warp w copies the rows w, w + 8, w + 16, w + 24 of the frame; a row is 66 pieces of 16 bytes = lanes 0..31 twice
    // plus lanes 0, 1 once more.  Piece i of a row sits (i >> 1) cells of 4 KB + 16 (i & 1) bytes behind the row's first cell.
    const unsigned s_base = (unsigned)__cvta_generic_to_shared(S0);
    const unsigned po0 = (unsigned)((lane >> 1) * (TMR * 32) + 16 * (lane & 1));           // piece `lane`; piece lane + 32 is 16 cells on
    auto stage = [&](int64_t n, int b) {
        const int64_t tile = n / TMR;
        const int row = (int)(n - tile * TMR);
        const uint8_t* ib = vs_idx + n * 32;
        const unsigned char* src0 = reinterpret_cast<const unsigned char*>(VsT) + ((size_t)tile * (R + 1) * NBG * TMR + row) * 32;
#pragma unroll
        for (int k = 0; k < (R + 7) / 8; ++k) {
            const int r = wid + 8 * k;
            if (r < R) {
                const unsigned char* src = src0 + (size_t)__ldg(ib + r) * (NBG * TMR * 32) + po0;
                const unsigned dst = s_base + (unsigned)((b * R + r) * HG7_ROWW * 4 + 16 * lane);
                hg7_cp16(dst, src);
                hg7_cp16(dst + 512, src + 16 * (TMR * 32));
                if (lane < 2) hg7_cp16(dst + 1024, src + 32 * (TMR * 32));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(nb, 0);
    if (t >= K && t < KT) { hs_old[t] = 0.f; hs[t] = 0.f; }

    int buf = 0;
    for (int64_t n = nb; n < ne; ++n, buf ^= 1) {
        // per-frame scalars first: their global-memory latency runs under the wait for the staged frame
        const float gg = g[n];
        const f32x2 p2 = *reinterpret_cast<const f32x2*>(P + n * ld + f2);
        const float pX = P[n * ld + 512];
        if (t < K) hs_old[t] = H[n * K + t];                            // the frame's activations reach every thread through shared memory
        asm volatile("cp.async.wait_group 0;" ::: "memory");            // this thread's pieces of frame n (requested one frame earlier)
        __syncthreads();                                                // frame n is staged for every thread, and every thread is done with frame n - 1:
        if (n + 1 < ne) stage(n + 1, buf ^ 1);                          // only now may its buffer be refilled (the cost pass reads it to the end)
        float h[KT];
        {
            const float4 h0 = *reinterpret_cast<const float4*>(hs_old), h1 = *reinterpret_cast<const float4*>(hs_old + 4);
            const float2 h2 = *reinterpret_cast<const float2*>(hs_old + 8);
            h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w; h[8] = h2.x; h[9] = h2.y;
        }
        const uint32_t* S = S0 + buf * R * HG7_ROWW + t;                // this thread's word of sample 0
        const float sX = xl ? vst_lo(S0[(buf * R + t) * HG7_ROWW + 256]) : 0.f;     // bin 512: this thread's single sample (without E)

        float p_lo, p_hi;
        upk2(p2, p_lo, p_hi);
        const f32x2 ga = pk2(gg * E0, gg * E0), gb = pk2(gg * E1, gg * E1);     // (g E) of bin 2t / 2t+1 for both sample lanes
        const float ggX = gg * EX;
        // even bin (a): lanes = samples (r, r+1) from two shifts; odd bin (b): the raw words
#define HG7_VA(r) pk2(vst_lo(S[(r) * HG7_ROWW]), vst_lo(S[((r) + 1) * HG7_ROWW]))
#define HG7_VB(r) pk2(vst_hi(S[(r) * HG7_ROWW]), vst_hi(S[((r) + 1) * HG7_ROWW]))
        // ---- H update (Vb1 = W_new H_old)
        f32x2 vb2 = 0ull;
        float vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on the threads t < R of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        float vb_lo, vb_hi;
        upk2(vb2, vb_lo, vb_hi);
        f32x2 va2 = pk2(vb_lo, vb_lo), vbb2 = pk2(vb_hi, vb_hi);
        f32x2 a1a = 0ull, a2a = 0ull, a1b = 0ull, a2b = 0ull;
        constexpr int R4 = R & ~3;
#pragma unroll
        for (int r = 0; r < R4; r += 4) {
            {
                const f32x2 x0 = fma2(ga, HG7_VA(r), va2), x1 = fma2(ga, HG7_VA(r + 2), va2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                a1a = add2(a1a, add2(i0, i1));
                a2a = fma2(i0, i0, fma2(i1, i1, a2a));
            }
            {
                const f32x2 x0 = fma2(gb, HG7_VB(r), vbb2), x1 = fma2(gb, HG7_VB(r + 2), vbb2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                a1b = add2(a1b, add2(i0, i1));
                a2b = fma2(i0, i0, fma2(i1, i1, a2b));
            }
        }
        if (R4 < R) {                               // R % 4 == 2: one sample pair left, one reciprocal per lane
            const f32x2 ia = rcp2(fma2(ga, HG7_VA(R4), va2)), ibb = rcp2(fma2(gb, HG7_VB(R4), vbb2));
            a1a = add2(a1a, ia); a2a = fma2(ia, ia, a2a);
            a1b = add2(a1b, ibb); a2b = fma2(ibb, ibb, a2b);
        }
        f32x2 a1, a2;                               // back to (bin 2t, bin 2t+1)
        {
            float l0, l1, m0, m1;
            upk2(a1a, l0, l1); upk2(a1b, m0, m1);
            a1 = pk2(l0 + l1, m0 + m1);
            upk2(a2a, l0, l1); upk2(a2b, m0, m1);
            a2 = pk2(l0 + l1, m0 + m1);
        }
        float a1X = 0.f, a2X = 0.f;
        if (xl) { const float ix = rcp_fast(fmaf(ggX, sX, vbX)); a1X = ix; a2X = ix * ix; }
        {
            const f32x2 q2 = mul2(p2, a2);
            const float qX = pX * a2X;
            // 20 warp sums with 30 shuffles: fold over lane bit 4 (each lane keeps half of the values), then bit 3, then a
            // butterfly over the remaining 8 lanes; lanes 0, 8, 16, 24 end up with five totals each
            const bool b4 = lane & 16, b3 = lane & 8;
            float a[10];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[10];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = 5 * hf + j;
                    float nlo, nhi, dlo, dhi;
                    upk2(mul2(w2[k], q2), nlo, nhi);
                    upk2(mul2(w2[k], a1), dlo, dhi);
                    v[2 * j] = nlo + nhi;
                    v[2 * j + 1] = dlo + dhi;
                }
                if (wid == 0) {                     // the other warps hold no sample of bin 512 (qX = a1X = 0 there)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        v[2 * j] = fmaf(wX[5 * hf + j], qX, v[2 * j]);
                        v[2 * j + 1] = fmaf(wX[5 * hf + j], a1X, v[2 * j + 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float send = b4 ? v[j] : v[5 + j], keep = b4 ? v[5 + j] : v[j];
                    a[5 * hf + j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            float o5[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float send = b3 ? a[j] : a[5 + j], keep = b3 ? a[5 + j] : a[j];
                o5[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1)
#pragma unroll
                for (int j = 0; j < 5; ++j) o5[j] += __shfl_xor_sync(0xffffffffu, o5[j], o);
            if ((lane & 7) == 0) {                  // value index = 10 * bit3 + 5 * bit4 + j
                float* dst = red + wid * HG3_NV + 10 * ((lane >> 3) & 1) + 5 * (lane >> 4);
#pragma unroll
                for (int j = 0; j < 5; ++j) dst[j] = o5[j];
            }
        }
        __syncthreads();
        if (t < K) {
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { num += red[w * HG3_NV + 2 * t]; den += red[w * HG3_NV + 2 * t + 1]; }
            hs[t] = hs_old[t] * sqrtf(num / den);
        }
        __syncthreads();
        {
            const float4 h0 = *reinterpret_cast<const float4*>(hs), h1 = *reinterpret_cast<const float4*>(hs + 4);
            const float2 h2 = *reinterpret_cast<const float2*>(hs + 8);
            h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w; h[8] = h2.x; h[9] = h2.y;
        }

        // ---- g update (Vb2 = W_new H_new, kept as the model's Vb)
        vb2 = 0ull; vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        *reinterpret_cast<f32x2*>(Vb + n * ld + f2) = vb2;
        if (t == 0) Vb[n * ld + 512] = vbX;
        upk2(vb2, vb_lo, vb_hi);
        va2 = pk2(vb_lo, vb_lo); vbb2 = pk2(vb_hi, vb_hi);
        f32x2 s1a = 0ull, s2a = 0ull, s1b = 0ull, s2b = 0ull;
#pragma unroll
        for (int r = 0; r < R4; r += 4) {
            {
                const f32x2 v0 = HG7_VA(r), v1 = HG7_VA(r + 2);
                const f32x2 x0 = fma2(ga, v0, va2), x1 = fma2(ga, v1, va2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                s1a = add2(s1a, add2(t0, t1));
                s2a = fma2(t0, i0, fma2(t1, i1, s2a));
            }
            {
                const f32x2 v0 = HG7_VB(r), v1 = HG7_VB(r + 2);
                const f32x2 x0 = fma2(gb, v0, vbb2), x1 = fma2(gb, v1, vbb2);
                const f32x2 rr = rcp2(mul2(x0, x1));
                const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                s1b = add2(s1b, add2(t0, t1));
                s2b = fma2(t0, i0, fma2(t1, i1, s2b));
            }
        }
        if (R4 < R) {
            const f32x2 v0 = HG7_VA(R4), v1 = HG7_VB(R4);
            const f32x2 i0 = rcp2(fma2(ga, v0, va2)), i1 = rcp2(fma2(gb, v1, vbb2));
            const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
            s1a = add2(s1a, t0); s2a = fma2(t0, i0, s2a);
            s1b = add2(s1b, t1); s2b = fma2(t1, i1, s2b);
        }
        float s1X = 0.f, s2X = 0.f;
        if (xl) { const float ix = rcp_fast(fmaf(ggX, sX, vbX)); s1X = EX * sX * ix; s2X = s1X * ix; }
        {
            float l0, l1, m0, m1;
            upk2(s1a, l0, l1); upk2(s1b, m0, m1);
            const float s1lo = E0 * (l0 + l1), s1hi = E1 * (m0 + m1);       // Vs = E v
            upk2(s2a, l0, l1); upk2(s2b, m0, m1);
            const float s2lo = E0 * (l0 + l1), s2hi = E1 * (m0 + m1);
            const float v2 = warp_sum(fmaf(pX, s2X, fmaf(p_lo, s2lo, p_hi * s2hi))), v1 = warp_sum(s1lo + s1hi + s1X);
            if (lane == 0) red2[wid] = make_float2(v2, v1);
        }
        __syncthreads();
        float t2 = 0.f, t1s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float2 rr = red2[w]; t2 += rr.x; t1s += rr.y; }
        const float gnew = gg * sqrtf(t2 / t1s);

        // ---- cost with Vx = g_new Vs + Vb2: FOUR samples of a bin share one reciprocal and one log2.  With y_i = c Vx_i
        //      (c = 2^8 keeps the four-fold products inside FP32 for variances from 1e-10 to 1e7; it rides on g_new and Vb2):
        //      sum_i log2 Vx_i = log2(y0 y1 y2 y3) - 32,  sum_i 1 / Vx_i = c ((y0 + y2) y1 y3 + (y1 + y3) y0 y2) / (y0 y1 y2 y3)
        f32x2 gca, gcb, vca, vcb;
        float fix, pca, pcb;                           // -R (log2 c_a + log2 c_b);  c_a P_a,  c_b P_b
        {                                              // per-bin scale (bin_scale) from the bin's first sample
            float lca, lcb;
            const float ca = bin_scale(fmaf(gnew * E0, vst_lo(S[0]), vb_lo), lca), cb = bin_scale(fmaf(gnew * E1, vst_hi(S[0]), vb_hi), lcb);
            gca = pk2(gnew * ca * E0, gnew * ca * E0); gcb = pk2(gnew * cb * E1, gnew * cb * E1);
            vca = pk2(vb_lo * ca, vb_lo * ca); vcb = pk2(vb_hi * cb, vb_hi * cb);
            fix = -(float)R * (lca + lcb);
            pca = p_lo * ca; pcb = p_hi * cb;
        }
        float cla = 0.f, cpa = 0.f, clb = 0.f, cpb = 0.f;
#pragma unroll
        for (int r = 0; r < R4; r += 4) {
            {
                const f32x2 y01 = fma2(gca, HG7_VA(r), vca), y23 = fma2(gca, HG7_VA(r + 2), vca);
                float s_lo, s_hi, q_lo, q_hi;
                upk2(add2(y01, y23), s_lo, s_hi);
                upk2(mul2(y01, y23), q_lo, q_hi);
                const float m = q_lo * q_hi;
                cla += lg2_fast(m);
                cpa = fmaf(fmaf(s_lo, q_hi, s_hi * q_lo), rcp_fast(m), cpa);
            }
            {
                const f32x2 y01 = fma2(gcb, HG7_VB(r), vcb), y23 = fma2(gcb, HG7_VB(r + 2), vcb);
                float s_lo, s_hi, q_lo, q_hi;
                upk2(add2(y01, y23), s_lo, s_hi);
                upk2(mul2(y01, y23), q_lo, q_hi);
                const float m = q_lo * q_hi;
                clb += lg2_fast(m);
                cpb = fmaf(fmaf(s_lo, q_hi, s_hi * q_lo), rcp_fast(m), cpb);
            }
        }
        if (R4 < R) {                                  // one sample pair left: sum 1 / y = (y0 + y1) / (y0 y1)
            float y0, y1;
            upk2(fma2(gca, HG7_VA(R4), vca), y0, y1);
            float m = y0 * y1;
            cla += lg2_fast(m);
            cpa = fmaf(y0 + y1, rcp_fast(m), cpa);
            upk2(fma2(gcb, HG7_VB(R4), vcb), y0, y1);
            m = y0 * y1;
            clb += lg2_fast(m);
            cpb = fmaf(y0 + y1, rcp_fast(m), cpb);
        }
#undef HG7_VA
#undef HG7_VB
        float cX = 0.f;
        if (xl) { const float x0 = fmaf(gnew * EX, sX, vbX); cX = fmaf(0.6931471805599453f, lg2_fast(x0), pX * rcp_fast(x0)); }
        // every sample of a bin carries the bin's log2 c; the reciprocal sums carry a factor 1 / c
        cost_d += (double)(fmaf(0.6931471805599453f, (cla + clb) + fix, fmaf(pca, cpa, pcb * cpb)) + cX);

        if (t < K) H[n * K + t] = hs[t] * norm[u * K + t];
        if (t == 0) g[n] = gnew;
    }
    cost_d = warp_sum_d(cost_d);
    if (lane == 0) redd[wid] = cost_d;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += redd[w];
        cost_part[(int64_t)u * gridDim.x + blockIdx.x] = sum / ((double)R * 513.0 * (double)(n1 - n0));
    }
}

// ----------------------------------------------------------------------------- H, g, cost on the emission, many samples per frame
// hg7 for frames that hold C chains x KEEP kept samples (multi-chain runs, BASELINE configs[3]: 16 x 10 = 160 samples, 169 KB
// per frame): the frame's samples form one flat list (sample j = chain j / KEEP, kept index j % KEEP; the chains of a frame are
// adjacent rows of one emission tile) that is walked in windows of up to 30 samples, once per pass - three reads of the frame, the
// first from HBM, the others from L2.  Same thread layout, packed-lane math and double-buffered cp.async staging as hg7; the
// per-bin partial sums of a pass live in registers across its windows.
template <int KEEP>
__global__ void __launch_bounds__(HG3_THREADS, 3) nmf_hg7w_kernel(const float* __restrict__ P, const uint4* __restrict__ VsT,
                                                                  const uint8_t* __restrict__ vs_idx, const float* __restrict__ bias_log2,
                                                                  const float* __restrict__ Wtmp, const float* __restrict__ norm,
                                                                  float* __restrict__ H, float* __restrict__ g,
                                                                  float* __restrict__ Vb, double* __restrict__ cost_part,
                                                                  const int64_t* __restrict__ fr_off, int K, int C) {
    constexpr int KT = HG2_KT;
    constexpr int ld = HG5_LD;
    constexpr int NBG = 33, TMR = 128, RWMAX = 30;
    static_assert(KEEP % 2 == 0 && KEEP <= 30, "hg7w: an even number of kept samples per chain");
    extern __shared__ __align__(16) uint32_t smw[];
    uint32_t* S0 = smw;                              // [2][RWMAX][HG7_ROWW] the current window and the next one being staged
    float* red = reinterpret_cast<float*>(smw + 2 * RWMAX * HG7_ROWW);  // [8][HG3_NV] per-warp partials of the H sums
    __shared__ float2 red2[8];
    __shared__ double redd[8];
    __shared__ __align__(16) float hs[12];          // KT = 10 activations, read back as two 16-byte and one 8-byte load
    __shared__ __align__(16) float hs_old[12];
    static_assert(KT == 10, "hg7w reads its ten activations as float4 + float4 + float2");
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    const int64_t nb = n0 + (int64_t)blockIdx.x * HG3_FPB;
    if (nb >= n1) {
        if (threadIdx.x == 0) cost_part[(int64_t)u * gridDim.x + blockIdx.x] = 0.0;
        return;
    }
    const int64_t ne = (nb + HG3_FPB < n1) ? nb + HG3_FPB : n1;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int f2 = 2 * t;
    const int Rtot = C * KEEP, NW = (Rtot + RWMAX - 1) / RWMAX;
    f32x2 w2[KT];
    float wX[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const float* wk = Wtmp + ((int64_t)u * K + k) * ld;
        w2[k] = (k < K) ? *reinterpret_cast<const f32x2*>(wk + f2) : 0ull;
        wX[k] = (k < K) ? wk[512] : 0.f;
    }
    const float E0 = exp2f(__ldg(bias_log2 + f2)), E1 = exp2f(__ldg(bias_log2 + f2 + 1)), EX = exp2f(__ldg(bias_log2 + 512));
    double cost_d = 0.0;

    const unsigned s_base = (unsigned)__cvta_generic_to_shared(S0);
    const unsigned po0 = (unsigned)((lane >> 1) * (TMR * 32) + 16 * (lane & 1));
    auto rows_of = [&](int w) { const int left = Rtot - RWMAX * w; return left < RWMAX ? left : RWMAX; };
    auto stage = [&](int64_t n, int w, int b) {       // window w of frame n -> buffer b; a warp copies the rows wid, wid + 8, ...
        const int nr = rows_of(w);
#pragma unroll
        for (int k = 0; k < (RWMAX + 7) / 8; ++k) {
            const int rl = wid + 8 * k;
            if (rl < nr) {
                const int j = RWMAX * w + rl, chain = j / KEEP, r = j - chain * KEEP;
                const int64_t m = n * C + chain;
                const int64_t tile = m / TMR;
                const int row = (int)(m - tile * TMR);
                const unsigned slot = __ldg(vs_idx + m * 32 + r);
                const unsigned char* src = reinterpret_cast<const unsigned char*>(VsT) +
                                           (((size_t)tile * (KEEP + 1) + slot) * NBG * TMR + row) * 32 + po0;
                const unsigned dst = s_base + (unsigned)((b * RWMAX + rl) * HG7_ROWW * 4 + 16 * lane);
                hg7_cp16(dst, src);
                hg7_cp16(dst + 512, src + 16 * (TMR * 32));
                if (lane < 2) hg7_cp16(dst + 1024, src + 32 * (TMR * 32));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(nb, 0, 0);
    if (t >= K && t < KT) { hs_old[t] = 0.f; hs[t] = 0.f; }
    int buf = 0;

    // one pass over the frame's windows: body(S = this thread's word of row 0, SX = word of bin 512 of row 0, rows)
    auto walk = [&](int64_t n, bool last_pass, auto&& body) {
        for (int w = 0; w < NW; ++w) {
            int64_t nn = n;
            int nw = w + 1;
            if (nw == NW) { nw = 0; if (last_pass) nn = n + 1; }
            if (nn < ne) {
                stage(nn, nw, buf ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();                                            // the window is staged for every thread
            body(S0 + buf * RWMAX * HG7_ROWW + t, S0 + buf * RWMAX * HG7_ROWW + 256, rows_of(w));
            __syncthreads();                                            // the window is consumed: its buffer may be refilled
            buf ^= 1;
        }
    };

    for (int64_t n = nb; n < ne; ++n) {
        const float gg = g[n];
        const f32x2 p2 = *reinterpret_cast<const f32x2*>(P + n * ld + f2);
        const float pX = P[n * ld + 512];
        if (t < K) hs_old[t] = H[n * K + t];
        __syncthreads();
        float h[KT];
        {
            const float4 h0 = *reinterpret_cast<const float4*>(hs_old), h1 = *reinterpret_cast<const float4*>(hs_old + 4);
            const float2 h2 = *reinterpret_cast<const float2*>(hs_old + 8);
            h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w; h[8] = h2.x; h[9] = h2.y;
        }
        float p_lo, p_hi;
        upk2(p2, p_lo, p_hi);
        const f32x2 ga = pk2(gg * E0, gg * E0), gb = pk2(gg * E1, gg * E1);
        const float ggX = gg * EX;
#define HG7W_VA(r) pk2(vst_lo(S[(r) * HG7_ROWW]), vst_lo(S[((r) + 1) * HG7_ROWW]))
#define HG7W_VB(r) pk2(vst_hi(S[(r) * HG7_ROWW]), vst_hi(S[((r) + 1) * HG7_ROWW]))
        // ---- H update (Vb1 = W_new H_old)
        f32x2 vb2 = 0ull;
        float vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {                             // bin 512 lives on the threads t < 30 of warp 0 only (warp-uniform branch)
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        float vb_lo, vb_hi;
        upk2(vb2, vb_lo, vb_hi);
        f32x2 va2 = pk2(vb_lo, vb_lo), vbb2 = pk2(vb_hi, vb_hi);
        f32x2 a1a = 0ull, a2a = 0ull, a1b = 0ull, a2b = 0ull;
        float a1X = 0.f, a2X = 0.f;
        walk(n, false, [&](const uint32_t* S, const uint32_t* SX, int nr) {
            int r = 0;
            for (; r + 4 <= nr; r += 4) {
                {
                    const f32x2 x0 = fma2(ga, HG7W_VA(r), va2), x1 = fma2(ga, HG7W_VA(r + 2), va2);
                    const f32x2 rr = rcp2(mul2(x0, x1));
                    const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                    a1a = add2(a1a, add2(i0, i1));
                    a2a = fma2(i0, i0, fma2(i1, i1, a2a));
                }
                {
                    const f32x2 x0 = fma2(gb, HG7W_VB(r), vbb2), x1 = fma2(gb, HG7W_VB(r + 2), vbb2);
                    const f32x2 rr = rcp2(mul2(x0, x1));
                    const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                    a1b = add2(a1b, add2(i0, i1));
                    a2b = fma2(i0, i0, fma2(i1, i1, a2b));
                }
            }
            if (r < nr) {
                const f32x2 ia = rcp2(fma2(ga, HG7W_VA(r), va2)), ibb = rcp2(fma2(gb, HG7W_VB(r), vbb2));
                a1a = add2(a1a, ia); a2a = fma2(ia, ia, a2a);
                a1b = add2(a1b, ibb); a2b = fma2(ibb, ibb, a2b);
            }
            if (t < nr) { const float ix = rcp_fast(fmaf(ggX, vst_lo(SX[t * HG7_ROWW]), vbX)); a1X += ix; a2X = fmaf(ix, ix, a2X); }
        });
        f32x2 a1, a2;
        {
            float l0, l1, m0, m1;
            upk2(a1a, l0, l1); upk2(a1b, m0, m1);
            a1 = pk2(l0 + l1, m0 + m1);
            upk2(a2a, l0, l1); upk2(a2b, m0, m1);
            a2 = pk2(l0 + l1, m0 + m1);
        }
        {
            const f32x2 q2 = mul2(p2, a2);
            const float qX = pX * a2X;
            const bool b4 = lane & 16, b3 = lane & 8;
            float a[10];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float v[10];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = 5 * hf + j;
                    float nlo, nhi, dlo, dhi;
                    upk2(mul2(w2[k], q2), nlo, nhi);
                    upk2(mul2(w2[k], a1), dlo, dhi);
                    v[2 * j] = nlo + nhi;
                    v[2 * j + 1] = dlo + dhi;
                }
                if (wid == 0) {                     // the other warps hold no sample of bin 512 (qX = a1X = 0 there)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        v[2 * j] = fmaf(wX[5 * hf + j], qX, v[2 * j]);
                        v[2 * j + 1] = fmaf(wX[5 * hf + j], a1X, v[2 * j + 1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float send = b4 ? v[j] : v[5 + j], keep = b4 ? v[5 + j] : v[j];
                    a[5 * hf + j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
            }
            float o5[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float send = b3 ? a[j] : a[5 + j], keep = b3 ? a[5 + j] : a[j];
                o5[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1)
#pragma unroll
                for (int j = 0; j < 5; ++j) o5[j] += __shfl_xor_sync(0xffffffffu, o5[j], o);
            if ((lane & 7) == 0) {
                float* dst = red + wid * HG3_NV + 10 * ((lane >> 3) & 1) + 5 * (lane >> 4);
#pragma unroll
                for (int j = 0; j < 5; ++j) dst[j] = o5[j];
            }
        }
        __syncthreads();
        if (t < K) {
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { num += red[w * HG3_NV + 2 * t]; den += red[w * HG3_NV + 2 * t + 1]; }
            hs[t] = hs_old[t] * sqrtf(num / den);
        }
        __syncthreads();
        {
            const float4 h0 = *reinterpret_cast<const float4*>(hs), h1 = *reinterpret_cast<const float4*>(hs + 4);
            const float2 h2 = *reinterpret_cast<const float2*>(hs + 8);
            h[0] = h0.x; h[1] = h0.y; h[2] = h0.z; h[3] = h0.w; h[4] = h1.x; h[5] = h1.y; h[6] = h1.z; h[7] = h1.w; h[8] = h2.x; h[9] = h2.y;
        }

        // ---- g update (Vb2 = W_new H_new, kept as the model's Vb)
        vb2 = 0ull; vbX = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) vb2 = fma2(w2[k], pk2(h[k], h[k]), vb2);
        if (wid == 0) {
#pragma unroll
            for (int k = 0; k < KT; ++k) vbX = fmaf(wX[k], h[k], vbX);
        }
        *reinterpret_cast<f32x2*>(Vb + n * ld + f2) = vb2;
        if (t == 0) Vb[n * ld + 512] = vbX;
        upk2(vb2, vb_lo, vb_hi);
        va2 = pk2(vb_lo, vb_lo); vbb2 = pk2(vb_hi, vb_hi);
        f32x2 s1a = 0ull, s2a = 0ull, s1b = 0ull, s2b = 0ull;
        float s1X = 0.f, s2X = 0.f;
        walk(n, false, [&](const uint32_t* S, const uint32_t* SX, int nr) {
            int r = 0;
            for (; r + 4 <= nr; r += 4) {
                {
                    const f32x2 v0 = HG7W_VA(r), v1 = HG7W_VA(r + 2);
                    const f32x2 x0 = fma2(ga, v0, va2), x1 = fma2(ga, v1, va2);
                    const f32x2 rr = rcp2(mul2(x0, x1));
                    const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                    const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                    s1a = add2(s1a, add2(t0, t1));
                    s2a = fma2(t0, i0, fma2(t1, i1, s2a));
                }
                {
                    const f32x2 v0 = HG7W_VB(r), v1 = HG7W_VB(r + 2);
                    const f32x2 x0 = fma2(gb, v0, vbb2), x1 = fma2(gb, v1, vbb2);
                    const f32x2 rr = rcp2(mul2(x0, x1));
                    const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
                    const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                    s1b = add2(s1b, add2(t0, t1));
                    s2b = fma2(t0, i0, fma2(t1, i1, s2b));
                }
            }
            if (r < nr) {
                const f32x2 v0 = HG7W_VA(r), v1 = HG7W_VB(r);
                const f32x2 i0 = rcp2(fma2(ga, v0, va2)), i1 = rcp2(fma2(gb, v1, vbb2));
                const f32x2 t0 = mul2(v0, i0), t1 = mul2(v1, i1);
                s1a = add2(s1a, t0); s2a = fma2(t0, i0, s2a);
                s1b = add2(s1b, t1); s2b = fma2(t1, i1, s2b);
            }
            if (t < nr) {
                const float sx = vst_lo(SX[t * HG7_ROWW]);
                const float ix = rcp_fast(fmaf(ggX, sx, vbX)), tt = EX * sx * ix;
                s1X += tt;
                s2X = fmaf(tt, ix, s2X);
            }
        });
        {
            float l0, l1, m0, m1;
            upk2(s1a, l0, l1); upk2(s1b, m0, m1);
            const float s1lo = E0 * (l0 + l1), s1hi = E1 * (m0 + m1);
            upk2(s2a, l0, l1); upk2(s2b, m0, m1);
            const float s2lo = E0 * (l0 + l1), s2hi = E1 * (m0 + m1);
            const float v2 = warp_sum(fmaf(pX, s2X, fmaf(p_lo, s2lo, p_hi * s2hi))), v1 = warp_sum(s1lo + s1hi + s1X);
            if (lane == 0) red2[wid] = make_float2(v2, v1);
        }
        __syncthreads();
        float t2 = 0.f, t1s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const float2 rr = red2[w]; t2 += rr.x; t1s += rr.y; }
        const float gnew = gg * sqrtf(t2 / t1s);

        // ---- cost with Vx = g_new Vs + Vb2 (four samples of a bin share one reciprocal and one log2, c = 2^8: see hg5)
        float lca = 0.f, lcb = 0.f, ca = 0.f, cb = 0.f;       // per-bin scale (bin_scale) from the bin's first sample
        f32x2 gca = 0ull, gcb = 0ull, vca = 0ull, vcb = 0ull;
        bool scaled = false;
        float cla = 0.f, cpa = 0.f, clb = 0.f, cpb = 0.f, cX = 0.f;
        walk(n, true, [&](const uint32_t* S, const uint32_t* SX, int nr) {
            if (!scaled) {                              // first window
                ca = bin_scale(fmaf(gnew * E0, vst_lo(S[0]), vb_lo), lca);
                cb = bin_scale(fmaf(gnew * E1, vst_hi(S[0]), vb_hi), lcb);
                gca = pk2(gnew * ca * E0, gnew * ca * E0); gcb = pk2(gnew * cb * E1, gnew * cb * E1);
                vca = pk2(vb_lo * ca, vb_lo * ca); vcb = pk2(vb_hi * cb, vb_hi * cb);
                scaled = true;
            }
            int r = 0;
            for (; r + 4 <= nr; r += 4) {
                {
                    const f32x2 y01 = fma2(gca, HG7W_VA(r), vca), y23 = fma2(gca, HG7W_VA(r + 2), vca);
                    float s_lo, s_hi, q_lo, q_hi;
                    upk2(add2(y01, y23), s_lo, s_hi);
                    upk2(mul2(y01, y23), q_lo, q_hi);
                    const float m = q_lo * q_hi;
                    cla += lg2_fast(m);
                    cpa = fmaf(fmaf(s_lo, q_hi, s_hi * q_lo), rcp_fast(m), cpa);
                }
                {
                    const f32x2 y01 = fma2(gcb, HG7W_VB(r), vcb), y23 = fma2(gcb, HG7W_VB(r + 2), vcb);
                    float s_lo, s_hi, q_lo, q_hi;
                    upk2(add2(y01, y23), s_lo, s_hi);
                    upk2(mul2(y01, y23), q_lo, q_hi);
                    const float m = q_lo * q_hi;
                    clb += lg2_fast(m);
                    cpb = fmaf(fmaf(s_lo, q_hi, s_hi * q_lo), rcp_fast(m), cpb);
                }
            }
            if (r < nr) {
                float y0, y1;
                upk2(fma2(gca, HG7W_VA(r), vca), y0, y1);
                float m = y0 * y1;
                cla += lg2_fast(m);
                cpa = fmaf(y0 + y1, rcp_fast(m), cpa);
                upk2(fma2(gcb, HG7W_VB(r), vcb), y0, y1);
                m = y0 * y1;
                clb += lg2_fast(m);
                cpb = fmaf(y0 + y1, rcp_fast(m), cpb);
            }
            if (t < nr) {
                const float x0 = fmaf(gnew * EX, vst_lo(SX[t * HG7_ROWW]), vbX);
                cX += fmaf(0.6931471805599453f, lg2_fast(x0), pX * rcp_fast(x0));
            }
        });
#undef HG7W_VA
#undef HG7W_VB
        cost_d += (double)(fmaf(0.6931471805599453f, fmaf(-(float)Rtot, lca, cla) + fmaf(-(float)Rtot, lcb, clb),
                                fmaf(p_lo * ca, cpa, p_hi * cb * cpb)) + cX);

        if (t < K) H[n * K + t] = hs[t] * norm[u * K + t];
        if (t == 0) g[n] = gnew;
    }
    cost_d = warp_sum_d(cost_d);
    if (lane == 0) redd[wid] = cost_d;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += redd[w];
        cost_part[(int64_t)u * gridDim.x + blockIdx.x] = sum / ((double)Rtot * 513.0 * (double)(n1 - n0));
    }
}

// cost[u] = sum of the per-CTA partials in block order (deterministic, unlike an atomic accumulation)
__global__ void cost_reduce_kernel(const double* __restrict__ cost_part, int nblk, double* __restrict__ cost, int* __restrict__ status) {
    const int u = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32) s += cost_part[(int64_t)u * nblk + i];
    s = warp_sum_d(s);
    if (threadIdx.x == 0) {
        cost[u] = s;
        if (status && !(fabs(s) <= 1.0e300)) atomicOr(status, DVAE_STATUS_NONFINITE);       // NaN / Inf guard
    }
}

// ----------------------------------------------------------------------------- Wiener masks (mcem.py:325-327)
__global__ void wiener_accum_kernel(const float* __restrict__ Vs, int R, const float* __restrict__ Vb,
                                    const float* __restrict__ g, int64_t NT, int F, int ld, float* __restrict__ WFs,
                                    float* __restrict__ WFn, int first) {
    const int64_t total = NT * ld;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / ld;
        const int f = (int)(i - n * ld);
        if (f >= F) continue;
        const float gg = g[n], vb = Vb[i];
        const float* v = Vs + (n * R) * (int64_t)ld + f;
        float s = first ? 0.f : WFs[i], q = first ? 0.f : WFn[i];
        for (int r = 0; r < R; ++r) {
            const float vs = __fmul_rn(gg, v[(int64_t)r * ld]);
            const float vx = __fadd_rn(vs, vb);
            s += __fdiv_rn(vs, vx);
            q += __fdiv_rn(vb, vx);
        }
        WFs[i] = s;
        WFn[i] = q;
    }
}

__global__ void wiener_apply_kernel(const float2* __restrict__ X, const float* __restrict__ WFs,
                                    const float* __restrict__ WFn, float inv_R, int64_t NT, int F, int ld,
                                    float2* __restrict__ S_hat, float2* __restrict__ N_hat) {
    const int64_t total = NT * ld;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % ld);
        float2 s = make_float2(0.f, 0.f), q = s;
        if (f < F) {
            const float2 x = X[i];
            const float ws = WFs[i] * inv_R, wn = WFn[i] * inv_R;
            s = make_float2(ws * x.x, ws * x.y);
            q = make_float2(wn * x.x, wn * x.y);
        }
        S_hat[i] = s;
        N_hat[i] = q;
    }
}

static int grid_for(int64_t n, int block) {
    const int64_t b = (n + block - 1) / block;
    return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

}  // namespace dvae

using namespace dvae;

extern "C" int dvae_nmf_init(uint64_t seed, const int32_t* utt_ids, const int64_t* fr_off, int B, int64_t NT, int F,
                             int K, int ld, float eps, float* W, float* H, float* g, void* stream) {
    DVAE_REQUIRE(utt_ids && fr_off && W && H && g, "dvae_nmf_init: null pointer");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && F >= 1 && ld >= F && K >= 1 && K <= KMAX, "dvae_nmf_init: bad sizes (K<=%d)", KMAX);
    const int64_t total = (int64_t)B * K * ld + NT * K + NT;
    nmf_init_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32),
                                                                          utt_ids, fr_off, B, NT, F, K, ld, eps, W, H, g);
    return check_launch("nmf_init_kernel");
}

extern "C" int dvae_nmf_vb(const float* W, const float* H, const int32_t* frame_utt, int64_t NT, int F, int K, int ld,
                           float* Vb, void* stream) {
    DVAE_REQUIRE(W && H && frame_utt && Vb, "dvae_nmf_vb: null pointer");
    DVAE_REQUIRE(NT >= 0 && F >= 1 && ld >= F && K >= 1 && K <= KMAX, "dvae_nmf_vb: bad sizes (K<=%d)", KMAX);
    if (NT == 0) return 0;
    nmf_vb_kernel<<<(unsigned)NT, 128, 0, (cudaStream_t)stream>>>(W, H, frame_utt, NT, F, K, ld, Vb);
    return check_launch("nmf_vb_kernel");
}

static int64_t hg_blocks(int max_frames) { return (max_frames > 0 ? (max_frames + FPB - 1) / FPB : 1) + 1; }   // generic path (+ 1 for hg7's aligned blocks): upper bound for all

extern "C" int64_t dvae_nmf_workspace_floats(int B, int K, int ld, int max_frames) {
    if (B <= 0 || K <= 0 || ld <= 0 || max_frames < 0) return 0;
    // un-normalised W_new + column norms (+1 pad to keep the doubles 8-byte aligned) + per-CTA cost partials (doubles)
    int64_t n = (int64_t)B * K * ld + (int64_t)B * K;
    n += n & 1;
    return n + 2 * (int64_t)B * hg_blocks(max_frames);
}

extern "C" int dvae_nmf_mstep(const float* P, const float* Vs, int R, float* W, float* H, float* g, float* Vb,
                              double* cost, const int64_t* fr_off, const int32_t* frame_utt, int B, int64_t NT, int F,
                              int K, int ld, int max_frames, float* ws, const float* wstat, int* status, void* stream) {
    (void)frame_utt;
    DVAE_REQUIRE(P && Vs && W && H && g && Vb && cost && fr_off && ws, "dvae_nmf_mstep: null pointer");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && F >= 1 && ld >= F && R >= 1, "dvae_nmf_mstep: bad sizes");
    DVAE_REQUIRE(K >= 1 && K <= KMAX, "dvae_nmf_mstep: K=%d out of range (<=%d)", K, KMAX);
    DVAE_REQUIRE(max_frames >= 0, "dvae_nmf_mstep: max_frames < 0");
    cudaStream_t st = (cudaStream_t)stream;
    float* Wtmp = ws;
    float* norm = ws + (int64_t)B * K * ld;
    int64_t off = (int64_t)B * K * ld + (int64_t)B * K;
    off += off & 1;
    double* cost_part = reinterpret_cast<double*>(ws + off);
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(cost_part) & 7) == 0, "dvae_nmf_mstep: workspace must be 8-byte aligned");
    const int nblk = (int)hg_blocks(max_frames) - 1;
    int rc;
    if (wstat) {                                   // per-frame reciprocal sums A1 | A2 from dvae_decode_stats_tc
        rc = dvae_nmf_w_from_frame_stats(wstat, wstat + NT * (int64_t)ld, P, H, W, fr_off, B, F, K, ld, Wtmp, stream);
    } else {
        nmf_w_kernel<<<dim3((F + 127) / 128, B), 128, 0, st>>>(P, Vs, R, W, H, g, Vb, fr_off, F, K, ld, Wtmp);
        rc = check_launch("nmf_w_kernel");
    }
    if (rc) return rc;
    nmf_norm_kernel<<<B, 256, 0, st>>>(Wtmp, F, K, ld, W, norm);
    rc = check_launch("nmf_norm_kernel");
    if (rc) return rc;
    if (max_frames == 0) {
        cudaMemsetAsync(cost, 0, sizeof(double) * B, st);
        return 0;
    }
    int nblk_used = nblk;
    const bool aligned = (reinterpret_cast<uintptr_t>(Vs) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wtmp) & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(P) & 7) == 0 && (reinterpret_cast<uintptr_t>(Vb) & 7) == 0;
    if (F == 513 && ld == HG5_LD && K <= HG2_KT && R > 30 && R % 10 == 0 && aligned) {
        // many samples per frame (multi-chain runs): windows of 30 when they divide R, else of 10
        nblk_used = (max_frames + HG3_FPB - 1) / HG3_FPB;
        if (R % 30 == 0) {
            const size_t smem6 = sizeof(float) * ((size_t)30 * ld + (size_t)HG3_NV * 8);
            cudaFuncSetAttribute(nmf_hg6_kernel<30, HG5_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6);
            nmf_hg6_kernel<30, HG5_LD><<<dim3(nblk_used, B), HG3_THREADS, smem6, st>>>(P, Vs, Wtmp, norm, H, g, Vb, cost_part, fr_off, K, R);
        } else {
            const size_t smem6 = sizeof(float) * ((size_t)10 * ld + (size_t)HG3_NV * 8);
            cudaFuncSetAttribute(nmf_hg6_kernel<10, HG5_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem6);
            nmf_hg6_kernel<10, HG5_LD><<<dim3(nblk_used, B), HG3_THREADS, smem6, st>>>(P, Vs, Wtmp, norm, H, g, Vb, cost_part, fr_off, K, R);
        }
        rc = check_launch("nmf_hg6_kernel");
    } else if (F == 513 && ld == HG5_LD && K <= HG2_KT && (R == 10 || R == 30) && aligned) {
        nblk_used = (max_frames + HG3_FPB - 1) / HG3_FPB;
        const size_t smem3 = sizeof(float) * ((size_t)R * ld + (size_t)HG3_NV * 8);
        if (R == 10) {
            cudaFuncSetAttribute(nmf_hg5_kernel<10, HG5_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
            nmf_hg5_kernel<10, HG5_LD><<<dim3(nblk_used, B), HG3_THREADS, smem3, st>>>(P, Vs, Wtmp, norm, H, g, Vb, cost_part, fr_off, K);
        } else {
            cudaFuncSetAttribute(nmf_hg5_kernel<30, HG5_LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
            nmf_hg5_kernel<30, HG5_LD><<<dim3(nblk_used, B), HG3_THREADS, smem3, st>>>(P, Vs, Wtmp, norm, H, g, Vb, cost_part, fr_off, K);
        }
        rc = check_launch("nmf_hg5_kernel");
    } else {
        const size_t smem = sizeof(float) * (size_t)K * ld;
        DVAE_REQUIRE(smem <= 200 * 1024, "dvae_nmf_mstep: K*ld too large for shared memory");
        cudaFuncSetAttribute(nmf_hg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nmf_hg_kernel<<<dim3(nblk, B), 256, smem, st>>>(P, Vs, R, Wtmp, norm, H, g, Vb, cost_part, fr_off, F, K, ld);
        rc = check_launch("nmf_hg_kernel");
    }
    if (rc) return rc;
    cost_reduce_kernel<<<B, 32, 0, st>>>(cost_part, nblk_used, cost, status);
    return check_launch("cost_reduce_kernel");
}

// The same M-step on the sampler's BF16 emission (vst.cu) instead of materialised FP32 variances: F = 513, ld = 520,
// K <= 10, R in {10, 30}; fstat = A1 | A2 from dvae_vst_frame_stats.
extern "C" int dvae_nmf_mstep_vst(const DvaeMlp* dec, const void* image, int L, int y_dim, const float* P, const void* VsT,
                                  const uint8_t* vs_idx, int R, float* W, float* H, float* g, float* Vb, double* cost,
                                  const int64_t* fr_off, int B, int64_t NT, int K, int ld, int max_frames, float* ws,
                                  const float* fstat, const float* wpart, const int32_t* utt_seg, int n_chains, int* status,
                                  void* stream) {
    tc::Dims d;
    int rc = tc::check_dims(dec, L, y_dim, "dvae_nmf_mstep_vst", &d);
    if (rc) return rc;
    DVAE_REQUIRE(image && P && VsT && vs_idx && W && H && g && Vb && cost && fr_off && ws, "dvae_nmf_mstep_vst: null pointer");
    DVAE_REQUIRE((fstat != nullptr) != (wpart != nullptr) && (!wpart || utt_seg),
                 "dvae_nmf_mstep_vst: exactly one of fstat (A1 | A2) and wpart (+ utt_seg) must be given");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && max_frames >= 0, "dvae_nmf_mstep_vst: bad sizes");
    DVAE_REQUIRE(d.F == 513 && ld == HG5_LD && K >= 1 && K <= HG2_KT && (R == 10 || R == 30),
                 "dvae_nmf_mstep_vst: needs F = 513, ld = %d, K <= %d, R in {10, 30}", HG5_LD, HG2_KT);
    DVAE_REQUIRE(n_chains >= 1 && n_chains <= 128 && (n_chains & (n_chains - 1)) == 0,
                 "dvae_nmf_mstep_vst: the chains of a frame must share an emission tile (n_chains a power of two <= 128)");
    DVAE_REQUIRE(n_chains == 1 || wpart, "dvae_nmf_mstep_vst: several chains per frame need the segment partial sums (wpart)");
    const float* bias_log2 = reinterpret_cast<const float*>((const unsigned char*)image + d.off_bias) + (d.n_hidden == 2 ? tc::HID : 0);
    cudaStream_t st = (cudaStream_t)stream;
    float* Wtmp = ws;
    float* norm = ws + (int64_t)B * K * ld;
    int64_t off = (int64_t)B * K * ld + (int64_t)B * K;
    off += off & 1;
    double* cost_part = reinterpret_cast<double*>(ws + off);
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(cost_part) & 7) == 0 && (reinterpret_cast<uintptr_t>(VsT) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(P) & 7) == 0 && (reinterpret_cast<uintptr_t>(Vb) & 7) == 0,
                 "dvae_nmf_mstep_vst: workspace / VsT / P / Vb alignment");
    if (fstat) rc = dvae_nmf_w_from_frame_stats(fstat, fstat + NT * (int64_t)ld, P, H, W, fr_off, B, d.F, K, ld, Wtmp, stream);
    else rc = dvae_nmf_w_from_partials(wpart, utt_seg, W, B, d.F, K, ld, Wtmp, stream);
    if (rc) return rc;
    nmf_norm_kernel<<<B, 256, 0, st>>>(Wtmp, d.F, K, ld, W, norm);
    rc = check_launch("nmf_norm_kernel");
    if (rc) return rc;
    if (max_frames == 0) {
        cudaMemsetAsync(cost, 0, sizeof(double) * B, st);
        return 0;
    }
    const int nblk = (max_frames + HG3_FPB - 1) / HG3_FPB + 1;          // + 1: hg7's blocks are aligned in the global frame index
    const size_t smem7 = 4 * ((size_t)2 * R * HG7_ROWW + (size_t)HG3_NV * 8);
    if (n_chains > 1) {                            // R kept samples per chain, n_chains x R per frame: windowed kernel
        const size_t smem7w = 4 * ((size_t)2 * 30 * HG7_ROWW + (size_t)HG3_NV * 8);
        if (R == 10) {
            cudaFuncSetAttribute(nmf_hg7w_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7w);
            nmf_hg7w_kernel<10><<<dim3(nblk, B), HG3_THREADS, smem7w, st>>>(P, (const uint4*)VsT, vs_idx, bias_log2, Wtmp, norm, H, g, Vb, cost_part, fr_off, K, n_chains);
        } else {
            cudaFuncSetAttribute(nmf_hg7w_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7w);
            nmf_hg7w_kernel<30><<<dim3(nblk, B), HG3_THREADS, smem7w, st>>>(P, (const uint4*)VsT, vs_idx, bias_log2, Wtmp, norm, H, g, Vb, cost_part, fr_off, K, n_chains);
        }
    } else if (R == 10) {
        cudaFuncSetAttribute(nmf_hg7_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7);
        nmf_hg7_kernel<10><<<dim3(nblk, B), HG3_THREADS, smem7, st>>>(P, (const uint4*)VsT, vs_idx, bias_log2, Wtmp, norm, H, g, Vb, cost_part, fr_off, K);
    } else {
        cudaFuncSetAttribute(nmf_hg7_kernel<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7);
        nmf_hg7_kernel<30><<<dim3(nblk, B), HG3_THREADS, smem7, st>>>(P, (const uint4*)VsT, vs_idx, bias_log2, Wtmp, norm, H, g, Vb, cost_part, fr_off, K);
    }
    rc = check_launch("nmf_hg7_kernel");
    if (rc) return rc;
    cost_reduce_kernel<<<B, 32, 0, st>>>(cost_part, nblk, cost, status);
    return check_launch("cost_reduce_kernel");
}

extern "C" int dvae_wiener_accum(const float* Vs, int R, const float* Vb, const float* g, int64_t NT, int F, int ld,
                                 float* WFs, float* WFn, int first, void* stream) {
    DVAE_REQUIRE(Vs && Vb && g && WFs && WFn, "dvae_wiener_accum: null pointer");
    DVAE_REQUIRE(R >= 1 && NT >= 0 && F >= 1 && ld >= F, "dvae_wiener_accum: bad sizes");
    if (NT == 0) return 0;
    wiener_accum_kernel<<<grid_for(NT * ld, 256), 256, 0, (cudaStream_t)stream>>>(Vs, R, Vb, g, NT, F, ld, WFs, WFn, first);
    return check_launch("wiener_accum_kernel");
}

extern "C" int dvae_wiener_apply(const void* X, const float* WFs, const float* WFn, int R_total, int64_t NT, int F,
                                 int ld, void* S_hat, void* N_hat, void* stream) {
    DVAE_REQUIRE(X && WFs && WFn && S_hat && N_hat, "dvae_wiener_apply: null pointer");
    DVAE_REQUIRE(R_total >= 1 && NT >= 0 && F >= 1 && ld >= F, "dvae_wiener_apply: bad sizes");
    if (NT == 0) return 0;
    wiener_apply_kernel<<<grid_for(NT * ld, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)X, WFs, WFn, 1.0f / (float)R_total,
                                                                                NT, F, ld, (float2*)S_hat, (float2*)N_hat);
    return check_launch("wiener_apply_kernel");
}

// P = |X|^2 (mcem.py:47: torch.tensor(np.abs(X)**2)) for spectra that did not come from dvae_stft_f32
namespace dvae {
__global__ void power_kernel(const float2* __restrict__ X, float* __restrict__ P, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 x = X[i];
        const float a = hypotf(x.x, x.y);          // np.abs of complex64, then squared
        P[i] = a * a;
    }
}
}  // namespace dvae

extern "C" int dvae_power(const void* X, float* P, int64_t n, void* stream) {
    DVAE_REQUIRE(X && P && n >= 0, "dvae_power: bad arguments");
    if (n == 0) return 0;
    power_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const float2*)X, P, n);
    return check_launch("power_kernel");
}
