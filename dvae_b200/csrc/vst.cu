// The E-step's kept-sample variances in the sampler's own format ("VsT": BF16, chain tiles) and the kernels that read it.
//
// The tcgen05 sampler (mh_tc2.cu) evaluates the decoder for every proposal; on kept iterations it writes the proposal's
// output 2^v (= Vs without the output-layer bias) straight to global memory, so compute_Vs (mcem.py:280-290) needs no
// second decode.  Layout (one 32-byte cell = 16 consecutive bins of one chain in one slot):
//
//     VsT[tile = chain / 128][slot 0 .. R][bg = bin / 16 (33 groups)][row = chain % 128][8 words]
//     one word = the variances of bins 2j, 2j+1 at BF16 precision (common.cuh, "VsT word": bf16(v_2j) in the low half, the
//     whole word read as FP32 is v_2j+1 to the same precision: one shift unpacks the even bin, nothing the odd one)
//     vs_idx[chain][32] bytes: byte r = slot holding kept sample r  (0 = the state the kept phase started from,
//                                                                     1 + j = the proposal of kept iteration j)
//
// A rejected proposal leaves its slot unreferenced; a sample whose proposal was rejected points at the slot of the last
// accepted one, so nothing is ever copied.  Vs[n][r][f] = E[f] * VsT[...][vs_idx[n][r]][...] with E[f] = exp(b3[f]), the
// output-layer bias the sampler folds into its P / Vb stream instead (pack_pv_kernel).
//
//   vst_frame_stats_kernel   A1[n][f] = sum_r 1 / Vx, A2[n][f] = sum_r 1 / Vx^2 (inner sums of mcem.py:108-110) with the
//                            E-step's g and Vb: thread = (chain, 8 bins), a warp reads 512 contiguous bytes per slot.
//   vst_unpack / vst_pack    conversion to / from the dense FP32 layout Vs[NT][R][ld] (the `.Vs` attribute of the MCEM
//                            shims, parity tests).
// The H / g / cost kernel on this format is nmf_hg7_kernel (nmf.cu).
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int NBG = NPAD / 16;                  // 33 bin groups
constexpr int VST_IDX_PITCH = 32;

__device__ __forceinline__ f32x2 vst_rcp2(f32x2 a) { float lo, hi; upk2(a, lo, hi); return pk2(rcp_approx(lo), rcp_approx(hi)); }

// Thread = (chain, half of a cell): 8 bins = one 16-byte load per slot; the two threads of a chain are adjacent lanes, so a warp
// reads 16 chains x 32 bytes = 512 contiguous bytes per slot.  64 registers -> 1 024 threads per SM, each with two sample pairs
// (64 bytes) in flight ahead of the one it reduces: the kernel is a pure HBM stream (R x 1 056 bytes per frame in, 2 x 2 052 out).
// RT > 0: compile-time sample count (even), slot indices held in registers; RT = 0: run-time count r_rt (any parity)
// WPART: instead of writing A1 / A2, the CTA multiplies them with the frames' activations and reduces over the frames of each
// utterance segment inside its tile: Wpart[seg][k][0][f] = sum_n P A2 H[n][k], Wpart[seg][k][1][f] = sum_n A1 H[n][k] (the sums of
// the W update, mcem.py:108-110).  A segment is a maximal run of frames that lies in one tile AND one utterance (host table);
// every (segment, bin) has exactly one writer and the final sum over an utterance's <= 3-4 segments runs in a fixed order, so
// the result is deterministic.  Saves the A1 / A2 round trip (2 x 4 x ld x NT bytes out, 3 x in) and the separate W kernel.
// With 2^cshift chains per frame (rows = chains, frame = row >> cshift) the same sums run over the chain rows: the frame's
// quantities are looked up per row and the segment table is cut on the row axis.
struct WPartArgs {
    const float* P;            // [NT][ld]
    const float* H;            // [NT][K]
    int K;
    const int64_t* seg_start;  // [S + 1] frame boundaries of the segments (ascending)
    const int32_t* tile_seg;   // [n_tiles + 1] first segment of every tile
    float* Wpart;              // [S][K][2][ld]
};

template <int RT, bool WPART>
__global__ void __launch_bounds__(256, 4) vst_frame_stats_kernel(const uint4* __restrict__ VsT, const uint8_t* __restrict__ idx, int r_rt,
                                                                 const float* __restrict__ bias_log2, const float* __restrict__ Vb,
                                                                 const float* __restrict__ g, int64_t rows, int cshift, int F, int ld,
                                                                 float* __restrict__ A1, float* __restrict__ A2, WPartArgs wp) {
    __shared__ float sA1[WPART ? TM : 1][17], sPA2[WPART ? TM : 1][17], sH[WPART ? TM : 1][12];
    const int R = RT > 0 ? RT : r_rt;
    const int64_t tile = blockIdx.x;
    const int bg = blockIdx.y, row = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int64_t m = tile * TM + row;
    // Order of the requests matters more than anything else here: a CTA lives for a few microseconds, and the first version
    // waited for three or four DRAM round trips one after the other (H -> shared memory, then g / Vb / bias, then the slot
    // indices, then the cells, then P).  Now: the activations go to shared memory asynchronously (cp.async, nobody waits before
    // the tail), the slot indices are requested first, the first cells right behind them, and only then the frame's scalars.
    if (WPART) {
        for (int i = threadIdx.x; i < TM * 12; i += 256) {
            const int r = i / 12, k = i - 12 * r;
            const int64_t n = tile * TM + r;
            if (n < rows && k < wp.K) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&sH[r][k]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(wp.H + (n >> cshift) * wp.K + k) : "memory");
            } else {
                sH[r][k] = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (m >= rows) {
        if (!WPART) return;
#pragma unroll
        for (int j = 0; j < 8; ++j) { sA1[row][8 * half + j] = 0.f; sPA2[row][8 * half + j] = 0.f; }
    }
    const bool rowlive = m < rows;
    const int64_t mm = rowlive ? m : 0;
    const int f0 = 16 * bg + 8 * half;
    const size_t slot_stride = (size_t)NBG * TM * 2;
    const uint4* cell = VsT + ((size_t)tile * (R + 1) * NBG + bg) * (TM * 2) + row * 2 + half;
    const uint8_t* ib = idx + mm * VST_IDX_PITCH;
    uint32_t iw[8];
    if (RT > 0) {
        const uint4 i0 = __ldg(reinterpret_cast<const uint4*>(ib)), i1 = __ldg(reinterpret_cast<const uint4*>(ib) + 1);
        iw[0] = i0.x; iw[1] = i0.y; iw[2] = i0.z; iw[3] = i0.w; iw[4] = i1.x; iw[5] = i1.y; iw[6] = i1.z; iw[7] = i1.w;
    }
    auto slot_of = [&](int r) -> uint32_t {
        if (RT > 0) return (iw[r >> 2] >> (8 * (r & 3))) & 255u;
        return (uint32_t)__ldg(ib + r);
    };
    // software pipeline over sample pairs (RT > 0): pairs k+1 and k+2 are in flight while pair k is reduced
    constexpr int NP = RT / 2;
    uint4 c[2], n1[2], n2[2];
    if (RT > 0) {
        c[0] = __ldcs(cell + slot_of(0) * slot_stride); c[1] = __ldcs(cell + slot_of(1) * slot_stride);
        if (NP > 1) { n1[0] = __ldcs(cell + slot_of(2) * slot_stride); n1[1] = __ldcs(cell + slot_of(3) * slot_stride); }
    }
    f32x2 ge2[4], vb2[4], a1[4], a2[4];
    {
        const int64_t fr = mm >> cshift;                   // the row's frame: 2^cshift chains per frame (WPART only), else row = frame
        const float gg = __ldg(g + fr);
        float vb[8];
        if (f0 + 8 <= ld) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(Vb + fr * ld + f0) + j);
                vb[4 * j] = t.x; vb[4 * j + 1] = t.y; vb[4 * j + 2] = t.z; vb[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) vb[j] = (f0 + j < F) ? __ldg(Vb + fr * ld + f0 + j) : 1.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = f0 + 2 * j;
            const float e0 = (f < F) ? exp2f(__ldg(bias_log2 + f)) : 0.f, e1 = (f + 1 < F) ? exp2f(__ldg(bias_log2 + f + 1)) : 0.f;
            ge2[j] = pk2(gg * e0, gg * e1);
            vb2[j] = pk2((f < F) ? vb[2 * j] : 1.f, (f + 1 < F) ? vb[2 * j + 1] : 1.f);      // padding bins: Vx = 1, never stored
            a1[j] = 0ull;
            a2[j] = 0ull;
        }
    }
    auto word_pair = [&](uint32_t w0, uint32_t w1, int j) {
        // two samples share one reciprocal: 1 / X0 = X1 / (X0 X1)
        const f32x2 x0 = fma2(ge2[j], pk2(vst_lo(w0), vst_hi(w0)), vb2[j]), x1 = fma2(ge2[j], pk2(vst_lo(w1), vst_hi(w1)), vb2[j]);
        const f32x2 rr = vst_rcp2(mul2(x0, x1));
        const f32x2 i0 = mul2(x1, rr), i1 = mul2(x0, rr);
        a1[j] = add2(a1[j], add2(i0, i1));
        a2[j] = fma2(i0, i0, fma2(i1, i1, a2[j]));
    };
    auto pair = [&](const uint4& p, const uint4& q) {
        word_pair(p.x, q.x, 0); word_pair(p.y, q.y, 1); word_pair(p.z, q.z, 2); word_pair(p.w, q.w, 3);
    };
    float pp[8];                                           // WPART: the frame's observation for the tail, requested before the last pairs are reduced
    auto load_p = [&]() {
        if (WPART && rowlive) {
            const int64_t fr = m >> cshift;
            if (f0 + 8 <= ld) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(wp.P + fr * ld + f0) + j);
                    pp[4 * j] = t.x; pp[4 * j + 1] = t.y; pp[4 * j + 2] = t.z; pp[4 * j + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) pp[j] = (f0 + j < F) ? __ldg(wp.P + fr * ld + f0 + j) : 0.f;
            }
        }
    };
    if (RT > 0) {
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            if (k + 2 < NP) { n2[0] = __ldcs(cell + slot_of(2 * k + 4) * slot_stride); n2[1] = __ldcs(cell + slot_of(2 * k + 5) * slot_stride); }
            if (k == (NP > 2 ? NP - 2 : 0)) load_p();      // the last cells are on their way: n2's registers are free from here on
            pair(c[0], c[1]);
            c[0] = n1[0]; c[1] = n1[1];
            n1[0] = n2[0]; n1[1] = n2[1];
        }
    } else {
        load_p();
        const int R2 = R & ~1;
        for (int r = 0; r < R2; r += 2) pair(__ldcs(cell + slot_of(r) * slot_stride), __ldcs(cell + slot_of(r + 1) * slot_stride));
        if (R & 1) {
            const uint4 p = __ldcs(cell + slot_of(R - 1) * slot_stride);
            const uint32_t w0[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const f32x2 i0 = vst_rcp2(fma2(ge2[j], pk2(vst_lo(w0[j]), vst_hi(w0[j])), vb2[j]));
                a1[j] = add2(a1[j], i0);
                a2[j] = fma2(i0, i0, a2[j]);
            }
        }
    }
    float o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { upk2(a1[j], o1[2 * j], o1[2 * j + 1]); upk2(a2[j], o2[2 * j], o2[2 * j + 1]); }
    if (WPART) {
        if (rowlive) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = f0 + j < F;
                sA1[row][8 * half + j] = ok ? o1[j] : 0.f;
                sPA2[row][8 * half + j] = ok ? pp[j] * o2[j] : 0.f;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int j = threadIdx.x & 15, k = threadIdx.x >> 4;               // threads 0 .. 16 K - 1: (bin of the group, rank)
        const int f = 16 * bg + j;
        if (k < wp.K && f < F) {
            const int64_t tile0 = tile * TM;
            for (int sg = wp.tile_seg[tile]; sg < wp.tile_seg[tile + 1]; ++sg) {
                const int lo = (int)(wp.seg_start[sg] - tile0), hi = (int)(wp.seg_start[sg + 1] - tile0);
                float num = 0.f, den = 0.f, num1 = 0.f, den1 = 0.f;      // two accumulator pairs: the loads of 8 frames in flight
                int n = lo;
                for (; n + 8 <= hi; n += 8) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const float h0 = sH[n + i][k], h1 = sH[n + i + 1][k];
                        num = fmaf(sPA2[n + i][j], h0, num);
                        den = fmaf(sA1[n + i][j], h0, den);
                        num1 = fmaf(sPA2[n + i + 1][j], h1, num1);
                        den1 = fmaf(sA1[n + i + 1][j], h1, den1);
                    }
                }
                for (; n < hi; ++n) {
                    const float hk = sH[n][k];
                    num = fmaf(sPA2[n][j], hk, num);
                    den = fmaf(sA1[n][j], hk, den);
                }
                num += num1;
                den += den1;
                float* dst = wp.Wpart + ((size_t)sg * wp.K + k) * 2 * ld + f;
                dst[0] = num;
                dst[ld] = den;
            }
        }
        return;
    }
    if (f0 + 8 <= F) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            reinterpret_cast<float4*>(A1 + m * ld + f0)[j] = make_float4(o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
            reinterpret_cast<float4*>(A2 + m * ld + f0)[j] = make_float4(o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (f0 + j < F) { A1[m * ld + f0 + j] = o1[j]; A2[m * ld + f0 + j] = o2[j]; }
    }
}

// W <- W sqrt(num / den) from the per-segment partial sums (mcem.py:108-111); one thread per bin, an utterance's segments in order
__global__ void __launch_bounds__(128) w_from_partials_kernel(const float* __restrict__ Wpart, const int32_t* __restrict__ utt_seg,
                                                              const float* __restrict__ W, int F, int K, int ld, float* __restrict__ Wtmp) {
    const int u = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int s0 = utt_seg[u], s1 = utt_seg[u + 1];
    for (int k = 0; k < K; ++k) {
        float num = 0.f, den = 0.f;
        for (int sg = s0; sg < s1; ++sg) {
            const float* src = Wpart + ((size_t)sg * K + k) * 2 * ld + f;
            num += src[0];
            den += src[ld];
        }
        const int64_t i = ((int64_t)u * K + k) * ld + f;
        Wtmp[i] = W[i] * sqrtf(num / den);
    }
}

// Vs[n][r][f] = E[f] * VsT[..][vs_idx[n][r]][..]   (one thread per (frame, sample, bin pair))
__global__ void vst_unpack_kernel(const uint32_t* __restrict__ VsT, const uint8_t* __restrict__ idx, int R, const float* __restrict__ bias_log2,
                                  int64_t rows, int F, int ld, float* __restrict__ Vs) {
    const int64_t total = rows * R * (NPAD / 2);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int pw = (int)(i % (NPAD / 2));
        const int64_t mr = i / (NPAD / 2);
        const int r = (int)(mr % R);
        const int64_t m = mr / R;
        const int f = 2 * pw;
        if (f >= F) continue;
        const int64_t tile = m / TM;
        const int row = (int)(m % TM), slot = idx[m * VST_IDX_PITCH + r], bg = f >> 4;
        const uint32_t w = VsT[((((size_t)tile * (R + 1) + slot) * NBG + bg) * TM + row) * 8 + ((f & 15) >> 1)];
        float* dst = Vs + (m * R + r) * (int64_t)ld + f;
        dst[0] = exp2f(bias_log2[f]) * vst_lo(w);
        if (f + 1 < F) dst[1] = exp2f(bias_log2[f + 1]) * vst_hi(w);
    }
}

// the inverse: sample r of frame n goes to slot 1 + r (slot 0 and the padding bins are zero), vs_idx[n][r] = 1 + r
__global__ void vst_pack_kernel(const float* __restrict__ Vs, int R, const float* __restrict__ bias_log2, int64_t rows, int F, int ld,
                                uint32_t* __restrict__ VsT, uint8_t* __restrict__ idx) {
    const int64_t n_tiles = (rows + TM - 1) / TM;
    const int64_t total = n_tiles * (R + 1) * NBG * TM * 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int wd = (int)(i & 7);
        const int row = (int)((i >> 3) % TM);
        const int bg = (int)((i / (8 * TM)) % NBG);
        const int slot = (int)((i / (8 * TM * NBG)) % (R + 1));
        const int64_t tile = i / ((int64_t)8 * TM * NBG * (R + 1));
        const int64_t m = tile * TM + row;
        const int f = 16 * bg + 2 * wd;
        uint32_t w = 0u;
        if (m < rows && slot > 0 && f < F) {
            const float* src = Vs + (m * R + slot - 1) * (int64_t)ld + f;
            const float lo = src[0] * exp2f(-bias_log2[f]);
            const float hi = (f + 1 < F) ? src[1] * exp2f(-bias_log2[f + 1]) : 1.f;
            w = vst_word(lo, hi);
        }
        VsT[i] = w;
        if (bg == 0 && wd == 0 && m < rows && slot > 0) idx[m * VST_IDX_PITCH + slot - 1] = (uint8_t)slot;
    }
}

}  // namespace tc
}  // namespace dvae

using namespace dvae;
using namespace dvae::tc;

static const float* vst_bias(const DvaeMlp* dec, const void* image, int L, int y_dim, const char* who, Dims* d, int* rc) {
    *rc = check_dims(dec, L, y_dim, who, d);
    if (*rc) return nullptr;
    // layer-3 bias (log2 domain) inside the decoder image: after the hidden-2 bias if there is one
    return reinterpret_cast<const float*>((const unsigned char*)image + d->off_bias) + (d->n_hidden == 2 ? HID : 0);
}

extern "C" int dvae_vst_frame_stats(const DvaeMlp* dec, const void* image, int L, int y_dim, const void* VsT, const uint8_t* vs_idx,
                                    int R, const float* Vb, const float* g, int64_t NT, int ld, float* A1, float* A2, void* stream) {
    Dims d;
    int rc;
    DVAE_REQUIRE(image != nullptr, "dvae_vst_frame_stats: null pointer");
    const float* bias_log2 = vst_bias(dec, image, L, y_dim, "dvae_vst_frame_stats", &d, &rc);
    if (rc) return rc;
    DVAE_REQUIRE(VsT && vs_idx && Vb && g && A1 && A2, "dvae_vst_frame_stats: null pointer");
    DVAE_REQUIRE(R >= 1 && R <= 31 && NT >= 0 && ld >= d.F && (ld & 3) == 0, "dvae_vst_frame_stats: bad sizes (1 <= R <= 31, ld %% 4 == 0)");
    DVAE_REQUIRE(((reinterpret_cast<uintptr_t>(VsT) | reinterpret_cast<uintptr_t>(vs_idx) | reinterpret_cast<uintptr_t>(Vb) |
                   reinterpret_cast<uintptr_t>(A1) | reinterpret_cast<uintptr_t>(A2)) & 15) == 0, "dvae_vst_frame_stats: 16-byte alignment required");
    if (NT == 0) return 0;
    const int64_t n_tiles = (NT + TM - 1) / TM;
    DVAE_REQUIRE(n_tiles < (1ll << 31), "dvae_vst_frame_stats: too many frames");
    const dim3 grid((unsigned)n_tiles, NBG);
    cudaStream_t st = (cudaStream_t)stream;
    const WPartArgs none{};
    if (R == 30) vst_frame_stats_kernel<30, false><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, NT, 0, d.F, ld, A1, A2, none);
    else if (R == 10) vst_frame_stats_kernel<10, false><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, NT, 0, d.F, ld, A1, A2, none);
    else vst_frame_stats_kernel<0, false><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, NT, 0, d.F, ld, A1, A2, none);
    return check_launch("vst_frame_stats_kernel");
}

extern "C" int64_t dvae_vst_w_partial_floats(int64_t n_segments, int K, int ld) {
    if (n_segments <= 0 || K <= 0 || ld <= 0) return 0;
    return n_segments * K * 2 * (int64_t)ld;
}

extern "C" int dvae_vst_w_partials(const DvaeMlp* dec, const void* image, int L, int y_dim, const void* VsT, const uint8_t* vs_idx,
                                   int R, const float* P, const float* Vb, const float* g, const float* H, int K, int64_t NT, int n_chains,
                                   int ld, const int64_t* seg_start, const int32_t* tile_seg, float* Wpart, void* stream) {
    Dims d;
    int rc;
    DVAE_REQUIRE(image != nullptr, "dvae_vst_w_partials: null pointer");
    const float* bias_log2 = vst_bias(dec, image, L, y_dim, "dvae_vst_w_partials", &d, &rc);
    if (rc) return rc;
    DVAE_REQUIRE(VsT && vs_idx && P && Vb && g && H && seg_start && tile_seg && Wpart, "dvae_vst_w_partials: null pointer");
    DVAE_REQUIRE(R >= 1 && R <= 31 && K >= 1 && K <= 12 && NT >= 0 && ld >= d.F && (ld & 3) == 0,
                 "dvae_vst_w_partials: bad sizes (1 <= R <= 31, K <= 12, ld %% 4 == 0)");
    DVAE_REQUIRE(((reinterpret_cast<uintptr_t>(VsT) | reinterpret_cast<uintptr_t>(vs_idx) | reinterpret_cast<uintptr_t>(Vb) |
                   reinterpret_cast<uintptr_t>(P)) & 15) == 0, "dvae_vst_w_partials: 16-byte alignment required");
    DVAE_REQUIRE(n_chains >= 1 && n_chains <= TM && (n_chains & (n_chains - 1)) == 0,
                 "dvae_vst_w_partials: n_chains must be a power of two <= 128 (the chains of a frame share an emission tile)");
    if (NT == 0) return 0;
    int cshift = 0;
    while ((1 << cshift) < n_chains) ++cshift;
    const int64_t rows = NT * n_chains;
    const int64_t n_tiles = (rows + TM - 1) / TM;
    DVAE_REQUIRE(n_tiles < (1ll << 31), "dvae_vst_w_partials: too many frames");
    const dim3 grid((unsigned)n_tiles, NBG);
    cudaStream_t st = (cudaStream_t)stream;
    const WPartArgs wp{P, H, K, seg_start, tile_seg, Wpart};
    if (R == 30) vst_frame_stats_kernel<30, true><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, rows, cshift, d.F, ld, nullptr, nullptr, wp);
    else if (R == 10) vst_frame_stats_kernel<10, true><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, rows, cshift, d.F, ld, nullptr, nullptr, wp);
    else vst_frame_stats_kernel<0, true><<<grid, 2 * TM, 0, st>>>((const uint4*)VsT, vs_idx, R, bias_log2, Vb, g, rows, cshift, d.F, ld, nullptr, nullptr, wp);
    return check_launch("vst_frame_stats_kernel<wpart>");
}

extern "C" int dvae_nmf_w_from_partials(const float* Wpart, const int32_t* utt_seg, const float* W, int B, int F, int K, int ld,
                                        float* Wtmp, void* stream) {
    DVAE_REQUIRE(Wpart && utt_seg && W && Wtmp && B >= 1 && K >= 1 && ld >= F, "dvae_nmf_w_from_partials: bad arguments");
    w_from_partials_kernel<<<dim3((F + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(Wpart, utt_seg, W, F, K, ld, Wtmp);
    return check_launch("w_from_partials_kernel");
}

extern "C" int dvae_vst_unpack(const DvaeMlp* dec, const void* image, int L, int y_dim, const void* VsT, const uint8_t* vs_idx, int R,
                               int64_t NT, int ld, float* Vs, void* stream) {
    Dims d;
    int rc;
    DVAE_REQUIRE(image != nullptr, "dvae_vst_unpack: null pointer");
    const float* bias_log2 = vst_bias(dec, image, L, y_dim, "dvae_vst_unpack", &d, &rc);
    if (rc) return rc;
    DVAE_REQUIRE(VsT && vs_idx && Vs && R >= 1 && R <= 31 && NT >= 0 && ld >= d.F, "dvae_vst_unpack: bad arguments");
    if (NT == 0) return 0;
    vst_unpack_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)VsT, vs_idx, R, bias_log2, NT, d.F, ld, Vs);
    return check_launch("vst_unpack_kernel");
}

extern "C" int dvae_vst_pack(const DvaeMlp* dec, const void* image, int L, int y_dim, const float* Vs, int R, int64_t NT, int ld,
                             void* VsT, uint8_t* vs_idx, void* stream) {
    Dims d;
    int rc;
    DVAE_REQUIRE(image != nullptr, "dvae_vst_pack: null pointer");
    const float* bias_log2 = vst_bias(dec, image, L, y_dim, "dvae_vst_pack", &d, &rc);
    if (rc) return rc;
    DVAE_REQUIRE(VsT && vs_idx && Vs && R >= 1 && R <= 31 && NT >= 0 && ld >= d.F, "dvae_vst_pack: bad arguments");
    if (NT == 0) return 0;
    vst_pack_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(Vs, R, bias_log2, NT, d.F, ld, (uint32_t*)VsT, vs_idx);
    return check_launch("vst_pack_kernel");
}
