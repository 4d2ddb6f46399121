// Shared pieces of the tcgen05 kernels (sampler, decode, decode + frame statistics): operand layout, PTX wrappers,
// the hidden-layer epilogue and the MMA issue loop.  See tc_decode.cu / mh_tc2.cu for the design notes.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dvae {
namespace tc {

constexpr int TM = 128;             // rows per tile
constexpr int HID = 128;            // hidden width (both hidden layers)
constexpr int NPAD = 528;           // output bins padded: 4*128 + 16
constexpr int NQ = NPAD / 4;        // bin quads in the packed P / Vb layout
constexpr int NTHREADS = 288;
constexpr int A_BYTES = 32768;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Dims {
    int L, y_dim, n_hidden, F, nkb1;
    int off_w2, off_w3, off_bias, image_bytes;      // byte offsets inside the image
};

__host__ __device__ inline Dims make_dims(int L, int y_dim, int n_hidden, int F) {
    Dims d;
    d.L = L; d.y_dim = y_dim; d.n_hidden = n_hidden; d.F = F;
    const int k1 = 2 * L + 2 * y_dim + 1;
    d.nkb1 = (k1 + 63) / 64;
    d.off_w2 = d.nkb1 * 16384;
    d.off_w3 = d.off_w2 + (n_hidden == 2 ? 2 * HID * 128 : 0);
    d.off_bias = d.off_w3 + 2 * NPAD * 128;
    d.image_bytes = d.off_bias + 4 * ((n_hidden == 2 ? HID : 0) + NPAD);
    return d;
}

// byte offset of element (row n, column k) of a K-major SWIZZLE_128B operand with `rows` rows
__host__ __device__ inline int sw128_offset(int rows, int n, int k) {
    const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
    return kb * rows * 128 + n * 128 + ((c ^ (n & 7)) << 4) + e * 2;
}

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B: start>>4 | LBO=1 (ignored) | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t umma_idesc(int N) {
    // kind::f16: D=F32 (bit 4), A=BF16 (bits 7-9 = 1), B=BF16 (bits 10-12 = 1), K-major A and B, N>>3 at 17, M>>4 at 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

// Bounded wait: a broken pipeline must never hang the GPU.  The bound is wall time (2 s on the global timer, polled every
// 4096 failed probes), not a probe count, so time slicing or an attached profiler cannot produce a false timeout.  On
// timeout the CTA-wide `dead` flag makes every later wait fall through and the host sees DVAE_STATUS_TIMEOUT.
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* dead, int* status) {
    if (*dead) return;
    unsigned long long t0 = 0;
    for (uint32_t spin = 1;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 4095u) == 0) {
            if (*dead) return;
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) break;
        }
    }
    *dead = 1;
    atomicOr(status, DVAE_STATUS_TIMEOUT);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // first source -> upper half
    return r;
}


// D[tmem] (+)= A[128 x 64*nkb] * B[N x 64*nkb]^T as 4*nkb tcgen05.mma of K=16; both operands K-major SWIZZLE_128B,
// K block kb of A / B starts a_kb_stride / b_kb_stride bytes after block kb-1.
// acc0 = 1 accumulates onto what the accumulator already holds (a bias preloaded with tcgen05.st).
__device__ __forceinline__ void issue_gemm2(uint32_t a_addr, int a_kb_stride, uint32_t b_addr, int b_kb_stride, int nkb,
                                            uint32_t d_tmem, int N, uint32_t acc0 = 0) {
    const uint32_t idesc = umma_idesc(N);
    uint32_t acc = acc0;
    for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            umma(d_tmem, umma_desc(a_addr + kb * a_kb_stride + 32 * k), umma_desc(b_addr + kb * b_kb_stride + 32 * k), idesc, acc);
            acc = 1;
        }
    }
}

// one row of the layer-1 operand: [bf16 hi(z) | bf16 lo(z) | y hi,lo ... | 1 | 0 ...] as 8-element swizzled chunks
__device__ __forceinline__ void write_a1_row(const Dims& d, unsigned char* A, int row, const float* zin, const float* yrow, bool valid) {
    const int K1 = 64 * d.nkb1;
    for (int ch = 0; ch < K1 / 8; ++ch) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = 8 * ch + i;
            float x = 0.f;
            if (valid) {
                if (k < d.L) x = __bfloat162float(__float2bfloat16_rn(zin[k]));
                else if (k < 2 * d.L) { const float zz = zin[k - d.L]; x = zz - __bfloat162float(__float2bfloat16_rn(zz)); }
                else if (k < 2 * d.L + 2 * d.y_dim) {
                    const float yy = yrow[(k - 2 * d.L) >> 1];
                    const float hi = __bfloat162float(__float2bfloat16_rn(yy));
                    x = ((k - 2 * d.L) & 1) ? (yy - hi) : hi;
                } else if (k == 2 * d.L + 2 * d.y_dim) x = 1.f;
            }
            e[i] = x;
        }
        const int kb = ch >> 3, cc = ch & 7;
        uint4 pk = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ (row & 7)) << 4)) = pk;
    }
}

// hidden-layer epilogue of the thread owning TMEM lane `row` (quadrant q) and column half h:
// D12[row][64h .. 64h+64) -> tanh(+bias) -> bf16 -> K-block h of the activation operand
__device__ __forceinline__ void hidden_epilogue_rows(uint32_t tmem, unsigned char* A, int q, int h, int row, const float* bias) {
    float v[32];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        const int col0 = 64 * h + 32 * part;
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + col0, v);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            float t[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float x = v[8 * cc + e];
                if (bias) x += bias[col0 + 8 * cc + e];
                t[e] = tanh_approx(x);
            }
            const int chunk = 4 * part + cc;
            uint4 pk = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
            *reinterpret_cast<uint4*>(A + h * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = pk;
        }
    }
}

__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// layer-1 operand row with a static layout: [hi(z) (L) | lo(z) (L) | y hi,lo ... | 1 | 0 ...]
template <int L>
__device__ __forceinline__ void write_a1_static(int y_dim, int nkb1, unsigned char* A, int row, const float (&z)[L], float y0, float y1,
                                                float y2, bool valid) {
    constexpr int CH = L / 8;                  // chunks of hi (and of lo)
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 2 * CH; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float zz = valid ? z[(c % CH) * 8 + i] : 0.f;
            const float hi = bf16_hi(zz);
            e[i] = (c < CH) ? hi : (zz - hi);
        }
        const int kb = c >> 3, cc = c & 7;
        uint4 pk = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = pk;
    }
    {   // chunk 2*CH: labels (hi, lo pairs, y_dim <= 3) and the constant one
        const float h0 = bf16_hi(y0), h1 = bf16_hi(y1), h2 = bf16_hi(y2);
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, e4 = 0.f, e5 = 0.f, e6 = 0.f;
        if (valid) {
            if (y_dim == 0) { e0 = 1.f; }
            else if (y_dim == 1) { e0 = h0; e1 = y0 - h0; e2 = 1.f; }
            else if (y_dim == 2) { e0 = h0; e1 = y0 - h0; e2 = h1; e3 = y1 - h1; e4 = 1.f; }
            else { e0 = h0; e1 = y0 - h0; e2 = h1; e3 = y1 - h1; e4 = h2; e5 = y2 - h2; e6 = 1.f; }
        }
        constexpr int c = 2 * CH;
        const int kb = c >> 3, cc = c & 7;
        uint4 pk = make_uint4(pack_bf16x2(e0, e1), pack_bf16x2(e2, e3), pack_bf16x2(e4, e5), pack_bf16x2(e6, 0.f));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = pk;
    }
    const int K1c = 8 * nkb1;                  // remaining chunks of the K blocks in use are zero
    for (int c = 2 * CH + 1; c < K1c; ++c) {
        const int kb = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// The same operand row written by two threads: half h holds the latent dimensions [h L/2, (h+1) L/2) in zh and writes their
// hi and lo chunks; half 0 adds the label / constant chunk, half 1 the zero chunks.
template <int L>
__device__ __forceinline__ void write_a1_half(int y_dim, int nkb1, unsigned char* A, int row, int h, const float (&zh)[L / 2], float y0,
                                              float y1, float y2, bool valid) {
    constexpr int CH = L / 8, CHH = CH / 2;    // chunks of hi (and of lo) per row / per half
    static_assert(L % 16 == 0, "write_a1_half: whole 8-element chunks per half");
    const int sw = row & 7;
#pragma unroll
    for (int j = 0; j < 2 * CHH; ++j) {
        const bool lo = j >= CHH;
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float zz = valid ? zh[(j % CHH) * 8 + i] : 0.f;
            const float hi = bf16_hi(zz);
            e[i] = lo ? (zz - hi) : hi;
        }
        const int c = (lo ? CH : 0) + h * CHH + (j % CHH);
        const int kb = c >> 3, cc = c & 7;
        uint4 pk = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = pk;
    }
    if (h == 0) {   // chunk 2*CH: labels (hi, lo pairs, y_dim <= 3) and the constant one
        const float h0 = bf16_hi(y0), h1 = bf16_hi(y1), h2 = bf16_hi(y2);
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f, e4 = 0.f, e5 = 0.f, e6 = 0.f;
        if (valid) {
            if (y_dim == 0) { e0 = 1.f; }
            else if (y_dim == 1) { e0 = h0; e1 = y0 - h0; e2 = 1.f; }
            else if (y_dim == 2) { e0 = h0; e1 = y0 - h0; e2 = h1; e3 = y1 - h1; e4 = 1.f; }
            else { e0 = h0; e1 = y0 - h0; e2 = h1; e3 = y1 - h1; e4 = h2; e5 = y2 - h2; e6 = 1.f; }
        }
        constexpr int c = 2 * CH;
        const int kb = c >> 3, cc = c & 7;
        uint4 pk = make_uint4(pack_bf16x2(e0, e1), pack_bf16x2(e2, e3), pack_bf16x2(e4, e5), pack_bf16x2(e6, 0.f));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = pk;
    } else {
        const int K1c = 8 * nkb1;              // remaining chunks of the K blocks in use are zero
        for (int c = 2 * CH + 1; c < K1c; ++c) {
            const int kb = c >> 3, cc = c & 7;
            *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

__device__ __forceinline__ void mbar_arrive2(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// 16 bins of the log-likelihood with the observation / noise variance streamed at BF16 precision (pack_pv_kernel): half
// the L2 traffic and half the L2 working set of an FP32 stream; the rounding (2^-9 relative, identical for l(z) and
// l(z')) perturbs the log ratio far less than the BF16 decoder weights do.
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// Four bins share ONE reciprocal and ONE logarithm:
//   ln(v0 v1 v2 v3)   and   sum_i p_i / v_i = (n01 * p23 + n23 * p01) / (p01 * p23),   p01 = v0 v1,  n01 = p0 v1 + p1 v0.
// The pair products are scaled by 2^15 so that the product of four variances stays inside the FP32 range for
// Vx in [1e-10, 1e7] (the observation is |STFT|^2 of audio in [-1, 1]: at most 2.6e5); the scale leaves `acc` short
// of a factor 2^15 (applied once per row by the caller: kQuadScale) and adds a constant to `accl` that cancels in
// l(z) - l(z').  -> 1.5 MUFU operations per bin instead of 2.
// Range: the stream keeps every factor within +-30 octaves of the frame's level (pack_pv floors Vb' 90 dB below it and gives the
// observation-free padding bins the constant 1 in the quad's scale), so the product cannot underflow and this loop carries no check.
// The scale is a power of two chosen PER FRAME by row_scale_kernel (k^2 X^4 ~ 1 for the frame's typical X = Vx 2^-b): with the
// fixed 2^15 of round 1 a decoder whose output bias is very small in some bins (2^-b ~ 1e10, e.g. a prior without energy above
// 3 kHz) overflowed the four-fold product, l(z') became Inf and those chains never moved.
constexpr float kDefaultQuadScale = 32768.0f;
// 2^x for a packed pair on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f through the 1.5 * 2^23 magic
// add, degree-3 minimax polynomial for 2^f on [-1/2, 1/2] (max relative error 7.5e-5, far inside the BF16 noise of the
// layer-3 pre-activation), exponent patched with an integer add.  Valid for |x| < 125: callers must know the range
// (DVAE_TC_POLY_EX2, dvae_tc_decoder_exponent_bound).
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 x) {
    const f32x2 t = add2(x, pk2(12582912.0f, 12582912.0f));
    const f32x2 n = add2(t, pk2(-12582912.0f, -12582912.0f));
    const f32x2 f = fma2(n, pk2(-1.0f, -1.0f), x);
    f32x2 p = fma2(pk2(0.0551716685f, 0.0551716685f), f, pk2(0.242611125f, 0.242611125f));
    p = fma2(p, f, pk2(0.693260968f, 0.693260968f));
    p = fma2(p, f, pk2(0.999928057f, 0.999928057f));
    float plo, phi, tlo, thi;
    upk2(p, plo, phi);
    upk2(t, tlo, thi);
    return pk2(__uint_as_float(__float_as_uint(plo) + (__float_as_uint(tlo) << 23)),
               __uint_as_float(__float_as_uint(phi) + (__float_as_uint(thi) << 23)));
}

// POLY: bins 2,3 of every quad take 2^v from ex2_poly2 instead of MUFU.EX2.  The sampler is bound by the MUFU pipe
// (4 lanes / clk / scheduler: profiles/r01_tc_ncu_mh2.txt shows XU at 51 % with the issue slots at 28 %), so moving a
// third of its transcendental work to the idle FMA pipe shortens the layer-3 epilogue.
// POLY = 2 (DVAE_TC_POLY_EX2_ALL): all four bins of a quad from the polynomial -> 0.5 MUFU operations per bin (rcp, lg2 only).
template <int POLY>
__device__ __forceinline__ void loglik16_pv(const float* v, const uint4* pv, float g_row, float k_row, float& acc, float& accl) {
    // Stream format: see pack_pv_kernel (bias folded in, word j = P'_j with bf16(Vb'_j) in its low half, quad scale k
    // pre-applied to Vb' of bins 0,1 and 1/k to P' of bins 2,3).  Packed FP32 pairs:
    //   A = k (X0, X1),  B = (X2, X3),  X_j = g 2^v_j + Vb'_j
    //   M = A * B = (k X0 X2, k X1 X3)
    //   N = P_A * B + (P_B / k) * A = (P0 X2 + P2 X0, P1 X3 + P3 X1)
    // so that  sum_j P'_j / X_j = (N.lo M.hi + N.hi M.lo) / (M.lo M.hi) / k  and  sum_j log2 X_j = log2(M.lo M.hi) - 30.
    const f32x2 g2 = pk2(g_row, g_row), g2k = pk2(g_row * k_row, g_row * k_row);
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        const uint4 w = pv[qd];
        const f32x2 e01 = (POLY >= 2) ? ex2_poly2(pk2(v[4 * qd + 0], v[4 * qd + 1])) : pk2(ex2_approx(v[4 * qd + 0]), ex2_approx(v[4 * qd + 1]));
        const f32x2 A = fma2(g2k, e01, pk2(__uint_as_float(w.x << 16), __uint_as_float(w.y << 16)));
        const f32x2 e23 = POLY ? ex2_poly2(pk2(v[4 * qd + 2], v[4 * qd + 3])) : pk2(ex2_approx(v[4 * qd + 2]), ex2_approx(v[4 * qd + 3]));
        const f32x2 B = fma2(g2, e23, pk2(__uint_as_float(w.z << 16), __uint_as_float(w.w << 16)));
        const f32x2 M = mul2(A, B);
        const f32x2 N = fma2(pk2(__uint_as_float(w.x), __uint_as_float(w.y)), B,
                             mul2(pk2(__uint_as_float(w.z), __uint_as_float(w.w)), A));
        float mlo, mhi, nlo, nhi;
        upk2(M, mlo, mhi);
        upk2(N, nlo, nhi);
        const float pq = mlo * mhi;
        acc = fmaf(fmaf(nlo, mhi, nhi * mlo), rcp_approx(pq), acc);
        accl += lg2_approx(pq);
    }
}


// Variant with packed BF16 tanh: the pre-activation pair is rounded to bf16x2 first and one MUFU operation
// (tanh.approx.bf16x2) yields both activations already in the operand's storage format -> half the MUFU work of the
// hidden layers.  Used by the samplers (the decode keeps FP32 tanh).
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) { uint32_t y; asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ void hidden_rows_bf_part(const float* v, unsigned char* A, int h, int row, int part, const float* bias) {
    const int col0 = 64 * h + 32 * part;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x0 = v[8 * cc + 2 * e], x1 = v[8 * cc + 2 * e + 1];
            if (bias) { x0 += bias[col0 + 8 * cc + 2 * e]; x1 += bias[col0 + 8 * cc + 2 * e + 1]; }
            w[e] = tanh_bf16x2(pack_bf16x2(x0, x1));
        }
        const int chunk = 4 * part + cc;
        *reinterpret_cast<uint4*>(A + h * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
// The second half of the thread's columns is requested as soon as the first has arrived: its TMEM latency runs under the
// tanh / pack / store work of the first half.
__device__ __forceinline__ void hidden_epilogue_rows_bf(uint32_t tmem, unsigned char* A, int q, int h, int row, const float* bias) {
    float v0[32], v1[32];
    const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + 64 * h;
    tmem_ld32(t0, v0);
    tmem_wait_ld();
    tmem_ld32(t0 + 32, v1);
    hidden_rows_bf_part(v0, A, h, row, 0, bias);
    tmem_wait_ld();
    hidden_rows_bf_part(v1, A, h, row, 1, bias);
}

// Preload of a layer's bias into the thread's own 64 accumulator columns (lanes 32q.., columns 64h..): the following
// GEMM accumulates onto it, so the epilogue has no bias loads / adds (they cost the layer-2 epilogue of the sampler 0.5 k of
// its 1.7 k cycles, measured with per-phase clock stamps in round 1).  bias64: 64 floats in shared memory, 16-byte aligned.  Asynchronous: the caller
// runs tmem_wait_st() (and the tcgen05 fence) before handing the columns over.
__device__ __forceinline__ void tmem_preload_bias64(uint32_t taddr, const float* bias64) {
    const uint32_t bias_s = smem_u32(bias64);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // explicit shared-space loads (through the generic pointer they compiled to LD.E and queued behind global traffic)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[4 * i]), "=r"(r[4 * i + 1]), "=r"(r[4 * i + 2]), "=r"(r[4 * i + 3])
                         : "r"(bias_s + 16u * (uint32_t)(8 * part + i)));
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            :: "r"(taddr + 32 * part), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
               "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
               "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
    }
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

size_t smem_bytes(const Dims& d);
int check_dims(const DvaeMlp* dec, int L, int y_dim, const char* who, Dims* out);

}  // namespace tc
}  // namespace dvae
