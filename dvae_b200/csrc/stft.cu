// STFT / ISTFT for the reference's only configuration: n_fft = 1024, periodic Hann, center=False.
//
// Replaces packages/processing/stft.py:13-60 / 63-99 (thin wrappers over librosa.core.stft / istft).
//
// One 1024-point real FFT is computed as a 512-point complex FFT of the even/odd-packed frame followed by the
// standard split post-processing.  The 512-point FFT is three radix-8 passes (512 = 8*8*8) executed by 64 threads
// that hold 8 points each in registers; the two digit exchanges go through shared memory (two conflict-free layouts with
// constant per-access offsets, see kXRow), twiddles come from a 1024-entry table built once per device in double.
// A CTA of 256 threads works on four frames at a time.
//
// HBM traffic per utterance (T samples, N frames, F = 513): STFT reads 4T (frame overlap is served by L1/L2),
// writes 8FN (+4FN for |X|^2); ISTFT reads 8FN, writes 4T.
#include <float.h>

#include <mutex>

#include "common.cuh"

namespace dvae {

constexpr int kNfft = 1024;
constexpr int kHalf = 512;
constexpr int kRow = 64;                       // natural-order row stride (the STFT's post-processing view of the buffer)
// The two digit exchanges of the 512-point FFT use two different layouts of the same buffer, both chosen so that EVERY access
// of a thread is "thread-constant base + compile-time offset" (one instruction) and conflict-free per half-warp of 8-byte
// accesses (16 banks of float2):
//   exchange 1, element (k1, c = t):            slot = 72 k1 + c
//       stage-1 stores: k1 fixed, 16 consecutive c;  stage-2 loads: rows k1 in {2a, 2a+1} (72 = 8 mod 16 banks apart) x c = 8 m1 + 0..7
//   exchange 2, element (j1, k1, m2):           slot = 72 j1 + k1 + 16 (k1 >> 1) + 2 m2
//       stage-2 stores: j1 fixed, (k1 & 1, m2) -> banks k1 + 2 m2 + const, all distinct;  stage-3 loads: m2 fixed, k1 = 0..7 and
//       j1 in {2b, 2b+1} -> banks k1 + 8 j1 + const, all distinct.  (k1 + 16 (k1 >> 1) + 2 m2 is injective, <= 69 < 72.)
// The first versions used one XOR-swizzled layout for both exchanges: conflict-free too, but every access cost a LOP3 and an
// address add on top of the load / store (3.6 instructions per access, ncu) - a tenth of the kernels' instructions.
constexpr int kXRow = 72;
constexpr int kBuf = 8 * kXRow;                // float2 per frame group

__device__ float2 g_tw[kNfft];                 // exp(-2*pi*i*k/1024)
__device__ float g_win[kNfft];                 // periodic Hann
__device__ double g_win2[kNfft];               // Hann^2 in double (librosa window_sumsquare accumulates these into f32)
// Window-sum-square values for hop = 256 (four frames overlap a sample).  A sample s covered by the frames jlo..jhi gets
// wss = (((0 + w2[o]) + w2[o - 256]) + ...) with o = s - jlo * 256, accumulated in ascending frame order into float32
// exactly like librosa's window_sumsquare; that value only depends on q = o / 256, r = o % 256 and the number of frames
// cnt = jhi - jlo + 1 <= q + 1, so all of them fit a 10 x 256 table: entry (q (q + 1) / 2 + cnt - 1, r).
__device__ float g_wss[10 * 256];
__device__ float g_wss_inv[10 * 256];            // 1 / wss where wss > FLT_MIN (the reference divides only there), else 1

__global__ void init_tables_kernel() {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kNfft) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)kNfft, &s, &c);
    g_tw[k] = make_float2((float)c, (float)(-s));
    const double w = 0.5 - 0.5 * c;
    g_win[k] = (float)w;
    g_win2[k] = w * w;
    if (k < 256) {
        for (int q = 0; q < 4; ++q)
            for (int cnt = 1; cnt <= q + 1; ++cnt) {
                float wss = 0.f;
                for (int i = 0; i < cnt; ++i) {
                    double s2, c2;
                    sincospi(2.0 * (double)((q - i) * 256 + k) / (double)kNfft, &s2, &c2);
                    const double wi = 0.5 - 0.5 * c2;
                    wss = (float)((double)wss + wi * wi);
                }
                g_wss[(q * (q + 1) / 2 + cnt - 1) * 256 + k] = wss;
                g_wss_inv[(q * (q + 1) / 2 + cnt - 1) * 256 + k] = (wss > FLT_MIN) ? 1.0f / wss : 1.0f;
            }
    }
}

static std::mutex g_init_mutex;
static bool g_inited[64] = {false};

int ensure_tables(cudaStream_t stream) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    std::lock_guard<std::mutex> lock(g_init_mutex);
    if (dev < 64 && g_inited[dev]) return 0;
    init_tables_kernel<<<kNfft / 256, 256, 0, stream>>>();
    int rc = check_launch("init_tables");
    if (rc) return rc;
    // one-off: the tables must be complete before the device is marked initialised, or a later call on ANOTHER stream
    // could read them half-written
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_error("init_tables: %s", cudaGetErrorString(e)); return (int)e; }
    if (dev < 64) g_inited[dev] = true;
    return 0;
}

// ----------------------------------------------------------------------------- complex helpers
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// complex add / subtract as ONE packed FP32 instruction (FADD2 / FFMA2 on the (re, im) register pair)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 o;
    upk2(add2(pk2(a.x, a.y), pk2(b.x, b.y)), o.x, o.y);
    return o;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 o;
    upk2(fma2(pk2(b.x, b.y), pk2(-1.0f, -1.0f), pk2(a.x, a.y)), o.x, o.y);
    return o;
}
__device__ __forceinline__ float2 cadd_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }    // a + b * (-i)
__device__ __forceinline__ float2 cadd_pi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }    // a + b * (+i)

// 8-point forward DFT (e^{-2 pi i nk/8}), natural order in and out, in registers.
__device__ __forceinline__ void fft8(float2* a) {
    const float s = 0.70710678118654752440f;
    const float2 b0 = cadd(a[0], a[4]), b1 = csub(a[0], a[4]);
    const float2 b2 = cadd(a[2], a[6]), b3 = csub(a[2], a[6]);
    const float2 b4 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    const float2 b6 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    const float2 e0 = cadd(b0, b2), e1 = cadd_mi(b1, b3), e2 = csub(b0, b2), e3 = cadd_pi(b1, b3);
    const float2 o0 = cadd(b4, b6), o1 = cadd_mi(b5, b7), o2 = csub(b4, b6), o3 = cadd_pi(b5, b7);
    const float2 t1 = make_float2(s * (o1.x + o1.y), s * (o1.y - o1.x));      // o1 * (1-i)/sqrt2
    const float2 t3 = make_float2(s * (o3.y - o3.x), -s * (o3.x + o3.y));     // o3 * (-1-i)/sqrt2
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, t1); a[5] = csub(e1, t1);
    a[2] = cadd_mi(e2, o2); a[6] = cadd_pi(e2, o2);                           // e2 -+ i o2
    a[3] = cadd(e3, t3); a[7] = csub(e3, t3);
}

// 512-point forward complex FFT by the 64 threads (two warps) of frame group `grp`.
// in : a[n1] = z[t + 64*n1]            (t = thread in group)
// out: a[j2] = Z[t + 64*j2]
// `buf` is the group's kBuf-float2 exchange buffer.  The groups of a CTA are independent: they synchronise on their own named
// barrier (id 1 + grp, 64 threads), not on the CTA barrier.
__device__ __forceinline__ void group_bar(int grp) {                        // immediate ids: a register id makes ptxas reserve all 16 barriers
    switch (grp) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}

// Inter-stage twiddles w^k (k = 1..7) are powers of ONE thread-constant value per exchange, formed in registers (4 multiply-adds
// each).  History: the first version fetched all of them from the 1024-entry table with strides of up to 14 entries (8-way bank
// conflicts, shared-memory pipe 73 % busy); conflict-free per-thread tables (14 loads per frame and thread) were faster than
// that, but once the exchanges themselves were down to one instruction per access the bound of both kernels was the
// shared-memory / L1 data pipe itself (l1tex__data_pipe_lsu_wavefronts 87 % of peak, one 128-byte wavefront per clock and SM,
// profiles/r02_ncu_stft.txt): a 64-bit load costs that pipe as much time as eight instructions cost the four schedulers.
__device__ __forceinline__ void twiddle_powers(float2 w, float2* wp) {     // wp[k] = w^k, k = 1..7 (wp[0] unused)
    wp[1] = w;
    wp[2] = cmul(w, w);
    wp[3] = cmul(wp[2], w);
    wp[4] = cmul(wp[2], wp[2]);
    wp[5] = cmul(wp[4], w);
    wp[6] = cmul(wp[4], wp[2]);
    wp[7] = cmul(wp[4], wp[3]);
}

struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `after_stage1()` runs right after the first group barrier, i.e. once every thread of the group has consumed its input
// registers' sources (the ISTFT issues the next pass's asynchronous row copies there).
// w1 = exp(-2 pi i t / 512), w2 = exp(-2 pi i (t & 7) / 64): the thread's twiddle bases (g_tw[2 t], g_tw[16 (t & 7)])
template <class Hook = NoHook>
__device__ __forceinline__ void fft512(float2* a, float2* buf, float2 w1, float2 w2, int t, int grp, Hook after_stage1 = Hook()) {
    fft8(a);
    {
        float2 wp[8];
        twiddle_powers(w1, wp);
        float2* st = buf + t;
        st[0] = a[0];
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) st[kXRow * k1] = cmul(a[k1], wp[k1]);      // exp(-2 pi i t k1 / 512)
    }
    group_bar(grp);
    after_stage1();
    {
        const int k1 = t >> 3, m2 = t & 7;
        const float2* ld = buf + kXRow * k1 + m2;
#pragma unroll
        for (int m1 = 0; m1 < 8; ++m1) a[m1] = ld[8 * m1];
        group_bar(grp);                            // exchange 1 is consumed: the buffer may be rewritten in the second layout
        fft8(a);
        float2 wp[8];
        twiddle_powers(w2, wp);
        float2* st = buf + k1 + 16 * (k1 >> 1) + 2 * m2;
        st[0] = a[0];
#pragma unroll
        for (int j1 = 1; j1 < 8; ++j1) st[kXRow * j1] = cmul(a[j1], wp[j1]);      // exp(-2 pi i m2 j1 / 64)
    }
    group_bar(grp);
    {
        const int k1 = t & 7, j1 = t >> 3;
        const float2* ld = buf + kXRow * j1 + k1 + 16 * (k1 >> 1);
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) a[m2] = ld[2 * m2];
        fft8(a);                                   // a[j2] = Z[k1 + 8*j1 + 64*j2] = Z[t + 64*j2]
    }
    group_bar(grp);                                // buf may be reused by the caller
}

__device__ __forceinline__ int find_utt(const int64_t* __restrict__ fr_off, int B, int64_t n) {
    {   // equal-length batches (the common case): the proportional guess is right and costs one round of two loads
        // instead of a chain of log2(B) dependent ones
        // (single-precision quotient: no 64-bit division; a wrong guess only falls through to the search below)
        const float NTf = (float)fr_off[B];
        int g = (int)((float)n * (float)B / fmaxf(NTf, 1.0f));
        g = min(max(g, 0), B - 1);
        if (fr_off[g] <= n && n < fr_off[g + 1]) return g;
        if (g + 1 < B && fr_off[g + 1] <= n && n < fr_off[g + 2]) return g + 1;
        if (g > 0 && fr_off[g - 1] <= n && n < fr_off[g]) return g - 1;
    }
    int lo = 0, hi = B;                            // largest u with fr_off[u] <= n
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (fr_off[mid] <= n) lo = mid; else hi = mid;
    }
    return lo;
}

// ----------------------------------------------------------------------------- STFT
// 80 registers (three resident CTAs per SM): with 64 the two look-ahead stages spill, and the spill-free variants without them are
// slower (A/B at B = 4 096: 64 registers with spills 1.30 ms, without the utterance look-ahead 1.35, this one 1.20).
__global__ void __launch_bounds__(256, 3) stft_kernel(const float* __restrict__ x, const int64_t* __restrict__ x_off,
                                                   const int32_t* __restrict__ x_len, int B, float2* __restrict__ X,
                                                   float* __restrict__ P, const int64_t* __restrict__ fr_off,
                                                   int64_t NT, int hop, int ld) {
    __shared__ __align__(8) float win[kNfft];
    __shared__ float2 bufs[4][kBuf];
    for (int i = threadIdx.x; i < kNfft; i += blockDim.x) win[i] = g_win[i];
    __syncthreads();

    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
    float2* buf = bufs[grp];
    const float2 w1 = g_tw[2 * t], w2 = g_tw[16 * (t & 7)], wt = g_tw[t];
    const int64_t n_pass = (NT + 3) / 4;
    // raw samples of frame n (zeros beyond the signal): 8 even / odd pairs per thread.  The frame of the NEXT pass is
    // requested before the current one is transformed, so its DRAM latency (half of all stall samples in the first
    // version, profiles/r01_stft_microbench.json) hides behind the FFT; the utterance of the frame AFTER that is looked up
    // speculatively at the same time (proportional guess g, then fr_off[g], fr_off[g+1], x_off[g], x_len[g] in one round of
    // independent loads that nobody waits for until the next pass), which takes the chain of dependent loads
    // frame -> utterance -> signal pointer -> samples out of the loop's critical path.
    struct Look { int len; int64_t lo, hi, xo; };
    auto look = [&](int64_t n) {
        Look L;
        L.len = 0; L.lo = 1; L.hi = 0; L.xo = 0;
        if (n < NT) {
            // single-precision quotient: no 64-bit division; a wrong guess falls back to the search in fetch()
            int g = (int)((float)n * (float)B / (float)NT);
            g = min(max(g, 0), B - 1);
            L.lo = __ldg(fr_off + g); L.hi = __ldg(fr_off + g + 1); L.xo = __ldg(x_off + g); L.len = __ldg(x_len + g);
        }
        return L;
    };
    auto fetch = [&](int64_t n, const Look& L, float2* v) {
        if (n < NT) {
            int64_t lo = L.lo, xo = L.xo, len = L.len;
            if (!(L.lo <= n && n < L.hi)) {
                const int u = find_utt(fr_off, B, n);
                lo = fr_off[u]; xo = x_off[u]; len = x_len[u];
            }
            const float* xu = x + xo;
            const int64_t s0 = (n - lo) * (int64_t)hop;
            // one 8-byte load per pair when the frame start is 8-byte aligned and the whole frame lies inside the signal
            const bool fast = ((reinterpret_cast<uintptr_t>(xu + s0) & 7) == 0) && (s0 + kNfft <= len);
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int64_t q = s0 + 2 * (t + 64 * n1);
                if (fast) {
                    v[n1] = __ldg(reinterpret_cast<const float2*>(xu + q));
                } else {
                    v[n1].x = (q < len) ? __ldg(xu + q) : 0.f;
                    v[n1].y = (q + 1 < len) ? __ldg(xu + q + 1) : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) v[n1] = make_float2(0.f, 0.f);
        }
    };
    float2 nxt[8];
    fetch((int64_t)blockIdx.x * 4 + grp, look((int64_t)blockIdx.x * 4 + grp), nxt);
    Look ahead = look(((int64_t)blockIdx.x + gridDim.x) * 4 + grp);
    for (int64_t pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
        const int64_t n = pass * 4 + grp;
        const bool live = n < NT;
        float2 a[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const float2 w2 = *reinterpret_cast<const float2*>(win + 2 * (t + 64 * n1));
            upk2(mul2(pk2(nxt[n1].x, nxt[n1].y), pk2(w2.x, w2.y)), a[n1].x, a[n1].y);       // one packed multiply per sample pair
        }
        fetch((pass + gridDim.x) * 4 + grp, ahead, nxt);
        ahead = look((pass + 2 * (int64_t)gridDim.x) * 4 + grp);
        fft512(a, buf, w1, w2, t, grp);
        // Split step.  X[k] = (s - i W^k d) / 2 with s = Z[k] + conj Z[512-k], d = Z[k] - conj Z[512-k]; the mirrored bin uses the
        // same two values: X[512-k] = conj(s + i W^k d) / 2 (W^(512-k) = -conj W^k), so a thread produces the pairs (k, 512 - k) of
        // its own k = t + 64 j, j < 4 (registers a[0..3]) and only needs Z[512 - k] = a[7 - j] of thread 64 - t: the upper halves
        // a[4..7] go through the buffer (natural order: Z[k] at slot k - 256), nothing else does.
#pragma unroll
        for (int j2 = 4; j2 < 8; ++j2) buf[(j2 - 4) * kRow + t] = a[j2];
        group_bar(grp);
        if (live) {
            float2* Xn = X + n * (int64_t)ld;
            float* Pn = P ? P + n * (int64_t)ld : nullptr;
            auto bins = [&](int k, float2 wk, float2 zk, float2 zm, bool both) {
                zm.y = -zm.y;                                               // conj(Z[512-k])
                const float2 s = cadd(zk, zm), d = csub(zk, zm);
                const float2 wd = cmul(wk, d);                              // W^k (Zk - conj Zm)
                const float2 r = make_float2(0.5f * (s.x + wd.y), 0.5f * (s.y - wd.x));
                Xn[k] = r;
                if (Pn) Pn[k] = r.x * r.x + r.y * r.y;
                if (both) {
                    const float2 q = make_float2(0.5f * (s.x - wd.y), -0.5f * (s.y + wd.x));
                    Xn[kHalf - k] = q;
                    if (Pn) Pn[kHalf - k] = q.x * q.x + q.y * q.y;
                }
            };
            // Z[512 - k] sits at slot 256 - k; k = 0 pairs with Z[512] = Z[0], the thread's own a[0]
            // W^k = W^t W^(64 j): a register pair times a compile-time constant (no table load, see twiddle_powers)
            const float cr[4] = {1.f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f};
            const float ci[4] = {0.f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f};
            bins(t, wt, a[0], t == 0 ? a[0] : buf[kHalf / 2 - t], true);
#pragma unroll
            for (int j = 1; j < 4; ++j)                                     // k = 64 .. 255 and 448 .. 257
                bins(t + 64 * j, cmul(wt, make_float2(cr[j], ci[j])), a[j], buf[kHalf / 2 - t - 64 * j], true);
            if (t == 0) bins(kHalf / 2, make_float2(0.f, -1.f), a[4], a[4], false);      // k = 256, its own mirror (W^256 = -i)
        }
        group_bar(grp);
    }
}

// ----------------------------------------------------------------------------- ISTFT
// hop = 256 = CTA size: thread i owns sample i of every hop.  A pass inverts four consecutive frames (one per frame
// group) and overlap-adds them in ascending frame order - the reference's float32 accumulation order - into SEVEN
// register accumulators per thread (the hops jp .. jp+6 those frames touch).  Hops jp .. jp+3 are final after the pass
// (no later frame reaches them): they are normalised by the window sum-square table and stored, the other three slide
// down.  No accumulator lives in shared memory, a CTA can stream a whole utterance (no frame is transformed twice), and
// only short batches are cut into segments (three lead-in frames each) to fill the GPU.
// MASKED: the frame is multiplied by a real per-bin mask (mask * mask_scale, the Wiener filter of mcem.py:176-177) as it is
// loaded, so that S_hat / N_hat need not be materialised between the filter and the inverse transform.
template <bool MASKED>
__global__ void __launch_bounds__(256, 4) istft_kernel(const float2* __restrict__ X, const float* __restrict__ mask, float mask_scale,
                                                       const int64_t* __restrict__ fr_off,
                                                       float* __restrict__ y, const int64_t* __restrict__ y_off,
                                                       const int32_t* __restrict__ y_len, int /*hop: 256*/, int ld, int seg_hops) {
    constexpr int hop = 256;                                                  // = CTA size (checked by the host wrapper): index math in shifts
    __shared__ __align__(8) float win[kNfft];
    __shared__ __align__(16) float2 bufs[4 * kBuf];                           // FFT exchange buffers; then the windowed frames
    float* fbuf = reinterpret_cast<float*>(bufs);

    const int u = blockIdx.y;
    const int64_t f0 = fr_off[u];
    const int N = (int)(fr_off[u + 1] - f0);
    const int len = y_len[u];
    const int h0 = blockIdx.x * seg_hops;                                     // first output hop of this CTA
    if (h0 * hop >= len) return;
    const int h1 = min(h0 + seg_hops, (len + hop - 1) / hop);                 // one past the last output hop
    const int jstart = max(0, h0 - 3);
    const int jend = min(N - 1, h1 - 1);                                      // last frame that reaches the segment; may be < jstart
    const int sig_len = (N > 0) ? kNfft + hop * (N - 1) : 0;

    for (int i = threadIdx.x; i < kNfft; i += blockDim.x) {
        // synthesis window with the transform's 1 / 1024 (a power of two: the products are the same numbers) and the sign of the
        // conjugation that turns the forward FFT into the inverse one folded in: sample pair m = (Re a * win[2m], -Im a * win[2m+1])
        win[i] = g_win[i] * ((i & 1) ? -1.0f / 1024.0f : 1.0f / 1024.0f);
    }
    __syncthreads();

    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63, i = threadIdx.x;
    const float2 w1 = g_tw[2 * t], w2 = g_tw[16 * (t & 7)], wt = g_tw[t];
    float2* buf = bufs + grp * kBuf;
    float* yu = y + y_off[u];
    const float inv_interior = __ldg(g_wss_inv + (9 << 8) + i);                // four frames cover the sample: q = 3, cnt = 4
    float v[7];
#pragma unroll
    for (int m = 0; m < 7; ++m) v[m] = 0.f;
    // The group's spectrum of the NEXT pass is requested with prefetch instructions (one 128-byte line per thread, 33 lines per
    // row) while the current one is transformed: a quarter of all stall samples of the first version waited for these rows at the
    // top of the pass (profiles/r02_ncu_istft.txt).  Staging the rows in shared memory with 8-byte cp.async one pass ahead was
    // slower than the plain loads it replaced (1.36 against 1.27 ms at B = 4 096: more work for the load / store pipe);
    // prefetching into L1 instead of L2 makes no difference.
    auto prefetch_row = [&](int jq) {
        const int j = jq + grp;
        if (jq < h1 && j <= jend && t < 33) {
            const float2* Xn = X + (f0 + j) * (int64_t)ld + 16 * t;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Xn));
        }
    };
    for (int jp = jstart; jp < h1; jp += 4) {
        if (jp <= jend) {                                                     // CTA-uniform: at least one frame of this pass exists
            const int j = jp + grp;
            const bool live = j <= jend;
            float2 a[8];
            if (live) {
                const float2* Xn = X + (f0 + j) * (int64_t)ld;
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) {
                    const int k = t + 64 * n1;
                    float2 xk = __ldg(Xn + k), xm = __ldg(Xn + kHalf - k);
                    if (MASKED) {                                             // same products as wiener_apply_kernel
                        const float* Mn = mask + (f0 + j) * (int64_t)ld;
                        const float mk = __ldg(Mn + k) * mask_scale, mm = __ldg(Mn + kHalf - k) * mask_scale;
                        xk = make_float2(mk * xk.x, mk * xk.y);
                        xm = make_float2(mm * xm.x, mm * xm.y);
                    }
                    if (k == 0) { xk.y = 0.f; xm.y = 0.f; }                   // irfft ignores Im of DC and Nyquist
                    xm.y = -xm.y;                                             // conj(X[512-k])
                    const float2 s = cadd(xk, xm), d = csub(xk, xm);
                    // W^k = W^t W^(64 n1): the second factor is a compile-time constant, the first sits in two registers - no
                    // table load (the shared-memory / L1 data pipe is the tighter bound here, see twiddle_powers)
                    const float cr[8] = {1.f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f, 0.f,
                                         -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f};
                    const float ci[8] = {0.f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f, -1.f,
                                         -0.92387953251128675613f, -0.70710678118654752440f, -0.38268343236508977173f};
                    float2 w = (n1 == 0) ? wt : cmul(wt, make_float2(cr[n1], ci[n1]));
                    w.y = -w.y;                                               // conj(W^k)
                    const float2 wd = cmul(w, d);
                    const float2 zb = make_float2(s.x - wd.y, s.y + wd.x);    // s + i*wd
                    a[n1] = make_float2(zb.x, -zb.y);                         // conj -> inverse via forward FFT
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) a[n1] = make_float2(0.f, 0.f);
            }
            fft512(a, buf, w1, w2, t, grp, [&] { prefetch_row(jp + 4); });
            {
                float2* fb = buf;                                             // the group's own exchange buffer (1 024 of its 1 152 floats)
#pragma unroll
                for (int j2 = 0; j2 < 8; ++j2) {
                    const int m = t + 64 * j2;
                    const float2 w2 = *reinterpret_cast<const float2*>(win + 2 * m);
                    float2 o;                                                 // a group without a frame transformed zeros: stores +-0
                    upk2(mul2(pk2(a[j2].x, a[j2].y), pk2(w2.x, w2.y)), o.x, o.y);
                    fb[m] = o;
                }
            }
            __syncthreads();
#pragma unroll
            for (int g = 0; g < 4; ++g) {                                     // ascending frame order, like the reference's loop
                if (jp + g <= jend) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[g + c] += fbuf[g * (2 * kBuf) + 256 * c + i];
                }
            }
            __syncthreads();                                                  // fbuf is the next pass's exchange buffer
        }
        // hops jp .. jp+3 are complete
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int hh = jp + m;
            const int s = hh * hop + i;
            if (hh >= h0 && hh < h1 && s < len) {
                float r = 0.f;
                if (hh >= 3 && hh < N) {                                      // frames hh-3 .. hh all exist (the bulk of an utterance)
                    r = v[m] * inv_interior;
                } else if (s < sig_len) {
                    r = v[m];
                    const int jlo = max(0, (s - kNfft + hop) / hop), jhi = min(N - 1, s / hop);
                    const int o = s - jlo * hop, q = o >> 8, cnt = jhi - jlo + 1;
                    // one multiplication by the tabulated reciprocal (<= 1.5 ulp from the reference's division)
                    r *= __ldg(g_wss_inv + ((q * (q + 1) / 2 + cnt - 1) << 8) + (o & 255));
                }
                yu[s] = r;
            }
        }
        v[0] = v[4]; v[1] = v[5]; v[2] = v[6];
        v[3] = 0.f; v[4] = 0.f; v[5] = 0.f; v[6] = 0.f;
    }
}

}  // namespace dvae

using namespace dvae;

extern "C" int dvae_stft_f32(const float* x, const int64_t* x_off, const int32_t* x_len, int B, void* X, float* P,
                             const int64_t* fr_off, int64_t NT, int n_fft, int hop, int ld, void* stream) {
    DVAE_REQUIRE(n_fft == kNfft, "dvae_stft_f32: only n_fft=1024 is implemented (got %d)", n_fft);
    DVAE_REQUIRE(hop >= 1 && hop <= kNfft, "dvae_stft_f32: bad hop %d", hop);
    DVAE_REQUIRE(B >= 1 && NT >= 0 && ld >= kHalf + 1, "dvae_stft_f32: bad sizes B=%d NT=%lld ld=%d", B, (long long)NT, ld);
    DVAE_REQUIRE(x && x_off && x_len && X && fr_off, "dvae_stft_f32: null pointer");
    if (NT == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_tables(st);
    if (rc) return rc;
    const int64_t n_pass = (NT + 3) / 4;
    const int grid = (int)(n_pass < 148 * 6 ? n_pass : 148 * 6);          // three resident CTAs per SM (80 registers, 24 KB), two rounds
    stft_kernel<<<grid, 256, 0, st>>>(x, x_off, x_len, B, (float2*)X, P, fr_off, NT, hop, ld);
    return check_launch("stft_kernel");
}

static int istft_launch(const char* who, const void* X, const float* mask, float mask_scale, const int64_t* fr_off, int B, float* y,
                        const int64_t* y_off, const int32_t* y_len, int max_y_len, int n_fft, int hop, int ld, void* stream) {
    DVAE_REQUIRE(n_fft == kNfft, "%s: only n_fft=1024 is implemented (got %d)", who, n_fft);
    DVAE_REQUIRE(hop == 256, "%s: only hop=256 is implemented (got %d)", who, hop);
    DVAE_REQUIRE(B >= 1 && B <= 65535 && max_y_len >= 0 && ld >= kHalf + 1, "%s: bad sizes (B <= 65535)", who);
    DVAE_REQUIRE(X && fr_off && y && y_off && y_len, "%s: null pointer", who);
    if (max_y_len == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_tables(st);
    if (rc) return rc;
    // whole utterances per CTA when the batch alone fills the GPU (4 resident CTAs per SM), otherwise segments of at least
    // 8 hops (each pays three lead-in frames)
    const int total_hops = (max_y_len + hop - 1) / hop;
    int n_seg = (148 * 4 + B - 1) / B;
    if (n_seg > (total_hops + 7) / 8) n_seg = (total_hops + 7) / 8;
    if (n_seg < 1) n_seg = 1;
    const int seg_hops = (total_hops + n_seg - 1) / n_seg;
    n_seg = (total_hops + seg_hops - 1) / seg_hops;
    if (mask)
        istft_kernel<true><<<dim3(n_seg, B), 256, 0, st>>>((const float2*)X, mask, mask_scale, fr_off, y, y_off, y_len, hop, ld, seg_hops);
    else
        istft_kernel<false><<<dim3(n_seg, B), 256, 0, st>>>((const float2*)X, nullptr, 1.0f, fr_off, y, y_off, y_len, hop, ld, seg_hops);
    return check_launch("istft_kernel");
}

extern "C" int dvae_istft_f32(const void* X, const int64_t* fr_off, int B, float* y, const int64_t* y_off,
                              const int32_t* y_len, int max_y_len, int n_fft, int hop, int ld, void* stream) {
    return istft_launch("dvae_istft_f32", X, nullptr, 1.0f, fr_off, B, y, y_off, y_len, max_y_len, n_fft, hop, ld, stream);
}

extern "C" int dvae_istft_masked_f32(const void* X, const float* mask, float mask_scale, const int64_t* fr_off, int B, float* y,
                                     const int64_t* y_off, const int32_t* y_len, int max_y_len, int n_fft, int hop, int ld,
                                     void* stream) {
    DVAE_REQUIRE(mask, "dvae_istft_masked_f32: null mask");
    return istft_launch("dvae_istft_masked_f32", X, mask, mask_scale, fr_off, B, y, y_off, y_len, max_y_len, n_fft, hop, ld, stream);
}
