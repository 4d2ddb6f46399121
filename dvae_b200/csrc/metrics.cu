// Scale-invariant energy ratios on the device: SI-SDR, SI-SIR, SI-SAR of a ragged batch.
//
// Replaces packages/metrics.py:12-82 (si_sdr_components / energy_ratios / si_sdr_leroux), which works on one pair of
// host vectors at a time.  With the six inner products  a = <s_hat,s>, b = <s_hat,n>, c = <s,s>, d = <n,n>,
// e = <s_hat,s_hat>, f = <s,n>  (accumulated in double) the reference's vectors never have to be formed:
//
//   alpha_s = a / c, alpha_n = b / d
//   |s_target|^2        = alpha_s^2 c
//   |e_noise|^2         = alpha_n^2 d
//   |e_noise + e_art|^2 = |s_hat - s_target|^2 = e - 2 alpha_s a + alpha_s^2 c
//   |e_art|^2           = e + alpha_s^2 c + alpha_n^2 d - 2 alpha_s a - 2 alpha_n b + 2 alpha_s alpha_n f
//
// One CTA per utterance; per-thread double partial sums in a fixed order, fixed-order tree reduction: deterministic.
// HBM-bound: 3 x 4 T bytes read per utterance.
#include "common.cuh"

namespace dvae {

__global__ void __launch_bounds__(256) energy_ratios_kernel(const float* __restrict__ s_hat, const float* __restrict__ s,
                                                            const float* __restrict__ n, const int64_t* __restrict__ off,
                                                            const int32_t* __restrict__ len, double* __restrict__ out) {
    __shared__ double red[6][8];
    const int u = blockIdx.x;
    const int64_t o = off[u];
    const int T = len[u];
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const double x = (double)s_hat[o + i], y = (double)s[o + i], z = n ? (double)n[o + i] : 0.0;
        acc[0] = fma(x, y, acc[0]);
        acc[1] = fma(x, z, acc[1]);
        acc[2] = fma(y, y, acc[2]);
        acc[3] = fma(z, z, acc[3]);
        acc[4] = fma(x, x, acc[4]);
        acc[5] = fma(y, z, acc[5]);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double v = warp_sum_d(acc[k]);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = 0.0;
            for (int w = 0; w < 8; ++w) v += red[k][w];
            t[k] = v;
        }
        const double a = t[0], b = t[1], c = t[2], d = t[3], e = t[4], f = t[5];
        const double as = a / c;
        const double tgt = as * as * c;
        const double dist = e - 2.0 * as * a + tgt;                      // |s_hat - s_target|^2
        double* r = out + 3 * (int64_t)u;
        r[0] = 10.0 * log10(tgt / dist);
        if (n) {
            const double an = b / d;
            const double noise = an * an * d;
            const double art = e + tgt + noise - 2.0 * as * a - 2.0 * an * b + 2.0 * as * an * f;
            r[1] = 10.0 * log10(tgt / noise);
            r[2] = 10.0 * log10(tgt / art);
        } else {
            r[1] = nan("");
            r[2] = nan("");
        }
    }
}

}  // namespace dvae

using namespace dvae;

extern "C" int dvae_energy_ratios(const float* s_hat, const float* s, const float* n, const int64_t* off, const int32_t* len,
                                  int B, double* out, void* stream) {
    DVAE_REQUIRE(s_hat && s && off && len && out && B >= 0, "dvae_energy_ratios: bad arguments");
    if (B == 0) return 0;
    energy_ratios_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(s_hat, s, n, off, len, out);
    return check_launch("energy_ratios_kernel");
}
