// Label front end on the device: time-domain voice-activity labels and ideal binary masks.
//
// Replaces packages/processing/target.py:5-56 (clean_speech_VAD), 58-70 (clean_speech_IBM) and 72-105
// (noise_robust_clean_speech_IBM) for the configuration the evaluation scripts use (center=False, end-pad rule of the
// STFT): labels are produced in the frame-major ragged layout of the enhancement path, so y for the M2 / M2-info models
// never leaves the GPU.
//
//   VAD : power[j] = sum of squares of frame j of the (zero-padded) signal; label = power > 10^thr * min_j power
//   IBM : mask[j][f] = 20 log10(|S[j][f]| + eps) > max_{j,f} 20 log10(|S| + eps) - thr_db   (optionally times the VAD label)
//
// Frame powers and the dB comparison run in double: the reference works on float64 signals, and a float32 power would
// flip labels of frames that sit on the threshold.
#include "common.cuh"

namespace dvae {

// one warp per frame
__global__ void __launch_bounds__(256) frame_power_kernel(const float* __restrict__ x, const int64_t* __restrict__ x_off,
                                                          const int32_t* __restrict__ x_len, const int32_t* __restrict__ frame_utt,
                                                          const int64_t* __restrict__ fr_off, int64_t NT, int n_fft, int hop,
                                                          double* __restrict__ power) {
    const int lane = threadIdx.x & 31;
    const int64_t n = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= NT) return;
    const int u = frame_utt[n];
    const int64_t s0 = (n - fr_off[u]) * (int64_t)hop;
    const float* xu = x + x_off[u];
    const int64_t len = x_len[u];
    double acc = 0.0;
    for (int i = lane; i < n_fft; i += 32) {
        const int64_t q = s0 + i;
        const double v = (q < len) ? (double)xu[q] : 0.0;          // samples beyond the signal are the end padding
        acc = fma(v, v, acc);
    }
    acc = warp_sum_d(acc);
    if (lane == 0) power[n] = acc;
}

// one CTA per utterance: minimum frame power, then the labels
__global__ void __launch_bounds__(256) vad_threshold_kernel(const double* __restrict__ power, const int64_t* __restrict__ fr_off,
                                                            double factor, float* __restrict__ vad) {
    __shared__ double red[8];
    const int u = blockIdx.x;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    double m = 1.0e300;
    for (int64_t n = n0 + threadIdx.x; n < n1; n += blockDim.x) m = fmin(m, power[n]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmin(m, red[w]);
    const double thr = factor * m;
    for (int64_t n = n0 + threadIdx.x; n < n1; n += blockDim.x) vad[n] = (power[n] > thr) ? 1.f : 0.f;
}

// per-utterance maximum magnitude (blockIdx.y = utterance, partial maxima combined with an ordered-int atomic)
__global__ void __launch_bounds__(256) max_mag_kernel(const float2* __restrict__ S, const int64_t* __restrict__ fr_off, int F,
                                                      int ld, unsigned int* __restrict__ max_bits) {
    const int u = blockIdx.y;
    const int64_t n0 = fr_off[u], n1 = fr_off[u + 1];
    float m = 0.f;
    for (int64_t n = n0 + blockIdx.x; n < n1; n += gridDim.x)
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const float2 v = S[n * ld + f];
            m = fmaxf(m, hypotf(v.x, v.y));
        }
    m = fmaxf(m, 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits + u, __float_as_uint(m));      // non-negative floats order like their bits
}

__global__ void __launch_bounds__(256) ibm_kernel(const float2* __restrict__ S, const int32_t* __restrict__ frame_utt,
                                                  const unsigned int* __restrict__ max_bits, const float* __restrict__ vad,
                                                  int64_t NT, int F, int ld, double eps, double ratio, float* __restrict__ mask) {
    const int64_t total = NT * ld;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / ld;
        const int f = (int)(i - n * ld);
        float r = 0.f;
        if (f < F) {
            const float2 v = S[i];
            // 20 log10(mag + eps) > 20 log10(max + eps) - thr  <=>  mag + eps > (max + eps) 10^(-thr / 20)
            const double lhs = (double)hypotf(v.x, v.y) + eps;
            const double rhs = ((double)__uint_as_float(max_bits[frame_utt[n]]) + eps) * ratio;
            r = (lhs > rhs) ? 1.f : 0.f;
            if (vad) r *= vad[n];
        }
        mask[i] = r;
    }
}

}  // namespace dvae

using namespace dvae;

extern "C" int64_t dvae_vad_workspace_bytes(int64_t NT) { return NT > 0 ? NT * (int64_t)sizeof(double) : 0; }

extern "C" int dvae_vad_labels(const float* x, const int64_t* x_off, const int32_t* x_len, int B, const int32_t* frame_utt,
                               const int64_t* fr_off, int64_t NT, int n_fft, int hop, float vad_threshold, float* vad,
                               void* ws, void* stream) {
    DVAE_REQUIRE(x && x_off && x_len && frame_utt && fr_off && vad && ws, "dvae_vad_labels: null pointer");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && n_fft >= 1 && hop >= 1, "dvae_vad_labels: bad sizes");
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "dvae_vad_labels: workspace must be 8-byte aligned");
    if (NT == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    double* power = reinterpret_cast<double*>(ws);
    frame_power_kernel<<<(unsigned)((NT + 7) / 8), 256, 0, st>>>(x, x_off, x_len, frame_utt, fr_off, NT, n_fft, hop, power);
    int rc = check_launch("frame_power_kernel");
    if (rc) return rc;
    vad_threshold_kernel<<<B, 256, 0, st>>>(power, fr_off, pow(10.0, (double)vad_threshold), vad);
    return check_launch("vad_threshold_kernel");
}

extern "C" int dvae_ibm_labels(const void* S, const int32_t* frame_utt, const int64_t* fr_off, int B, int64_t NT, int F, int ld,
                               float eps, float ibm_threshold_db, const float* vad, float* mask, void* ws, void* stream) {
    DVAE_REQUIRE(S && frame_utt && fr_off && mask && ws, "dvae_ibm_labels: null pointer");
    DVAE_REQUIRE(B >= 1 && NT >= 0 && F >= 1 && ld >= F, "dvae_ibm_labels: bad sizes");
    if (NT == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* max_bits = reinterpret_cast<unsigned int*>(ws);                      // B words
    cudaError_t e = cudaMemsetAsync(max_bits, 0, sizeof(unsigned int) * B, st);
    if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    max_mag_kernel<<<dim3(16, B), 256, 0, st>>>((const float2*)S, fr_off, F, ld, max_bits);
    int rc = check_launch("max_mag_kernel");
    if (rc) return rc;
    const int64_t blocks = (NT * ld + 255) / 256;
    ibm_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>((const float2*)S, frame_utt, max_bits, vad, NT, F, ld,
                                                                                 (double)eps, pow(10.0, -(double)ibm_threshold_db / 20.0), mask);
    return check_launch("ibm_kernel");
}
