// Second-generation tcgen05 Metropolis-Hastings sampler (the default "tc" sampler).
//
// Same math, operand images and TMEM plan as mh_tc.cu (see the notes there); what changes is the schedule inside
// the CTA, driven by the first ncu capture (profiles/r01_tc_v1_*): v1 ran at 0.7 IPC per SM because every phase was
// serialised behind CTA barriers with only two warps per scheduler and a long single-warp "owner" phase.
//
//   * 16 epilogue warps (4 per TMEM lane quadrant, each 32 columns of a chunk) + the control warp: 4 warps per
//     scheduler hide the L2 latency of the P / Vb loads and keep the MUFU pipe fed;
//   * the layer-3 chunks are handed over with mbarriers only (chunk ready: tcgen05.commit; buffer free: 512 thread
//     arrivals), no CTA barrier inside the chunk loop, so warps drift and MMA / epilogue overlap;
//   * the chain state (z, z', l) lives in registers (latent size is a template parameter) and the layer-1 operand
//     row is written with a static layout;
//   * the Philox draws of iteration it+1 are produced by ALL 16 warps (one 4-word block each) right after the last
//     chunk of iteration it and parked in the idle half of the activation buffer.
#include "tc_common.cuh"

namespace dvae {
namespace tc {

constexpr int MH2_THREADS = 544;
constexpr int MH2_EPI = 512;

struct Mh2Params {
    Dims d;
    const unsigned char* image;
    int64_t rows;                 // NT*C chains
    int C;
    const float* y;
    const float4* Ppk;
    const float4* Vbpk;
    const float* g;
    float* Z;
    float* Zs;
    const int32_t* frame_gid;
    const int32_t* frame_idx;
    const float* inj_eps;
    const float* inj_u;
    uint32_t* n_accept;
    float* a_trace;
    int n_burn, n_keep;
    uint32_t seed_lo, seed_hi, iter0;
    float sd;
    int* status;
};

__device__ __forceinline__ void mbar_arrive2(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// hidden-layer epilogue, 32 columns per thread: D12[row][32s .. 32s+32) -> tanh(+bias) -> bf16 -> A operand
__device__ __forceinline__ void hidden_epilogue32(uint32_t tmem, unsigned char* A, int q, int s, int row, const float* bias) {
    float v[32];
    const int col0 = 32 * s;
    tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + col0, v);
    tmem_wait_ld();
    const int kb = s >> 1, cbase = 4 * (s & 1);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float x = v[8 * cc + e];
            if (bias) x += bias[col0 + 8 * cc + e];
            t[e] = tanh_approx(x);
        }
        uint4 pk = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + (((cbase + cc) ^ (row & 7)) << 4)) = pk;
    }
}

__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// layer-1 operand row with a static layout: [hi(z) (L) | lo(z) (L) | y hi,lo ... | 1 | 0 ...]
template <int L>
__device__ __forceinline__ void write_a1_static(const Dims& d, unsigned char* A, int row, const float* z, const float* yrow, bool valid) {
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
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = 0.f;
        if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (i < d.y_dim) { const float hi = bf16_hi(yrow[i]); e[2 * i] = hi; e[2 * i + 1] = yrow[i] - hi; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i == 2 * d.y_dim) e[i] = 1.f;
        }
        constexpr int c = 2 * CH;
        const int kb = c >> 3, cc = c & 7;
        uint4 pk = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = pk;
    }
    const int K1c = 8 * d.nkb1;                // remaining chunks of the K blocks in use are zero
    for (int c = 2 * CH + 1; c < K1c; ++c) {
        const int kb = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(A + kb * 16384 + row * 128 + ((cc ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int L>
__global__ void __launch_bounds__(MH2_THREADS, 1) mh2_kernel(Mh2Params p) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bars[5];                       // layer 1/2 ready, chunk ready x2, buffer free x2
    __shared__ uint32_t tmem_slot;
    __shared__ int dead_flag;

    const Dims& d = p.d;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* A = base + ((d.image_bytes + 1023) & ~1023);
    float* red = reinterpret_cast<float*>(A + A_BYTES);          // [3][128] partial l + [128] uniforms
    float* uS = red + 3 * TM;
    float* epsS = reinterpret_cast<float*>(A + 16384);          // [128][L] draws of the next iteration
    const uint32_t bar12 = smem_u32(&bars[0]);
    const uint32_t bar3[2] = {smem_u32(&bars[1]), smem_u32(&bars[2])};
    const uint32_t barf[2] = {smem_u32(&bars[3]), smem_u32(&bars[4])};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, s = (warp >> 2) & 3;
    const int row = 32 * q + lane;
    const bool epi = warp < 16, owner = warp < 4, ctrl = (warp == 16 && lane == 0);
    uint32_t ph12 = 0, ph3[2] = {0, 0}, phf[2] = {0, 0};

    {
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(base);
        for (int i = threadIdx.x; i < d.image_bytes / 16; i += MH2_THREADS) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x == 0) {
        dead_flag = 0;
        mbar_init(bar12, 1);
        mbar_init(bar3[0], 1);
        mbar_init(bar3[1], 1);
        mbar_init(barf[0], MH2_EPI);
        mbar_init(barf[1], MH2_EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
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
    const int n_iter = p.n_burn + p.n_keep;
    const int64_t n_tiles = (p.rows + TM - 1) / TM;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row_g = tile * TM + row;
        const bool valid = epi && row_g < p.rows;
        const int64_t fr = valid ? row_g / p.C : 0;
        const float g_row = valid ? p.g[fr] : 1.f;
        uint32_t utt = 0, fc = 0;
        if (valid && !p.inj_eps) {
            utt = (uint32_t)p.frame_gid[fr];
            fc = (uint32_t)p.frame_idx[fr] | ((uint32_t)(row_g - fr * p.C) << 20);
        }
        float z[L], zp[L], yrow[3] = {0.f, 0.f, 0.f};
        float ll_cur = 0.f, u_cur = 0.5f;
        uint32_t n_acc = 0;
        if (owner) {
#pragma unroll
            for (int l = 0; l < L; ++l) z[l] = valid ? p.Z[row_g * L + l] : 0.f;
            if (valid)
                for (int i = 0; i < d.y_dim; ++i) yrow[i] = p.y[fr * d.y_dim + i];
        }

        // eval -1 scores the start state; eval it >= 0 scores proposal `it`
        for (int it = -1; it < n_iter; ++it) {
            if (owner) {
                if (it >= 0) {
                    u_cur = uS[row];
#pragma unroll
                    for (int l = 0; l < L; ++l) zp[l] = __fadd_rn(z[l], __fmul_rn(p.sd, epsS[row * L + l]));
                    write_a1_static<L>(d, A, row, zp, yrow, valid);
                } else {
                    write_a1_static<L>(d, A, row, z, yrow, valid);
                }
            }
            fence_async_smem();
            __syncthreads();                                                    // S1: layer-1 operand ready

            if (ctrl) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w1_addr, 16384, d.nkb1, tmem, HID);
                umma_commit(bar12);
            }
            if (epi) {
                mbar_wait(bar12, ph12, dead, p.status);
                tc_fence_after();
                hidden_epilogue32(tmem, A, q, s, row, nullptr);
                fence_async_smem();
                tc_fence_before();
            }
            ph12 ^= 1;
            __syncthreads();                                                    // S2
            if (d.n_hidden == 2) {
                if (ctrl) {
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w2_addr, 16384, 2, tmem, HID);
                    umma_commit(bar12);
                }
                if (epi) {
                    mbar_wait(bar12, ph12, dead, p.status);
                    tc_fence_after();
                    hidden_epilogue32(tmem, A, q, s, row, b2);
                    fence_async_smem();
                    tc_fence_before();
                }
                ph12 ^= 1;
                __syncthreads();                                                // S3
            }

            // ---- layer 3: 4 chunks of 128 bins + bin 512, TMEM double buffer, mbarrier hand-over
            float part = 0.f;
            if (ctrl) {
                tc_fence_after();
                issue_gemm2(a_addr, 16384, w3_addr, NPAD * 128, 2, tmem + 128, 128);
                umma_commit(bar3[0]);
                issue_gemm2(a_addr, 16384, w3_addr + 16384, NPAD * 128, 2, tmem + 256, 128);
                umma_commit(bar3[1]);
                for (int j = 2; j < 5; ++j) {
                    const int b = j & 1;
                    mbar_wait(barf[b], phf[b], dead, p.status);
                    phf[b] ^= 1;
                    tc_fence_after();
                    issue_gemm2(a_addr, 16384, w3_addr + j * 16384, NPAD * 128, 2, tmem + 128 + 128 * b, j < 4 ? 128 : 16);
                    umma_commit(bar3[b]);
                }
                mbar_wait(barf[1], phf[1], dead, p.status); phf[1] ^= 1;       // chunk 3 drained
                mbar_wait(barf[0], phf[0], dead, p.status); phf[0] ^= 1;       // chunk 4 drained
            }
            if (epi) {
                float acc = 0.f, accl = 0.f;
#pragma unroll 1
                for (int j = 0; j < 5; ++j) {
                    const int b = j & 1;
                    mbar_wait(bar3[b], ph3[b], dead, p.status);
                    ph3[b] ^= 1;
                    tc_fence_after();
                    const uint32_t tbuf = tmem + 128 + 128 * b + ((uint32_t)(32 * q) << 16);
                    if (j < 4) {
                        float v[32];
                        const int f0 = 128 * j + 32 * s;
                        const float4* Pq = p.Ppk + (tile * NQ + (f0 >> 2)) * TM + row;
                        const float4* Vq = p.Vbpk + (tile * NQ + (f0 >> 2)) * TM + row;
                        float4 pp[8], vb[8];
#pragma unroll
                        for (int qd = 0; qd < 8; ++qd) { pp[qd] = __ldg(Pq + qd * TM); vb[qd] = __ldg(Vq + qd * TM); }
                        tmem_ld32(tbuf + 32 * s, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int qd = 0; qd < 8; ++qd) {
                            const float4 bb = *reinterpret_cast<const float4*>(b3 + f0 + 4 * qd);
                            const float v0 = fmaf(g_row, ex2_approx(v[4 * qd + 0] + bb.x), vb[qd].x);
                            const float v1 = fmaf(g_row, ex2_approx(v[4 * qd + 1] + bb.y), vb[qd].y);
                            const float v2 = fmaf(g_row, ex2_approx(v[4 * qd + 2] + bb.z), vb[qd].z);
                            const float v3 = fmaf(g_row, ex2_approx(v[4 * qd + 3] + bb.w), vb[qd].w);
                            const float p01 = v0 * v1, p23 = v2 * v3;
                            const float n01 = fmaf(pp[qd].y, v0, pp[qd].x * v1), n23 = fmaf(pp[qd].w, v2, pp[qd].z * v3);
                            acc = fmaf(n01, rcp_approx(p01), acc);
                            acc = fmaf(n23, rcp_approx(p23), acc);
                            accl += lg2_approx(p01) + lg2_approx(p23);
                        }
                    } else if (s == 0) {                                   // bin 512
                        float v[4];
                        tmem_ld4(tbuf, v);
                        tmem_wait_ld();
                        const float4 pp = __ldg(p.Ppk + (tile * NQ + 128) * TM + row);
                        const float4 vb = __ldg(p.Vbpk + (tile * NQ + 128) * TM + row);
                        const float v0 = fmaf(g_row, ex2_approx(v[0] + b3[512]), vb.x);
                        acc = fmaf(pp.x, rcp_approx(v0), acc);
                        accl += lg2_approx(v0);
                    }
                    tc_fence_before();
                    mbar_arrive2(barf[b]);
                }
                part = fmaf(kLn2, accl, acc);
                if (s > 0) red[(s - 1) * TM + row] = part;

                // ---- draws of the next iteration (A is idle: every layer-3 MMA of this eval has completed)
                const int nxt = it + 1;
                if (nxt < n_iter) {
                    if (p.inj_eps) {
                        if (valid) {
                            const float* e = p.inj_eps + ((int64_t)nxt * p.rows + row_g) * L;
                            for (int l = s; l < L; l += 4) epsS[row * L + l] = e[l];
                            if (s == 0) uS[row] = p.inj_u[(int64_t)nxt * p.rows + row_g];
                        }
                    } else {
                        for (int blk = s; blk < L / 4; blk += 4) {
                            const Philox4 r = philox4x32_10(utt, fc, p.iter0 + (uint32_t)nxt, (uint32_t)blk, p.seed_lo, p.seed_hi);
                            float n0, n1, n2, n3;
                            box_muller(r.x, r.y, n0, n1);
                            box_muller(r.z, r.w, n2, n3);
                            *reinterpret_cast<float4*>(epsS + row * L + 4 * blk) = make_float4(n0, n1, n2, n3);
                        }
                        if (s == 3) {
                            const Philox4 r = philox4x32_10(utt, fc, p.iter0 + (uint32_t)nxt, (uint32_t)(L / 4), p.seed_lo, p.seed_hi);
                            uS[row] = u01(r.x);
                        }
                    }
                }
            }
            __syncthreads();                                                    // S4: partials and draws visible

            if (owner) {
                const float ll_prop = part + red[row] + red[TM + row] + red[2 * TM + row];
                if (it < 0) {
                    ll_cur = ll_prop;
                } else if (valid) {
                    float prior = 0.f;
#pragma unroll
                    for (int l = 0; l < L; ++l) prior += __fsub_rn(__fmul_rn(z[l], z[l]), __fmul_rn(zp[l], zp[l]));
                    const float a = (ll_cur - ll_prop) + 0.5f * prior;
                    if (p.a_trace) p.a_trace[(int64_t)it * p.rows + row_g] = a;
                    if (__logf(u_cur) < a) {
#pragma unroll
                        for (int l = 0; l < L; ++l) z[l] = zp[l];
                        ll_cur = ll_prop;
                        ++n_acc;
                    }
                    if (it >= p.n_burn) {
                        float4* dst = reinterpret_cast<float4*>(p.Zs + (row_g * p.n_keep + (it - p.n_burn)) * L);
#pragma unroll
                        for (int l = 0; l < L / 4; ++l) dst[l] = make_float4(z[4 * l], z[4 * l + 1], z[4 * l + 2], z[4 * l + 3]);
                    }
                }
            }
        }
        if (owner && valid) {
#pragma unroll
            for (int l = 0; l < L; ++l) p.Z[row_g * L + l] = z[l];
            if (p.n_accept) p.n_accept[row_g] += n_acc;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace tc
}  // namespace dvae

using namespace dvae;
using namespace dvae::tc;

extern "C" int dvae_mh_chain_tc2(const DvaeMlp* dec, const void* image, const float* Ppk, const float* Vbpk, const float* g,
                                 const float* y, int y_dim, const int32_t* frame_utt, const int32_t* frame_idx, float* Z,
                                 float* Zs, int64_t NT, int L, int n_chains, int n_burn, int n_keep, float var_rw,
                                 const DvaeRng* rng, uint32_t* n_accept, float* a_trace, int* status, void* stream) {
    Mh2Params p{};
    int rc = check_dims(dec, L, y_dim, "dvae_mh_chain_tc2", &p.d);
    if (rc) return rc;
    DVAE_REQUIRE(L == 16 || L == 32, "dvae_mh_chain_tc2: latent size must be 16 or 32 (got %d)", L);
    DVAE_REQUIRE(y_dim <= 3, "dvae_mh_chain_tc2: at most 3 label inputs");
    DVAE_REQUIRE(image && Ppk && Vbpk && g && Z && Zs && rng && status, "dvae_mh_chain_tc2: null pointer");
    DVAE_REQUIRE(y_dim == 0 || y, "dvae_mh_chain_tc2: y_dim=%d but y is null", y_dim);
    DVAE_REQUIRE(NT >= 0 && n_chains >= 1 && n_chains < 4096 && n_burn >= 0 && n_keep >= 1 && var_rw > 0.f, "dvae_mh_chain_tc2: bad sizes");
    DVAE_REQUIRE((rng->eps == nullptr) == (rng->u == nullptr), "dvae_mh_chain_tc2: eps and u must be injected together");
    DVAE_REQUIRE(rng->eps || (frame_utt && frame_idx), "dvae_mh_chain_tc2: Philox mode needs frame_utt/frame_idx");
    DVAE_REQUIRE((reinterpret_cast<uintptr_t>(Zs) & 15) == 0, "dvae_mh_chain_tc2: Zs must be 16-byte aligned");
    if (NT == 0) return 0;
    p.image = (const unsigned char*)image;
    p.rows = NT * n_chains; p.C = n_chains; p.y = y;
    p.Ppk = (const float4*)Ppk; p.Vbpk = (const float4*)Vbpk; p.g = g; p.Z = Z; p.Zs = Zs;
    p.frame_gid = frame_utt; p.frame_idx = frame_idx; p.inj_eps = rng->eps; p.inj_u = rng->u;
    p.n_accept = n_accept; p.a_trace = a_trace; p.n_burn = n_burn; p.n_keep = n_keep;
    p.seed_lo = (uint32_t)(rng->seed & 0xffffffffu); p.seed_hi = (uint32_t)(rng->seed >> 32); p.iter0 = rng->iter0;
    p.sd = sqrtf(var_rw);
    p.status = status;
    const size_t smem = smem_bytes(p.d) + 2048;
    DVAE_REQUIRE(smem <= 227 * 1024, "dvae_mh_chain_tc2: shared memory budget exceeded");
    const int64_t n_tiles = (p.rows + TM - 1) / TM;
    const int grid = (int)(n_tiles < 148 ? n_tiles : 148);
    cudaStream_t st = (cudaStream_t)stream;
    if (L == 16) {
        cudaFuncSetAttribute(mh2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mh2_kernel<16><<<grid, MH2_THREADS, smem, st>>>(p);
    } else {
        cudaFuncSetAttribute(mh2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mh2_kernel<32><<<grid, MH2_THREADS, smem, st>>>(p);
    }
    return check_launch("mh2_kernel");
}
