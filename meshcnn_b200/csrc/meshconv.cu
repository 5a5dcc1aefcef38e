// MeshConv: fused 1-ring gather + symmetric functions + (5*Cin) x Cout contraction, forward and backward.
// Exact-fp32 SIMT path (FFMA); every layer shape goes through it, the tensor-core path (meshconv_tc.cu) takes
// over only where the contraction is a real dense GEMM.
//
// Reference semantics: models/layers/mesh_conv.py:20-84 (forward), autograd of the same (backward); restated
// in SURVEY.md Appendix A.  Layout: x[B][E][Cin], out[B][E][Cout] (edge-major), gemm[B][E][4] int32, ec[B].
#include "common.cuh"

namespace mcnn {

constexpr int TM = 64;          // edges per CTA tile
constexpr int TN = 64;          // output channels per CTA tile (forward)
constexpr int CK = 16;          // input channels per K chunk
constexpr int KK = CK * 5;      // K extent of one chunk (channel-major: kk = c_local * 5 + k)
constexpr int LDT = TM + 4;     // padded leading dimension of the [kk][m] tiles
constexpr int NTHREADS = 256;

// Resolve the five gather rows of TM consecutive edges into flat row indices (b*E + edge) or -1.
//   e <  ec[b] : slot 0 = e itself, slots 1..4 = gemm[b][e][0..3] (-1 stays -1: zero feature)
//   e >= ec[b] : all five slots = edge 0 of mesh b  (pad value 0 before the +1 shift)
__device__ __forceinline__ void resolve_rows(int (*nb)[5], int m0, int M, int E, const int32_t* __restrict__ gemm,
                                             const int32_t* __restrict__ ec) {
    for (int i = threadIdx.x; i < TM * 5; i += NTHREADS) {
        int r = i / 5, k = i - r * 5;
        int m = m0 + r;
        int idx = -1;
        if (m < M) {
            int b = m / E, e = m - b * E;
            if (e >= ec[b]) {
                idx = b * E;
            } else if (k == 0) {
                idx = m;
            } else {
                int g = gemm[(size_t)m * 4 + (k - 1)];
                idx = (g < 0) ? -1 : b * E + g;
            }
        }
        nb[r][k] = idx;
    }
}

__device__ __forceinline__ float4 load4(const float* __restrict__ base, int row, int C, int c, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < 0) return v;
    const float* p = base + (size_t)row * C + c;
    if (vec) {
        if (c < C) v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
        if (c + 0 < C) v.x = __ldg(p + 0);
        if (c + 1 < C) v.y = __ldg(p + 1);
        if (c + 2 < C) v.z = __ldg(p + 2);
        if (c + 3 < C) v.w = __ldg(p + 3);
    }
    return v;
}

// G tile for rows of this CTA and channels [c0, c0+CK): Gs[(cl*5+k) * ld + r]  (kmajor=true)
//                                                   or  Gs[r * ld + cl*5+k]    (kmajor=false)
template <bool KMAJOR>
__device__ __forceinline__ void build_g_tile(float* Gs, int ld, const int (*nb)[5], const float* __restrict__ x,
                                             int Cin, int c0, bool vec) {
    int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    int c = c0 + q * 4;
    float4 f0 = load4(x, nb[r][0], Cin, c, vec);
    float4 f1 = load4(x, nb[r][1], Cin, c, vec);
    float4 f2 = load4(x, nb[r][2], Cin, c, vec);
    float4 f3 = load4(x, nb[r][3], Cin, c, vec);
    float4 f4 = load4(x, nb[r][4], Cin, c, vec);
    const float a0[4] = {f0.x, f0.y, f0.z, f0.w};
    const float a1[4] = {f1.x + f3.x, f1.y + f3.y, f1.z + f3.z, f1.w + f3.w};
    const float a2[4] = {f2.x + f4.x, f2.y + f4.y, f2.z + f4.z, f2.w + f4.w};
    const float a3[4] = {fabsf(f1.x - f3.x), fabsf(f1.y - f3.y), fabsf(f1.z - f3.z), fabsf(f1.w - f3.w)};
    const float a4[4] = {fabsf(f2.x - f4.x), fabsf(f2.y - f4.y), fabsf(f2.z - f4.z), fabsf(f2.w - f4.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int kk = (q * 4 + j) * 5;
        if (KMAJOR) {
            Gs[(kk + 0) * ld + r] = a0[j];
            Gs[(kk + 1) * ld + r] = a1[j];
            Gs[(kk + 2) * ld + r] = a2[j];
            Gs[(kk + 3) * ld + r] = a3[j];
            Gs[(kk + 4) * ld + r] = a4[j];
        } else {
            float* g = Gs + r * ld + kk;
            g[0] = a0[j]; g[1] = a1[j]; g[2] = a2[j]; g[3] = a3[j]; g[4] = a4[j];
        }
    }
}

// weight [Cout][Cin*5] -> wt [Cin*5][Cout]
__global__ void weight_kmajor_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int Cout) {
    pdl_enter();
    __shared__ float t[32][33];
    int k0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int o = o0 + i, k = k0 + threadIdx.x;
        t[i][threadIdx.x] = (o < Cout && k < K) ? w[(size_t)o * K + k] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int k = k0 + i, o = o0 + threadIdx.x;
        if (k < K && o < Cout) wt[(size_t)k * Cout + o] = t[threadIdx.x][i];
    }
}

__global__ void __launch_bounds__(NTHREADS)
meshconv_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                    const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                    int B, int E, int Cin, int Cout) {
    pdl_enter();
    __shared__ int nb[TM][5];
    __shared__ __align__(16) float As[KK * LDT];
    __shared__ __align__(16) float Ws[KK * (TN + 4)];
    const int M = B * E;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const bool vec = (Cin & 3) == 0;
    const bool nvec = (Cout & 3) == 0;
    resolve_rows(nb, m0, M, E, gemm, ec);
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < Cin; c0 += CK) {
        build_g_tile<true>(As, LDT, nb, x, Cin, c0, vec);
        const int kmax = min(CK, Cin - c0) * 5;
        // W tile: rows kk of wt (contiguous in o)
        for (int i = threadIdx.x; i < KK * (TN / 4); i += NTHREADS) {
            int kk = i / (TN / 4), nq = i - kk * (TN / 4);
            int n = n0 + nq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kk < kmax) {
                const float* p = wt + (size_t)(c0 * 5 + kk) * Cout + n;
                if (nvec) {
                    if (n < Cout) v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (n + 0 < Cout) v.x = __ldg(p + 0);
                    if (n + 1 < Cout) v.y = __ldg(p + 1);
                    if (n + 2 < Cout) v.z = __ldg(p + 2);
                    if (n + 3 < Cout) v.w = __ldg(p + 3);
                }
            }
            *reinterpret_cast<float4*>(&Ws[kk * (TN + 4) + nq * 4]) = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < kmax; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk * LDT + ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Ws[kk * (TN + 4) + tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int n = n0 + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n + j < Cout) bv[j] = __ldg(bias + n + j);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float* o = out + (size_t)m * Cout + n;
        if (nvec && n + 3 < Cout) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < Cout) o[j] = acc[i][j] + bv[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, data:  dG = dout * W  (per edge, [Cout] -> [Cin][5]), sign logic, scatter-add to the gather rows
// ---------------------------------------------------------------------------------------------------------
constexpr int OK = 32;          // output channels per K chunk in the backward kernels

__global__ void __launch_bounds__(NTHREADS)
meshconv_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ gemm,
                         const int32_t* __restrict__ ec, const float* __restrict__ w, float* __restrict__ dx,
                         int B, int E, int Cin, int Cout) {
    pdl_enter();
    __shared__ int nb[TM][5];
    __shared__ __align__(16) float As[OK * LDT];          // dout tile, [oo][r]
    __shared__ float Ws[OK * (KK + 1)];                    // weight tile, [oo][cl*5+k]
    const int M = B * E;
    const int m0 = blockIdx.x * TM, c0 = blockIdx.y * CK;
    const bool ovec = (Cout & 3) == 0;
    resolve_rows(nb, m0, M, E, gemm, ec);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int ncl = min(CK, Cin - c0);
    float acc[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[i][k] = 0.f;

    for (int o0 = 0; o0 < Cout; o0 += OK) {
        {   // dout tile: 64 rows x 32 channels, one float4 per thread for 2 passes
            for (int i = threadIdx.x; i < TM * (OK / 4); i += NTHREADS) {
                int r = i / (OK / 4), oq = i - r * (OK / 4);
                int m = m0 + r, o = o0 + oq * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M) {
                    const float* p = dout + (size_t)m * Cout + o;
                    if (ovec) {
                        if (o < Cout) v = __ldg(reinterpret_cast<const float4*>(p));
                    } else {
                        if (o + 0 < Cout) v.x = __ldg(p + 0);
                        if (o + 1 < Cout) v.y = __ldg(p + 1);
                        if (o + 2 < Cout) v.z = __ldg(p + 2);
                        if (o + 3 < Cout) v.w = __ldg(p + 3);
                    }
                }
                As[(oq * 4 + 0) * LDT + r] = v.x;
                As[(oq * 4 + 1) * LDT + r] = v.y;
                As[(oq * 4 + 2) * LDT + r] = v.z;
                As[(oq * 4 + 3) * LDT + r] = v.w;
            }
            for (int i = threadIdx.x; i < OK * KK; i += NTHREADS) {
                int oo = i / KK, kk = i - oo * KK;
                int o = o0 + oo;
                float v = 0.f;
                if (o < Cout && kk < ncl * 5) v = __ldg(w + (size_t)o * Cin * 5 + (size_t)c0 * 5 + kk);
                Ws[oo * (KK + 1) + kk] = v;
            }
        }
        __syncthreads();
        const int omax = min(OK, Cout - o0);
#pragma unroll 4
        for (int oo = 0; oo < omax; ++oo) {
            float4 a = *reinterpret_cast<const float4*>(&As[oo * LDT + ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            float wv[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) wv[k] = Ws[oo * (KK + 1) + tx * 5 + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 5; ++k) acc[i][k] = fmaf(av[i], wv[k], acc[i][k]);
        }
        __syncthreads();
    }
    const int c = c0 + tx;
    if (c >= Cin) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        int m = m0 + r;
        if (m >= M) continue;
        int b = m / E, e = m - b * E;
        const float g0 = acc[i][0], g1 = acc[i][1], g2 = acc[i][2], g3 = acc[i][3], g4 = acc[i][4];
        if (e >= __ldg(ec + b)) {
            // padded slot: all five gathers are edge 0, |f1-f3| = |f2-f4| = 0 with sign(0) = 0
            atomicAdd(dx + (size_t)(b * E) * Cin + c, g0 + 2.f * g1 + 2.f * g2);
            continue;
        }
        const int n1 = nb[r][1], n2 = nb[r][2], n3 = nb[r][3], n4 = nb[r][4];
        const float f1 = n1 >= 0 ? __ldg(x + (size_t)n1 * Cin + c) : 0.f;
        const float f2 = n2 >= 0 ? __ldg(x + (size_t)n2 * Cin + c) : 0.f;
        const float f3 = n3 >= 0 ? __ldg(x + (size_t)n3 * Cin + c) : 0.f;
        const float f4 = n4 >= 0 ? __ldg(x + (size_t)n4 * Cin + c) : 0.f;
        const float s13 = sgn(f1 - f3), s24 = sgn(f2 - f4);
        atomicAdd(dx + (size_t)m * Cin + c, g0);
        if (n1 >= 0) atomicAdd(dx + (size_t)n1 * Cin + c, g1 + s13 * g3);
        if (n3 >= 0) atomicAdd(dx + (size_t)n3 * Cin + c, g1 - s13 * g3);
        if (n2 >= 0) atomicAdd(dx + (size_t)n2 * Cin + c, g2 + s24 * g4);
        if (n4 >= 0) atomicAdd(dx + (size_t)n4 * Cin + c, g2 - s24 * g4);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward, weight:  dW[o][c][k] = sum_m dout[m][o] * G_k[m][c]   (all E slots, padded ones included)
// ---------------------------------------------------------------------------------------------------------
constexpr int MSPLIT_MAX = 1024;    // most rows reduced by one CTA before the atomic flush

__global__ void __launch_bounds__(NTHREADS)
meshconv_bwd_weight_kernel(const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ gemm,
                           const int32_t* __restrict__ ec, float* __restrict__ dw, float* __restrict__ dbias,
                           int B, int E, int Cin, int Cout, int msplit) {
    pdl_enter();
    __shared__ int nb[TM][5];
    __shared__ __align__(16) float As[TM * (TN + 4)];     // dout tile [mm][oo]
    __shared__ float Gs[TM * (KK + 1)];                   // G tile    [mm][cl*5+k]
    const int M = B * E;
    const int o0 = blockIdx.y * TN, c0 = blockIdx.z * CK;
    const bool vec = (Cin & 3) == 0, ovec = (Cout & 3) == 0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][5];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[i][k] = 0.f;
    const int mbeg = blockIdx.x * msplit, mend = min(M, mbeg + msplit);
    for (int m0 = mbeg; m0 < mend; m0 += TM) {
        resolve_rows(nb, m0, mend, E, gemm, ec);           // rows >= mend resolve to -1 -> zero G
        for (int i = threadIdx.x; i < TM * (TN / 4); i += NTHREADS) {
            int r = i / (TN / 4), oq = i - r * (TN / 4);
            int m = m0 + r, o = o0 + oq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mend) {
                const float* p = dout + (size_t)m * Cout + o;
                if (ovec) {
                    if (o < Cout) v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (o + 0 < Cout) v.x = __ldg(p + 0);
                    if (o + 1 < Cout) v.y = __ldg(p + 1);
                    if (o + 2 < Cout) v.z = __ldg(p + 2);
                    if (o + 3 < Cout) v.w = __ldg(p + 3);
                }
            }
            *reinterpret_cast<float4*>(&As[r * (TN + 4) + oq * 4]) = v;
        }
        __syncthreads();
        build_g_tile<false>(Gs, KK + 1, nb, x, Cin, c0, vec);
        __syncthreads();
#pragma unroll 4
        for (int mm = 0; mm < TM; ++mm) {
            float4 a = *reinterpret_cast<const float4*>(&As[mm * (TN + 4) + ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            float gv[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) gv[k] = Gs[mm * (KK + 1) + tx * 5 + k];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += av[i];
#pragma unroll
                for (int k = 0; k < 5; ++k) acc[i][k] = fmaf(av[i], gv[k], acc[i][k]);
            }
        }
        __syncthreads();
    }
    const int c = c0 + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int o = o0 + ty * 4 + i;
        if (o >= Cout) continue;
        if (c < Cin) {
            float* p = dw + ((size_t)o * Cin + c) * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) atomicAdd(p + k, acc[i][k]);
        }
        if (dbias != nullptr && blockIdx.z == 0 && tx == 0) atomicAdd(dbias + o, bsum[i]);
    }
}

}  // namespace mcnn

using namespace mcnn;

extern "C" int mcnn_meshconv_fwd(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                 const float* bias, float* out, float* wt_ws, int B, int E, int Cin, int Cout,
                                 void* stream) {
    if (!x || !gemm || !ec || !weight || !out || !wt_ws || B <= 0 || E <= 0 || Cin <= 0 || Cout <= 0) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int K = Cin * 5;
    launch_k(weight_kmajor_kernel, dim3(ceil_div(K, 32), ceil_div(Cout, 32)), dim3(32, 8), 0, st, weight, wt_ws, K, Cout);
    int rc = after_launch();
    if (rc) return rc;
    dim3 grid(ceil_div(B * E, TM), ceil_div(Cout, TN));
    launch_k(meshconv_fwd_kernel, grid, NTHREADS, 0, st, x, gemm, ec, wt_ws, bias, out, B, E, Cin, Cout);
    return after_launch();
}

extern "C" int mcnn_meshconv_bwd_data(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                      const float* weight, float* dx, int B, int E, int Cin, int Cout, void* stream) {
    if (!dout || !x || !gemm || !ec || !weight || !dx || B <= 0 || E <= 0 || Cin <= 0 || Cout <= 0) return MCNN_E_BADARG;
    dim3 grid(ceil_div(B * E, TM), ceil_div(Cin, CK));
    launch_k(meshconv_bwd_data_kernel, grid, NTHREADS, 0, (cudaStream_t)stream, dout, x, gemm, ec, weight, dx, B, E, Cin, Cout);
    return after_launch();
}

extern "C" int mcnn_meshconv_bwd_weight(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                        float* dweight, float* dbias, int B, int E, int Cin, int Cout, void* stream) {
    if (!dout || !x || !gemm || !ec || !dweight || B <= 0 || E <= 0 || Cin <= 0 || Cout <= 0) return MCNN_E_BADARG;
    // rows per CTA: about two CTAs per SM in total (a narrow layer has a single (Cout, Cin) tile: 1024-row slabs left
    // two thirds of the SMs idle), at least one 64-row tile, at most 1024 rows
    const int tiles = ceil_div(Cout, TN) * ceil_div(Cin, CK);
    int msplit = ceil_div(ceil_div(B * E, max(1, 296 / tiles)), TM) * TM;
    msplit = min(MSPLIT_MAX, max(TM, msplit));
    dim3 grid(ceil_div(B * E, msplit), ceil_div(Cout, TN), ceil_div(Cin, CK));
    launch_k(meshconv_bwd_weight_kernel, grid, NTHREADS, 0, (cudaStream_t)stream, dout, x, gemm, ec, dweight, dbias, B, E, Cin, Cout, msplit);
    return after_launch();
}
