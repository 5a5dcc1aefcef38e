// MeshConv forward on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), with fp32
// accuracy recovered by the 3xTF32 split  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (a_hi = rna_tf32(a), a_lo = a - a_hi).
// Used where the contraction is a real dense GEMM: Cin % 32 == 0 and Cout % 16 == 0 (K = 5*Cin >= 160).
//
// One CTA = 128 edges x N_TILE output channels.  The A operand cannot come from TMA -- it is *gathered* (self + 4
// one-ring neighbours) and pushed through the symmetric functions -- so 8 producer warps build it: coalesced
// 16-byte row-segment loads -> G values -> hi/lo split -> st.shared into the canonical K-major SWIZZLE_128B tile,
// fence.proxy.async, mbarrier arrive.  One elected thread of warp 8 issues the MMAs and commits each stage back to
// the producers; after the last K block the 8 producer warps become the epilogue (tcgen05.ld -> +bias -> global).
//
// K is permuted so that every 32-wide K block needs one gather pass (the weight is re-laid to match, once per call):
//   for each block of 32 input channels cb:   blk 0: G0[c 0..31]
//                                              blk 1: G1[c 0..15]  | G3[c 0..15]      (both need f1, f3)
//                                              blk 2: G1[c 16..31] | G3[c 16..31]
//                                              blk 3: G2[c 0..15]  | G4[c 0..15]      (both need f2, f4)
//                                              blk 4: G2[c 16..31] | G4[c 16..31]
#include "common.cuh"
#include "tc_common.cuh"

namespace mcnn {

constexpr int TC_M = 128;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1) * 32;

// weight [Cout][Cin][5] -> wt[Cout][5*Cin] in the permuted K order above
__global__ void weight_tc_layout_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cin, int Cout) {
    const int K = 5 * Cin;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Cout * K) return;
    const int o = idx / K, kk = idx - o * K;
    const int kb = kk >> 5, p = kk & 31;
    const int cb = kb / 5, blk = kb - cb * 5;
    int k5, c;
    if (blk == 0) { k5 = 0; c = cb * 32 + p; }
    else {
        const int first = (blk <= 2) ? 1 : 2;
        k5 = (p < 16) ? first : first + 2;
        c = cb * 32 + (((blk - 1) & 1) << 4) + (p & 15);
    }
    wt[idx] = w[((size_t)o * Cin + c) * 5 + k5];
}

template <int N_TILE>
struct TcSmem {
    static constexpr int A_BYTES = TC_M * 128;            // 128 rows x 32 fp32
    static constexpr int B_BYTES = N_TILE * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (N_TILE >= 256) ? 2 : ((N_TILE >= 128) ? 3 : 4);
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int NB_OFF = TILE_BYTES;             // int nb[128][5]
    static constexpr int BAR_OFF = NB_OFF + TC_M * 5 * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;   // + barriers + tmem slot + alignment slack
};

template <int N_TILE>
__global__ void __launch_bounds__(TC_THREADS, 1)
meshconv_fwd_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                       const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                       int B, int E, int Cin, int Cout) {
    using S = TcSmem<N_TILE>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int (*nb)[5] = reinterpret_cast<int (*)[5]>(smem + S::NB_OFF);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* accum_bar = empty_bar + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int m0 = blockIdx.x * TC_M, n0 = blockIdx.y * N_TILE;
    const int K = 5 * Cin;
    const int num_kb = K >> 5;

    // ---- setup ------------------------------------------------------------------------------------------------
    for (int i = tid; i < TC_M * 5; i += TC_THREADS) {
        const int r = i / 5, k = i - r * 5;
        const int m = m0 + r;
        int idx = -1;
        if (m < M) {
            const int b = m / E, e = m - b * E;
            if (e >= __ldg(ec + b)) idx = b * E;                                  // padded slot -> edge 0
            else if (k == 0) idx = m;
            else { const int g = __ldg(gemm + (size_t)m * 4 + (k - 1)); idx = g < 0 ? -1 : b * E + g; }
        }
        nb[r][k] = idx;
    }
    if (warp == TC_PRODUCER_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, TC_PRODUCER_WARPS * 32); tc::mbar_init(empty_bar + s, 1); }
            tc::mbar_init(accum_bar, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<N_TILE>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        // ======================================== producers ========================================================
        const int pt = tid;                                     // 0..255
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            tc::mbar_wait(empty_bar + s, ph ^ 1);
            unsigned char* a_hi = smem + s * S::STAGE_BYTES;
            unsigned char* a_lo = a_hi + S::A_BYTES;
            unsigned char* b_hi = a_lo + S::A_BYTES;
            unsigned char* b_lo = b_hi + S::B_BYTES;
            const int cb = kb / 5, blk = kb - cb * 5;
            // ---- B tile: rows n0..n0+N_TILE of wt, K block kb (32 floats = 8 chunks) ------------------------------
#pragma unroll
            for (int it = 0; it < (N_TILE * 8) / (TC_PRODUCER_WARPS * 32); ++it) {
                const int item = pt + it * (TC_PRODUCER_WARPS * 32);
                const int r = item >> 3, q = item & 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + r < Cout) v = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(n0 + r) * K + (kb << 5) + (q << 2)));
                float4 hi, lo;
                tc::split4(v, hi, lo);
                const uint32_t off = tc::sw128_off(r, q);
                *reinterpret_cast<float4*>(b_hi + off) = hi;
                *reinterpret_cast<float4*>(b_lo + off) = lo;
            }
            // ---- A tile ---------------------------------------------------------------------------------------------
            if (blk == 0) {
                const int c0 = cb * 32;
#pragma unroll
                for (int it = 0; it < (TC_M * 8) / (TC_PRODUCER_WARPS * 32); ++it) {
                    const int item = pt + it * (TC_PRODUCER_WARPS * 32);
                    const int r = item >> 3, q = item & 7;
                    const int src = nb[r][0];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)src * Cin + c0 + (q << 2)));
                    float4 hi, lo;
                    tc::split4(v, hi, lo);
                    const uint32_t off = tc::sw128_off(r, q);
                    *reinterpret_cast<float4*>(a_hi + off) = hi;
                    *reinterpret_cast<float4*>(a_lo + off) = lo;
                }
            } else {
                const int ka = (blk <= 2) ? 1 : 2;                         // neighbour slots (ka, ka+2)
                const int c0 = cb * 32 + (((blk - 1) & 1) << 4);
#pragma unroll
                for (int it = 0; it < (TC_M * 4) / (TC_PRODUCER_WARPS * 32); ++it) {
                    const int item = pt + it * (TC_PRODUCER_WARPS * 32);
                    const int r = item >> 2, q = item & 3;
                    const int sa = nb[r][ka], sb = nb[r][ka + 2];
                    float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
                    if (sa >= 0) fa = __ldg(reinterpret_cast<const float4*>(x + (size_t)sa * Cin + c0 + (q << 2)));
                    if (sb >= 0) fb = __ldg(reinterpret_cast<const float4*>(x + (size_t)sb * Cin + c0 + (q << 2)));
                    const float4 sum = make_float4(fa.x + fb.x, fa.y + fb.y, fa.z + fb.z, fa.w + fb.w);
                    const float4 dif = make_float4(fabsf(fa.x - fb.x), fabsf(fa.y - fb.y), fabsf(fa.z - fb.z), fabsf(fa.w - fb.w));
                    float4 hi, lo;
                    tc::split4(sum, hi, lo);
                    uint32_t off = tc::sw128_off(r, q);
                    *reinterpret_cast<float4*>(a_hi + off) = hi;
                    *reinterpret_cast<float4*>(a_lo + off) = lo;
                    tc::split4(dif, hi, lo);
                    off = tc::sw128_off(r, q + 4);
                    *reinterpret_cast<float4*>(a_hi + off) = hi;
                    *reinterpret_cast<float4*>(a_lo + off) = lo;
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(full_bar + s);
        }
        // ======================================== epilogue ==========================================================
        tc::mbar_wait(accum_bar, 0);
        tc::tc_fence_after();
        const int lane_grp = warp & 3;                           // TMEM lanes 32*lane_grp .. +31
        const int r = lane_grp * 32 + lane;
        const int m = m0 + r;
        constexpr int COLS_PER_WARP = N_TILE / 2;                // warps w and w+4 share a lane group
        const int col0 = (warp >> 2) * COLS_PER_WARP;
#pragma unroll 1
        for (int cc = 0; cc < COLS_PER_WARP; cc += 32) {
            float v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(col0 + cc), v);
            if (m < M) {
                const int n = n0 + col0 + cc;
                float* o = out + (size_t)m * Cout + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (n + j < Cout) {
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(bias + n + j));
                        *reinterpret_cast<float4*>(o + j) = make_float4(v[j] + bv.x, v[j + 1] + bv.y, v[j + 2] + bv.z, v[j + 3] + bv.w);
                    }
                }
            }
        }
        tc::tc_fence_before();
    } else {
        // ======================================== MMA issuer ========================================================
        constexpr uint32_t idesc = tc::make_idesc_tf32(N_TILE);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            tc::mbar_wait(full_bar + s, ph);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t a_lo = a_hi + S::A_BYTES;
                const uint32_t b_hi = a_lo + S::A_BYTES;
                const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {                 // 4 x (K = 8 tf32 = 32 bytes)
                    const uint64_t dah = tc::make_desc_sw128(a_hi + kk * 32), dal = tc::make_desc_sw128(a_lo + kk * 32);
                    const uint64_t dbh = tc::make_desc_sw128(b_hi + kk * 32), dbl = tc::make_desc_sw128(b_lo + kk * 32);
                    tc::mma_tf32(tmem_base, dah, dbh, idesc, (kb | kk) != 0);
                    tc::mma_tf32(tmem_base, dah, dbl, idesc, 1);
                    tc::mma_tf32(tmem_base, dal, dbh, idesc, 1);
                }
                tc::mma_commit(empty_bar + s);                   // frees the stage when these MMAs have read it
                if (kb == num_kb - 1) tc::mma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS) {
        tc::tc_fence_after();
        tc::tmem_dealloc<N_TILE>(tmem_base);
    }
}

template <int N_TILE>
static int launch_fwd_tc(const float* x, const int32_t* gemm, const int32_t* ec, const float* wt, const float* bias,
                         float* out, int B, int E, int Cin, int Cout, cudaStream_t st) {
    using S = TcSmem<N_TILE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(meshconv_fwd_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) { g_last_error = e; return MCNN_E_CUDA; }
        configured = true;
    }
    dim3 grid(ceil_div(B * E, TC_M), ceil_div(Cout, N_TILE));
    meshconv_fwd_tc_kernel<N_TILE><<<grid, TC_THREADS, S::TOTAL, st>>>(x, gemm, ec, wt, bias, out, B, E, Cin, Cout);
    return after_launch();
}

}  // namespace mcnn

using namespace mcnn;

extern "C" int mcnn_meshconv_tc_supported(int Cin, int Cout) { return (Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 64 == 0) ? 1 : 0; }

extern "C" int mcnn_meshconv_fwd_tc(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                    const float* bias, float* out, float* wt_ws, int B, int E, int Cin, int Cout,
                                    void* stream) {
    if (!x || !gemm || !ec || !weight || !out || !wt_ws || B <= 0 || E <= 0) return MCNN_E_BADARG;
    if (!mcnn_meshconv_tc_supported(Cin, Cout)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = Cout * 5 * Cin;
    weight_tc_layout_kernel<<<ceil_div(total, 256), 256, 0, st>>>(weight, wt_ws, Cin, Cout);
    int rc = after_launch();
    if (rc) return rc;
    const int mtiles = ceil_div(B * E, TC_M);
    if (Cout % 256 == 0 && mtiles >= 222) return launch_fwd_tc<256>(x, gemm, ec, wt_ws, bias, out, B, E, Cin, Cout, st);
    if (Cout % 128 == 0) return launch_fwd_tc<128>(x, gemm, ec, wt_ws, bias, out, B, E, Cin, Cout, st);
    return launch_fwd_tc<64>(x, gemm, ec, wt_ws, bias, out, B, E, Cin, Cout, st);
}
