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

// debug knobs for bottleneck attribution (tools/tc_knobs.py): bit 0 = producers skip global loads, bit 1 = MMA warp skips
// the tcgen05.mma issue, bit 2 = producers skip the st.shared, bit 3 = no bulk copies of the weight tile.  0 in normal operation.
__device__ int g_tc_debug = 0;
__device__ long long g_tc_prof[16];        // -DMCNN_TC_PROF: cycle counters of CTA 0 (tools/tc_prof.py)
__device__ long long g_sk_prof[64];        // -DMCNN_TC_PROF: clock64 stamps of one CTA of the stream-K kernel (tools/sk_prof.py)
__device__ int g_sk_prof_cta = 0;
__device__ unsigned long long g_sk_cta_ns[256][2];   // -DMCNN_TC_PROF: %globaltimer at entry / exit of every CTA
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef MCNN_TC_PROF
#define SK_STAMP(cond, idx) do { if ((cond) && blockIdx.x == g_sk_prof_cta && (idx) < 64) g_sk_prof[(idx)] = clock64(); } while (0)
#else
#define SK_STAMP(cond, idx) do { } while (0)
#endif
static int g_tc_force_ntile = 0;          // host side: 0 = heuristic
static int g_tc_pw16 = 0;                 // host side: 16 producer warps in the forward kernel (N >= 128)
static int g_tc_pair = 0;                 // host side: use the CTA-pair (cta_group::2) forward kernel for N = 256

constexpr int TC_M = 128;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 1) * 32;

// Row / chunk of item i (0 .. 511) of a "pair" K block (two 16-channel halves per row: 4 + 4 chunks, written by two
// separate stores).  The shared-memory pipeline serves a 16-byte-per-lane store in quarter-warps of 8 lanes; with the
// 128-byte swizzle the 4 chunks of row r land in chunk slots {0..3} ^ (r & 7), so a quarter-warp is conflict-free iff its
// two rows differ in bit 2: rows (r, r + 4) instead of (r, r + 1).
__device__ __forceinline__ int pair_item_row(int i) { return ((i >> 5) << 3) + (((i >> 2) & 1) << 2) + ((i >> 3) & 3); }

// Position inside a 32-wide K block of its p-th logical entry (blk 0: input channel p of the block; pair blocks: p < 16 ->
// first tap, channel p; p >= 16 -> second tap, channel p - 16).  Permuted so that the producers of the persistent kernel
// can write the A operand straight into TENSOR MEMORY in the mma-fragment layout (tcgen05.st .16x256b: thread t holds K
// columns 8i + 2(t%4) + {0,1}): in a G0 block thread t owns channels 8(t%4) .. +7, in a pair block channels 4(t%4) .. +3 of
// both taps -- each a contiguous piece of a gathered row.  A and B use the same order, so the product is unchanged.
__host__ __device__ constexpr int fw_kpos(int blk, int p) {
    return blk == 0 ? (8 * ((p & 7) >> 1) + 2 * (p >> 3) + (p & 1))
                    : ((p & 16) + 8 * (((p & 15) & 3) >> 1) + 2 * ((p & 15) >> 2) + (p & 1));
}

// weight [Cout][Cin][5] -> the B operand exactly as the tensor core wants it in shared memory, split into TF32 head and
// remainder:  img[hi|lo][kb][Cout rows x 128 bytes, canonical K-major SWIZZLE_128B atoms], K in the permuted order above.
// The rows n0 .. n0+N_TILE of one K block are N_TILE*128 contiguous bytes, so a CTA fetches its B tile with one bulk copy
// per half instead of pushing it through registers.
__global__ void weight_tc_layout_kernel(const float* __restrict__ w, float* __restrict__ img, int Cin, int Cout) {
    pdl_enter();
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
    float hi, lo;
    tc::split_tf32(w[((size_t)o * Cin + c) * 5 + k5], hi, lo);
    const int pk = fw_kpos(blk, p);
    const size_t off = (size_t)kb * Cout * 32 + (tc::sw128_off(o, pk >> 2) >> 2) + (pk & 3);
    img[off] = hi;
    img[(size_t)Cout * K + off] = lo;
}

// four consecutive logical K entries (a float4 of channels) under fw_kpos: K columns c1, c1+1 and c1+8, c1+9 of row r
__device__ __forceinline__ void sts_kperm(uint32_t a_hi, uint32_t a_lo, int r, int c1, const float4& hi, const float4& lo) {
    const uint32_t o1 = tc::sw128_off(r, c1 >> 2) + (uint32_t)((c1 & 3) << 2);
    const uint32_t o2 = tc::sw128_off(r, (c1 + 8) >> 2) + (uint32_t)((c1 & 3) << 2);
    tc::sts64(a_hi + o1, hi.x, hi.y);
    tc::sts64(a_hi + o2, hi.z, hi.w);
    tc::sts64(a_lo + o1, lo.x, lo.y);
    tc::sts64(a_lo + o2, lo.z, lo.w);
}

template <int N_TILE>
struct TcSmem {
    static constexpr int A_BYTES = TC_M * 128;            // 128 rows x 32 fp32
    static constexpr int B_BYTES = N_TILE * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (N_TILE >= 256) ? 2 : ((N_TILE >= 128) ? 3 : 4);
    static_assert(STAGES * STAGE_BYTES <= 200 * 1024, "stage ring too large");
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int NB_OFF = TILE_BYTES;             // int nb[128][5]
    static constexpr int BAR_OFF = NB_OFF + TC_M * 5 * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;   // + barriers + tmem slot + alignment slack
};

template <int N_TILE, int PW>
__global__ void __launch_bounds__((PW + 1) * 32, 1)
meshconv_fwd_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                       const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                       int8_t* __restrict__ signs, int B, int E, int Cin, int Cout) {
    pdl_enter();
    using S = TcSmem<N_TILE>;
    constexpr int NTHREADS = (PW + 1) * 32;      // PW producer / epilogue warps + the MMA warp
    constexpr int PT = PW * 32;                  // producer threads
    constexpr int IT0 = (TC_M * 8) / PT;         // 16-byte items per thread in a G0 block (128 rows x 8 chunks)
    constexpr int ITP = (TC_M * 4) / PT;         // items per thread in a pair block (128 rows x 4 chunks, two gathers each)
    static_assert(IT0 == 2 * ITP && ITP >= 1, "producer mapping");
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
    const int dbg = g_tc_debug;

    // ---- setup ------------------------------------------------------------------------------------------------
    for (int i = tid; i < TC_M * 5; i += NTHREADS) {
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
    if (warp == PW) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, PW); tc::mbar_init(empty_bar + s, 1); }
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

    if (warp < PW) {
        // ======================================== producers ========================================================
        // Each thread owns the same tile rows for every K block, so its gather row ids live in registers and the global
        // loads of K block kb+2 are issued before the shared-memory stores of block kb (register prefetch, distance 2):
        // the L2 latency of the gathers is hidden behind two blocks of split / st.shared / MMA work.
        const int pt = tid;                                     // 0..255
        int src0[IT0], srcA[ITP][2], srcB[ITP][2];                    // blk 0: 4 items (row, chunk); pair blocks: 2 items x 2 rows
#pragma unroll
        for (int it = 0; it < IT0; ++it) src0[it] = nb[(pt + it * PT) >> 3][0];
#pragma unroll
        for (int it = 0; it < ITP; ++it) {
            const int r = pair_item_row(pt + it * PT);
            srcA[it][0] = nb[r][1]; srcB[it][0] = nb[r][3];     // pair (f1, f3)
            srcA[it][1] = nb[r][2]; srcB[it][1] = nb[r][4];     // pair (f2, f4)
        }
        struct Regs { float4 a[IT0]; };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_block = [&](int kb, Regs& R) {
            const int cb = kb / 5, blk = kb - cb * 5;
            if (dbg & 1) {
#pragma unroll
                for (int it = 0; it < IT0; ++it) R.a[it] = zero4;
                return;
            }
            if (blk == 0) {
                const int c = cb * 32 + ((pt & 7) << 2);
#pragma unroll
                for (int it = 0; it < IT0; ++it)
                    R.a[it] = src0[it] >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)src0[it] * Cin + c)) : zero4;
            } else {
                const int pr = (blk <= 2) ? 0 : 1;
                const int c = cb * 32 + (((blk - 1) & 1) << 4) + ((pt & 3) << 2);
#pragma unroll
                for (int it = 0; it < ITP; ++it) {
                    const int sa = pr ? srcA[it][1] : srcA[it][0], sb = pr ? srcB[it][1] : srcB[it][0];   // selects, not a dynamic index (local memory)
                    R.a[2 * it] = sa >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sa * Cin + c)) : zero4;
                    R.a[2 * it + 1] = sb >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sb * Cin + c)) : zero4;
                }
            }
        };
        auto store_block = [&](int kb, const Regs& R) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
#ifdef MCNN_TC_PROF
            const long long tp0 = clock64();
#endif
            if (lane == 0) tc::mbar_wait(empty_bar + s, ph ^ 1);      // one poller per warp
            __syncwarp();
#ifdef MCNN_TC_PROF
            const long long tp1 = clock64();
            if (blockIdx.x == 0 && tid == 0) { g_tc_prof[0] += tp1 - tp0; g_tc_prof[6] += 1; }
#endif
            unsigned char* a_hi = smem + s * S::STAGE_BYTES;
            unsigned char* a_lo = a_hi + S::A_BYTES;
            unsigned char* b_hi = a_lo + S::A_BYTES;
            unsigned char* b_lo = b_hi + S::B_BYTES;
            const int blk = kb % 5;
            float4 hi, lo;
            if (tid == 0 && !(dbg & 8)) {                                 // the B tile of this K block: two bulk copies (TMA engine)
                tc::mbar_expect_tx(full_bar + s, 2 * S::B_BYTES);
                const float* src = wt + ((size_t)kb * Cout + n0) * 32;
                tc::bulk_g2s(b_hi, src, S::B_BYTES, full_bar + s);
                tc::bulk_g2s(b_lo, src + (size_t)Cout * K, S::B_BYTES, full_bar + s);
            }
            if (dbg & 4) {
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(full_bar + s);
                return;
            }
            if (blk == 0) {
#pragma unroll
                for (int it = 0; it < IT0; ++it) {
                    const int item = pt + it * PT;
                    tc::split4(R.a[it], hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), item >> 3, 16 * (item & 1) + 2 * ((item & 7) >> 1), hi, lo);
                }
            } else {
#pragma unroll
                for (int it = 0; it < ITP; ++it) {
                    const int item = pt + it * PT;
                    const int r = pair_item_row(item), q = item & 3;
                    const float4 fa = R.a[2 * it], fb = R.a[2 * it + 1];
                    const float4 sum = make_float4(fa.x + fb.x, fa.y + fb.y, fa.z + fb.z, fa.w + fb.w);
                    const float4 dd = make_float4(fa.x - fb.x, fa.y - fb.y, fa.z - fb.z, fa.w - fb.w);
                    const float4 dif = make_float4(fabsf(dd.x), fabsf(dd.y), fabsf(dd.z), fabsf(dd.w));
                    if (signs != nullptr && blockIdx.y == 0 && m0 + r < M) {
                        // sign(f1 - f3) / sign(f2 - f4) of this edge, kept for the data-gradient kernel (abs backward)
                        const int cb_ = kb / 5, c_ = cb_ * 32 + (((blk - 1) & 1) << 4) + (q << 2);
                        const size_t so = ((size_t)(blk <= 2 ? 0 : 1) * M + (m0 + r)) * Cin + c_;
                        *reinterpret_cast<char4*>(signs + so) = make_char4((signed char)sgn(dd.x), (signed char)sgn(dd.y), (signed char)sgn(dd.z), (signed char)sgn(dd.w));
                    }
                    tc::split4(sum, hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), r, 2 * q, hi, lo);
                    tc::split4(dif, hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), r, 16 + 2 * q, hi, lo);
                }
            }
#ifdef MCNN_TC_PROF
            const long long tp2 = clock64();
#endif
            tc::fence_proxy_async_smem();                                  // every writer fences its own stores ...
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(full_bar + s);                  // ... then one arrival per warp
#ifdef MCNN_TC_PROF
            if (blockIdx.x == 0 && tid == 0) { const long long tp3 = clock64(); g_tc_prof[1] += tp2 - tp1; g_tc_prof[2] += tp3 - tp2; }
#endif
        };
        {                                                        // prefetch distance 2 (three register sets)
            // three register sets, prefetch distance 3.  The gathers of block kb + 3 are issued right AFTER block kb has been
            // handed over: fence.proxy.async waits for the thread's outstanding loads, so a load issued just before a
            // store_block would put the whole L2 latency on the critical path of every K block.
            Regs R0, R1, R2;
            load_block(0, R0);
            if (num_kb > 1) load_block(1, R1);
            if (num_kb > 2) load_block(2, R2);
            for (int kb = 0; kb < num_kb; kb += 3) {
                store_block(kb, R0);
                if (kb + 3 < num_kb) load_block(kb + 3, R0);
                if (kb + 1 < num_kb) {
                    store_block(kb + 1, R1);
                    if (kb + 4 < num_kb) load_block(kb + 4, R1);
                }
                if (kb + 2 < num_kb) {
                    store_block(kb + 2, R2);
                    if (kb + 5 < num_kb) load_block(kb + 5, R2);
                }
            }
        }
        // ======================================== epilogue ==========================================================
#ifdef MCNN_TC_PROF
        const long long te0 = clock64();
#endif
        tc::mbar_wait(accum_bar, 0);
        tc::tc_fence_after();
#ifdef MCNN_TC_PROF
        if (blockIdx.x == 0 && tid == 0) g_tc_prof[4] += clock64() - te0;
#endif
        const int lane_grp = warp & 3;                           // TMEM lanes 32*lane_grp .. +31
        const int r = lane_grp * 32 + lane;
        const int m = m0 + r;
        constexpr int COLS_PER_WARP = N_TILE / (PW / 4);         // the PW / 4 warps of a lane group split the columns
        static_assert(COLS_PER_WARP % 32 == 0, "epilogue reads 32 columns at a time");
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
#ifdef MCNN_TC_PROF
            const long long tm0 = clock64();
#endif
            tc::mbar_wait(full_bar + s, ph);
            tc::tc_fence_after();
#ifdef MCNN_TC_PROF
            if (blockIdx.x == 0 && lane == 0) g_tc_prof[3] += clock64() - tm0;
#endif
            if (tc::elect_one()) {
                const uint32_t a_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t a_lo = a_hi + S::A_BYTES;
                const uint32_t b_hi = a_lo + S::A_BYTES;
                const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                for (int kk = 0; kk < ((dbg & 2) ? 0 : 4); ++kk) {                 // 4 x (K = 8 tf32 = 32 bytes)
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
    if (warp == PW) {
        tc::tc_fence_after();
        tc::tmem_dealloc<N_TILE>(tmem_base);
    }
}

// =============================================================================================================
// Forward, persistent "stream-K" form.  The plain kernel above launches one CTA per 128-edge tile, and the tile counts of
// this workload sit just above a multiple of the SM count (150 / 225 / 300 / 375 tiles on 148 SMs): the last, nearly
// empty wave costs as much as a full one.  Here the work is cut into (tile, K block) units, unit u = tile * num_kb + kb,
// and CTA c of G <= #SMs resident CTAs takes the contiguous range [total*c/G, total*(c+1)/G): a tail of one tile, whole
// tiles, a head of the next.  A CTA that starts in the middle of a tile stores its partial accumulator to its own slot
// of a workspace and raises its flag; the CTA that holds the tile's K block 0 (the owner: it reaches that tile LAST, the
// others reach it FIRST, so it practically never waits) adds the partials in ascending CTA order, the bias, and writes
// the tile.  Deterministic: the split depends only on the shape and the grid.  The owner lowers each flag it consumed,
// so the workspace needs zeroing once, at allocation; one workspace must not be shared by concurrent launches.
// =============================================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int SK_FLAG_BYTES = 4096;
constexpr int SK_MAX_CTAS = 192;                 // workspace: flags + two 128 x 256 fp32 partial slots per CTA
constexpr int SK_EPI_WARPS = 4;                  // one per TMEM lane quarter
constexpr int SK_THREADS = (TC_PRODUCER_WARPS + 1 + SK_EPI_WARPS) * 32;

// Shared memory holds only the B (weight) tiles; the A operand is written to tensor memory by the producers.  TMEM (512
// columns): accumulator b at column b * N_TILE (two for N <= 128, one for N = 256), A stage s at 256 + 64 s (hi) / + 32 (lo).
template <int N_TILE>
struct SkSmem {
    static constexpr int B_BYTES = N_TILE * 128;
    static constexpr int STAGE_BYTES = 2 * B_BYTES;
    static constexpr int STAGES = 3;                                  // = A stages in TMEM
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int NB_OFF = TILE_BYTES;                         // int nb[2][128][5]: gather table of this and the next tile
    static constexpr int BAR_OFF = NB_OFF + 2 * TC_M * 5 * 4;
    static constexpr int NBARS = 2 * STAGES + 8;                      // full / empty per stage, accumulator and table full / empty x 2
    static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16 + 1024;
    static constexpr int NACC = N_TILE <= 128 ? 2 : 1;                // accumulators: with two, the epilogue of a segment overlaps the next one
    static constexpr int TMEM_COLS = 512;
    static constexpr uint32_t A_COL0 = 256;
};

// The segments of one CTA's unit range, in the order they are processed.  A tile that the range boundaries cut into pieces
// is OWNED by the CTA holding its longest piece (the first such on ties): every CTA works through its non-owned pieces
// first (each is published as a partial as soon as its accumulator is complete, without waiting for anybody), then its
// whole tiles, and last the pieces it owns -- by then the partials it has to add were written long ago.
struct SkSeg { int tile, ka, ke, kind, slot; };   // kind 0: whole tile, 1: partial (goes to `slot`), 2: owner of a cut tile

struct SkPlan {
    long long total, u0, u1;
    int G, cta, nkb;
    int n_early, n_whole, n_late, w0;
    int e_tile[2], e_ka[2], e_ke[2], e_slot[2];
    int l_tile[2], l_ka[2], l_ke[2];

    __device__ __forceinline__ long long start_of(int c) const { return total * c / G; }
    // first CTA whose range intersects `tile`, given one CTA that does
    __device__ int first_cta_of(int tile, int hint) const {
        const long long ts = (long long)tile * nkb;
        int c = hint;
        while (c > 0 && start_of(c) > ts) --c;
        while (start_of(c + 1) <= ts) ++c;
        return c;
    }
    __device__ int owner_of(int tile, int hint) const {
        const long long ts = (long long)tile * nkb, te = ts + nkb;
        int best = -1;
        long long best_len = 0;
        for (int c = first_cta_of(tile, hint); c < G; ++c) {
            const long long a = max(start_of(c), ts), b = min(start_of(c + 1), te);
            if (a >= te) break;
            if (b - a > best_len) { best_len = b - a; best = c; }
        }
        return best;
    }
    __device__ void add_piece(int tile, int ka, int ke, int at_start) {
        if (owner_of(tile, cta) == cta) { l_tile[n_late] = tile; l_ka[n_late] = ka; l_ke[n_late] = ke; ++n_late; }
        else { e_tile[n_early] = tile; e_ka[n_early] = ka; e_ke[n_early] = ke; e_slot[n_early] = 2 * cta + (at_start ? 0 : 1); ++n_early; }
    }
    __device__ void build(long long total_, int G_, int cta_, int nkb_) {
        total = total_; G = G_; cta = cta_; nkb = nkb_;
        u0 = start_of(cta); u1 = start_of(cta + 1);
        n_early = n_whole = n_late = 0; w0 = 0;
        const int tf = (int)(u0 / nkb), tl = (int)((u1 - 1) / nkb);
        const int ka = (int)(u0 - (long long)tf * nkb), ke = (int)(u1 - (long long)tl * nkb);       // first piece starts at ka, last ends at ke
        if (tf == tl) {
            if (ka == 0 && ke == nkb) { w0 = tf; n_whole = 1; }
            else add_piece(tf, ka, ke, 1);
            return;
        }
        int wa = tf, wb = tl;                                   // whole tiles wa .. wb
        if (ka > 0) { add_piece(tf, ka, nkb, 1); wa = tf + 1; }
        if (ke < nkb) { add_piece(tl, 0, ke, 0); wb = tl - 1; }
        w0 = wa; n_whole = wb - wa + 1;
    }
    __device__ __forceinline__ int nseg() const { return n_early + n_whole + n_late; }
    __device__ SkSeg seg(int i) const {
        SkSeg r;
        if (i < n_early) {
            const int j = i;                                     // selects, not dynamic indexing (would go to local memory)
            r.tile = j ? e_tile[1] : e_tile[0]; r.ka = j ? e_ka[1] : e_ka[0]; r.ke = j ? e_ke[1] : e_ke[0];
            r.slot = j ? e_slot[1] : e_slot[0]; r.kind = 1;
        } else if (i < n_early + n_whole) {
            r.tile = w0 + (i - n_early); r.ka = 0; r.ke = nkb; r.kind = 0; r.slot = -1;
        } else {
            const int j = i - n_early - n_whole;
            r.tile = j ? l_tile[1] : l_tile[0]; r.ka = j ? l_ka[1] : l_ka[0]; r.ke = j ? l_ke[1] : l_ke[0];
            r.kind = 2; r.slot = -1;
        }
        return r;
    }
};

// Warp roles: 0-7 producers (gather -> symmetric functions -> hi/lo split -> st.shared), 8 MMA issuer, 9-12 epilogue.
// The epilogue warps also prepare the gather table of the NEXT segment's tile while the current one is produced, and the
// accumulator is double-buffered in TMEM, so producers and tensor core run from one segment straight into the next.
template <int N_TILE>
__global__ void __launch_bounds__(SK_THREADS, 1)
meshconv_fwd_tc_sk_kernel(const float* __restrict__ x, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                          const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                          int8_t* __restrict__ signs, float* part, int* flags, int B, int E, int Cin, int Cout, int one_pass) {
    pdl_enter();
    using S = SkSmem<N_TILE>;
    using SS = SkSmem<N_TILE>;
    constexpr int PW = TC_PRODUCER_WARPS, PT = PW * 32;
    constexpr int NACC = SS::NACC;
    static_assert(PW == 8, "two producer warps per TMEM lane quarter");
    constexpr int ET = SK_EPI_WARPS * 32;
    SK_STAMP(threadIdx.x == 0, 61);
#ifdef MCNN_TC_PROF
    if (threadIdx.x == 0 && blockIdx.x < 256) g_sk_cta_ns[blockIdx.x][0] = globaltimer_ns();
#endif
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int (*nb)[TC_M][5] = reinterpret_cast<int (*)[TC_M][5]>(smem + SS::NB_OFF);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SS::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* acc_full = empty_bar + S::STAGES;              // [2]
    uint64_t* acc_empty = acc_full + 2;                      // [2]
    uint64_t* nb_full = acc_empty + 2;                       // [2]
    uint64_t* nb_empty = nb_full + 2;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nb_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int K = 5 * Cin;
    const int num_kb = K >> 5;
    const int ntn = Cout / N_TILE;
    const long long total = (long long)ceil_div(M, TC_M) * ntn * num_kb;
    const int G = gridDim.x, cta = blockIdx.x;
    SkPlan plan;
    plan.build(total, G, cta, num_kb);
    const int nseg = plan.nseg();
    SK_STAMP(tid == 0, 62);

    if (warp == PW) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, PW); tc::mbar_init(empty_bar + s, 1); }
            for (int i = 0; i < 2; ++i) {
                tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, SK_EPI_WARPS);
                tc::mbar_init(nb_full + i, SK_EPI_WARPS); tc::mbar_init(nb_empty + i, PW);
            }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<SS::TMEM_COLS>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    SK_STAMP(tid == 0, 0);

    // Epilogue of a whole tile or of the piece this CTA owns: accumulator (+ the other CTAs' pieces, ascending K order) + bias
    // -> out.  The calling warp reads its TMEM lane quarter and the 32-column chunks chunk0, chunk0 + step, ...
    auto finish_tile = [&](const SkSeg& sg, int seg, int chunk0, int step, bool all_warps) {
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int tile = sg.tile, buf = seg % NACC;
        const int mt = tile / ntn, nt = tile - mt * ntn;
        const int m0 = mt * TC_M, n0 = nt * N_TILE;
        const int lane_grp = warp & 3;                           // the TMEM lane quarter this warp may read
        const int r = lane_grp * 32 + lane;
        const int m = m0 + r;
        const uint32_t acc = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * N_TILE);
        int c_first = 0, n_cov = 0;                              // CTAs c_first .. c_first + n_cov - 1 hold pieces of this tile
        const long long ts = (long long)tile * num_kb, te = ts + num_kb;
        if (sg.kind == 2) {
            c_first = plan.first_cta_of(tile, cta);
            while (c_first + n_cov < G && plan.start_of(c_first + n_cov) < te) ++n_cov;
            if (tid == (PW + 1) * 32) {                          // first epilogue thread: wait for the pieces, lower their flags
                for (int j = 0; j < n_cov; ++j) {
                    const int c2 = c_first + j;
                    if (c2 == cta) continue;
                    volatile int* f = flags + 2 * c2 + (plan.start_of(c2) >= ts ? 0 : 1);
                    while (*f == 0) __nanosleep(64);
                    *f = 0;
                }
                __threadfence();
            }
            if (all_warps) named_bar_sync(3, PT + ET); else named_bar_sync(2, ET);
            SK_STAMP(tid == (PW + 1) * 32, 22 + 4 * seg);
        }
        auto slot_base = [&](int j) -> const float4* {
            const int c2 = c_first + j;
            const int slot = 2 * c2 + (plan.start_of(c2) >= ts ? 0 : 1);
            return reinterpret_cast<const float4*>(part) + (size_t)slot * (N_TILE / 4) * TC_M + r;
        };
        int j_first = -1;
        for (int j = n_cov - 1; j >= 0; --j) if (c_first + j != cta) j_first = j;
        const float4* p_first = (j_first >= 0 && m < M) ? slot_base(j_first) : nullptr;
        float4 nxt[8];                                           // the first foreign piece is fetched one chunk ahead
        if (p_first != nullptr && chunk0 * 32 < N_TILE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) nxt[j] = __ldcg(p_first + (size_t)(chunk0 * 8 + j) * TC_M);
        }
        tc::mbar_wait(acc_full + buf, (seg / NACC) & 1);
        tc::tc_fence_after();
        SK_STAMP(tid == (PW + 1) * 32, 20 + 4 * seg);
#pragma unroll 1
        for (int cc = chunk0 * 32; cc < N_TILE; cc += 32 * step) {
            float v[32];
            tc::tmem_ld32(acc + (uint32_t)cc, v);
            if (m < M) {
                if (p_first != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[4 * j] += nxt[j].x; v[4 * j + 1] += nxt[j].y; v[4 * j + 2] += nxt[j].z; v[4 * j + 3] += nxt[j].w; }
                    const int cn = cc + 32 * step;
                    if (cn < N_TILE) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) nxt[j] = __ldcg(p_first + (size_t)((cn >> 2) + j) * TC_M);
                    }
                    for (int jj = j_first + 1; jj < n_cov; ++jj) {          // a tile cut into 3+ pieces (more CTAs than tiles)
                        if (c_first + jj == cta) continue;
                        const float4* q4 = slot_base(jj) + (size_t)(cc >> 2) * TC_M;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = __ldcg(q4 + (size_t)j * TC_M);
                            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                        }
                    }
                }
                const int n = n0 + cc;
                float* o = out + (size_t)m * Cout + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 bv = zero4;
                    if (bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(bias + n + j));
                    *reinterpret_cast<float4*>(o + j) = make_float4(v[j] + bv.x, v[j + 1] + bv.y, v[j + 2] + bv.z, v[j + 3] + bv.w);
                }
            }
        }
    };

    if (warp < PW) {
        // ======================================== producers ========================================================
        // Warp w owns 16 rows of the tile: TMEM lanes 32 (w % 4) + 16 (w / 4) + [0, 16); thread t holds rows t/4 and t/4 + 8 of
        // them.  Per K block it gathers, for each of its two rows, 32 contiguous bytes (G0 block: channels 8 (t%4) .. +7) or
        // 16 + 16 (pair block: channels 4 (t%4) .. +3 of both gathered rows), forms the symmetric functions, splits and writes
        // the values into tensor memory with tcgen05.st .16x256b.x4 -- the registers are already in that layout (fw_kpos).
        const int rowA = (warp & 3) * 32 + (warp >> 2) * 16 + (lane >> 2);       // and rowA + 8
        const int q4 = lane & 3;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32 + (warp >> 2) * 16) << 16;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        struct Regs { float4 a[4]; };            // G0: rowA ch 0-3, 4-7, rowB ch 0-3, 4-7;  pair: rowA first, second row; rowB first, second
        uint32_t g = 0;                                          // K blocks this CTA has pushed through the stage ring
        for (int seg = 0; seg < nseg; ++seg) {
            const SkSeg sg = plan.seg(seg);
            const int tile = sg.tile, ka = sg.ka, ke = sg.ke;
            const int mt = tile / ntn, nt = tile - mt * ntn;
            const int m0 = mt * TC_M, n0 = nt * N_TILE;
            const int nbuf = seg & 1;
            tc::mbar_wait(nb_full + nbuf, (seg >> 1) & 1);
            SK_STAMP(tid == 0, 1 + 4 * seg);
            int src[2][5];                                       // gather rows of rowA / rowA + 8: self, then the four neighbours
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int kq = 0; kq < 5; ++kq) src[h][kq] = nb[nbuf][rowA + 8 * h][kq];
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(nb_empty + nbuf);
            auto load_block = [&](int kb, Regs& R) {
                const int cb = kb / 5, blk = kb - cb * 5;
                if (blk == 0) {
                    const int c = cb * 32 + (q4 << 3);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float* p = x + (size_t)src[h][0] * Cin + c;
                        R.a[2 * h] = src[h][0] >= 0 ? __ldg(reinterpret_cast<const float4*>(p)) : zero4;
                        R.a[2 * h + 1] = src[h][0] >= 0 ? __ldg(reinterpret_cast<const float4*>(p + 4)) : zero4;
                    }
                } else {
                    const int pr = (blk <= 2) ? 0 : 1;
                    const int c = cb * 32 + (((blk - 1) & 1) << 4) + (q4 << 2);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int sa = pr ? src[h][2] : src[h][1], sb = pr ? src[h][4] : src[h][3];     // selects, not a dynamic index
                        R.a[2 * h] = sa >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sa * Cin + c)) : zero4;
                        R.a[2 * h + 1] = sb >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sb * Cin + c)) : zero4;
                    }
                }
            };
            auto store_block = [&](int kb, const Regs& R) {
                const uint32_t s = g % S::STAGES;
                const uint32_t ph = (g / S::STAGES) & 1;
                ++g;
                if (lane == 0) tc::mbar_wait(empty_bar + s, ph ^ 1);
                __syncwarp();
                tc::tc_fence_after();
                const int blk = kb % 5;
                if (tid == 0) {
                    unsigned char* b_hi = smem + s * S::STAGE_BYTES;
                    tc::mbar_expect_tx(full_bar + s, one_pass ? S::B_BYTES : 2 * S::B_BYTES);
                    const float* srcw = wt + ((size_t)kb * Cout + n0) * 32;
                    tc::bulk_g2s(b_hi, srcw, S::B_BYTES, full_bar + s);
                    if (!one_pass) tc::bulk_g2s(b_hi + S::B_BYTES, srcw + (size_t)Cout * K, S::B_BYTES, full_bar + s);
                }
                float va[8], vb[8];                              // the 8 K entries of rowA / rowA + 8 owned by this thread
                if (blk == 0) {
                    va[0] = R.a[0].x; va[1] = R.a[0].y; va[2] = R.a[0].z; va[3] = R.a[0].w;
                    va[4] = R.a[1].x; va[5] = R.a[1].y; va[6] = R.a[1].z; va[7] = R.a[1].w;
                    vb[0] = R.a[2].x; vb[1] = R.a[2].y; vb[2] = R.a[2].z; vb[3] = R.a[2].w;
                    vb[4] = R.a[3].x; vb[5] = R.a[3].y; vb[6] = R.a[3].z; vb[7] = R.a[3].w;
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 fa = R.a[2 * h], fb = R.a[2 * h + 1];
                        const float dd[4] = {fa.x - fb.x, fa.y - fb.y, fa.z - fb.z, fa.w - fb.w};
                        float* v = h ? vb : va;
                        v[0] = fa.x + fb.x; v[1] = fa.y + fb.y; v[2] = fa.z + fb.z; v[3] = fa.w + fb.w;
                        v[4] = fabsf(dd[0]); v[5] = fabsf(dd[1]); v[6] = fabsf(dd[2]); v[7] = fabsf(dd[3]);
                        const int m = m0 + rowA + 8 * h;
                        if (signs != nullptr && nt == 0 && m < M) {
                            // sign(f1 - f3) / sign(f2 - f4) of this edge, kept for the data-gradient kernel (abs backward)
                            const int c_ = (kb / 5) * 32 + (((blk - 1) & 1) << 4) + (q4 << 2);
                            const size_t so = ((size_t)(blk <= 2 ? 0 : 1) * M + m) * Cin + c_;
                            *reinterpret_cast<char4*>(signs + so) = make_char4((signed char)sgn(dd[0]), (signed char)sgn(dd[1]), (signed char)sgn(dd[2]), (signed char)sgn(dd[3]));
                        }
                    }
                }
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    tc::split_tf32(va[2 * i], hi[4 * i], lo[4 * i]);
                    tc::split_tf32(va[2 * i + 1], hi[4 * i + 1], lo[4 * i + 1]);
                    tc::split_tf32(vb[2 * i], hi[4 * i + 2], lo[4 * i + 2]);
                    tc::split_tf32(vb[2 * i + 1], hi[4 * i + 3], lo[4 * i + 3]);
                }
                const uint32_t acol = tmem_base + lane_field + SS::A_COL0 + 64u * s;
                tc::tmem_st_16x256b_x4(acol, hi);
                if (!one_pass) tc::tmem_st_16x256b_x4(acol + 32u, lo);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(full_bar + s);
            };
            {
                Regs R0, R1, R2;
                load_block(ka, R0);
                if (ka + 1 < ke) load_block(ka + 1, R1);
                if (ka + 2 < ke) load_block(ka + 2, R2);
                for (int kb = ka; kb < ke; kb += 3) {
                    store_block(kb, R0);
                    if (kb + 3 < ke) load_block(kb + 3, R0);
                    if (kb + 1 < ke) {
                        store_block(kb + 1, R1);
                        if (kb + 4 < ke) load_block(kb + 4, R1);
                    }
                    if (kb + 2 < ke) {
                        store_block(kb + 2, R2);
                        if (kb + 5 < ke) load_block(kb + 5, R2);
                    }
                }
            }
            SK_STAMP(tid == 0, 2 + 4 * seg);
        }
        // nothing left to produce: help the epilogue warps with the CTA's last tile (two more warps per TMEM lane quarter)
        const SkSeg lastseg = plan.seg(nseg - 1);
        if (lastseg.kind != 1) {
            finish_tile(lastseg, nseg - 1, 1 + (warp >> 2), 3, true);
            tc::tc_fence_before();
        }
    } else if (warp == PW) {
        // ======================================== MMA issuer ========================================================
        constexpr uint32_t idesc = tc::make_idesc_tf32(N_TILE);
        uint32_t g = 0;
        for (int seg = 0; seg < nseg; ++seg) {
            const SkSeg sg = plan.seg(seg);
            const int ka = sg.ka, ke = sg.ke;
            const int buf = seg % NACC;
            const uint32_t acc = tmem_base + (uint32_t)(buf * N_TILE);
            if (seg >= NACC) {                                   // the epilogue of segment seg - NACC has drained this accumulator
                tc::mbar_wait(acc_empty + buf, ((seg / NACC) - 1) & 1);
                tc::tc_fence_after();
            }
            for (int kb = ka; kb < ke; ++kb, ++g) {
                const uint32_t s = g % S::STAGES;
                const uint32_t ph = (g / S::STAGES) & 1;
                tc::mbar_wait(full_bar + s, ph);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_hi = tmem_base + SS::A_COL0 + 64u * s, a_lo = a_hi + 32u;
                    const uint32_t b_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
                    const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t dbh = tc::make_desc_sw128(b_hi + kk * 32), dbl = tc::make_desc_sw128(b_lo + kk * 32);
                        tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbh, idesc, (kb != ka) || (kk != 0));
                        if (!one_pass) {
                            tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbl, idesc, 1);
                            tc::mma_tf32_ts(acc, a_lo + 8 * kk, dbh, idesc, 1);
                        }
                    }
                    tc::mma_commit(empty_bar + s);
                    if (kb == ke - 1) tc::mma_commit(acc_full + buf);
                }
                __syncwarp();
            }
            SK_STAMP(lane == 0, 3 + 4 * seg);
        }
        tc::tc_fence_before();
    } else {
        // ======================================== epilogue + gather tables =========================================
        const int et = tid - (PW + 1) * 32;                      // 0 .. 127
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fill_table = [&](int tile_, int buf_) {
            const int m0_ = (tile_ / ntn) * TC_M;
            for (int i = et; i < TC_M * 5; i += ET) {
                const int r = i / 5, k = i - r * 5;
                const int m = m0_ + r;
                int idx = -1;
                if (m < M) {
                    const int b = m / E, e = m - b * E;
                    if (e >= __ldg(ec + b)) idx = b * E;                               // padded slot -> edge 0
                    else if (k == 0) idx = m;
                    else { const int gg = __ldg(gemm + (size_t)m * 4 + (k - 1)); idx = gg < 0 ? -1 : b * E + gg; }
                }
                nb[buf_][r][k] = idx;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(nb_full + buf_);
        };
        fill_table(plan.seg(0).tile, 0);
        for (int seg = 0; seg < nseg; ++seg) {
            const SkSeg sg = plan.seg(seg);
            const int tile = sg.tile;
            const int mt = tile / ntn, nt = tile - mt * ntn;
            const int m0 = mt * TC_M, n0 = nt * N_TILE;
            const int nbuf = seg & 1, buf = seg % NACC;
            SK_STAMP(et == 0, 23 + 4 * seg);
            if (seg + 1 < nseg) {                                // gather table of the next segment's tile
                if (seg >= 1) tc::mbar_wait(nb_empty + (nbuf ^ 1), ((seg - 1) >> 1) & 1);   // producers copied segment seg - 1's table out
                fill_table(plan.seg(seg + 1).tile, nbuf ^ 1);
            }
            const int lane_grp = warp & 3;                       // the TMEM lane quarter this warp may read
            const int r = lane_grp * 32 + lane;
            const int m = m0 + r;
            const uint32_t acc = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * N_TILE);
            if (sg.kind == 1) {
                // ---- a piece of a tile another CTA owns: publish the raw accumulator ----
                tc::mbar_wait(acc_full + buf, (seg / NACC) & 1);
                tc::tc_fence_after();
                SK_STAMP(et == 0, 20 + 4 * seg);
#pragma unroll 1
                for (int cc = 0; cc < N_TILE; cc += 32) {
                    float v[32];
                    tc::tmem_ld32(acc + (uint32_t)cc, v);
                    if (m < M) {
                        // slot layout [column / 4][row] of float4: the 32 lanes of a store (32 rows) touch 512 contiguous bytes
                        float4* o = reinterpret_cast<float4*>(part) + ((size_t)sg.slot * (N_TILE / 4) + (cc >> 2)) * TC_M + r;
#pragma unroll
                        for (int j = 0; j < 8; ++j) __stcg(o + j * TC_M, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(acc_empty + buf);
                __threadfence();                                 // every writer fences, then one flag store
                named_bar_sync(2, ET);
                if (et == 0) { volatile int* f = flags + sg.slot; *f = 1; }
                SK_STAMP(et == 0, 21 + 4 * seg);
                continue;
            }
            // ---- whole tile, or the owned piece of a cut tile; the producer warps help with the CTA's last one ----
            const bool all_warps = seg == nseg - 1;
            finish_tile(sg, seg, 0, all_warps ? 3 : 1, all_warps);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_empty + buf);     // accumulator free for segment seg + 2
            SK_STAMP(et == 0, 21 + 4 * seg);
        }
    }
    SK_STAMP(tid == 0, 60);
    __syncthreads();
    if (warp == PW) {
        tc::tc_fence_after();
        tc::tmem_dealloc<SS::TMEM_COLS>(tmem_base);
        SK_STAMP(lane == 0, 59);
#ifdef MCNN_TC_PROF
        if (lane == 0 && blockIdx.x < 256) g_sk_cta_ns[blockIdx.x][1] = globaltimer_ns();
#endif
    }
}

// =============================================================================================================
// Forward on a CTA PAIR (tcgen05 cta_group::2): two CTAs of one cluster compute 256 edges x 256 output channels.  Each CTA
// gathers / builds the A operand of its own 128 edges and fetches only HALF of the weight tile (128 of the 256 rows of B);
// the pair's tensor cores share both halves.  Per SM this halves the B traffic from L2 and the B reads from shared memory
// -- the two things that bound the single-CTA kernel at N = 256 -- and the smaller stage (64 KB) buys a third pipeline stage.
// Hand-over: the producers of BOTH CTAs arrive on the LEADER's full barrier (remote arrive through the cluster window); the
// leader's elected thread issues the MMAs for the pair and its commit multicasts the stage-free signal to both CTAs.
// =============================================================================================================
struct Tc2Smem {
    static constexpr int A_BYTES = TC_M * 128;            // 128 rows x 32 fp32
    static constexpr int B_BYTES = 128 * 128;              // this CTA's half of the 256-row B tile
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 65536
    static constexpr int STAGES = 3;
    static constexpr int NB_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = NB_OFF + TC_M * 5 * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
meshconv_fwd_tc2_kernel(const float* __restrict__ x, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                        const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ out,
                        int8_t* __restrict__ signs, int B, int E, int Cin, int Cout) {
    pdl_enter();
    using S = Tc2Smem;
    constexpr int N_TILE = 256;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int (*nb)[5] = reinterpret_cast<int (*)[5]>(smem + S::NB_OFF);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);      // used in the leader: 16 warp arrivals + its own B bytes
    uint64_t* empty_bar = full_bar + S::STAGES;                                // in each CTA: one multicast commit
    uint64_t* bfull_bar = empty_bar + S::STAGES;                               // peer only: its half of B has landed
    uint64_t* accum_bar = bfull_bar + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int M = B * E;
    const int m0 = blockIdx.x * TC_M, n0 = blockIdx.y * N_TILE;
    const int K = 5 * Cin;
    const int num_kb = K >> 5;

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
            for (int s = 0; s < S::STAGES; ++s) {
                tc::mbar_init(full_bar + s, 2 * TC_PRODUCER_WARPS);
                tc::mbar_init(empty_bar + s, 1);
                tc::mbar_init(bfull_bar + s, 1);
            }
            tc::mbar_init(accum_bar, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc_pair<N_TILE>(tmem_slot);
    }
    tc::tc_fence_before();
    tc::cluster_sync_all();                                  // both CTAs: barriers initialised, TMEM allocated
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        const int pt = tid;
        int src0[4], srcA[2][2], srcB[2][2];
#pragma unroll
        for (int it = 0; it < 4; ++it) src0[it] = nb[(pt + it * 256) >> 3][0];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int r = pair_item_row(pt + it * 256);
            srcA[it][0] = nb[r][1]; srcB[it][0] = nb[r][3];
            srcA[it][1] = nb[r][2]; srcB[it][1] = nb[r][4];
        }
        struct Regs { float4 a[4]; };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_block = [&](int kb, Regs& R) {
            const int cb = kb / 5, blk = kb - cb * 5;
            if (blk == 0) {
                const int c = cb * 32 + ((pt & 7) << 2);
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    R.a[it] = src0[it] >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)src0[it] * Cin + c)) : zero4;
            } else {
                const int pr = (blk <= 2) ? 0 : 1;
                const int c = cb * 32 + (((blk - 1) & 1) << 4) + ((pt & 3) << 2);
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int sa = pr ? srcA[it][1] : srcA[it][0], sb = pr ? srcB[it][1] : srcB[it][0];
                    R.a[2 * it] = sa >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sa * Cin + c)) : zero4;
                    R.a[2 * it + 1] = sb >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)sb * Cin + c)) : zero4;
                }
            }
        };
        auto store_block = [&](int kb, const Regs& R) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            if (lane == 0) tc::mbar_wait_cluster(empty_bar + s, ph ^ 1);
            __syncwarp();
            unsigned char* a_hi = smem + s * S::STAGE_BYTES;
            unsigned char* a_lo = a_hi + S::A_BYTES;
            unsigned char* b_hi = a_lo + S::A_BYTES;
            unsigned char* b_lo = b_hi + S::B_BYTES;
            const int blk = kb % 5;
            float4 hi, lo;
            if (tid == 0) {                                               // this CTA's 128 rows of the B tile
                const float* src = wt + ((size_t)kb * Cout + n0 + rank * 128) * 32;
                uint64_t* bbar = rank == 0 ? full_bar + s : bfull_bar + s;
                if (rank == 0) tc::mbar_expect_tx(bbar, 2 * S::B_BYTES);
                else tc::mbar_arrive_expect_tx(bbar, 2 * S::B_BYTES);
                tc::bulk_g2s(b_hi, src, S::B_BYTES, bbar);
                tc::bulk_g2s(b_lo, src + (size_t)Cout * K, S::B_BYTES, bbar);
            }
            if (blk == 0) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int item = pt + it * 256;
                    tc::split4(R.a[it], hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), item >> 3, 16 * (item & 1) + 2 * ((item & 7) >> 1), hi, lo);
                }
            } else {
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int item = pt + it * 256;
                    const int r = pair_item_row(item), q = item & 3;
                    const float4 fa = R.a[2 * it], fb = R.a[2 * it + 1];
                    const float4 sum = make_float4(fa.x + fb.x, fa.y + fb.y, fa.z + fb.z, fa.w + fb.w);
                    const float4 dd = make_float4(fa.x - fb.x, fa.y - fb.y, fa.z - fb.z, fa.w - fb.w);
                    const float4 dif = make_float4(fabsf(dd.x), fabsf(dd.y), fabsf(dd.z), fabsf(dd.w));
                    if (signs != nullptr && blockIdx.y == 0 && m0 + r < M) {
                        // sign(f1 - f3) / sign(f2 - f4) of this edge, kept for the data-gradient kernel (abs backward)
                        const int cb_ = kb / 5, c_ = cb_ * 32 + (((blk - 1) & 1) << 4) + (q << 2);
                        const size_t so = ((size_t)(blk <= 2 ? 0 : 1) * M + (m0 + r)) * Cin + c_;
                        *reinterpret_cast<char4*>(signs + so) = make_char4((signed char)sgn(dd.x), (signed char)sgn(dd.y), (signed char)sgn(dd.z), (signed char)sgn(dd.w));
                    }
                    tc::split4(sum, hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), r, 2 * q, hi, lo);
                    tc::split4(dif, hi, lo);
                    sts_kperm(tc::smem_u32(a_hi), tc::smem_u32(a_lo), r, 16 + 2 * q, hi, lo);
                }
            }
            tc::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (rank != 0 && warp == 0) tc::mbar_wait(bfull_bar + s, ph);   // the peer vouches for its half of B as well
                tc::mbar_arrive_leader(full_bar + s);
            }
        };
        {
            // three register sets, prefetch distance 3.  The gathers of block kb + 3 are issued right AFTER block kb has been
            // handed over: fence.proxy.async waits for the thread's outstanding loads, so a load issued just before a
            // store_block would put the whole L2 latency on the critical path of every K block.
            Regs R0, R1, R2;
            load_block(0, R0);
            if (num_kb > 1) load_block(1, R1);
            if (num_kb > 2) load_block(2, R2);
            for (int kb = 0; kb < num_kb; kb += 3) {
                store_block(kb, R0);
                if (kb + 3 < num_kb) load_block(kb + 3, R0);
                if (kb + 1 < num_kb) {
                    store_block(kb + 1, R1);
                    if (kb + 4 < num_kb) load_block(kb + 4, R1);
                }
                if (kb + 2 < num_kb) {
                    store_block(kb + 2, R2);
                    if (kb + 5 < num_kb) load_block(kb + 5, R2);
                }
            }
        }
        // ---- epilogue: this CTA's 128 rows of the accumulator ----
        tc::mbar_wait_cluster(accum_bar, 0);
        tc::tc_fence_after();
        const int lane_grp = warp & 3;
        const int r = lane_grp * 32 + lane;
        const int m = m0 + r;
        constexpr int COLS_PER_WARP = N_TILE / 2;
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
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(bias + n + j));
                    *reinterpret_cast<float4*>(o + j) = make_float4(v[j] + bv.x, v[j + 1] + bv.y, v[j + 2] + bv.z, v[j + 3] + bv.w);
                }
            }
        }
        tc::tc_fence_before();
    } else if (rank == 0) {
        // ======================================== MMA issuer (leader CTA only) ========================================
        constexpr uint32_t idesc = tc::make_idesc_tf32_pair(N_TILE);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            tc::mbar_wait_cluster(full_bar + s, ph);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t a_lo = a_hi + S::A_BYTES;
                const uint32_t b_hi = a_lo + S::A_BYTES;
                const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t dah = tc::make_desc_sw128(a_hi + kk * 32), dal = tc::make_desc_sw128(a_lo + kk * 32);
                    const uint64_t dbh = tc::make_desc_sw128(b_hi + kk * 32), dbl = tc::make_desc_sw128(b_lo + kk * 32);
                    tc::mma_tf32_pair(tmem_base, dah, dbh, idesc, (kb | kk) != 0);
                    tc::mma_tf32_pair(tmem_base, dah, dbl, idesc, 1);
                    tc::mma_tf32_pair(tmem_base, dal, dbh, idesc, 1);
                }
                tc::mma_commit_pair(empty_bar + s);
                if (kb == num_kb - 1) tc::mma_commit_pair(accum_bar);
            }
            __syncwarp();
        }
        tc::tc_fence_before();
    }
    tc::cluster_sync_all();                                  // both epilogues done before the pair's TMEM goes away
    if (warp == TC_PRODUCER_WARPS) {
        tc::tc_fence_after();
        tc::tmem_dealloc_pair<N_TILE>(tmem_base);
    }
}

static int launch_fwd_tc2(const float* x, const int32_t* gemm, const int32_t* ec, const float* wt, const float* bias,
                          float* out, int8_t* signs, int B, int E, int Cin, int Cout, cudaStream_t st) {
    using S = Tc2Smem;
    static DynSmemGuard configured;
    if (int grc = configured.ensure(meshconv_fwd_tc2_kernel, S::TOTAL)) return grc;
    const int mtiles = ceil_div(B * E, TC_M);
    dim3 grid(2 * ceil_div(mtiles, 2), Cout / 256);          // whole CTA pairs; a surplus CTA works on rows >= M (all zero)
    launch_k(meshconv_fwd_tc2_kernel, grid, TC_THREADS, S::TOTAL, st, x, gemm, ec, wt, bias, out, signs, B, E, Cin, Cout);
    return after_launch();
}

template <int N_TILE, int PW>
static int launch_fwd_tc(const float* x, const int32_t* gemm, const int32_t* ec, const float* wt, const float* bias,
                         float* out, int8_t* signs, int B, int E, int Cin, int Cout, cudaStream_t st) {
    using S = TcSmem<N_TILE>;
    static DynSmemGuard configured;
    if (int grc = configured.ensure(meshconv_fwd_tc_kernel<N_TILE, PW>, S::TOTAL)) return grc;
    dim3 grid(ceil_div(B * E, TC_M), ceil_div(Cout, N_TILE));
    launch_k(meshconv_fwd_tc_kernel<N_TILE, PW>, grid, (PW + 1) * 32, S::TOTAL, st, x, gemm, ec, wt, bias, out, signs, B, E, Cin, Cout);
    return after_launch();
}


static int g_tc_sk_ctas = 0;               // debug: force the stream-K grid size (0 = heuristic)
static int g_tc_bd_tile = 0;               // debug: 1 = one CTA per item for the data gradient instead of the persistent kernel

template <int N_TILE>
static int launch_fwd_tc_sk(const float* x, const int32_t* gemm, const int32_t* ec, const float* wt, const float* bias,
                            float* out, int8_t* signs, void* sk_ws, int B, int E, int Cin, int Cout, int one_pass, cudaStream_t st) {
    using S = SkSmem<N_TILE>;
    static DynSmemGuard configured;
    if (int grc = configured.ensure(meshconv_fwd_tc_sk_kernel<N_TILE>, S::TOTAL)) return grc;
    const long long total = (long long)ceil_div(B * E, TC_M) * (Cout / N_TILE) * (5 * Cin / 32);
    long long G = min((long long)min(sm_count(), SK_MAX_CTAS), max(1LL, total / 4));     // at least 4 K blocks per CTA
    if (g_tc_sk_ctas > 0) G = min((long long)g_tc_sk_ctas, total);
    int* flags = reinterpret_cast<int*>(sk_ws);
    float* part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sk_ws) + SK_FLAG_BYTES);
    launch_k(meshconv_fwd_tc_sk_kernel<N_TILE>, (int)G, SK_THREADS, S::TOTAL, st, x, gemm, ec, wt, bias, out, signs, part, flags, B, E, Cin, Cout, one_pass);
    return after_launch();
}

// =============================================================================================================
// backward, data:   dG[m][(k5, c)] = sum_o dout[m][o] * W[o][c][k5]   on tcgen05, then the abs/add backward and the
// scatter-add to the gather rows in the epilogue.  One CTA = 128 edges x (5 x 32 channels) = N 160; K = Cout.
//   A = dout rows (K-major as they lie in memory), B = weight re-laid as wtb[(cb*160 + k5*32 + cl)][o] (K-major).
// =============================================================================================================
constexpr int BD_N = 160;
// Accumulator column (within a tap's block of 32) of input channel cl (within the CTA's block of 32).  Permuted so that
// the mma-fragment TMEM load (tcgen05.ld .16x256b.x2 at column 0 or 16 of the block) hands thread t of a warp the four
// CONSECUTIVE channels 16*h + 4*(t%4) .. +3 of an edge row: the four threads of a quad then scatter 64 contiguous bytes of
// one dx row with one red.v4 each, and a warp instruction touches 8 rows instead of 32 (the scatter is bound by the
// number of distinct lines per instruction, one L1 wavefront each).
// K position (within a 32-wide K block = 32 output channels) of output channel p of the block.  Permuted so that the
// producers of the persistent kernel can write the A operand (dout rows) straight into TENSOR MEMORY in the mma-fragment
// layout (tcgen05.st .16x256b: thread t holds columns 8i + 2(t%4) + {0,1}) from 32 CONTIGUOUS bytes of a row per thread:
// memory channel 8q + j  <->  K column 8(j/2) + 2q + (j%2).  A and B use the same order, so the product is unchanged.
__host__ __device__ constexpr int bd_kperm(int p) { return 8 * ((p & 7) >> 1) + 2 * (p >> 3) + (p & 1); }
__host__ __device__ constexpr int bd_col(int cl) { return 8 * (2 * (cl >> 4) + ((cl >> 1) & 1)) + 2 * ((cl & 15) >> 2) + (cl & 1); }

struct BdSmem {
    static constexpr int A_BYTES = TC_M * 128;
    static constexpr int B_BYTES = BD_N * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 73728
    static constexpr int STAGES = 3;
    static constexpr int NB_OFF = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = NB_OFF + TC_M * 5 * 4;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;
};

// weight [Cout][Cin][5] -> B operand image of the data-gradient GEMM, TF32 head / remainder:
//   img[hi|lo][cb][kb][160 rows x 128 bytes swizzled],  row n = k5*32 + bd_col(cl) (c = cb*32 + cl),  K = output channels
__global__ void weight_tc_bwd_layout_kernel(const float* __restrict__ w, float* __restrict__ img, int Cin, int Cout) {
    pdl_enter();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 5 * Cin * Cout) return;
    const int row = idx / Cout, o = idx - row * Cout;
    const int cb = row / BD_N, n = row - cb * BD_N;
    const int k5 = n >> 5, c = cb * 32 + (n & 31);
    const int kb = o >> 5, p = o & 31;
    float hi, lo;
    tc::split_tf32(w[((size_t)o * Cin + c) * 5 + k5], hi, lo);
    const int nrow = k5 * 32 + bd_col(n & 31);                 // B row = accumulator column of (tap, channel)
    const int pk = bd_kperm(p);                                 // K position of this output channel inside its block
    const size_t off = ((size_t)cb * (Cout >> 5) + kb) * (BD_N * 32) + (tc::sw128_off(nrow, pk >> 2) >> 2) + (pk & 3);
    img[off] = hi;
    img[(size_t)5 * Cin * Cout + off] = lo;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
meshconv_bwd_data_tc_kernel(const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ gemm,
                            const int32_t* __restrict__ ec, const float* __restrict__ wtb, float* __restrict__ dx,
                            const int8_t* __restrict__ signs, int B, int E, int Cin, int Cout) {
    pdl_enter();
    using S = BdSmem;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    int (*nb)[5] = reinterpret_cast<int (*)[5]>(smem + S::NB_OFF);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* accum_bar = empty_bar + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int m0 = blockIdx.x * TC_M, cb = blockIdx.y;
    const int num_kb = Cout >> 5;
    const int dbg = g_tc_debug;            // diagnostics: bit 4 = no scatter, bit 5 = no sign loads (results wrong)

    for (int i = tid; i < TC_M * 5; i += TC_THREADS) {
        const int r = i / 5, k = i - r * 5;
        const int m = m0 + r;
        int idx = -1;
        if (m < M) {
            const int b = m / E, e = m - b * E;
            if (e >= __ldg(ec + b)) idx = b * E;
            else if (k == 0) idx = m;
            else { const int g = __ldg(gemm + (size_t)m * 4 + (k - 1)); idx = g < 0 ? -1 : b * E + g; }
        }
        nb[r][k] = idx;
    }
    if (warp == TC_PRODUCER_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, TC_PRODUCER_WARPS); tc::mbar_init(empty_bar + s, 1); }
            tc::mbar_init(accum_bar, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<256>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        const int pt = tid;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            if (lane == 0) tc::mbar_wait(empty_bar + s, ph ^ 1);      // one poller per warp
            __syncwarp();
            unsigned char* a_hi = smem + s * S::STAGE_BYTES;
            unsigned char* a_lo = a_hi + S::A_BYTES;
            unsigned char* b_hi = a_lo + S::A_BYTES;
            unsigned char* b_lo = b_hi + S::B_BYTES;
#pragma unroll
            for (int it = 0; it < (TC_M * 8) / 256; ++it) {
                const int item = pt + it * 256;
                const int r = item >> 3, q = item & 7;
                const int m = m0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < M) v = __ldg(reinterpret_cast<const float4*>(dout + (size_t)m * Cout + (kb << 5) + (q << 2)));
                float4 hi, lo;
                tc::split4(v, hi, lo);
                // permuted K order (bd_kperm): the chunk's channels 4q .. 4q+3 are K columns c1, c1+1 and c1+8, c1+9
                const int c1 = 16 * (q & 1) + 2 * (q >> 1);
                const uint32_t o1 = tc::sw128_off(r, c1 >> 2) + (uint32_t)((c1 & 3) << 2);
                const uint32_t o2 = tc::sw128_off(r, (c1 + 8) >> 2) + (uint32_t)((c1 & 3) << 2);
                tc::sts64(tc::smem_u32(a_hi) + o1, hi.x, hi.y);
                tc::sts64(tc::smem_u32(a_hi) + o2, hi.z, hi.w);
                tc::sts64(tc::smem_u32(a_lo) + o1, lo.x, lo.y);
                tc::sts64(tc::smem_u32(a_lo) + o2, lo.z, lo.w);
            }
            if (tid == 0) {                                               // B tile: two bulk copies of the pre-built image
                tc::mbar_expect_tx(full_bar + s, 2 * S::B_BYTES);
                const float* src = wtb + ((size_t)cb * num_kb + kb) * (BD_N * 32);
                tc::bulk_g2s(b_hi, src, S::B_BYTES, full_bar + s);
                tc::bulk_g2s(b_lo, src + (size_t)5 * Cin * Cout, S::B_BYTES, full_bar + s);
            }
            tc::fence_proxy_async_smem();                                  // every writer fences its own stores ...
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(full_bar + s);                  // ... then one arrival per warp
        }
        // ---- epilogue: abs/add backward + scatter ----------------------------------------------
        tc::mbar_wait(accum_bar, 0);
        tc::tc_fence_after();
        const int lane_grp = warp & 3;
        const int r = lane_grp * 32 + lane;
        const int m = m0 + r;
        const int cl0 = (warp >> 2) * 16;                     // the two warps of a lane group split the 32 channels
        float g[5][16];                                        // g[k5][bd_col(co)] = tap k5, channel c0 + co (columns are permuted)
#pragma unroll
        for (int k5 = 0; k5 < 5; ++k5) tc::tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(k5 * 32 + cl0), g[k5]);
        if (m < M && !(dbg & 16)) {
            const int b = m / E, e = m - b * E;
            const int c0 = cb * 32 + cl0;
            if (e >= __ldg(ec + b)) {
                float* d = dx + (size_t)(b * E) * Cin + c0;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    tc::red_add_v4(d + j, g[0][bd_col(j)] + 2.f * (g[1][bd_col(j)] + g[2][bd_col(j)]), g[0][bd_col(j + 1)] + 2.f * (g[1][bd_col(j + 1)] + g[2][bd_col(j + 1)]),
                                   g[0][bd_col(j + 2)] + 2.f * (g[1][bd_col(j + 2)] + g[2][bd_col(j + 2)]), g[0][bd_col(j + 3)] + 2.f * (g[1][bd_col(j + 3)] + g[2][bd_col(j + 3)]));
            } else {
                const int n1 = nb[r][1], n2 = nb[r][2], n3 = nb[r][3], n4 = nb[r][4];
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float s13[4], s24[4];
                    if (signs != nullptr) {                                 // kept by the forward kernel: 2 x 4 bytes instead of 4 x 16
                        const char4 a13 = *reinterpret_cast<const char4*>(signs + (size_t)m * Cin + c0 + j);
                        const char4 a24 = *reinterpret_cast<const char4*>(signs + ((size_t)M + m) * Cin + c0 + j);
                        s13[0] = (float)a13.x; s13[1] = (float)a13.y; s13[2] = (float)a13.z; s13[3] = (float)a13.w;
                        s24[0] = (float)a24.x; s24[1] = (float)a24.y; s24[2] = (float)a24.z; s24[3] = (float)a24.w;
                    } else {
                        const float4 f1 = (n1 >= 0 && !(dbg & 32)) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)n1 * Cin + c0 + j)) : z;
                        const float4 f2 = (n2 >= 0 && !(dbg & 32)) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)n2 * Cin + c0 + j)) : z;
                        const float4 f3 = (n3 >= 0 && !(dbg & 32)) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)n3 * Cin + c0 + j)) : z;
                        const float4 f4 = (n4 >= 0 && !(dbg & 32)) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)n4 * Cin + c0 + j)) : z;
                        s13[0] = sgn(f1.x - f3.x); s13[1] = sgn(f1.y - f3.y); s13[2] = sgn(f1.z - f3.z); s13[3] = sgn(f1.w - f3.w);
                        s24[0] = sgn(f2.x - f4.x); s24[1] = sgn(f2.y - f4.y); s24[2] = sgn(f2.z - f4.z); s24[3] = sgn(f2.w - f4.w);
                    }
                    tc::red_add_v4(dx + (size_t)m * Cin + c0 + j, g[0][bd_col(j)], g[0][bd_col(j + 1)], g[0][bd_col(j + 2)], g[0][bd_col(j + 3)]);
                    if (n1 >= 0) tc::red_add_v4(dx + (size_t)n1 * Cin + c0 + j, g[1][bd_col(j)] + s13[0] * g[3][bd_col(j)], g[1][bd_col(j + 1)] + s13[1] * g[3][bd_col(j + 1)],
                                                g[1][bd_col(j + 2)] + s13[2] * g[3][bd_col(j + 2)], g[1][bd_col(j + 3)] + s13[3] * g[3][bd_col(j + 3)]);
                    if (n3 >= 0) tc::red_add_v4(dx + (size_t)n3 * Cin + c0 + j, g[1][bd_col(j)] - s13[0] * g[3][bd_col(j)], g[1][bd_col(j + 1)] - s13[1] * g[3][bd_col(j + 1)],
                                                g[1][bd_col(j + 2)] - s13[2] * g[3][bd_col(j + 2)], g[1][bd_col(j + 3)] - s13[3] * g[3][bd_col(j + 3)]);
                    if (n2 >= 0) tc::red_add_v4(dx + (size_t)n2 * Cin + c0 + j, g[2][bd_col(j)] + s24[0] * g[4][bd_col(j)], g[2][bd_col(j + 1)] + s24[1] * g[4][bd_col(j + 1)],
                                                g[2][bd_col(j + 2)] + s24[2] * g[4][bd_col(j + 2)], g[2][bd_col(j + 3)] + s24[3] * g[4][bd_col(j + 3)]);
                    if (n4 >= 0) tc::red_add_v4(dx + (size_t)n4 * Cin + c0 + j, g[2][bd_col(j)] - s24[0] * g[4][bd_col(j)], g[2][bd_col(j + 1)] - s24[1] * g[4][bd_col(j + 1)],
                                                g[2][bd_col(j + 2)] - s24[2] * g[4][bd_col(j + 2)], g[2][bd_col(j + 3)] - s24[3] * g[4][bd_col(j + 3)]);
                }
            }
        }
        tc::tc_fence_before();
    } else {
        constexpr uint32_t idesc = tc::make_idesc_tf32(BD_N);
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
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t dah = tc::make_desc_sw128(a_hi + kk * 32), dal = tc::make_desc_sw128(a_lo + kk * 32);
                    const uint64_t dbh = tc::make_desc_sw128(b_hi + kk * 32), dbl = tc::make_desc_sw128(b_lo + kk * 32);
                    tc::mma_tf32(tmem_base, dah, dbh, idesc, (kb | kk) != 0);
                    tc::mma_tf32(tmem_base, dah, dbl, idesc, 1);
                    tc::mma_tf32(tmem_base, dal, dbh, idesc, 1);
                }
                tc::mma_commit(empty_bar + s);
                if (kb == num_kb - 1) tc::mma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS) {
        tc::tc_fence_after();
        tc::tmem_dealloc<256>(tmem_base);
    }
}

// =============================================================================================================
// backward, data, persistent form.  The kernel above spends its time in three places that use different parts of the SM
// and the chip -- per-CTA set-up, a short main loop (2-8 K blocks on the tensor core), and a long epilogue (sign loads,
// then 5 x 128 x 32 fp32 atomics through L2) -- strictly one after the other, 1200 times per launch.  Here <= one CTA per
// SM stays resident and walks a contiguous range of (128-edge tile, 32-channel block) items with dedicated warps:
//   warps 0-7  producers: dout rows of the item's K blocks -> hi/lo split -> st.shared (register prefetch, distance 2)
//   warp  8    MMA issuer; the accumulator (160 TMEM columns) is double-buffered
//   warps 9-16 epilogue: tcgen05.ld -> abs/add backward with the kept signs -> red.global.add.v4 scatter
// so the scatter of item i runs under the main loop of item i + 1, and set-up is paid once per CTA.
// =============================================================================================================
constexpr int BDP_PRODUCER_WARPS = 8;
constexpr int BDP_EPI_WARPS = 8;
constexpr int BDP_THREADS = (BDP_PRODUCER_WARPS + 1 + BDP_EPI_WARPS) * 32;

// Shared memory holds only the B (weight) tiles: the A operand lives in tensor memory.  TMEM columns (512 allocated):
//   accumulators  [0,160) and [256,416);   A stage s: hi / lo = 32 columns each at  s0: 160 / 192,  s1: 416 / 448,  s2: 224 / 480
struct BdpSmem {
    static constexpr int B_BYTES = BD_N * 128;
    static constexpr int STAGE_BYTES = 2 * B_BYTES;                   // 40960
    static constexpr int STAGES = 3;                                  // = A stages that fit in TMEM next to two accumulators
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;
};
__device__ __forceinline__ uint32_t bdp_a_hi_col(uint32_t s) { return s == 0 ? 160u : (s == 1 ? 416u : 224u); }
__device__ __forceinline__ uint32_t bdp_a_lo_col(uint32_t s) { return s == 0 ? 192u : (s == 1 ? 448u : 480u); }

__global__ void __launch_bounds__(BDP_THREADS, 1)
meshconv_bwd_data_tc_p_kernel(const float* __restrict__ dout, const int32_t* __restrict__ gemm, const int32_t* __restrict__ ec,
                              const float* __restrict__ wtb, float* __restrict__ dx, const int8_t* __restrict__ signs,
                              int B, int E, int Cin, int Cout, int one_pass) {
    pdl_enter();
    using S = BdpSmem;
    constexpr int PW = BDP_PRODUCER_WARPS, PT = PW * 32, ET = BDP_EPI_WARPS * 32;
    static_assert(PW == 8, "two producer warps per TMEM lane quarter");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* acc_full = empty_bar + S::STAGES;              // [2]
    uint64_t* acc_empty = acc_full + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int num_kb = Cout >> 5, ncb = Cin >> 5;
    const long long n_items_all = (long long)ceil_div(M, TC_M) * ncb;
    const int G = gridDim.x, cta = blockIdx.x;
    const int item0 = (int)(n_items_all * cta / G), item1 = (int)(n_items_all * (cta + 1) / G);
    const int dbg = g_tc_debug;            // diagnostics (results wrong): 1 no dout loads, 2 no MMA, 4 no TMEM stores, 8 no B copies, 16 no scatter, 32 no sign loads

    if (warp == PW) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, PW / 2); tc::mbar_init(empty_bar + s, 1); }
            for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, BDP_EPI_WARPS); }
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<512>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

#ifdef MCNN_TC_PROF
#define BDP_ACC(cond, idx, expr) do { const long long t0_ = clock64(); expr; if ((cond) && blockIdx.x == g_sk_prof_cta) g_sk_prof[(idx)] += clock64() - t0_; } while (0)
#else
#define BDP_ACC(cond, idx, expr) do { expr; } while (0)
#endif
    const long long t_begin = clock64();
    if (warp < PW) {
        // ======================================== producers ========================================================
        // Two groups of four warps take alternate (item, K block) steps, so the serial chain of a step (wait for the stage,
        // split, tcgen05.st, wait::st, hand-over) may last two step periods; the gathers of a group's next step fly while
        // the other group works.  Inside a group warp w owns the TMEM lane quarter w % 4 (32 accumulator rows, as two 16-lane
        // halves); thread t holds rows t/4 and t/4 + 8 of each half and, per K block, the 8 consecutive output channels
        // 8 (t%4) .. +7 of each row (two 16-byte loads per row; a quad reads one 128-byte line) -- exactly the registers
        // tcgen05.st .16x256b.x4 wants under the bd_kperm K order.
        const int n_q = (item1 - item0) * num_kb;                // (item, K block) steps of this CTA
        const int grp = warp >> 2;
        const int row0 = (warp & 3) * 32 + (lane >> 2);          // rows row0 + {0, 8, 16, 24}
        const int q8 = (lane & 3) << 3;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        struct Regs { float4 a[8]; };                            // row k (0..3): channels 0-3 in a[2k], 4-7 in a[2k+1]
        // (item, K block) of the step being loaded / stored are carried along instead of divided out of q: an integer
        // division by a run-time value is a ~100-cycle dependent chain, and a producer warp has its scheduler to itself
        struct Cursor {
            int m0, cb, kb;
            __device__ __forceinline__ void next(int num_kb, int ncb) {
                if (++kb == num_kb) { kb = 0; if (++cb == ncb) { cb = 0; m0 += TC_M; } }
            }
        };
        Cursor lc;                                               // this group's next step to load; the store cursor follows it
        lc.m0 = (item0 / ncb) * TC_M; lc.cb = item0 % ncb; lc.kb = 0;
        if (grp) lc.next(num_kb, ncb);
        Cursor sc = lc;
        uint32_t stage = (uint32_t)grp, phase = 0;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_step = [&](Regs& R) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int m = lc.m0 + row0 + 8 * k;
                const float* p = dout + (size_t)m * Cout + (lc.kb << 5) + q8;
                const bool ok = m < M && !(dbg & 1);
                R.a[2 * k] = ok ? __ldg(reinterpret_cast<const float4*>(p)) : zero4;
                R.a[2 * k + 1] = ok ? __ldg(reinterpret_cast<const float4*>(p + 4)) : zero4;
            }
            lc.next(num_kb, ncb); lc.next(num_kb, ncb);
        };
        auto store_step = [&](const Regs& R) {
            const uint32_t s = stage, ph = phase;
            stage += 2;
            if (stage >= (uint32_t)S::STAGES) { stage -= S::STAGES; phase ^= 1; }
            BDP_ACC(tid == 0, 1, if (lane == 0) tc::mbar_wait(empty_bar + s, ph ^ 1); __syncwarp());
            tc::tc_fence_after();
            if ((tid & 127) == 0 && !(dbg & 8)) {                         // B tile: two bulk copies of the pre-built image
                unsigned char* b_hi = smem + s * S::STAGE_BYTES;
                tc::mbar_expect_tx(full_bar + s, 2 * S::B_BYTES);
                const float* src = wtb + ((size_t)sc.cb * num_kb + sc.kb) * (BD_N * 32);
                tc::bulk_g2s(b_hi, src, S::B_BYTES, full_bar + s);
                tc::bulk_g2s(b_hi + S::B_BYTES, src + (size_t)5 * Cin * Cout, S::B_BYTES, full_bar + s);
            }
            sc.next(num_kb, ncb); sc.next(num_kb, ncb);
            if (!(dbg & 4)) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // fragment registers: v[4i + {0,1}] = row (2h) channels 2i, 2i+1;  v[4i + {2,3}] = row (2h + 1), the same channels
                    const float4 a0 = R.a[4 * h], a1 = R.a[4 * h + 1], b0 = R.a[4 * h + 2], b1 = R.a[4 * h + 3];
                    const float fa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float fb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    float hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        tc::split_tf32(fa[2 * i], hi[4 * i], lo[4 * i]);
                        tc::split_tf32(fa[2 * i + 1], hi[4 * i + 1], lo[4 * i + 1]);
                        tc::split_tf32(fb[2 * i], hi[4 * i + 2], lo[4 * i + 2]);
                        tc::split_tf32(fb[2 * i + 1], hi[4 * i + 3], lo[4 * i + 3]);
                    }
                    const uint32_t lf = lane_field + ((uint32_t)(16 * h) << 16);
                    tc::tmem_st_16x256b_x4(tmem_base + lf + bdp_a_hi_col(s), hi);
                    if (!one_pass) tc::tmem_st_16x256b_x4(tmem_base + lf + bdp_a_lo_col(s), lo);
                }
                tc::tmem_wait_st();
            }
            BDP_ACC(tid == 0, 2, tc::tc_fence_before(); __syncwarp(); if (lane == 0) tc::mbar_arrive(full_bar + s));
        };
        Regs R;
        if (grp < n_q) load_step(R);
        for (int q = grp; q < n_q; q += 2) {
            store_step(R);
            if (q + 2 < n_q) load_step(R);
        }
#ifdef MCNN_TC_PROF
        if (tid == 0 && blockIdx.x == g_sk_prof_cta) { g_sk_prof[0] += clock64() - t_begin; g_sk_prof[3] += n_q; }
#endif
    } else if (warp == PW) {
        // ======================================== MMA issuer ========================================================
        constexpr uint32_t idesc = tc::make_idesc_tf32(BD_N);
        uint32_t stage = 0, phase = 0;
        for (int item = item0; item < item1; ++item) {
            const int i = item - item0, buf = i & 1;
            const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
            if (i >= 2) {
                BDP_ACC(lane == 0, 5, tc::mbar_wait(acc_empty + buf, ((i >> 1) - 1) & 1));
                tc::tc_fence_after();
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                const uint32_t s = stage, ph = phase;
                if (++stage == S::STAGES) { stage = 0; phase ^= 1; }
                BDP_ACC(lane == 0, 6, tc::mbar_wait(full_bar + s, ph));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_hi = tmem_base + bdp_a_hi_col(s), a_lo = tmem_base + bdp_a_lo_col(s);
                    const uint32_t b_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
                    const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                    for (int kk = 0; kk < ((dbg & 2) ? 0 : 4); ++kk) {
                        const uint64_t dbh = tc::make_desc_sw128(b_hi + kk * 32), dbl = tc::make_desc_sw128(b_lo + kk * 32);
                        tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbh, idesc, (kb | kk) != 0);
                        if (!one_pass) {
                            tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbl, idesc, 1);
                            tc::mma_tf32_ts(acc, a_lo + 8 * kk, dbh, idesc, 1);
                        }
                    }
                    tc::mma_commit(empty_bar + s);
                    if (kb == num_kb - 1) tc::mma_commit(acc_full + buf);
                }
                __syncwarp();
            }
        }
#ifdef MCNN_TC_PROF
        if (lane == 0 && blockIdx.x == g_sk_prof_cta) g_sk_prof[4] += clock64() - t_begin;
#endif
        tc::tc_fence_before();
    } else {
        // ======================================== epilogue: abs/add backward + scatter =========
        // Warp (lane quarter lq, channel half hb) reads its accumulator lanes in two passes of 16 with the mma-fragment
        // load: thread t holds rows rsub = t/4 and rsub + 8 of the pass, channels c0 .. c0+3 with c0 = 16 hb + 4 (t%4).
        const int lq = warp & 3, hb = (warp - (PW + 1)) >> 2;
        const int rsub = lane >> 2, q4 = lane & 3;
        int mt = item0 / ncb, cb = item0 % ncb;
        int cur_mt = -1;
        int tgt[4][5];                                           // per row (pass, +0 / +8): dx row of self, n1, n2, n3, n4 (-1: none)
        bool pad[4];
        for (int item = item0; item < item1; ++item, cb = (cb + 1 == ncb) ? 0 : cb + 1, mt += (cb == 0)) {
            const int i = item - item0, buf = i & 1;
            const int m0 = mt * TC_M;
            if (mt != cur_mt) {                                  // scatter targets of this thread's four edges (once per edge tile)
                cur_mt = mt;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int m = m0 + lq * 32 + (rr >> 1) * 16 + rsub + (rr & 1) * 8;
                    tgt[rr][0] = tgt[rr][1] = tgt[rr][2] = tgt[rr][3] = tgt[rr][4] = -1; pad[rr] = false;
                    if (m < M) {
                        const int b = m / E, e = m - b * E;
                        if (e >= __ldg(ec + b)) { pad[rr] = true; tgt[rr][0] = b * E; }
                        else {
                            const int4 gq = __ldg(reinterpret_cast<const int4*>(gemm) + m);
                            tgt[rr][0] = m;
                            tgt[rr][1] = gq.x < 0 ? -1 : b * E + gq.x; tgt[rr][2] = gq.y < 0 ? -1 : b * E + gq.y;
                            tgt[rr][3] = gq.z < 0 ? -1 : b * E + gq.z; tgt[rr][4] = gq.w < 0 ? -1 : b * E + gq.w;
                        }
                    }
                }
            }
            const int c0 = cb * 32 + hb * 16 + 4 * q4;
            // the signs do not depend on the accumulator: fetch them while the tensor core still works on this item
            uint32_t sg13[4], sg24[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                sg13[rr] = sg24[rr] = 0;
                if (tgt[rr][0] >= 0 && !pad[rr] && !(dbg & 32)) {
                    sg13[rr] = __ldg(reinterpret_cast<const uint32_t*>(signs + (size_t)tgt[rr][0] * Cin + c0));
                    sg24[rr] = __ldg(reinterpret_cast<const uint32_t*>(signs + ((size_t)M + tgt[rr][0]) * Cin + c0));
                }
            }
            BDP_ACC(lane == 0 && warp == PW + 1, 8, tc::mbar_wait(acc_full + buf, (i >> 1) & 1));
            tc::tc_fence_after();
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                float g[5][8];
                const uint32_t acc = tmem_base + ((uint32_t)(lq * 32 + pass * 16) << 16) + (uint32_t)(buf * 256 + hb * 16);
#pragma unroll
                for (int k5 = 0; k5 < 5; ++k5) tc::tmem_ld_16x256b_x2_nowait(acc + (uint32_t)(k5 * 32), g[k5]);
                tc::tmem_wait_ld();
                if (pass == 1) {                                 // accumulator drained: the tensor core may start item i + 2
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(acc_empty + buf);
                }
                if (dbg & 16) continue;
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2) {
                    const int rr = pass * 2 + j2;
                    if (tgt[rr][0] < 0) continue;
                    // channels c0 .. c0+3 of this row: registers (0, 1, 4, 5) for the row at +0, (2, 3, 6, 7) for the row at +8
                    float v[5][4];
#pragma unroll
                    for (int k5 = 0; k5 < 5; ++k5) { v[k5][0] = g[k5][2 * j2]; v[k5][1] = g[k5][2 * j2 + 1]; v[k5][2] = g[k5][4 + 2 * j2]; v[k5][3] = g[k5][5 + 2 * j2]; }
                    float* d0 = dx + c0;
                    if (pad[rr]) {
                        tc::red_add_v4(d0 + (size_t)tgt[rr][0] * Cin, v[0][0] + 2.f * (v[1][0] + v[2][0]), v[0][1] + 2.f * (v[1][1] + v[2][1]),
                                       v[0][2] + 2.f * (v[1][2] + v[2][2]), v[0][3] + 2.f * (v[1][3] + v[2][3]));
                        continue;
                    }
                    float s13[4], s24[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        s13[t] = (float)(int)(signed char)((sg13[rr] >> (8 * t)) & 0xff);
                        s24[t] = (float)(int)(signed char)((sg24[rr] >> (8 * t)) & 0xff);
                    }
                    tc::red_add_v4(d0 + (size_t)tgt[rr][0] * Cin, v[0][0], v[0][1], v[0][2], v[0][3]);
                    if (tgt[rr][1] >= 0) tc::red_add_v4(d0 + (size_t)tgt[rr][1] * Cin, v[1][0] + s13[0] * v[3][0], v[1][1] + s13[1] * v[3][1],
                                                        v[1][2] + s13[2] * v[3][2], v[1][3] + s13[3] * v[3][3]);
                    if (tgt[rr][3] >= 0) tc::red_add_v4(d0 + (size_t)tgt[rr][3] * Cin, v[1][0] - s13[0] * v[3][0], v[1][1] - s13[1] * v[3][1],
                                                        v[1][2] - s13[2] * v[3][2], v[1][3] - s13[3] * v[3][3]);
                    if (tgt[rr][2] >= 0) tc::red_add_v4(d0 + (size_t)tgt[rr][2] * Cin, v[2][0] + s24[0] * v[4][0], v[2][1] + s24[1] * v[4][1],
                                                        v[2][2] + s24[2] * v[4][2], v[2][3] + s24[3] * v[4][3]);
                    if (tgt[rr][4] >= 0) tc::red_add_v4(d0 + (size_t)tgt[rr][4] * Cin, v[2][0] - s24[0] * v[4][0], v[2][1] - s24[1] * v[4][1],
                                                        v[2][2] - s24[2] * v[4][2], v[2][3] - s24[3] * v[4][3]);
                }
            }
        }
#ifdef MCNN_TC_PROF
        if (lane == 0 && warp == PW + 1 && blockIdx.x == g_sk_prof_cta) g_sk_prof[7] += clock64() - t_begin;
#endif
        (void)ET;
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == PW) {
        tc::tc_fence_after();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

// =============================================================================================================
// backward, weight:  dW[o][c][k5] = sum_m dout[m][o] * G_k5[m][c]   (reduction over all edge slots, padded ones included)
// One CTA = 128 output channels x (5 x 32 input channels) over a slab of edges; K = edges.  Both operands are
// MN-major exactly as they are produced: A[k=edge][o] = dout rows, B[k=edge][(k5, cl)] = G rows.  Slab partials go
// to a workspace and are reduced by a second kernel (deterministic, no atomics).
// =============================================================================================================
// OT = output-channel tiles of 128 per CTA.  With two, the gathered operand (five row loads, symmetric functions, five
// splits and ten 16-byte stores per thread and K block -- the expensive one: this kernel is bound by operand PRODUCTION)
// is built once for 256 output channels instead of once per 128.
template <int OT>
struct BwSmem {
    static constexpr int A_BYTES = 32 * 128 * 4;            // 32 edges x 128 o, per output tile
    static constexpr int B_BYTES = 32 * BD_N * 4;            // 32 edges x 160 cols
    static constexpr int STAGE_BYTES = OT * 2 * A_BYTES + 2 * B_BYTES;     // 73728 / 106496
    static constexpr int STAGES = OT == 1 ? 3 : 2;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 16 * 8 + 16 + 1024;
    static constexpr uint32_t LBO = 4096;                    // 4 K-groups of 1024 bytes per M/N atom
    static constexpr int TMEM_COLS = OT == 1 ? 256 : 512;    // accumulator t at column 256 t
};

template <int OT>
__global__ void __launch_bounds__(TC_THREADS, 1)
meshconv_bwd_weight_tc_kernel(const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ gemm,
                              const int32_t* __restrict__ ec, float* __restrict__ ws, int B, int E, int Cin, int Cout,
                              int rows_per_split) {
    pdl_enter();
    using S = BwSmem<OT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_bar = full_bar + S::STAGES;
    uint64_t* accum_bar = empty_bar + S::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int split = blockIdx.x, o0 = blockIdx.y * (128 * OT), cb = blockIdx.z;
    const int mbeg = split * rows_per_split, mend = min(M, mbeg + rows_per_split);
    const int num_kb = (mend > mbeg) ? (mend - mbeg + 31) >> 5 : 0;
    const int dbg = g_tc_debug;            // diagnostics (results wrong): 1 no global loads, 2 no MMA, 4 no st.shared

    if (warp == TC_PRODUCER_WARPS) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_bar + s, TC_PRODUCER_WARPS); tc::mbar_init(empty_bar + s, 1); }
            tc::mbar_init(accum_bar, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<S::TMEM_COLS>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PRODUCER_WARPS) {
        const int pt = tid;
        // Software pipeline: the gather is two dependent global loads (one-ring ids, then five rows), so the ids run three K
        // blocks ahead and the rows / dout chunks two; every load is issued right after a hand-over (never just before the
        // proxy fence of a store, which would wait for it).
        struct Idx { int s0, s1, s2, s3, s4; };
        struct Data { float4 a[4 * OT]; float4 f[5]; };
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const int bk = pt >> 3, bq = pt & 7;                                // this thread's edge-in-block and channel chunk (B operand)
        // load_idx is called for kb = 0, 1, 2, ... in order: the (mesh, edge) of this thread's row is carried along instead
        // of divided out of m every K block (an integer division by a run-time E is a ~100-cycle dependent chain)
        int nb_b = (mbeg + bk) / E, nb_e = (mbeg + bk) - nb_b * E;
        auto load_idx = [&](int kb) {
            Idx I; I.s0 = I.s1 = I.s2 = I.s3 = I.s4 = -1;
            const int m = mbeg + (kb << 5) + bk;
            const int b = nb_b, e = nb_e;
            nb_e += 32;
            while (nb_e >= E) { nb_e -= E; ++nb_b; }
            if (m < mend) {
                if (e >= __ldg(ec + b)) { I.s0 = I.s1 = I.s2 = I.s3 = I.s4 = b * E; }
                else {
                    const int4 g4 = __ldg(reinterpret_cast<const int4*>(gemm + (size_t)m * 4));
                    I.s0 = m;
                    I.s1 = g4.x < 0 ? -1 : b * E + g4.x; I.s2 = g4.y < 0 ? -1 : b * E + g4.y;
                    I.s3 = g4.z < 0 ? -1 : b * E + g4.z; I.s4 = g4.w < 0 ? -1 : b * E + g4.w;
                }
            }
            return I;
        };
        auto load_data = [&](int kb, const Idx& I) {
            Data D;
            if (dbg & 1) {
#pragma unroll
                for (int it = 0; it < 4 * OT; ++it) D.a[it] = z;
#pragma unroll
                for (int it = 0; it < 5; ++it) D.f[it] = z;
                return D;
            }
            const int mb = mbeg + (kb << 5);
#pragma unroll
            for (int it = 0; it < 4 * OT; ++it) {                           // A: dout[m][o0 + 4q ..] for 32 edges x 32 OT chunks
                const int item = pt + it * 256;
                const int m = mb + item / (32 * OT), o = o0 + ((item % (32 * OT)) << 2);
                D.a[it] = (m < mend && o < Cout) ? __ldg(reinterpret_cast<const float4*>(dout + (size_t)m * Cout + o)) : z;
            }
            const int c = cb * 32 + (bq << 2);
            D.f[0] = I.s0 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s0 * Cin + c)) : z;
            D.f[1] = I.s1 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s1 * Cin + c)) : z;
            D.f[2] = I.s2 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s2 * Cin + c)) : z;
            D.f[3] = I.s3 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s3 * Cin + c)) : z;
            D.f[4] = I.s4 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s4 * Cin + c)) : z;
            return D;
        };
        auto store_block = [&](int kb, const Data& D) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            if (lane == 0) tc::mbar_wait(empty_bar + s, ph ^ 1);      // one poller per warp
            __syncwarp();
            const uint32_t a_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);     // per output tile t: hi at + 2 t A_BYTES, lo behind it
            const uint32_t b_hi = a_hi + OT * 2 * S::A_BYTES;
            const uint32_t b_lo = b_hi + S::B_BYTES;
            if (dbg & 4) {
                tc::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(full_bar + s);
                return;
            }
#pragma unroll
            for (int it = 0; it < 4 * OT; ++it) {
                const int item = pt + it * 256;
                float4 hi, lo;
                tc::split4(D.a[it], hi, lo);
                const int ch = item % (32 * OT);
                const uint32_t off = (uint32_t)((ch >> 5) * 2 * S::A_BYTES) + tc::sw128_mn_off(item / (32 * OT), ch & 31, S::LBO);
                tc::sts128(a_hi + off, hi);
                tc::sts128(a_hi + S::A_BYTES + off, lo);
            }
            const float4 f0 = D.f[0], f1 = D.f[1], f2 = D.f[2], f3 = D.f[3], f4 = D.f[4];
            float4 gk[5];
            gk[0] = f0;
            gk[1] = make_float4(f1.x + f3.x, f1.y + f3.y, f1.z + f3.z, f1.w + f3.w);
            gk[2] = make_float4(f2.x + f4.x, f2.y + f4.y, f2.z + f4.z, f2.w + f4.w);
            gk[3] = make_float4(fabsf(f1.x - f3.x), fabsf(f1.y - f3.y), fabsf(f1.z - f3.z), fabsf(f1.w - f3.w));
            gk[4] = make_float4(fabsf(f2.x - f4.x), fabsf(f2.y - f4.y), fabsf(f2.z - f4.z), fabsf(f2.w - f4.w));
#pragma unroll
            for (int k5 = 0; k5 < 5; ++k5) {
                float4 hi, lo;
                tc::split4(gk[k5], hi, lo);
                const uint32_t off = tc::sw128_mn_off(bk, k5 * 8 + bq, S::LBO);
                tc::sts128(b_hi + off, hi);
                tc::sts128(b_lo + off, lo);
            }
            tc::fence_proxy_async_smem();                                  // every writer fences its own stores ...
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(full_bar + s);                  // ... then one arrival per warp
        };
        if (num_kb > 0) {
            Data D0 = load_data(0, load_idx(0));
            Data D1;
            if (num_kb > 1) D1 = load_data(1, load_idx(1));
            Idx In;
            if (num_kb > 2) In = load_idx(2);
            for (int kb = 0; kb < num_kb; kb += 2) {
                store_block(kb, D0);
                if (kb + 2 < num_kb) D0 = load_data(kb + 2, In);
                if (kb + 3 < num_kb) In = load_idx(kb + 3);
                if (kb + 1 < num_kb) {
                    store_block(kb + 1, D1);
                    if (kb + 3 < num_kb) D1 = load_data(kb + 3, In);
                    if (kb + 4 < num_kb) In = load_idx(kb + 4);
                }
            }
        }
        // ---- epilogue: slab partial -> ws[split][o][c][k5] ---------------------------------------------------------
        const int lane_grp = warp & 3;
        const int cl0 = (warp >> 2) * 16;
        if (num_kb > 0) {
            tc::mbar_wait(accum_bar, 0);
            tc::tc_fence_after();
        }
#pragma unroll 1
        for (int t = 0; t < OT; ++t) {
            const int o = o0 + t * 128 + lane_grp * 32 + lane;
            float g[5][16];
            if (num_kb > 0) {
#pragma unroll
                for (int k5 = 0; k5 < 5; ++k5) tc::tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * 256 + k5 * 32 + cl0), g[k5]);
            } else {
#pragma unroll
                for (int k5 = 0; k5 < 5; ++k5)
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[k5][j] = 0.f;
            }
            if (o < Cout) {
                float* dst = ws + (((size_t)split * Cout + o) * Cin + cb * 32 + cl0) * 5;      // 80 contiguous floats
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    // (c, k5) interleave: 4 channels x 5 taps = 20 floats = 5 float4
                    float tt[20];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int k5 = 0; k5 < 5; ++k5) tt[jj * 5 + k5] = g[k5][j + jj];
#pragma unroll
                    for (int v = 0; v < 5; ++v)
                        *reinterpret_cast<float4*>(dst + j * 5 + v * 4) = make_float4(tt[v * 4], tt[v * 4 + 1], tt[v * 4 + 2], tt[v * 4 + 3]);
                }
            }
        }
        tc::tc_fence_before();
    } else {
        constexpr uint32_t idesc = tc::make_idesc_tf32_mn(BD_N);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            tc::mbar_wait(full_bar + s, ph);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t a0 = tc::smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t b_hi = a0 + OT * 2 * S::A_BYTES;
                const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
                for (int kk = 0; kk < ((dbg & 2) ? 0 : 4); ++kk) {                 // one 8-edge K group per MMA: +1024 bytes
                    const uint64_t dbh = tc::make_desc_sw128_mn(b_hi + kk * 1024, S::LBO), dbl = tc::make_desc_sw128_mn(b_lo + kk * 1024, S::LBO);
#pragma unroll
                    for (int t = 0; t < OT; ++t) {
                        const uint32_t a_hi = a0 + t * 2 * S::A_BYTES, a_lo = a_hi + S::A_BYTES;
                        const uint64_t dah = tc::make_desc_sw128_mn(a_hi + kk * 1024, S::LBO), dal = tc::make_desc_sw128_mn(a_lo + kk * 1024, S::LBO);
                        tc::mma_tf32(tmem_base + (uint32_t)(t * 256), dah, dbh, idesc, (kb | kk) != 0);
                        tc::mma_tf32(tmem_base + (uint32_t)(t * 256), dah, dbl, idesc, 1);
                        tc::mma_tf32(tmem_base + (uint32_t)(t * 256), dal, dbh, idesc, 1);
                    }
                }
                tc::mma_commit(empty_bar + s);
                if (kb == num_kb - 1) tc::mma_commit(accum_bar);
            }
            __syncwarp();
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == TC_PRODUCER_WARPS) {
        tc::tc_fence_after();
        tc::tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

// -------------------------------------------------------------------------------------------------------------
// The same contraction with the A operand (dout, transposed: rows = output channels, K = edges) in TENSOR MEMORY.
// The kernel above is bound by the shared-memory port: per 32-edge K block 256 threads write 106 KB of operands (A and B,
// hi and lo) and the MMAs read 216 KB back.  dout^T is exactly what a TMEM A operand wants -- lane = output channel,
// column = edge -- so each producer thread owns ONE output channel, reads its 32 values of the K block (a warp instruction
// reads 32 consecutive channels of one edge: a coalesced 128-byte line), splits them and writes them into its own TMEM
// lane with tcgen05.st.32x32b: no transpose, no shared memory, no proxy fence for A.  Shared memory holds only the
// gathered G operand (40 KB per stage, four stages), built once per K block for all 128 * OT output channels.
// TMEM (512 columns): accumulators [0,160) and [256,416); three A slots of 32 + 32 columns (hi, lo) in the gaps, the
// column plan of the persistent data-gradient kernel.  A slot holds one (K block, output tile) unit; units go round the
// three slots, so A production runs up to 1.5 K blocks ahead of the tensor core.
// -------------------------------------------------------------------------------------------------------------
template <int OT>
struct BwaSmem {
    static constexpr int B_BYTES = 32 * BD_N * 4;            // 32 edges x 160 cols
    static constexpr int STAGE_BYTES = 2 * B_BYTES;          // hi + lo = 40960
    static constexpr int STAGES = 4;
    static constexpr int A_SLOTS = 3;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 32 * 8 + 16 + 1024;
    static constexpr uint32_t LBO = 4096;
};

// Warp roles (17 warps): 0-7 feed the A operand (warp w: output tile w / 4 [OT == 2] or edge half w / 4 [OT == 1], TMEM lane
// quarter w % 4) and run the epilogue; 8-15 build the gathered B operand in shared memory; 16 issues the MMAs.  A producer
// thread retires roughly one dependent instruction per 5 cycles, so the per-K-block chain of a warp that did BOTH operands
// (~250 instructions, two such warps per scheduler) was the kernel's period, not the tensor pipe (51 % active): two
// specialised warp sets halve the chain and put four warps on every scheduler.
constexpr int BWA_A_WARPS = 8, BWA_B_WARPS = 8;
constexpr int BWA_THREADS = (BWA_A_WARPS + BWA_B_WARPS + 1) * 32;

template <int OT>
__global__ void __launch_bounds__(BWA_THREADS, 1)
meshconv_bwd_weight_tca_kernel(const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ gemm,
                               const int32_t* __restrict__ ec, float* __restrict__ ws, int B, int E, int Cin, int Cout,
                               int rows_per_split, int one_pass) {
    pdl_enter();
    using S = BwaSmem<OT>;
    constexpr int AV = 16 * OT;                              // dout values per A thread and K block (OT == 1: half of the edges)
    constexpr int MMA_WARP = BWA_A_WARPS + BWA_B_WARPS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_b = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* empty_b = full_b + S::STAGES;
    uint64_t* full_a = empty_b + S::STAGES;
    uint64_t* empty_a = full_a + S::A_SLOTS;
    uint64_t* accum_bar = empty_a + S::A_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = B * E;
    const int split = blockIdx.x, o0 = blockIdx.y * (128 * OT), cb = blockIdx.z;
    const int mbeg = split * rows_per_split, mend = min(M, mbeg + rows_per_split);
    const int num_kb = (mend > mbeg) ? (mend - mbeg + 31) >> 5 : 0;

    if (warp == MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(full_b + s, BWA_B_WARPS); tc::mbar_init(empty_b + s, 1); }
            for (int s = 0; s < S::A_SLOTS; ++s) { tc::mbar_init(full_a + s, OT == 2 ? 4 : 8); tc::mbar_init(empty_a + s, 1); }
            tc::mbar_init(accum_bar, 1);
            tc::fence_barrier_init();
        }
        __syncwarp();
        tc::tmem_alloc<512>(tmem_slot);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < BWA_A_WARPS) {
        // ======================================== A producers: dout^T into tensor memory ==============================
        const int a_tile = OT == 2 ? (warp >> 2) : 0;                       // OT == 2: warps 0-3 feed tile 0, warps 4-7 tile 1
        const int a_e0 = OT == 2 ? 0 : (warp >> 2) * 16;                    // OT == 1: the two warp groups split the 32 edges
        const int a_o = o0 + a_tile * 128 + (warp & 3) * 32 + lane;         // this thread's output channel = its TMEM lane
        const uint32_t a_lane = (uint32_t)((warp & 3) * 32) << 16;
        struct AData { float av[AV]; };
        auto load_a = [&](int kb) {
            AData D;
            const int mb = mbeg + (kb << 5) + a_e0;
            const float* ap = dout + (size_t)mb * Cout + a_o;
#pragma unroll
            for (int j = 0; j < AV; ++j)
                D.av[j] = (mb + j < mend && a_o < Cout) ? __ldg(ap + (size_t)j * Cout) : 0.f;
            return D;
        };
        auto store_a = [&](int kb, const AData& D) {
            const int u = kb * OT + a_tile;                                // (K block, output tile) unit -> A slot
            const uint32_t slot = (uint32_t)(u % S::A_SLOTS), ph = (uint32_t)((u / S::A_SLOTS) & 1);
            if (lane == 0) tc::mbar_wait(empty_a + slot, ph ^ 1);
            __syncwarp();
            tc::tc_fence_after();
            const uint32_t c_hi = tmem_base + a_lane + bdp_a_hi_col(slot) + (uint32_t)a_e0;
            const uint32_t c_lo = tmem_base + a_lane + bdp_a_lo_col(slot) + (uint32_t)a_e0;
#pragma unroll
            for (int h = 0; h < AV / 16; ++h) {
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) tc::split_tf32(D.av[16 * h + j], hi[j], lo[j]);
                tc::tmem_st_32x32b_x16(c_hi + 16 * h, hi);
                if (!one_pass) tc::tmem_st_32x32b_x16(c_lo + 16 * h, lo);
            }
            tc::tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(full_a + slot);
        };
        if (num_kb > 0) {
            AData D0 = load_a(0);
            AData D1;
            if (num_kb > 1) D1 = load_a(1);
            for (int kb = 0; kb < num_kb; kb += 2) {
                store_a(kb, D0);
                if (kb + 2 < num_kb) D0 = load_a(kb + 2);
                if (kb + 1 < num_kb) {
                    store_a(kb + 1, D1);
                    if (kb + 3 < num_kb) D1 = load_a(kb + 3);
                }
            }
        }
        // ---- epilogue: slab partial -> ws[split][o][c][k5] ---------------------------------------------------------
        const int lane_grp = warp & 3;
        const int cl0 = (warp >> 2) * 16;
        if (num_kb > 0) {
            tc::mbar_wait(accum_bar, 0);
            tc::tc_fence_after();
        }
#pragma unroll 1
        for (int t = 0; t < OT; ++t) {
            const int o = o0 + t * 128 + lane_grp * 32 + lane;
            float g[5][16];
            if (num_kb > 0) {
#pragma unroll
                for (int k5 = 0; k5 < 5; ++k5) tc::tmem_ld16(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * 256 + k5 * 32 + cl0), g[k5]);
            } else {
#pragma unroll
                for (int k5 = 0; k5 < 5; ++k5)
#pragma unroll
                    for (int j = 0; j < 16; ++j) g[k5][j] = 0.f;
            }
            if (o < Cout) {
                float* dst = ws + (((size_t)split * Cout + o) * Cin + cb * 32 + cl0) * 5;      // 80 contiguous floats
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float tt[20];                                        // (c, k5) interleave: 4 channels x 5 taps = 5 float4
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                        for (int k5 = 0; k5 < 5; ++k5) tt[jj * 5 + k5] = g[k5][j + jj];
#pragma unroll
                    for (int v = 0; v < 5; ++v)
                        *reinterpret_cast<float4*>(dst + j * 5 + v * 4) = make_float4(tt[v * 4], tt[v * 4 + 1], tt[v * 4 + 2], tt[v * 4 + 3]);
                }
            }
        }
        tc::tc_fence_before();
    } else if (warp < MMA_WARP) {
        // ======================================== B producers: the gathered operand into shared memory ==================
        const int pt = tid - BWA_A_WARPS * 32;                              // 0 .. 255
        struct Idx { int s0, s1, s2, s3, s4; };
        struct BData { float4 f[5]; };
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const int bk = pt >> 3, bq = pt & 7;                                // this thread's edge-in-block and channel chunk
        int nb_b = (mbeg + bk) / E, nb_e = (mbeg + bk) - nb_b * E;
        auto load_idx = [&](int kb) {
            Idx I; I.s0 = I.s1 = I.s2 = I.s3 = I.s4 = -1;
            const int m = mbeg + (kb << 5) + bk;
            const int b = nb_b, e = nb_e;
            nb_e += 32;
            while (nb_e >= E) { nb_e -= E; ++nb_b; }
            if (m < mend) {
                if (e >= __ldg(ec + b)) { I.s0 = I.s1 = I.s2 = I.s3 = I.s4 = b * E; }
                else {
                    const int4 g4 = __ldg(reinterpret_cast<const int4*>(gemm + (size_t)m * 4));
                    I.s0 = m;
                    I.s1 = g4.x < 0 ? -1 : b * E + g4.x; I.s2 = g4.y < 0 ? -1 : b * E + g4.y;
                    I.s3 = g4.z < 0 ? -1 : b * E + g4.z; I.s4 = g4.w < 0 ? -1 : b * E + g4.w;
                }
            }
            return I;
        };
        auto load_b = [&](const Idx& I) {
            BData D;
            const int c = cb * 32 + (bq << 2);
            D.f[0] = I.s0 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s0 * Cin + c)) : z;
            D.f[1] = I.s1 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s1 * Cin + c)) : z;
            D.f[2] = I.s2 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s2 * Cin + c)) : z;
            D.f[3] = I.s3 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s3 * Cin + c)) : z;
            D.f[4] = I.s4 >= 0 ? __ldg(reinterpret_cast<const float4*>(x + (size_t)I.s4 * Cin + c)) : z;
            return D;
        };
        auto store_b = [&](int kb, const BData& D) {
            const int s = kb % S::STAGES;
            const uint32_t ph = (kb / S::STAGES) & 1;
            if (lane == 0) tc::mbar_wait(empty_b + s, ph ^ 1);            // one poller per warp
            __syncwarp();
            const uint32_t b_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
            const uint32_t b_lo = b_hi + S::B_BYTES;
            const float4 f0 = D.f[0], f1 = D.f[1], f2 = D.f[2], f3 = D.f[3], f4 = D.f[4];
            float4 gk[5];
            gk[0] = f0;
            gk[1] = make_float4(f1.x + f3.x, f1.y + f3.y, f1.z + f3.z, f1.w + f3.w);
            gk[2] = make_float4(f2.x + f4.x, f2.y + f4.y, f2.z + f4.z, f2.w + f4.w);
            gk[3] = make_float4(fabsf(f1.x - f3.x), fabsf(f1.y - f3.y), fabsf(f1.z - f3.z), fabsf(f1.w - f3.w));
            gk[4] = make_float4(fabsf(f2.x - f4.x), fabsf(f2.y - f4.y), fabsf(f2.z - f4.z), fabsf(f2.w - f4.w));
#pragma unroll
            for (int k5 = 0; k5 < 5; ++k5) {
                float4 hi, lo;
                tc::split4(gk[k5], hi, lo);
                const uint32_t off = tc::sw128_mn_off(bk, k5 * 8 + bq, S::LBO);
                tc::sts128(b_hi + off, hi);
                if (!one_pass) tc::sts128(b_lo + off, lo);
            }
            tc::fence_proxy_async_smem();                                  // every writer fences its own stores ...
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(full_b + s);                    // ... then one arrival per warp
        };
        if (num_kb > 0) {
            // the gather is two dependent global loads (one-ring ids, then five rows): ids run three K blocks ahead, rows two
            BData D0 = load_b(load_idx(0));
            BData D1;
            if (num_kb > 1) D1 = load_b(load_idx(1));
            Idx In;
            if (num_kb > 2) In = load_idx(2);
            for (int kb = 0; kb < num_kb; kb += 2) {
                store_b(kb, D0);
                if (kb + 2 < num_kb) D0 = load_b(In);
                if (kb + 3 < num_kb) In = load_idx(kb + 3);
                if (kb + 1 < num_kb) {
                    store_b(kb + 1, D1);
                    if (kb + 3 < num_kb) D1 = load_b(In);
                    if (kb + 4 < num_kb) In = load_idx(kb + 4);
                }
            }
        }
    } else {
        // ======================================== MMA issuer ============================================================
        constexpr uint32_t idesc = tc::make_idesc_tf32_bmn(BD_N);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % S::STAGES;
            tc::mbar_wait(full_b + s, (uint32_t)((kb / S::STAGES) & 1));
            const uint32_t b_hi = tc::smem_u32(smem + s * S::STAGE_BYTES);
            const uint32_t b_lo = b_hi + S::B_BYTES;
#pragma unroll
            for (int t = 0; t < OT; ++t) {
                const int u = kb * OT + t;
                const uint32_t slot = (uint32_t)(u % S::A_SLOTS);
                tc::mbar_wait(full_a + slot, (uint32_t)((u / S::A_SLOTS) & 1));
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_hi = tmem_base + bdp_a_hi_col(slot), a_lo = tmem_base + bdp_a_lo_col(slot);
                    const uint32_t acc = tmem_base + (uint32_t)(t * 256);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {                     // one 8-edge K group per MMA: +1024 bytes in B, +8 columns in A
                        const uint64_t dbh = tc::make_desc_sw128_mn(b_hi + kk * 1024, S::LBO), dbl = tc::make_desc_sw128_mn(b_lo + kk * 1024, S::LBO);
                        tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbh, idesc, (kb | kk) != 0);
                        if (!one_pass) {
                            tc::mma_tf32_ts(acc, a_hi + 8 * kk, dbl, idesc, 1);
                            tc::mma_tf32_ts(acc, a_lo + 8 * kk, dbh, idesc, 1);
                        }
                    }
                    tc::mma_commit(empty_a + slot);
                    if (t == OT - 1) tc::mma_commit(empty_b + s);
                    if (t == OT - 1 && kb == num_kb - 1) tc::mma_commit(accum_bar);
                }
                __syncwarp();
            }
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc::tc_fence_after();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

// dW[i] = sum_s ws[s][i]
__global__ void reduce_splits_kernel(const float* __restrict__ ws, float* __restrict__ dst, int n, int splits) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(size_t)s * n + i];
    dst[i] = acc;
}

// column sums of dout [M][Cout] in two deterministic stages: part[p][o], then reduce_splits_kernel
constexpr int COLSUM_PARTS = 64;
__global__ void colsum_part_kernel(const float* __restrict__ dout, float* __restrict__ part, int M, int Cout) {
    pdl_enter();
    __shared__ float red[8][33];
    const int o = blockIdx.x * 32 + threadIdx.x;
    const int p = blockIdx.y;
    const int rows = (M + COLSUM_PARTS - 1) / COLSUM_PARTS;
    const int mb = p * rows, me = min(M, mb + rows);
    float acc = 0.f;
    if (o < Cout)
        for (int m = mb + threadIdx.y; m < me; m += 8) acc += __ldg(dout + (size_t)m * Cout + o);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && o < Cout) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
        part[(size_t)p * Cout + o] = t;
    }
}

static int g_tc_bw_ot = 0;                 // debug: force the output tiles per CTA of the weight gradient (0 = heuristic)
static int g_tc_bw_tmem_a = 1;             // debug: 0 = the weight-gradient kernel with both operands in shared memory
static int bw_ot(int Cout) {
    if (g_tc_bw_ot == 1 || g_tc_bw_ot == 2) return g_tc_bw_ot;
    return Cout % 256 == 0 ? 2 : 1;
}

static int bw_splits(int M, int Cin, int Cout, int* rows_per_split) {
    const int tiles = ceil_div(Cout, 128 * bw_ot(Cout)) * (Cin / 32);
    // one CTA per SM (222 KB of shared memory each): fill exactly one wave of the 148 SMs -- a 149th CTA would cost a
    // whole second round
    int S = 148 / tiles;
    const int maxS = max(1, M / 256);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    int rows = ceil_div(ceil_div(M, S), 32) * 32;
    S = ceil_div(M, rows);
    *rows_per_split = rows;
    return S;
}

}  // namespace mcnn

using namespace mcnn;

extern "C" int mcnn_meshconv_tc_supported(int Cin, int Cout) { return (Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 64 == 0) ? 1 : 0; }

extern "C" int mcnn_meshconv_tc_weight_images(const float* weight, float* fwd_img, float* bwd_img, int Cin, int Cout, void* stream) {
    if (!weight || (!fwd_img && !bwd_img) || !(Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 32 == 0)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = Cout * 5 * Cin;
    int rc = MCNN_OK;
    if (fwd_img != nullptr) {
        launch_k(weight_tc_layout_kernel, ceil_div(total, 256), 256, 0, st, weight, fwd_img, Cin, Cout);
        if ((rc = after_launch())) return rc;
    }
    if (bwd_img != nullptr) {
        launch_k(weight_tc_bwd_layout_kernel, ceil_div(total, 256), 256, 0, st, weight, bwd_img, Cin, Cout);
        rc = after_launch();
    }
    return rc;
}

extern "C" int mcnn_meshconv_fwd_tc(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                    const float* bias, float* out, float* wt_ws, int B, int E, int Cin, int Cout,
                                    void* stream) {
    return mcnn_meshconv_fwd_tc_signs(x, gemm, ec, weight, bias, out, wt_ws, nullptr, B, E, Cin, Cout, stream);
}

extern "C" int mcnn_meshconv_fwd_tc_signs(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                          const float* bias, float* out, float* wt_ws, int8_t* signs, int B, int E, int Cin,
                                          int Cout, void* stream) {
    if (!x || !gemm || !ec || !out || !wt_ws || B <= 0 || E <= 0) return MCNN_E_BADARG;
    if (!mcnn_meshconv_tc_supported(Cin, Cout)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = Cout * 5 * Cin;
    int rc = MCNN_OK;
    if (weight != nullptr) {                                 // NULL: wt_ws already holds the image (mcnn_meshconv_tc_weight_images)
        launch_k(weight_tc_layout_kernel, ceil_div(total, 256), 256, 0, st, weight, wt_ws, Cin, Cout);
        if ((rc = after_launch())) return rc;
    }
    const int mtiles = ceil_div(B * E, TC_M);
    int nt = (Cout % 256 == 0) ? 256 : ((Cout % 128 == 0) ? 128 : 64);
    (void)mtiles;
    if (g_tc_force_ntile && Cout % g_tc_force_ntile == 0) nt = g_tc_force_ntile;
    if (nt == 256 && g_tc_pair) return launch_fwd_tc2(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
    if (g_tc_pw16) {                                          // 16 producer warps (the kernel is bound by operand production)
        if (nt == 256) return launch_fwd_tc<256, 16>(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
        if (nt == 128) return launch_fwd_tc<128, 16>(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
    }
    if (nt == 256) return launch_fwd_tc<256, 8>(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
    if (nt == 128) return launch_fwd_tc<128, 8>(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
    return launch_fwd_tc<64, 8>(x, gemm, ec, wt_ws, bias, out, signs, B, E, Cin, Cout, st);
}


/* bytes of the stream-K workspace (flags + one partial tile per CTA); the caller zeroes it ONCE after allocating it */
extern "C" size_t mcnn_meshconv_fwd_tc_sk_ws(void) { return (size_t)SK_FLAG_BYTES + (size_t)2 * SK_MAX_CTAS * TC_M * 256 * sizeof(float); }

extern "C" int mcnn_meshconv_fwd_tc_sk_p(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                         const float* bias, float* out, float* wt_ws, int8_t* signs, void* sk_ws, int B, int E,
                                         int Cin, int Cout, int precision, void* stream) {
    if (precision != MCNN_PREC_FP32 && precision != MCNN_PREC_TF32) return MCNN_E_BADARG;
    const int one_pass = precision == MCNN_PREC_TF32;
    if (!x || !gemm || !ec || !out || !wt_ws || !sk_ws || B <= 0 || E <= 0) return MCNN_E_BADARG;
    if (!mcnn_meshconv_tc_supported(Cin, Cout)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = MCNN_OK;
    if (weight != nullptr) {
        launch_k(weight_tc_layout_kernel, ceil_div(Cout * 5 * Cin, 256), 256, 0, st, weight, wt_ws, Cin, Cout);
        if ((rc = after_launch())) return rc;
    }
    int nt = (Cout % 256 == 0) ? 256 : ((Cout % 128 == 0) ? 128 : 64);
    if (g_tc_force_ntile && Cout % g_tc_force_ntile == 0) nt = g_tc_force_ntile;
    if (nt == 256) return launch_fwd_tc_sk<256>(x, gemm, ec, wt_ws, bias, out, signs, sk_ws, B, E, Cin, Cout, one_pass, st);
    if (nt == 128) return launch_fwd_tc_sk<128>(x, gemm, ec, wt_ws, bias, out, signs, sk_ws, B, E, Cin, Cout, one_pass, st);
    return launch_fwd_tc_sk<64>(x, gemm, ec, wt_ws, bias, out, signs, sk_ws, B, E, Cin, Cout, one_pass, st);
}

extern "C" int mcnn_meshconv_fwd_tc_sk(const float* x, const int32_t* gemm, const int32_t* ec, const float* weight,
                                       const float* bias, float* out, float* wt_ws, int8_t* signs, void* sk_ws, int B, int E,
                                       int Cin, int Cout, void* stream) {
    return mcnn_meshconv_fwd_tc_sk_p(x, gemm, ec, weight, bias, out, wt_ws, signs, sk_ws, B, E, Cin, Cout, MCNN_PREC_FP32, stream);
}

extern "C" int mcnn_debug_set_tc_bw_ot(int ot) { g_tc_bw_ot = (ot == 1 || ot == 2) ? ot : 0; return MCNN_OK; }
extern "C" int mcnn_debug_set_tc_bw_tmem_a(int on) { g_tc_bw_tmem_a = on ? 1 : 0; return MCNN_OK; }

extern "C" int mcnn_debug_set_tc_bd_tile(int on) { g_tc_bd_tile = on ? 1 : 0; return MCNN_OK; }

extern "C" int mcnn_debug_set_tc_sk_ctas(int n) { g_tc_sk_ctas = n > 0 ? (n > SK_MAX_CTAS ? SK_MAX_CTAS : n) : 0; return MCNN_OK; }


extern "C" int mcnn_meshconv_bwd_data_tc(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                         const float* weight, float* dx, float* wtb_ws, int B, int E, int Cin, int Cout,
                                         void* stream) {
    return mcnn_meshconv_bwd_data_tc_signs(dout, x, nullptr, gemm, ec, weight, dx, wtb_ws, B, E, Cin, Cout, stream);
}

extern "C" int mcnn_meshconv_bwd_data_tc_signs_p(const float* dout, const float* x, const int8_t* signs, const int32_t* gemm,
                                                 const int32_t* ec, const float* weight, float* dx, float* wtb_ws, int B, int E,
                                                 int Cin, int Cout, int precision, void* stream);
extern "C" int mcnn_meshconv_bwd_data_tc_signs(const float* dout, const float* x, const int8_t* signs, const int32_t* gemm,
                                               const int32_t* ec, const float* weight, float* dx, float* wtb_ws, int B, int E,
                                               int Cin, int Cout, void* stream) {
    return mcnn_meshconv_bwd_data_tc_signs_p(dout, x, signs, gemm, ec, weight, dx, wtb_ws, B, E, Cin, Cout, MCNN_PREC_FP32, stream);
}
extern "C" int mcnn_meshconv_bwd_data_tc_signs_p(const float* dout, const float* x, const int8_t* signs, const int32_t* gemm,
                                                 const int32_t* ec, const float* weight, float* dx, float* wtb_ws, int B, int E,
                                                 int Cin, int Cout, int precision, void* stream) {
    if (precision != MCNN_PREC_FP32 && precision != MCNN_PREC_TF32) return MCNN_E_BADARG;
    // single-pass TF32 exists in the persistent kernel only (it needs the signs the forward kept)
    const int one_pass = (precision == MCNN_PREC_TF32 && signs != nullptr && !g_tc_bd_tile) ? 1 : 0;
    if (!dout || (!x && !signs) || !gemm || !ec || !dx || !wtb_ws || B <= 0 || E <= 0) return MCNN_E_BADARG;
    if (!(Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 32 == 0)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int total = Cout * 5 * Cin;
    int rc = MCNN_OK;
    if (weight != nullptr) {                                 // NULL: wtb_ws already holds the image
        launch_k(weight_tc_bwd_layout_kernel, ceil_div(total, 256), 256, 0, st, weight, wtb_ws, Cin, Cout);
        if ((rc = after_launch())) return rc;
    }
    static DynSmemGuard configured;
    if (int grc = configured.ensure(meshconv_bwd_data_tc_kernel, BdSmem::TOTAL)) return grc;
    if (signs != nullptr && !g_tc_bd_tile) {                 // persistent kernel (needs the signs kept by the forward)
        static DynSmemGuard configured_p;
        if (int grc = configured_p.ensure(meshconv_bwd_data_tc_p_kernel, BdpSmem::TOTAL)) return grc;
        const long long items = (long long)ceil_div(B * E, TC_M) * (Cin / 32);
        const int G = (int)min((long long)sm_count(), items);
        launch_k(meshconv_bwd_data_tc_p_kernel, G, BDP_THREADS, BdpSmem::TOTAL, st, dout, gemm, ec, wtb_ws, dx, signs, B, E, Cin, Cout, one_pass);
        return after_launch();
    }
    dim3 grid(ceil_div(B * E, TC_M), Cin / 32);
    launch_k(meshconv_bwd_data_tc_kernel, grid, TC_THREADS, BdSmem::TOTAL, st, dout, x, gemm, ec, wtb_ws, dx, signs, B, E, Cin, Cout);
    return after_launch();
}

/* floats of workspace mcnn_meshconv_bwd_weight_tc needs */
extern "C" size_t mcnn_meshconv_bwd_weight_tc_ws(int B, int E, int Cin, int Cout) {
    int rows;
    const int S = bw_splits(B * E, Cin, Cout, &rows);
    return (size_t)S * Cout * Cin * 5 + (size_t)COLSUM_PARTS * Cout;
}

extern "C" int mcnn_meshconv_bwd_weight_tc_p(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                             float* dweight, float* dbias, float* ws, int B, int E, int Cin, int Cout,
                                             int precision, void* stream);
extern "C" int mcnn_meshconv_bwd_weight_tc(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                           float* dweight, float* dbias, float* ws, int B, int E, int Cin, int Cout,
                                           void* stream) {
    return mcnn_meshconv_bwd_weight_tc_p(dout, x, gemm, ec, dweight, dbias, ws, B, E, Cin, Cout, MCNN_PREC_FP32, stream);
}
extern "C" int mcnn_meshconv_bwd_weight_tc_p(const float* dout, const float* x, const int32_t* gemm, const int32_t* ec,
                                             float* dweight, float* dbias, float* ws, int B, int E, int Cin, int Cout,
                                             int precision, void* stream) {
    if (precision != MCNN_PREC_FP32 && precision != MCNN_PREC_TF32) return MCNN_E_BADARG;
    const int one_pass = (precision == MCNN_PREC_TF32 && g_tc_bw_tmem_a) ? 1 : 0;    // single-pass TF32: TMEM-operand kernel only
    if (!dout || !x || !gemm || !ec || !dweight || !ws || B <= 0 || E <= 0) return MCNN_E_BADARG;
    if (!(Cin > 0 && Cin % 32 == 0 && Cout > 0 && Cout % 4 == 0)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int M = B * E;
    int rows;
    const int S = bw_splits(M, Cin, Cout, &rows);
    if (g_tc_bw_tmem_a) {                                   // A operand in tensor memory (default)
        if (bw_ot(Cout) == 2) {
            static DynSmemGuard cfg2;
            if (int grc = cfg2.ensure(meshconv_bwd_weight_tca_kernel<2>, BwaSmem<2>::TOTAL)) return grc;
            dim3 grid(S, ceil_div(Cout, 256), Cin / 32);
            launch_k(meshconv_bwd_weight_tca_kernel<2>, grid, BWA_THREADS, BwaSmem<2>::TOTAL, st, dout, x, gemm, ec, ws, B, E, Cin, Cout, rows, one_pass);
        } else {
            static DynSmemGuard cfg1;
            if (int grc = cfg1.ensure(meshconv_bwd_weight_tca_kernel<1>, BwaSmem<1>::TOTAL)) return grc;
            dim3 grid(S, ceil_div(Cout, 128), Cin / 32);
            launch_k(meshconv_bwd_weight_tca_kernel<1>, grid, BWA_THREADS, BwaSmem<1>::TOTAL, st, dout, x, gemm, ec, ws, B, E, Cin, Cout, rows, one_pass);
        }
    } else if (bw_ot(Cout) == 2) {
        static DynSmemGuard configured2;
        if (int grc = configured2.ensure(meshconv_bwd_weight_tc_kernel<2>, BwSmem<2>::TOTAL)) return grc;
        dim3 grid(S, ceil_div(Cout, 256), Cin / 32);
        launch_k(meshconv_bwd_weight_tc_kernel<2>, grid, TC_THREADS, BwSmem<2>::TOTAL, st, dout, x, gemm, ec, ws, B, E, Cin, Cout, rows);
    } else {
        static DynSmemGuard configured;
        if (int grc = configured.ensure(meshconv_bwd_weight_tc_kernel<1>, BwSmem<1>::TOTAL)) return grc;
        dim3 grid(S, ceil_div(Cout, 128), Cin / 32);
        launch_k(meshconv_bwd_weight_tc_kernel<1>, grid, TC_THREADS, BwSmem<1>::TOTAL, st, dout, x, gemm, ec, ws, B, E, Cin, Cout, rows);
    }
    int rc = after_launch();
    if (rc) return rc;
    const int n = Cout * Cin * 5;
    launch_k(reduce_splits_kernel, ceil_div(n, 256), 256, 0, st, ws, dweight, n, S);
    rc = after_launch();
    if (rc) return rc;
    if (dbias != nullptr) {
        float* part = ws + (size_t)S * n;
        launch_k(colsum_part_kernel, dim3(ceil_div(Cout, 32), COLSUM_PARTS), dim3(32, 8), 0, st, dout, part, M, Cout);
        rc = after_launch();
        if (rc) return rc;
        launch_k(reduce_splits_kernel, ceil_div(Cout, 256), 256, 0, st, part, dbias, Cout, COLSUM_PARTS);
        rc = after_launch();
    }
    return rc;
}


extern "C" int mcnn_debug_tc_prof(long long* out16, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(long long) * 16);
    if (e == cudaSuccess && reset) {
        long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
    }
    return e == cudaSuccess ? MCNN_OK : MCNN_E_CUDA;
}

/* diagnostics (prof build): %globaltimer (ns) at entry and exit of every CTA of the last stream-K forward launch */
extern "C" int mcnn_debug_sk_cta_times(unsigned long long* out512) {
    cudaError_t e = cudaMemcpyFromSymbol(out512, g_sk_cta_ns, sizeof(unsigned long long) * 512);
    return e == cudaSuccess ? MCNN_OK : MCNN_E_CUDA;
}

/* diagnostics (prof build): clock64 stamps of CTA `cta` of the last stream-K forward launch; selects the CTA for the next */
extern "C" int mcnn_debug_sk_prof(long long* out64, int cta) {
    cudaError_t e = cudaMemcpyFromSymbol(out64, g_sk_prof, sizeof(long long) * 64);
    if (e == cudaSuccess) {
        long long z[64] = {0};
        e = cudaMemcpyToSymbol(g_sk_prof, z, sizeof(z));
    }
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_sk_prof_cta, &cta, sizeof(int));
    return e == cudaSuccess ? MCNN_OK : MCNN_E_CUDA;
}

extern "C" int mcnn_debug_set_tc_pair(int on) {
    g_tc_pw16 = (on & 2) ? 1 : 0;
    g_tc_pair = (on & 1) ? 1 : 0;
    return MCNN_OK;
}

extern "C" int mcnn_debug_set_tc_ntile(int ntile) {
    g_tc_force_ntile = (ntile == 64 || ntile == 128 || ntile == 256) ? ntile : 0;
    return MCNN_OK;
}

extern "C" int mcnn_debug_set_tc(int flags) {
    cudaError_t e = cudaMemcpyToSymbol(g_tc_debug, &flags, sizeof(int));
    return e == cudaSuccess ? MCNN_OK : MCNN_E_CUDA;
}
