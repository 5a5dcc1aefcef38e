// Classifier head of MeshConvNet (models/networks.py:143-157 of the reference):
//     x = AvgPool1d(res[-1])(x)  ->  x.view(-1, k[-1])  ->  relu(fc1(x))  ->  fc2(x)
// on the edge-major activation x[B][E][C].  In torch this is ~17 small launches per training step (mean, two addmm, relu,
// and in the backward four mm, two column sums, the relu mask, the broadcast of the pooled gradient, four gradient
// accumulations); here it is one forward and two backward kernels.  fp32 FFMA, deterministic (no atomics).  Everything is
// tiny except the gradient of the average pooling, which writes the whole [B][E][C] tensor once.
#include <cooperative_groups.h>

#include "common.cuh"

namespace mcnn {

constexpr int HEAD_NT = 256;

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// grid = (B, HEAD_S), one thread-block CLUSTER of HEAD_S CTAs per mesh: the average over the edge axis is the only real memory
// traffic (the whole activation, once) and wants many loads in flight, so every CTA of the cluster sums a slice of the rows;
// the partial sums meet through distributed shared memory, every CTA then computes its share of the hidden units and hands
// them to CTA 0, which finishes with fc2.
//     pooled[b][c] = mean_e x[b][e][c];  h[b][j] = relu(b1[j] + W1[j] . pooled[b]);  logits[b][k] = b2[k] + W2[k] . h[b]
constexpr int HEAD_S = 4;

__global__ void __launch_bounds__(HEAD_NT, 2) head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2,
                                                           const float* __restrict__ b2, float* __restrict__ pooled,
                                                           float* __restrict__ h, float* __restrict__ logits, int E, int C, int H,
                                                           int K) {
    pdl_enter();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float sm[];    // red[rpp][C] | part[C] | pooled[C] | h[H]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();    // == blockIdx.y
    const int cq = C >> 2;
    const int lanes_q = min(cq, HEAD_NT);
    const int rpp = HEAD_NT / lanes_q;              // row lanes
    const int tx = tid % lanes_q, ty = tid / lanes_q;
    float* red = sm;
    float* part = sm + (size_t)rpp * C;
    float* ps = part + C;
    float* hs = ps + C;
    const int per = (E + HEAD_S - 1) / HEAD_S;
    const int e0 = rank * per, e1 = min(E, e0 + per);
    const float* xb = x + (size_t)b * E * C;
    for (int q = tx; q < cq && ty < rpp; q += lanes_q) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = e0 + ty;
        for (; e + 7 * rpp < e1; e += 8 * rpp) {                 // eight independent rows in flight
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)(e + k * rpp) * C) + q);
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        }
        for (; e < e1; e += rpp) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)e * C) + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(red + (size_t)ty * C + 4 * q) = acc;
    }
    __syncthreads();
    for (int c = tid; c < C; c += HEAD_NT) {
        float s = 0.f;
        for (int y = 0; y < rpp; ++y) s += red[(size_t)y * C + c];
        part[c] = s;
    }
    cluster.sync();                                              // every CTA's slice sum is in its `part`
    const float inv = 1.f / (float)E;
    for (int c = tid; c < C; c += HEAD_NT) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < HEAD_S; ++r) s += cluster.map_shared_rank(part, r)[c];        // fixed order: deterministic
        s *= inv;
        ps[c] = s;
        if (rank == 0) pooled[(size_t)b * C + c] = s;
    }
    __syncthreads();
    // hidden units rank, rank + S, ...: one warp per unit, kept in CTA 0's shared memory for fc2
    const int warp = tid >> 5, lane = tid & 31;
    float* hs0 = cluster.map_shared_rank(hs, 0);
    for (int j = rank + HEAD_S * warp; j < H; j += HEAD_S * (HEAD_NT / 32)) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w1 + (size_t)j * C + c), ps[c], s);
        s = warp_sum_f(s);
        if (lane == 0) {
            s = fmaxf(s + (b1 != nullptr ? __ldg(b1 + j) : 0.f), 0.f);
            hs0[j] = s;
            h[(size_t)b * H + j] = s;
        }
    }
    cluster.sync();                                              // h complete in CTA 0; nobody reads remote memory after this
    if (rank != 0) return;
    for (int k = warp; k < K; k += HEAD_NT / 32) {
        float s = 0.f;
        for (int j = lane; j < H; j += 32) s = fmaf(__ldg(w2 + (size_t)k * H + j), hs[j], s);
        s = warp_sum_f(s);
        if (lane == 0) logits[(size_t)b * K + k] = s + (b2 != nullptr ? __ldg(b2 + k) : 0.f);
    }
}

// grid = (B, S).  dh[b][j] = [h > 0] * sum_k dlogits[b][k] W2[k][j];  dpooled[c] = sum_j dh[j] W1[j][c];
// dx[b][e][c] = dpooled[c] / E for the rows e of slice blockIdx.y (every slice recomputes the two small products; the
// broadcast over the edge axis is the only real memory traffic and wants more than B CTAs)
__global__ void __launch_bounds__(HEAD_NT, 2) head_bwd_data_kernel(const float* __restrict__ dlogits, const float* __restrict__ h,
                                                                const float* __restrict__ w1, const float* __restrict__ w2,
                                                                float* __restrict__ dh, float* __restrict__ dx, int E, int C, int H,
                                                                int K) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];    // dp[C] (read back as float4) | part[JG][C] | dl[K] | dh[H]
    constexpr int JG = 4;                          // groups of hidden units summed side by side
    const int b = blockIdx.x, tid = threadIdx.x;
    float* dp = sm;
    float* part = dp + C;
    float* dl = part + (size_t)JG * C;
    float* dhs = dl + K;
    for (int k = tid; k < K; k += HEAD_NT) dl[k] = __ldg(dlogits + (size_t)b * K + k);
    __syncthreads();
    for (int j = tid; j < H; j += HEAD_NT) {
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < K; ++k) s = fmaf(dl[k], __ldg(w2 + (size_t)k * H + j), s);
        s = (__ldg(h + (size_t)b * H + j) > 0.f) ? s : 0.f;
        dhs[j] = s;
        if (blockIdx.y == 0) dh[(size_t)b * H + j] = s;
    }
    __syncthreads();
    const int cq = C >> 2;
    {   // dpooled: thread = (quad of channels, group of hidden units); the JG partial sums meet in shared memory
        const int hj = (H + JG - 1) / JG;
        for (int i = tid; i < cq * JG; i += HEAD_NT) {
            const int q = i % cq, g = i / cq;
            const int j0 = g * hj, j1 = min(H, j0 + hj);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int j = j0; j < j1; ++j) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)j * C) + q);
                const float d = dhs[j];
                acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
            }
            *reinterpret_cast<float4*>(part + (size_t)g * C + 4 * q) = acc;
        }
    }
    __syncthreads();
    const float inv = 1.f / (float)E;
    for (int c = tid; c < C; c += HEAD_NT) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < JG; ++g) s += part[(size_t)g * C + c];
        dp[c] = s * inv;
    }
    __syncthreads();
    if (dx != nullptr) {
        const int per = (E + gridDim.y - 1) / gridDim.y;
        const int e0 = blockIdx.y * per, e1 = min(E, e0 + per);
        float4* xb = reinterpret_cast<float4*>(dx + ((size_t)b * E + e0) * C);
        const int total = max(0, e1 - e0) * cq;
        for (int i = tid; i < total; i += HEAD_NT) xb[i] = *reinterpret_cast<const float4*>(dp + 4 * (i % cq));
    }
}

// grid = H + K.  CTA j < H: dW1[j][c] = sum_b dh[b][j] pooled[b][c], db1[j] = sum_b dh[b][j];
//                CTA H + k: dW2[k][j] = sum_b dlogits[b][k] h[b][j], db2[k] = sum_b dlogits[b][k]
__global__ void __launch_bounds__(HEAD_NT, 2) head_bwd_weight_kernel(const float* __restrict__ dlogits, const float* __restrict__ dh,
                                                                  const float* __restrict__ pooled, const float* __restrict__ h,
                                                                  float* __restrict__ dw1, float* __restrict__ db1,
                                                                  float* __restrict__ dw2, float* __restrict__ db2, int B, int C,
                                                                  int H, int K) {
    pdl_enter();
    const int tid = threadIdx.x;
    const bool first = (int)blockIdx.x < H;
    const int row = first ? blockIdx.x : blockIdx.x - H;
    const float* g = first ? dh : dlogits;          // [B][H] / [B][K]: the factor that is constant over the row
    const float* act = first ? pooled : h;          // [B][C] / [B][H]
    const int gs = first ? H : K, n = first ? C : H;
    float* dw = (first ? dw1 : dw2) + (size_t)row * n;
    for (int c = tid; c < n; c += HEAD_NT) {
        float s = 0.f;
#pragma unroll 16
        for (int b = 0; b < B; ++b) s = fmaf(__ldg(g + (size_t)b * gs + row), __ldg(act + (size_t)b * n + c), s);
        dw[c] = s;
    }
    float* db = first ? db1 : db2;
    if (db != nullptr && tid < 32) {
        float s = 0.f;
        for (int b = tid; b < B; b += 32) s += __ldg(g + (size_t)b * gs + row);
        s = warp_sum_f(s);
        if (tid == 0) db[row] = s;
    }
}

}  // namespace mcnn

using namespace mcnn;

static bool head_args_ok(int B, int E, int C, int H, int K) {
    return B > 0 && E > 0 && C > 0 && (C % 4) == 0 && C <= 2048 && H > 0 && H <= 4096 && K > 0 && K <= 4096;
}

extern "C" int mcnn_head_fwd(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* pooled,
                             float* h, float* logits, int B, int E, int C, int H, int K, void* stream) {
    if (!x || !w1 || !w2 || !pooled || !h || !logits || !head_args_ok(B, E, C, H, K)) return MCNN_E_BADARG;
    const int rpp = HEAD_NT / ((C / 4) < HEAD_NT ? (C / 4) : HEAD_NT);
    const size_t smem = ((size_t)rpp * C + 2 * (size_t)C + H) * sizeof(float);
    static DynSmemGuard configured;
    if (smem > 48 * 1024)
        if (int rc = configured.ensure(head_fwd_kernel, smem)) return rc;
    launch_cluster(head_fwd_kernel, dim3(B, HEAD_S), HEAD_NT, smem, (cudaStream_t)stream, dim3(1, HEAD_S, 1), x, w1, b1, w2, b2, pooled, h,
                   logits, E, C, H, K);
    return after_launch();
}

extern "C" int mcnn_head_bwd(const float* dlogits, const float* pooled, const float* h, const float* w1, const float* w2,
                             float* dh, float* dx, float* dw1, float* db1, float* dw2, float* db2, int B, int E, int C, int H,
                             int K, void* stream) {
    if (!dlogits || !pooled || !h || !w1 || !w2 || !dh || !dw1 || !dw2 || !head_args_ok(B, E, C, H, K)) return MCNN_E_BADARG;
    const size_t smem = ((size_t)K + H + 5 * (size_t)C) * sizeof(float);
    static DynSmemGuard configured;
    if (smem > 48 * 1024)
        if (int rc = configured.ensure(head_bwd_data_kernel, smem)) return rc;
    int slices = ceil_div(2 * sm_count(), B);                  // ~2 CTAs per SM write the broadcast gradient
    slices = slices < 1 ? 1 : (slices > 8 ? 8 : slices);
    if (dx == nullptr) slices = 1;
    launch_k(head_bwd_data_kernel, dim3(B, slices), HEAD_NT, smem, (cudaStream_t)stream, dlogits, h, w1, w2, dh, dx, E, C, H, K);
    if (int rc = after_launch()) return rc;
    launch_k(head_bwd_weight_kernel, H + K, HEAD_NT, 0, (cudaStream_t)stream, dlogits, dh, pooled, h, dw1, db1, dw2, db2, B, C, H, K);
    return after_launch();
}
