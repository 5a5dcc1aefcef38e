// Fused normalisation blocks on edge-major activations x[rows][C] (row N1 of SURVEY.md §8f: the elementwise /
// normalisation passes around MeshConv, networks.py:149, :171-179, :222-233, :276-288 of the reference).
//
//     z = pre(x [+ a])            pre  = relu or identity
//     y = post(gamma * (z - mean) * rstd + beta [+ r])        post = relu or identity
//
// with statistics over (block of `rows_per_block` consecutive rows) x (group of C/G consecutive channels):
//     BatchNorm2d    : one block of all B*E rows, G = C      (training-mode batch statistics; padded slots included, F8)
//     GroupNorm(g)   : one block per mesh (E rows), G = g
//     InstanceNorm2d : one block per mesh, G = C, gamma = 1, beta = 0
// Three kernels forward (slab partial sums -> finalize in fp64 -> apply), three backward; every reduction is two-stage and
// deterministic.  All HBM-bound: each activation is read twice and written once per direction.
#include "common.cuh"

namespace mcnn {

constexpr int NT = 256;

struct NormShape {
    int nblocks, rows_per_block, C, G, slabs, rows_per_slab;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// per-slab column sums of (u, v); FN supplies the two quantities for one float4 of columns
template <typename FN>
__device__ __forceinline__ void slab_colsums(const NormShape s, float* __restrict__ part, FN fn) {
    extern __shared__ float sm[];                 // [rpp][2][C]
    const int cq = s.C >> 2;
    const int rpp = NT / cq;                       // rows per pass
    const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
    const int blk = blockIdx.x, slab = blockIdx.y;
    const int r0 = slab * s.rows_per_slab, r1 = min(s.rows_per_block, r0 + s.rows_per_slab);
    float u[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ty < rpp) {
        int r = r0 + ty;
        for (; r + 3 * rpp < r1; r += 4 * rpp) {          // four independent rows in flight
            float4 a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                fn(((size_t)blk * s.rows_per_block + r + k * rpp) * s.C + (tx << 2), tx << 2, a[k], b[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u[0] += a[k].x; u[1] += a[k].y; u[2] += a[k].z; u[3] += a[k].w;
                v[0] += b[k].x; v[1] += b[k].y; v[2] += b[k].z; v[3] += b[k].w;
            }
        }
        for (; r < r1; r += rpp) {
            const size_t off = ((size_t)blk * s.rows_per_block + r) * s.C + (tx << 2);
            float4 a, b;
            fn(off, tx << 2, a, b);
            u[0] += a.x; u[1] += a.y; u[2] += a.z; u[3] += a.w;
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        *reinterpret_cast<float4*>(sm + (ty * 2 + 0) * s.C + (tx << 2)) = make_float4(u[0], u[1], u[2], u[3]);
        *reinterpret_cast<float4*>(sm + (ty * 2 + 1) * s.C + (tx << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * s.C; i += NT) {
        const int which = i / s.C, c = i - which * s.C;
        float acc = 0.f;
        for (int y = 0; y < rpp; ++y) acc += sm[(y * 2 + which) * s.C + c];
        part[(((size_t)blk * s.slabs + slab) * 2 + which) * s.C + c] = acc;
    }
}

__global__ void __launch_bounds__(NT) norm_stats_kernel(const float* __restrict__ x, const float* __restrict__ a, int pre_relu,
                                                        float* __restrict__ part, NormShape s) {
    pdl_enter();
    slab_colsums(s, part, [&](size_t off, int, float4& u, float4& v) {
        float4 z = ld4(x + off);
        if (a != nullptr) { const float4 t = ld4(a + off); z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w; }
        if (pre_relu) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
        u = z;
        v = make_float4(z.x * z.x, z.y * z.y, z.z * z.z, z.w * z.w);
    });
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// mean / rstd per (block, group) from the slab partials, in fp64 (one warp per output); optional BatchNorm running-stat update
__global__ void norm_finalize_kernel(const float* __restrict__ part, float* __restrict__ mean, float* __restrict__ rstd,
                                     float* running_mean, float* running_var, float momentum, float eps, NormShape s) {
    pdl_enter();
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= s.nblocks * s.G) return;
    const int blk = idx / s.G, g = idx - blk * s.G;
    const int cpg = s.C / s.G;
    double su = 0.0, sv = 0.0;
    const int items = s.slabs * cpg;
#pragma unroll 4
    for (int it = lane; it < items; it += 32) {
        const int sl = it / cpg, c = g * cpg + (it - sl * cpg);
        su += (double)part[(((size_t)blk * s.slabs + sl) * 2 + 0) * s.C + c];
        sv += (double)part[(((size_t)blk * s.slabs + sl) * 2 + 1) * s.C + c];
    }
    su = warp_sum(su);
    sv = warp_sum(sv);
    if (lane != 0) return;
    const double n = (double)s.rows_per_block * cpg;
    const double m = su / n;
    double var = sv / n - m * m;
    if (var < 0.0) var = 0.0;
    mean[idx] = (float)m;
    rstd[idx] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {                       // BatchNorm2d training: unbiased variance into the running buffer
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[g] = (1.f - momentum) * running_mean[g] + momentum * (float)m;
        running_var[g] = (1.f - momentum) * running_var[g] + momentum * (float)unb;
    }
}

// y = post(gamma * (pre(x [+ a]) - mean) * rstd + beta [+ r]).  Same (block, row slab) decomposition as the statistics:
// a thread owns one float4 of columns, so mean / rstd / gamma / beta are loaded once and folded into scale / shift; the
// row loop keeps four rows in flight.
__global__ void __launch_bounds__(NT) norm_apply_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                        const float* __restrict__ r, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, float* __restrict__ y, int pre_relu,
                                                        int post_relu, NormShape s) {
    pdl_enter();
    const int cq = s.C >> 2;
    const int rpp = NT / cq;
    const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
    if (ty >= rpp) return;
    const int blk = blockIdx.x, slab = blockIdx.y;
    const int r0 = slab * s.rows_per_slab, r1 = min(s.rows_per_block, r0 + s.rows_per_slab);
    const int c = tx << 2;
    const int cpg = s.C / s.G;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = (c + j) / cpg;
        const float m = __ldg(mean + blk * s.G + g), rs = __ldg(rstd + blk * s.G + g);
        const float gm = gamma != nullptr ? __ldg(gamma + c + j) : 1.f;
        const float bt = gamma != nullptr ? __ldg(beta + c + j) : 0.f;
        sc[j] = rs * gm;
        sh[j] = bt - m * rs * gm;
    }
    const size_t base = (size_t)blk * s.rows_per_block * s.C + c;
    auto one = [&](const float4 z4, const float4 a4, const float4 r4) {
        float z[4] = {z4.x + a4.x, z4.y + a4.y, z4.z + a4.z, z4.w + a4.w};
        const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v = pre_relu ? fmaxf(z[j], 0.f) : z[j];
            const float w = fmaf(v, sc[j], sh[j]) + rr[j];
            o[j] = post_relu ? fmaxf(w, 0.f) : w;
        }
        return make_float4(o[0], o[1], o[2], o[3]);
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int row = r0 + ty;
    for (; row + 3 * rpp < r1; row += 4 * rpp) {
        float4 z[4], av[4], rv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t off = base + (size_t)(row + k * rpp) * s.C;
            z[k] = ld4(x + off);
            av[k] = a != nullptr ? ld4(a + off) : zero;
            rv[k] = r != nullptr ? ld4(r + off) : zero;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(y + base + (size_t)(row + k * rpp) * s.C) = one(z[k], av[k], rv[k]);
    }
    for (; row < r1; row += rpp) {
        const size_t off = base + (size_t)row * s.C;
        *reinterpret_cast<float4*>(y + off) = one(ld4(x + off), a != nullptr ? ld4(a + off) : zero, r != nullptr ? ld4(r + off) : zero);
    }
}

// backward stage 1: per-slab column sums of (dyh, dyh * xhat), dyh = dy * [y > 0 if post_relu]
__global__ void __launch_bounds__(NT) norm_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ a,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            int pre_relu, int post_relu, float* __restrict__ part, NormShape s) {
    pdl_enter();
    const int cpg = s.C / s.G;
    const int blk = blockIdx.x;
    slab_colsums(s, part, [&](size_t off, int c, float4& u, float4& v) {
        float4 d = ld4(dy + off);
        if (post_relu) {
            const float4 yy = ld4(y + off);
            d.x = yy.x > 0.f ? d.x : 0.f; d.y = yy.y > 0.f ? d.y : 0.f; d.z = yy.z > 0.f ? d.z : 0.f; d.w = yy.w > 0.f ? d.w : 0.f;
        }
        float4 z = ld4(x + off);
        if (a != nullptr) { const float4 t = ld4(a + off); z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w; }
        if (pre_relu) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
        const float zz[4] = {z.x, z.y, z.z, z.w};
        float xh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = (c + j) / cpg;
            xh[j] = (zz[j] - __ldg(mean + blk * s.G + g)) * __ldg(rstd + blk * s.G + g);
        }
        u = d;
        v = make_float4(d.x * xh[0], d.y * xh[1], d.z * xh[2], d.w * xh[3]);
    });
}

// backward stage 2 (one warp per output): per (block, group) s1 = sum gamma*A / n, s2 = sum gamma*Bc / n;
// then dgamma[c] = sum_blocks Bc, dbeta[c] = sum_blocks A
__global__ void norm_bwd_finalize_kernel(const float* __restrict__ part, const float* __restrict__ gamma, float* __restrict__ s1,
                                         float* __restrict__ s2, float* __restrict__ dgamma, float* __restrict__ dbeta, NormShape s) {
    pdl_enter();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int cpg = s.C / s.G;
    const int ngroups = s.nblocks * s.G;
    if (w < ngroups) {
        const int blk = w / s.G, g = w - blk * s.G;
        double a1 = 0.0, a2 = 0.0;
        const int items = s.slabs * cpg;
#pragma unroll 4
        for (int it = lane; it < items; it += 32) {
            const int sl = it / cpg, c = g * cpg + (it - sl * cpg);
            const double gm = gamma != nullptr ? (double)gamma[c] : 1.0;
            a1 += gm * (double)part[(((size_t)blk * s.slabs + sl) * 2 + 0) * s.C + c];
            a2 += gm * (double)part[(((size_t)blk * s.slabs + sl) * 2 + 1) * s.C + c];
        }
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
            const double n = (double)s.rows_per_block * cpg;
            s1[w] = (float)(a1 / n);
            s2[w] = (float)(a2 / n);
        }
    } else if (dgamma != nullptr && w < ngroups + s.C) {
        const int c = w - ngroups;
        double A = 0.0, Bc = 0.0;
        const int items = s.nblocks * s.slabs;
#pragma unroll 4
        for (int it = lane; it < items; it += 32) {
            A += (double)part[((size_t)it * 2 + 0) * s.C + c];
            Bc += (double)part[((size_t)it * 2 + 1) * s.C + c];
        }
        A = warp_sum(A);
        Bc = warp_sum(Bc);
        if (lane == 0) { dgamma[c] = (float)Bc; dbeta[c] = (float)A; }
    }
}

// backward stage 3: dz = rstd * (gamma*dyh - s1 - xhat*s2); dx = da = dz * [z_pre > 0 if pre_relu]; dr = dyh.
// Slab form as above: per-column constants hoisted (xhat = z*rs - m*rs), two rows in flight.
__global__ void __launch_bounds__(NT) norm_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ a,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ s1,
                                                            const float* __restrict__ s2, float* __restrict__ dx,
                                                            float* __restrict__ dr, const float* __restrict__ dacc, int pre_relu,
                                                            int post_relu, NormShape s) {
    pdl_enter();
    const int cq = s.C >> 2;
    const int rpp = NT / cq;
    const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
    if (ty >= rpp) return;
    const int blk = blockIdx.x, slab = blockIdx.y;
    const int r0 = slab * s.rows_per_slab, r1 = min(s.rows_per_block, r0 + s.rows_per_slab);
    const int c = tx << 2;
    const int cpg = s.C / s.G;
    float rs[4], mrs[4], gm[4], k1[4], k2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = (c + j) / cpg;
        rs[j] = __ldg(rstd + blk * s.G + g);
        mrs[j] = __ldg(mean + blk * s.G + g) * rs[j];
        gm[j] = gamma != nullptr ? __ldg(gamma + c + j) : 1.f;
        k1[j] = __ldg(s1 + blk * s.G + g);
        k2[j] = __ldg(s2 + blk * s.G + g);
    }
    const size_t base = (size_t)blk * s.rows_per_block * s.C + c;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    auto one = [&](size_t off, const float4 d4, const float4 y4, const float4 z4, const float4 a4) {
        float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
        const float zr[4] = {z4.x + a4.x, z4.y + a4.y, z4.z + a4.z, z4.w + a4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (post_relu) d[j] = yv[j] > 0.f ? d[j] : 0.f;
            const float z = pre_relu ? fmaxf(zr[j], 0.f) : zr[j];
            const float xh = z * rs[j] - mrs[j];
            float dz = rs[j] * (gm[j] * d[j] - k1[j] - xh * k2[j]);
            if (pre_relu && !(zr[j] > 0.f)) dz = 0.f;
            o[j] = dz;
        }
        if (dr != nullptr) *reinterpret_cast<float4*>(dr + off) = make_float4(d[0], d[1], d[2], d[3]);
        if (dacc != nullptr) {                                   // a second gradient of x (it also feeds a residual branch)
            const float4 t = ld4(dacc + off);
            o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
        }
        *reinterpret_cast<float4*>(dx + off) = make_float4(o[0], o[1], o[2], o[3]);
    };
    int row = r0 + ty;
    for (; row + rpp < r1; row += 2 * rpp) {
        float4 d[2], yy[2], z[2], av[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const size_t off = base + (size_t)(row + k * rpp) * s.C;
            d[k] = ld4(dy + off);
            yy[k] = post_relu ? ld4(y + off) : zero;
            z[k] = ld4(x + off);
            av[k] = a != nullptr ? ld4(a + off) : zero;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) one(base + (size_t)(row + k * rpp) * s.C, d[k], yy[k], z[k], av[k]);
    }
    for (; row < r1; row += rpp) {
        const size_t off = base + (size_t)row * s.C;
        one(off, ld4(dy + off), post_relu ? ld4(y + off) : zero, ld4(x + off), a != nullptr ? ld4(a + off) : zero);
    }
}

// ---- fused per-mesh forms (GroupNorm / InstanceNorm2d: statistics never cross a mesh) ------------------------------------
// One CTA = one block of rows (a mesh) x CW consecutive channels (whole groups).  The pre-activation z = pre(x [+ a]) of the
// CTA's slab is kept in shared memory, so each activation is read ONCE: partial sums -> block reduction -> group statistics
// (fp64) -> normalise from shared memory.  One launch instead of three per direction (+ the small dgamma / dbeta reduction
// in the backward), and no statistics round trip through global memory between dependent launches -- at the sizes of this
// workload the three-kernel form is bound by exactly those dependent launches, not by HBM.
struct FusedShape { int nblocks, rows, C, G, CW; };

__device__ __forceinline__ void fused_reduce_columns(float* red, float* tot, const float (&u)[4], const float (&v)[4], int tx, int ty,
                                                     int rl, int CW) {
    // red[ty][2][CW] -> tot[2][CW]
    // one 16-byte store per thread and quantity: a quarter warp writes 128 contiguous bytes (the scalar form put the four
    // rows of a warp on the same banks: 4-way conflicts on every store)
    *reinterpret_cast<float4*>(red + (ty * 2 + 0) * CW + (tx << 2)) = make_float4(u[0], u[1], u[2], u[3]);
    *reinterpret_cast<float4*>(red + (ty * 2 + 1) * CW + (tx << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * CW; i += NT) {
        const int which = i / CW, c = i - which * CW;
        float acc = 0.f;
        for (int y = 0; y < rl; ++y) acc += red[(y * 2 + which) * CW + c];
        tot[i] = acc;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(NT) norm_fused_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                            const float* __restrict__ r, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, float eps,
                                                            int pre_relu, int post_relu, FusedShape s) {
    pdl_enter();
    extern __shared__ float sm[];
    const int CW = s.CW, cq = CW >> 2, rl = NT / cq;
    float* zs = sm;                                   // [rows][CW]
    float* red = zs + (size_t)s.rows * CW;            // [rl][2][CW]
    float* tot = red + rl * 2 * CW;                   // [2][CW]
    float* stat = tot + 2 * CW;                       // [2][CW / cpg]: mean, rstd of the CTA's groups
    const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
    const int blk = blockIdx.x, c0 = blockIdx.y * CW;
    const int cpg = s.C / s.G;
    const size_t base = (size_t)blk * s.rows * s.C + c0 + (tx << 2);
    float u[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
    auto take = [&](int row, float4 z, const float4 t) {
        z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w;
        if (pre_relu) { z.x = fmaxf(z.x, 0.f); z.y = fmaxf(z.y, 0.f); z.z = fmaxf(z.z, 0.f); z.w = fmaxf(z.w, 0.f); }
        *reinterpret_cast<float4*>(zs + (size_t)row * CW + (tx << 2)) = z;
        u[0] += z.x; u[1] += z.y; u[2] += z.z; u[3] += z.w;
        v[0] += z.x * z.x; v[1] += z.y * z.y; v[2] += z.z * z.z; v[3] += z.w * z.w;
    };
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    int row = ty;
    for (; row + 3 * rl < s.rows; row += 4 * rl) {            // four independent rows in flight
        float4 z[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t off = base + (size_t)(row + k * rl) * s.C;
            z[k] = ld4(x + off);
            t[k] = a != nullptr ? ld4(a + off) : zero;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) take(row + k * rl, z[k], t[k]);
    }
    for (; row < s.rows; row += rl) {
        const size_t off = base + (size_t)row * s.C;
        take(row, ld4(x + off), a != nullptr ? ld4(a + off) : zero);
    }
    fused_reduce_columns(red, tot, u, v, tx, ty, rl, CW);
    const int ngrp = CW / cpg;
    if (threadIdx.x < ngrp) {
        const int g = threadIdx.x;
        double su = 0.0, sv = 0.0;
        for (int j = 0; j < cpg; ++j) { su += (double)tot[g * cpg + j]; sv += (double)tot[CW + g * cpg + j]; }
        const double n = (double)s.rows * cpg;
        const double m = su / n;
        double var = sv / n - m * m;
        if (var < 0.0) var = 0.0;
        const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)eps));
        stat[g] = mf; stat[ngrp + g] = rf;
        mean[blk * s.G + c0 / cpg + g] = mf;
        rstd[blk * s.G + c0 / cpg + g] = rf;
    }
    __syncthreads();
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cl = (tx << 2) + j, g = cl / cpg;
        const float m = stat[g], rs = stat[ngrp + g];
        const float gm = gamma != nullptr ? __ldg(gamma + c0 + cl) : 1.f;
        const float bt = gamma != nullptr ? __ldg(beta + c0 + cl) : 0.f;
        sc[j] = rs * gm;
        sh[j] = bt - m * rs * gm;
    }
    auto put = [&](int row_, const float4 r4) {
        const float4 z = *reinterpret_cast<const float4*>(zs + (size_t)row_ * CW + (tx << 2));
        float o[4] = {fmaf(z.x, sc[0], sh[0]) + r4.x, fmaf(z.y, sc[1], sh[1]) + r4.y, fmaf(z.z, sc[2], sh[2]) + r4.z, fmaf(z.w, sc[3], sh[3]) + r4.w};
        if (post_relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        *reinterpret_cast<float4*>(y + base + (size_t)row_ * s.C) = make_float4(o[0], o[1], o[2], o[3]);
    };
    row = ty;
    if (r != nullptr) {
        for (; row + 3 * rl < s.rows; row += 4 * rl) {
            float4 rv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rv[k] = ld4(r + base + (size_t)(row + k * rl) * s.C);
#pragma unroll
            for (int k = 0; k < 4; ++k) put(row + k * rl, rv[k]);
        }
        for (; row < s.rows; row += rl) put(row, ld4(r + base + (size_t)row * s.C));
    } else {
        for (; row < s.rows; row += rl) put(row, zero);
    }
}

// backward: pass 1 keeps the raw z = x [+ a] of the slab in shared memory and accumulates A_c = sum dyh, B_c = sum dyh * xhat;
// the per-mesh column sums go to part[blk][2][C] (dgamma / dbeta are reduced over the meshes by norm_bwd_finalize_kernel);
// pass 2 re-reads dy (and y) from L2 and writes dx (and dr).
__global__ void __launch_bounds__(NT) norm_fused_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ a,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ dx,
                                                            float* __restrict__ dr, float* __restrict__ part, int pre_relu,
                                                            int post_relu, FusedShape s) {
    pdl_enter();
    extern __shared__ float sm[];
    const int CW = s.CW, cq = CW >> 2, rl = NT / cq;
    float* zs = sm;
    float* red = zs + (size_t)s.rows * CW;
    float* tot = red + rl * 2 * CW;
    float* stat = tot + 2 * CW;                       // [2][ngrp]: s1, s2
    const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
    const int blk = blockIdx.x, c0 = blockIdx.y * CW;
    const int cpg = s.C / s.G;
    const int ngrp = CW / cpg;
    const size_t base = (size_t)blk * s.rows * s.C + c0 + (tx << 2);
    float rs[4], mrs[4], gm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cl = (tx << 2) + j, g = (c0 + cl) / cpg;
        rs[j] = __ldg(rstd + blk * s.G + g);
        mrs[j] = __ldg(mean + blk * s.G + g) * rs[j];
        gm[j] = gamma != nullptr ? __ldg(gamma + c0 + cl) : 1.f;
    }
    float u[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    auto take = [&](int row_, const float4 d4, const float4 y4, float4 z, const float4 t) {
        z.x += t.x; z.y += t.y; z.z += t.z; z.w += t.w;
        *reinterpret_cast<float4*>(zs + (size_t)row_ * CW + (tx << 2)) = z;          // raw: pass 2 needs the sign for pre_relu
        float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
        const float zr[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (post_relu) d[j] = yv[j] > 0.f ? d[j] : 0.f;
            const float zc = pre_relu ? fmaxf(zr[j], 0.f) : zr[j];
            const float xh = zc * rs[j] - mrs[j];
            u[j] += d[j];
            v[j] += d[j] * xh;
        }
    };
    int row = ty;
    for (; row + rl < s.rows; row += 2 * rl) {
        float4 d[2], yy[2], z[2], t[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const size_t off = base + (size_t)(row + k * rl) * s.C;
            d[k] = ld4(dy + off);
            yy[k] = post_relu ? ld4(y + off) : zero;
            z[k] = ld4(x + off);
            t[k] = a != nullptr ? ld4(a + off) : zero;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) take(row + k * rl, d[k], yy[k], z[k], t[k]);
    }
    for (; row < s.rows; row += rl) {
        const size_t off = base + (size_t)row * s.C;
        take(row, ld4(dy + off), post_relu ? ld4(y + off) : zero, ld4(x + off), a != nullptr ? ld4(a + off) : zero);
    }
    fused_reduce_columns(red, tot, u, v, tx, ty, rl, CW);
    for (int i = threadIdx.x; i < 2 * CW; i += NT) {           // per-mesh column sums for dbeta (A) / dgamma (B)
        const int which = i / CW, c = i - which * CW;
        part[((size_t)blk * 2 + which) * s.C + c0 + c] = tot[i];
    }
    if (threadIdx.x < ngrp) {
        const int g = threadIdx.x;
        double a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < cpg; ++j) {
            const double gmj = gamma != nullptr ? (double)__ldg(gamma + c0 + g * cpg + j) : 1.0;
            a1 += gmj * (double)tot[g * cpg + j];
            a2 += gmj * (double)tot[CW + g * cpg + j];
        }
        const double n = (double)s.rows * cpg;
        stat[g] = (float)(a1 / n);
        stat[ngrp + g] = (float)(a2 / n);
    }
    __syncthreads();
    float k1[4], k2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = ((tx << 2) + j) / cpg;
        k1[j] = stat[g];
        k2[j] = stat[ngrp + g];
    }
    auto put = [&](int row_, const float4 d4, const float4 y4) {
        const size_t off = base + (size_t)row_ * s.C;
        const float4 z4 = *reinterpret_cast<const float4*>(zs + (size_t)row_ * CW + (tx << 2));
        float d[4] = {d4.x, d4.y, d4.z, d4.w};
        const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
        const float zr[4] = {z4.x, z4.y, z4.z, z4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (post_relu) d[j] = yv[j] > 0.f ? d[j] : 0.f;
            const float z = pre_relu ? fmaxf(zr[j], 0.f) : zr[j];
            const float xh = z * rs[j] - mrs[j];
            float dz = rs[j] * (gm[j] * d[j] - k1[j] - xh * k2[j]);
            if (pre_relu && !(zr[j] > 0.f)) dz = 0.f;
            o[j] = dz;
        }
        if (dr != nullptr) *reinterpret_cast<float4*>(dr + off) = make_float4(d[0], d[1], d[2], d[3]);
        *reinterpret_cast<float4*>(dx + off) = make_float4(o[0], o[1], o[2], o[3]);
    };
    row = ty;
    for (; row + 3 * rl < s.rows; row += 4 * rl) {
        float4 d[4], yy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const size_t off = base + (size_t)(row + k * rl) * s.C;
            d[k] = ld4(dy + off);
            yy[k] = post_relu ? ld4(y + off) : zero;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) put(row + k * rl, d[k], yy[k]);
    }
    for (; row < s.rows; row += rl) {
        const size_t off = base + (size_t)row * s.C;
        put(row, ld4(dy + off), post_relu ? ld4(y + off) : zero);
    }
}

static int g_norm_fused = 1;                 // debug: 0 = always the three-kernel form

// channel width of the fused form for a shape, 0 if it does not apply (statistics crossing meshes, odd group sizes, or a
// slab that does not fit in shared memory)
static int fused_cw(int nblocks, int rows, int C, int G, size_t* smem_bytes) {
    if (!g_norm_fused || nblocks < 8) return 0;
    const int cpg = C / G;
    if (cpg & (cpg - 1)) return 0;
    int cw = 32;
    if (cw < cpg) cw = cpg;
    if (cw > C) cw = C;
    if ((cw & (cw - 1)) || C % cw || cw < 4 || cw > 256) return 0;
    // measured: worth it only while two CTAs fit on an SM and a row segment is at least 64 bytes (humanseg's 2280-row slabs
    // were 4 % slower per step in the one-launch form: one 256-thread CTA per SM cannot keep enough loads in flight)
    for (; cw >= 16 && cw >= cpg; cw >>= 1) {
        const int rl = NT / (cw >> 2);
        const size_t b = ((size_t)rows * cw + (size_t)rl * 2 * cw + 2 * cw + 2 * (cw / cpg)) * sizeof(float);
        if (b <= 100 * 1024) { *smem_bytes = b; return cw; }
    }
    return 0;
}

// ---- flat Adam over contiguous parameter / gradient / moment buffers (torch.optim.Adam semantics, no amsgrad) ----------
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    const float pi = p[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

// CUDA-graph friendly variant: the hyper-parameters and the step counter live in device memory, so a captured launch stays
// valid when the learning rate changes or the step count advances.   state = {lr, beta1, beta2, eps, weight_decay, step,
// 1 - beta1^step, sqrt(1 - beta2^step)}
__global__ void adam_tick_kernel(float* __restrict__ state) {
    pdl_enter();
    const float step = state[5] + 1.f;
    state[5] = step;
    state[6] = 1.f - powf(state[1], step);
    state[7] = sqrtf(1.f - powf(state[2], step));
}
__global__ void adam_flat_state_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                       long long n, const float* __restrict__ state) {
    pdl_enter();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lr = state[0], beta1 = state[1], beta2 = state[2], eps = state[3], weight_decay = state[4];
    const float bc1 = state[6], bc2_sqrt = state[7];
    float gi = g[i];
    const float pi = p[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}

static NormShape make_shape(int nblocks, int rows_per_block, int C, int G) {
    NormShape s;
    s.nblocks = nblocks; s.rows_per_block = rows_per_block; s.C = C; s.G = G;
    int slabs = ceil_div(nblocks == 1 ? 296 : 592, nblocks); // 2-4 CTAs per SM in flight; fewer slabs for BatchNorm: the finalize
                                                             // kernels walk them with one warp per channel (latency bound)
    const int max_slabs = max(1, rows_per_block / 64);
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs > 512) slabs = 512;
    s.rows_per_slab = ceil_div(rows_per_block, slabs);
    s.slabs = ceil_div(rows_per_block, s.rows_per_slab);
    return s;
}

}  // namespace mcnn

using namespace mcnn;

static bool norm_args_ok(int nblocks, int rows, int C, int G) {
    return nblocks > 0 && rows > 0 && C > 0 && (C % 4) == 0 && C <= 1024 && G > 0 && (C % G) == 0;
}

/* floats needed in `part` for a given shape */
extern "C" size_t mcnn_norm_ws(int nblocks, int rows_per_block, int C, int G) {
    if (!norm_args_ok(nblocks, rows_per_block, C, G)) return 0;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    return (size_t)nblocks * s.slabs * 2 * C;
}

extern "C" int mcnn_norm_fwd(const float* x, const float* a, const float* r, const float* gamma, const float* beta, float* y,
                             float* mean, float* rstd, float* part, float* running_mean, float* running_var, float momentum,
                             float eps, int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu,
                             int stats_given, void* stream) {
    if (!x || !y || !mean || !rstd || !part || !norm_args_ok(nblocks, rows_per_block, C, G)) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    const size_t smem = (size_t)(NT / (C / 4)) * 2 * C * sizeof(float);
    int rc;
    size_t fsm = 0;
    const int cw = (!stats_given && running_mean == nullptr) ? fused_cw(nblocks, rows_per_block, C, G, &fsm) : 0;
    if (cw) {
        static DynSmemGuard configured;
        if (int grc = configured.ensure(norm_fused_fwd_kernel, 200 * 1024)) return grc;
        FusedShape f; f.nblocks = nblocks; f.rows = rows_per_block; f.C = C; f.G = G; f.CW = cw;
        launch_k(norm_fused_fwd_kernel, dim3(nblocks, C / cw), NT, fsm, st, x, a, r, gamma, beta, y, mean, rstd, eps, pre_relu, post_relu, f);
        return after_launch();
    }
    if (!stats_given) {
        launch_k(norm_stats_kernel, dim3(nblocks, s.slabs), NT, smem, st, x, a, pre_relu, part, s);
        if ((rc = after_launch())) return rc;
        launch_k(norm_finalize_kernel, ceil_div(nblocks * G * 32, 256), 256, 0, st, part, mean, rstd, running_mean, running_var, momentum, eps, s);
        if ((rc = after_launch())) return rc;
    }
    launch_k(norm_apply_kernel, dim3(nblocks, s.slabs), NT, 0, st, x, a, r, gamma, beta, mean, rstd, y, pre_relu, post_relu, s);
    return after_launch();
}

extern "C" int mcnn_norm_bwd_acc(const float* dy, const float* y, const float* x, const float* a, const float* gamma,
                                 const float* mean, const float* rstd, const float* dacc, float* dx, float* dr, float* dgamma,
                                 float* dbeta, float* part, float* s1, float* s2, int nblocks, int rows_per_block, int C, int G,
                                 int pre_relu, int post_relu, void* stream);

extern "C" int mcnn_norm_bwd(const float* dy, const float* y, const float* x, const float* a, const float* gamma,
                             const float* mean, const float* rstd, float* dx, float* dr, float* dgamma, float* dbeta,
                             float* part, float* s1, float* s2, int nblocks, int rows_per_block, int C, int G, int pre_relu,
                             int post_relu, void* stream) {
    return mcnn_norm_bwd_acc(dy, y, x, a, gamma, mean, rstd, nullptr, dx, dr, dgamma, dbeta, part, s1, s2, nblocks, rows_per_block, C, G,
                             pre_relu, post_relu, stream);
}

/* the same with dx = (gradient through the block) + dacc: x also feeds another branch whose gradient is already known */
extern "C" int mcnn_norm_bwd_acc(const float* dy, const float* y, const float* x, const float* a, const float* gamma,
                                 const float* mean, const float* rstd, const float* dacc, float* dx, float* dr, float* dgamma,
                                 float* dbeta, float* part, float* s1, float* s2, int nblocks, int rows_per_block, int C, int G,
                                 int pre_relu, int post_relu, void* stream) {
    if (!dy || !x || !mean || !rstd || !dx || !part || !s1 || !s2 || !norm_args_ok(nblocks, rows_per_block, C, G)) return MCNN_E_BADARG;
    if (post_relu && !y) return MCNN_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    const size_t smem = (size_t)(NT / (C / 4)) * 2 * C * sizeof(float);
    int rc;
    size_t fsm = 0;
    const int cw = dacc != nullptr ? 0 : fused_cw(nblocks, rows_per_block, C, G, &fsm);   // the one-launch form has no dacc input
    if (cw) {
        static DynSmemGuard configured;
        if (int grc = configured.ensure(norm_fused_bwd_kernel, 200 * 1024)) return grc;
        FusedShape f; f.nblocks = nblocks; f.rows = rows_per_block; f.C = C; f.G = G; f.CW = cw;
        launch_k(norm_fused_bwd_kernel, dim3(nblocks, C / cw), NT, fsm, st, dy, y, x, a, gamma, mean, rstd, dx, dr, part, pre_relu, post_relu, f);
        if ((rc = after_launch())) return rc;
        if (dgamma != nullptr) {                               // dgamma / dbeta: the per-mesh column sums reduced over the meshes
            NormShape s1s = s; s1s.slabs = 1; s1s.rows_per_slab = rows_per_block;
            const int nfin1 = nblocks * G + C;
            launch_k(norm_bwd_finalize_kernel, ceil_div(nfin1 * 32, 256), 256, 0, st, part, gamma, s1, s2, dgamma, dbeta, s1s);
            rc = after_launch();
        }
        return rc;
    }
    launch_k(norm_bwd_stats_kernel, dim3(nblocks, s.slabs), NT, smem, st, dy, y, x, a, mean, rstd, pre_relu, post_relu, part, s);
    if ((rc = after_launch())) return rc;
    const int nfin = nblocks * G + C;
    launch_k(norm_bwd_finalize_kernel, ceil_div(nfin * 32, 256), 256, 0, st, part, gamma, s1, s2, dgamma, dbeta, s);
    if ((rc = after_launch())) return rc;
    launch_k(norm_bwd_apply_kernel, dim3(nblocks, s.slabs), NT, 0, st, dy, y, x, a, gamma, mean, rstd, s1, s2, dx, dr, dacc, pre_relu, post_relu, s);
    return after_launch();
}

/* ---- the statistics and apply stages on their own: what a cross-rank (SyncBN) reduction sits between ------------------ */
extern "C" int mcnn_norm_slabs(int nblocks, int rows_per_block, int C, int G) {
    if (!norm_args_ok(nblocks, rows_per_block, C, G)) return 0;
    return make_shape(nblocks, rows_per_block, C, G).slabs;
}

extern "C" int mcnn_norm_stats(const float* x, const float* a, float* part, int nblocks, int rows_per_block, int C, int G,
                               int pre_relu, void* stream) {
    if (!x || !part || !norm_args_ok(nblocks, rows_per_block, C, G)) return MCNN_E_BADARG;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    const size_t smem = (size_t)(NT / (C / 4)) * 2 * C * sizeof(float);
    launch_k(norm_stats_kernel, dim3(nblocks, s.slabs), NT, smem, (cudaStream_t)stream, x, a, pre_relu, part, s);
    return after_launch();
}

extern "C" int mcnn_norm_bwd_stats(const float* dy, const float* y, const float* x, const float* a, const float* mean,
                                   const float* rstd, float* part, int nblocks, int rows_per_block, int C, int G, int pre_relu,
                                   int post_relu, void* stream) {
    if (!dy || !x || !mean || !rstd || !part || !norm_args_ok(nblocks, rows_per_block, C, G)) return MCNN_E_BADARG;
    if (post_relu && !y) return MCNN_E_BADARG;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    const size_t smem = (size_t)(NT / (C / 4)) * 2 * C * sizeof(float);
    launch_k(norm_bwd_stats_kernel, dim3(nblocks, s.slabs), NT, smem, (cudaStream_t)stream, dy, y, x, a, mean, rstd, pre_relu, post_relu, part, s);
    return after_launch();
}

extern "C" int mcnn_norm_bwd_apply(const float* dy, const float* y, const float* x, const float* a, const float* gamma,
                                   const float* mean, const float* rstd, const float* s1, const float* s2, float* dx, float* dr,
                                   int nblocks, int rows_per_block, int C, int G, int pre_relu, int post_relu, void* stream) {
    if (!dy || !x || !mean || !rstd || !s1 || !s2 || !dx || !norm_args_ok(nblocks, rows_per_block, C, G)) return MCNN_E_BADARG;
    if (post_relu && !y) return MCNN_E_BADARG;
    const NormShape s = make_shape(nblocks, rows_per_block, C, G);
    launch_k(norm_bwd_apply_kernel, dim3(nblocks, s.slabs), NT, 0, (cudaStream_t)stream, dy, y, x, a, gamma, mean, rstd, s1, s2, dx, dr, (const float*)nullptr, pre_relu, post_relu, s);
    return after_launch();
}

extern "C" int mcnn_debug_set_norm_fused(int on) { g_norm_fused = on ? 1 : 0; return MCNN_OK; }

extern "C" int mcnn_adam_flat_state(float* p, const float* g, float* m, float* v, long long n, float* state, void* stream) {
    if (!p || !g || !m || !v || !state || n <= 0) return MCNN_E_BADARG;
    launch_k(adam_tick_kernel, 1, 1, 0, (cudaStream_t)stream, state);
    int rc = after_launch();
    if (rc) return rc;
    launch_k(adam_flat_state_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, p, g, m, v, n, state);
    return after_launch();
}

extern "C" int mcnn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, float weight_decay, int step, void* stream) {
    if (!p || !g || !m || !v || n <= 0 || step <= 0) return MCNN_E_BADARG;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2 = 1.f - powf(beta2, (float)step);
    launch_k(adam_flat_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                                     bc1, sqrtf(bc2));
    return after_launch();
}
