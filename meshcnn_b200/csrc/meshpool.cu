// MeshPool on the GPU: one CTA per mesh.
//
// The reference (models/layers/mesh_pool.py:41-203, mesh.py:26-71, mesh_union.py) pops edges off a heap in
// ascending (squared feature norm, id) order and collapses them one at a time in Python, mutating NumPy
// topology arrays, Python vertex->edge lists and a dense n x n group matrix.  Here the whole state of one mesh
// lives in the shared memory of one CTA:
//
//   gemm / sides / mask      int16[4n] / int8[4n] / uint8[n]      one-ring tables of the incoming level
//   ev + rep                 original endpoints + a vertex union-find: `edges[edges == v1] = v0` over ALL rows
//                            (mesh.py:35-37) is exactly rep[v1] = v0
//   ve lists                 chains of slabs.  Every vertex owns the contiguous slab it arrives with (the CSR row of
//                            the incoming level); `ve[v0].extend(ve[v1])` links v1's chain behind v0's (O(1)).
//                            `list.remove(x)` turns the entry into a tombstone: every edge knows where its two entries
//                            sit (epos), so a removal is two stores; only id 0 -- which the previous level's re-indexing
//                            gives to every stale entry -- can occur more than once in a list and is
//                            searched for its first occurrence.  Stale entries are kept, as in the reference.  Slabs
//                            are arrays, so a warp walks a list 32 entries at a time (one-ring test).
//   MeshUnion                only the (source -> target) log of `groups[t] += groups[s]`.  Every source dies within
//                            the attempt that pours it and receives nothing between its (at most two) pours, so the
//                            rows are the reachability sets of a DAG with out-degree <= 2: row t = {ancestors of t},
//                            un-clamped occurrences of m = number of paths starting at m.  Both are
//                            evaluated after the collapse by all threads.
//
// Phases: load -> stable radix sort of the norm bits (ties stay in id order) -> the collapse, in ROUNDS over a window of the next
// pending candidates (one per warp): every warp inspects its candidate read-only on the committed state and reserves the
// rows a collapse would write; the longest prefix of the window whose members touch nothing an earlier member writes is
// committed in parallel -- each of them read exactly what it would have read in the strict one-at-a-time order, so pops,
// outcomes, topology, lists and the union log are those of the reference, bit for bit.  A candidate that has to rewire
// around a valence-3 vertex (__get_invalids finds a shared edge; a few per cent) runs alone through the sequential code
// path, failed attempts and their persistent side effects included -> all threads run Mesh.clean (compaction, re-indexing
// with dead -> 0) and emit the groups as CSR + CSC + un-clamped occurrences.
#include "common.cuh"

namespace mcnn {

// Two instantiations of the same kernel:
//   Narrow  16-bit ids, state in dynamic shared memory, 256 threads            (meshes up to a few thousand edges)
//   Wide    32-bit ids, state in a per-mesh global-memory workspace (L2), 512 threads  (no size limit: 170 k-edge meshes)
struct Narrow {
    typedef int16_t idx_t;
    typedef uint16_t uidx_t;
    typedef uint32_t pair_t;
    static constexpr unsigned NIL = 0xFFFFu;
    static constexpr int SHIFT = 16;
    static constexpr int THREADS = 256;
    static constexpr bool WIDE = false;
};
struct Wide {
    typedef int32_t idx_t;
    typedef uint32_t uidx_t;
    typedef unsigned long long pair_t;
    static constexpr unsigned NIL = 0xFFFFFFFFu;
    static constexpr int SHIFT = 32;
    static constexpr int THREADS = 512;
    static constexpr bool WIDE = true;
};
template <class P> __device__ __forceinline__ typename P::pair_t pr_pack(int a, int b) {
    return (typename P::pair_t)(typename P::uidx_t)a | ((typename P::pair_t)(typename P::uidx_t)b << P::SHIFT);
}
template <class P> __device__ __forceinline__ int pr_lo(typename P::pair_t w) { return (int)(typename P::uidx_t)w; }
template <class P> __device__ __forceinline__ int pr_hi(typename P::pair_t w) { return (int)(typename P::uidx_t)(w >> P::SHIFT); }

constexpr int POOL_RSET_CAP = 256;       // vertex ids one inspection may read through the two vertex lists (2 per entry); more -> runs alone
constexpr int POOL_BFS_CAP = 96;         // members of one MeshUnion row / column walked per thread (implementation limit)

struct PoolLayout {
    size_t vs, gemm, evp, sides, mask, order, rep, own_e, own_v, wstamp, rset, slab_start, slab_len, slab_next, chain_tail, vmask, ve, eposa,
        eposb, ulog, cnt32, keys, scan, total;
    int npow2, ulog_cap;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <class P>
__host__ __device__ inline PoolLayout pool_layout(int n, int V, int ve_cap, bool with_vs) {
    typedef typename P::idx_t idx_t;
    typedef typename P::uidx_t uidx_t;
    typedef typename P::pair_t pair_t;
    PoolLayout L;
    size_t o = 0;
    L.vs = o; o += with_vs ? sizeof(double) * 3 * (size_t)V : 0; o = align_up(o, 16);
    L.gemm = o; o += sizeof(idx_t) * 4 * (size_t)n; o = align_up(o, 16);          // later: column member pool (u16[3n]) + col_off
    L.evp = o; o += sizeof(pair_t) * (size_t)n; o = align_up(o, 16);            // evp + sides, later: row member pool (u16[3n]) + ...
    L.sides = o; o += 4 * (size_t)n; o = align_up(o, 16);
    L.mask = o; o += (size_t)n; o = align_up(o, 16);
    L.order = o; o += sizeof(uidx_t) * (size_t)n; o = align_up(o, 16);
    L.rep = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.own_e = o; o += sizeof(uint32_t) * (size_t)n; o = align_up(o, 16);          // write reservations of the current round (edge rows)
    L.own_v = o; o += sizeof(uint32_t) * (size_t)V; o = align_up(o, 16);          //   ... (vertex rows: list, union-find root, v_mask)
    L.wstamp = o; o += sizeof(uint16_t) * (size_t)V * (P::THREADS / 32); o = align_up(o, 16);   // one-ring stamps, one array per warp
    L.rset = o; o += sizeof(uidx_t) * (size_t)POOL_RSET_CAP * (P::THREADS / 32); o = align_up(o, 16);   // vertices an inspection read
    L.slab_start = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.slab_len = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.slab_next = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.chain_tail = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.vmask = o; o += (size_t)V; o = align_up(o, 16);
    L.ve = o; o += sizeof(uidx_t) * ((size_t)ve_cap + 3 * (size_t)n); o = align_up(o, 16);   // + room for re-packed lists
    L.eposa = o; o += sizeof(uidx_t) * (size_t)n; o = align_up(o, 16);
    L.eposb = o; o += sizeof(uidx_t) * (size_t)n; o = align_up(o, 16);
    L.ulog_cap = 2 * n + 16;                     // every union's source dies within the same attempt: <= 2 per dead edge
    L.ulog = o; o += sizeof(pair_t) * (size_t)L.ulog_cap; o = align_up(o, 16);
    L.cnt32 = o; o += sizeof(int) * 2 * (size_t)n; o = align_up(o, 16);           // out-degree, row length / cursor
    L.npow2 = 0;
    // the radix-sort ping-pong buffers (2 x (u32 key + id)[n]) are dead before the group evaluation needs occ (int32[n]) +
    // out0 / out1 (id[n] each), which live in the same place
    L.keys = o; o += (2 * sizeof(uint32_t) + 2 * sizeof(uidx_t)) * (size_t)n; o = align_up(o, 16);
    L.scan = o; o += sizeof(int) * 48; o = align_up(o, 16);
    L.total = o;
    return L;
}

// exclusive block scan of one int per thread; returns the prefix and (to all threads) the block total
template <int NT>
__device__ __forceinline__ int block_scan_excl(int v, int* scratch, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) scratch[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < NT / 32) ? scratch[lane] : 0;
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        scratch[lane] = wi - w;                 // exclusive warp offsets
        if (lane == 31) scratch[32] = wi;       // block total
    }
    __syncthreads();
    int res = scratch[wid] + incl - v;
    total = scratch[32];
    __syncthreads();
    return res;
}


// One row of the one-ring table with a single load: 4 x int16 as one 64-bit word (Narrow) or 4 x int32 as an int4 (Wide).
template <class P> struct Row4;
template <> struct Row4<Narrow> {
    typedef unsigned long long T;
    static __device__ __forceinline__ T load(const int16_t* g, int e) { return *reinterpret_cast<const unsigned long long*>(g + 4 * e); }
    static __device__ __forceinline__ bool neg(T r) { return (r & 0x8000800080008000ULL) != 0ULL; }       // any entry == -1
    static __device__ __forceinline__ T join(T a, T b) { return a | b; }
    static __device__ __forceinline__ int get(T r, int j) { return (int)((r >> (16 * j)) & 0xFFFFULL); }  // entry j (non-negative)
};
template <> struct Row4<Wide> {
    typedef int4 T;
    static __device__ __forceinline__ T load(const int32_t* g, int e) { return *reinterpret_cast<const int4*>(g + 4 * e); }
    static __device__ __forceinline__ bool neg(T r) { return (r.x | r.y | r.z | r.w) < 0; }
    static __device__ __forceinline__ T join(T a, T b) { return make_int4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
    static __device__ __forceinline__ int get(T r, int j) { return j == 0 ? r.x : (j == 1 ? r.y : (j == 2 ? r.z : r.w)); }
};

// Per-mesh collapse state (shared memory for Narrow, the L2-resident workspace for Wide).
template <class P>
struct PoolState {
    typename P::idx_t* gemm;
    int8_t* sides;
    uint8_t* mask;
    typename P::pair_t *evp, *ulog;
    typename P::uidx_t *rep, *slab_start, *slab_len, *slab_next, *chain_tail, *ve, *eposa, *eposb;
    uint8_t* vmask;
    double* vs;               // or nullptr
    uint32_t *own_e, *own_v;  // write reservations: (round << 5) | (31 - window position)
    int* bump;                // free room behind the lists (entries), used by warp_repack
    int bump_end;
};

template <class U>
__device__ __forceinline__ int uf_find(U* rep, int v) {
    int p = rep[v];
    while (p != v) {
        const int gp = rep[p];
        rep[v] = (U)gp;                                 // path halving; concurrent readers only ever write ancestors
        v = gp;
        p = rep[v];
    }
    return v;
}
template <class P>
__device__ __forceinline__ void edge_ends(const typename P::pair_t* evp, typename P::uidx_t* rep, int e, int& v0, int& v1) {
    const typename P::pair_t w = evp[e];
    const int a = pr_lo<P>(w), b = pr_hi<P>(w);
    const int ra = (int)rep[a], rb = (int)rep[b];                 // both loads in flight before either branch
    v0 = (ra == a) ? a : uf_find(rep, a);
    v1 = (rb == b) ? b : uf_find(rep, b);
}

// A list that has absorbed another one (`extend`) is a chain of slabs; every later walk pays one dependent round trip per
// slab.  The warp that commits a merge re-packs the chain into ONE slab in the free room behind the lists: order preserved,
// tombstones dropped, the positions of the moved entries (epos) updated.  The caller owns vertex v (nobody else reads it).
template <class P>
__device__ __forceinline__ void warp_repack(const PoolState<P>& w, int lane, int v) {
    if ((unsigned)w.slab_next[v] == P::NIL) return;
    int need = 0;
    for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)w.slab_next[s]) need += (int)w.slab_len[s];
    int out = -1;
    if (lane == 0) {
        out = atomicAdd(w.bump, need);
        if (out + need > w.bump_end) { atomicSub(w.bump, need); out = -1; }     // no room left: keep the chain
    }
    out = __shfl_sync(0xffffffffu, out, 0);
    if (out < 0) return;
    const unsigned lt = (1u << lane) - 1u;
    int total = 0;
    for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)w.slab_next[s]) {
        const int start = w.slab_start[s], len = w.slab_len[s];
        for (int base = 0; base < len; base += 32) {
            const int i = base + lane;
            const unsigned x = (i < len) ? (unsigned)w.ve[start + i] : P::NIL;
            const bool live = x != P::NIL;
            const unsigned m = __ballot_sync(0xffffffffu, live);
            if (live) {
                const int dst = out + total + __popc(m & lt);
                w.ve[dst] = (typename P::uidx_t)x;
                if (x != 0u) {
                    if ((int)w.eposa[x] == start + i) w.eposa[x] = (typename P::uidx_t)dst;
                    else if ((int)w.eposb[x] == start + i) w.eposb[x] = (typename P::uidx_t)dst;
                }
            }
            total += __popc(m);
        }
    }
    __syncwarp();
    if (lane == 0) {
        w.slab_start[v] = (typename P::uidx_t)out;
        w.slab_len[v] = (typename P::uidx_t)total;
        w.slab_next[v] = (typename P::uidx_t)P::NIL;
        w.chain_tail[v] = (typename P::uidx_t)v;
    }
    __syncwarp();
}

// __is_one_ring_valid (mesh_pool.py:95-100), by one warp, read-only on the mesh state: the number of distinct vertices other
// than v0 / v1 that occur as an endpoint of an entry of BOTH lists (the lists may hold stale / dead edge ids).
// stamp: this warp's own array (tag = endpoint seen from v0; tag | 1 = already counted), so warps do not disturb each other.
// Every endpoint read is recorded in rset (the vertices this inspection depends on) unless rset is NULL; returns -1 if they
// do not fit.
template <class P>
__device__ __forceinline__ int warp_one_ring(const PoolState<P>& w, uint16_t* stamp, unsigned tagA, int lane, int v0, int v1,
                                             typename P::uidx_t* rset, int& nrset) {
    typedef typename P::uidx_t uidx_t;
    const uint16_t tA = (uint16_t)tagA, tB = (uint16_t)(tagA | 1u);
    const unsigned lt = (1u << lane) - 1u;
    int step = 0;
    bool over = false;
    for (unsigned s = (unsigned)v0; s != P::NIL; s = (unsigned)w.slab_next[s]) {
        const int start = w.slab_start[s], len = w.slab_len[s];
        for (int base = 0; base < len; base += 32, ++step) {
            const int i = base + lane;
            const unsigned x = (i < len) ? (unsigned)w.ve[start + i] : P::NIL;
            int a = (int)P::NIL, b = (int)P::NIL;
            if (x != P::NIL) {
                edge_ends<P>(w.evp, w.rep, (int)x, a, b);
                stamp[a] = tA;
                stamp[b] = tA;
            }
            if (rset != nullptr) {
                if (step * 64 + 64 <= POOL_RSET_CAP) { rset[step * 64 + 2 * lane] = (uidx_t)a; rset[step * 64 + 2 * lane + 1] = (uidx_t)b; }
                else over = true;
            }
        }
    }
    __syncwarp();
    int shared = 0;
    for (unsigned s = (unsigned)v1; s != P::NIL; s = (unsigned)w.slab_next[s]) {
        const int start = w.slab_start[s], len = w.slab_len[s];
        for (int base = 0; base < len; base += 32, ++step) {
            const int i = base + lane;
            const unsigned x = (i < len) ? (unsigned)w.ve[start + i] : P::NIL;
            int a = (int)P::NIL, b = (int)P::NIL;
            if (x != P::NIL) edge_ends<P>(w.evp, w.rep, (int)x, a, b);
            if (rset != nullptr) {
                if (step * 64 + 64 <= POOL_RSET_CAP) { rset[step * 64 + 2 * lane] = (uidx_t)a; rset[step * 64 + 2 * lane + 1] = (uidx_t)b; }
                else over = true;
            }
#pragma unroll
            for (int which = 0; which < 2; ++which) {                      // first endpoints, then second endpoints
                const int c = which ? b : a;
                const bool hit = (x != P::NIL) && c != v0 && c != v1 && stamp[c] == tA;
                const unsigned peers = __match_any_sync(0xffffffffu, hit ? c : -1 - lane);
                const bool lead = hit && (peers & lt) == 0u;               // one lane per distinct vertex counts it ...
                shared += __popc(__ballot_sync(0xffffffffu, lead));
                if (lead) stamp[c] = tB;                                   // ... once
                __syncwarp();
            }
        }
    }
    nrset = min(step * 64, POOL_RSET_CAP);
    return over ? -1 : shared;
}

// The common case of the above: both lists are ONE slab of at most 32 entries (the committing warp re-packs a merged list,
// so this holds until the spare room runs out).  The two lists are fetched and resolved side by side -- their dependent load
// chains (descriptor -> entry -> endpoints -> union-find roots) overlap -- and the endpoints stay in registers: a0, b0 /
// a1, b1 are the roots this lane read through entry `lane` of ve[v0] / ve[v1] (P::NIL: no entry), x0 / x1 the entries.
// Returns the shared-vertex count, or -1 if the lists do not have that form (the caller takes the general walk).
template <class P>
__device__ __forceinline__ int warp_one_ring_fast(const PoolState<P>& w, uint16_t* stamp, unsigned tagA, int lane, int v0, int v1,
                                                  int& a0, int& b0, int& a1, int& b1, int& start0, int& len0, int& start1, int& len1) {
    const uint16_t tA = (uint16_t)tagA, tB = (uint16_t)(tagA | 1u);
    const unsigned lt = (1u << lane) - 1u;
    const unsigned next0 = (unsigned)w.slab_next[v0], next1 = (unsigned)w.slab_next[v1];
    start0 = w.slab_start[v0]; len0 = w.slab_len[v0];
    start1 = w.slab_start[v1]; len1 = w.slab_len[v1];
    if (next0 != P::NIL || next1 != P::NIL || len0 > 32 || len1 > 32) return -1;
    const unsigned x0 = (lane < len0) ? (unsigned)w.ve[start0 + lane] : P::NIL;
    const unsigned x1 = (lane < len1) ? (unsigned)w.ve[start1 + lane] : P::NIL;
    const typename P::pair_t w0 = w.evp[x0 != P::NIL ? x0 : 0u], w1 = w.evp[x1 != P::NIL ? x1 : 0u];   // unconditional, safe index
    a0 = pr_lo<P>(w0); b0 = pr_hi<P>(w0); a1 = pr_lo<P>(w1); b1 = pr_hi<P>(w1);
    // the four roots together, one round trip per level of the union-find forest for the whole warp (four separate find loops
    // are four divergent chains of three dependent loads each)
    while (true) {
        const int ra0 = (int)w.rep[a0], rb0 = (int)w.rep[b0], ra1 = (int)w.rep[a1], rb1 = (int)w.rep[b1];
        const bool fixed = (ra0 == a0) & (rb0 == b0) & (ra1 == a1) & (rb1 == b1);
        a0 = ra0; b0 = rb0; a1 = ra1; b1 = rb1;
        if (__all_sync(0xffffffffu, fixed)) break;
    }
    if (x0 != P::NIL) { stamp[a0] = tA; stamp[b0] = tA; } else { a0 = (int)P::NIL; b0 = (int)P::NIL; }
    if (x1 == P::NIL) { a1 = (int)P::NIL; b1 = (int)P::NIL; }
    __syncwarp();
    // every lane tests its two endpoints; then each hitting (lane, endpoint) writes its own mark over the stamp and the one
    // mark that survives per vertex counts it: distinct vertices are counted once without a warp-wide match
    const bool hit_a = (x1 != P::NIL) && a1 != v0 && a1 != v1 && stamp[a1] == tA;
    const bool hit_b = (x1 != P::NIL) && b1 != v0 && b1 != v1 && stamp[b1] == tA;
    __syncwarp();
    const uint16_t mark = (uint16_t)(tagA | 64u | (unsigned)(2 * lane));
    if (hit_a) stamp[a1] = mark;
    if (hit_b) stamp[b1] = (uint16_t)(mark | 1u);
    __syncwarp();
    const bool win_a = hit_a && stamp[a1] == mark, win_b = hit_b && stamp[b1] == (uint16_t)(mark | 1u);
    (void)tB; (void)lt;
    return __popc(__ballot_sync(0xffffffffu, win_a)) + __popc(__ballot_sync(0xffffffffu, win_b));
}

// The committing warp's re-pack of a list it has just merged, for two single-slab lists (see warp_one_ring_fast): the live
// entries of the slab of v0, then those of v1, go to fresh room as one slab of v0.  Falls back to the general chain walk.
template <class P>
__device__ __forceinline__ void warp_repack_pair(const PoolState<P>& w, int lane, int v0, int v1, int start0, int len0, int start1,
                                                 int len1, int& priv_bump, int priv_end) {
    typedef typename P::uidx_t uidx_t;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned x0 = (lane < len0) ? (unsigned)w.ve[start0 + lane] : P::NIL;
    const unsigned x1 = (lane < len1) ? (unsigned)w.ve[start1 + lane] : P::NIL;
    const unsigned m0 = __ballot_sync(0xffffffffu, x0 != P::NIL), m1 = __ballot_sync(0xffffffffu, x1 != P::NIL);
    const int n0 = __popc(m0), need = n0 + __popc(m1);
    int out = -1;
    if (priv_bump + need <= priv_end) {                                          // this warp's own room: no atomics
        out = priv_bump;
        priv_bump += need;
    } else {
        if (lane == 0) {
            out = atomicAdd(w.bump, need);
            if (out + need > w.bump_end) { atomicSub(w.bump, need); out = -1; } // no room left: keep the chain
        }
        out = __shfl_sync(0xffffffffu, out, 0);
        if (out < 0) return;
    }
    if (x0 != P::NIL) {
        const int dst = out + __popc(m0 & lt);
        w.ve[dst] = (uidx_t)x0;
        if (x0 != 0u) {
            if ((int)w.eposa[x0] == start0 + lane) w.eposa[x0] = (uidx_t)dst;
            else if ((int)w.eposb[x0] == start0 + lane) w.eposb[x0] = (uidx_t)dst;
        }
    }
    if (x1 != P::NIL) {
        const int dst = out + n0 + __popc(m1 & lt);
        w.ve[dst] = (uidx_t)x1;
        if (x1 != 0u) {
            if ((int)w.eposa[x1] == start1 + lane) w.eposa[x1] = (uidx_t)dst;
            else if ((int)w.eposb[x1] == start1 + lane) w.eposb[x1] = (uidx_t)dst;
        }
        // edges[edges == v1] = v0 (mesh.py:33) applied to the stored endpoints of v1's edges right away, so that later walks find
        // roots at the first look-up (the union-find stays authoritative; only this candidate's reservation covers these words)
        const typename P::pair_t w1 = w.evp[x1];
        const int lo = pr_lo<P>(w1), hi = pr_hi<P>(w1);
        if (lo == v1 || hi == v1) w.evp[x1] = pr_pack<P>(lo == v1 ? v0 : lo, hi == v1 ? v0 : hi);
    }
    if (lane == 0) {
        w.slab_start[v0] = (uidx_t)out;
        w.slab_len[v0] = (uidx_t)need;
        w.slab_next[v0] = (uidx_t)P::NIL;
        w.chain_tail[v0] = (uidx_t)v0;
    }
    __syncwarp();
}

// list.remove(x) on ve[v] (mesh.py:48) by search (one thread): the first occurrence in list order becomes a tombstone.
// Only needed for id 0 -- which the previous level's re-indexing gives to every stale entry -- and for inconsistent input.
template <class P>
__device__ bool list_remove_scan(const PoolState<P>& w, int v, int x) {
    for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)w.slab_next[s]) {
        const int start = w.slab_start[s], len = w.slab_len[s];
        for (int i = 0; i < len; ++i)
            if ((int)w.ve[start + i] == x) { w.ve[start + i] = (typename P::uidx_t)P::NIL; return true; }
    }
    return false;
}

// The reference's control flow for ONE candidate, run by one thread on state nobody else touches meanwhile.
template <class P>
struct Collapse {
    typedef typename P::idx_t idx_t;
    typedef typename P::uidx_t uidx_t;
    typedef typename P::pair_t pair_t;
    PoolState<P> st;
    idx_t* gemm;
    int8_t* sides;
    uint8_t* mask;
    int32_t* log_unions;      // global or nullptr
    int union_cap;
    int nunions, ulog_cap;    // nunions: where this candidate's unions go in the log (strict order)
    int ec, T, status;

    __device__ __forceinline__ void ends(int e, int& v0, int& v1) { edge_ends<P>(st.evp, st.rep, e, v0, v1); }

    // MeshUnion.union (mesh_union.py:11-12) is only logged here; the rows are evaluated after the collapse
    __device__ __forceinline__ void group_union(int s, int t) {
        if (log_unions != nullptr && nunions < union_cap) { log_unions[2 * nunions] = s; log_unions[2 * nunions + 1] = t; }
        if (s == t) status = MCNN_ST_DEGENERATE;
        if (nunions < ulog_cap) st.ulog[nunions] = pr_pack<P>(s, t);
        else status = MCNN_ST_GROUP_OVERFLOW;
        ++nunions;
    }

    // MeshPool.has_boundaries (mesh_pool.py:87-92)
    __device__ bool has_boundaries(int e) const {
        if (!P::WIDE) {
            const unsigned long long row = *reinterpret_cast<const unsigned long long*>(gemm + 4 * e);
            if ((row & 0x8000800080008000ULL) != 0ULL) return true;
            const unsigned long long r0 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)(row & 0xFFFF));
            const unsigned long long r1 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 16) & 0xFFFF));
            const unsigned long long r2 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 32) & 0xFFFF));
            const unsigned long long r3 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 48) & 0xFFFF));
            return ((r0 | r1 | r2 | r3) & 0x8000800080008000ULL) != 0ULL;
        } else {
            const int4 row = *reinterpret_cast<const int4*>(gemm + 4 * e);
            if ((row.x | row.y | row.z | row.w) < 0) return true;
            const int4 r0 = *reinterpret_cast<const int4*>(gemm + 4 * row.x);
            const int4 r1 = *reinterpret_cast<const int4*>(gemm + 4 * row.y);
            const int4 r2 = *reinterpret_cast<const int4*>(gemm + 4 * row.z);
            const int4 r3 = *reinterpret_cast<const int4*>(gemm + 4 * row.w);
            return ((r0.x | r0.y | r0.z | r0.w | r1.x | r1.y | r1.z | r1.w | r2.x | r2.y | r2.z | r2.w | r3.x | r3.y | r3.z | r3.w) < 0);
        }
    }

    struct Face { int ka, kb, sa, sb, osa, osb, oka0, oka1, okb0, okb1; };

    // __get_face_info (mesh_pool.py:160-170)
    __device__ __forceinline__ Face face_info(int e, int side) const {
        Face f;
        f.ka = gemm[4 * e + side];
        f.kb = gemm[4 * e + side + 1];
        f.sa = sides[4 * e + side];
        f.sb = sides[4 * e + side + 1];
        f.osa = (f.sa - (f.sa & 1) + 2) & 3;
        f.osb = (f.sb - (f.sb & 1) + 2) & 3;
        f.oka0 = gemm[4 * f.ka + f.osa];
        f.oka1 = gemm[4 * f.ka + f.osa + 1];
        f.okb0 = gemm[4 * f.kb + f.osb];
        f.okb1 = gemm[4 * f.kb + f.osb + 1];
        return f;
    }

    // __redirect_edges (mesh_pool.py:140-145)
    __device__ __forceinline__ void redirect(int a, int sa, int b, int sb) {
        gemm[4 * a + sa] = (idx_t)b;
        gemm[4 * b + sb] = (idx_t)a;
        sides[4 * a + sa] = (int8_t)sb;
        sides[4 * b + sb] = (int8_t)sa;
    }

    __device__ __forceinline__ static int other_side(int s) { return s + 1 - 2 * (s & 1); }

    // __get_invalids (mesh_pool.py:115-138): returns 0 or 3 edges; rewires around a valence-3 vertex
    __device__ int get_invalids(int e, int side, int& t0, int& t1, int& t2) {
        const Face f = face_info(e, side);
        int nsh = 0, mid = 0, ua = 0, ub = 0, usa = 0, usb = 0;
        // shared items of (oka0, oka1) x (okb0, okb1), first match in the reference's iteration order
        if (f.oka0 == f.okb0) { if (nsh == 0) { mid = f.oka0; ua = f.oka1; ub = f.okb1; usa = 1; usb = 1; } ++nsh; }
        if (f.oka0 == f.okb1) { if (nsh == 0) { mid = f.oka0; ua = f.oka1; ub = f.okb0; usa = 1; usb = 0; } ++nsh; }
        if (f.oka1 == f.okb0) { if (nsh == 0) { mid = f.oka1; ua = f.oka0; ub = f.okb1; usa = 0; usb = 1; } ++nsh; }
        if (f.oka1 == f.okb1) { if (nsh == 0) { mid = f.oka1; ua = f.oka0; ub = f.okb0; usa = 0; usb = 0; } ++nsh; }
        if (nsh == 0) return 0;
        if (nsh != 1) { status = MCNN_ST_SHARED_ASSERT; return 0; }
        usa = sides[4 * f.ka + f.osa + usa];
        usb = sides[4 * f.kb + f.osb + usb];
        redirect(e, side, ua, usa);
        redirect(e, side + 1, ub, usb);
        redirect(ua, other_side(usa), ub, other_side(usb));
        group_union(f.ka, e);
        group_union(f.kb, e);
        group_union(f.ka, ua);
        group_union(mid, ua);
        group_union(f.kb, ub);
        group_union(mid, ub);
        t0 = f.ka; t1 = f.kb; t2 = mid;
        return 3;
    }

    // __remove_triplete (mesh_pool.py:172-182): no remove_edge, the ve entries go stale
    __device__ void remove_triplet(int t0, int t1, int t2) {
        int a0, a1, b0, b1, c0, c1;
        ends(t0, a0, a1);
        ends(t1, b0, b1);
        ends(t2, c0, c1);
        mask[t0] = 0; mask[t1] = 0; mask[t2] = 0;
        ec -= 3;
        int count = 0, v = -1;
        if ((a0 == b0 || a0 == b1) && (a0 == c0 || a0 == c1)) { ++count; v = a0; }
        if (a1 != a0 && (a1 == b0 || a1 == b1) && (a1 == c0 || a1 == c1)) { ++count; v = a1; }
        if (count != 1) { status = MCNN_ST_VERTEX_ASSERT; return; }
        st.vmask[v] = 0;
    }

    // __clean_side (mesh_pool.py:74-85)
    __device__ bool clean_side(int e, int side) {
        if (ec <= T) return false;
        int t0, t1, t2;
        int n = get_invalids(e, side, t0, t1, t2);
        while (n != 0 && ec > T) {
            if (status) return false;
            remove_triplet(t0, t1, t2);
            if (status) return false;
            if (ec <= T) return false;
            if (has_boundaries(e)) return false;
            n = get_invalids(e, side, t0, t1, t2);
        }
        return status == 0;
    }

    // Mesh.remove_edge (mesh.py:42-48): both list entries of x become tombstones
    __device__ void remove_edge(int x) {
        const unsigned ia = st.eposa[x], ib = st.eposb[x];
        if (x != 0 && ia != P::NIL && ib != P::NIL && (int)st.ve[ia] == x && (int)st.ve[ib] == x) {
            st.ve[ia] = (uidx_t)P::NIL;
            st.ve[ib] = (uidx_t)P::NIL;
            return;
        }
        int a, b;
        ends(x, a, b);
        if (!list_remove_scan<P>(st, a, x) || !list_remove_scan<P>(st, b, x)) status = MCNN_ST_VE_REMOVE;
    }

    // __pool_side (mesh_pool.py:102-113)
    __device__ void pool_side(int e, int side) {
        const Face f = face_info(e, side);
        const int base = f.sa - (f.sa & 1);
        redirect(f.ka, base, f.okb0, sides[4 * f.kb + f.osb]);
        redirect(f.ka, base + 1, f.okb1, sides[4 * f.kb + f.osb + 1]);     // read after the first rewiring, as the reference does
        group_union(f.kb, f.ka);
        group_union(e, f.ka);
        mask[f.kb] = 0;
        remove_edge(f.kb);
        ec -= 1;
    }

    // Mesh.merge_vertices (mesh.py:26-37); returns the surviving vertex (its list is now a chain) or -1
    __device__ int merge_vertices(int e) {
        remove_edge(e);
        int v0, v1;
        ends(e, v0, v1);
        if (v0 == v1) { status = MCNN_ST_DEGENERATE; return -1; }
        if (st.vs != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) st.vs[3 * v0 + k] = (st.vs[3 * v0 + k] + st.vs[3 * v1 + k]) / 2;
        }
        st.vmask[v1] = 0;
        // ve[v0].extend(ve[v1]) : v1's chain moves behind v0's (ve[v1] is unreachable once every row holding v1 reads v0)
        st.slab_next[(unsigned)st.chain_tail[v0]] = (uidx_t)v1;
        st.chain_tail[v0] = st.chain_tail[v1];
        st.rep[v1] = (uidx_t)v0;         // edges[edges == v1] = v0 over every row
        return v0;
    }

    // the tail of __pool_edge (mesh_pool.py:67-71) for a candidate that passed every test
    __device__ int collapse(int e, int& merged) {
        merged = -1;
        pool_side(e, 0);
        if (status) return MCNN_POP_COLLAPSED;
        pool_side(e, 2);
        if (status) return MCNN_POP_COLLAPSED;
        merged = merge_vertices(e);
        mask[e] = 0;
        ec -= 1;
        return MCNN_POP_COLLAPSED;
    }

    // __pool_edge (mesh_pool.py:58-72), complete: used for the candidates that run alone.  ring_shared: the one-ring count,
    // taken beforehand -- it depends on the vertex lists, the endpoints and the union-find only, none of which
    // __clean_side touches.
    __device__ int pool_edge(int e, int ring_shared, int& merged) {
        merged = -1;
        if (has_boundaries(e)) return MCNN_POP_REJ_BOUNDARY;
        if (!clean_side(e, 0)) return MCNN_POP_REJ_SIDE0;
        if (!clean_side(e, 2)) return MCNN_POP_REJ_SIDE2;
        if (ring_shared != 2) return MCNN_POP_REJ_ONE_RING;
        return collapse(e, merged);
    }
};

// verdicts of the read-only inspection of a candidate
enum { INSP_REJ_BOUNDARY = 1, INSP_REJ_ONE_RING = 2, INSP_COLLAPSE = 3, INSP_ALONE = 4 };
constexpr int POOL_MAX_WARPS = 32;

struct RoundCtl {                       // shared by the warps of the CTA
    int verdict[POOL_MAX_WARPS];        // INSP_* of the window members, 0 = no member
    int ok[POOL_MAX_WARPS];             // 1 = touches nothing an earlier member writes
    int pos[POOL_MAX_WARPS];            // position of the member in the sorted order (its pop index)
    unsigned bits[2][4];                // per round parity: members | COLLAPSE verdicts | ALONE verdicts | clash-free members
    int ec, nunions, status, bump;
};

template <class P>
__global__ void __launch_bounds__(P::THREADS, 1) pool_collapse_kernel(const mcnn_pool_args a) {
    pdl_enter();
    typedef typename P::idx_t idx_t;
    typedef typename P::uidx_t uidx_t;
    typedef typename P::pair_t pair_t;
    constexpr int POOL_THREADS = P::THREADS;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    unsigned char* smem = P::WIDE ? reinterpret_cast<unsigned char*>(a.workspace) + (size_t)blockIdx.x * a.workspace_stride : dyn_smem;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int E_in = a.E_in, T = a.T, V = a.V;
    const int n = a.ec_in[b];                      // rows of the incoming level
    const bool with_vs = a.vs != nullptr;
    const PoolLayout L = pool_layout<P>(E_in, V, a.ve_cap, with_vs);

    double* vs_s = with_vs ? reinterpret_cast<double*>(smem + L.vs) : nullptr;
    idx_t* gemm = reinterpret_cast<idx_t*>(smem + L.gemm);
    pair_t* evp = reinterpret_cast<pair_t*>(smem + L.evp);
    int8_t* sides = reinterpret_cast<int8_t*>(smem + L.sides);
    uint8_t* mask = smem + L.mask;
    uidx_t* order = reinterpret_cast<uidx_t*>(smem + L.order);
    uidx_t* rep = reinterpret_cast<uidx_t*>(smem + L.rep);
    uint32_t* own_e = reinterpret_cast<uint32_t*>(smem + L.own_e);
    uint32_t* own_v = reinterpret_cast<uint32_t*>(smem + L.own_v);
    uint16_t* wstamp = reinterpret_cast<uint16_t*>(smem + L.wstamp);
    uidx_t* rset_all = reinterpret_cast<uidx_t*>(smem + L.rset);
    uidx_t* slab_start = reinterpret_cast<uidx_t*>(smem + L.slab_start);
    uidx_t* slab_len = reinterpret_cast<uidx_t*>(smem + L.slab_len);
    uidx_t* slab_next = reinterpret_cast<uidx_t*>(smem + L.slab_next);
    uidx_t* chain_tail = reinterpret_cast<uidx_t*>(smem + L.chain_tail);
    uint8_t* vmask_s = smem + L.vmask;
    uidx_t* ve = reinterpret_cast<uidx_t*>(smem + L.ve);
    uidx_t* eposa = reinterpret_cast<uidx_t*>(smem + L.eposa);
    uidx_t* eposb = reinterpret_cast<uidx_t*>(smem + L.eposb);
    pair_t* ulog = reinterpret_cast<pair_t*>(smem + L.ulog);
    int* scan = reinterpret_cast<int*>(smem + L.scan);

    const int32_t* g_gemm = a.gemm_in + (size_t)b * E_in * 4;
    const int8_t* g_sides = a.sides_in + (size_t)b * E_in * 4;
    int32_t* g_edges = a.edges + (size_t)b * a.edges_stride * 2;
    int32_t* g_ve_ptr = a.ve_ptr + (size_t)b * (V + 1);
    int32_t* g_ve_idx = a.ve_idx + (size_t)b * a.ve_cap;
    uint8_t* g_vmask = a.v_mask + (size_t)b * V;
    double* g_vs = with_vs ? a.vs + (size_t)b * V * 3 : nullptr;
    const float* g_norms = a.norms + (size_t)b * E_in;

    long long* clk = a.dbg_clocks ? a.dbg_clocks + (size_t)b * 24 : nullptr;
    if (clk && tid == 0) clk[0] = clock64();
    // ---- phase A: load -----------------------------------------------------------------------------------
    for (int i = tid; i < 4 * n; i += POOL_THREADS) {
        gemm[i] = (idx_t)g_gemm[i];
        sides[i] = g_sides[i];
    }
    for (int i = tid; i < n; i += POOL_THREADS) {
        evp[i] = pr_pack<P>(g_edges[2 * i], g_edges[2 * i + 1]);
        mask[i] = 1;
        own_e[i] = 0u;
        eposa[i] = (uidx_t)P::NIL;
        eposb[i] = (uidx_t)P::NIL;
    }
    for (int v = tid; v < V; v += POOL_THREADS) {
        rep[v] = (uidx_t)v;
        own_v[v] = 0u;
        vmask_s[v] = g_vmask[v];
        const int beg = g_ve_ptr[v], end = g_ve_ptr[v + 1];
        slab_start[v] = (uidx_t)beg;
        slab_len[v] = (uidx_t)(end - beg);
        slab_next[v] = (uidx_t)P::NIL;
        chain_tail[v] = (uidx_t)v;
    }
    {
        const int nve = min(g_ve_ptr[V], a.ve_cap);
        for (int i = tid; i < nve; i += POOL_THREADS) ve[i] = (uidx_t)g_ve_idx[i];
    }
    if (with_vs)
        for (int i = tid; i < 3 * V; i += POOL_THREADS) vs_s[i] = g_vs[i];
    for (int i = tid; i < V * (POOL_THREADS / 32); i += POOL_THREADS) wstamp[i] = 0;
    if (clk && tid == 0) clk[1] = clock64();
    // ---- phase B: sort (norm bits, id) ascending == heapify + heappop order (mesh_pool.py:184-192, F2) ----
    // Squared norms are non-negative floats: their bit patterns order like unsigned integers.  Stable LSD radix sort, four
    // 8-bit passes, starting from id order, so that equal norms stay in ascending id order (the heap's tie-break).  Every
    // warp owns a contiguous chunk of the current sequence; `match.any` ranks equal digits inside a group of 32.
    {
        constexpr int NW = POOL_THREADS / 32;
        __shared__ int s_hist[NW][256];
        uint32_t* kA = reinterpret_cast<uint32_t*>(smem + L.keys);
        uint32_t* kB = kA + n;
        uidx_t* iA = reinterpret_cast<uidx_t*>(kB + n);
        uidx_t* iB = iA + n;
        for (int i = tid; i < n; i += POOL_THREADS) { kA[i] = __float_as_uint(g_norms[i]); iA[i] = (uidx_t)i; }
        const int wid = tid >> 5, lane = tid & 31;
        const int chunk = ((n + NW - 1) / NW + 31) & ~31;
        const int cbeg = min(n, wid * chunk), cend = min(n, cbeg + chunk);
        const unsigned lt = (1u << lane) - 1u;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = pass * 8;
            for (int i = tid; i < NW * 256; i += POOL_THREADS) (&s_hist[0][0])[i] = 0;
            __syncthreads();
            for (int base = cbeg; base < cend; base += 32) {             // count
                const int i = base + lane;
                const bool ok = i < cend;
                const unsigned d = ok ? ((kA[i] >> shift) & 255u) : 256u;
                const unsigned grp = __match_any_sync(0xffffffffu, d);
                if (ok && (grp & lt) == 0u) s_hist[wid][d] += __popc(grp);
            }
            __syncthreads();
            // exclusive scan in (digit, warp) order: one thread per digit walks the NW counters, block scan over digit totals
            {
                int tot = 0;
                if (tid < 256)
                    for (int w = 0; w < NW; ++w) { const int c = s_hist[w][tid]; s_hist[w][tid] = tot; tot += c; }
                int all;
                const int pre = block_scan_excl<POOL_THREADS>(tid < 256 ? tot : 0, scan, all);
                if (tid < 256)
                    for (int w = 0; w < NW; ++w) s_hist[w][tid] += pre;
            }
            __syncthreads();
            for (int base = cbeg; base < cend; base += 32) {             // scatter, in order
                const int i = base + lane;
                const bool ok = i < cend;
                const uint32_t key = ok ? kA[i] : 0u;
                const unsigned d = ok ? ((key >> shift) & 255u) : 256u;
                const unsigned grp = __match_any_sync(0xffffffffu, d);
                int pos = 0;
                if (ok) pos = s_hist[wid][d] + __popc(grp & lt);
                __syncwarp();
                if (ok && (grp & lt) == 0u) s_hist[wid][d] += __popc(grp);
                __syncwarp();
                if (ok) { kB[pos] = key; iB[pos] = iA[i]; }
            }
            __syncthreads();
            uint32_t* tk = kA; kA = kB; kB = tk;
            uidx_t* ti = iA; iA = iB; iB = ti;
        }
        for (int i = tid; i < n; i += POOL_THREADS) order[i] = iA[i];
    }
    // where the two list entries of every edge sit (the entry in the slab of its first / second endpoint)
    for (int v = tid; v < V; v += POOL_THREADS) {
        const int start = slab_start[v], len = slab_len[v];
        for (int i = start; i < start + len; ++i) {
            const int x = ve[i];
            if (x >= n) continue;
            const pair_t w = evp[x];
            if (pr_lo<P>(w) == v) eposa[x] = (uidx_t)i;
            else if (pr_hi<P>(w) == v) eposb[x] = (uidx_t)i;
        }
    }
    __shared__ int s_status, s_nunions, s_cursor, s_flag, s_ndead;
    if (tid == 0) { s_cursor = 0; s_flag = 0; s_ndead = 0; }
    __syncthreads();
    if (clk && tid == 0) clk[2] = clock64();
    // ---- phase D: the collapse (mesh_pool.py:41-53) in rounds; see the header comment ------------------------------------
    __shared__ RoundCtl ctl;
    constexpr int NWARPS = POOL_THREADS / 32;
    const int wid = tid >> 5, lane = tid & 31;
    PoolState<P> S;
    S.gemm = gemm; S.sides = sides; S.mask = mask; S.evp = evp; S.ulog = ulog; S.rep = rep; S.slab_start = slab_start;
    S.slab_len = slab_len; S.slab_next = slab_next; S.chain_tail = chain_tail; S.ve = ve; S.eposa = eposa; S.eposb = eposb;
    S.vmask = vmask_s; S.vs = vs_s; S.own_e = own_e; S.own_v = own_v; S.bump = &ctl.bump; S.bump_end = a.ve_cap + E_in;
    if (tid == 0) { ctl.ec = n; ctl.nunions = 0; ctl.status = 0; ctl.bump = min(g_ve_ptr[V], a.ve_cap); }
    if (tid < 8) (&ctl.bits[0][0])[tid] = 0u;
    __syncthreads();
    {
        uint16_t* my_stamp = wstamp + (size_t)wid * V;
        uidx_t* my_rset = rset_all + (size_t)wid * POOL_RSET_CAP;
        int32_t* lp = a.log_pops ? a.log_pops + (size_t)b * a.log_cap : nullptr;
        int8_t* lo = a.log_outcome ? a.log_outcome + (size_t)b * a.log_cap : nullptr;
        int32_t* lu = a.log_unions ? a.log_unions + (size_t)b * a.log_cap * 8 * 2 : nullptr;
        const int priv_size = (2 * E_in) / NWARPS;               // this warp's own room for re-packed lists (see warp_repack_pair)
        int priv_bump = a.ve_cap + E_in + wid * priv_size;
        const int priv_end = priv_bump + priv_size;
        unsigned ring_tag = 0;                                   // per warp; tags are (k << 7), k = 1 .. 510 (low bits: marks)
        unsigned round = 0;
        int ec = n, cursor = 0, nunions = 0, cstatus = 0;
        auto next_ring_tag = [&]() {
            ring_tag += 128;
            if (ring_tag >= 65408u) {                             // tags wrapped: clear this warp's stamps
                for (int i = lane; i < V; i += 32) my_stamp[i] = 0;
                ring_tag = 128;
                __syncwarp();
            }
            return ring_tag;
        };
        long long t_sel = 0, t_insp = 0, t_chk = 0, t_com = 0, n_alone = 0, n_commit = 0, t_i1 = 0, t_i2 = 0, t_i3 = 0, t_c1 = 0, t_c2 = 0, t_c3 = 0;
        long long s_i1 = 0, s_i2 = 0, s_i3 = 0, s_c1 = 0, s_c2 = 0, s_c3 = 0, n_ring = 0, n_ring_fast = 0, n_plain = 0, n_col = 0;
        while (ec > T && cstatus == 0) {
            ++round;
            const long long c0_ = clk ? clock64() : 0;
            // -- select: the next NWARPS unmasked entries of the sorted order, the w-th one for warp w.  Every warp scans the
            //    same entries and gets the same window, so no hand-over is needed. ----------------------------------------------
            int my_e = -1, my_pos = -1, found = 0;
            for (int p0 = cursor; p0 < n && found < NWARPS; p0 += 32) {
                const int p = p0 + lane;
                const int e = (p < n) ? (int)order[p] : 0;
                const bool alive = (p < n) && mask[e] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                const int cnt = __popc(m);
                const int want = wid - found;                     // which set bit of this chunk is this warp's (uniform)
                if (my_e < 0 && want >= 0 && want < cnt) {
                    const int src = (int)__fns(m, 0, want + 1);
                    my_e = __shfl_sync(0xffffffffu, e, src);
                    my_pos = p0 + src;
                }
                found += cnt;
            }
            if (found == 0) {                                     // the queue ran dry before the target was reached
                for (int p = cursor + tid; p < n && p < a.log_cap; p += POOL_THREADS) {
                    if (lp) lp[p] = order[p];
                    if (lo) lo[p] = MCNN_POP_SKIP_MASKED;
                }
                cursor = n;
                cstatus = MCNN_ST_QUEUE_EMPTY;
                break;
            }
            const long long c1_ = clk ? clock64() : 0;
            // -- inspect (read-only on the mesh state) + reserve the rows a collapse would write ----------------------------------
            const unsigned my_tag = (round << 5) | (unsigned)(31 - wid);
            int verdict = 0;
            int nre = 0;                                          // edge rows read: e (+ its four neighbours)
            int row[4] = {0, 0, 0, 0};
            int we[4] = {0, 0, 0, 0};                             // besides e and row[]: key_b's far-face neighbours, both sides
            int wv[4] = {0, 0, 0, 0};                             // besides v0, v1: the endpoints of both key_b
            int nrset = 0, v0 = 0, v1 = 0;
            int ra0 = (int)P::NIL, rb0 = (int)P::NIL, ra1 = (int)P::NIL, rb1 = (int)P::NIL;   // roots read through the lists (fast walk)
            int ls0 = 0, ll0 = 0, ls1 = 0, ll1 = 0;               // the two list slabs (fast walk), for the re-pack at commit
            bool fast_lists = false, plain = false;              // plain: the nine reserved edge rows are pairwise distinct
            unsigned sd4 = 0, sdkb[2] = {0, 0};                   // sides rows of e and of both key_b
            unsigned ep[6] = {0, 0, 0, 0, 0, 0};                  // list positions (eposa, eposb) of key_b (side 0), key_b (side 2), e
            if (my_e >= 0) {
                typedef Row4<P> RW;
                const int e = my_e;
                const typename RW::T R = RW::load(gemm, e);       // MeshPool.has_boundaries (mesh_pool.py:87-92)
                bool boundary = RW::neg(R);
                nre = 1;
                typename RW::T N[4];
                if (!boundary) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) row[k] = RW::get(R, k);
#pragma unroll
                    for (int k = 0; k < 4; ++k) N[k] = RW::load(gemm, row[k]);
                    boundary = RW::neg(RW::join(RW::join(N[0], N[1]), RW::join(N[2], N[3])));
                    nre = 5;
                }
                if (boundary) {
                    verdict = INSP_REJ_BOUNDARY;
                } else {
                    bool alone = false;
                    sd4 = *reinterpret_cast<const unsigned*>(sides + 4 * e);
#pragma unroll
                    for (int sd = 0; sd < 2; ++sd) {             // __get_invalids' test (mesh_pool.py:115-121) without its side effects
                        const int side = 2 * sd;
                        const int sa = (int)((sd4 >> (8 * side)) & 0xFFu), sb = (int)((sd4 >> (8 * side + 8)) & 0xFFu);
                        const int osa = (sa & 2) ^ 2, osb = (sb & 2) ^ 2;       // == (s - s % 2 + 2) % 4
                        const int oka0 = RW::get(N[side], osa), oka1 = RW::get(N[side], osa + 1);
                        const int okb0 = RW::get(N[side + 1], osb), okb1 = RW::get(N[side + 1], osb + 1);
                        alone |= (oka0 == okb0) | (oka0 == okb1) | (oka1 == okb0) | (oka1 == okb1);
                        we[2 * sd] = okb0; we[2 * sd + 1] = okb1;
                    }
                    edge_ends<P>(evp, rep, e, v0, v1);
                    alone |= v0 == v1;
                    if (clk) s_i1 = clock64();
                    int ring = 0;
                    if (!alone) {
                        ring = warp_one_ring_fast<P>(S, my_stamp, next_ring_tag(), lane, v0, v1, ra0, rb0, ra1, rb1, ls0, ll0, ls1, ll1);
                        fast_lists = ring >= 0;
                        if (!fast_lists) {
                            ra0 = rb0 = ra1 = rb1 = (int)P::NIL;
                            ring = warp_one_ring<P>(S, my_stamp, next_ring_tag(), lane, v0, v1, my_rset, nrset);
                            alone = ring < 0;                     // lists too long for the read-set buffer
                        }
                    }
                    if (clk) s_i2 = clock64();
                    if (alone) {
                        verdict = INSP_ALONE;
                    } else if (ring != 2) {
                        verdict = INSP_REJ_ONE_RING;
                    } else {
                        verdict = INSP_COLLAPSE;
                        edge_ends<P>(evp, rep, row[1], wv[0], wv[1]);     // remove_edge(key_b) edits the lists of key_b's endpoints
                        edge_ends<P>(evp, rep, row[3], wv[2], wv[3]);
                        // reserve: lanes 0..8 the edge rows {e, row[0..3], we[0..3]}, lanes 9..14 the vertex rows {v0, v1, wv[0..3]}
                        int x = e;
#pragma unroll
                        for (int k = 0; k < 4; ++k) { if (lane == 1 + k) x = row[k]; if (lane == 5 + k) x = we[k]; if (lane == 11 + k) x = wv[k]; }
                        if (lane == 9) x = v0;
                        if (lane == 10) x = v1;
                        unsigned prev = 0u;
                        if (lane < 9) prev = atomicMax(&own_e[x], my_tag);
                        else if (lane < 15) atomicMax(&own_v[x], my_tag);
                        // two of this candidate's own edge rows coincide (tiny meshes, valence-3 endpoints): general commit path
                        plain = !__any_sync(0xffffffffu, lane < 9 && prev == my_tag);
                        sdkb[0] = *reinterpret_cast<const unsigned*>(sides + 4 * row[1]);
                        sdkb[1] = *reinterpret_cast<const unsigned*>(sides + 4 * row[3]);
                        ep[0] = eposa[row[1]]; ep[1] = eposb[row[1]]; ep[2] = eposa[row[3]]; ep[3] = eposb[row[3]];
                        ep[4] = eposa[e]; ep[5] = eposb[e];
                    }
                }
            }
            if (clk) s_i3 = clock64();
            const int par = (int)(round & 1u);
            if (lane == 0) {
                ctl.pos[wid] = my_pos;
                if (verdict != 0) atomicOr(&ctl.bits[par][0], 1u << wid);
                if (verdict == INSP_COLLAPSE) atomicOr(&ctl.bits[par][1], 1u << wid);
                if (verdict == INSP_ALONE) atomicOr(&ctl.bits[par][2], 1u << wid);
            }
            __syncthreads();
            const long long c2_ = clk ? clock64() : 0;
            // -- check: does anything this candidate read, or will write, belong to an earlier member of the window? --------------
            {
                bool clash = false;
                auto earlier = [&](unsigned t) { return (t >> 5) == round && (31u - (t & 31u)) < (unsigned)wid; };
                if (verdict != 0 && verdict != INSP_ALONE) {
                    // lanes 0..8: edge rows (0..nre-1 read; 5..8 only written); lanes 9..14: vertex rows; every lane: its list roots
                    int x = my_e;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { if (lane == 1 + k) x = row[k]; if (lane == 5 + k) x = we[k]; if (lane == 11 + k) x = wv[k]; }
                    if (lane == 9) x = v0;
                    if (lane == 10) x = v1;
                    const bool col = verdict == INSP_COLLAPSE, ring_read = verdict != INSP_REJ_BOUNDARY;
                    if (lane < 9) clash = (lane < nre || col) && earlier(own_e[x]);
                    else if (lane < 15) clash = (col || (ring_read && lane < 11)) && earlier(own_v[x]);
                    if (ring_read) {
                        if (fast_lists) {
                            const unsigned nil = P::NIL;
                            const unsigned t0 = own_v[(unsigned)ra0 != nil ? ra0 : 0], t1 = own_v[(unsigned)rb0 != nil ? rb0 : 0];
                            const unsigned t2 = own_v[(unsigned)ra1 != nil ? ra1 : 0], t3 = own_v[(unsigned)rb1 != nil ? rb1 : 0];
                            clash |= ((unsigned)ra0 != nil && earlier(t0)) | ((unsigned)rb0 != nil && earlier(t1)) |
                                     ((unsigned)ra1 != nil && earlier(t2)) | ((unsigned)rb1 != nil && earlier(t3));
                        } else {
                            for (int i = lane; i < nrset; i += 32) {
                                const unsigned xx = my_rset[i];
                                const unsigned t = own_v[xx != P::NIL ? xx : 0u];
                                clash |= (xx != P::NIL) && earlier(t);
                            }
                        }
                    }
                }
                const bool any = __any_sync(0xffffffffu, clash);
                if (lane == 0 && !any) atomicOr(&ctl.bits[par][3], 1u << wid);
            }
            __syncthreads();
            const long long c3_ = clk ? clock64() : 0;
            // -- the prefix that commits (every thread derives it from the shared verdict masks) ----------------------------------
            const unsigned b_any = ctl.bits[par][0], b_col = ctl.bits[par][1], b_alone = ctl.bits[par][2], b_ok = ctl.bits[par][3];
            if (tid < 4) ctl.bits[par ^ 1][tid] = 0u;             // the other parity's words are free since the previous barrier
            int k, ncol;
            bool alone_round = false;
            if (b_alone & 1u) {                                   // the head of the window has to rewire: it runs alone
                k = 1; ncol = 0; alone_round = true;
            } else {
                const unsigned valid = b_any & b_ok & ~b_alone;   // members that may commit if every earlier one does
                k = __ffs(~valid) - 1;                            // >= 1: the head has nothing before it to clash with
                // mesh_pool.py:49: a member is popped only while edges_count > target; every earlier COLLAPSE removed three edges
                const int budget = (ec - T - 1) / 3;              // most collapses that may precede a member
                while (k > 1 && __popc(b_col & ((1u << (k - 1)) - 1u)) > budget) --k;
                ncol = __popc(b_col & ((1u << k) - 1u));
            }
            const int my_uoff = nunions + 4 * __popc(b_col & ((1u << wid) - 1u));
            const int ec_run = ec - 3 * ncol;
            const int new_cursor = ctl.pos[k - 1] + 1;            // the pops move just behind the last committed member
            if (clk) s_c1 = clock64();
            // -- commit ------------------------------------------------------------------------------------------------------------
            if (wid < k) {
                int outcome = -1, merged = -1;
                if (alone_round) {                               // (warp 0) the reference's sequential code path, by one thread
                    int a0, a1, ring = 0, dummy = 0;
                    edge_ends<P>(evp, rep, my_e, a0, a1);
                    if (a0 != a1) ring = warp_one_ring<P>(S, my_stamp, next_ring_tag(), lane, a0, a1, (uidx_t*)nullptr, dummy);
                    if (lane == 0) {
                        Collapse<P> c;
                        c.st = S; c.gemm = gemm; c.sides = sides; c.mask = mask; c.log_unions = lu; c.union_cap = a.log_cap * 8;
                        c.nunions = nunions; c.ulog_cap = L.ulog_cap; c.ec = ec; c.T = T; c.status = 0;
                        outcome = c.pool_edge(my_e, ring, merged);
                        ctl.ec = c.ec; ctl.nunions = c.nunions;
                        if (c.status) ctl.status = c.status;
                    }
                } else if (verdict == INSP_COLLAPSE) {
                    if (lane == 0) {
                        if (plain && my_uoff + 4 <= L.ulog_cap) {
                            // __pool_side x 2 + merge_vertices (mesh_pool.py:67-71, :102-113; mesh.py:26-37) for nine distinct rows,
                            // on the values the inspection read -- nothing has written these rows since -- so this is stores only
                            const int e = my_e;
                            bool slow_remove = false;
#pragma unroll
                            for (int sd = 0; sd < 2; ++sd) {
                                const int side = 2 * sd;
                                const int ka = row[side], kb = row[side + 1], okb0 = we[side], okb1 = we[side + 1];
                                const int sa = (int)((sd4 >> (8 * side)) & 0xFFu), sb = (int)((sd4 >> (8 * side + 8)) & 0xFFu);
                                const int osb = (sb & 2) ^ 2, base = sa & 2;
                                const int s0 = (int)((sdkb[sd] >> (8 * osb)) & 0xFFu), s1 = (int)((sdkb[sd] >> (8 * osb + 8)) & 0xFFu);
                                gemm[4 * ka + base] = (idx_t)okb0;     gemm[4 * okb0 + s0] = (idx_t)ka;
                                sides[4 * ka + base] = (int8_t)s0;     sides[4 * okb0 + s0] = (int8_t)base;
                                gemm[4 * ka + base + 1] = (idx_t)okb1; gemm[4 * okb1 + s1] = (idx_t)ka;
                                sides[4 * ka + base + 1] = (int8_t)s1; sides[4 * okb1 + s1] = (int8_t)(base + 1);
                                ulog[my_uoff + side] = pr_pack<P>(kb, ka);          // union(key_b -> key_a), union(e -> key_a)
                                ulog[my_uoff + side + 1] = pr_pack<P>(e, ka);
                                if (lu != nullptr && my_uoff + side + 1 < a.log_cap * 8) {
                                    lu[2 * (my_uoff + side)] = kb; lu[2 * (my_uoff + side) + 1] = ka;
                                    lu[2 * (my_uoff + side + 1)] = e; lu[2 * (my_uoff + side + 1) + 1] = ka;
                                }
                                mask[kb] = 0;
                                const unsigned ia = ep[2 * sd], ib = ep[2 * sd + 1];      // Mesh.remove_edge(key_b), mesh.py:42-48
                                if (kb != 0 && ia != P::NIL && ib != P::NIL && (int)ve[ia] == kb && (int)ve[ib] == kb) {
                                    ve[ia] = (uidx_t)P::NIL; ve[ib] = (uidx_t)P::NIL;
                                } else {
                                    slow_remove = true;
                                    int ea, eb;
                                    edge_ends<P>(evp, rep, kb, ea, eb);
                                    if (!list_remove_scan<P>(S, ea, kb) || !list_remove_scan<P>(S, eb, kb)) atomicMax(&ctl.status, MCNN_ST_VE_REMOVE);
                                }
                            }
                            {                                                 // Mesh.merge_vertices(e), mesh.py:26-37
                                const unsigned ia = ep[4], ib = ep[5];
                                if (e != 0 && ia != P::NIL && ib != P::NIL && (int)ve[ia] == e && (int)ve[ib] == e) {
                                    ve[ia] = (uidx_t)P::NIL; ve[ib] = (uidx_t)P::NIL;
                                } else {
                                    slow_remove = true;
                                    if (!list_remove_scan<P>(S, v0, e) || !list_remove_scan<P>(S, v1, e)) atomicMax(&ctl.status, MCNN_ST_VE_REMOVE);
                                }
                                if (vs_s != nullptr) {
#pragma unroll
                                    for (int q = 0; q < 3; ++q) vs_s[3 * v0 + q] = (vs_s[3 * v0 + q] + vs_s[3 * v1 + q]) / 2;
                                }
                                vmask_s[v1] = 0;
                                slab_next[(unsigned)chain_tail[v0]] = (uidx_t)v1;
                                chain_tail[v0] = chain_tail[v1];
                                rep[v1] = (uidx_t)v0;
                                mask[e] = 0;
                            }
                            (void)slow_remove;
                            merged = v0;
                            outcome = MCNN_POP_COLLAPSED;
                        } else {
                            Collapse<P> c;
                            c.st = S; c.gemm = gemm; c.sides = sides; c.mask = mask; c.log_unions = lu; c.union_cap = a.log_cap * 8;
                            c.nunions = my_uoff; c.ulog_cap = L.ulog_cap; c.ec = ec; c.T = T; c.status = 0;
                            outcome = c.collapse(my_e, merged);
                            if (c.status) atomicMax(&ctl.status, c.status);
                        }
                    }
                } else {
                    outcome = (verdict == INSP_REJ_BOUNDARY) ? MCNN_POP_REJ_BOUNDARY : MCNN_POP_REJ_ONE_RING;
                }
                merged = __shfl_sync(0xffffffffu, merged, 0);
                if (clk) s_c2 = clock64();
                if (merged >= 0) {                               // keep the merged list one slab (see warp_one_ring_fast)
                    if (!alone_round && fast_lists) warp_repack_pair<P>(S, lane, merged, v1, ls0, ll0, ls1, ll1, priv_bump, priv_end);
                    else warp_repack<P>(S, lane, merged);
                }
                if (lane == 0) {
                    if (lp && my_pos < a.log_cap) lp[my_pos] = my_e;
                    if (lo && my_pos < a.log_cap) lo[my_pos] = (int8_t)outcome;
                }
            }
            if (lp || lo) {
                // the entries between the members were masked when the round began (and stay masked): skipped pops
                for (int p = cursor + tid; p < new_cursor && p < a.log_cap; p += POOL_THREADS) {
                    bool member = false;
                    for (int j = 0; j < k; ++j) member |= ctl.pos[j] == p;
                    if (!member) { if (lp) lp[p] = order[p]; if (lo) lo[p] = MCNN_POP_SKIP_MASKED; }
                }
            }
            if (!alone_round && tid == 0) { ctl.ec = ec_run; ctl.nunions = nunions + 4 * ncol; }
            if (clk) s_c3 = clock64();
            __syncthreads();
            if (clk) { const long long c4_ = clock64(); t_sel += c1_ - c0_; t_insp += c2_ - c1_; t_chk += c3_ - c2_; t_com += c4_ - c3_; n_alone += alone_round; n_commit += k;
                       if (verdict == INSP_COLLAPSE || verdict == INSP_REJ_ONE_RING) { t_i1 += s_i1 - c1_; t_i2 += s_i2 - s_i1; } t_i3 += s_i3 - c1_;
                       t_c1 += s_c1 - c3_; t_c2 += s_c2 - s_c1; t_c3 += s_c3 - s_c2;
                       if (verdict == INSP_COLLAPSE || verdict == INSP_REJ_ONE_RING) { ++n_ring; n_ring_fast += fast_lists; } if (verdict == INSP_COLLAPSE) { ++n_col; n_plain += plain; } }
            cursor = new_cursor;
            ec = ctl.ec;
            nunions = ctl.nunions;
            cstatus = ctl.status;
        }
        if (tid == 0) {
            if (a.log_npops) a.log_npops[b] = cursor;
            if (a.log_nunions) a.log_nunions[b] = nunions;
            s_status = cstatus;
            s_nunions = min(nunions, L.ulog_cap);
            if (clk) { clk[3] = clock64(); clk[7] = cursor; clk[8] = round; clk[9] = t_sel; clk[10] = t_insp; clk[11] = t_chk; clk[12] = t_com; clk[13] = n_alone; clk[14] = n_commit; clk[15] = t_i1; clk[16] = t_i2; clk[17] = t_i3; clk[18] = t_c1; clk[5] = t_c2; clk[6] = t_c3; clk[19 - 19 + 9 + 0] = clk[9]; }
            if (clk) { a.dbg_clocks[(size_t)b * 24 + 0] = clk[0]; }
            if (clk) { clk[13] = n_alone * 1000000LL + n_ring * 1000LL + n_ring_fast; clk[14] = n_commit * 1000000LL + n_col * 1000LL + n_plain; }
        }
    }
    __syncthreads();
    int status = s_status;
    const int U = s_nunions;
    // ---- phase E: Mesh.clean (mesh.py:50-71) + group export ---------------------------------------------------
    // group evaluation scratch: occ / out-adjacency live where the sort keys were
    int* occ = reinterpret_cast<int*>(smem + L.keys);                       // int32[n]
    uidx_t* out0 = reinterpret_cast<uidx_t*>(occ + n);                      // [n]
    uidx_t* out1 = out0 + n;                                                 // [n]
    int* outdeg = reinterpret_cast<int*>(smem + L.cnt32);                   // int32[n]  (later: column lengths)
    int* rowlen = outdeg + n;                                                // int32[n]  (row lengths, then fill cursors)
    for (int i = tid; i < n; i += POOL_THREADS) { occ[i] = 1; outdeg[i] = 0; rowlen[i] = 0; }
    // E1: survivor ranks; dead -> 0
    uidx_t* newid = order;
    int carry = 0, total = 0;
    for (int base = 0; base < n; base += POOL_THREADS) {
        const int i = base + tid;
        const int m = (i < n) ? mask[i] : 0;
        int tot;
        const int pre = block_scan_excl<POOL_THREADS>(m, scan, tot);
        if (i < n) newid[i] = (uidx_t)(m ? carry + pre : 0);
        carry += tot;
    }
    total = carry;
    __syncthreads();
    // out-adjacency of the pour DAG: every source has at most two targets
    for (int k = tid; k < U; k += POOL_THREADS) {
        const pair_t w = ulog[k];
        const int s = pr_lo<P>(w);
        const int slot = atomicAdd(&outdeg[s], 1);
        if (slot == 0) out0[s] = (uidx_t)pr_hi<P>(w);
        else if (slot == 1) out1[s] = (uidx_t)pr_hi<P>(w);
        else s_flag = MCNN_ST_GROUP_OVERFLOW;
    }
    if (clk) { __syncthreads(); if (tid == 0) clk[19] = clock64(); }
    int32_t* o_gemm = a.gemm_out + (size_t)b * T * 4;
    int8_t* o_sides = a.sides_out + (size_t)b * T * 4;
    if (total > T) { status = status ? status : MCNN_ST_QUEUE_EMPTY; total = T; }   // only after an aborted collapse
    // E2: compacted one-ring tables and edge endpoints
    for (int i = tid; i < n; i += POOL_THREADS) {
        if (!mask[i]) continue;
        const int r = newid[i];
        if (r >= T) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = gemm[4 * i + j];
            o_gemm[4 * r + j] = g < 0 ? -1 : (int)newid[g];
            o_sides[4 * r + j] = sides[4 * i + j];
        }
        const pair_t w = evp[i];
        int v0 = pr_lo<P>(w), v1 = pr_hi<P>(w);
        while ((int)rep[v0] != v0) v0 = (int)rep[v0];
        while ((int)rep[v1] != v1) v1 = (int)rep[v1];
        g_edges[2 * r] = v0;
        g_edges[2 * r + 1] = v1;
    }
    for (int i = total * 4 + tid; i < T * 4; i += POOL_THREADS) { o_gemm[i] = 0; o_sides[i] = 0; }
    if (a.edge_mask_out)
        for (int i = tid; i < E_in; i += POOL_THREADS) a.edge_mask_out[(size_t)b * E_in + i] = (i < n) ? mask[i] : 0;
    if (with_vs)
        for (int i = tid; i < 3 * V; i += POOL_THREADS) g_vs[i] = vs_s[i];
    for (int v = tid; v < V; v += POOL_THREADS) g_vmask[v] = vmask_s[v];
    if (clk) { __syncthreads(); if (tid == 0) clk[20] = clock64(); }
    // E3: ve chains -> CSR with re-indexed ids (the list of a merged-away vertex is dropped: nothing can reach it)
    carry = 0;
    for (int base = 0; base < V; base += POOL_THREADS) {
        const int v = base + tid;
        int len = 0;
        const bool root = (v < V) && ((int)rep[v] == v);
        if (root)
            for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)slab_next[s]) {
                const int start = slab_start[s], sl = slab_len[s];
                for (int i = 0; i < sl; ++i) len += (ve[start + i] != P::NIL);
            }
        int tot;
        const int pre = block_scan_excl<POOL_THREADS>(len, scan, tot);
        if (v < V) {
            int pos = carry + pre;
            g_ve_ptr[v] = pos;
            if (root)
                for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)slab_next[s]) {
                    const int start = slab_start[s], sl = slab_len[s];
                    for (int i = 0; i < sl; ++i) {
                        const unsigned x = ve[start + i];
                        if (x != P::NIL) g_ve_idx[pos++] = newid[x];
                    }
                }
        }
        carry += tot;
    }
    if (tid == 0) g_ve_ptr[V] = carry;
    __syncthreads();                       // gemm / sides / evp are dead from here on; out-adjacency complete
    if (clk) { __syncthreads(); if (tid == 0) clk[21] = clock64(); }
    // E4: MeshUnion.  (a) occurrences = number of paths starting at each edge:  occ[s] = 1 + sum over the pours s -> t of
    // occ[t].  A target is alive when it is poured into, so the pours form a DAG; sweeping all edges in parallel until
    // nothing changes reaches its unique fixed point after (depth + 1) rounds.
    uidx_t* cpool = reinterpret_cast<uidx_t*>(smem + L.gemm);               // [3n] column members, allocated by cursor
    uidx_t* coff = cpool + 3 * n;                                            // [n]
    uidx_t* rpool = reinterpret_cast<uidx_t*>(smem + L.evp);                // [<= 3n]: rows at their final CSR offsets
    int* collen = reinterpret_cast<int*>(smem + L.ve);                      // int32[n]; the ve slabs were exported in E3
    const int pool_cap = min(a.nnz_cap, 3 * n);
    for (int round = 0; round <= n; ++round) {
        int changed = 0;
        for (int j = tid; j < n; j += POOL_THREADS) {
            const int d = outdeg[j];
            if (d == 0) continue;
            int v = 1 + occ[out0[j]];
            if (d > 1) v += occ[out1[j]];
            if (v != occ[j]) { occ[j] = v; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (clk) { __syncthreads(); if (tid == 0) clk[22] = clock64(); }
    // (b) the CSC column of every old edge: a survivor sits in its own row; a dead edge is walked forward through the DAG
    // to the surviving rows that contain it.  Dead edges are compacted first so that every thread gets at most a few.
    uidx_t* deadlist = reinterpret_cast<uidx_t*>(smem + L.ulog);            // [n]; the log is consumed
    for (int j = tid; j < n; j += POOL_THREADS) {
        if (mask[j]) {
            const int nr = newid[j];
            int rn = 0;
            if (nr < T) rn = 1;
            const int off = atomicAdd(&s_cursor, rn);
            collen[j] = rn;
            coff[j] = (uidx_t)min((unsigned)off, P::NIL);
            if (rn && off + rn <= pool_cap) { cpool[off] = (uidx_t)nr; atomicAdd(&rowlen[nr], 1); }
        } else {
            deadlist[atomicAdd(&s_ndead, 1)] = (uidx_t)j;
        }
    }
    __syncthreads();
    {
        uidx_t q[POOL_BFS_CAP], r[POOL_BFS_CAP];
        const int ndead = s_ndead;
        for (int di = tid; di < ndead; di += POOL_THREADS) {
            const int j = deadlist[di];
            int rn = 0;
            {
                int qn = 1;
                q[0] = (uidx_t)j;
                bool over = false;
                for (int h = 0; h < qn; ++h) {
                    const int x = q[h];
                    const int d = min(outdeg[x], 2);
                    for (int o = 0; o < d; ++o) {
                        const int t = o ? out1[x] : out0[x];
                        if (mask[t]) {
                            const int nr = newid[t];
                            if (nr >= T) continue;
                            int k = 0;
                            while (k < rn && r[k] != nr) ++k;
                            if (k == rn) { if (rn < POOL_BFS_CAP) r[rn++] = (uidx_t)nr; else over = true; }
                        } else {
                            int k = 0;
                            while (k < qn && q[k] != t) ++k;
                            if (k == qn) { if (qn < POOL_BFS_CAP) q[qn++] = (uidx_t)t; else over = true; }
                        }
                    }
                }
                if (over) s_flag = MCNN_ST_GROUP_OVERFLOW;
                for (int p = 1; p < rn; ++p) {                       // ascending pooled id
                    const uidx_t v = r[p];
                    int k = p;
                    while (k > 0 && r[k - 1] > v) { r[k] = r[k - 1]; --k; }
                    r[k] = v;
                }
            }
            const int off = atomicAdd(&s_cursor, rn);
            collen[j] = rn;
            coff[j] = (uidx_t)min((unsigned)off, P::NIL);
            if (off + rn <= pool_cap)
                for (int k = 0; k < rn; ++k) { cpool[off + k] = r[k]; atomicAdd(&rowlen[r[k]], 1); }
        }
    }
    __syncthreads();
    if (status == 0 && s_flag) status = s_flag;
    const int nnz = s_cursor;
    const bool nnz_ok = nnz <= pool_cap;
    if (!nnz_ok && status == 0) status = MCNN_ST_NNZ_OVERFLOW;
    if (clk) { __syncthreads(); if (tid == 0) clk[23] = clock64(); }
    float* o_occ = a.occ + (size_t)b * E_in;
    for (int i = tid; i < E_in; i += POOL_THREADS) o_occ[i] = (i < n) ? (float)occ[i] : 1.f;
    int32_t* o_rowptr = a.grp_rowptr + (size_t)b * (T + 1);
    int32_t* o_cols = a.grp_cols + (size_t)b * a.nnz_cap;
    int32_t* o_colptr = a.grp_colptr + (size_t)b * (E_in + 1);
    int32_t* o_rows = a.grp_rows + (size_t)b * a.nnz_cap;
    // CSC: column pointers + members (already ascending)
    carry = 0;
    for (int base = 0; base < n; base += POOL_THREADS) {
        const int j = base + tid;
        const int cnt = (j < n) ? collen[j] : 0;
        int tot;
        const int pre = block_scan_excl<POOL_THREADS>(cnt, scan, tot);
        if (j < n) {
            o_colptr[j] = carry + pre;
            if (nnz_ok) {
                const int off = coff[j];
                for (int k = 0; k < cnt; ++k) o_rows[carry + pre + k] = cpool[off + k];
            }
            collen[j] = carry + pre;                    // keep the column start for the row fill below
        }
        carry += tot;
    }
    for (int i = n + tid; i <= E_in; i += POOL_THREADS) o_colptr[i] = nnz;
    // CSR: row pointers over the surviving rows, fill through per-row cursors, sort each row, write out
    carry = 0;
    for (int base = 0; base < total; base += POOL_THREADS) {
        const int r = base + tid;
        const int len = (r < total) ? rowlen[r] : 0;
        int tot;
        const int pre = block_scan_excl<POOL_THREADS>(len, scan, tot);
        if (r < total) {
            o_rowptr[r] = carry + pre;
            occ[r] = carry + pre;                       // occ is exported: reuse as row start ...
            rowlen[r] = carry + pre;                    // ... and rowlen as the fill cursor
        }
        carry += tot;
    }
    for (int r = total + tid; r <= T; r += POOL_THREADS) o_rowptr[r] = nnz;
    __syncthreads();
    if (nnz_ok) {
        for (int j = tid; j < n; j += POOL_THREADS) {
            const int off = coff[j];
            const int cnt = ((j + 1 < n) ? collen[j + 1] : nnz) - collen[j];
            for (int k = 0; k < cnt; ++k) {
                const int pos = atomicAdd(&rowlen[cpool[off + k]], 1);
                rpool[pos] = (uidx_t)j;
            }
        }
        __syncthreads();
        // every (column j, row r) pair finds the rank of j inside row r and writes it to its sorted place
        for (int j = tid; j < n; j += POOL_THREADS) {
            const int off = coff[j];
            const int cnt = ((j + 1 < n) ? collen[j + 1] : nnz) - collen[j];
            for (int k = 0; k < cnt; ++k) {
                const int r = cpool[off + k];
                const int beg = occ[r], end = rowlen[r];
                int rank = 0;
                for (int p = beg; p < end; ++p) rank += (rpool[p] < j);
                o_cols[beg + rank] = j;
            }
        }
    }
    if (clk) __syncthreads();
    if (tid == 0) {
        a.ec_out[b] = total;
        a.status[b] = status;
        if (clk) clk[4] = clock64();
    }
}

// ---------------------------------------------------------------------------------------------------------
// feature kernels
// ---------------------------------------------------------------------------------------------------------
// squared feature norm of every edge row (the collapse priority, mesh_pool.py:197).  Eight lanes per row: a row is read as
// contiguous 128-byte pieces; the eight partial sums are combined in a fixed butterfly order (deterministic).
__global__ void pool_norms_kernel(const float* __restrict__ x, float* __restrict__ norms, int M, int C) {
    pdl_enter();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = (int)(t >> 3), l = (int)(t & 7);
    float s = 0.f;
    if (m < M) {
        const float* p = x + (size_t)m * C;
        if ((C & 3) == 0) {
            for (int c = l << 2; c < C; c += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p + c));
                s = __fadd_rn(s, __fmul_rn(v.x, v.x));
                s = __fadd_rn(s, __fmul_rn(v.y, v.y));
                s = __fadd_rn(s, __fmul_rn(v.z, v.z));
                s = __fadd_rn(s, __fmul_rn(v.w, v.w));
            }
        } else {
            for (int c = l; c < C; c += 8) {
                const float v = __ldg(p + c);
                s = __fadd_rn(s, __fmul_rn(v, v));
            }
        }
    }
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));
    s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));
    if (m < M && l == 0) norms[m] = s;
}

// one thread per (row, channel quad); used for the four segment kernels below
template <bool VEC>
__device__ __forceinline__ void acc_row(float4& acc, const float* __restrict__ p, int C, int c, float w) {
    if (VEC) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + c));
        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    } else {
        if (c + 0 < C) acc.x += w * __ldg(p + c + 0);
        if (c + 1 < C) acc.y += w * __ldg(p + c + 1);
        if (c + 2 < C) acc.z += w * __ldg(p + c + 2);
        if (c + 3 < C) acc.w += w * __ldg(p + c + 3);
    }
}

template <bool VEC>
__device__ __forceinline__ void store_row(float* __restrict__ p, int C, int c, float4 v) {
    if (VEC) {
        *reinterpret_cast<float4*>(p + c) = v;
    } else {
        if (c + 0 < C) p[c + 0] = v.x;
        if (c + 1 < C) p[c + 1] = v.y;
        if (c + 2 < C) p[c + 2] = v.z;
        if (c + 3 < C) p[c + 3] = v.w;
    }
}

// Generic "sparse rows" gather:  out[b][r] = scale(r) * sum_{j in seg(r)} w(j) * in[b][idx[j]]
//   MODE 0 (segmean fwd) : seg = CSR row r,    w = 1,            scale = 1/len,   valid r < ec_out
//   MODE 1 (segmean bwd) : seg = CSC column j, w = 1/len(row),   scale = 1,       valid j < ec_in
//   MODE 2 (unpool fwd)  : seg = CSC column i, w = 1,            scale = 1/occ_i, valid i < ec_hi
//   MODE 3 (unpool bwd)  : seg = CSR row r,    w = 1/occ(col),   scale = 1,       valid r < ec_lo
constexpr int SEG_NR = 4;      // rows per thread, their dependent load chains (ptr -> idx -> weight -> data) issued side by side

template <int MODE, bool VEC>
__global__ void segment_kernel(const float* __restrict__ in, float* __restrict__ out, const int32_t* __restrict__ ptr,
                               const int32_t* __restrict__ idx, const int32_t* __restrict__ aux_ptr,
                               const float* __restrict__ occ, const int32_t* __restrict__ valid, int B, int R_out,
                               int R_in, int ptr_stride, int aux_stride, int occ_stride, int C, int nnz_cap) {
    pdl_enter();
    // grid (row blocks, meshes); a CTA covers SEG_NR * (256 / cq) rows, one thread per (SEG_NR rows, channel quad).  A row's
    // segment is 1-3 entries long, so a thread is a chain of four dependent global loads; with one row per thread the kernel
    // was bound by that latency (three waves of CTAs, each a full chain) -- four rows per thread run their chains together.
    const int cq = (C + 3) >> 2;
    const int rpc = blockDim.x / cq;                             // rows per CTA and pass
    const int lr = threadIdx.x / cq, q = threadIdx.x - lr * cq;
    const int b = blockIdx.y;
    if (lr >= rpc) return;
    const int c = q * 4;
    const int nvalid = __ldg(valid + b);
    const int32_t* pp = ptr + (size_t)b * ptr_stride;
    const int32_t* ii = idx + (size_t)b * nnz_cap;
    const float* src = in + (size_t)b * R_in * C;
    auto weight_of = [&](int j) -> float {
        if (MODE == 1) {
            const int32_t* ap = aux_ptr + (size_t)b * aux_stride;
            return 1.f / (float)(__ldg(ap + j + 1) - __ldg(ap + j));
        }
        if (MODE == 3) return 1.f / __ldg(occ + (size_t)b * occ_stride + j);
        return 1.f;
    };
    int r[SEG_NR], beg[SEG_NR], end[SEG_NR];
#pragma unroll
    for (int k = 0; k < SEG_NR; ++k) {
        r[k] = (blockIdx.x * SEG_NR + k) * rpc + lr;
        const bool live = r[k] < R_out && r[k] < nvalid && r[k] < ptr_stride - 1;
        beg[k] = live ? __ldg(pp + r[k]) : 0;
        end[k] = live ? __ldg(pp + r[k] + 1) : 0;
    }
    int j0[SEG_NR], j1[SEG_NR];
#pragma unroll
    for (int k = 0; k < SEG_NR; ++k) {
        j0[k] = (beg[k] < end[k]) ? __ldg(ii + beg[k]) : 0;
        j1[k] = (beg[k] + 1 < end[k]) ? __ldg(ii + beg[k] + 1) : 0;
    }
    float w0[SEG_NR], w1[SEG_NR];
#pragma unroll
    for (int k = 0; k < SEG_NR; ++k) {
        w0[k] = (beg[k] < end[k]) ? weight_of(j0[k]) : 0.f;
        w1[k] = (beg[k] + 1 < end[k]) ? weight_of(j1[k]) : 0.f;
    }
    float4 a0[SEG_NR], a1[SEG_NR];
#pragma unroll
    for (int k = 0; k < SEG_NR; ++k) {
        a0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        a1[k] = a0[k];
        if (beg[k] < end[k]) acc_row<VEC>(a0[k], src + (size_t)j0[k] * C, C, c, w0[k]);
        if (beg[k] + 1 < end[k]) acc_row<VEC>(a1[k], src + (size_t)j1[k] * C, C, c, w1[k]);
    }
#pragma unroll
    for (int k = 0; k < SEG_NR; ++k) {
        if (r[k] >= R_out) continue;
        // the summation order of the one-row-at-a-time form: pairs (0 + a, 0 + b) added in turn, a trailing single entry directly
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int len = end[k] - beg[k];
        if (len >= 2) {
            acc.x += a0[k].x; acc.y += a0[k].y; acc.z += a0[k].z; acc.w += a0[k].w;
            acc.x += a1[k].x; acc.y += a1[k].y; acc.z += a1[k].z; acc.w += a1[k].w;
            int p = beg[k] + 2;
            for (; p + 1 < end[k]; p += 2) {
                const int ja = __ldg(ii + p), jb = __ldg(ii + p + 1);
                const float wa = weight_of(ja), wb = weight_of(jb);
                float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
                acc_row<VEC>(ta, src + (size_t)ja * C, C, c, wa);
                acc_row<VEC>(tb, src + (size_t)jb * C, C, c, wb);
                acc.x += ta.x; acc.y += ta.y; acc.z += ta.z; acc.w += ta.w;
                acc.x += tb.x; acc.y += tb.y; acc.z += tb.z; acc.w += tb.w;
            }
            if (p < end[k]) {
                const int j = __ldg(ii + p);
                acc_row<VEC>(acc, src + (size_t)j * C, C, c, weight_of(j));
            }
        } else if (len == 1) {
            acc = a0[k];                                          // == 0 + w * v, what acc_row on a zero accumulator gives
        }
        if (len > 0) {
            float s = 1.f;
            if (MODE == 0) s = 1.f / (float)len;
            if (MODE == 2) s = 1.f / __ldg(occ + (size_t)b * occ_stride + r[k]);
            acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s;
        }
        store_row<VEC>(out + ((size_t)b * R_out + r[k]) * C, C, c, acc);
    }
}

template <int MODE>
static int launch_segment(const float* in, float* out, const int32_t* ptr, const int32_t* idx, const int32_t* aux_ptr,
                          const float* occ, const int32_t* valid, int B, int R_out, int R_in, int ptr_stride,
                          int aux_stride, int occ_stride, int C, int nnz_cap, cudaStream_t st) {
    const int cq = (C + 3) / 4;
    const int threads = cq <= 256 ? 256 : ((cq + 31) / 32) * 32;
    if (threads > 1024) return MCNN_E_BADARG;
    const int rpc = threads / cq;
    dim3 grid((unsigned)ceil_div(R_out, rpc * SEG_NR), (unsigned)B);
    if ((C & 3) == 0)
        launch_k(segment_kernel<MODE, true>, grid, threads, 0, st, in, out, ptr, idx, aux_ptr, occ, valid, B, R_out, R_in,
                                                             ptr_stride, aux_stride, occ_stride, C, nnz_cap);
    else
        launch_k(segment_kernel<MODE, false>, grid, threads, 0, st, in, out, ptr, idx, aux_ptr, occ, valid, B, R_out, R_in,
                                                              ptr_stride, aux_stride, occ_stride, C, nnz_cap);
    return after_launch();
}

}  // namespace mcnn

namespace mcnn {
namespace seq { int launch(const mcnn_pool_args* a, size_t* smem_needed, cudaStream_t st, bool query_only); }   // meshpool_seq.cu
// Up to this many edges in the incoming level the sequential kernel is used.  On 750-edge meshes the window of the round
// kernel holds ~2 committable candidates, which only pays for the rounds' fixed cost; at 2280 edges (human-seg / COSEG, real
// conv features) the reworked sequential kernel -- two helper warps, DESIGN 4.3 -- is 3 % of the step ahead as well (4.82 vs
// 4.97 ms, tools/pool_mode_ab.py).  Above, the window fills and the rounds win (2.6x at 170 k edges).
constexpr int POOL_SEQ_MAX_EDGES = 3000;
static int g_pool_mode = 0;                // debug: 0 = by size, 1 = always the round kernel, 2 = always the sequential kernel (on chip)
}  // namespace mcnn

using namespace mcnn;

extern "C" int mcnn_debug_set_pool_mode(int mode) { g_pool_mode = (mode == 1 || mode == 2) ? mode : 0; return MCNN_OK; }

extern "C" size_t mcnn_pool_smem_bytes(int E_in, int V, int ve_cap, int with_vs) {
    return pool_layout<Narrow>(E_in, V, ve_cap, with_vs != 0).total + 64;
}

extern "C" size_t mcnn_pool_workspace_bytes(int E_in, int V, int ve_cap, int with_vs) {
    return align_up(pool_layout<Wide>(E_in, V, ve_cap, with_vs != 0).total + 64, 256);
}

extern "C" size_t mcnn_pool_smem_limit(void) { return 227 * 1024 - 9 * 1024; }   /* dynamic part: the kernel also has ~8.6 KB of static shared memory */

extern "C" int mcnn_pool_norms(const float* x, float* norms, int B, int E, int C, void* stream) {
    if (!x || !norms || B <= 0 || E <= 0 || C <= 0) return MCNN_E_BADARG;
    const int M = B * E;
    launch_k(pool_norms_kernel, ceil_div(M, 32), 256, 0, (cudaStream_t)stream, x, norms, M, C);
    return after_launch();
}

extern "C" int mcnn_pool_collapse(const mcnn_pool_args* a, void* stream) {
    if (!a || a->B <= 0 || a->E_in <= 0 || a->T <= 0 || a->V <= 0 || !a->gemm_in || !a->sides_in || !a->ec_in ||
        !a->norms || !a->edges || !a->ve_ptr || !a->ve_idx || !a->v_mask || !a->gemm_out || !a->sides_out ||
        !a->ec_out || !a->grp_rowptr || !a->grp_cols || !a->grp_colptr || !a->grp_rows || !a->occ || !a->status)
        return MCNN_E_BADARG;
    if (a->ve_cap < 2 * a->E_in) return MCNN_E_BADARG;                                        // the column lengths reuse the ve slabs
    const size_t smem = pool_layout<Narrow>(a->E_in, a->V, a->ve_cap, a->vs != nullptr).total;
    const bool narrow_ok = a->E_in <= 10900 && a->V <= 32767 && a->ve_cap <= 65535 && smem <= mcnn_pool_smem_limit();   // 16-bit ids on chip
    if (narrow_ok && !(a->workspace != nullptr && a->force_workspace)) {
        if (g_pool_mode == 2 || (g_pool_mode == 0 && a->E_in <= POOL_SEQ_MAX_EDGES)) return seq::launch(a, nullptr, (cudaStream_t)stream, false);
        static DynSmemGuard configured;
        if (int grc = configured.ensure(pool_collapse_kernel<Narrow>, smem)) return grc;
        launch_k(pool_collapse_kernel<Narrow>, a->B, Narrow::THREADS, smem, (cudaStream_t)stream, *a);
        return after_launch();
    }
    // large meshes: the same algorithm with 32-bit ids on a global-memory workspace (one slice per mesh)
    if (a->workspace == nullptr) return MCNN_E_TOO_LARGE;
    if ((size_t)a->workspace_stride < mcnn_pool_workspace_bytes(a->E_in, a->V, a->ve_cap, a->vs != nullptr)) return MCNN_E_BADARG;
    launch_k(pool_collapse_kernel<Wide>, a->B, Wide::THREADS, 0, (cudaStream_t)stream, *a);
    return after_launch();
}

extern "C" int mcnn_segmean_fwd(const float* x, const int32_t* grp_rowptr, const int32_t* grp_cols, const int32_t* ec_out,
                                float* out, int B, int E_in, int T, int C, int nnz_cap, void* stream) {
    if (!x || !grp_rowptr || !grp_cols || !ec_out || !out || B <= 0 || E_in <= 0 || T <= 0 || C <= 0) return MCNN_E_BADARG;
    return launch_segment<0>(x, out, grp_rowptr, grp_cols, nullptr, nullptr, ec_out, B, T, E_in, T + 1, 0, 0, C, nnz_cap,
                             (cudaStream_t)stream);
}

extern "C" int mcnn_segmean_bwd(const float* dout, const int32_t* grp_rowptr, const int32_t* grp_colptr,
                                const int32_t* grp_rows, const int32_t* ec_in, float* dx, int B, int E_in, int T, int C,
                                int nnz_cap, void* stream) {
    if (!dout || !grp_rowptr || !grp_colptr || !grp_rows || !ec_in || !dx || B <= 0 || E_in <= 0 || T <= 0 || C <= 0)
        return MCNN_E_BADARG;
    return launch_segment<1>(dout, dx, grp_colptr, grp_rows, grp_rowptr, nullptr, ec_in, B, E_in, T, E_in + 1, T + 1, 0, C,
                             nnz_cap, (cudaStream_t)stream);
}

extern "C" int mcnn_unpool_fwd(const float* f, const int32_t* grp_colptr, const int32_t* grp_rows, const float* occ,
                               const int32_t* ec_hi, float* out, int B, int E_lo, int E_hi, int E_grp, int C, int nnz_cap,
                               void* stream) {
    if (!f || !grp_colptr || !grp_rows || !occ || !ec_hi || !out || B <= 0 || E_lo <= 0 || E_hi < E_grp || E_grp <= 0 || C <= 0)
        return MCNN_E_BADARG;
    return launch_segment<2>(f, out, grp_colptr, grp_rows, nullptr, occ, ec_hi, B, E_hi, E_lo, E_grp + 1, 0, E_grp, C,
                             nnz_cap, (cudaStream_t)stream);
}

extern "C" int mcnn_unpool_bwd(const float* dout, const int32_t* grp_rowptr, const int32_t* grp_cols, const float* occ,
                               const int32_t* ec_lo, float* df, int B, int E_lo, int E_hi, int E_grp, int C, int nnz_cap,
                               void* stream) {
    if (!dout || !grp_rowptr || !grp_cols || !occ || !ec_lo || !df || B <= 0 || E_lo <= 0 || E_hi < E_grp || E_grp <= 0 || C <= 0)
        return MCNN_E_BADARG;
    return launch_segment<3>(dout, df, grp_rowptr, grp_cols, nullptr, occ, ec_lo, B, E_lo, E_hi, E_lo + 1, 0, E_grp, C,
                             nnz_cap, (cudaStream_t)stream);
}
