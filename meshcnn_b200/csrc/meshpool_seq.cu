// MeshPool collapse, SEQUENTIAL form: one CTA per mesh, thread 0 runs the reference's pop loop one candidate at a time and two
// helper warps serve the vertex-list work (R: one-ring walks and searches, M: list merges).  Used for meshes of up to 3000
// edges: the round-based kernel in meshpool.cu commits the conflict-free prefix of a window of candidates per round, and on a
// 750-edge mesh that prefix is ~2 candidates (measured: 2.4 / 1.9 / 1.9 / 1.9 per round on the cubes pool chain) while a round
// costs ~3x a sequential pop; at 2280 edges (3.9 / 3.2 / 2.4 per round) this kernel is still ahead on real conv features.  Where
// the window fills (~15 per round at 170 k edges) the round kernel wins (2.6x).  mcnn_pool_collapse picks by the size of the
// incoming level.  Same results bit for bit (tests/test_pool_gpu.py runs every golden chain through both).
//
// Everything below lives in namespace mcnn::seq; the data structures are described in meshpool.cu.
#include "common.cuh"

namespace mcnn {
namespace seq {

// Two instantiations of the same kernel:
//   Narrow  16-bit ids, state in dynamic shared memory, 256 threads            (meshes up to a few thousand edges)
//   Wide    32-bit ids, state in a per-mesh global-memory workspace (L2), 1024 threads  (no size limit: 170 k-edge meshes)
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
    static constexpr int THREADS = 1024;
    static constexpr bool WIDE = true;
};
template <class P> __device__ __forceinline__ typename P::pair_t pr_pack(int a, int b) {
    return (typename P::pair_t)(typename P::uidx_t)a | ((typename P::pair_t)(typename P::uidx_t)b << P::SHIFT);
}
template <class P> __device__ __forceinline__ int pr_lo(typename P::pair_t w) { return (int)(typename P::uidx_t)w; }
template <class P> __device__ __forceinline__ int pr_hi(typename P::pair_t w) { return (int)(typename P::uidx_t)(w >> P::SHIFT); }

// release / acquire between the warps of a CTA that talk through shared memory (__threadfence_block() is the sequentially
// consistent MEMBAR.SC.CTA; the hand-over protocol below only needs acquire-release ordering)
__device__ __forceinline__ void cta_fence() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

constexpr int POOL_BFS_CAP = 96;         // members of one MeshUnion row / column walked per thread (implementation limit)

struct PoolLayout {
    size_t vs, gemm, evp, sides, mask, order, rep, stamp, slab_start, slab_len, slab_next, chain_tail, vmask, ve, eposa, eposb, ulog, cnt32,
        keys, scan, total;
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
    L.stamp = o; o += sizeof(uint32_t) * (size_t)V; o = align_up(o, 16);
    L.slab_start = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.slab_len = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.slab_next = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.chain_tail = o; o += sizeof(uidx_t) * (size_t)V; o = align_up(o, 16);
    L.vmask = o; o += (size_t)V; o = align_up(o, 16);
    L.ve = o; o += sizeof(uidx_t) * ((size_t)ve_cap + 2 * (size_t)n); o = align_up(o, 16);   // + room for re-packed lists
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


// The sequential collapse runs on thread 0.  Walks over vertex -> edge lists (the one-ring test, and the rare search for
// the first occurrence of id 0) are handed to helper warp R (warp 1) through a small shared-memory ring: it walks a list 32
// entries at a time and the collapse thread waits for the answer; list merges go to helper warp M (warp 2) through a second
// ring and are not waited for (see Collapse::wait_merges).
template <class P>
struct WarpCtx {
    typename P::uidx_t *ve, *slab_start, *slab_len, *slab_next, *chain_tail, *rep, *eposa, *eposb;
    typename P::pair_t* evp;
    uint32_t* stamp;
    int bump, bump_end;        // free room behind the lists (entries), used by wop_repack
};

enum { WOP_EXIT = 0, WOP_ONE_RING = 1, WOP_REMOVE_SCAN = 2, WOP_MERGE = 3 };
constexpr int WQ_CAP = 32;

struct WorkQueue {
    // one 64-bit word per operation, written with ONE store and polled by the helper warp itself:
    //   bits 0..3 op | 4..15 sequence number (1-based, mod 4096) | 16..31 a0 | 32..47 a1 | 48..63 a2     (16-bit ids)
    unsigned long long ent[WQ_CAP];
    // the helper's answer: (sequence number << 2) | result -- one word, so the collapse thread needs one poll and no fence pair;
    // every operation publishes (also the ones nobody waits for): the same word is the ring's flow control
    unsigned res;
};

template <class U>
__device__ __forceinline__ int uf_find(U* rep, int v) {
    int p = rep[v];
    while (p != v) {
        const int gp = rep[p];
        rep[v] = (U)gp;                                 // path halving; concurrent lanes only ever write ancestors
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

// A list that has absorbed other lists (`extend`) is a chain of slabs; every later walk pays one dependent round trip per
// slab.  While the collapse thread waits for a one-ring answer anyway, the helper warp re-packs such a chain into ONE slab in
// the free room behind the lists: order preserved, tombstones dropped, the positions of the moved entries (epos) updated.
template <class P>
__device__ __forceinline__ void wop_repack(WarpCtx<P>& w, int lane, int v) {
    if ((unsigned)w.slab_next[v] == P::NIL) return;
    int need = 0;
    for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)w.slab_next[s]) need += (int)w.slab_len[s];
    if (w.bump + need > w.bump_end) return;                                  // no room left: keep the chain
    const unsigned lt = (1u << lane) - 1u;
    const int out = w.bump;
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
    w.bump = out + total;
    __syncwarp();
}

// ve[v0].extend(ve[v1]) (mesh.py:35), posted by the collapse thread WITHOUT waiting: the helper builds the merged list while
// thread 0 goes on to the next candidate (whose one-ring request queues up behind this operation; thread 0 touches no list
// state before it has collected that answer).  Common case -- both lists are single slabs of <= 32 entries: both are read in
// one go and written, tombstones dropped, as ONE slab into the free room, so the one-ring walks never meet a chain.
// Otherwise: link the chains (O(1)) and re-pack if there is room.
template <class P>
__device__ __forceinline__ void wop_merge(WarpCtx<P>& w, int lane, int v0, int v1) {
    typedef typename P::uidx_t uidx_t;
    const unsigned n0 = (unsigned)w.slab_next[v0], n1 = (unsigned)w.slab_next[v1];
    const int s0 = w.slab_start[v0], l0 = w.slab_len[v0], s1 = w.slab_start[v1], l1 = w.slab_len[v1];
    if (n0 == P::NIL && n1 == P::NIL && l0 <= 32 && l1 <= 32 && w.bump + l0 + l1 <= w.bump_end) {
        const unsigned x0 = (lane < l0) ? (unsigned)w.ve[s0 + lane] : P::NIL;
        const unsigned x1 = (lane < l1) ? (unsigned)w.ve[s1 + lane] : P::NIL;
        const unsigned lt = (1u << lane) - 1u;
        const unsigned m0 = __ballot_sync(0xffffffffu, x0 != P::NIL), m1 = __ballot_sync(0xffffffffu, x1 != P::NIL);
        const int out = w.bump, c0 = __popc(m0), c1 = __popc(m1);
        // all dependent loads of both lists side by side (unconditional, with a safe index)
        const unsigned i0 = x0 != P::NIL ? x0 : 0u, i1 = x1 != P::NIL ? x1 : 0u;
        const int pa0 = (int)w.eposa[i0], pb0 = (int)w.eposb[i0], pa1 = (int)w.eposa[i1], pb1 = (int)w.eposb[i1];
        const typename P::pair_t w1 = w.evp[i1];
        if (x0 != P::NIL) {
            const int dst = out + __popc(m0 & lt);
            w.ve[dst] = (uidx_t)x0;
            if (x0 != 0u) {
                if (pa0 == s0 + lane) w.eposa[x0] = (uidx_t)dst;
                else if (pb0 == s0 + lane) w.eposb[x0] = (uidx_t)dst;
            }
        }
        if (x1 != P::NIL) {
            const int dst = out + c0 + __popc(m1 & lt);
            w.ve[dst] = (uidx_t)x1;
            if (x1 != 0u) {
                if (pa1 == s1 + lane) w.eposa[x1] = (uidx_t)dst;
                else if (pb1 == s1 + lane) w.eposb[x1] = (uidx_t)dst;
            }
            // edges[edges == v1] = v0 (mesh.py:33) applied to the stored endpoints of v1's edges right away: later walks then find
            // roots at the first look-up (the union-find stays authoritative; an entry that is stale or already redirected keeps its word)
            const int lo = pr_lo<P>(w1), hi = pr_hi<P>(w1);
            if (lo == v1 || hi == v1) w.evp[x1] = pr_pack<P>(lo == v1 ? v0 : lo, hi == v1 ? v0 : hi);
        }
        if (lane == 0) { w.slab_start[v0] = (uidx_t)out; w.slab_len[v0] = (uidx_t)(c0 + c1); }
        w.bump = out + c0 + c1;
        __syncwarp();
        return;
    }
    if (lane == 0) {
        w.slab_next[(unsigned)w.chain_tail[v0]] = (uidx_t)v1;
        w.chain_tail[v0] = w.chain_tail[v1];
    }
    __syncwarp();
    wop_repack<P>(w, lane, v0);
}

// __is_one_ring_valid (mesh_pool.py:95-100); the lists may hold stale / dead edge ids.  stamp[w] = tagA marks w as an
// endpoint seen from v0; the lane that flips it to tagB counts it once.  The first slab of both lists (the whole list
// unless a vertex was merged within this level) is fetched and resolved together, so the two dependent load chains overlap.
template <class P>
__device__ __forceinline__ int wop_one_ring(WarpCtx<P>& w, int lane, int v0, int v1, uint32_t tagA, long long* hp) {
    const uint32_t tagB = tagA | 1u;
#ifdef MCNN_POOL_PROF
    const long long t0_ = clock64();
#endif
    // both lists' descriptors in one round trip; a chained list is re-packed first (rare)
    if ((unsigned)w.slab_next[v0] != P::NIL) wop_repack<P>(w, lane, v0);
    if ((unsigned)w.slab_next[v1] != P::NIL) wop_repack<P>(w, lane, v1);
    const int start0 = w.slab_start[v0], len0 = w.slab_len[v0];
    const int start1 = w.slab_start[v1], len1 = w.slab_len[v1];
    const unsigned next0 = (unsigned)w.slab_next[v0], next1 = (unsigned)w.slab_next[v1];
    const unsigned x0 = (lane < len0) ? (unsigned)w.ve[start0 + lane] : P::NIL;
    const unsigned x1 = (lane < len1) ? (unsigned)w.ve[start1 + lane] : P::NIL;
    // the two dependent chains (entry -> endpoints -> union-find roots) run side by side: unconditional loads with a safe index
    const typename P::pair_t w0 = w.evp[x0 != P::NIL ? x0 : 0u], w1 = w.evp[x1 != P::NIL ? x1 : 0u];
    int a0 = pr_lo<P>(w0), b0 = pr_hi<P>(w0), a1 = pr_lo<P>(w1), b1 = pr_hi<P>(w1);
    // the four roots together, one round trip per level of the union-find forest for the whole warp (four separate find loops
    // serialise: each is a divergent chain of three dependent loads)
    while (true) {
        const int ra0 = (int)w.rep[a0], rb0 = (int)w.rep[b0], ra1 = (int)w.rep[a1], rb1 = (int)w.rep[b1];
        const bool fixed = (ra0 == a0) & (rb0 == b0) & (ra1 == a1) & (rb1 == b1);
        a0 = ra0; b0 = rb0; a1 = ra1; b1 = rb1;
        if (__all_sync(0xffffffffu, fixed)) break;
    }
    if (x0 != P::NIL) { w.stamp[a0] = tagA; w.stamp[b0] = tagA; }
#ifdef MCNN_POOL_PROF
    const long long t1_ = clock64();
#endif
    // the rest of the first list: entries 32.. of the first slab, then the slabs linked behind it
    {
        unsigned s = (unsigned)v0;
        int start = start0, len = len0, first = 32;
        while (true) {
            for (int i = first + lane; i < len; i += 32) {
                const unsigned x = w.ve[start + i];
                if (x == P::NIL) continue;
                int a, b;
                edge_ends<P>(w.evp, w.rep, (int)x, a, b);
                w.stamp[a] = tagA;
                w.stamp[b] = tagA;
            }
            s = (s == (unsigned)v0) ? next0 : (unsigned)(unsigned)w.slab_next[s];
            if (s == P::NIL) break;
            start = w.slab_start[s]; len = w.slab_len[s]; first = 0;
        }
    }
    __syncwarp();
#ifdef MCNN_POOL_PROF
    const long long t2_ = clock64();
#endif
    // every lane tests its two endpoints; each hitting (lane, endpoint) then writes its own mark over the stamp and the one mark
    // that survives per vertex counts it: distinct vertices are counted once without atomics (tags are multiples of 128, marks
    // have bit 6 set, so a marked vertex fails the compare-and-swap of the remainder walk below: counted once there too)
    bool ha = (x1 != P::NIL) && a1 != v0 && a1 != v1 && w.stamp[a1] == tagA;
    bool hb = (x1 != P::NIL) && b1 != v0 && b1 != v1 && w.stamp[b1] == tagA;
    __syncwarp();
    const uint32_t mark = tagA | 64u | (uint32_t)(2 * lane);
    if (ha) w.stamp[a1] = mark;
    if (hb) w.stamp[b1] = mark | 1u;
    __syncwarp();
    ha = ha && w.stamp[a1] == mark;
    hb = hb && w.stamp[b1] == (mark | 1u);
    int shared = __popc(__ballot_sync(0xffffffffu, ha)) + __popc(__ballot_sync(0xffffffffu, hb));
#ifdef MCNN_POOL_PROF
    const long long t3_ = clock64();
    if (lane == 0 && hp) { hp[0] += t1_ - t0_; hp[1] += t2_ - t1_; hp[2] += t3_ - t2_; }
#endif
    {
        unsigned s = (unsigned)v1;
        int start = start1, len = len1, first = 32;
        while (true) {
            for (int base = first; base < len; base += 32) {
                const int i = base + lane;
                const unsigned x = (i < len) ? (unsigned)w.ve[start + i] : P::NIL;
                ha = false; hb = false;
                if (x != P::NIL) {
                    int a, b;
                    edge_ends<P>(w.evp, w.rep, (int)x, a, b);
                    if (a != v0 && a != v1) ha = atomicCAS(&w.stamp[a], tagA, tagB) == tagA;
                    if (b != v0 && b != v1) hb = atomicCAS(&w.stamp[b], tagA, tagB) == tagA;
                }
                shared += __popc(__ballot_sync(0xffffffffu, ha)) + __popc(__ballot_sync(0xffffffffu, hb));
            }
            s = (s == (unsigned)v1) ? next1 : (unsigned)(unsigned)w.slab_next[s];
            if (s == P::NIL) break;
            start = w.slab_start[s]; len = w.slab_len[s]; first = 0;
        }
    }
    return shared;
}

// list.remove(x) on ve[v] (mesh.py:48) by search: the first occurrence in list order becomes a tombstone
template <class P>
__device__ __forceinline__ bool wop_list_remove(const WarpCtx<P>& w, int lane, int v, int x) {
    for (unsigned s = (unsigned)v; s != P::NIL; s = (unsigned)w.slab_next[s]) {
        const int start = w.slab_start[s], len = w.slab_len[s];
        for (int base = 0; base < len; base += 32) {
            const int i = base + lane;
            const bool hit = (i < len) && ((int)w.ve[start + i] == x);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m != 0u) {
                if (lane == __ffs(m) - 1) w.ve[start + i] = (typename P::uidx_t)P::NIL;
                __syncwarp();
                return true;
            }
        }
    }
    return false;
}

// Service loop of the helper warp: executes the posted operations in order until WOP_EXIT.
template <class P>
__device__ void helper_warp_loop(WarpCtx<P>& w, WorkQueue* q, int lane, long long* hprof) {
    static_assert(!P::WIDE, "the operation ring packs 16-bit ids");
    unsigned seq = 0;                                        // operations finished so far
    while (true) {
        const unsigned want = (seq + 1u) & 0xFFFu;
        volatile unsigned long long* ep = &q->ent[seq % WQ_CAP];
        unsigned long long rq;
        do { rq = *ep; } while ((((unsigned)rq >> 4) & 0xFFFu) != want);      // all lanes poll the same word: no divergence
        cta_fence();
        ++seq;
        const int op = (int)(rq & 15ull), a0 = (int)((rq >> 16) & 0xFFFFull), a1 = (int)((rq >> 32) & 0xFFFFull), a2 = (int)(rq >> 48);
        if (op == WOP_EXIT) break;
#ifdef MCNN_POOL_PROF
        const long long th_ = clock64();
#endif
        int r = 0;
        if (op == WOP_ONE_RING) {                            // tags: multiples of 128, never reused within a launch (< 2^25 operations)
            r = min(wop_one_ring<P>(w, lane, a0, a1, seq << 7, hprof ? hprof + 3 : nullptr), 3);    // 2 = valid (mesh_pool.py:99)
        } else if (op == WOP_MERGE) {                        // nobody waits for this one
            wop_merge<P>(w, lane, a0, a1);
        } else {                                             // WOP_REMOVE_SCAN: x leaves ve[a], then ve[b] (mesh.py:42-48)
            r = wop_list_remove<P>(w, lane, a1, a0) ? 1 : 0;
            if (r) r = wop_list_remove<P>(w, lane, a2, a0) ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) {
            cta_fence();                           // the list writes above before the answer
            *(volatile unsigned*)&q->res = (seq << 2) | (unsigned)r;
        }
#ifdef MCNN_POOL_PROF
        if (lane == 0 && hprof) hprof[op - 1] += clock64() - th_;
#endif
    }
}

// The collapse state as seen by lane 0 of warp 0 (the only thread that runs the reference's control flow).
template <class P>
struct Collapse {
    typedef typename P::idx_t idx_t;
    typedef typename P::uidx_t uidx_t;
    typedef typename P::pair_t pair_t;
    WorkQueue* q;             // list walks (one-ring tests, searches) go to helper warp R ...
    unsigned posted;
    WorkQueue* qm;            // ... list merges to helper warp M, which nobody waits for unless the next operation needs its result
    unsigned posted_m;
    int last_mv0;             // the vertex whose list the most recent merge writes (-1: none yet)
    idx_t* gemm;
    uidx_t *rep, *slab_next, *chain_tail, *ve, *eposa, *eposb;
    pair_t *evp, *ulog;
    int8_t* sides;
    uint8_t* mask;
    double* vs;               // shared copy or nullptr
    uint8_t* vmask;           // shared copy of v_mask
    int32_t* log_unions;      // global or nullptr
    int union_cap;
    int nunions, ulog_cap;
    int ec, T, status;

#ifdef MCNN_POOL_PROF
    long long prof[7] = {0, 0, 0, 0, 0, 0, 0};
#define PROF_T(i, stmt) { const long long t_ = clock64(); stmt; prof[i] += clock64() - t_; }
#else
#define PROF_T(i, stmt) { stmt; }
#endif
    __device__ __forceinline__ void ends(int e, int& v0, int& v1) { edge_ends<P>(evp, rep, e, v0, v1); }

    __device__ __forceinline__ void post_to(WorkQueue* qq, unsigned& cnt, int op, int a0, int a1, int a2) {
        while (cnt - (*(volatile unsigned*)&qq->res >> 2) >= (unsigned)WQ_CAP) {}
#ifdef MCNN_POOL_PROF
        const long long tp_ = clock64();
#endif
        const unsigned slot = cnt % WQ_CAP;
        ++cnt;
        cta_fence();                               // everything this thread wrote before the operation becomes visible
        *(volatile unsigned long long*)&qq->ent[slot] = (unsigned long long)(unsigned)op | ((unsigned long long)(cnt & 0xFFFu) << 4) |
            ((unsigned long long)(unsigned)(a0 & 0xFFFF) << 16) | ((unsigned long long)(unsigned)(a1 & 0xFFFF) << 32) |
            ((unsigned long long)(unsigned)(a2 & 0xFFFF) << 48);
#ifdef MCNN_POOL_PROF
        prof[5] += clock64() - tp_;
#endif
    }
    __device__ __forceinline__ void post(int op, int a0, int a1, int a2) { post_to(q, posted, op, a0, a1, a2); }

    // Two helper warps.  M builds the merged list of a committed collapse while R already walks the one-ring of the next
    // candidate: the walk reads the lists of the candidate's two endpoints only (plus evp / rep words, where the merge's
    // redirection and the union-find agree), the merge writes the list of ITS surviving vertex and the entry positions of edges
    // incident to it -- so the two may overlap unless the candidate touches that vertex (then the request waits for the merge).
    // Everything thread 0 itself does to the lists (tombstones of the next commit, searches) waits for the merge first; by then
    // it has long finished.  At most one merge is in flight.
    __device__ __forceinline__ void wait_merges() {
        while ((*(volatile unsigned*)&qm->res >> 2) != posted_m) {}
        cta_fence();
    }
    __device__ __forceinline__ void post_merge(int v0, int v1) {
        post_to(qm, posted_m, WOP_MERGE, v0, v1, 0);
        last_mv0 = v0;
    }
    __device__ __forceinline__ void post_ring(int v0, int v1) {
        if (v0 == last_mv0 || v1 == last_mv0) wait_merges();
        post(WOP_ONE_RING, v0, v1, 0);
    }

    // MeshUnion.union (mesh_union.py:11-12) is only logged here; the rows are evaluated after the collapse
    __device__ __forceinline__ void group_union(int s, int t) {
        if (log_unions != nullptr && nunions < union_cap) { log_unions[2 * nunions] = s; log_unions[2 * nunions + 1] = t; }
        if (s == t) status = MCNN_ST_DEGENERATE;
        if (nunions < ulog_cap) ulog[nunions] = pr_pack<P>(s, t);
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
        vmask[v] = 0;
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

    __device__ __forceinline__ int wait_result() {
#ifdef MCNN_POOL_PROF
        const long long tw_ = clock64();
#endif
        unsigned r;
        do { r = *(volatile unsigned*)&q->res; } while ((r >> 2) != posted);
        cta_fence();
#ifdef MCNN_POOL_PROF
        prof[6] += clock64() - tw_;
#endif
        return (int)(r & 3u);
    }

    // __is_one_ring_valid (mesh_pool.py:95-100), walked by the helper warp.  The answer depends on the vertex lists, the
    // endpoints and the union-find only -- none of which __clean_side touches -- so it is requested before the side
    // cleaning and collected after it.
    __device__ void one_ring_request(int e) {
        int v0, v1;
        ends(e, v0, v1);
        post_ring(v0, v1);
    }

    // Mesh.remove_edge (mesh.py:42-48): both list entries of x become tombstones
    __device__ void remove_edge(int x) {
        const unsigned ia = eposa[x], ib = eposb[x];
        if (x != 0 && ia != P::NIL && ib != P::NIL && (int)ve[ia] == x && (int)ve[ib] == x) {
            ve[ia] = (uidx_t)P::NIL;
            ve[ib] = (uidx_t)P::NIL;
            return;
        }
        int a, b;
        ends(x, a, b);
        post(WOP_REMOVE_SCAN, x, a, b);
        if (wait_result() != 1) status = MCNN_ST_VE_REMOVE;
    }

    // __pool_side (mesh_pool.py:102-113)
    __device__ void pool_side(int e, int side) { pool_side(e, side, face_info(e, side)); }

    // the same with the face record already fetched (the caller reads it while it waits for the one-ring answer)
    __device__ void pool_side(int e, int side, const Face f) {
        (void)side;
        const int base = f.sa - (f.sa & 1);
        redirect(f.ka, base, f.okb0, sides[4 * f.kb + f.osb]);
        redirect(f.ka, base + 1, f.okb1, sides[4 * f.kb + f.osb + 1]);     // read after the first rewiring, as the reference does
        group_union(f.kb, f.ka);
        group_union(e, f.ka);
        mask[f.kb] = 0;
        remove_edge(f.kb);
        ec -= 1;
    }

    // Mesh.merge_vertices (mesh.py:26-37)
    __device__ void merge_vertices(int e) {
        remove_edge(e);
        int v0, v1;
        ends(e, v0, v1);
        if (v0 == v1) { status = MCNN_ST_DEGENERATE; return; }
        // ve[v0].extend(ve[v1]) : the helper warp moves v1's list behind v0's (ve[v1] is unreachable once every row holding v1
        // reads v0); not waited for -- the next list operation is queued behind it
        post_merge(v0, v1);
        if (vs != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) vs[3 * v0 + k] = (vs[3 * v0 + k] + vs[3 * v1 + k]) / 2;
        }
        vmask[v1] = 0;
        rep[v1] = (uidx_t)v0;         // edges[edges == v1] = v0 over every row
    }

    // __pool_edge (mesh_pool.py:58-72), general form
    __device__ int pool_edge_generic(int e) {
        bool r;
        PROF_T(0, r = has_boundaries(e));
        if (r) return MCNN_POP_REJ_BOUNDARY;
        // the one-ring walk is the long pole of an attempt: it goes to the helper warp before the side cleaning.  An attempt that
        // the side cleaning rejects does not collect the answer (every answer carries its sequence number; the next supersedes it)
        PROF_T(2, one_ring_request(e));
        return pool_edge_tail(e);
    }

    // the same from the side cleaning on (the one-ring request is on its way)
    __device__ int pool_edge_tail(int e) {
        bool r, side2 = false;
        PROF_T(1, r = clean_side(e, 0));
        if (r) PROF_T(1, r = clean_side(e, 2) ? true : (side2 = true, false));
        if (!r) return side2 ? MCNN_POP_REJ_SIDE2 : MCNN_POP_REJ_SIDE0;
        // nothing writes the one-ring tables between here and __pool_side(0): fetch its face record under the wait
        const Face f0 = face_info(e, 0);
        const bool ring_ok = wait_result() == 2;
        if (!ring_ok) return MCNN_POP_REJ_ONE_RING;
        wait_merges();                                           // the commit edits lists and entry positions itself
        return commit_generic(e, f0);
    }

    __device__ int commit_generic(int e, const Face& f0) {
        PROF_T(3, pool_side(e, 0, f0));
        if (status) return MCNN_POP_COLLAPSED;
        PROF_T(3, pool_side(e, 2));
        if (status) return MCNN_POP_COLLAPSED;
        PROF_T(4, merge_vertices(e); mask[e] = 0);
        ec -= 1;
        return MCNN_POP_COLLAPSED;
    }

    // 16-bit ids: a one-ring row is ONE 64-bit word, its side row one 32-bit word
    __device__ __forceinline__ static int rg(unsigned long long row, int j) { return (int)((row >> (16 * j)) & 0xFFFFull); }
    __device__ __forceinline__ static int rs(uint32_t sd, int j) { return (int)((sd >> (8 * j)) & 0xFFu); }

    // face record of side S (0 / 2) from words that are already in registers
    template <int S>
    __device__ __forceinline__ static Face face_from(unsigned long long rowe, uint32_t sde, unsigned long long rowka,
                                                     unsigned long long rowkb) {
        Face f;
        f.ka = rg(rowe, S);
        f.kb = rg(rowe, S + 1);
        f.sa = rs(sde, S);
        f.sb = rs(sde, S + 1);
        f.osa = (f.sa - (f.sa & 1) + 2) & 3;
        f.osb = (f.sb - (f.sb & 1) + 2) & 3;
        f.oka0 = rg(rowka, f.osa);
        f.oka1 = rg(rowka, f.osa + 1);
        f.okb0 = rg(rowkb, f.osb);
        f.okb1 = rg(rowkb, f.osb + 1);
        return f;
    }
    __device__ __forceinline__ static bool shares(const Face& f) {
        return f.oka0 == f.okb0 || f.oka0 == f.okb1 || f.oka1 == f.okb0 || f.oka1 == f.okb1;
    }

    // __pool_edge for 16-bit ids, arranged around what a single thread is bound by -- dependent shared-memory round trips:
    //  * the row of e and, in ONE further round trip, the rows of its four neighbours give has_boundaries, both face records
    //    (mesh_pool.py:160-170) and both shared-edge tests of __get_invalids; the endpoints of e are resolved alongside, so the
    //    one-ring request leaves two round trips after the pop;
    //  * an attempt whose sides are clean (94 % of them) never calls __clean_side; everything the commit will need that the
    //    helper warp cannot change -- the side words of the two key_b rows, the vertex coordinates -- is fetched while the
    //    one-ring answer is awaited, and the nine rows involved are checked to be distinct (then no store of the commit can
    //    change a word another part of it reads, and the reference's read-after-write order is immaterial);
    //  * the commit is then stores only, plus the six list positions of the three dying edges fetched together.
    // Anything else (a shared edge, rows that coincide, a list entry that is not where it was recorded) takes the general code.
    __device__ int pool_edge_narrow(int e) {
#ifdef MCNN_POOL_PROF
        const long long t0_ = clock64();
#endif
        const unsigned long long HI = 0x8000800080008000ULL;
        const unsigned long long rowe = *reinterpret_cast<const unsigned long long*>(gemm + 4 * e);
        const uint32_t sde = *reinterpret_cast<const uint32_t*>(sides + 4 * e);
        const pair_t we = evp[e];
        if ((rowe & HI) != 0ULL) return MCNN_POP_REJ_BOUNDARY;
        const int n0 = rg(rowe, 0), n1 = rg(rowe, 1), n2 = rg(rowe, 2), n3 = rg(rowe, 3);
        const unsigned long long r0 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * n0);
        const unsigned long long r1 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * n1);
        const unsigned long long r2 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * n2);
        const unsigned long long r3 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * n3);
        const uint32_t d1 = *reinterpret_cast<const uint32_t*>(sides + 4 * n1);
        const uint32_t d3 = *reinterpret_cast<const uint32_t*>(sides + 4 * n3);
        const int ea = pr_lo<P>(we), eb = pr_hi<P>(we);
        const int ra = (int)rep[ea], rb = (int)rep[eb];
        if (((r0 | r1 | r2 | r3) & HI) != 0ULL) return MCNN_POP_REJ_BOUNDARY;
        const int v0 = (ra == ea) ? ea : uf_find(rep, ea);
        const int v1 = (rb == eb) ? eb : uf_find(rep, eb);
#ifdef MCNN_POOL_PROF
        const long long t1_ = clock64();
        prof[0] += t1_ - t0_;
#endif
        post_ring(v0, v1);
#ifdef MCNN_POOL_PROF
        const long long t2_ = clock64();
        prof[2] += t2_ - t1_;
#endif
        const Face f0 = face_from<0>(rowe, sde, r0, r1), f2 = face_from<2>(rowe, sde, r2, r3);
        if (shares(f0) || shares(f2)) return pool_edge_tail(e);          // rewiring around a valence-3 vertex: general code
        // ---- under the wait for the one-ring answer ----
        const int ssb00 = rs(d1, f0.osb), ssb01 = rs(d1, f0.osb + 1), ssb20 = rs(d3, f2.osb), ssb21 = rs(d3, f2.osb + 1);
        bool distinct;
        {
            const int q[9] = {e, f0.ka, f0.kb, f0.okb0, f0.okb1, f2.ka, f2.kb, f2.okb0, f2.okb1};
            bool same = false;
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = i + 1; j < 9; ++j) same |= (q[i] == q[j]);
            distinct = !same && v0 != v1;
        }
        double p0[3] = {0.0, 0.0, 0.0}, p1[3] = {0.0, 0.0, 0.0};
        if (vs != nullptr && distinct) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { p0[k] = vs[3 * v0 + k]; p1[k] = vs[3 * v1 + k]; }
        }
#ifdef MCNN_POOL_PROF
        prof[1] += clock64() - t2_;
#endif
        const bool ring_ok = wait_result() == 2;
        if (!ring_ok) return MCNN_POP_REJ_ONE_RING;
        wait_merges();                                           // the commit edits lists and entry positions itself
        if (!distinct) return commit_generic(e, f0);
#ifdef MCNN_POOL_PROF
        const long long t3_ = clock64();
#endif
        // ---- commit: __pool_side(0), __pool_side(2) (mesh_pool.py:102-113), merge_vertices (mesh.py:26-37) ----
        {
            const int b0 = f0.sa - (f0.sa & 1), b2 = f2.sa - (f2.sa & 1);
            redirect(f0.ka, b0, f0.okb0, ssb00);
            redirect(f0.ka, b0 + 1, f0.okb1, ssb01);
            redirect(f2.ka, b2, f2.okb0, ssb20);
            redirect(f2.ka, b2 + 1, f2.okb1, ssb21);
        }
        group_union(f0.kb, f0.ka);
        group_union(e, f0.ka);
        group_union(f2.kb, f2.ka);
        group_union(e, f2.ka);
        mask[f0.kb] = 0; mask[f2.kb] = 0; mask[e] = 0;
        ec -= 3;
        {   // Mesh.remove_edge x 3 (mesh.py:42-48): the six list positions together, then the six entries
            const int x[3] = {f0.kb, f2.kb, e};
            unsigned ia[3], ib[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { ia[k] = eposa[x[k]]; ib[k] = eposb[x[k]]; }
            int ta[3], tb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ta[k] = (ia[k] != P::NIL) ? (int)ve[ia[k]] : -1;
                tb[k] = (ib[k] != P::NIL) ? (int)ve[ib[k]] : -1;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (x[k] != 0 && ta[k] == x[k] && tb[k] == x[k]) {
                    ve[ia[k]] = (uidx_t)P::NIL;
                    ve[ib[k]] = (uidx_t)P::NIL;
                } else {
                    remove_edge(x[k]);                           // id 0 / stale position: search, as the general code does
                }
            }
        }
        if (status) return MCNN_POP_COLLAPSED;
        post_merge(v0, v1);                                      // ve[v0].extend(ve[v1]) on helper warp M, not waited for
        if (vs != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) vs[3 * v0 + k] = (p0[k] + p1[k]) / 2;
        }
        vmask[v1] = 0;
        rep[v1] = (uidx_t)v0;
#ifdef MCNN_POOL_PROF
        prof[3] += clock64() - t3_;
#endif
        return MCNN_POP_COLLAPSED;
    }

    __device__ __forceinline__ int pool_edge(int e) {
        if constexpr (P::WIDE) return pool_edge_generic(e);
        else return pool_edge_narrow(e);
    }
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
    uint32_t* stamp = reinterpret_cast<uint32_t*>(smem + L.stamp);
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
        eposa[i] = (uidx_t)P::NIL;
        eposb[i] = (uidx_t)P::NIL;
    }
    for (int v = tid; v < V; v += POOL_THREADS) {
        rep[v] = (uidx_t)v;
        stamp[v] = 0u;
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
    // ---- phase D: the sequential collapse (mesh_pool.py:41-53) on thread 0; warps 1 and 2 serve the list operations ------
    __shared__ WorkQueue s_queue, s_queue_m;
    if (tid < WQ_CAP) { s_queue.ent[tid] = 0ull; s_queue_m.ent[tid] = 0ull; }
    if (tid == 0) { s_queue.res = 0u; s_queue_m.res = 0u; }
    __syncthreads();
    if (tid == 0) {
        Collapse<P> c;
        c.q = &s_queue; c.posted = 0u; c.qm = &s_queue_m; c.posted_m = 0u; c.last_mv0 = -1;
        c.gemm = gemm; c.rep = rep; c.slab_next = slab_next; c.chain_tail = chain_tail; c.ve = ve; c.eposa = eposa; c.eposb = eposb;
        c.evp = evp; c.ulog = ulog;
        c.sides = sides; c.mask = mask; c.vs = vs_s; c.vmask = vmask_s;
        c.log_unions = a.log_unions ? a.log_unions + (size_t)b * a.log_cap * 8 * 2 : nullptr;
        c.union_cap = a.log_cap * 8; c.nunions = 0; c.ulog_cap = L.ulog_cap;
        c.ec = n; c.T = T; c.status = 0;
        int32_t* lp = a.log_pops ? a.log_pops + (size_t)b * a.log_cap : nullptr;
        int8_t* lo = a.log_outcome ? a.log_outcome + (size_t)b * a.log_cap : nullptr;
        int p = 0;
        while (c.ec > T) {
            if (p == n) { c.status = MCNN_ST_QUEUE_EMPTY; break; }
            const int e = order[p];
            const int outcome = mask[e] ? c.pool_edge(e) : MCNN_POP_SKIP_MASKED;
            if (lp && p < a.log_cap) lp[p] = e;
            if (lo && p < a.log_cap) lo[p] = (int8_t)outcome;
            ++p;
            if (c.status) break;
        }
        c.post(WOP_EXIT, 0, 0, 0);
        c.post_to(c.qm, c.posted_m, WOP_EXIT, 0, 0, 0);
        if (a.log_npops) a.log_npops[b] = p;
        if (a.log_nunions) a.log_nunions[b] = c.nunions;
        s_status = c.status;
        s_nunions = min(c.nunions, c.ulog_cap);
        if (clk) {
            clk[3] = clock64();
            clk[7] = p;
#ifdef MCNN_POOL_PROF
            for (int i = 0; i < 5; ++i) clk[8 + i] = c.prof[i];
            clk[5] = c.prof[5];
            clk[6] = c.prof[6];
#endif
        }
    } else if (tid >= 32 && tid < 96) {
        // warp 1 = R (one-ring walks, searches), warp 2 = M (merges); each re-packs lists into its own part of the spare room
        const bool is_m = tid >= 64;
        WarpCtx<P> w;
        w.ve = ve; w.slab_start = slab_start; w.slab_len = slab_len; w.slab_next = slab_next; w.chain_tail = chain_tail;
        w.rep = rep; w.evp = evp; w.stamp = stamp; w.eposa = eposa; w.eposb = eposb;
        const int room0 = min(g_ve_ptr[V], a.ve_cap), room1 = a.ve_cap + 2 * E_in;
        const int cut = room1 - (room1 - room0) / 4;             // M: [room0, cut), R: [cut, room1)
        w.bump = is_m ? room0 : cut; w.bump_end = is_m ? cut : room1;
        helper_warp_loop<P>(w, is_m ? &s_queue_m : &s_queue, tid & 31, clk ? clk + 13 : nullptr);
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
int launch(const mcnn_pool_args* a, size_t* smem_needed, cudaStream_t st, bool query_only) {
    const size_t smem = pool_layout<Narrow>(a->E_in, a->V, a->ve_cap, a->vs != nullptr).total;
    if (smem_needed) *smem_needed = smem;
    if (query_only) return MCNN_OK;
    static DynSmemGuard configured;
    if (int grc = configured.ensure(pool_collapse_kernel<Narrow>, smem)) return grc;
    launch_k(pool_collapse_kernel<Narrow>, a->B, Narrow::THREADS, smem, st, *a);
    return after_launch();
}

}  // namespace seq
}  // namespace mcnn
