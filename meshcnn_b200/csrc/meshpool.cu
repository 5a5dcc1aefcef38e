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
//   ve lists                 singly linked nodes (first-occurrence removal, O(1) tail concatenation); stale
//                            entries are kept, as in the reference
//   MeshUnion rows           linked (member, multiplicity) nodes -- `groups[t] += groups[s]` keeps
//                            multiplicities because the unpool occurrences need them
//
// Phases: load -> bitonic sort of (norm bits, id) keys -> thread 0 runs the sequential collapse (order and side
// effects identical to the reference, failed attempts included) -> all threads run Mesh.clean (compaction,
// re-indexing with dead -> 0) and emit the groups as CSR + CSC + un-clamped occurrences.
#include "common.cuh"

namespace mcnn {

typedef int16_t idx_t;           // shared-memory variant: n, V, node counts < 32768 (0xFFFF = nil)
typedef uint16_t cnt_t;

constexpr int POOL_THREADS = 256;
constexpr int GROUP_NODE_FACTOR = 3;     // group node pool = 3 n nodes

struct PoolLayout {
    size_t vs, keys, gemm, evp, sides, mask, order, rep, vhead, vtail, stamp, vn, ghead, gtail, glen, gn, ulog, scan, total;
    int npow2, gn_cap, ulog_cap;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline PoolLayout pool_layout(int n, int V, int ve_cap, bool with_vs) {
    PoolLayout L;
    size_t o = 0;
    L.vs = o; o += with_vs ? sizeof(double) * 3 * (size_t)V : 0; o = align_up(o, 16);
    L.gemm = o; o += sizeof(idx_t) * 4 * (size_t)n; o = align_up(o, 16);
    L.evp = o; o += sizeof(uint32_t) * (size_t)n; o = align_up(o, 16);
    L.sides = o; o += 4 * (size_t)n; o = align_up(o, 16);
    L.mask = o; o += (size_t)n; o = align_up(o, 16);
    L.order = o; o += sizeof(idx_t) * (size_t)n; o = align_up(o, 16);
    L.rep = o; o += sizeof(idx_t) * (size_t)V; o = align_up(o, 16);
    L.vhead = o; o += sizeof(idx_t) * (size_t)V; o = align_up(o, 16);
    L.vtail = o; o += sizeof(idx_t) * (size_t)V; o = align_up(o, 16);
    L.stamp = o; o += sizeof(cnt_t) * (size_t)V; o = align_up(o, 16);
    L.vn = o; o += sizeof(uint32_t) * (size_t)ve_cap; o = align_up(o, 16);
    L.ghead = o; o += sizeof(idx_t) * (size_t)n; o = align_up(o, 16);
    L.gtail = o; o += sizeof(idx_t) * (size_t)n; o = align_up(o, 16);
    L.glen = o; o += sizeof(cnt_t) * (size_t)n; o = align_up(o, 16);
    L.ulog_cap = 2 * n + 16;                     // every union's source dies within the same attempt: <= 2 per dead edge
    L.ulog = o; o += sizeof(uint32_t) * (size_t)L.ulog_cap; o = align_up(o, 16);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    L.npow2 = np2;
    L.gn_cap = GROUP_NODE_FACTOR * n;
    // the sort keys (uint64[npow2], npow2 < 2n) are dead before the group nodes (uint64[3n]) are born
    L.keys = o;
    L.gn = o; o += sizeof(unsigned long long) * (size_t)L.gn_cap; o = align_up(o, 16);
    if (o < L.keys + sizeof(unsigned long long) * (size_t)np2) o = align_up(L.keys + sizeof(unsigned long long) * (size_t)np2, 16);
    L.scan = o; o += sizeof(int) * 48; o = align_up(o, 16);
    L.total = o;
    return L;
}

// exclusive block scan of one int per thread; returns the prefix and (to all threads) the block total
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
        int w = (lane < POOL_THREADS / 32) ? scratch[lane] : 0;
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

constexpr unsigned NIL16 = 0xFFFFu;

// group node: member [0,16) | multiplicity [16,32) | next [32,48)
__device__ __forceinline__ unsigned long long gn_pack(unsigned member, unsigned mult, unsigned next) {
    return (unsigned long long)member | ((unsigned long long)mult << 16) | ((unsigned long long)next << 32);
}

// MeshUnion rows, owned by the consumer thread (warp 1) while the collapse runs
struct Groups {
    idx_t *ghead, *gtail;
    cnt_t* glen;
    unsigned long long* gn;
    int gn_count, gn_cap, status;

    // MeshUnion.union (mesh_union.py:11-12): row t += row s, multiplicities kept
    __device__ void row_add(int s, int t) {
        if (s == t) {
            for (unsigned nt = (unsigned short)ghead[t]; nt != NIL16;) {
                const unsigned long long w = gn[nt];
                const unsigned nx = (unsigned)(w >> 32) & 0xFFFFu;
                gn[nt] = gn_pack((unsigned)w & 0xFFFFu, (((unsigned)w >> 16) & 0xFFFFu) * 2u, nx);
                nt = nx;
            }
            return;
        }
        for (unsigned ns = (unsigned short)ghead[s]; ns != NIL16;) {
            const unsigned long long ws = gn[ns];
            const unsigned m = (unsigned)ws & 0xFFFFu, c = ((unsigned)ws >> 16) & 0xFFFFu;
            ns = (unsigned)(ws >> 32) & 0xFFFFu;
            bool found = false;
            for (unsigned nt = (unsigned short)ghead[t]; nt != NIL16;) {
                const unsigned long long wt = gn[nt];
                if (((unsigned)wt & 0xFFFFu) == m) {
                    gn[nt] = wt + ((unsigned long long)c << 16);
                    found = true;
                    break;
                }
                nt = (unsigned)(wt >> 32) & 0xFFFFu;
            }
            if (!found) {
                if (gn_count >= gn_cap) { status = MCNN_ST_GROUP_OVERFLOW; return; }
                const unsigned id = (unsigned)gn_count++;
                gn[id] = gn_pack(m, c, NIL16);
                const unsigned tl = (unsigned short)gtail[t];
                gn[tl] = (gn[tl] & ~(0xFFFFull << 32)) | ((unsigned long long)id << 32);
                gtail[t] = (idx_t)id;
                glen[t] = (cnt_t)(glen[t] + 1);
            }
        }
    }
};

struct Collapse {
    idx_t *gemm, *rep, *vhead, *vtail;
    uint32_t *evp, *vn, *ulog;
    cnt_t* stamp;
    int8_t* sides;
    uint8_t* mask;
    double* vs;               // shared copy or nullptr
    uint8_t* vmask_g;         // global v_mask of this mesh
    int32_t* log_unions;      // global or nullptr
    int union_cap;
    int nunions, ulog_cap;
    int ec, T, status;
    unsigned stamp_id;

    __device__ __forceinline__ int find(int v) {
        int p = rep[v];
        while (p != v) {
            const int gp = rep[p];
            rep[v] = (idx_t)gp;
            v = gp;
            p = rep[v];
        }
        return v;
    }
    __device__ __forceinline__ void ends(int e, int& v0, int& v1) {
        const uint32_t w = evp[e];
        const int a = (int)(w & 0xFFFFu), b = (int)(w >> 16);
        const int ra = rep[a], rb = rep[b];                 // both loads in flight before either branch
        v0 = (ra == a) ? a : find(a);
        v1 = (rb == b) ? b : find(b);
    }

    // the union itself is applied by the consumer warp; the collapse only logs (source, target)
    __device__ __forceinline__ void group_union(int s, int t) {
        if (log_unions != nullptr && nunions < union_cap) { log_unions[2 * nunions] = s; log_unions[2 * nunions + 1] = t; }
        if (nunions < ulog_cap) ulog[nunions] = (uint32_t)s | ((uint32_t)t << 16);
        else status = MCNN_ST_GROUP_OVERFLOW;
        ++nunions;
    }

    __device__ __forceinline__ bool row_has_hole(int e) const {
        const unsigned long long row = *reinterpret_cast<const unsigned long long*>(gemm + 4 * e);
        return (row & 0x8000800080008000ULL) != 0ULL;
    }

    // MeshPool.has_boundaries (mesh_pool.py:87-92)
    __device__ bool has_boundaries(int e) const {
        const unsigned long long row = *reinterpret_cast<const unsigned long long*>(gemm + 4 * e);
        if ((row & 0x8000800080008000ULL) != 0ULL) return true;
        const unsigned long long r0 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)(row & 0xFFFF));
        const unsigned long long r1 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 16) & 0xFFFF));
        const unsigned long long r2 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 32) & 0xFFFF));
        const unsigned long long r3 = *reinterpret_cast<const unsigned long long*>(gemm + 4 * (int)((row >> 48) & 0xFFFF));
        return ((r0 | r1 | r2 | r3) & 0x8000800080008000ULL) != 0ULL;
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
    __device__ int get_invalids(int e, int side, int tri[3]) {
        const Face f = face_info(e, side);
        const int oka[2] = {f.oka0, f.oka1}, okb[2] = {f.okb0, f.okb1};
        int nsh = 0, si = 0, sj = 0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (oka[i] == okb[j]) {
                    if (nsh == 0) { si = i; sj = j; }
                    ++nsh;
                }
        if (nsh == 0) return 0;
        if (nsh != 1) { status = MCNN_ST_SHARED_ASSERT; return 0; }
        const int mid = oka[si];
        const int ua = oka[1 - si], ub = okb[1 - sj];
        const int usa = sides[4 * f.ka + f.osa + 1 - si];
        const int usb = sides[4 * f.kb + f.osb + 1 - sj];
        redirect(e, side, ua, usa);
        redirect(e, side + 1, ub, usb);
        redirect(ua, other_side(usa), ub, other_side(usb));
        group_union(f.ka, e);
        group_union(f.kb, e);
        group_union(f.ka, ua);
        group_union(mid, ua);
        group_union(f.kb, ub);
        group_union(mid, ub);
        tri[0] = f.ka; tri[1] = f.kb; tri[2] = mid;
        return 3;
    }

    // __remove_triplete (mesh_pool.py:172-182): no remove_edge, the ve entries go stale
    __device__ void remove_triplet(const int tri[3]) {
        int a0, a1, b0, b1, c0, c1;
        ends(tri[0], a0, a1);
        ends(tri[1], b0, b1);
        ends(tri[2], c0, c1);
        mask[tri[0]] = 0; mask[tri[1]] = 0; mask[tri[2]] = 0;
        ec -= 3;
        int count = 0, v = -1;
        if ((a0 == b0 || a0 == b1) && (a0 == c0 || a0 == c1)) { ++count; v = a0; }
        if (a1 != a0 && (a1 == b0 || a1 == b1) && (a1 == c0 || a1 == c1)) { ++count; v = a1; }
        if (count != 1) { status = MCNN_ST_VERTEX_ASSERT; return; }
        vmask_g[v] = 0;
    }

    // __clean_side (mesh_pool.py:74-85)
    __device__ bool clean_side(int e, int side) {
        if (ec <= T) return false;
        int tri[3];
        int n = get_invalids(e, side, tri);
        while (n != 0 && ec > T) {
            if (status) return false;
            remove_triplet(tri);
            if (status) return false;
            if (ec <= T) return false;
            if (has_boundaries(e)) return false;
            n = get_invalids(e, side, tri);
        }
        return status == 0;
    }

    // __is_one_ring_valid (mesh_pool.py:95-100); the lists may hold stale / dead edge ids.
    // stamp[w] = (id << 2) | 1 marks w as a member of the first set, bit 1 = already counted in the intersection.
    __device__ bool one_ring_valid(int e) {
        int v0, v1;
        ends(e, v0, v1);
        ++stamp_id;
        const unsigned tagA = (stamp_id << 2) | 1u;
        unsigned nd = (unsigned short)vhead[v0];
        if (nd != NIL16) {
            uint32_t node = vn[nd];
            while (true) {
                const unsigned nx = node >> 16;
                const uint32_t nextnode = (nx != NIL16) ? vn[nx] : 0u;      // next node in flight while this one is processed
                int a, b;
                ends((int)(node & 0xFFFFu), a, b);
                stamp[a] = (cnt_t)tagA;
                stamp[b] = (cnt_t)tagA;
                if (nx == NIL16) break;
                node = nextnode;
            }
        }
        int shared = 0;
        nd = (unsigned short)vhead[v1];
        if (nd != NIL16) {
            uint32_t node = vn[nd];
            while (true) {
                const unsigned nx = node >> 16;
                const uint32_t nextnode = (nx != NIL16) ? vn[nx] : 0u;
                int a, b;
                ends((int)(node & 0xFFFFu), a, b);
                const unsigned ta = stamp[a], tb = stamp[b];
                if (a != v0 && a != v1 && ta == tagA) { stamp[a] = (cnt_t)(tagA | 2u); ++shared; }
                if (b != a && b != v0 && b != v1 && tb == tagA) { stamp[b] = (cnt_t)(tagA | 2u); ++shared; }
                if (nx == NIL16) break;
                node = nextnode;
            }
        }
        return shared == 2;
    }

    // list.remove(x) on ve[v] (mesh.py:48): first occurrence
    __device__ bool list_remove(int v, int x) {
        unsigned prev = NIL16;
        unsigned nd = (unsigned short)vhead[v];
        uint32_t node = (nd != NIL16) ? vn[nd] : 0u;
        while (nd != NIL16) {
            const unsigned nx = node >> 16;
            if ((int)(node & 0xFFFFu) == x) {
                if (prev == NIL16) vhead[v] = (idx_t)nx;
                else vn[prev] = (vn[prev] & 0xFFFFu) | (nx << 16);
                if ((unsigned short)vtail[v] == nd) vtail[v] = (idx_t)prev;
                return true;
            }
            const uint32_t nextnode = (nx != NIL16) ? vn[nx] : 0u;
            prev = nd;
            nd = nx;
            node = nextnode;
        }
        status = MCNN_ST_VE_REMOVE;
        return false;
    }

    // Mesh.remove_edge (mesh.py:42-48)
    __device__ bool remove_edge(int x) {
        int a, b;
        ends(x, a, b);
        if (!list_remove(a, x)) return false;
        return list_remove(b, x);
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

    // Mesh.merge_vertices (mesh.py:26-37)
    __device__ void merge_vertices(int e) {
        if (!remove_edge(e)) return;
        int v0, v1;
        ends(e, v0, v1);
        if (v0 == v1) { status = MCNN_ST_DEGENERATE; return; }
        if (vs != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) vs[3 * v0 + k] = (vs[3 * v0 + k] + vs[3 * v1 + k]) / 2;
        }
        vmask_g[v1] = 0;
        // ve[v0].extend(ve[v1]) : the nodes move (ve[v1] is unreachable once every row holding v1 reads v0)
        const unsigned h1 = (unsigned short)vhead[v1];
        if (h1 != NIL16) {
            if ((unsigned short)vhead[v0] == NIL16) vhead[v0] = (idx_t)h1;
            else { const unsigned t0 = (unsigned short)vtail[v0]; vn[t0] = (vn[t0] & 0xFFFFu) | (h1 << 16); }
            vtail[v0] = vtail[v1];
            vhead[v1] = (idx_t)NIL16; vtail[v1] = (idx_t)NIL16;
        }
        rep[v1] = (idx_t)v0;          // edges[edges == v1] = v0 over every row
    }

    // __pool_edge (mesh_pool.py:58-72)
    __device__ int pool_edge(int e) {
        if (has_boundaries(e)) return MCNN_POP_REJ_BOUNDARY;
        if (!clean_side(e, 0)) return MCNN_POP_REJ_SIDE0;
        if (!clean_side(e, 2)) return MCNN_POP_REJ_SIDE2;
        if (!one_ring_valid(e)) return MCNN_POP_REJ_ONE_RING;
        pool_side(e, 0);
        if (status) return MCNN_POP_COLLAPSED;
        pool_side(e, 2);
        if (status) return MCNN_POP_COLLAPSED;
        merge_vertices(e);
        mask[e] = 0;
        ec -= 1;
        return MCNN_POP_COLLAPSED;
    }
};

__global__ void __launch_bounds__(POOL_THREADS, 1) pool_collapse_kernel(const mcnn_pool_args a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int E_in = a.E_in, T = a.T, V = a.V;
    const int n = a.ec_in[b];                      // rows of the incoming level
    const bool with_vs = a.vs != nullptr;
    const PoolLayout L = pool_layout(E_in, V, a.ve_cap, with_vs);

    double* vs_s = with_vs ? reinterpret_cast<double*>(smem + L.vs) : nullptr;
    idx_t* gemm = reinterpret_cast<idx_t*>(smem + L.gemm);
    uint32_t* evp = reinterpret_cast<uint32_t*>(smem + L.evp);
    int8_t* sides = reinterpret_cast<int8_t*>(smem + L.sides);
    uint8_t* mask = smem + L.mask;
    idx_t* order = reinterpret_cast<idx_t*>(smem + L.order);
    idx_t* rep = reinterpret_cast<idx_t*>(smem + L.rep);
    idx_t* vhead = reinterpret_cast<idx_t*>(smem + L.vhead);
    idx_t* vtail = reinterpret_cast<idx_t*>(smem + L.vtail);
    cnt_t* stamp = reinterpret_cast<cnt_t*>(smem + L.stamp);
    uint32_t* vn = reinterpret_cast<uint32_t*>(smem + L.vn);
    idx_t* ghead = reinterpret_cast<idx_t*>(smem + L.ghead);
    idx_t* gtail = reinterpret_cast<idx_t*>(smem + L.gtail);
    cnt_t* glen = reinterpret_cast<cnt_t*>(smem + L.glen);
    uint32_t* ulog = reinterpret_cast<uint32_t*>(smem + L.ulog);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.keys);
    unsigned long long* gn = reinterpret_cast<unsigned long long*>(smem + L.gn);
    int* scan = reinterpret_cast<int*>(smem + L.scan);

    const int32_t* g_gemm = a.gemm_in + (size_t)b * E_in * 4;
    const int8_t* g_sides = a.sides_in + (size_t)b * E_in * 4;
    int32_t* g_edges = a.edges + (size_t)b * a.edges_stride * 2;
    int32_t* g_ve_ptr = a.ve_ptr + (size_t)b * (V + 1);
    int32_t* g_ve_idx = a.ve_idx + (size_t)b * a.ve_cap;
    uint8_t* g_vmask = a.v_mask + (size_t)b * V;
    double* g_vs = with_vs ? a.vs + (size_t)b * V * 3 : nullptr;
    const float* g_norms = a.norms + (size_t)b * E_in;

    // ---- phase A: load -----------------------------------------------------------------------------------
    for (int i = tid; i < 4 * n; i += POOL_THREADS) {
        gemm[i] = (idx_t)g_gemm[i];
        sides[i] = g_sides[i];
    }
    for (int i = tid; i < n; i += POOL_THREADS) {
        evp[i] = (uint32_t)(g_edges[2 * i] & 0xFFFF) | ((uint32_t)(g_edges[2 * i + 1] & 0xFFFF) << 16);
        mask[i] = 1;
    }
    for (int v = tid; v < V; v += POOL_THREADS) {
        rep[v] = (idx_t)v;
        stamp[v] = 0;
        const int beg = g_ve_ptr[v], end = g_ve_ptr[v + 1];
        vhead[v] = (idx_t)(beg < end ? beg : (int)NIL16);
        vtail[v] = (idx_t)(beg < end ? end - 1 : (int)NIL16);
        for (int i = beg; i < end; ++i)
            vn[i] = ((uint32_t)g_ve_idx[i] & 0xFFFFu) | ((uint32_t)(i + 1 < end ? i + 1 : (int)NIL16) << 16);
    }
    if (with_vs)
        for (int i = tid; i < 3 * V; i += POOL_THREADS) vs_s[i] = g_vs[i];
    // ---- phase B: sort (norm bits, id) ascending == heapify + heappop order (mesh_pool.py:184-192, F2) ----
    const int np2 = L.npow2;
    for (int i = tid; i < np2; i += POOL_THREADS) {
        unsigned long long k = ~0ULL;
        if (i < n) k = ((unsigned long long)__float_as_uint(g_norms[i]) << 32) | (unsigned)i;
        keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < np2; i += POOL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = keys[i], y = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += POOL_THREADS) order[i] = (idx_t)(keys[i] & 0xffffffffULL);
    __syncthreads();
    // ---- phase C: MeshUnion = identity (mesh_union.py:9) -----------------------------------------------------
    for (int i = tid; i < n; i += POOL_THREADS) {
        ghead[i] = (idx_t)i; gtail[i] = (idx_t)i; glen[i] = 1;
        gn[i] = gn_pack((unsigned)i, 1u, NIL16);
    }
    __shared__ int s_status, s_ec, s_gstatus;
    __shared__ volatile int s_pub, s_done;
    if (tid == 0) { s_pub = 0; s_done = 0; s_gstatus = 0; }
    __syncthreads();
    // ---- phase D: the sequential collapse (mesh_pool.py:41-53); thread 32 applies the logged unions concurrently ----
    if (tid == 0) {
        Collapse c;
        c.gemm = gemm; c.rep = rep; c.vhead = vhead; c.vtail = vtail; c.evp = evp; c.vn = vn; c.ulog = ulog;
        c.stamp = stamp; c.sides = sides; c.mask = mask; c.vs = vs_s; c.vmask_g = g_vmask;
        c.log_unions = a.log_unions ? a.log_unions + (size_t)b * a.log_cap * 8 * 2 : nullptr;
        c.union_cap = a.log_cap * 8; c.nunions = 0; c.ulog_cap = L.ulog_cap;
        c.ec = n; c.T = T; c.status = 0; c.stamp_id = 0;
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
            if (outcome == MCNN_POP_COLLAPSED || c.nunions != s_pub) {
                __threadfence_block();
                s_pub = min(c.nunions, c.ulog_cap);
            }
            if (c.status) break;
        }
        __threadfence_block();
        s_pub = min(c.nunions, c.ulog_cap);
        __threadfence_block();
        s_done = 1;
        if (a.log_npops) a.log_npops[b] = p;
        if (a.log_nunions) a.log_nunions[b] = c.nunions;
        s_status = c.status;
        s_ec = c.ec;
    } else if (tid == 32) {
        Groups g;
        g.ghead = ghead; g.gtail = gtail; g.glen = glen; g.gn = gn;
        g.gn_count = n; g.gn_cap = L.gn_cap; g.status = 0;
        int seen = 0;
        while (true) {
            const int d = s_done;
            const int avail = s_pub;
            while (seen < avail) {
                const uint32_t w = ulog[seen++];
                if (g.status == 0) g.row_add((int)(w & 0xFFFFu), (int)(w >> 16));
            }
            if (d) break;
            __nanosleep(64);
        }
        s_gstatus = g.status;
    }
    __syncthreads();
    int status = s_status ? s_status : s_gstatus;
    // ---- phase E: Mesh.clean (mesh.py:50-71) + group export ---------------------------------------------------
    // E1: survivor ranks; dead -> 0
    idx_t* newid = order;
    int carry = 0, total = 0;
    for (int base = 0; base < n; base += POOL_THREADS) {
        const int i = base + tid;
        const int m = (i < n) ? mask[i] : 0;
        int tot;
        const int pre = block_scan_excl(m, scan, tot);
        if (i < n) newid[i] = (idx_t)(m ? carry + pre : 0);
        carry += tot;
    }
    total = carry;
    __syncthreads();
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
        const uint32_t w = evp[i];
        int v0 = (int)(w & 0xFFFFu), v1 = (int)(w >> 16);
        while (rep[v0] != v0) v0 = rep[v0];
        while (rep[v1] != v1) v1 = rep[v1];
        g_edges[2 * r] = v0;
        g_edges[2 * r + 1] = v1;
    }
    for (int i = total * 4 + tid; i < T * 4; i += POOL_THREADS) { o_gemm[i] = 0; o_sides[i] = 0; }
    if (a.edge_mask_out)
        for (int i = tid; i < E_in; i += POOL_THREADS) a.edge_mask_out[(size_t)b * E_in + i] = (i < n) ? mask[i] : 0;
    if (with_vs)
        for (int i = tid; i < 3 * V; i += POOL_THREADS) g_vs[i] = vs_s[i];
    // E3: ve lists -> CSR with re-indexed ids
    carry = 0;
    for (int base = 0; base < V; base += POOL_THREADS) {
        const int v = base + tid;
        int len = 0;
        if (v < V)
            for (unsigned nd = (unsigned short)vhead[v]; nd != NIL16; nd = vn[nd] >> 16) ++len;
        int tot;
        const int pre = block_scan_excl(len, scan, tot);
        if (v < V) {
            int pos = carry + pre;
            g_ve_ptr[v] = pos;
            for (unsigned nd = (unsigned short)vhead[v]; nd != NIL16;) {
                const uint32_t w = vn[nd];
                g_ve_idx[pos++] = newid[w & 0xFFFFu];
                nd = w >> 16;
            }
        }
        carry += tot;
    }
    if (tid == 0) g_ve_ptr[V] = carry;
    __syncthreads();
    // E4: groups.  the ve node array is free now: occ_cnt / col_cnt live there (2 * int32[n] <= 4 * ve_cap bytes)
    int* occ_cnt = reinterpret_cast<int*>(smem + L.vn);
    int* col_cnt = occ_cnt + n;
    for (int i = tid; i < 2 * n; i += POOL_THREADS) occ_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += POOL_THREADS) {
        const bool live = mask[i];
        for (unsigned nd = (unsigned short)ghead[i]; nd != NIL16;) {
            const unsigned long long w = gn[nd];
            const int m = (int)((unsigned)w & 0xFFFFu);
            atomicAdd(&occ_cnt[m], (int)(((unsigned)w >> 16) & 0xFFFFu));       // every row, un-clamped
            if (live) atomicAdd(&col_cnt[m], 1);
            nd = (unsigned)(w >> 32) & 0xFFFFu;
        }
    }
    __syncthreads();
    float* o_occ = a.occ + (size_t)b * E_in;
    for (int i = tid; i < E_in; i += POOL_THREADS) o_occ[i] = (i < n) ? (float)occ_cnt[i] : 1.f;
    __syncthreads();
    // CSR row pointers over the surviving rows (in new order == old order restricted to survivors)
    int32_t* o_rowptr = a.grp_rowptr + (size_t)b * (T + 1);
    int32_t* o_cols = a.grp_cols + (size_t)b * a.nnz_cap;
    int32_t* o_colptr = a.grp_colptr + (size_t)b * (E_in + 1);
    int32_t* o_rows = a.grp_rows + (size_t)b * a.nnz_cap;
    carry = 0;
    for (int base = 0; base < n; base += POOL_THREADS) {
        const int i = base + tid;
        const int len = (i < n && mask[i] && newid[i] < T) ? glen[i] : 0;
        int tot;
        const int pre = block_scan_excl(len, scan, tot);
        if (len) occ_cnt[i] = carry + pre;          // row start, reusing occ_cnt (already exported)
        carry += tot;
    }
    const int nnz = carry;
    __syncthreads();
    const bool nnz_ok = nnz <= a.nnz_cap;
    if (!nnz_ok && status == 0) status = MCNN_ST_NNZ_OVERFLOW;
    for (int i = tid; i < n; i += POOL_THREADS) {
        if (!(mask[i] && newid[i] < T)) continue;
        const int r = newid[i];
        const int start = occ_cnt[i];
        o_rowptr[r] = start;
        if (!nnz_ok) continue;
        int len = 0;
        for (unsigned nd = (unsigned short)ghead[i]; nd != NIL16;) {      // insertion sort, ascending member id
            const unsigned long long w = gn[nd];
            const int m = (int)((unsigned)w & 0xFFFFu);
            int p = start + len;
            while (p > start && o_cols[p - 1] > m) { o_cols[p] = o_cols[p - 1]; --p; }
            o_cols[p] = m;
            ++len;
            nd = (unsigned)(w >> 32) & 0xFFFFu;
        }
    }
    for (int r = total + tid; r <= T; r += POOL_THREADS) o_rowptr[r] = nnz;
    __syncthreads();
    // CSC: column pointers, then a counting fill followed by a per-column insertion sort (ascending pooled id)
    carry = 0;
    for (int base = 0; base < n; base += POOL_THREADS) {
        const int i = base + tid;
        const int cnt = (i < n) ? col_cnt[i] : 0;
        int tot;
        const int pre = block_scan_excl(cnt, scan, tot);
        if (i < n) { o_colptr[i] = carry + pre; occ_cnt[i] = carry + pre; }   // occ_cnt = fill cursor
        carry += tot;
    }
    for (int i = n + tid; i <= E_in; i += POOL_THREADS) o_colptr[i] = nnz;
    __syncthreads();
    if (nnz_ok) {
        for (int i = tid; i < n; i += POOL_THREADS) {
            if (!(mask[i] && newid[i] < T)) continue;
            const int r = newid[i];
            for (unsigned nd = (unsigned short)ghead[i]; nd != NIL16;) {
                const unsigned long long w = gn[nd];
                const int pos = atomicAdd(&occ_cnt[(int)((unsigned)w & 0xFFFFu)], 1);
                o_rows[pos] = r;
                nd = (unsigned)(w >> 32) & 0xFFFFu;
            }
        }
        __syncthreads();
        for (int j = tid; j < n; j += POOL_THREADS) {
            const int beg = o_colptr[j], end = beg + col_cnt[j];
            for (int p = beg + 1; p < end; ++p) {
                const int v = o_rows[p];
                int q = p;
                while (q > beg && o_rows[q - 1] > v) { o_rows[q] = o_rows[q - 1]; --q; }
                o_rows[q] = v;
            }
        }
    }
    if (tid == 0) {
        a.ec_out[b] = total;
        a.status[b] = status;
    }
}

// ---------------------------------------------------------------------------------------------------------
// feature kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void pool_norms_kernel(const float* __restrict__ x, float* __restrict__ norms, int M, int C) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* p = x + (size_t)m * C;
    float s = 0.f;
    if ((C & 3) == 0) {
        for (int c = 0; c < C; c += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + c));
            s = __fadd_rn(s, __fmul_rn(v.x, v.x));
            s = __fadd_rn(s, __fmul_rn(v.y, v.y));
            s = __fadd_rn(s, __fmul_rn(v.z, v.z));
            s = __fadd_rn(s, __fmul_rn(v.w, v.w));
        }
    } else {
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(p + c);
            s = __fadd_rn(s, __fmul_rn(v, v));
        }
    }
    norms[m] = s;
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
template <int MODE, bool VEC>
__global__ void segment_kernel(const float* __restrict__ in, float* __restrict__ out, const int32_t* __restrict__ ptr,
                               const int32_t* __restrict__ idx, const int32_t* __restrict__ aux_ptr,
                               const float* __restrict__ occ, const int32_t* __restrict__ valid, int B, int R_out,
                               int R_in, int ptr_stride, int aux_stride, int occ_stride, int C, int nnz_cap) {
    const int cq = (C + 3) >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)B * R_out * cq) return;
    const int q = (int)(t % cq);
    const long long br = t / cq;
    const int r = (int)(br % R_out), b = (int)(br / R_out);
    const int c = q * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < __ldg(valid + b) && r < ptr_stride - 1) {
        const int32_t* pp = ptr + (size_t)b * ptr_stride;
        const int beg = __ldg(pp + r), end = __ldg(pp + r + 1);
        const int32_t* ii = idx + (size_t)b * nnz_cap;
        const float* src = in + (size_t)b * R_in * C;
        for (int p = beg; p < end; ++p) {
            const int j = __ldg(ii + p);
            float w = 1.f;
            if (MODE == 1) {
                const int32_t* ap = aux_ptr + (size_t)b * aux_stride;
                w = 1.f / (float)(__ldg(ap + j + 1) - __ldg(ap + j));
            } else if (MODE == 3) {
                w = 1.f / __ldg(occ + (size_t)b * occ_stride + j);
            }
            acc_row<VEC>(acc, src + (size_t)j * C, C, c, w);
        }
        float s = 1.f;
        if (MODE == 0) s = 1.f / (float)(end - beg);
        if (MODE == 2) s = 1.f / __ldg(occ + (size_t)b * occ_stride + r);
        acc.x *= s; acc.y *= s; acc.z *= s; acc.w *= s;
    }
    store_row<VEC>(out + ((size_t)b * R_out + r) * C, C, c, acc);
}

template <int MODE>
static int launch_segment(const float* in, float* out, const int32_t* ptr, const int32_t* idx, const int32_t* aux_ptr,
                          const float* occ, const int32_t* valid, int B, int R_out, int R_in, int ptr_stride,
                          int aux_stride, int occ_stride, int C, int nnz_cap, cudaStream_t st) {
    const long long total = (long long)B * R_out * ((C + 3) / 4);
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if ((C & 3) == 0)
        segment_kernel<MODE, true><<<blocks, threads, 0, st>>>(in, out, ptr, idx, aux_ptr, occ, valid, B, R_out, R_in,
                                                                ptr_stride, aux_stride, occ_stride, C, nnz_cap);
    else
        segment_kernel<MODE, false><<<blocks, threads, 0, st>>>(in, out, ptr, idx, aux_ptr, occ, valid, B, R_out, R_in,
                                                                 ptr_stride, aux_stride, occ_stride, C, nnz_cap);
    return after_launch();
}

}  // namespace mcnn

using namespace mcnn;

extern "C" size_t mcnn_pool_smem_bytes(int E_in, int V, int ve_cap, int with_vs) {
    return pool_layout(E_in, V, ve_cap, with_vs != 0).total + 64;
}

extern "C" size_t mcnn_pool_smem_limit(void) { return 227 * 1024; }

extern "C" int mcnn_pool_norms(const float* x, float* norms, int B, int E, int C, void* stream) {
    if (!x || !norms || B <= 0 || E <= 0 || C <= 0) return MCNN_E_BADARG;
    const int M = B * E;
    pool_norms_kernel<<<ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(x, norms, M, C);
    return after_launch();
}

extern "C" int mcnn_pool_collapse(const mcnn_pool_args* a, void* stream) {
    if (!a || a->B <= 0 || a->E_in <= 0 || a->T <= 0 || a->V <= 0 || !a->gemm_in || !a->sides_in || !a->ec_in ||
        !a->norms || !a->edges || !a->ve_ptr || !a->ve_idx || !a->v_mask || !a->gemm_out || !a->sides_out ||
        !a->ec_out || !a->grp_rowptr || !a->grp_cols || !a->grp_colptr || !a->grp_rows || !a->occ || !a->status)
        return MCNN_E_BADARG;
    if (a->E_in > 10900 || a->V > 32767 || a->ve_cap > 32767 || a->ve_cap < 2 * 0) return MCNN_E_TOO_LARGE;
    if ((size_t)a->ve_cap * sizeof(uint32_t) < (size_t)a->E_in * 2 * sizeof(int)) return MCNN_E_BADARG;  // occ/col counters alias the ve nodes
    const size_t smem = pool_layout(a->E_in, a->V, a->ve_cap, a->vs != nullptr).total;
    if (smem > mcnn_pool_smem_limit()) return MCNN_E_TOO_LARGE;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(pool_collapse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_error = e; return MCNN_E_CUDA; }
        configured = smem;
    }
    pool_collapse_kernel<<<a->B, POOL_THREADS, smem, (cudaStream_t)stream>>>(*a);
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
