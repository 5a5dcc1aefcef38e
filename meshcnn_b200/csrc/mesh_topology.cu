// Host-side topology builder (no device code): the sequential parts of the reference's mesh preparation
// (models/layers/mesh_prepare.py:90-113 remove_non_manifolds, :116-161 build_gemm) for one mesh, as plain C++.
// SURVEY.md section 8f row N2.  Integer work only, plus the face-order accumulation of the edge areas, so the results are
// bit-identical to the NumPy formulation in meshcnn_b200/mesh_prepare.py (tests/test_mesh_prepare.py compares both with the
// reference's fixtures); the floating-point features stay in NumPy.
#include <stdint.h>

#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/meshcnn_b200.h"

// faces [nf][3] int64, face_areas [nf] double (computed by the caller).  keep [nf]: 1 for the faces that survive
// remove_non_manifolds (zero area, or a DIRECTED edge that an accepted face already uses -> dropped).  Returns the number kept.
extern "C" int mcnn_host_drop_non_manifold(const int64_t* faces, const double* face_areas, int nf, int nv, uint8_t* keep) {
    if (!faces || !face_areas || !keep || nf < 0 || nv <= 0) return MCNN_E_BADARG;
    std::unordered_set<uint64_t> seen;
    seen.reserve((size_t)nf * 3 * 2);
    const uint64_t m = (uint64_t)nv + 1;
    int kept = 0;
    for (int f = 0; f < nf; ++f) {
        keep[f] = 0;
        if (face_areas[f] == 0.0) continue;
        const uint64_t a = (uint64_t)faces[3 * f], b = (uint64_t)faces[3 * f + 1], c = (uint64_t)faces[3 * f + 2];
        const uint64_t k0 = a * m + b, k1 = b * m + c, k2 = c * m + a;
        if (seen.count(k0) || seen.count(k1) || seen.count(k2)) continue;
        seen.insert(k0); seen.insert(k1); seen.insert(k2);
        keep[f] = 1;
        ++kept;
    }
    return kept;
}

// build_gemm: edges in order of first appearance over (face, corner); per edge the (up to) two faces' neighbour pairs in
// gemm / sides; vertex -> incident edges (CSR, edge creation order); per-edge area sums accumulated in face order.
// Outputs are sized for the worst case E = 3 * nf: edges [3nf][2] int32, gemm / sides [3nf][4] int64 (-1 = boundary),
// ve_ptr [nv + 1] int32, ve_idx [6 nf] int32, edge_area_sum [3nf] double (sum of face_area / 3 over the edge's faces).
// Returns E, or MCNN_E_BADARG; -100 when an edge has more than two faces (the reference's build_gemm would index out of range).
extern "C" int mcnn_host_build_gemm(const int64_t* faces, const double* face_areas, int nf, int nv, int32_t* edges, int64_t* gemm,
                                    int64_t* sides, int32_t* ve_ptr, int32_t* ve_idx, double* edge_area_sum) {
    if (!faces || !face_areas || !edges || !gemm || !sides || !ve_ptr || !ve_idx || !edge_area_sum || nf < 0 || nv <= 0)
        return MCNN_E_BADARG;
    std::unordered_map<uint64_t, int32_t> edge_of;
    edge_of.reserve((size_t)nf * 3);
    std::vector<uint8_t> nb_count;
    nb_count.reserve((size_t)nf * 3);
    const uint64_t m = (uint64_t)nv + 1;
    int E = 0;
    for (int f = 0; f < nf; ++f) {
        int32_t he[3];
        for (int k = 0; k < 3; ++k) {
            const int64_t a = faces[3 * f + k], b = faces[3 * f + (k + 1) % 3];
            if (a < 0 || b < 0 || a >= nv || b >= nv) return MCNN_E_BADARG;
            const int64_t lo = a < b ? a : b, hi = a < b ? b : a;
            const uint64_t key = (uint64_t)lo * m + (uint64_t)hi;
            auto it = edge_of.find(key);
            if (it == edge_of.end()) {
                it = edge_of.emplace(key, E).first;
                edges[2 * E] = (int32_t)lo;
                edges[2 * E + 1] = (int32_t)hi;
                for (int j = 0; j < 4; ++j) { gemm[4 * (size_t)E + j] = -1; sides[4 * (size_t)E + j] = -1; }
                edge_area_sum[E] = 0.0;
                nb_count.push_back(0);
                ++E;
            }
            he[k] = it->second;
            edge_area_sum[he[k]] += face_areas[f] / 3;
        }
        int base[3];
        for (int k = 0; k < 3; ++k) {
            if (nb_count[he[k]] >= 4) return -100;
            base[k] = nb_count[he[k]];
        }
        for (int k = 0; k < 3; ++k) {
            const int e = he[k], n1 = he[(k + 1) % 3], n2 = he[(k + 2) % 3];
            gemm[4 * (size_t)e + base[k]] = n1;
            gemm[4 * (size_t)e + base[k] + 1] = n2;
            nb_count[e] += 2;
        }
        for (int k = 0; k < 3; ++k) {                 // in next's row e is the next-next entry, in next-next's row the next entry
            const int e = he[k];
            sides[4 * (size_t)e + base[k]] = base[(k + 1) % 3] + 1;
            sides[4 * (size_t)e + base[k] + 1] = base[(k + 2) % 3];
        }
    }
    // vertex -> edges, edge creation order (counting sort over the endpoints)
    for (int v = 0; v <= nv; ++v) ve_ptr[v] = 0;
    for (int e = 0; e < E; ++e) { ++ve_ptr[edges[2 * e] + 1]; ++ve_ptr[edges[2 * e + 1] + 1]; }
    for (int v = 0; v < nv; ++v) ve_ptr[v + 1] += ve_ptr[v];
    std::vector<int32_t> cur(ve_ptr, ve_ptr + nv);
    for (int e = 0; e < E; ++e) {
        ve_idx[cur[edges[2 * e]]++] = e;
        ve_idx[cur[edges[2 * e + 1]]++] = e;
    }
    return E;
}

// ---- edge features up to the arccos (models/layers/mesh_prepare.py:303-427) -------------------------------------------------
// Same operations in the same order as the NumPy formulation in meshcnn_b200/mesh_prepare.py (products, then left-to-right sums
// of three terms; IEEE division and square root are correctly rounded in both), so the results are bit-identical; the arccos
// itself stays in NumPy, whose implementation differs between machines.  Compiled without FMA contraction (Makefile).
#include <math.h>

namespace {
struct V3 { double x, y, z; };
inline V3 ld(const double* vs, int i) { return V3{vs[3 * i], vs[3 * i + 1], vs[3 * i + 2]}; }
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm3(V3 a) { return sqrt((a.x * a.x + a.y * a.y) + a.z * a.z); }
inline V3 cross3(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 divs(V3 a, double d) { return V3{a.x / d, a.y / d, a.z / d}; }
inline double clip1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }   // NaN passes through, as np.clip does
}  // namespace

// edges [E][2] int32, gemm [E][4] int64, vs [nv][3] double.  Outputs (all [E] double unless noted): lengths; dih_arg = clipped
// cosine between the two face normals; opp_arg [2][E] = clipped cosines of the two opposite angles; ratio [2][E] = height / base
// of the two faces.  Returns 0, MCNN_E_BADARG, or -101 when a division NumPy would flag under errstate(divide='raise') occurs
// (a non-zero numerator over a zero base length).
extern "C" int mcnn_host_edge_features(const double* vs, int nv, const int32_t* edges, const int64_t* gemm, int E, double* lengths,
                                       double* dih_arg, double* opp_arg, double* ratio) {
    if (!vs || !edges || !gemm || !lengths || !dih_arg || !opp_arg || !ratio || nv <= 0 || E < 0) return MCNN_E_BADARG;
    int flagged = 0;
    for (int e = 0; e < E; ++e) {
        const int64_t* g = gemm + 4 * (size_t)e;
        // get_edge_points (mesh_prepare.py:361-399): boundary edges take the face that exists for both sides
        const int64_t b_id = g[0] == -1 ? g[2] : g[0], c_id = g[0] == -1 ? g[3] : g[1];
        const int64_t d_id = g[2] == -1 ? g[0] : g[2], e_id = g[2] == -1 ? g[1] : g[3];
        if (b_id < 0 || c_id < 0 || d_id < 0 || e_id < 0 || b_id >= E || c_id >= E || d_id >= E || e_id >= E) return MCNN_E_BADARG;
        const int32_t *ea = edges + 2 * (size_t)e, *eb = edges + 2 * b_id, *ec = edges + 2 * c_id, *ed = edges + 2 * d_id,
                      *ee = edges + 2 * e_id;
        const int first = (ea[1] == eb[0] || ea[1] == eb[1]) ? 1 : 0;
        const int second = (eb[1] == ec[0] || eb[1] == ec[1]) ? 1 : 0;
        const int third = (ed[1] == ee[0] || ed[1] == ee[1]) ? 1 : 0;
        const int p[4] = {ea[first], ea[1 - first], eb[second], ed[third]};
        for (int k = 0; k < 4; ++k)
            if (p[k] < 0 || p[k] >= nv) return MCNN_E_BADARG;
        const V3 P0 = ld(vs, p[0]), P1 = ld(vs, p[1]);
        lengths[e] = norm3(sub(P0, P1));
        V3 nrm[2];
        for (int s = 0; s < 2; ++s) {
            const V3 Ps = ld(vs, p[s]), Po = ld(vs, p[s + 2]), Pt = ld(vs, p[1 - s]);
            // face normal of side s
            const V3 n = cross3(sub(Po, Ps), sub(Pt, Ps));
            nrm[s] = divs(n, norm3(n) + 0.1);
            // opposite angle: between the two edges that meet at the opposite vertex
            V3 ua = sub(Ps, Po), ub = sub(Pt, Po);
            ua = divs(ua, norm3(ua) + 0.1);
            ub = divs(ub, norm3(ub) + 0.1);
            opp_arg[(size_t)s * E + e] = clip1(dot3(ua, ub));
            // height / base ratio
            const double base_len = norm3(sub(Ps, Pt));
            const V3 ab = sub(Pt, Ps);
            const double proj = dot3(ab, sub(Po, Ps)) / (norm3(ab) + 0.1);
            if (base_len == 0.0 && proj != 0.0 && proj == proj) flagged = 1;
            const double t = proj / base_len;
            const V3 closest = V3{Ps.x + t * ab.x, Ps.y + t * ab.y, Ps.z + t * ab.z};
            const double h = norm3(sub(Po, closest));
            if (base_len == 0.0 && h != 0.0 && h == h) flagged = 1;
            ratio[(size_t)s * E + e] = h / base_len;
        }
        dih_arg[e] = clip1(dot3(nrm[0], nrm[1]));
    }
    return flagged ? -101 : 0;
}
