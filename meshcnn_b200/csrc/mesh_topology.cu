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
