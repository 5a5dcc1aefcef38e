"""Edge-topology builder and the five input edge features (row N2 of SURVEY.md §8f, needed here to
make benchmark / test inputs without the reference on the GPU box).

Conventions reproduced (the reference is models/layers/mesh_prepare.py):
  * faces are visited in file order, an edge id is assigned the first time its (sorted) vertex pair is
    seen                                                              (mesh_prepare.py:129-146)
  * ``gemm_edges[e]`` holds two slot pairs, (0,1) for the first face containing ``e`` and (2,3) for the
    second; inside a pair: the face's next and next-next edge in winding order        (:148-152)
  * ``sides[e, k]`` = slot of ``e`` inside ``gemm_edges[gemm_edges[e, k]]``            (:153-156)
  * ``ve[v]`` lists the incident edge ids in creation order                            (:142-143)
  * edge_areas[e] = sum over its faces of area/3, normalised by total area             (:147, :161)
  * the 5 features: dihedral angle, 2 sorted opposite angles, 2 sorted height/base ratios (:310-427)

This is an array-at-a-time restatement (no per-face dict loop), not a copy: edge ids come from a
first-occurrence ``np.unique`` over packed vertex-pair keys and the slot tables from occurrence ranks.
"""
import ntpath

import os

import numpy as np


class MeshData:
    """Plain picklable container for one mesh's topology + features."""
    __slots__ = ('vs', 'edges', 'gemm_edges', 'sides', 've', 'v_mask', 'edges_count', 'filename',
                 'edge_lengths', 'edge_areas', 'features')


def read_obj(path):
    """OBJ reader: 'v x y z' and triangular 'f a b c' records, 1-based or negative indices, a/b/c forms
    (mesh_prepare.py:65-87)."""
    vs, faces = [], []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                vs.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == 'f':
                ids = [int(t.split('/')[0]) for t in tok[1:]]
                if len(ids) != 3:
                    raise AssertionError('non-triangular face in %s' % path)
                faces.append([i - 1 if i >= 0 else len(vs) + i for i in ids])
    vs = np.asarray(vs, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    if faces.size and not ((faces >= 0) & (faces < len(vs))).all():
        raise AssertionError('face index out of range in %s' % path)
    return vs, faces


def face_areas_of(vs, faces):
    n = np.cross(vs[faces[:, 1]] - vs[faces[:, 0]], vs[faces[:, 2]] - vs[faces[:, 1]])
    return 0.5 * np.sqrt((n ** 2).sum(1))


# The sequential parts (the face sweep and build_gemm) run in the library's host-side C++ (csrc/mesh_topology.cu); the NumPy /
# Python formulations below are the same algorithm and stay as the cross-check (tests/test_mesh_prepare.py) and for
# MESHCNN_B200_NATIVE_PREPARE=0.
NATIVE = os.environ.get('MESHCNN_B200_NATIVE_PREPARE', '1') != '0'


def drop_non_manifold(vs, faces):
    """Drop zero-area faces and faces repeating an already-used *directed* edge (mesh_prepare.py:90-113).
    The acceptance of a face depends on the faces accepted before it, so this is a sequential sweep."""
    if not NATIVE:
        return drop_non_manifold_py(vs, faces)
    from . import _lib
    faces = np.ascontiguousarray(faces, dtype=np.int64)
    areas = np.ascontiguousarray(face_areas_of(vs, faces), dtype=np.float64) if len(faces) else np.zeros(0)
    keep = np.zeros(len(faces), dtype=np.uint8)
    rc = _lib.lib().mcnn_host_drop_non_manifold(faces.ctypes.data, areas.ctypes.data, len(faces), int(vs.shape[0]), keep.ctypes.data)
    if rc < 0:
        raise RuntimeError('mcnn_host_drop_non_manifold failed (%d)' % rc)
    keep = keep.astype(bool)
    return faces[keep], areas[keep]


def drop_non_manifold_py(vs, faces):
    areas = face_areas_of(vs, faces)
    keep = np.ones(len(faces), dtype=bool)
    seen = set()
    nv = int(vs.shape[0]) + 1
    for fi in range(len(faces)):
        if areas[fi] == 0:
            keep[fi] = False
            continue
        a, b, c = (int(t) for t in faces[fi])
        k0, k1, k2 = a * nv + b, b * nv + c, c * nv + a
        if k0 in seen or k1 in seen or k2 in seen:
            keep[fi] = False
            continue
        seen.add(k0); seen.add(k1); seen.add(k2)
    return faces[keep], areas[keep]


def build_topology(vs, faces, face_areas=None):
    """Returns edges[E,2] int32, gemm_edges[E,4] int64, sides[E,4] int64, ve (list of lists),
    edge_areas[E] float32."""
    if not NATIVE:
        return build_topology_py(vs, faces, face_areas)
    from . import _lib
    faces = np.ascontiguousarray(faces, dtype=np.int64)
    nf, nv = len(faces), int(vs.shape[0])
    if face_areas is None:
        face_areas = face_areas_of(vs, faces)
    face_areas = np.ascontiguousarray(face_areas, dtype=np.float64)
    cap = max(1, 3 * nf)
    edges = np.empty((cap, 2), dtype=np.int32)
    gemm = np.empty((cap, 4), dtype=np.int64)
    sides = np.empty((cap, 4), dtype=np.int64)
    ve_ptr = np.empty(nv + 1, dtype=np.int32)
    ve_idx = np.empty(2 * cap, dtype=np.int32)
    ea = np.empty(cap, dtype=np.float64)
    E = _lib.lib().mcnn_host_build_gemm(faces.ctypes.data, face_areas.ctypes.data, nf, nv, edges.ctypes.data, gemm.ctypes.data,
                                        sides.ctypes.data, ve_ptr.ctypes.data, ve_idx.ctypes.data, ea.ctypes.data)
    if E == -100:
        raise ValueError('edge shared by more than two faces (non-manifold input)')
    if E < 0:
        raise RuntimeError('mcnn_host_build_gemm failed (%d)' % E)
    idx = ve_idx[:2 * E].tolist()
    ptr = ve_ptr.tolist()
    ve = [idx[ptr[v]:ptr[v + 1]] for v in range(nv)]
    # float32 array / float64 scalar: the result dtype follows numpy's promotion rules, as in the reference
    edge_areas = ea[:E].astype(np.float32) / np.sum(face_areas)
    return edges[:E].copy(), gemm[:E].copy(), sides[:E].copy(), ve, edge_areas


def build_topology_py(vs, faces, face_areas=None):
    nf = len(faces)
    if face_areas is None:
        face_areas = face_areas_of(vs, faces)
    nv = int(vs.shape[0])
    # half-edges in (face, corner) order == the reference's visiting order
    a = faces.ravel()
    b = faces[:, [1, 2, 0]].ravel()
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * (nv + 1) + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind='stable')           # unique keys by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    he_edge = rank[inv]                                 # edge id of each half-edge
    E = len(uniq)
    edges = np.stack([lo[first[order]], hi[first[order]]], 1).astype(np.int32)
    # occurrence rank of each half-edge among the half-edges of its edge (0 = first face, 1 = second)
    srt = np.argsort(he_edge, kind='stable')
    start = np.r_[0, np.cumsum(np.bincount(he_edge, minlength=E))[:-1]]
    occ = np.empty(3 * nf, dtype=np.int64)
    occ[srt] = np.arange(3 * nf) - start[he_edge[srt]]
    if occ.max(initial=0) > 1:
        raise ValueError('edge shared by more than two faces (non-manifold input)')
    he = he_edge.reshape(nf, 3)
    base = (2 * occ).reshape(nf, 3)                     # slot-pair base of this face in each edge's row
    nxt, nxt2 = he[:, [1, 2, 0]], he[:, [2, 0, 1]]
    bn, bn2 = base[:, [1, 2, 0]], base[:, [2, 0, 1]]
    gemm = -np.ones((E, 4), dtype=np.int64)
    sides = -np.ones((E, 4), dtype=np.int64)
    gemm[he.ravel(), base.ravel()] = nxt.ravel()
    gemm[he.ravel(), base.ravel() + 1] = nxt2.ravel()
    sides[he.ravel(), base.ravel()] = bn.ravel() + 1    # in next's row, e is the next-next entry
    sides[he.ravel(), base.ravel() + 1] = bn2.ravel()   # in next-next's row, e is the next entry
    # vertex -> incident edges in edge creation order
    ev = edges.astype(np.int64).ravel()
    eid = np.repeat(np.arange(E), 2)
    o = np.argsort(ev, kind='stable')
    cnt = np.bincount(ev, minlength=nv)
    splits = np.cumsum(cnt)[:-1]
    ve = [seg.tolist() for seg in np.split(eid[o], splits)]
    # edge areas: sequential float64 accumulation order of the reference is face order; np.add.at keeps it
    ea = np.zeros(E, dtype=np.float64)
    np.add.at(ea, he.ravel(), np.repeat(face_areas / 3, 3))
    # float32 array / float64 scalar: the result dtype follows numpy's promotion rules, as in the reference
    edge_areas = ea.astype(np.float32) / np.sum(face_areas)
    return edges, gemm, sides, ve, edge_areas


def _div_eps(x, eps):
    x = x.copy()
    if eps == 0:
        x[x == 0] = 0.1
    else:
        x += eps
    return x


def edge_points(edges, gemm):
    """Four vertex ids per edge: its two endpoints (oriented) and the opposite vertex of each incident face
    (mesh_prepare.py:361-399)."""
    g = gemm
    b_id = np.where(g[:, 0] == -1, g[:, 2], g[:, 0])
    c_id = np.where(g[:, 0] == -1, g[:, 3], g[:, 1])
    d_id = np.where(g[:, 2] == -1, g[:, 0], g[:, 2])
    e_id = np.where(g[:, 2] == -1, g[:, 1], g[:, 3])
    ea, eb, ec_, ed, ee = edges, edges[b_id], edges[c_id], edges[d_id], edges[e_id]
    first = ((ea[:, 1] == eb[:, 0]) | (ea[:, 1] == eb[:, 1])).astype(np.int64)
    second = ((eb[:, 1] == ec_[:, 0]) | (eb[:, 1] == ec_[:, 1])).astype(np.int64)
    third = ((ed[:, 1] == ee[:, 0]) | (ed[:, 1] == ee[:, 1])).astype(np.int64)
    r = np.arange(len(edges))
    return np.stack([ea[r, first], ea[r, 1 - first], eb[r, second], ed[r, third]], 1).astype(np.int32)


def extract_features(vs, edges, gemm):
    """[5, E] float64 features + edge lengths (mesh_prepare.py:303-427)."""
    if NATIVE:
        return extract_features_native(vs, edges, gemm)
    return extract_features_py(vs, edges, gemm)


def extract_features_native(vs, edges, gemm):
    """Everything up to the arccos in the library's host-side C++ (mcnn_host_edge_features: same operations in the same order,
    bit-identical); arccos, the pair sorts and the stacking in NumPy."""
    from . import _lib
    vs = np.ascontiguousarray(vs, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.int32)
    gemm = np.ascontiguousarray(gemm, dtype=np.int64)
    E = len(edges)
    lengths = np.empty(E, dtype=np.float64)
    dih = np.empty(E, dtype=np.float64)
    opp = np.empty((2, E), dtype=np.float64)
    rat = np.empty((2, E), dtype=np.float64)
    rc = _lib.lib().mcnn_host_edge_features(vs.ctypes.data, int(vs.shape[0]), edges.ctypes.data, gemm.ctypes.data, E,
                                            lengths.ctypes.data, dih.ctypes.data, opp.ctypes.data, rat.ctypes.data)
    if rc == -101:
        raise ValueError('bad features')
    if rc != 0:
        raise RuntimeError('mcnn_host_edge_features failed (%d)' % rc)
    dihedral = (np.pi - np.arccos(dih))[None]
    ang = np.sort(np.arccos(opp), axis=0)
    rat = np.sort(rat, axis=0)
    return np.concatenate([dihedral, ang, rat], 0), lengths


def extract_features_py(vs, edges, gemm):
    ep = edge_points(edges, gemm)
    lengths = np.linalg.norm(vs[ep[:, 0]] - vs[ep[:, 1]], axis=1)

    def normals(side):
        s = side // 2
        n = np.cross(vs[ep[:, s + 2]] - vs[ep[:, s]], vs[ep[:, 1 - s]] - vs[ep[:, s]])
        return n / _div_eps(np.linalg.norm(n, axis=1), 0.1)[:, None]

    def opposite(side):
        s = side // 2
        ua = vs[ep[:, s]] - vs[ep[:, s + 2]]
        ub = vs[ep[:, 1 - s]] - vs[ep[:, s + 2]]
        ua = ua / _div_eps(np.linalg.norm(ua, axis=1), 0.1)[:, None]
        ub = ub / _div_eps(np.linalg.norm(ub, axis=1), 0.1)[:, None]
        return np.arccos(np.sum(ua * ub, axis=1).clip(-1, 1))

    def ratio(side):
        s = side // 2
        base_len = np.linalg.norm(vs[ep[:, s]] - vs[ep[:, 1 - s]], axis=1)
        po, pa, pb = vs[ep[:, s + 2]], vs[ep[:, s]], vs[ep[:, 1 - s]]
        ab = pb - pa
        proj = np.sum(ab * (po - pa), axis=1) / _div_eps(np.linalg.norm(ab, axis=1), 0.1)
        closest = pa + (proj / base_len)[:, None] * ab
        return np.linalg.norm(po - closest, axis=1) / base_len

    with np.errstate(divide='raise'):
        try:
            dihedral = (np.pi - np.arccos(np.sum(normals(0) * normals(3), axis=1).clip(-1, 1)))[None]
            ang = np.sort(np.stack([opposite(0), opposite(3)], 0), axis=0)
            rat = np.sort(np.stack([ratio(0), ratio(3)], 0), axis=0)
        except FloatingPointError as e:
            raise ValueError('bad features') from e
    return np.concatenate([dihedral, ang, rat], 0), lengths


def dihedral_angles(vs, edges, gemm):
    """The first feature channel alone (mesh_prepare.py:325-330); slide_vertices ranks vertices by it."""
    ep = edge_points(edges, gemm)

    def normals(s):
        n = np.cross(vs[ep[:, s + 2]] - vs[ep[:, s]], vs[ep[:, 1 - s]] - vs[ep[:, s]])
        return n / _div_eps(np.linalg.norm(n, axis=1), 0.1)[:, None]
    return np.pi - np.arccos(np.sum(normals(0) * normals(1), axis=1).clip(-1, 1))


# ---------------------------------------------------------------------------------------------------------------
# data augmentation (mesh_prepare.py:175-294).  Every random draw is made with the same numpy call, in the same order,
# on the same generator (the global np.random unless `rng` is given) as the reference makes it, so that for an equal
# generator state the augmented mesh is bit-identical (tests/test_mesh_prepare.py pins this against reference-made goldens).
# ---------------------------------------------------------------------------------------------------------------
def scale_vertices(vs, rng=np.random, mean=1.0, var=0.1):
    """Anisotropic scaling: one normal draw per coordinate axis (mesh_prepare.py:204-206).  In place."""
    for axis in range(vs.shape[1]):
        vs[:, axis] = vs[:, axis] * rng.normal(mean, var)


def _edge_table(faces):
    """Edges in first-seen order over (face, corner): rows [v_lo, v_hi, first face, second face or -1] and the
    {(v_lo, v_hi): edge id} index (mesh_prepare.py:275-291)."""
    index, rows = {}, []
    for fi in range(len(faces)):
        f = faces[fi]
        for c in range(3):
            a, b = int(f[c]), int(f[(c + 1) % 3])
            key = (a, b) if a < b else (b, a)
            k = index.get(key)
            if k is None:
                index[key] = len(rows)
                rows.append([key[0], key[1], fi, -1])
            elif rows[k][2] == -1:
                rows[k][2] = fi
            else:
                rows[k][3] = fi
    return np.array(rows, dtype=np.int64).reshape(-1, 4), index


def flip_edges(vs, faces, prct, rng=np.random):
    """Random flips of flat interior edges, at most int(prct * #edges) of them (mesh_prepare.py:225-262).  `faces` is
    edited in place and returned.  The dihedral angles are those of the input mesh (not refreshed between flips) and an edge
    without a second face reads the LAST face of the array, both as in the reference."""
    table, index = _edge_table(faces)
    n_edges = len(table)
    unit = []
    for col in (2, 3):
        f = faces[table[:, col]]
        n = np.cross(vs[f[:, 2]] - vs[f[:, 1]], vs[f[:, 1]] - vs[f[:, 0]])
        unit.append(n / _div_eps(np.linalg.norm(n, axis=1), 0)[:, None])
    dihedral = np.pi - np.arccos(np.sum(unit[0] * unit[1], axis=1).clip(-1, 1))
    order = rng.permutation(n_edges)
    target = int(prct * n_edges)
    done = 0
    for k in order:
        if done == target:
            break
        if not dihedral[k] > 2.7:
            continue
        v0, v1, fa, fb = (int(t) for t in table[k])
        if fb == -1:
            continue
        opp = tuple(sorted(set(faces[fa].tolist()) ^ set(faces[fb].tolist())))
        if opp in index:
            continue
        cand = np.array([[v1, opp[0], opp[1]], [v0, opp[0], opp[1]]])
        if not (face_areas_of(vs, cand) > 0).all():
            continue
        del index[(v0, v1)]
        table[k, 0], table[k, 1] = opp[0], opp[1]
        index[opp] = k
        for fid, tri in ((fa, cand[0]), (fb, cand[1])):           # swap the corner that is not part of the new triangle
            face = faces[fid]
            tri_l, face_l = tri.tolist(), face.tolist()
            fresh = list(set(tri_l) - set(face_l))[0]
            for c in range(3):
                if face_l[c] not in tri_l:
                    face[c] = fresh
                    break
        for me, other in ((fa, fb), (fb, fa)):                     # the outer edges that changed face
            face = faces[me]
            for c in range(3):
                a, b = int(face[c]), int(face[(c + 1) % 3])
                key = (a, b) if a < b else (b, a)
                if key != opp:
                    row = table[index[key]]
                    for col in (2, 3):
                        if row[col] == other:
                            row[col] = me
        done += 1
    return faces


def slide_vertices(vs, edges, gemm, ve, prct, rng=np.random):
    """Slide up to int(prct * #vertices) vertices whose incident edges are all flat (dihedral > 2.65) 20-50 % of the way
    along a random incident edge (mesh_prepare.py:186-202).  In place; returns the fraction of vertices moved."""
    dihedral = dihedral_angles(vs, edges, gemm)
    order = rng.permutation(len(ve))
    target = int(prct * len(order))
    moved = 0
    for vi in order:
        if moved >= target:
            break
        inc = ve[vi]
        if min(dihedral[inc]) > 2.65:
            e = edges[rng.choice(inc)]
            other = e[1] if vi == e[0] else e[0]
            vs[vi] = vs[vi] + rng.uniform(0.2, 0.5) * (vs[other] - vs[vi])
            moved += 1
    return moved / len(ve)


def from_arrays(vs, faces, filename='synthetic', opt=None, rng=np.random):
    """Full preparation of one mesh from vertex / face arrays (mesh_prepare.from_scratch, :39-63).  `opt` (the reference's
    option namespace) switches the augmentations on: `num_aug > 1` and `scale_verts` / `flip_edges` / `slide_verts`."""
    vs = np.array(vs, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    faces, areas = drop_non_manifold(vs, faces)
    augment = opt is not None and getattr(opt, 'num_aug', 1) > 1
    if augment:                                                     # mesh_prepare.py:175-180
        if getattr(opt, 'scale_verts', False):
            scale_vertices(vs, rng)
        if getattr(opt, 'flip_edges', 0):
            faces = flip_edges(vs, faces, opt.flip_edges, rng)
    d = MeshData()
    d.vs = vs
    d.v_mask = np.ones(len(vs), dtype=bool)
    d.filename = filename
    d.edges, d.gemm_edges, d.sides, d.ve, d.edge_areas = build_topology(vs, faces, areas)
    d.edges_count = int(len(d.edges))
    if augment and getattr(opt, 'slide_verts', 0):                  # mesh_prepare.py:183-185
        slide_vertices(vs, d.edges, d.gemm_edges, d.ve, opt.slide_verts, rng)
    d.features, d.edge_lengths = extract_features(vs, d.edges, d.gemm_edges)
    return d


def from_obj(path, opt=None):
    vs, faces = read_obj(path)
    return from_arrays(vs, faces, filename=ntpath.split(path)[1], opt=opt)


# ---------------------------------------------------------------------------------------------------------------
# the .npz cache (mesh_prepare.py:6-37): <dir>/cache/<name>_<k>.npz with k = np.random.randint(0, num_aug), same keys as
# the reference's files.  `ve` is ragged; the reference leaves the conversion to np.savez_compressed, which modern numpy
# refuses -- it is stored as the 1-D object array of lists that older numpy produced, so files interchange.
# ---------------------------------------------------------------------------------------------------------------
CACHE_KEYS = ('gemm_edges', 'vs', 'edges', 'edges_count', 've', 'v_mask', 'filename', 'sides', 'edge_lengths', 'edge_areas', 'features')


def cache_path(file, num_aug, rng=np.random):
    import os
    stem = os.path.splitext(file)[0]
    folder = os.path.join(os.path.dirname(stem), 'cache')
    name = '%s_%03d.npz' % (os.path.basename(stem), rng.randint(0, num_aug))
    if not os.path.isdir(folder):
        os.makedirs(folder, exist_ok=True)
    return os.path.join(folder, name)


def fill(file, opt):
    """fill_mesh (mesh_prepare.py:6-27): the cached preparation of `file` for a random augmentation slot, built and stored
    on a miss.  opt=None means "no augmentation, no cache" (a convenience the reference does not have)."""
    import os
    if opt is None:
        return from_obj(file)
    path = cache_path(file, opt.num_aug)
    if os.path.exists(path):
        z = np.load(path, encoding='latin1', allow_pickle=True)
        d = MeshData()
        for k in CACHE_KEYS:
            setattr(d, k, z[k])
        d.edges_count = int(d.edges_count)
        d.filename = str(d.filename)
        d.ve = [list(map(int, l)) for l in d.ve]
        return d
    d = from_obj(file, opt)
    ve = np.empty(len(d.ve), dtype=object)
    for i, l in enumerate(d.ve):
        ve[i] = list(l)
    np.savez_compressed(path, gemm_edges=d.gemm_edges, vs=d.vs, edges=d.edges, edges_count=d.edges_count, ve=ve, v_mask=d.v_mask,
                        filename=d.filename, sides=d.sides, edge_lengths=d.edge_lengths, edge_areas=d.edge_areas,
                        features=d.features)
    return d
