"""Synthetic manifold meshes for tests and benchmarks.

Closed, boundary-free triangle meshes with exactly predictable edge counts:

* ``icosphere(nu)``  : class-I geodesic sphere, V = 10 nu^2 + 2, F = 20 nu^2, E = 30 nu^2
                        (nu = 5  -> 750 edges, the SHREC / cubes resolution)
* ``torus(m, n)``    : m x n quad grid split into triangles, V = mn, F = 2mn, E = 3mn
                        (20 x 38 -> 2280 edges, the human-seg / COSEG resolution)
* ``open_grid(m, n)``: a planar patch with a boundary (exercises the ``-1`` neighbour path)

``perturb`` applies the per-mesh deterministic vertex noise named in BASELINE.md §4.
Nothing here reads the reference; the generators exist on the GPU box as well.
"""
import numpy as np


def _icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    vs = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
                   [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                   [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    faces = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                      [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                      [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                      [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    vs /= np.linalg.norm(vs, axis=1, keepdims=True)
    return vs, faces


def icosphere(nu):
    """Geodesic sphere of frequency ``nu``: every icosahedron face becomes nu^2 triangles."""
    base_vs, base_faces = _icosahedron()
    key2id = {}
    out_vs = []

    def vid(a, b, c, i, j, k):
        # barycentric lattice point (i, j, k)/nu of triangle (a, b, c); canonical key so that
        # points on shared icosahedron edges / corners are emitted once
        w = sorted(((a, i), (b, j), (c, k)))
        key = tuple((v, n) for v, n in w if n != 0)
        got = key2id.get(key)
        if got is None:
            p = (base_vs[a] * i + base_vs[b] * j + base_vs[c] * k) / nu
            got = len(out_vs)
            key2id[key] = got
            out_vs.append(p / np.linalg.norm(p))
        return got

    faces = []
    for a, b, c in base_faces:
        for i in range(nu):
            for j in range(nu - i):
                k = nu - i - j
                # "up" triangle
                faces.append([vid(a, b, c, k, j, i), vid(a, b, c, k - 1, j + 1, i), vid(a, b, c, k - 1, j, i + 1)])
                if j + i < nu - 1:
                    # "down" triangle
                    faces.append([vid(a, b, c, k - 1, j + 1, i), vid(a, b, c, k - 2, j + 1, i + 1),
                                  vid(a, b, c, k - 1, j, i + 1)])
    return np.asarray(out_vs, dtype=np.float64), np.asarray(faces, dtype=np.int64)


def torus(m, n, R=1.0, r=0.4):
    """m x n torus, E = 3 m n."""
    u = np.arange(m) * (2 * np.pi / m)
    v = np.arange(n) * (2 * np.pi / n)
    uu, vv = np.meshgrid(u, v, indexing='ij')
    vs = np.stack([(R + r * np.cos(vv)) * np.cos(uu), (R + r * np.cos(vv)) * np.sin(uu), r * np.sin(vv)], -1)
    vs = vs.reshape(-1, 3)
    ii, jj = np.meshgrid(np.arange(m), np.arange(n), indexing='ij')
    a = (ii * n + jj).ravel()
    b = (((ii + 1) % m) * n + jj).ravel()
    c = (((ii + 1) % m) * n + (jj + 1) % n).ravel()
    d = (ii * n + (jj + 1) % n).ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    # interleave so the two triangles of a quad are adjacent in file order
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return vs.astype(np.float64), faces.astype(np.int64)


def open_grid(m, n):
    """Planar (m+1) x (n+1) vertex patch with a boundary; E = 3mn + m + n."""
    ii, jj = np.meshgrid(np.arange(m + 1), np.arange(n + 1), indexing='ij')
    vs = np.stack([ii.ravel() / m, jj.ravel() / n, 0.15 * np.sin(3.0 * ii.ravel() / m) * np.cos(2.0 * jj.ravel() / n)], -1)
    qi, qj = np.meshgrid(np.arange(m), np.arange(n), indexing='ij')
    a = (qi * (n + 1) + qj).ravel()
    b = ((qi + 1) * (n + 1) + qj).ravel()
    c = ((qi + 1) * (n + 1) + qj + 1).ravel()
    d = (qi * (n + 1) + qj + 1).ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    faces = faces.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return vs.astype(np.float64), faces.astype(np.int64)


def perturb(vs, seed, a=0.1):
    """v <- v (1 + a n1) + 0.01 n3^3 with RandomState(seed)."""
    rs = np.random.RandomState(seed)
    n1 = rs.randn(len(vs), 1)
    n3 = rs.randn(len(vs), 3)
    return vs * (1.0 + a * np.clip(n1, -2.5, 2.5)) + 0.01 * n3 ** 3


def write_obj(path, vs, faces):
    with open(path, 'w') as f:
        for v in vs:
            f.write('v %.9f %.9f %.9f\n' % (v[0], v[1], v[2]))
        for t in faces:
            f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))


def make(kind, seed, **kw):
    """One perturbed synthetic mesh: returns (vs, faces)."""
    if kind == 'ico':
        vs, faces = icosphere(kw.get('nu', 5))
        return perturb(vs, seed, kw.get('a', 0.1)), faces
    if kind == 'torus':
        vs, faces = torus(kw.get('m', 20), kw.get('n', 38))
        return perturb(vs, seed, kw.get('a', 0.03)), faces
    if kind == 'grid':
        vs, faces = open_grid(kw.get('m', 10), kw.get('n', 12))
        return perturb(vs, seed, kw.get('a', 0.01)), faces
    raise ValueError(kind)


def write_classification_dataset(root, classes=2, n_train=4, n_test=2, nu=3, seed=0):
    """A tiny dataset in the layout the reference's ClassificationData expects (data/classification_data.py:47-60):
    root/<class>/{train,test}/*.obj.  Class k is an icosphere stretched along axis k."""
    import os
    s = seed
    for k in range(classes):
        for split, n in (('train', n_train), ('test', n_test)):
            d = os.path.join(root, 'class%d' % k, split)
            os.makedirs(d, exist_ok=True)
            for i in range(n):
                vs, faces = make('ico', s, nu=nu)
                vs = vs.copy()
                vs[:, k % 3] *= 1.8
                write_obj(os.path.join(d, 'm%d_%d.obj' % (k, i)), vs, faces)
                s += 1
    return root


def write_segmentation_dataset(root, n_train=4, n_test=2, nu=3, nclasses=3, seed=0):
    """Segmentation layout (data/segmentation_data.py:15-21): root/{train,test}/*.obj, root/seg/<name>.eseg (one label per
    edge, 1-based), root/sseg/<name>.seseg (soft labels [E, nclasses]), root/classes.txt."""
    import os
    from . import mesh_prepare
    os.makedirs(os.path.join(root, 'seg'), exist_ok=True)
    os.makedirs(os.path.join(root, 'sseg'), exist_ok=True)
    s = seed
    for split, n in (('train', n_train), ('test', n_test)):
        d = os.path.join(root, split)
        os.makedirs(d, exist_ok=True)
        for i in range(n):
            vs, faces = make('ico', s, nu=nu)
            name = '%s_%d' % (split, i)
            write_obj(os.path.join(d, name + '.obj'), vs, faces)
            m = mesh_prepare.from_obj(os.path.join(d, name + '.obj'))
            mid = vs[m.edges].mean(1)
            lab = np.minimum((np.clip((mid[:, 2] + 1.0) / 2.0, 0, 0.999) * nclasses).astype(np.int64), nclasses - 1)
            np.savetxt(os.path.join(root, 'seg', name + '.eseg'), lab + 1, fmt='%d')
            soft = np.zeros((len(lab), nclasses))
            soft[np.arange(len(lab)), lab] = 1
            np.savetxt(os.path.join(root, 'sseg', name + '.seseg'), soft, fmt='%f')
            s += 1
    with open(os.path.join(root, 'classes.txt'), 'w') as f:
        f.write('\n'.join(str(i + 1) for i in range(nclasses)) + '\n')
    return root
