"""Generate tests/golden/prepare_*.npz: the UNMODIFIED reference's mesh_prepare.from_scratch (features, edge areas / lengths,
topology; with and without the data augmentation of scripts/*/train.sh) on synthetic meshes, under a seeded global np.random.
Build container only:   python tools/make_golden_prepare.py
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from meshcnn_b200 import mesh_prepare, synth           # noqa: E402
from oracle import reference_arm                        # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
CASES = [
    # name, generator, kwargs, vertex scale (the reference's normal / angle features divide by |.| + 0.1: meshes must be
    # large against 0.1 for the flat-edge tests of flip_edges / slide_verts to fire), perturbation seed or None
    ('ico3', 'ico', dict(nu=3), 50.0, None),
    ('ico5p', 'ico', dict(nu=5), 1.0, 11),
    ('grid', 'grid', dict(m=10, n=12), 50.0, None),
    ('torus', 'torus', dict(m=10, n=25), 30.0, 5),
]
OPTS = [
    ('plain', dict(num_aug=1)),
    ('aug', dict(num_aug=20, scale_verts=True, flip_edges=0.2, slide_verts=0.2)),          # scripts/shrec/train.sh + scale
    ('flip', dict(num_aug=5, scale_verts=False, flip_edges=0.5, slide_verts=0.0)),
    ('slide', dict(num_aug=5, slide_verts=0.5)),
]


def ragged(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(l) for l in lists])
    return np.array([x for l in lists for x in l], dtype=np.int32), off


def main():
    ref = reference_arm.load()
    tmp = tempfile.mkdtemp(prefix='golden_prep_')
    for name, kind, kw, scale, pseed in CASES:
        if pseed is None:
            vs, faces = {'ico': lambda: synth.icosphere(kw['nu']), 'grid': lambda: synth.open_grid(kw['m'], kw['n']),
                         'torus': lambda: synth.torus(kw['m'], kw['n'])}[kind]()
        else:
            vs, faces = synth.make(kind, pseed, **kw)
        p = os.path.join(tmp, name + '.obj')
        synth.write_obj(p, vs * scale, faces)
        vs_in = mesh_prepare.read_obj(p)[0]
        for oname, okw in OPTS:
            seed = 1234 + len(name) + len(oname)
            np.random.seed(seed)
            d = ref.mesh_prepare.from_scratch(p, types.SimpleNamespace(**okw))
            draws_left = np.random.randint(0, 1 << 30)               # fingerprint of the generator position afterwards
            out = {'vs_in': vs_in, 'faces_in': faces.astype(np.int32), 'seed': np.int64(seed), 'next_draw': np.int64(draws_left),
                   'opt_keys': np.array(sorted(okw)), 'opt_vals': np.array([float(okw[k]) for k in sorted(okw)]),
                   'vs': np.asarray(d.vs, dtype=np.float64), 'edges': np.asarray(d.edges, dtype=np.int32),
                   'gemm': np.asarray(d.gemm_edges, dtype=np.int32), 'sides': np.asarray(d.sides, dtype=np.int8),
                   'edge_areas': np.asarray(d.edge_areas), 'edge_lengths': np.asarray(d.edge_lengths),
                   'features': np.asarray(d.features), 'ec': np.int32(d.edges_count),
                   'shifted': np.float64(getattr(d, 'shifted', -1.0))}
            out['ve_flat'], out['ve_off'] = ragged(d.ve)
            np.savez_compressed(os.path.join(OUT, 'prepare_%s_%s.npz' % (name, oname)), **out)
            print(name, oname, 'E', d.edges_count, 'shifted', out['shifted'])


if __name__ == '__main__':
    main()
