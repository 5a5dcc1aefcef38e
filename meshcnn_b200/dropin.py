"""Run the reference's own train.py / test.py on the CUDA path, unchanged.

    python -m meshcnn_b200.dropin /path/to/MeshCNN/train.py <the reference's usual flags>

The reference tree is put on sys.path, the CUDA-backed layer modules are registered under the reference's dotted
module names *before* models.networks / data.* import them, then the script runs as __main__.
"""
import importlib
import os
import runpy
import sys

_NAMES = ('mesh_conv', 'mesh_pool', 'mesh_unpool', 'mesh_union', 'mesh')


def install(reference_root):
    """Make `from models.layers.mesh_conv import MeshConv` (etc.) resolve to meshcnn_b200's classes."""
    reference_root = os.path.abspath(reference_root)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    importlib.import_module('models.layers')                 # the reference's (empty) package objects
    for n in _NAMES:
        mod = importlib.import_module('meshcnn_b200.' + n)
        sys.modules['models.layers.' + n] = mod
        setattr(sys.modules['models.layers'], n, mod)
    return [sys.modules['models.layers.' + n] for n in _NAMES]


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit('usage: python -m meshcnn_b200.dropin <reference>/train.py [reference flags]')
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
