"""Helpers shared by the -m gpu parity tests (product objects built from fixtures)."""
import numpy as np
import torch

from meshcnn_b200 import mesh_prepare
from meshcnn_b200.mesh import Mesh

DEV = 'cuda:0'


def product_mesh(vs, faces, hold_history=False, name='golden'):
    return Mesh(data=mesh_prepare.from_arrays(vs, faces, filename=name), hold_history=hold_history)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
