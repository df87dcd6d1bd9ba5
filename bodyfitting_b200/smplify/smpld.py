"""SMPL+D: per-vertex displacement fit of the SMPLify result to a scan
(reference: smplify/smplify.py:228-247 with smplify/loss.py:233-288 and utils/io_utils.py:405-428),
on the B200 grid kernels (include/bodyfit_b200_grid.h: bf_smpld_step / bf_smpld_run)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .. import constants as K
from ..engine import _stream


def vertex_face_csr(faces, num_verts):
    """CSR vertex -> incident faces (ascending face id), the fixed order of every per-vertex sum."""
    faces = np.asarray(faces, dtype=np.int64)
    v = faces.reshape(-1)
    f = np.repeat(np.arange(len(faces)), 3)
    o = np.lexsort((f, v))
    v, f = v[o], f[o]
    keep = np.ones(len(v), bool)
    keep[1:] = (v[1:] != v[:-1]) | (f[1:] != f[:-1])          # a degenerate face lists a vertex once
    v, f = v[keep], f[keep]
    ptr = np.zeros(num_verts + 1, dtype=np.int32)
    np.add.at(ptr, v + 1, 1)
    return np.cumsum(ptr).astype(np.int32), f.astype(np.int32)


class DisplacementFitter(object):
    def __init__(self, searcher, scan_face_normals, body_faces, num_verts, constant_scale, device='cuda'):
        self.searcher = searcher
        self.dev = torch.device(device)
        self.V, self.F = int(num_verts), int(len(body_faces))
        self.constant_scale = float(constant_scale)
        ptr, fid = vertex_face_csr(body_faces, num_verts)
        T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(self.dev)
        self.t = dict(faces=T(np.asarray(body_faces), torch.int32), vf_ptr=T(ptr, torch.int32), vf_face=T(fid, torch.int32),
                      scan_fn=T(np.asarray(scan_face_normals, dtype=np.float32), torch.float32))

    def run(self, body_vertices, num_iters, lr=5e-2, return_grad=False):
        """body_vertices [V,3] (world, detached) -> disp [V,3], trace [num_iters,4] = icp, normal, smooth, loss."""
        V, F, dev = self.V, self.F, self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        t = self.t
        b = dict(base=body_vertices.detach().to(**f32).reshape(V, 3).contiguous(), disp=torch.zeros(V, 3, **f32),
                 adam_m=torch.zeros(V, 3, **f32), adam_v=torch.zeros(V, 3, **f32), P=torch.empty(V, 3, **f32),
                 C=torch.empty(V, 3, **f32), near_faces=torch.empty(V, dtype=torch.int32, device=dev),
                 nhat=torch.empty(F, 3, **f32), nlen=torch.empty(F, **f32), m=torch.empty(V, 3, **f32),
                 Nlen=torch.empty(V, **f32), dN=torch.empty(V, 3, **f32), dcorner=torch.empty(F, 9, **f32),
                 partial=torch.empty(64, 3, **f32), totals=torch.empty(4, **f32),
                 trace=torch.zeros(max(num_iters, 1), 4, **f32))
        if return_grad:
            b['grad'] = torch.empty(V, 3, **f32)
        s = _lib.BfSmpld()
        for k, v in list(b.items()) + list(t.items()):
            setattr(s, k, v.data_ptr())
        s.lr, s.beta1, s.beta2, s.eps = float(lr), K.ADAM_BETAS[0], K.ADAM_BETAS[1], K.ADAM_EPS
        s.reg_scale = self.constant_scale * 0.1
        s.V, s.F, s.iter = V, F, 0
        _lib.check(_lib.lib().bf_smpld_run(C.byref(self.searcher.grid), C.byref(s), int(num_iters), _stream()), 'bf_smpld_run')
        self.buffers = b
        return b['disp'], b['trace']
