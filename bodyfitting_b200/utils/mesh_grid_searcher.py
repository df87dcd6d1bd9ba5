"""``MeshGridSearcher`` -- same interface as the reference's ``utils/mesh_grid_searcher.py:51-84``
(``set_mesh``, ``nearest_points`` -> (points [Q,3] f32, face ids [Q] i32), ``inside_mesh`` -> signs [Q] f32,
``intersects_any`` -> bool [Q]; :51-99) on the B200 grid kernels (include/bodyfit_b200_grid.h)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..engine import _stream


class MeshGridSearcher(object):
    def __init__(self, verts=None, faces=None, device=None):
        """``device``: where the grid lives; default = the device of ``verts`` if it is a CUDA tensor, else the CURRENT CUDA
        device (the reference hard-codes 'cuda:0', utils/mesh_grid_searcher.py:52,56 -- wrong on every rank but the first)."""
        self.device = None if device is None else torch.device(device)
        if verts is not None and faces is not None:
            self.set_mesh(verts, faces, device)

    def set_mesh(self, verts, faces, device=None):
        _lib.require_device()
        if device is None:
            device = verts.device if (torch.is_tensor(verts) and verts.is_cuda) else torch.device('cuda', torch.cuda.current_device())
        dev = torch.device(device)
        self.device = dev
        verts = torch.as_tensor(np.asarray(verts) if not torch.is_tensor(verts) else verts).float().to(dev).reshape(-1, 3).contiguous()
        faces = torch.as_tensor(np.asarray(faces) if not torch.is_tensor(faces) else faces).to(torch.int32).to(dev).reshape(-1, 3).contiguous()
        self.verts, self.faces = verts, faces
        # grid sizing exactly as the reference (mesh_grid_searcher.py:63-71), fp32 on the host
        v = verts.detach().cpu().numpy()
        _min, _max = v.min(0), v.max(0)
        ext = (_max - _min).astype(np.float32)
        step = np.float32((np.float32(np.prod(ext, dtype=np.float32)) / np.float32(len(v))) ** np.float32(1.0 / 3.0))
        l = np.maximum(np.floor(ext / step), 0) + 1
        c = (_max + _min) / np.float32(2)
        min_step = (c - step * l / np.float32(2)).astype(np.float32)
        self.step = float(step)
        self.num = [int(x) for x in l] + [int(np.prod(l))]
        self.minmax = np.concatenate([min_step, _max])
        g = _lib.BfGrid()
        g.verts, g.faces = verts.data_ptr(), faces.data_ptr()
        for d in range(3):
            g.min[d] = float(min_step[d])
            g.dim[d] = int(l[d])
        g.step = self.step
        g.ncell, g.Ns, g.Fs = self.num[3], verts.shape[0], faces.shape[0]
        self.cell_start = torch.zeros(g.ncell + 1, dtype=torch.int32, device=dev)
        scratch = torch.zeros(g.ncell, dtype=torch.int32, device=dev)
        g.cell_start = self.cell_start.data_ptr()
        L = _lib.lib()
        _lib.check(L.bf_grid_count(C.byref(g), scratch.data_ptr(), _stream()), 'bf_grid_count')
        total = int(self.cell_start[-1].item())                 # one host sync per mesh, as in the reference (:217)
        self.cell_tris = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
        g.cell_tris = self.cell_tris.data_ptr()
        _lib.check(L.bf_grid_fill(C.byref(g), scratch.data_ptr(), _stream()), 'bf_grid_fill')
        self.grid = g
        self.tri_num = self.cell_start[1:]                       # the reference's cumulative counts
        self.tri_idx = self.cell_tris

    def nearest_points(self, points, return_dist2=False):
        points = points.to(self.device).float().reshape(-1, 3).contiguous()
        Q = points.shape[0]
        near_pts = torch.empty(Q, 3, device=self.device)
        near_faces = torch.empty(Q, dtype=torch.int32, device=self.device)
        d2 = torch.empty(Q, device=self.device) if return_dist2 else None
        _lib.check(_lib.lib().bf_grid_nearest(C.byref(self.grid), points.detach().data_ptr(), Q, near_pts.data_ptr(),
                                              near_faces.data_ptr(), d2.data_ptr() if return_dist2 else None, _stream()),
                   'bf_grid_nearest')
        return (near_pts, near_faces, d2) if return_dist2 else (near_pts, near_faces)

    def inside_mesh(self, points):
        """+1 for query points inside the closed mesh, -1 outside (reference: mesh_grid_searcher.py:86-91)."""
        points = points.to(self.device).float().reshape(-1, 3).contiguous()
        Q = points.shape[0]
        signs = torch.empty(Q, dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().bf_grid_inside(C.byref(self.grid), points.detach().data_ptr(), Q, signs.data_ptr(), _stream()),
                   'bf_grid_inside')
        return signs

    def intersects_any(self, origins, directions):
        """bool [Q]: does the ray origin + t * direction (t >= 0) meet the mesh (reference: mesh_grid_searcher.py:93-99)."""
        origins = origins.to(self.device).float().reshape(-1, 3).contiguous()
        directions = directions.to(self.device).float().reshape(-1, 3).contiguous()
        assert origins.shape == directions.shape
        Q = origins.shape[0]
        hit = torch.empty(Q, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().bf_grid_intersects_any(C.byref(self.grid), origins.detach().data_ptr(), directions.detach().data_ptr(),
                                                     Q, hit.data_ptr(), _stream()), 'bf_grid_intersects_any')
        return hit.bool()
