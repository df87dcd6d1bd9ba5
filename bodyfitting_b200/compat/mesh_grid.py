"""Drop-in for the reference's native module ``mesh_grid`` (pybind11, thirdparty/mesh_grid/mesh_grid.cpp:129-136): the same
six functions with the same tensor signatures, ownership and in-place output conventions, on the B200 grid kernels of
``libbodyfit_b200.so`` (include/bodyfit_b200_grid.h).  The reference's own wrapper ``utils/mesh_grid_searcher.py`` binds to
it unmodified:

    import sys, bodyfitting_b200.compat.mesh_grid as mg
    sys.modules['mesh_grid'] = mg            # before `import utils.mesh_grid_searcher`

Conventions kept (mesh_grid.cpp:39-127, mesh_grid_kernel.cu:181-237,385-433): the caller allocates every tensor; outputs are
``resize_``d and overwritten in place; ``tri_num`` [cells] becomes the inclusive cumulative triangle count per cell and
``tri_idx`` is resized to the total and filled with the cell lists (here: 0-based face ids, ascending inside a cell -- the
reference stores id+1 in arrival order; both are only ever read back by this module); ``step`` arrives as a float (a 0-dim
tensor is converted, one device->host read as in the reference).  Differences by design: launches go to the CURRENT stream
(the reference uses the legacy default stream) and errors raise instead of being printed (mesh_grid_kernel.cu:210-212).
"""
import ctypes as C

import torch

from .. import _lib

__all__ = ['insert_grid_surface', 'search_nearest_point', 'search_inside_mesh', 'search_intersect', 'cumsum',
           'search_nearest_point_backward']


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(**tensors):
    for name, t in tensors.items():
        if not (torch.is_tensor(t) and t.is_cuda):
            raise RuntimeError('%s must be a CUDA tensor' % name)      # mesh_grid.cpp:35 CHECK_CUDA


def _grid(verts, faces, tri_num, tri_idx, num, minmax, step, build=False):
    _lib.require_device()
    if verts.dtype != torch.float32 or faces.dtype != torch.int32:
        raise RuntimeError('mesh_grid: verts must be float32 and faces int32')
    verts_c, faces_c = verts.reshape(-1, 3), faces.reshape(-1, 3)
    if not (verts_c.is_contiguous() and faces_c.is_contiguous()):
        raise RuntimeError('mesh_grid: verts / faces must be contiguous')
    # grid geometry lives in device tensors in the reference's interface; one small device->host read per call (the reference
    # pays the same for `step`, a 0-dim CUDA tensor converted to float by pybind: mesh_grid_searcher.py:65,76-79)
    nums, mm = num.detach().cpu().tolist(), minmax.detach().cpu().tolist()
    g = _lib.BfGrid()
    g.verts, g.faces = verts_c.data_ptr(), faces_c.data_ptr()
    for d in range(3):
        g.min[d] = float(mm[d])
        g.dim[d] = int(nums[d])
    g.step = float(step)
    g.ncell, g.Ns, g.Fs = int(nums[3]), verts_c.shape[0], faces_c.shape[0]
    keep = None
    if not build:
        # the kernels index [cells + 1] exclusive offsets; tri_num holds the inclusive counts (reference layout)
        keep = torch.cat([torch.zeros(1, dtype=torch.int32, device=tri_num.device), tri_num.reshape(-1)])
        g.cell_start, g.cell_tris = keep.data_ptr(), tri_idx.data_ptr()
    return g, verts_c, keep


def insert_grid_surface(verts, faces, minmax, num, step, tri_num, tri_idx):
    """mesh_grid.cpp:39-52 -> mesh_grid_kernel.cu:181-237: count per cell, scan, fill (deterministic: per-cell lists sorted)."""
    _check_cuda(verts=verts, faces=faces, minmax=minmax, num=num, tri_num=tri_num, tri_idx=tri_idx)
    with torch.cuda.device(verts.device):
        g, _, _ = _grid(verts, faces, tri_num, tri_idx, num, minmax, step, build=True)
        L = _lib.lib()
        start = torch.zeros(g.ncell + 1, dtype=torch.int32, device=verts.device)
        scratch = torch.zeros(g.ncell, dtype=torch.int32, device=verts.device)
        g.cell_start = start.data_ptr()
        _lib.check(L.bf_grid_count(C.byref(g), scratch.data_ptr(), _stream()), 'bf_grid_count')
        total = int(start[-1].item())                      # one host read per mesh, as the reference's cumsum + item (:214-219)
        tri_idx.resize_(max(total, 1))
        g.cell_tris = tri_idx.data_ptr()
        _lib.check(L.bf_grid_fill(C.byref(g), scratch.data_ptr(), _stream()), 'bf_grid_fill')
        tri_num.resize_(g.ncell)
        tri_num.copy_(start[1:])


def search_nearest_point(points, verts, faces, tri_num, tri_idx, num, minmax, step, near_faces, near_pts, coeff):
    """mesh_grid.cpp:54-72 -> mesh_grid_kernel.cu:385-433: closest point, its face and its barycentric coefficients."""
    _check_cuda(points=points, verts=verts, faces=faces, tri_num=tri_num, tri_idx=tri_idx, num=num, minmax=minmax,
                near_faces=near_faces, coeff=coeff)
    with torch.cuda.device(verts.device):
        g, _, offsets = _grid(verts, faces, tri_num, tri_idx, num, minmax, step)      # `offsets` stays referenced until the launches below are queued
        pts = points.detach().reshape(-1, 3).float().contiguous()
        Q = pts.shape[0]
        near_faces.resize_(Q)
        near_pts.resize_(Q, 3)
        coeff.resize_(Q, 3)
        if Q == 0:
            return
        L = _lib.lib()
        _lib.check(L.bf_grid_nearest(C.byref(g), pts.data_ptr(), Q, near_pts.data_ptr(), near_faces.data_ptr(), None, _stream()),
                   'bf_grid_nearest')
        _lib.check(L.bf_grid_barycentric(C.byref(g), pts.data_ptr(), near_faces.data_ptr(), Q, coeff.data_ptr(), _stream()),
                   'bf_grid_barycentric')


def search_inside_mesh(points, verts, faces, tri_num, tri_idx, num, minmax, step, signs):
    """mesh_grid.cpp:74-90: signs[Q] = +1 inside the closed mesh, -1 outside."""
    _check_cuda(points=points, verts=verts, faces=faces, tri_num=tri_num, tri_idx=tri_idx, num=num, minmax=minmax, signs=signs)
    with torch.cuda.device(verts.device):
        g, _, offsets = _grid(verts, faces, tri_num, tri_idx, num, minmax, step)      # `offsets` stays referenced until the launches below are queued
        pts = points.detach().reshape(-1, 3).float().contiguous()
        Q = pts.shape[0]
        signs.resize_(Q)
        if Q:
            _lib.check(_lib.lib().bf_grid_inside(C.byref(g), pts.data_ptr(), Q, signs.data_ptr(), _stream()), 'bf_grid_inside')


def search_intersect(origins, directions, verts, faces, tri_num, tri_idx, num, minmax, step, intersect):
    """mesh_grid.cpp:92-109: intersect[Q] (bool) = the ray origin + t direction, t >= 0, meets the mesh."""
    _check_cuda(origins=origins, directions=directions, verts=verts, faces=faces, tri_num=tri_num, tri_idx=tri_idx, num=num,
                minmax=minmax, intersect=intersect)
    with torch.cuda.device(verts.device):
        g, _, offsets = _grid(verts, faces, tri_num, tri_idx, num, minmax, step)      # `offsets` stays referenced until the launches below are queued
        o = origins.detach().reshape(-1, 3).float().contiguous()
        d = directions.detach().reshape(-1, 3).float().contiguous()
        Q = o.shape[0]
        intersect.resize_(Q)
        if Q:
            hit = intersect if intersect.dtype == torch.uint8 else torch.empty(Q, dtype=torch.uint8, device=o.device)
            _lib.check(_lib.lib().bf_grid_intersects_any(C.byref(g), o.data_ptr(), d.data_ptr(), Q, hit.data_ptr(), _stream()),
                       'bf_grid_intersects_any')
            if hit is not intersect:
                intersect.copy_(hit)


def cumsum(input):
    """mesh_grid.cpp:111-118: in-place cumulative sum along dim 0, returned reshaped to [1,1,-1]."""
    input.copy_(input.cumsum(0))             # same storage, same dtype (an int32 cumsum comes back as int64 from torch)
    return input.reshape(1, 1, -1)


def search_nearest_point_backward(points, verts, faces, near_faces, grad):
    """mesh_grid.cpp:119-127: grad is resized to [Q,3,3,3]; grad[q,i,j,k] = d near_pt[q][j] / d verts[faces[near_faces[q]][i]][k]
    (exact inside each Voronoi region).  The reference's kernel of this name is unfinished and never called
    (mesh_grid_kernel.cu:354-382, utils/mesh_grid_searcher.py:17-49)."""
    _check_cuda(points=points, verts=verts, faces=faces, near_faces=near_faces, grad=grad)
    with torch.cuda.device(verts.device):
        _lib.require_device()
        g = _lib.BfGrid()
        verts_c, faces_c = verts.reshape(-1, 3).contiguous(), faces.reshape(-1, 3).contiguous()
        g.verts, g.faces = verts_c.data_ptr(), faces_c.data_ptr()
        g.Ns, g.Fs = verts_c.shape[0], faces_c.shape[0]
        pts = points.detach().reshape(-1, 3).float().contiguous()
        Q = pts.shape[0]
        grad.resize_(Q, 3, 3, 3)
        if Q:
            _lib.check(_lib.lib().bf_grid_nearest_backward(C.byref(g), pts.data_ptr(), near_faces.data_ptr(), Q, grad.data_ptr(),
                                                           _stream()), 'bf_grid_nearest_backward')
