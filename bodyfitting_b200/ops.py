"""torch.autograd bindings of the stand-alone CUDA operators (include/bodyfit_b200_ops.h) and of the
all-vertex LBS operator.  Torch carries tensors and the autograd graph; the arithmetic is in the kernels."""
import numpy as np
import torch

from . import _lib
from .engine import FrameBuffers, _stream


def _f32c(t):
    return t.detach().to(dtype=torch.float32).contiguous()


def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args, _stream()), name)


def _need_cuda(*ts):
    _lib.require_device()
    for t in ts:
        if not t.is_cuda:
            raise _lib.BodyfitError('bodyfitting_b200 operators need CUDA tensors (no CPU path)')


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, rotation, translation, K):
        _need_cuda(points)
        p, R, t, Km = _f32c(points), _f32c(rotation), _f32c(translation), _f32c(K)
        B, N = p.shape[0], p.shape[1]
        nb = R.shape[0]
        uv = torch.empty(B, N, 2, device=p.device)
        _call('bf_op_project', p.data_ptr(), R.data_ptr(), t.data_ptr(), Km.data_ptr(), uv.data_ptr(), B, N, nb)
        ctx.save_for_backward(p, R, t, Km)
        return uv

    @staticmethod
    def backward(ctx, duv):
        p, R, t, Km = ctx.saved_tensors
        B, N = p.shape[0], p.shape[1]
        dp = torch.empty_like(p)
        _call('bf_op_project_backward', p.data_ptr(), R.data_ptr(), t.data_ptr(), Km.data_ptr(), _f32c(duv).data_ptr(),
              dp.data_ptr(), B, N, R.shape[0])
        return dp, None, None, None


class _Gmof(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sigma):
        _need_cuda(x)
        xc = _f32c(x)
        y = torch.empty_like(xc)
        _call('bf_op_gmof', xc.data_ptr(), y.data_ptr(), float(sigma), xc.numel())
        ctx.save_for_backward(xc)
        ctx.sigma = float(sigma)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        dx = torch.empty_like(xc)
        _call('bf_op_gmof_backward', xc.data_ptr(), _f32c(dy).data_ptr(), dx.data_ptr(), ctx.sigma, xc.numel())
        return dx, None


class _Reprojection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cord, cord_gt, weights, scale_coeff, sigma):
        _need_cuda(cord)
        c, g, w = _f32c(cord), _f32c(cord_gt), _f32c(weights)
        out = torch.empty(1, device=c.device)
        dc = torch.empty_like(c)
        _call('bf_op_reprojection', c.data_ptr(), g.data_ptr(), w.data_ptr(), float(scale_coeff), float(sigma),
              c.shape[0], out.data_ptr(), dc.data_ptr())
        ctx.save_for_backward(dc)
        return out[0]

    @staticmethod
    def backward(ctx, dout):
        (dc,) = ctx.saved_tensors
        return dc * dout, None, None, None, None


class _KeypointsWorld(torch.autograd.Function):
    """per-frame data term [B] on world joints [B,K,3]"""
    @staticmethod
    def forward(ctx, joints, kp_packed, cams, scale_coeff, sigma):
        _need_cuda(joints)
        j = _f32c(joints)
        B, K = j.shape[0], j.shape[1]
        Nv = kp_packed.shape[2]                       # [B,K,Nv,3] joint-major (engine.pack_keypoints)
        lbk = torch.empty(B, K, device=j.device)
        dJ = torch.empty_like(j)
        _call('bf_op_keypoints_world', j.data_ptr(), kp_packed.data_ptr(), cams.data_ptr(), B, K, Nv, float(scale_coeff),
              float(sigma), lbk.data_ptr(), dJ.data_ptr())
        ctx.save_for_backward(dJ)
        return lbk.sum(dim=1)

    @staticmethod
    def backward(ctx, dl):
        (dJ,) = ctx.saved_tensors
        return dJ * dl.view(-1, 1, 1), None, None, None, None


class _AnglePrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose):
        _need_cuda(pose)
        p = _f32c(pose)
        B, D = p.shape
        out = torch.empty(B, 4, device=p.device)
        d = torch.empty(B, 4, device=p.device)
        _call('bf_op_angle_prior', p.data_ptr(), B, D, out.data_ptr(), d.data_ptr())
        ctx.save_for_backward(d)
        ctx.D = D
        return out

    @staticmethod
    def backward(ctx, dout):
        (d,) = ctx.saved_tensors
        g = torch.zeros(d.shape[0], ctx.D, device=d.device)
        g[:, [52, 55, 9, 12]] = d * dout
        return g


class _GmmPose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, prepared):
        _need_cuda(pose)
        p = _f32c(pose)
        B, D = p.shape
        grad = torch.empty(B, 69, device=p.device)
        loss = torch.empty(B, device=p.device)
        _lib.check(_lib.lib().bf_op_gmm_pose(prepared.struct, p.data_ptr(), D, min(D, 69), B, 1.0, grad.data_ptr(),
                                             loss.data_ptr(), _stream()), 'bf_op_gmm_pose')
        ctx.save_for_backward(grad)
        ctx.D = D
        return loss

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        g = grad[:, :ctx.D] * dl.view(-1, 1)
        return g, None


class _Lbs(torch.autograd.Function):
    """theta [B,NP] -> vertices [B,V,3], joints [B,K_full,3], full_pose [B,3J] on the all-vertex set."""
    @staticmethod
    def forward(ctx, theta, prepared):
        _need_cuda(theta)
        B = theta.shape[0]
        fb = FrameBuffers(prepared, B, full=True, need_backward=True)
        fb.t['theta'].copy_(theta.detach())
        fb.call('bf_lbs_forward')
        ctx.fb = fb
        ctx.mark_non_differentiable(fb.t['full_pose'])
        return fb.t['verts'].view(B, prepared.V, 3), fb.t['joints'], fb.t['full_pose']

    @staticmethod
    def backward(ctx, dverts, djoints, _dfp):
        fb = ctx.fb
        B = fb.B
        if dverts is None:
            fb.t['dverts'].zero_()
        else:
            fb.t['dverts'].copy_(dverts.reshape(B, -1))
        fb.bind('djoints', _f32c(djoints) if djoints is not None else torch.zeros_like(fb.t['joints']))
        fb.call('bf_lbs_backward')
        return fb.t['grad'].clone(), None


_csr_cache = {}


def _faces_csr(faces):
    """int32 faces [F,3] on the device + the CSR vertex -> incident faces the atomics-free gathers walk (built on the host
    once per faces tensor)."""
    from .smplify.smpld import vertex_face_csr
    f = faces.detach().reshape(-1, 3)
    key = (f.data_ptr(), f.shape[0], str(f.device), f._version)
    hit = _csr_cache.get(key)
    if hit is None:
        fh = f.cpu().numpy().astype(np.int64)
        V = int(fh.max()) + 1
        ptr, fid = vertex_face_csr(fh, V)
        hit = (f.to(torch.int32).contiguous(), torch.from_numpy(ptr).to(f.device), torch.from_numpy(fid).to(f.device), V)
        if len(_csr_cache) > 16:
            _csr_cache.clear()
        _csr_cache[key] = hit
    return hit


class _VertexNormals(torch.autograd.Function):
    """utils/io_utils.py:410-428 compute_normal_torch."""
    @staticmethod
    def forward(ctx, vertices, faces):
        _need_cuda(vertices)
        f32, vf_ptr, vf_face, Vmin = _faces_csr(faces)
        v = _f32c(vertices).reshape(-1, 3)
        V, F = v.shape[0], f32.shape[0]
        assert V >= Vmin, 'faces index past the vertex array'
        if V > Vmin:                                   # trailing vertices without faces: extend the CSR pointer
            vf_ptr = torch.cat([vf_ptr, vf_ptr[-1:].expand(V - Vmin)]).contiguous()
        nhat, nlen = torch.empty(F, 3, device=v.device), torch.empty(F, device=v.device)
        N, Nlen = torch.empty(V, 3, device=v.device), torch.empty(V, device=v.device)
        _call('bf_op_vertex_normals', v.data_ptr(), f32.data_ptr(), vf_ptr.data_ptr(), vf_face.data_ptr(), V, F,
              nhat.data_ptr(), nlen.data_ptr(), N.data_ptr(), Nlen.data_ptr())
        ctx.save_for_backward(v, f32, vf_ptr, vf_face, nhat, nlen, N, Nlen)
        ctx.shape = vertices.shape
        return N

    @staticmethod
    def backward(ctx, dN):
        v, f32, vf_ptr, vf_face, nhat, nlen, N, Nlen = ctx.saved_tensors
        V, F = v.shape[0], f32.shape[0]
        dm, dcorner, dv = torch.empty(V, 3, device=v.device), torch.empty(F, 9, device=v.device), torch.empty(V, 3, device=v.device)
        _call('bf_op_vertex_normals_backward', v.data_ptr(), f32.data_ptr(), vf_ptr.data_ptr(), vf_face.data_ptr(), V, F,
              nhat.data_ptr(), nlen.data_ptr(), N.data_ptr(), Nlen.data_ptr(), _f32c(dN).reshape(V, 3).data_ptr(), dm.data_ptr(),
              dcorner.data_ptr(), dv.data_ptr())
        return dv.reshape(ctx.shape), None


class _PcLoss(torch.autograd.Function):
    """|points - closest|_F with the closest points held constant (smplify/loss.py:239-241)."""
    @staticmethod
    def forward(ctx, points, closest):
        _need_cuda(points)
        p, c = _f32c(points).reshape(-1), _f32c(closest).reshape(-1)
        out, dp = torch.empty(1, device=p.device), torch.empty_like(p)
        _call('bf_op_pc_loss', p.data_ptr(), c.data_ptr(), p.numel(), out.data_ptr(), dp.data_ptr())
        ctx.save_for_backward(dp)
        ctx.shape = points.shape
        return out[0]

    @staticmethod
    def backward(ctx, dout):
        (dp,) = ctx.saved_tensors
        return (dp * dout).reshape(ctx.shape), None


class _NormalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, near_faces, face_norm, point_norm):
        _need_cuda(point_norm)
        nf = near_faces.detach().to(torch.int32).contiguous()
        fn, pn = _f32c(face_norm).reshape(-1, 3), _f32c(point_norm).reshape(-1, 3)
        V = pn.shape[0]
        out, d = torch.empty(1, device=pn.device), torch.empty_like(pn)
        _call('bf_op_normal_loss', nf.data_ptr(), fn.data_ptr(), pn.data_ptr(), V, out.data_ptr(), d.data_ptr())
        ctx.save_for_backward(d)
        ctx.shape = point_norm.shape
        return out[0]

    @staticmethod
    def backward(ctx, dout):
        (d,) = ctx.saved_tensors
        return None, None, (d * dout).reshape(ctx.shape)


class _Laplacian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, norms, faces):
        _need_cuda(norms)
        f32, vf_ptr, vf_face, Vmin = _faces_csr(faces)
        n = _f32c(norms).reshape(-1, 3)
        V, F = n.shape[0], f32.shape[0]
        if V > Vmin:
            vf_ptr = torch.cat([vf_ptr, vf_ptr[-1:].expand(V - Vmin)]).contiguous()
        out, d = torch.empty(1, device=n.device), torch.empty_like(n)
        _call('bf_op_laplacian', n.data_ptr(), f32.data_ptr(), vf_ptr.data_ptr(), vf_face.data_ptr(), V, F, out.data_ptr(), d.data_ptr())
        ctx.save_for_backward(d)
        ctx.shape = norms.shape
        return out[0]

    @staticmethod
    def backward(ctx, dout):
        (d,) = ctx.saved_tensors
        return (d * dout).reshape(ctx.shape), None


class _MaskLoss(torch.autograd.Function):
    """smplify/loss.py:85-130 on world vertices [B,V,3]; ``term`` = smplify.mask.SilhouetteTerm (masks, contours, cameras)."""
    @staticmethod
    def forward(ctx, verts, term):
        import ctypes as C
        _need_cuda(verts)
        v = _f32c(verts)
        B, V = v.shape[0], v.shape[1]
        scratch = torch.empty(8 * B, device=v.device)
        loss, dv = torch.empty(B, device=v.device), torch.empty_like(v)
        _lib.check(_lib.lib().bf_op_mask_loss(v.data_ptr(), B, V, C.byref(term.struct), scratch.data_ptr(), loss.data_ptr(),
                                              dv.data_ptr(), _stream()), 'bf_op_mask_loss')
        ctx.save_for_backward(dv)
        return loss.sum()

    @staticmethod
    def backward(ctx, dout):
        (dv,) = ctx.saved_tensors
        return dv * dout, None


def _dense_to_csr(Wm):
    """dense [R,N] numpy -> (ptr, idx, w) int32/int32/float32 numpy, entries in ascending column order"""
    Wm = np.asarray(Wm, dtype=np.float32)
    r, c = np.nonzero(Wm)
    ptr = np.zeros(Wm.shape[0] + 1, dtype=np.int32)
    np.add.at(ptr, r + 1, 1)
    return np.cumsum(ptr).astype(np.int32), c.astype(np.int32), Wm[r, c].astype(np.float32)


class SparseRegressor(object):
    """A fixed joint regressor [R,N] (e.g. J_regressor_h36m) as CSR tables of W and W^T on the device."""

    def __init__(self, Wm, device):
        Wm = np.asarray(Wm, dtype=np.float32)
        self.R, self.N = Wm.shape
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a if len(a) else np.zeros(1, a.dtype))).to(device)
        self.fwd = tuple(up(a) for a in _dense_to_csr(Wm))
        self.bwd = tuple(up(a) for a in _dense_to_csr(Wm.T))


class _Regress(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, reg):
        _need_cuda(points)
        p = _f32c(points)
        B, N = p.shape[0], p.shape[1]
        assert N == reg.N, 'regressor is for %d vertices, got %d' % (reg.N, N)
        out = torch.empty(B, reg.R, 3, device=p.device)
        ptr, idx, w = reg.fwd
        _call('bf_op_regress_joints', p.data_ptr(), ptr.data_ptr(), idx.data_ptr(), w.data_ptr(), B, N, reg.R, out.data_ptr())
        ctx.reg = reg
        return out

    @staticmethod
    def backward(ctx, dout):
        reg = ctx.reg
        d = _f32c(dout)
        B = d.shape[0]
        dp = torch.empty(B, reg.N, 3, device=d.device)
        ptr, idx, w = reg.bwd
        _call('bf_op_regress_joints', d.data_ptr(), ptr.data_ptr(), idx.data_ptr(), w.data_ptr(), B, reg.R, reg.N, dp.data_ptr())
        return dp, None


project = _Project.apply
regress_joints = _Regress.apply
vertex_normals = _VertexNormals.apply
pc_loss = _PcLoss.apply
normal_loss = _NormalLoss.apply
laplacian = _Laplacian.apply
mask_loss = _MaskLoss.apply
gmof_op = _Gmof.apply
reprojection_op = _Reprojection.apply
keypoints_world = _KeypointsWorld.apply
angle_prior_op = _AnglePrior.apply
gmm_pose = _GmmPose.apply
lbs = _Lbs.apply
