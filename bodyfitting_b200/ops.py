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


project = _Project.apply
gmof_op = _Gmof.apply
reprojection_op = _Reprojection.apply
keypoints_world = _KeypointsWorld.apply
angle_prior_op = _AnglePrior.apply
gmm_pose = _GmmPose.apply
lbs = _Lbs.apply
