"""Per-batch device buffers and the calls into the C ABI.

Torch is used for device memory, the current stream and autograd plumbing only; every
arithmetic step of the path is a kernel in ``libbodyfit_b200.so``.
"""
import numpy as np
import torch

from . import _lib
from . import constants as K
from .model import PreparedModel


def _stream():
    return torch.cuda.current_stream().cuda_stream


class FrameBuffers(object):
    """Device buffers for B frames on one vertex set (``full=True``: all V vertices,
    else the active set).  Owns the ``BfFrames`` struct passed to the library."""

    def __init__(self, model: PreparedModel, B, full=False, Nv=0, need_backward=True, n_trace=0,
                 imsize=512.0, constant_scale=K.CONSTANT_SCALE_NO_SCAN, ext=None):
        _lib.require_device()
        self.model, self.B, self.full = model, int(B), bool(full)
        dev = model.device
        J, Kp, NP = model.J, model.Kp, model.NP
        self.ld_v = 3 * model.V if full else model.ld_act
        self.K_out = model.K_out if full else model.K_out_act
        f32 = dict(device=dev, dtype=torch.float32)
        t = {}
        ext = ext or {}

        def buf(name, *shape, zero=False, dtype=None):
            if name in ext and ext[name] is not None:
                t[name] = ext[name]
                return
            kw = dict(f32)
            if dtype is not None:
                kw['dtype'] = dtype
            t[name] = (torch.zeros if zero else torch.empty)(*shape, **kw)

        buf('theta', B, NP, zero=True)
        buf('pf', B, Kp)
        buf('A', B, J, 12)
        buf('Jtr', B, J, 3)
        buf('full_pose', B, 3 * J)
        buf('yaw', B, dtype=torch.int32, zero=True)
        buf('verts', B, self.ld_v)
        buf('joints', B, self.K_out, 3)
        buf('loss', B, zero=True)
        if need_backward:
            buf('grad', B, NP, zero=True)
            buf('adam_m', B, NP, zero=True)
            buf('adam_v', B, NP, zero=True)
            buf('dpf', B, Kp)
            buf('dA', B, J, 12)
            buf('dJtr', B, J, 3)
            buf('vposed', B, self.ld_v)
            buf('dverts', B, self.ld_v, zero=True)
            buf('dvp', B, self.ld_v)
            buf('loss_terms', B, 4, zero=True)
        if n_trace:
            buf('trace', n_trace, B, zero=True)
        self.t = t
        s = _lib.BfFrames()
        for name, ten in t.items():
            setattr(s, name, ten.data_ptr())
        s.B, s.Nv, s.ld_v, s.iter = self.B, int(Nv), self.ld_v, 0
        s.imsize, s.constant_scale, s.sigma = float(imsize), float(constant_scale), K.GMOF_SIGMA
        s.w_pose, s.w_angle, s.w_shape = K.POSE_PRIOR_WEIGHT, K.ANGLE_PRIOR_WEIGHT, K.SHAPE_PRIOR_WEIGHT
        s.lr_ts, s.lr = K.LR_TRANSL_SCALE, K.LR_DEFAULT
        s.beta1, s.beta2, s.eps = K.ADAM_BETAS[0], K.ADAM_BETAS[1], K.ADAM_EPS
        self.struct = s

    def bind(self, name, tensor):
        """Point a struct field at an externally owned tensor (kept alive here)."""
        self.t[name] = tensor
        setattr(self.struct, name, tensor.data_ptr() if tensor is not None else None)

    def call(self, fn_name, *extra):
        L = _lib.lib()
        rc = getattr(L, fn_name)(self.model.struct, self.struct, *extra, _stream())
        _lib.check(rc, fn_name)


def pack_cameras(c2ws, Ks):
    """c2w [Nv,4,4], K [Nv,3,3] -> [Nv,12] fp32 rows of K @ inv(c2w)[:3] (smplify/smplify.py:131-135,
    smplify/loss.py:38-41 pre-composed in fp64 on the host)."""
    c2ws = np.asarray([np.asarray(c.detach().cpu() if torch.is_tensor(c) else c, dtype=np.float64) for c in c2ws])
    Ks = np.asarray([np.asarray(k.detach().cpu() if torch.is_tensor(k) else k, dtype=np.float64) for k in Ks])
    w2c = np.linalg.inv(c2ws.astype(np.float32)).astype(np.float64)         # the reference inverts in fp32
    M = np.einsum('vij,vjk->vik', Ks, w2c[:, :3, :])
    return np.ascontiguousarray(M.reshape(len(c2ws), 12).astype(np.float32))


def pack_keypoints(kp, use_hand_face):
    """[B,Nv,K,3] (x, y, conf) -> (x, y, effective weight).  Body joints weigh conf^2; the
    reference passes hand / face confidences as [N,1], which broadcasts against the [N]
    residuals, so every joint of such a group weighs sum_i conf_i^2 of that group and view
    (smplify/loss.py:134 with :168,:173,:179)."""
    kp = torch.as_tensor(kp, dtype=torch.float32)
    out = kp.clone()
    c2 = kp[..., 2] ** 2
    w = c2.clone()
    if use_hand_face:
        for lo, hi in ((25, 46), (46, 67), (67, 135)):
            w[..., lo:hi] = c2[..., lo:hi].sum(-1, keepdim=True)
    out[..., 2] = w
    return out.contiguous()
