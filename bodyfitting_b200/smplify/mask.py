"""Silhouette term of the fitting loop (``use_mask=True``): host side.

``extract_contours`` mirrors smplify/loss.py:73-83 (OpenCV on the host, once per fit: the longest external contour of
every mask); ``SilhouetteTerm`` packs masks / contours / cameras for ``bf_mask_loss`` (include/bodyfit_b200_mask.h),
the CUDA replacement of ``multview_mask_loss`` (smplify/loss.py:85-130)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..engine import _stream


def extract_contours(masks, pick='reference'):
    """masks [Nm,H,W] (bool / 0-1 / 0-255) -> list of float32 [Nc,2] (x, y) contour pixels of one external contour per
    mask (cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE).  The reference unpacks OpenCV 3's three return values (loss.py:79);
    OpenCV 4 returns (contours, hierarchy): the contours are the second-to-last element in both.
    ``pick='reference'``: the reference keeps ``contour[argmax(a.shape[1] for a in contour)]`` (loss.py:80) -- OpenCV contours
    are [Nc,1,2], so every ``shape[1]`` is 1 and the argmax is 0: the FIRST contour OpenCV returns (for a mask with several
    blobs not necessarily the largest).  Reproduced as is; ``pick='longest'`` selects the longest contour instead (the
    evident intent), at the price of differing from the reference on multi-blob masks."""
    try:
        import cv2
    except ImportError as e:                                   # pragma: no cover
        raise _lib.BodyfitError('use_mask=True needs OpenCV (cv2) for cv2.findContours, as the reference does') from e
    out = []
    for mask in np.asarray(masks):
        res = cv2.findContours((np.asarray(mask) > 0).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        cs = res[-2]
        if len(cs) == 0:
            out.append(np.zeros((0, 2), np.float32))
            continue
        c = cs[int(np.argmax(np.array([a.shape[0] for a in cs])))] if pick == 'longest' else cs[0]
        out.append(np.ascontiguousarray(c.reshape(-1, 2), dtype=np.float32))
    return out


class SilhouetteTerm(object):
    """Device state of the silhouette term for B frames x Nm mask views."""

    def __init__(self, model, masks, mask_cams, imsize=512, epsilon=10.0, stride=4, device='cuda', contours=None,
                 num_verts=None, pick='reference'):
        """``contours`` (optional): per frame a list of [Nc,2] arrays, one per mask view (else extracted here);
        ``num_verts``: body vertex count when no model is given (stand-alone operator)."""
        dev = torch.device(device)
        masks = np.asarray(masks)
        if masks.ndim == 3:
            masks = masks[None]
        B, Nm, H, W = masks.shape
        binm = masks > 128                                                          # smplify.py:139
        self.masks = torch.from_numpy(binm.astype(np.float32)).to(dev).contiguous()
        cont, cptr, cown = [], [0], []
        for b in range(B):
            for m, c in enumerate(contours[b] if contours is not None else extract_contours(binm[b], pick=pick)):
                cont.append(c)
                cptr.append(cptr[-1] + len(c))
                cown += [b * Nm + m] * len(c)
        total = cptr[-1]
        self.contour = torch.from_numpy(np.concatenate(cont) if total else np.zeros((1, 2), np.float32)).to(dev).contiguous()
        self.cptr = torch.tensor(cptr, dtype=torch.int32, device=dev)
        self.cown = torch.tensor(cown if total else [0], dtype=torch.int32, device=dev)
        self.cams = torch.as_tensor(np.asarray(mask_cams, dtype=np.float32)).reshape(Nm, 12).to(dev).contiguous()
        Nq = ((model.V if model is not None else int(num_verts)) + stride - 1) // stride
        f32 = dict(device=dev, dtype=torch.float32)
        self.uv = torch.empty(B, Nm, Nq, 2, **f32)
        self.near_q = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        self.cdist = torch.empty(max(total, 1), **f32)
        self.cw = torch.empty(max(total, 1), **f32)
        self.dPw = torch.empty(B, Nq, 3, **f32)
        self.part = torch.empty(B, Nq, **f32)
        self.mask_loss = torch.zeros(B, **f32)
        self.model, self.B = model, B
        s = _lib.BfMask()
        for name in ('masks', 'cams', 'contour', 'cptr', 'cown', 'uv', 'near_q', 'cdist', 'cw', 'dPw', 'part', 'mask_loss'):
            setattr(s, name, getattr(self, name).data_ptr())
        s.Nm, s.H, s.W, s.Nq, s.stride, s.total = Nm, H, W, Nq, stride, total
        s.imsize, s.epsilon = float(imsize), float(epsilon)
        self.struct = s

    def add(self, fb, weight=5.0):
        """fb: all-vertex FrameBuffers after the forward pass.  loss += weight * term; gradient added to dverts / grad[:, :4]."""
        assert fb.B == self.B and fb.full
        _lib.check(_lib.lib().bf_mask_loss(self.model.struct, fb.struct, C.byref(self.struct), float(weight), _stream()), 'bf_mask_loss')
