"""Loss terms of the reference's ``smplify/loss.py`` with the same names, arguments and semantics,
running on the B200 kernels (CUDA tensors only; autograd supported through hand-written backward
kernels).  The fused fit loop (smplify.smplify.SMPLify) does not go through these functions.

Not provided: the unused ``point_cloud_loss_chamfer_naive`` (loss.py:245-258).
"""
import numpy as np
import torch

from .. import constants as C
from .. import ops
from ..engine import pack_cameras, pack_keypoints

SKELETON_LENGTH = C.SKELETON_LENGTH
HANDS_LENGTH = C.HANDS_LENGTH
FACE_LENGTH = C.FACE_LENGTH
FACE_MAPPING = C.FACE_MAPPING


def perspective_projection(points, rotation, translation, K):
    """points (bs,N,3), rotation (bs|1,3,3), translation (bs|1,3), K (3,3) -> (bs,N,2)  [loss.py:22-43]"""
    if isinstance(K, np.ndarray):
        K = torch.tensor(K, dtype=torch.float32, device=points.device)
    return ops.project(points, rotation, translation, K)


def gmof(x, sigma):
    """Geman-McClure error function [loss.py:45-51]"""
    return ops.gmof_op(x, sigma)


def angle_prior(pose):
    """exp(pose[:, [52,55,9,12]] * [1,-1,-1,-1]) ** 2 -> (B,4)  [loss.py:54-61]"""
    return ops.angle_prior_op(pose)


def reprojection_loss(cord, cord_gt, conf, scale_coeff, sigma):
    """[loss.py:132-136]; ``conf`` of shape [N] weighs joint j by conf_j^2; of shape [N,1] it broadcasts
    against the [N] residual sums exactly as in the reference and returns the [N] vector
    conf_i^2 * sum_j rho_j."""
    if conf.dim() == 2:
        total = ops.reprojection_op(cord, cord_gt, torch.ones_like(conf[:, 0]), scale_coeff, sigma)
        return (conf[:, 0] ** 2) * total
    return ops.reprojection_op(cord, cord_gt, conf ** 2, scale_coeff, sigma)


def multiview_keypoint_loss(w2cs, Ks, keypoints, model_joints, poses, betas, use_frames, pose_prior, sigma=100,
                            shape_prior_weight=5, angle_prior_weight=15.2, output='sum', debug=False, imsize=512,
                            pose_prior_weight=4.78, use_hand_face=False, output_folder=None, verts=None):
    """[loss.py:139-230]  model_joints (B,K,3) in world space; keypoints = list (per view) of OpenPose dicts
    (or None), or a packed [B,Nv,K,3] array.  Returns (total.sum(), dict of the four terms)."""
    from ..synthetic import openpose_to_keypoints
    dev = model_joints.device
    B = model_joints.shape[0]
    if isinstance(keypoints, (np.ndarray, torch.Tensor)):
        kp = torch.as_tensor(keypoints, dtype=torch.float32)
        kp = kp[None] if kp.dim() == 3 else kp
    else:
        kp = torch.from_numpy(openpose_to_keypoints(keypoints, 'smplx' if use_hand_face else 'smpl'))[None]
    Kn = kp.shape[2]
    kp = pack_keypoints(kp.to(dev), use_hand_face)                # -> [1 or B, K, Nv, 3]
    if torch.is_tensor(w2cs):
        c2ws = [np.linalg.inv(w.detach().cpu().numpy()) for w in w2cs]
    else:
        c2ws = [np.linalg.inv(np.asarray(w)) for w in w2cs]
    cams = torch.from_numpy(pack_cameras(c2ws, Ks)).to(dev)
    data = ops.keypoints_world(model_joints[:, :Kn], kp.expand(B, -1, -1, -1).contiguous(), cams, imsize / 1024, sigma)
    loss_2d = data
    if use_hand_face:
        poses = torch.cat([poses, torch.zeros_like(poses[:, :6])], dim=-1)
    pose_prior_loss = (pose_prior_weight ** 2) * pose_prior(poses, None)
    angle_prior_loss = (angle_prior_weight ** 2) * angle_prior(poses).sum(dim=-1)
    shape_prior_loss = (shape_prior_weight ** 2) * (betas ** 2).sum(dim=-1)
    total = loss_2d + pose_prior_loss + angle_prior_loss + shape_prior_loss
    cpu = lambda t: t.detach().cpu().numpy()
    losses = dict(reprojection_loss=cpu(loss_2d), pose_prior_loss=cpu(pose_prior_loss),
                  angle_prior_loss=cpu(angle_prior_loss), shape_prior_loss=cpu(shape_prior_loss))
    if output == 'sum':
        return total.sum(), losses
    return reprojection_loss, losses


def point_cloud_loss_mesh_grid(mesh_grid_searcher, points):
    """[loss.py:233-242] point-to-scan term: the Frobenius norm of (points - their closest scan points); the closest points
    are constants of the objective (the reference detaches them, and its search has no backward)."""
    pts = points.reshape(-1, 3)
    closest, _ = mesh_grid_searcher.nearest_points(pts.detach())
    return ops.pc_loss(pts, closest)


def normal_loss_mesh_grid(mesh_grid_searcher, points, face_norm_mesh, point_norm):
    """[loss.py:260-271] mean(1 - <face normal of the closest scan face, vertex normal>); differentiable w.r.t. ``point_norm``."""
    _, faces = mesh_grid_searcher.nearest_points(points.reshape(-1, 3).detach())
    return ops.normal_loss(faces, face_norm_mesh, point_norm)


def normal_laplacian_smoothness(norms, faces):
    """[loss.py:273-288] mean over faces of |na - nb|^2 + |nc - na|^2 + |nb - nc|^2."""
    return ops.laplacian(norms, faces)


def extract_countours(masks):
    """[loss.py:73-83] (the reference's spelling) external contour of every mask [Nm,H,W] -> list of float tensors [Nc,1,2]
    of (x, y) pixels on the masks' device.  As the reference does, the contour kept is ``contours[argmax(shape[1])]`` -- OpenCV
    contours are [Nc,1,2], so that is always the FIRST external contour it returns, not the longest."""
    from .mask import extract_contours
    dev = masks.device if torch.is_tensor(masks) else 'cpu'
    arr = masks.detach().cpu().numpy() if torch.is_tensor(masks) else np.asarray(masks)
    return [torch.from_numpy(c.reshape(-1, 1, 2)).to(dev) for c in extract_contours(arr)]


def multview_mask_loss(contours, masks, smpl_verts, smpl_faces, w2cs, Ks, mask_frames, epsilon=10, imsize=512):
    """[loss.py:85-130] (the reference's spelling) silhouette term for one frame: ``contours`` from extract_countours,
    ``masks`` [Nm,H,W] 0/1, ``smpl_verts`` [1,V,3] world vertices, ``w2cs`` / ``Ks`` of the mask views.  Differentiable w.r.t.
    ``smpl_verts``.  (``smpl_faces`` / ``mask_frames`` are accepted and unused, as in the reference.)"""
    from .mask import SilhouetteTerm
    dev = smpl_verts.device
    w2c = [w.detach().cpu().numpy() if torch.is_tensor(w) else np.asarray(w) for w in w2cs]
    c2ws = [np.linalg.inv(w.astype(np.float64)) for w in w2c]
    cams = pack_cameras(c2ws, [k.detach().cpu().numpy() if torch.is_tensor(k) else np.asarray(k) for k in Ks])
    mk = masks.detach().cpu().numpy() if torch.is_tensor(masks) else np.asarray(masks)
    cont = [c.detach().cpu().numpy().reshape(-1, 2).astype(np.float32) for c in contours]
    term = SilhouetteTerm(None, (mk > 0.5).astype(np.uint8) * 255, cams, imsize=imsize, epsilon=float(epsilon), device=dev,
                          contours=[cont], num_verts=smpl_verts.shape[-2])
    return ops.mask_loss(smpl_verts.reshape(1, -1, 3), term)
