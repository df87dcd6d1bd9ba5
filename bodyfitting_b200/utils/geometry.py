"""Init conversion in front of the fitting path: rotation matrices of the pose
regressor -> axis-angle vectors, same entry points as the reference's ``utils/geometry.py``
(``rotation_matrix_to_angle_axis`` :331-351, ``convert_hom_to_angle`` :483-493).  Runs once per frame on
whatever device the input lives on (torch ops; not part of the per-iteration hot path)."""
import torch


def rotation_matrix_to_angle_axis(rotation_matrix):
    """(N,3,3) or (N,3,4) rotation -> (N,3) Rodrigues vector, via the unit quaternion with the largest
    component chosen as pivot (stable for every rotation angle); the identity maps to exactly 0."""
    R = rotation_matrix[:, :3, :3]
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    q = torch.stack([
        torch.stack([1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01], 1),      # pivot w
        torch.stack([m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20], 1),      # pivot x
        torch.stack([m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21], 1),      # pivot y
        torch.stack([m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22], 1)], 1)  # pivot z   -> [N,4(cand),4(wxyz)]
    diag = torch.stack([q[:, 0, 0], q[:, 1, 1], q[:, 2, 2], q[:, 3, 3]], 1)
    best = diag.argmax(1)
    quat = q[torch.arange(R.shape[0], device=R.device), best]
    quat = quat / quat.norm(dim=1, keepdim=True)
    quat = torch.where(quat[:, :1] < 0, -quat, quat)                                  # w >= 0 -> angle in [0, pi]
    w, v = quat[:, 0].clamp(-1, 1), quat[:, 1:]
    s = v.norm(dim=1)
    angle = 2 * torch.atan2(s, w)
    k = torch.where(s > 1e-12, angle / s.clamp_min(1e-12), torch.full_like(s, 2.0))
    return v * k[:, None]


def convert_hom_to_angle(pred_rotmat, batch_size, device=None):
    """(B,24,3,3) predicted rotation matrices -> (B,72) axis-angle pose (the reference pads to 3x4 and patches
    the NaN its converter returns for the identity; this converter returns 0 there directly)."""
    R = pred_rotmat.detach().reshape(-1, 3, 3)
    if device is not None:
        R = R.to(device)
    return rotation_matrix_to_angle_axis(R).contiguous().view(batch_size, -1)


def world_init_from_camera_rotmat(pred_rotmat, c2w):
    """The step between the pose regressor and the fit (reference: smplify/body_fitting.py:70-73): the regressor predicts the
    root orientation in the KEYFRAME CAMERA's frame; ``c2w[:3,:3] @ R_root`` carries it into world coordinates, then all 24
    rotation matrices become the (1,72) axis-angle pose SMPLify starts from.  ``pred_rotmat`` (1,24,3,3) is not modified
    (the reference overwrites it in place)."""
    R = pred_rotmat.detach().clone()
    c2w = torch.as_tensor(c2w, dtype=R.dtype, device=R.device)
    R[0, 0] = c2w[:3, :3] @ R[0, 0]
    return convert_hom_to_angle(R, R.shape[0], R.device)
