"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

CPU (torch) restatement of the linear-blend-skinning arithmetic of the third-party
``smplx`` package, which the reference pins as ``smplx==0.1.13``
(/root/reference/requirements.txt:5) but does NOT vendor; it is absent from this
image.  The reference's call sites are ``models/smpl.py:3-6,60,71-72`` and
``smplify/smplify.py:7,59-80,179-187``.  The published algorithm (smplx/lbs.py,
smplx/body_models.py, smplx/vertex_joint_selector.py) is restated here from its
public description:

  v_shaped = v_template + sum_l betas_l * shapedirs[..., l]
  J        = J_regressor @ v_shaped
  R        = rodrigues(pose)    with angle = ||r + 1e-8||, R = I + sin K + (1-cos) K^2
  v_posed  = v_shaped + (R[1:] - I).flatten() @ posedirs
  chain    : T_i = T_parent(i) @ [R_i | J_i - J_parent(i)],  A_i = T_i - [0 | T_i @ J_i]
  verts    = (sum_j W_vj A_j) @ [v_posed; 1]
  joints   = chain translations ++ picked vertices (++ barycentric face landmarks, SMPL-X)

PARITY STATUS: **unpinned** against the real ``smplx`` package (no copy of it exists
offline, and the reference ships no golden vectors for this path -- SURVEY.md 8c).
The restatement is pinned only against itself (fp64 vs fp32, finite differences) and
is the single definition both the verbatim reference loop (via oracle/shim/smplx) and
the CUDA kernels are compared with.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# smplx.vertex_ids (hard-coded in the package; the official model files do not carry them) -- restated independently of the
# product's copy in bodyfitting_b200/constants.py; tests/test_host.py checks that the two agree
_EXTRA_VIDS = {
    'smpl': [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787, 2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133],
    'smplx': [9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635, 5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022],
}


def batch_rodrigues(rot_vecs):
    """[N,3] axis-angle -> [N,3,3]; smplx.lbs.batch_rodrigues (epsilon added to the
    vector before the norm, direction = r / angle)."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.unsqueeze(torch.cos(angle), dim=1)
    sin = torch.unsqueeze(torch.sin(angle), dim=1)
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype, device=rot_vecs.device)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def blend_shapes(betas, shape_disps):
    return torch.einsum('bl,mkl->bmk', betas, shape_disps)


def vertices2joints(J_regressor, vertices):
    return torch.einsum('bik,ji->bjk', vertices, J_regressor)


def batch_rigid_transform(rot_mats, joints, parents):
    """rot_mats [B,J,3,3], joints [B,J,3], parents [J] -> posed joints [B,J,3] and
    rest-pose-removed transforms [B,J,4,4] (smplx.lbs.batch_rigid_transform)."""
    B, J = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    top = torch.cat([rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)], dim=2)
    bottom = torch.zeros((B * J, 1, 4), dtype=joints.dtype, device=joints.device)
    bottom[:, 0, 3] = 1
    tm = torch.cat([top, bottom], dim=1).reshape(B, J, 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, J):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_h = F.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - F.pad(torch.matmul(transforms, joints_h), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    B = max(betas.shape[0], pose.shape[0])
    dtype, device = betas.dtype, betas.device
    v_shaped = v_template + blend_shapes(betas, shapedirs)
    J = vertices2joints(J_regressor, v_shaped)
    ident = torch.eye(3, dtype=dtype, device=device)
    rot_mats = batch_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:, :, :] - ident).reshape(B, -1)
    pose_offsets = torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    v_posed = pose_offsets + v_shaped
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents)
    W = lbs_weights.unsqueeze(0).expand(B, -1, -1)
    nj = J_regressor.shape[0]
    T = torch.matmul(W, A.view(B, nj, 16)).view(B, -1, 4, 4)
    ones = torch.ones((B, v_posed.shape[1], 1), dtype=dtype, device=device)
    v_homo = torch.matmul(T, torch.cat([v_posed, ones], dim=2).unsqueeze(-1))
    return v_homo[:, :, :3, 0], J_transformed


def rot_mat_to_euler(rot_mats):
    sy = torch.sqrt(rot_mats[:, 0, 0] * rot_mats[:, 0, 0] + rot_mats[:, 1, 0] * rot_mats[:, 1, 0])
    return torch.atan2(-rot_mats[:, 2, 0], sy)


def find_dynamic_lmk_idx_and_bcoords(vertices, pose, dynamic_lmk_faces_idx, dynamic_lmk_b_coords,
                                     neck_kin_chain):
    """Contour-landmark look-up by head yaw (smplx.lbs.find_dynamic_lmk_idx_and_bcoords):
    yaw of R_0 R_3 R_6 R_9 R_12 in degrees, negated, clamped to <=39, rounded; negative
    values map to 39-y (78 below -39)."""
    B = vertices.shape[0]
    aa = torch.index_select(pose.view(B, -1, 3), 1, neck_kin_chain)
    rot_mats = batch_rodrigues(aa.reshape(-1, 3)).view(B, -1, 3, 3)
    rel = torch.eye(3, dtype=vertices.dtype, device=vertices.device).unsqueeze(0).repeat(B, 1, 1)
    for idx in range(len(neck_kin_chain)):
        rel = torch.bmm(rot_mats[:, idx], rel)
    y = torch.round(torch.clamp(-rot_mat_to_euler(rel) * 180.0 / np.pi, max=39)).to(dtype=torch.long)
    neg_mask = y.lt(0).to(dtype=torch.long)
    mask = y.lt(-39).to(dtype=torch.long)
    neg_vals = mask * 78 + (1 - mask) * (39 - y)
    y = neg_mask * neg_vals + (1 - neg_mask) * y
    return torch.index_select(dynamic_lmk_faces_idx, 0, y), torch.index_select(dynamic_lmk_b_coords, 0, y)


def vertices2landmarks(vertices, faces, lmk_faces_idx, lmk_bary_coords):
    B, V = vertices.shape[:2]
    lmk_faces = torch.index_select(faces, 0, lmk_faces_idx.reshape(-1)).view(B, -1, 3)
    lmk_faces = lmk_faces + torch.arange(B, dtype=torch.long, device=vertices.device).view(-1, 1, 1) * V
    lmk_vertices = vertices.reshape(-1, 3)[lmk_faces].view(B, -1, 3, 3)
    return torch.einsum('blfi,blf->bli', lmk_vertices, lmk_bary_coords)


class _Output(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _load(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    d = np.load(path_or_dict, allow_pickle=True)
    return {k: d[k] for k in d.files}


def _parents_from(data):
    kt = np.asarray(data['kintree_table'])[0].astype(np.int64)
    kt[0] = -1
    return torch.tensor(kt, dtype=torch.long)


class SMPLLayer(nn.Module):
    """smplx.SMPL restated: 24 chain joints + 21 vertex-picked joints = 45."""
    NUM_BODY_JOINTS = 23

    def __init__(self, data, num_betas=10, dtype=torch.float32, joint_mapper=None, create_transl=False,
                 batch_size=1, age='adult', kid_template_path=''):
        super().__init__()
        data = _load(data)
        if age == 'kid':
            # smplx (>= 0.1.26) kid model [recalled]: v_template_smil = np.load(kid_template_path); mean-centred; its
            # difference to the adult template is appended to the first num_betas shape directions; num_betas += 1
            data = dict(data)
            smil = kid_template_path if not isinstance(kid_template_path, str) else np.load(kid_template_path, allow_pickle=True)
            smil = np.array(smil, dtype=np.float64)
            smil -= np.mean(smil, axis=0)
            diff = np.expand_dims(smil - np.asarray(data['v_template'], dtype=np.float64), axis=2)
            data['shapedirs'] = np.concatenate((np.asarray(data['shapedirs'])[:, :, :num_betas], diff), axis=2)
            num_betas = num_betas + 1
        self.num_betas = num_betas
        V = data['v_template'].shape[0]
        self.faces = np.asarray(data['f'])
        self.register_buffer('faces_tensor', torch.tensor(self.faces.astype(np.int64)))
        self.register_buffer('v_template', torch.tensor(np.asarray(data['v_template']), dtype=dtype))
        self.register_buffer('shapedirs', torch.tensor(np.asarray(data['shapedirs'])[:, :, :num_betas], dtype=dtype))
        P = data['posedirs'].shape[-1]
        self.register_buffer('posedirs', torch.tensor(np.reshape(np.asarray(data['posedirs']), [-1, P]).T.copy(), dtype=dtype))
        self.register_buffer('J_regressor', torch.tensor(np.asarray(data['J_regressor']), dtype=dtype))
        self.register_buffer('parents', _parents_from(data))
        self.register_buffer('lbs_weights', torch.tensor(np.asarray(data['weights']), dtype=dtype))
        self.register_buffer('extra_joints_idxs', torch.tensor(np.asarray(data['extra_vids'] if 'extra_vids' in data else _EXTRA_VIDS[getattr(self, 'MODEL_TYPE', 'smpl')]), dtype=torch.long))
        self.joint_mapper = joint_mapper
        if create_transl:
            self.register_parameter('transl', nn.Parameter(torch.zeros([batch_size, 3], dtype=dtype)))

    def _select(self, vertices, joints):
        return torch.cat([joints, torch.index_select(vertices, 1, self.extra_joints_idxs)], dim=1)

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, return_verts=True,
                return_full_pose=False, **kwargs):
        apply_trans = transl is not None or hasattr(self, 'transl')
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        vertices, joints = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs,
                               self.J_regressor, self.parents, self.lbs_weights)
        joints = self._select(vertices, joints)
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints)
        if apply_trans:
            joints = joints + transl.unsqueeze(dim=1)
            vertices = vertices + transl.unsqueeze(dim=1)
        return _Output(vertices=vertices if return_verts else None, global_orient=global_orient,
                       body_pose=body_pose, joints=joints, betas=betas,
                       full_pose=full_pose if return_full_pose else None)


class SMPLXLayer(SMPLLayer):
    """smplx.SMPLX restated: PCA hands (6 comps), pose_mean, expression, 55 chain joints
    + 21 picked + 51 static + 17 dynamic landmarks = 144 joints, then ``joint_mapper``."""
    NUM_BODY_JOINTS = 21
    NECK_IDX = 12
    MODEL_TYPE = 'smplx'

    def __init__(self, data, num_betas=10, num_expression_coeffs=10, num_pca_comps=6, dtype=torch.float32,
                 joint_mapper=None, use_face_contour=True, create_transl=False, batch_size=1):
        data = _load(data)
        super().__init__(data, num_betas=num_betas, dtype=dtype, joint_mapper=joint_mapper,
                         create_transl=create_transl, batch_size=batch_size)
        sd = np.asarray(data['shapedirs'])
        e0 = 300 if sd.shape[-1] >= 300 + num_expression_coeffs else num_betas     # official files: 300 shape + 100 expression dirs
        self.register_buffer('expr_dirs', torch.tensor(sd[:, :, e0:e0 + num_expression_coeffs], dtype=dtype))
        self.register_buffer('left_hand_components', torch.tensor(np.asarray(data['hands_componentsl'])[:num_pca_comps], dtype=dtype))
        self.register_buffer('right_hand_components', torch.tensor(np.asarray(data['hands_componentsr'])[:num_pca_comps], dtype=dtype))
        pose_mean = np.concatenate([np.zeros(3), np.zeros(63), np.zeros(3), np.zeros(3), np.zeros(3),
                                    np.asarray(data['hands_meanl']), np.asarray(data['hands_meanr'])])
        self.register_buffer('pose_mean', torch.tensor(pose_mean, dtype=dtype))
        self.register_buffer('lmk_faces_idx', torch.tensor(np.asarray(data['lmk_faces_idx']), dtype=torch.long))
        self.register_buffer('lmk_bary_coords', torch.tensor(np.asarray(data['lmk_bary_coords']), dtype=dtype))
        self.use_face_contour = use_face_contour
        self.register_buffer('dynamic_lmk_faces_idx', torch.tensor(np.asarray(data['dynamic_lmk_faces_idx']), dtype=torch.long))
        self.register_buffer('dynamic_lmk_bary_coords', torch.tensor(np.asarray(data['dynamic_lmk_bary_coords']), dtype=dtype))
        chain, curr = [], self.NECK_IDX
        while curr != -1:
            chain.append(curr)
            curr = int(self.parents[curr])
        self.register_buffer('neck_kin_chain', torch.tensor(chain, dtype=torch.long))
        # the module's own zero expression Parameter (create_expression=True, never optimised)
        self.register_parameter('expression', nn.Parameter(torch.zeros([batch_size, num_expression_coeffs], dtype=dtype)))

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None,
                right_hand_pose=None, transl=None, expression=None, jaw_pose=None, leye_pose=None,
                reye_pose=None, return_verts=True, return_full_pose=False, **kwargs):
        expression = expression if expression is not None else self.expression
        apply_trans = transl is not None or hasattr(self, 'transl')
        if transl is None and hasattr(self, 'transl'):
            transl = self.transl
        left_hand_pose = torch.einsum('bi,ij->bj', left_hand_pose, self.left_hand_components)
        right_hand_pose = torch.einsum('bi,ij->bj', right_hand_pose, self.right_hand_components)
        full_pose = torch.cat([global_orient.reshape(-1, 1, 3), body_pose.reshape(-1, self.NUM_BODY_JOINTS, 3),
                               jaw_pose.reshape(-1, 1, 3), leye_pose.reshape(-1, 1, 3), reye_pose.reshape(-1, 1, 3),
                               left_hand_pose.reshape(-1, 15, 3), right_hand_pose.reshape(-1, 15, 3)],
                              dim=1).reshape(-1, 165)
        full_pose = full_pose + self.pose_mean
        B = max(betas.shape[0], global_orient.shape[0], body_pose.shape[0])
        if expression.shape[0] != B:
            expression = expression.expand(B, -1)
        shape_components = torch.cat([betas, expression], dim=-1)
        shapedirs = torch.cat([self.shapedirs, self.expr_dirs], dim=-1)
        vertices, joints = lbs(shape_components, full_pose, self.v_template, shapedirs, self.posedirs,
                               self.J_regressor, self.parents, self.lbs_weights)
        lmk_faces_idx = self.lmk_faces_idx.unsqueeze(0).expand(B, -1).contiguous()
        lmk_bary = self.lmk_bary_coords.unsqueeze(0).repeat(B, 1, 1)
        if self.use_face_contour:
            dyn_idx, dyn_bary = find_dynamic_lmk_idx_and_bcoords(
                vertices, full_pose, self.dynamic_lmk_faces_idx, self.dynamic_lmk_bary_coords, self.neck_kin_chain)
            lmk_faces_idx = torch.cat([lmk_faces_idx, dyn_idx], 1)
            lmk_bary = torch.cat([lmk_bary, dyn_bary], 1)
        landmarks = vertices2landmarks(vertices, self.faces_tensor, lmk_faces_idx, lmk_bary)
        joints = self._select(vertices, joints)
        joints = torch.cat([joints, landmarks], dim=1)
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints=joints, vertices=vertices)
        if apply_trans:
            joints = joints + transl.unsqueeze(dim=1)
            vertices = vertices + transl.unsqueeze(dim=1)
        return _Output(vertices=vertices if return_verts else None, joints=joints, betas=betas,
                       expression=expression, global_orient=global_orient, body_pose=body_pose,
                       left_hand_pose=left_hand_pose, right_hand_pose=right_hand_pose, jaw_pose=jaw_pose,
                       full_pose=full_pose if return_full_pose else None)
