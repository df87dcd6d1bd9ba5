"""Seeded synthetic SMPL / SMPL-X shaped model tensors, GMM prior, cameras and
OpenPose-layout keypoints.

Official model files are not redistributable (and not available offline), so every
test / bench in this repo runs on random-init tensors with the official *shapes*
(SURVEY.md section 8d): 6890 / 10475 vertices, 24 / 55 joints, 207 / 486 pose
features, the official kinematic trees and the official extra-joint vertex ids.

This module is data synthesis only (numpy): it holds no model arithmetic.  The
keypoints of a scene are produced by projecting joints that the *caller* computed
(the CUDA path in bench.py, the CPU oracle in tests) -- see ``make_keypoints``.

File formats written by ``save_model_npz`` follow the key names of the official
SMPL / SMPL-X ``.npz`` files as consumed by ``smplx`` (reference call sites
``models/smpl.py:56-66`` and ``smplify/smplify.py:63-80``).
"""
import os
import pickle

import numpy as np

# --- official kinematic trees -------------------------------------------------
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def _smplx_parents():
    p = SMPL_PARENTS[:22] + [15, 15, 15]
    for wrist in (20, 21):                      # 5 fingers x 3 phalanges per hand
        base = len(p)
        for f in range(5):
            p += [wrist, base + 3 * f, base + 3 * f + 1]
    return p


SMPLX_PARENTS = _smplx_parents()

# extra joints picked from vertices, order = face(5), feet(6), left tips(5), right tips(5)
from .constants import EXTRA_VIDS as _EXTRA_VIDS  # noqa: E402
SMPL_EXTRA_VIDS, SMPLX_EXTRA_VIDS = _EXTRA_VIDS['smpl'], _EXTRA_VIDS['smplx']

# OpenPose face (70 pts) -> model landmark order (51 inner + 17 contour), smplify/loss.py:20
FACE_MAPPING = list(range(17, 17 + 51)) + list(range(0, 17))

# rough T-pose joint directions (x left, y up, z front) used to seed joint regressors
_SKEL24 = np.array([
    [0, 0, 0], [.07, -.09, 0], [-.07, -.09, 0], [0, .11, 0], [.10, -.47, 0], [-.10, -.47, 0],
    [0, .25, 0], [.09, -.87, -.03], [-.09, -.87, -.03], [0, .30, 0], [.11, -.93, .09],
    [-.11, -.93, .09], [0, .51, -.02], [.08, .42, 0], [-.08, .42, 0], [0, .58, .03],
    [.17, .45, 0], [-.17, .45, 0], [.43, .44, 0], [-.43, .44, 0], [.68, .44, 0], [-.68, .44, 0],
    [.76, .43, 0], [-.76, .43, 0]], dtype=np.float64)


def _skeleton_dirs(parents):
    J = len(parents)
    sk = np.zeros((J, 3))
    sk[:22] = _SKEL24[:22]
    if J == 24:
        sk[22:] = _SKEL24[22:]
        return sk
    sk[22] = [0, .55, .06]
    sk[23] = [.03, .62, .08]
    sk[24] = [-.03, .62, .08]
    for h, (wrist, sgn) in enumerate(((20, 1.0), (21, -1.0))):
        for f in range(5):
            for k in range(3):
                j = 25 + 15 * h + 3 * f + k
                sk[j] = sk[wrist] + np.array([sgn * (.05 + .03 * k), .01 * (f - 2), .015 * (f - 2)])
    return sk


def make_template(num_verts, seed=0):
    """Closed genus-0 triangle mesh with exactly ``num_verts`` vertices and
    ``2*num_verts-4`` faces: a Fibonacci sphere triangulated by its convex hull,
    then stretched to a body-sized, mildly bumpy star-shaped blob (the radial
    deformation keeps the triangulation valid)."""
    from scipy.spatial import ConvexHull
    i = np.arange(num_verts) + 0.5
    phi = np.arccos(1 - 2 * i / num_verts)
    th = np.pi * (1 + 5 ** 0.5) * i
    p = np.stack([np.cos(th) * np.sin(phi), np.cos(phi), np.sin(th) * np.sin(phi)], 1)
    hull = ConvexHull(p)
    faces = hull.simplices.astype(np.int64)
    # consistent outward orientation
    a, b, c = p[faces[:, 0]], p[faces[:, 1]], p[faces[:, 2]]
    flip = np.einsum('ij,ij->i', np.cross(b - a, c - a), a + b + c) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    faces = faces[np.lexsort((faces[:, 2], faces[:, 1], faces[:, 0]))]
    rng = np.random.RandomState(seed + 17)
    k = rng.normal(size=(4, 3))
    bump = 1.0 + 0.08 * np.sin(3.0 * p @ k[0]) * np.cos(2.0 * p @ k[1]) + 0.05 * np.sin(5.0 * p @ k[2])
    v = p * bump[:, None] * np.array([0.45, 0.85, 0.20])
    return v.astype(np.float32), faces.astype(np.int32)


def make_model(model_type='smpl', seed=0, num_betas=10, num_expression=10):
    """Random-init model dict with official shapes.  Keys (all numpy):
    v_template[V,3] f32, f[F,3] i32, shapedirs[V,3,nb(+ne)] f32, posedirs[V,3,P] f32
    (official on-disk layout; the loaders reshape to [P,3V]), J_regressor[J,V] f32
    (dense storage, ~32 non-zeros/row), weights[V,J] f32 (4 non-zeros/row),
    kintree_table[2,J] i64, extra_vids[21] i64; SMPL-X adds hands_components{l,r}[6,45],
    hands_mean{l,r}[45], lmk_faces_idx[51], lmk_bary_coords[51,3],
    dynamic_lmk_faces_idx[79,17], dynamic_lmk_bary_coords[79,17,3]."""
    assert model_type in ('smpl', 'smplx')
    rng = np.random.RandomState(seed)
    smplx = model_type == 'smplx'
    V = 10475 if smplx else 6890
    parents = SMPLX_PARENTS if smplx else SMPL_PARENTS
    J = len(parents)
    P = (J - 1) * 9
    nshape = num_betas + (num_expression if smplx else 0)
    v_template, faces = make_template(V, seed)

    shapedirs = (rng.standard_normal((V, 3, nshape)) * 0.01).astype(np.float32)
    posedirs = (rng.standard_normal((V, 3, P)).astype(np.float32) * np.float32(0.001))

    # joint regressors: ~32 nearest vertices of a seed vertex, positive weights summing to 1
    sk = _skeleton_dirs(parents)
    sk = sk * np.array([0.5, 0.8, 1.0])
    J_regressor = np.zeros((J, V), dtype=np.float32)
    vt = v_template.astype(np.float64)
    for j in range(J):
        d = np.linalg.norm(vt - sk[j], axis=1)
        seed_v = int(np.argmin(d))
        nn = np.argsort(np.linalg.norm(vt - vt[seed_v], axis=1))[:32]
        w = rng.uniform(0.1, 1.0, size=32)
        J_regressor[j, nn] = (w / w.sum()).astype(np.float32)
    J_rest = J_regressor.astype(np.float64) @ vt

    # skinning weights: 4 nearest rest joints, gaussian falloff
    d2 = ((vt[:, None, :] - J_rest[None]) ** 2).sum(-1)                   # [V,J]
    nn = np.argsort(d2, axis=1)[:, :4]
    dn = np.take_along_axis(d2, nn, 1)
    w = np.exp(-(dn - dn[:, :1]) / 0.01) + 0.02
    w = w / w.sum(1, keepdims=True)
    weights = np.zeros((V, J), dtype=np.float32)
    np.put_along_axis(weights, nn, w.astype(np.float32), 1)

    kintree = np.stack([np.array([2 ** 32 - 1 if p < 0 else p for p in parents], dtype=np.int64),
                        np.arange(J, dtype=np.int64)])
    model = dict(model_type=model_type, v_template=v_template, f=faces, shapedirs=shapedirs,
                 posedirs=posedirs, J_regressor=J_regressor, weights=weights,
                 kintree_table=kintree, num_betas=np.int64(num_betas),
                 extra_vids=np.array(SMPLX_EXTRA_VIDS if smplx else SMPL_EXTRA_VIDS, dtype=np.int64))
    if smplx:
        F = faces.shape[0]
        model['hands_componentsl'] = (rng.standard_normal((6, 45)) * 0.3).astype(np.float32)
        model['hands_componentsr'] = (rng.standard_normal((6, 45)) * 0.3).astype(np.float32)
        model['hands_meanl'] = (rng.standard_normal(45) * 0.1).astype(np.float32)
        model['hands_meanr'] = (rng.standard_normal(45) * 0.1).astype(np.float32)
        model['lmk_faces_idx'] = rng.choice(F, 51, replace=False).astype(np.int64)
        b = rng.uniform(0.05, 1.0, size=(51, 3))
        model['lmk_bary_coords'] = (b / b.sum(1, keepdims=True)).astype(np.float32)
        # contour landmarks slide over a small pool of faces as the head yaws
        pool = rng.choice(F, 17 * 6, replace=False).reshape(17, 6)
        dyn = np.zeros((79, 17), dtype=np.int64)
        for a in range(79):
            dyn[a] = pool[np.arange(17), np.minimum(5, (a + np.arange(17)) % 79 // 14)]
        model['dynamic_lmk_faces_idx'] = dyn
        b = rng.uniform(0.05, 1.0, size=(79, 17, 3))
        model['dynamic_lmk_bary_coords'] = (b / b.sum(-1, keepdims=True)).astype(np.float32)
    return model


def make_gmm(seed=0, num_gaussians=8, dim=69):
    """Synthetic stand-in for data/gmm_08.pkl (smplify/prior.py:119-133)."""
    rng = np.random.RandomState(seed + 101)
    means = rng.standard_normal((num_gaussians, dim)) * 0.2
    A = rng.standard_normal((num_gaussians, dim, dim)) * 0.1
    covars = A @ np.transpose(A, (0, 2, 1)) + 0.05 * np.eye(dim)
    w = rng.uniform(0.2, 1.0, size=num_gaussians)
    return dict(means=means, covars=covars, weights=w / w.sum())


def make_J_regressor_extra(num_verts=6890, rows=9, seed=0):
    """Synthetic stand-in for data/J_regressor_extra.npy (models/smpl.py:62)."""
    rng = np.random.RandomState(seed + 202)
    R = np.zeros((rows, num_verts), dtype=np.float32)
    for r in range(rows):
        nn = rng.choice(num_verts, 24, replace=False)
        w = rng.uniform(0.1, 1.0, size=24)
        R[r, nn] = (w / w.sum()).astype(np.float32)
    return R


def save_model_npz(model, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, **{k: v for k, v in model.items() if k != 'model_type'})


def write_data_dir(root, seed=0, model_types=('smpl', 'smplx')):
    """Lay out a ``data/`` folder the way the reference expects it (config.py:1-6,
    smplify/smplify.py:46,63): gmm_08.pkl, J_regressor_extra.npy, J_regressor_h36m.npy,
    smpl/SMPL_NEUTRAL.npz, smplx/SMPLX_NEUTRAL.npz."""
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, 'gmm_08.pkl'), 'wb') as f:
        pickle.dump(make_gmm(seed), f)
    np.save(os.path.join(root, 'J_regressor_extra.npy'), make_J_regressor_extra(seed=seed))
    np.save(os.path.join(root, 'J_regressor_h36m.npy'), make_J_regressor_extra(rows=17, seed=seed + 1))
    for mt in model_types:
        m = make_model(mt, seed)
        save_model_npz(m, os.path.join(root, mt, '%s_NEUTRAL.npz' % mt.upper()))
    if 'smpl' in model_types:                                  # config.SMIL_MODEL_DIR: the kid template (age='kid')
        os.makedirs(os.path.join(root, 'smil'), exist_ok=True)
        with open(os.path.join(root, 'smil', 'smil_web.pkl'), 'wb') as f:
            np.save(f, make_kid_template(seed))
    return root


def make_kid_template(seed=0):
    """A synthetic stand-in for the SMIL infant template (smplx ``kid_template_path``): the adult SMPL template shrunk
    and perturbed, [6890,3] float32 (smplx loads it with np.load, mean-centres it and uses its difference to the adult
    template as an 11th shape direction)."""
    vt = make_model('smpl', seed)['v_template']
    rng = np.random.RandomState(seed + 77)
    return (vt * np.array([0.55, 0.5, 0.6], np.float32) + 0.05 + rng.standard_normal(vt.shape).astype(np.float32) * 0.002).astype(np.float32)


# --- scene --------------------------------------------------------------------

def make_cameras(n_views, radius=1.2, focal=512.0, imsize=512, seed=0):
    """Ring of calibrated cameras looking at the origin (cf. utils/renderer.py:7-25).
    Returns c2ws [Nv,4,4] f32 and Ks [Nv,3,3] f32 (OpenCV convention: +z forward)."""
    rng = np.random.RandomState(seed + 303)
    c2ws, Ks = [], []
    for i in range(n_views):
        ang = 2 * np.pi * i / n_views + rng.uniform(-0.05, 0.05)
        eye = np.array([radius * np.sin(ang), rng.uniform(-0.1, 0.1), radius * np.cos(ang)])
        fwd = -eye / np.linalg.norm(eye)
        right = np.cross(np.array([0, -1.0, 0]), fwd)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
        c2ws.append(c2w)
        Ks.append(np.array([[focal, 0, imsize / 2], [0, focal, imsize / 2], [0, 0, 1.0]]))
    return np.stack(c2ws).astype(np.float32), np.stack(Ks).astype(np.float32)


def make_params(model_type, n_frames, seed=0, init_noise=True):
    """Ground-truth and initial (HMR stand-in) parameters for ``n_frames`` frames.
    Returns two dicts of [B,...] float32 arrays: keys betas[10], global_orient[3],
    body_pose[69|63]; gt additionally has transl[3], scale[1] and (smplx)
    left_hand_pose[6], right_hand_pose[6], leye_pose[3], reye_pose[3]."""
    nb = 69 if model_type == 'smpl' else 63
    rng = np.random.RandomState(1000 + seed)
    gt = dict(betas=rng.standard_normal((n_frames, 10)),
              global_orient=rng.standard_normal((n_frames, 3)) * 0.3,
              body_pose=rng.standard_normal((n_frames, nb)) * 0.2,
              transl=rng.standard_normal((n_frames, 3)) * 0.05,
              scale=1.0 + rng.standard_normal((n_frames, 1)) * 0.03)
    if model_type == 'smplx':
        gt['left_hand_pose'] = rng.standard_normal((n_frames, 6)) * 0.3
        gt['right_hand_pose'] = rng.standard_normal((n_frames, 6)) * 0.3
        gt['leye_pose'] = rng.standard_normal((n_frames, 3)) * 0.05
        gt['reye_pose'] = rng.standard_normal((n_frames, 3)) * 0.05
    init = dict(betas=gt['betas'] + (rng.standard_normal((n_frames, 10)) * 0.5 if init_noise else 0),
                global_orient=gt['global_orient'] + (rng.standard_normal((n_frames, 3)) * 0.1 if init_noise else 0),
                body_pose=gt['body_pose'] + (rng.standard_normal((n_frames, nb)) * 0.1 if init_noise else 0))
    f32 = lambda d: {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in d.items()}
    return f32(gt), f32(init)


def project_points(points_world, c2ws, Ks):
    """numpy pinhole projection used only to synthesise 2-D detections.
    points_world [B,K,3] -> uv [B,Nv,K,2]."""
    w2cs = np.linalg.inv(c2ws.astype(np.float64))
    pc = np.einsum('vij,bkj->bvki', w2cs[:, :3, :3], points_world.astype(np.float64)) + w2cs[:, None, :3, 3][None]
    p = np.einsum('vij,bvkj->bvki', Ks.astype(np.float64), pc)
    return p[..., :2] / p[..., 2:3]


def make_keypoints(joints_world, c2ws, Ks, seed=0, noise_px=2.0, drop=0.05):
    """Synthetic detections in *model joint order*: [B,Nv,K,3] (x, y, conf).
    ``joints_world`` are the scaled/translated model joints of the ground-truth
    parameters, computed by the caller (CUDA path or oracle)."""
    rng = np.random.RandomState(2000 + seed)
    uv = project_points(joints_world, c2ws, Ks)
    B, Nv, K, _ = uv.shape
    uv = uv + rng.standard_normal(uv.shape) * noise_px
    conf = rng.uniform(0.3, 1.0, size=(B, Nv, K, 1))
    kp = np.concatenate([uv, conf], -1)
    kp[rng.uniform(size=(B, Nv, K)) < drop] = 0.0
    return kp.astype(np.float32)


def keypoints_to_openpose(kp_frame, model_type):
    """One frame of packed keypoints [Nv,K,3] -> the list of per-view dicts that
    ``utils/io_utils.py:load_openpose`` returns and ``multiview_keypoint_loss``
    consumes (smplify/loss.py:160-181)."""
    out = []
    for v in range(kp_frame.shape[0]):
        k = kp_frame[v]
        d = {'pose': np.ascontiguousarray(k[:25])}
        if model_type == 'smplx':
            d['hand_left'] = np.ascontiguousarray(k[25:46])
            d['hand_right'] = np.ascontiguousarray(k[46:67])
            face = np.zeros((70, 3), dtype=k.dtype)
            face[FACE_MAPPING] = k[67:135]
            d['face'] = face
        out.append(d)
    return out


def openpose_to_keypoints(views, model_type):
    """Inverse of ``keypoints_to_openpose``: list of per-view dicts (or None for a
    view without detection) -> packed [Nv,K,3] in model joint order; missing views
    and missing parts get conf = 0, which contributes exactly 0 to loss and gradient."""
    K = 135 if model_type == 'smplx' else 25
    out = np.zeros((len(views), K, 3), dtype=np.float32)
    for v, d in enumerate(views):
        if d is None:
            continue
        out[v, :25] = np.asarray(d['pose'], dtype=np.float32)[:25]
        if model_type == 'smplx':
            if 'hand_left' in d:
                out[v, 25:46] = np.asarray(d['hand_left'], dtype=np.float32)
            if 'hand_right' in d:
                out[v, 46:67] = np.asarray(d['hand_right'], dtype=np.float32)
            if 'face' in d:
                out[v, 67:135] = np.asarray(d['face'], dtype=np.float32)[FACE_MAPPING]
    return out


def make_masks(verts_world, faces, c2ws, Ks, imsize=512):
    """Silhouette masks [Nv,imsize,imsize] uint8 (0 / 255) of a world-space mesh: every face rasterised with
    cv2.fillPoly in every view (stand-in for the segmentation masks of apps/genebody_fitting.py)."""
    import cv2
    v = np.asarray(verts_world, dtype=np.float64)
    out = np.zeros((len(c2ws), imsize, imsize), np.uint8)
    for i, (c2w, K) in enumerate(zip(c2ws, Ks)):
        w2c = np.linalg.inv(np.asarray(c2w, dtype=np.float64))
        pc = v @ w2c[:3, :3].T + w2c[:3, 3]
        uv = pc @ np.asarray(K, dtype=np.float64).T
        uv = uv[:, :2] / uv[:, 2:3]
        polys = np.round(uv[np.asarray(faces)]).astype(np.int32)
        cv2.fillPoly(out[i], list(polys), 255)
    return out
