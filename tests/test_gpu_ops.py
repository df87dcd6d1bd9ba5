"""GPU parity of the operator surface (models.smpl.SMPL, smplify.loss.*, smplify.prior.*) against the
oracle's CPU restatements, values and gradients."""
import numpy as np
import pytest
import torch

from bodyfitting_b200 import synthetic as syn
from oracle import fit_port as fp
from util import make_port, make_scene, perturbed_params, relerr

pytestmark = pytest.mark.gpu


def test_projection_gmof_reprojection_angle():
    from bodyfitting_b200.smplify import loss as L
    rng = np.random.RandomState(0)
    B, N = 3, 40
    pts = rng.randn(B, N, 3) * 0.2
    c2ws, Ks = syn.make_cameras(1)
    w2c = np.linalg.inv(c2ws[0].astype(np.float64))
    R, t, K = w2c[None, :3, :3], w2c[None, :3, 3], Ks[0]
    T = lambda a, g=False: torch.tensor(a, dtype=torch.float32, requires_grad=g)
    p_ref = T(pts, True)
    uv_ref = fp.project(p_ref, T(R), T(t), K)
    p_gpu = T(pts).cuda().requires_grad_(True)
    uv = L.perspective_projection(p_gpu, T(R).cuda(), T(t).cuda(), K)
    assert relerr(uv.detach().cpu().numpy(), uv_ref.detach().numpy()) < 1e-5
    wgt = rng.randn(B, N, 2).astype(np.float32)
    (uv_ref * T(wgt)).sum().backward()
    (uv * T(wgt).cuda()).sum().backward()
    assert relerr(p_gpu.grad.cpu().numpy(), p_ref.grad.numpy()) < 1e-4
    # per-frame rotation (bs,3,3)
    Rb = np.repeat(R, B, 0)
    tb = np.repeat(t, B, 0)
    uv2 = L.perspective_projection(T(pts).cuda(), T(Rb).cuda(), T(tb).cuda(), torch.tensor(K).cuda())
    assert relerr(uv2.cpu().numpy(), uv_ref.detach().numpy()) < 1e-5
    # gmof
    x = rng.randn(1000).astype(np.float32) * 300
    xr, xg = T(x, True), T(x).cuda().requires_grad_(True)
    fp.gmof(xr, 100.0).sum().backward()
    y = L.gmof(xg, 100.0)
    y.sum().backward()
    assert relerr(y.detach().cpu().numpy(), fp.gmof(T(x), 100.0).numpy()) < 1e-6
    assert relerr(xg.grad.cpu().numpy(), xr.grad.numpy()) < 1e-5
    # reprojection_loss with [N] and [N,1] confidences (the reference's broadcast)
    cord, gt, conf = rng.rand(21, 2) * 500, rng.rand(21, 2) * 500, rng.rand(21)
    for shape in ((21,), (21, 1)):
        cr, cg = T(cord, True), T(cord).cuda().requires_grad_(True)
        ref = fp.reprojection(cr, T(gt), T(conf.reshape(shape)), 0.5, 100.0)
        out = L.reprojection_loss(cg, T(gt).cuda(), T(conf.reshape(shape)).cuda(), 0.5, 100.0)
        assert out.shape == ref.shape
        assert relerr(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
        ref.sum().backward(); out.sum().backward()
        assert relerr(cg.grad.cpu().numpy(), cr.grad.numpy()) < 1e-4
    # angle prior
    pose = rng.randn(5, 69).astype(np.float32) * 0.3
    pr, pg = T(pose, True), T(pose).cuda().requires_grad_(True)
    a_ref, a = fp.angle_prior(pr), L.angle_prior(pg)
    assert relerr(a.detach().cpu().numpy(), a_ref.detach().numpy()) < 1e-6
    a_ref.sum().backward(); a.sum().backward()
    assert relerr(pg.grad.cpu().numpy(), pr.grad.numpy()) < 1e-6


def test_gmm_prior_module(assets):
    from bodyfitting_b200.smplify.prior import MaxMixturePrior, SMPLifyAnglePrior, L2Prior, create_prior
    gmm = assets('gmm')
    ref = fp.GMMPrior(gmm)
    prior = MaxMixturePrior(gmm=gmm)
    rng = np.random.RandomState(1)
    for D in (69, 63):
        pose = (rng.randn(70, D) * 0.3).astype(np.float32)
        pr = torch.tensor(np.pad(pose, ((0, 0), (0, 69 - D))), requires_grad=True)
        pg = torch.tensor(pose).cuda().requires_grad_(True)
        a, b = ref(pr, None), prior(pg, None)
        assert relerr(b.detach().cpu().numpy(), a.detach().numpy()) < 1e-5
        a.sum().backward(); b.sum().backward()
        assert relerr(pg.grad.cpu().numpy(), pr.grad.numpy()[:, :D]) < 1e-4
    assert isinstance(create_prior('angle'), SMPLifyAnglePrior) and isinstance(create_prior('l2'), L2Prior)
    with pytest.raises(ValueError):
        create_prior('nope')


def test_smpl_module_forward_backward(assets):
    """models.smpl.SMPL: vertices [B,6890,3], joints [B,49,3], joints_ori [B,45,3] + autograd."""
    from bodyfitting_b200.models.smpl import SMPL
    B = 4
    port = make_port(assets, 'smpl', dtype=torch.float64)
    m = SMPL(model_data=assets('smpl'), J_regressor_extra=assets('jx'))
    p = perturbed_params('smpl', B, seed=13)
    tr = {k: torch.tensor(p[k], dtype=torch.float64, requires_grad=True) for k in ('global_orient', 'body_pose', 'betas')}
    tg = {k: torch.tensor(p[k]).cuda().requires_grad_(True) for k in ('global_orient', 'body_pose', 'betas')}
    ref = port.model(**tr, return_full_pose=True)
    out = m(**tg, return_full_pose=True)
    assert out.vertices.shape == (B, 6890, 3) and out.joints.shape == (B, 49, 3) and out.joints_ori.shape == (B, 45, 3)
    assert relerr(out.vertices.detach().cpu().numpy(), ref.vertices.detach().numpy()) < 1e-5
    assert relerr(out.joints.detach().cpu().numpy(), ref.joints.detach().numpy()) < 1e-5
    assert relerr(out.joints_ori.detach().cpu().numpy(), ref.joints_ori.detach().numpy()) < 1e-5
    assert relerr(out.full_pose.cpu().numpy(), ref.full_pose.detach().numpy()) < 1e-6
    rng = np.random.RandomState(2)
    wv, wj, wo = rng.randn(B, 6890, 3), rng.randn(B, 49, 3), rng.randn(B, 45, 3)
    Td = lambda a: torch.tensor(a, dtype=torch.float64)
    ((ref.vertices * Td(wv)).sum() + (ref.joints * Td(wj)).sum() + (ref.joints_ori * Td(wo)).sum()).backward()
    Tf = lambda a: torch.tensor(a, dtype=torch.float32).cuda()
    ((out.vertices * Tf(wv)).sum() + (out.joints * Tf(wj)).sum() + (out.joints_ori * Tf(wo)).sum()).backward()
    for k in tr:
        assert relerr(tg[k].grad.cpu().numpy(), tr[k].grad.numpy()) < 1e-4, k


def test_multiview_keypoint_loss_composed(assets):
    """The stand-alone objective (SMPL-X module + loss.multiview_keypoint_loss + prior) equals the oracle's
    reference-style objective, value and gradients, for one frame given as OpenPose dicts."""
    from bodyfitting_b200.models.smpl import create_smplx
    from bodyfitting_b200.smplify.loss import multiview_keypoint_loss
    from bodyfitting_b200.smplify.prior import MaxMixturePrior
    mt, nv = 'smplx', 8
    port = make_port(assets, mt)
    sc = make_scene(port, mt, 1, nv, seed=17)
    views = syn.keypoints_to_openpose(sc['kp'][0], mt)
    p = perturbed_params(mt, 1, seed=18)
    # oracle
    pr = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    full = dict(pr)
    full['jaw_pose'] = torch.zeros(1, 1, 3)
    full['leye_pose'] = pr['leye_pose'].view(1, 1, 3); full['reye_pose'] = pr['reye_pose'].view(1, 1, 3)
    o = port.forward_model(full)
    jw = (o.joints + pr['global_transl']) * pr['body_scale'] * 0.3
    w2cs = torch.inverse(torch.tensor(np.array(sc['c2ws'])))
    ref, terms = fp.keypoint_objective(w2cs, [np.asarray(k) for k in sc['Ks']], views, jw, pr['body_pose'], pr['betas'],
                                       port.prior, 512, True)
    ref.backward()
    # ours
    pg = {k: torch.tensor(v).cuda().requires_grad_(True) for k, v in p.items()}
    model = create_smplx(model_data=assets(mt))
    out = model(global_orient=pg['global_orient'], body_pose=pg['body_pose'], betas=pg['betas'], leye_pose=pg['leye_pose'],
                reye_pose=pg['reye_pose'], left_hand_pose=pg['left_hand_pose'], right_hand_pose=pg['right_hand_pose'])
    jwg = (out.joints + pg['global_transl'][:, None]) * pg['body_scale'][:, None] * 0.3
    prior = MaxMixturePrior(gmm=assets('gmm'))
    tot, losses = multiview_keypoint_loss(w2cs, list(sc['Ks']), views, jwg, pg['body_pose'], pg['betas'], list(range(nv)),
                                          prior, imsize=512, use_hand_face=True)
    tot.backward()
    assert relerr(float(tot), float(ref)) < 1e-5
    for k in ('reprojection_loss', 'pose_prior_loss', 'angle_prior_loss', 'shape_prior_loss'):
        assert relerr(losses[k], terms[k].detach().numpy()) < 1e-5, k
    for k in p:
        assert relerr(pg[k].grad.cpu().numpy(), pr[k].grad.numpy()) < 2e-4, k


def test_scan_and_normal_terms_under_reference_names():
    """smplify.loss.{point_cloud_loss_mesh_grid, normal_loss_mesh_grid, normal_laplacian_smoothness} and
    utils.io_utils.compute_normal_torch (reference: smplify/loss.py:233-242,260-288, utils/io_utils.py:410-428): values and
    gradients against autograd of the oracle's torch restatements (fp64), composed as the SMPL+D loop composes them
    (smplify/smplify.py:236-245)."""
    from bodyfitting_b200.smplify import loss as L
    from bodyfitting_b200.utils.io_utils import compute_normal_torch
    from bodyfitting_b200.utils.mesh_grid_searcher import MeshGridSearcher
    from oracle import geometry_port as gp
    body, faces = syn.make_template(1200, 3)
    scan, sfaces = syn.make_template(2500, 11)
    scan = (scan * 1.03).astype(np.float32)
    rng = np.random.RandomState(1)
    body = (body + rng.randn(*body.shape) * 0.003).astype(np.float32)
    tris = scan[sfaces]
    fn = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]).astype(np.float32)
    searcher = MeshGridSearcher(scan, sfaces.astype(np.int32))
    # ---- ours
    v = torch.tensor(body, device='cuda', requires_grad=True)
    f = torch.tensor(faces.astype(np.int64), device='cuda')
    norms = compute_normal_torch(v, f)
    icp = L.point_cloud_loss_mesh_grid(searcher, v)
    nl = L.normal_loss_mesh_grid(searcher, v, torch.tensor(fn, device='cuda'), norms)
    sm = L.normal_laplacian_smoothness(norms, f)
    (icp + (nl + sm) * 0.3).backward()
    # ---- oracle (fp64, exact brute-force closest points)
    vr = torch.tensor(body, dtype=torch.float64, requires_grad=True)
    fr = torch.tensor(faces.astype(np.int64))
    nr = gp.compute_normal_torch(vr, fr)
    cp, cf, _ = gp.closest_points_bruteforce(body.astype(np.float64), scan.astype(np.float64), sfaces.astype(np.int64))
    icp_r = torch.norm(vr - torch.tensor(cp), p=2)
    nl_r = torch.mean(1 - torch.sum(torch.tensor(fn, dtype=torch.float64)[torch.tensor(cf)] * nr, dim=-1))
    sm_r = gp.normal_laplacian_smoothness(nr, fr)
    (icp_r + (nl_r + sm_r) * 0.3).backward()
    print('normals', relerr(norms.detach().cpu().numpy(), nr.detach().numpy()), 'icp', abs(float(icp) - float(icp_r)) / float(icp_r),
          'normal', abs(float(nl) - float(nl_r)) / abs(float(nl_r)), 'smooth', abs(float(sm) - float(sm_r)) / float(sm_r),
          'grad', relerr(v.grad.cpu().numpy(), vr.grad.numpy()))
    assert relerr(norms.detach().cpu().numpy(), nr.detach().numpy()) < 1e-5
    assert abs(float(icp) - float(icp_r)) / float(icp_r) < 1e-5
    assert abs(float(nl) - float(nl_r)) / abs(float(nl_r)) < 1e-4          # closest-face ties may differ on a handful of vertices
    assert abs(float(sm) - float(sm_r)) / float(sm_r) < 1e-5
    assert relerr(v.grad.cpu().numpy(), vr.grad.numpy()) < 1e-3


def test_mask_loss_under_reference_names(assets):
    """smplify.loss.{extract_countours, multview_mask_loss} (reference spelling, smplify/loss.py:73-130) on a mask with TWO
    blobs in one view: the contour kept is the one the reference keeps (the first OpenCV returns), value and vertex
    gradient match the oracle's restatement of the reference function."""
    from bodyfitting_b200.smplify import loss as L
    mt, nv = 'smpl', 4
    port = make_port(assets, mt)
    sc = make_scene(port, mt, 1, nv, seed=23)
    init = sc['init']
    ev = port.loss_and_grads(dict(global_orient=init['global_orient'], body_pose=init['body_pose'], betas=init['betas']),
                             sc['c2ws'], sc['Ks'], np.zeros((1, nv, 25, 3), np.float32))
    masks = syn.make_masks(ev['vertices'][0], port.faces, sc['c2ws'], sc['Ks'])[[1, 3]]
    masks[0, 20:60, 30:90] = 255                                       # a second, smaller blob
    masks[1, 440:500, 400:480] = 255
    mk = torch.as_tensor((masks > 128).astype(np.float32))
    w2cs = torch.inverse(torch.as_tensor(np.array(sc['c2ws']), dtype=torch.float32))[[1, 3]]
    Kt = torch.as_tensor(np.array(sc['Ks']), dtype=torch.float32)[[1, 3]]
    cont_ref = fp.extract_contours(mk)
    cont = L.extract_countours(mk.cuda())
    assert all(c.shape == r.shape and np.array_equal(c.cpu().numpy(), r.numpy()) for c, r in zip(cont, cont_ref))
    vr = torch.tensor(ev['vertices'][:1], requires_grad=True)
    val_r = fp.mask_objective(cont_ref, mk, vr, list(w2cs), list(Kt), 512, exact_cdist=True)
    val_r.backward()
    vg = torch.tensor(ev['vertices'][:1], device='cuda', requires_grad=True)
    val = L.multview_mask_loss(cont, mk.cuda(), vg, port.faces[None], list(w2cs.cuda()), list(Kt.cuda()), [1, 3], imsize=512)
    val.backward()
    print('mask loss', float(val), float(val_r), 'grad rel', relerr(vg.grad.cpu().numpy(), vr.grad.numpy()))
    assert abs(float(val) - float(val_r)) / float(val_r) < 2e-5
    assert relerr(vg.grad.cpu().numpy(), vr.grad.numpy()) < 1e-4


def test_get_joints_h36m(assets):
    """models/smpl.py:85-87: J_regressor_h36m @ vertices, value and gradient against the dense einsum."""
    from bodyfitting_b200.models.smpl import SMPL
    rng = np.random.RandomState(4)
    Wm = np.zeros((17, 6890), np.float32)
    for r in range(17):
        cols = rng.choice(6890, 30, replace=False)
        Wm[r, cols] = rng.dirichlet(np.ones(30)).astype(np.float32)
    smpl = SMPL(model_data=assets('smpl'), J_regressor_extra=assets('jx'), J_regressor_h36m=Wm)
    v = torch.tensor(rng.randn(3, 6890, 3).astype(np.float32), device='cuda', requires_grad=True)
    j = smpl.get_joints_h36m(v)
    ref_v = torch.tensor(v.detach().cpu().numpy(), dtype=torch.float64, requires_grad=True)
    jr = torch.einsum('bik,ji->bjk', ref_v, torch.tensor(Wm, dtype=torch.float64))
    assert j.shape == (3, 17, 3) and relerr(j.detach().cpu().numpy(), jr.detach().numpy()) < 1e-6
    g = rng.randn(3, 17, 3).astype(np.float32)
    (j * torch.tensor(g, device='cuda')).sum().backward()
    (jr * torch.tensor(g, dtype=torch.float64)).sum().backward()
    assert relerr(v.grad.cpu().numpy(), ref_v.grad.numpy()) < 1e-6
