"""GPU parity of the scan path: keypoints + point-to-scan main loop (use_mesh=True) and the SMPL+D
displacement loop, against the oracle restatements with an exact fp64 brute-force closest point."""
import numpy as np
import pytest
import torch

from bodyfitting_b200 import synthetic as syn
from oracle import geometry_port as gp
from util import make_port, make_scene, relerr

pytestmark = pytest.mark.gpu


def _scan(scale=0.3, n=1000, seed=7):
    v, f = syn.make_template(n, seed)
    rng = np.random.RandomState(seed)
    v = (v * np.array([1.05, 1.0, 1.1]) + rng.randn(*v.shape) * 0.004) * scale
    return v.astype(np.float64), f.astype(np.int64)


def test_smpld_gradient_and_loop(assets):
    from bodyfitting_b200.smplify.smpld import DisplacementFitter
    from bodyfitting_b200.utils.mesh_grid_searcher import MeshGridSearcher
    body_v, body_f = syn.make_template(2000, 3)
    body_v = (body_v * 0.3).astype(np.float32)
    sv, sf = _scan()
    tris = sv[sf]
    fn = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
    cs = float((sv.max(0) - sv.min(0))[1]) / 1.7
    searcher = MeshGridSearcher(sv.astype(np.float32), sf.astype(np.int32))
    fit = DisplacementFitter(searcher, fn, body_f, len(body_v), cs)
    # one step: loss terms and gradient w.r.t. the displacement vs autograd (fp64, exact closest points)
    disp1, tr1 = fit.run(torch.from_numpy(body_v).cuda(), 1, return_grad=True)
    g = fit.buffers['grad'].cpu().numpy()
    bv = torch.tensor(body_v, dtype=torch.float64)
    d = torch.zeros_like(bv, requires_grad=True)
    faces = torch.tensor(body_f.astype(np.int64))
    norms = gp.compute_normal_torch(bv + d, faces)
    cp, cf, _ = gp.closest_points_bruteforce(body_v, sv, sf)
    icp = torch.norm((bv + d) - torch.tensor(cp), p=2)
    nl = torch.mean(1 - torch.sum(torch.tensor(fn)[torch.tensor(cf)] * norms, dim=-1))
    sm = gp.normal_laplacian_smoothness(norms, faces)
    loss = icp + (nl + sm) * cs * 0.1
    loss.backward()
    t = tr1.cpu().numpy()[0]
    print('smpld terms', t, [float(icp), float(nl), float(sm), float(loss)])
    assert relerr(t, [float(icp), float(nl), float(sm), float(loss)]) < 1e-5
    print('smpld grad rel err', relerr(g, d.grad.numpy()))
    assert relerr(g, d.grad.numpy()) < 1e-4
    # loop
    N = 10
    disp, tr = fit.run(torch.from_numpy(body_v).cuda(), N)
    rdisp, rtr = gp.smpld_loop(body_v, body_f, sv, sf, cs, N)
    # Adam with lr 5e-2 moves every coordinate by ~lr whatever the size of its gradient, so coordinates whose
    # gradient is ~0 flip sign between any two floating-point evaluations: the yardstick is the drift of the
    # oracle itself between fp32 and fp64
    rdisp32, rtr32 = gp.smpld_loop(body_v, body_f, sv, sf, cs, N, dtype=torch.float32)
    rel = np.abs(tr.cpu().numpy() - rtr) / np.abs(rtr)
    rel32 = np.abs(rtr32 - rtr) / np.abs(rtr)
    print('smpld loop: trace max rel', rel.max(), '(fp32 oracle vs fp64:', rel32.max(), ') icp', rtr[0, 0], '->', rtr[-1, 0])
    assert rel[:3].max() < 1e-4
    assert rel.max() < max(1e-3, 5 * rel32.max())
    dd = np.abs(disp.cpu().numpy() - rdisp)
    dd32 = np.abs(rdisp32 - rdisp)
    print('   disp fraction within 1e-3: ours %.5f, fp32 oracle %.5f' % ((dd < 1e-3).mean(), (dd32 < 1e-3).mean()))
    assert (dd < 1e-3).mean() > min(0.999, (dd32 < 1e-3).mean() - 0.01)


def test_fit_to_scan_main_loop(assets):
    from bodyfitting_b200.smplify.smplify import SMPLify
    mt, nv, B, N = 'smpl', 4, 2, 12
    port = make_port(assets, mt)
    sc = make_scene(port, mt, B, nv, seed=23)
    sv, sf = _scan()
    ref, trace, pcs = port.fit_batched_scan(sc['init_betas'], sc['init_pose'], sc['c2ws'], sc['Ks'], sc['kp'], sv, sf, num_iters=N)
    fit = SMPLify(smpl_type=mt, num_iters=N, gender='neutral', model_data=assets(mt), gmm=assets('gmm'),
                  J_regressor_extra=assets('jx'))
    out = fit((sc['init_betas'], sc['init_pose']), list(sc['c2ws']), list(sc['Ks']), sc['kp'], None, imsize=512,
              use_mesh=True, meshfile=(sv, sf))
    tr = fit.last_trace.cpu().numpy()
    rel = np.abs(tr - trace) / np.abs(trace)
    print('scan fit: loss trace max rel', rel.max(), 'pc', pcs[0], pcs[-1], 'ours', fit.last_pc_loss.cpu().numpy())
    assert rel.max() < 1e-4
    assert relerr(fit.last_pc_loss.cpu().numpy(), pcs[-1]) < 1e-4
    for k in ('pose', 'betas', 'global_orient', 'global_transl', 'scale'):
        d = np.abs(np.asarray(out[k]).reshape(B, -1) - np.asarray(ref[k]).reshape(B, -1)).max()
        print('   %-14s %.3e' % (k, d))
        assert d < 1e-3, k
    assert relerr(out['vertices'], ref['vertices']) < 1e-3


def test_fit_to_scan_with_displacement_output(assets, tmp_path):
    """reference call form: meshfile = OBJ path, displacement=True -> 'displacement' [V,3] in the result dict"""
    from bodyfitting_b200.smplify.smplify import SMPLify
    from bodyfitting_b200.utils.io_utils import load_obj_mesh, save_obj_mesh
    mt, nv, N = 'smpl', 4, 6
    port = make_port(assets, mt)
    sc = make_scene(port, mt, 1, nv, seed=29)
    sv, sf = _scan()
    fn = str(tmp_path / 'scan.obj')
    save_obj_mesh(fn, sv, sf)
    v2, f2 = load_obj_mesh(fn)
    assert np.abs(v2 - sv).max() < 1e-4 and np.array_equal(f2, sf)
    fit = SMPLify(smpl_type=mt, num_iters=N, gender='neutral', model_data=assets(mt), gmm=assets('gmm'),
                  J_regressor_extra=assets('jx'))
    out = fit((sc['init_betas'], sc['init_pose']), list(sc['c2ws']), list(sc['Ks']),
              syn.keypoints_to_openpose(sc['kp'][0], mt), None, use_frames=list(range(nv)), imsize=512,
              use_mesh=True, meshfile=fn, displacement=True)
    assert out['displacement'].shape == (6890, 3) and out['vertices'].shape == (6890, 3)
    assert np.isfinite(out['displacement']).all() and np.abs(out['displacement']).max() > 0
    tr = fit.last_disp_trace.cpu().numpy()
    assert np.isfinite(tr).all() and tr.shape == (N, 4)
