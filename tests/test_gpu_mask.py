"""GPU parity of the silhouette term (use_mask=True; bf_mask_loss, include/bodyfit_b200_mask.h) against the oracle's
restatement of smplify/loss.py:73-130 (oracle/fit_port.py::mask_objective, itself pinned bit-exactly against the
verbatim reference in tests/test_oracle.py).  The kernel measures contour distances directly, so the oracle is run
with exact_cdist=True (torch's default cdist uses a matmul expansion whose fp32 cancellation can flip arg-mins)."""
import numpy as np
import pytest
import torch

from bodyfitting_b200 import synthetic as syn
from oracle import fit_port as fp
from util import gt_param_dict, make_port, make_scene, relerr

pytestmark = pytest.mark.gpu


def _scene_with_masks(assets, mt, B, nv, mask_frames, seed):
    port = make_port(assets, mt)
    sc = make_scene(port, mt, B, nv, seed=seed)
    K = 135 if mt == 'smplx' else 25
    ev = port.loss_and_grads(gt_param_dict(sc['gt'], mt), sc['c2ws'], sc['Ks'], np.zeros((B, nv, K, 3), np.float32))
    masks = np.stack([syn.make_masks(ev['vertices'][b], port.faces, sc['c2ws'], sc['Ks'])[mask_frames] for b in range(B)])
    return port, sc, masks


@pytest.mark.parametrize('mt', ['smpl', 'smplx'])
def test_mask_term_value_and_vertex_gradient(assets, mt):
    """One evaluation at the initial parameters: per-frame value and d/d(world vertices) against autograd of the oracle."""
    from bodyfitting_b200.engine import FrameBuffers, pack_cameras
    from bodyfitting_b200.model import PreparedModel
    from bodyfitting_b200.smplify.mask import SilhouetteTerm, extract_contours
    B, nv, mask_frames = 2, 4, [0, 2]
    port, sc, masks = _scene_with_masks(assets, mt, B, nv, mask_frames, seed=21)
    pm = PreparedModel(mt, assets(mt), gmm=assets('gmm'), J_regressor_extra=assets('jx'), device='cuda')
    init = sc['init']
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    fb = FrameBuffers(pm, B, full=True, Nv=nv)
    fb.t['theta'].copy_(pm.pack_theta(T(init['global_orient']), T(init['body_pose']), T(init['betas'])))
    fb.call('bf_pose_forward')
    fb.call('bf_skin_forward', 1)
    fb.t['loss'].zero_(); fb.t['grad'].zero_(); fb.t['dverts'].zero_()
    cams = pack_cameras(list(sc['c2ws']), list(sc['Ks']))
    sil = SilhouetteTerm(pm, masks, cams[mask_frames], imsize=512)
    sil.add(fb, 1.0)
    torch.cuda.synchronize()
    # oracle: the same world vertices (from the oracle's own forward), autograd wrt them
    p = {k: torch.tensor(v) for k, v in dict(global_orient=init['global_orient'], body_pose=init['body_pose'], betas=init['betas']).items()}
    ev = port.loss_and_grads({k: v.numpy() for k, v in p.items()}, sc['c2ws'], sc['Ks'],
                             np.zeros((B, nv, pm.K_used, 3), np.float32))
    w2cs = torch.inverse(torch.as_tensor(np.array(sc['c2ws']), dtype=torch.float32))
    Kt = torch.as_tensor(np.array(sc['Ks']), dtype=torch.float32)
    mk = torch.as_tensor((masks > 128).astype(np.float32))
    vals, grads = [], []
    for b in range(B):
        v = torch.tensor(ev['vertices'][b:b + 1], requires_grad=True)
        cont = fp.extract_contours(mk[b])
        ours = extract_contours(masks[b] > 128)
        assert all(np.array_equal(c.numpy().reshape(-1, 2), o) for c, o in zip(cont, ours))      # same contours as the oracle's
        val = fp.mask_objective(cont, mk[b], v, [w2cs[f] for f in mask_frames], [Kt[f] for f in mask_frames], 512, exact_cdist=True)
        val.backward()
        vals.append(float(val)); grads.append(v.grad[0, ::4].numpy())
    got = sil.mask_loss.cpu().numpy()
    print(mt, 'mask term', got, 'oracle', vals)
    assert relerr(got, np.array(vals)) < 2e-5
    assert relerr(fb.t['loss'].cpu().numpy(), np.array(vals)) < 2e-5
    g = sil.dPw.cpu().numpy()
    gref = np.stack(grads)
    print(mt, 'vertex gradient rel err', relerr(g, gref))
    assert relerr(g, gref) < 1e-4
    # chained into the model-space vertex gradient: dverts[::4] = dPw * scale * constant_scale, other vertices untouched
    dv = fb.t['dverts'].view(B, -1, 3).cpu().numpy()
    assert relerr(dv[:, ::4], gref * fb.struct.constant_scale) < 1e-4 and not dv[:, 1::4].any()


def test_fit_with_masks_trajectory(assets):
    """use_mask=True through SMPLify.__call__: per-iteration loss against the oracle's batched loop.  The term switches
    arg-mins / the 10x outside penalty on sub-pixel changes, so long trajectories are compared loosely."""
    from bodyfitting_b200.smplify.smplify import SMPLify
    mt, B, nv, N, mask_frames = 'smpl', 2, 4, 9, [1, 3]
    port, sc, masks = _scene_with_masks(assets, mt, B, nv, mask_frames, seed=5)
    ref, trace, mls = port.fit_batched_mask(sc['init_betas'], sc['init_pose'], sc['c2ws'], sc['Ks'], sc['kp'], masks, mask_frames,
                                            num_iters=N, exact_cdist=True)
    fit = SMPLify(smpl_type=mt, num_iters=N, gender='neutral', model_data=assets(mt), gmm=assets('gmm'), J_regressor_extra=assets('jx'))
    out = fit((sc['init_betas'], sc['init_pose']), list(sc['c2ws']), list(sc['Ks']), sc['kp'], None, use_mask=True, masks=masks,
              use_frames=list(range(nv)), mask_frames=mask_frames, imsize=512)
    tr = fit.last_trace.cpu().numpy()
    rel = np.abs(tr - trace) / np.abs(trace)
    print('mask fit: per-iteration rel err', rel.max(1))
    first = N // 3 + 2                                                   # keypoint-only iterations + the first mask iteration
    assert rel[:first].max() < 1e-4
    assert rel.max() < 5e-2
    assert (tr[N // 3 + 1:] > 2 * tr[:N // 3 + 1].min()).all()            # the term is active from iteration N//3+1 on
    assert np.abs(np.array(out['pose']) - ref['pose']).max() < 5e-3
    assert out['vertices'].shape == (B, 6890, 3)
    # single frame, reference-style call (masks [Nm,H,W]) squeezes the batch dimension
    one = fit((sc['init_betas'][:1], sc['init_pose'][:1]), list(sc['c2ws']), list(sc['Ks']), sc['kp'][:1], None, use_mask=True,
              masks=list(masks[0]), use_frames=list(range(nv)), mask_frames=mask_frames, imsize=512)
    assert one['vertices'].shape == (6890, 3)
    assert np.array_equal(one['pose'], np.array(out['pose'])[0])
