"""Silhouette term (bf_mask_loss) timing: SMPL-X, B frames x 4 mask views of 512x512, per-kernel and total, plus the
all-vertex iteration it lives in (pose fwd, blend + skin, keypoint loss, mask term, dense backward, pose backward + Adam)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FrameBuffers, pack_cameras, pack_keypoints
from bodyfitting_b200.model import PreparedModel
from bodyfitting_b200.smplify.mask import SilhouetteTerm

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mt, nv, mask_frames = 'smplx', 8, [0, 2, 4, 6]
pm = PreparedModel(mt, syn.make_model(mt, 0), gmm=syn.make_gmm(0), device='cuda')
wl = bench.build_workload(pm, B, seed=3)
c2ws, Ks = wl['c2ws'], wl['Ks']
fb = FrameBuffers(pm, B, full=True, Nv=nv, n_trace=4)
fb.bind('kp', pack_keypoints(torch.from_numpy(wl['kp']).cuda(), True))
cams = pack_cameras(list(c2ws), list(Ks))
fb.bind('cams', torch.from_numpy(cams).cuda())
poses = torch.from_numpy(wl['init_pose']).cuda()
fb.t['theta'].copy_(pm.pack_theta(poses[:, :3], poses[:, 3:3 + pm.nbody], torch.from_numpy(wl['init_betas']).cuda()))
fb.call('bf_pose_forward'); fb.call('bf_skin_forward', 1)
# masks: silhouettes of the initial meshes (any mask works for timing), rasterised on the host
verts = fb.t['verts'].view(B, pm.V, 3).cpu().numpy() * 0.3
masks = np.stack([syn.make_masks(verts[b], pm.faces, c2ws, Ks)[mask_frames] for b in range(B)])
sil = SilhouetteTerm(pm, masks, cams[mask_frames], imsize=512)

def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))

def iteration(with_mask):
    fb.call('bf_pose_forward'); fb.call('bf_skin_forward', 1); fb.call('bf_keypoint_loss', 1)
    if with_mask:
        sil.add(fb, 5.0)
    fb.call('bf_gmm_prior'); fb.call('bf_skin_backward', 1); fb.call('bf_pose_backward', 1 | 4)

out = {'frames': B, 'mask_views': len(mask_frames), 'contour_pixels_total': int(sil.struct.total), 'sampled_vertices': int(sil.struct.Nq),
       'mask_term_ms': ev(lambda: sil.add(fb, 5.0)), 'iteration_ms_without': ev(lambda: iteration(False)),
       'iteration_ms_with': ev(lambda: iteration(True))}
out['pairs_per_s'] = out['contour_pixels_total'] * out['sampled_vertices'] / out['mask_term_ms'] / 1e-3
print(json.dumps(out))
