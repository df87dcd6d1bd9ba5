"""Where the end-to-end time of SMPLify.__call__ goes (host packing, H2D, fit, D2H)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import pack_cameras, pack_keypoints
from bodyfitting_b200.smplify.smplify import SMPLify

F = 10000
fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=syn.make_model('smplx', 0), gmm=syn.make_gmm(0), concurrent_parts=1)
pm = fit.model
wl = bench.build_workload(pm, F, seed=100)
args = ((wl['init_betas'], wl['init_pose']), list(wl['c2ws']), list(wl['Ks']), wl['kp'], None)
for _ in range(2):
    fit(*args, imsize=512)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = T()
    ib, ip, kp = fit._pack_inputs(args[0], args[3]); t1 = T()
    sess = fit.session(F, 8, 512, True); t2 = T()
    kp_dev = fit._h2d('kp', kp); t3 = T()
    poses_dev = fit._h2d('poses', ip); betas_dev = fit._h2d('betas', ib)
    cams = fit._h2d('cams', torch.from_numpy(pack_cameras(args[1], args[2]))); t4 = T()
    sess.set_inputs(pack_keypoints(kp_dev, True), cams); t5 = T()
    theta0 = pm.pack_theta(poses_dev[:, :3], poses_dev[:, 3:3 + pm.nbody], betas_dev); t6 = T()
    sess.run(theta0); t7 = T()
    out = sess.results(); t8 = T()
    res = fit._d2h(out); t9 = T()
    print('pack_inputs %.1f | session %.1f | h2d kp %.1f | h2d rest+cams %.1f | pack_kp %.1f | theta %.1f | fit %.1f | results %.1f | d2h %.1f | total %.1f ms'
          % tuple(1e3 * x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6, t8-t7, t9-t8, t9-t0)))
    print('d2h bytes', fit.d2h_bytes, 'GB/s', fit.d2h_bytes / (t9 - t8) / 1e9)
