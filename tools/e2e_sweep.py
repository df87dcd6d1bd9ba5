"""End-to-end (host numpy -> host numpy) time of a 10,000-frame fit for part layouts: parts x taper.
    python tools/e2e_sweep.py [B]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.smplify.smplify import SMPLify

B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
model, gmm = syn.make_model('smplx', 0), syn.make_gmm(0)
wl = pin = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for parts in (3, 4, 5, 6):
    for taper in (0.5, 1.0, 2.0):
        os.environ['BODYFIT_TAPER'] = str(taper)
        fit = SMPLify(smpl_type='smplx', num_iters=100, gender='neutral', model_data=model, gmm=gmm, concurrent_parts=parts,
                      concurrent_min_part=512)
        if wl is None:
            wl = bench.build_workload(fit.model, 10000, 100, 0, B)
            pin = {k: torch.from_numpy(wl[k]).pin_memory() for k in ('kp', 'init_pose', 'init_betas')}
        args = ((pin['init_betas'].numpy(), pin['init_pose'].numpy()), list(wl['c2ws']), list(wl['Ks']), pin['kp'].numpy(), None)
        for _ in range(2):
            fit(*args, use_frames=list(range(8)), imsize=512)
        ws = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter(); fit(*args, use_frames=list(range(8)), imsize=512); ws.append(1e3 * (time.perf_counter() - t0))
        sess = fit.session(B, 8, 512, True)
        print(json.dumps({'B': B, 'parts': parts, 'taper': taper, 'ranges': [hi - lo for lo, hi in getattr(sess, 'ranges', [(0, B)])],
                          'e2e_ms': round(float(np.median(ws)), 2), 'e2e_frames_per_s': round(B / float(np.median(ws)) * 1e3)}), flush=True)
        del fit, sess
        torch.cuda.empty_cache()
