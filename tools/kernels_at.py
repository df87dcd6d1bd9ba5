"""Per-kernel CUDA-event times of one fit iteration at a given batch size (one batch, one stream, direct launches).
    python tools/kernels_at.py 1250 2500 10000"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.engine import FitSession, pack_cameras
from bodyfitting_b200.model import PreparedModel

pm = PreparedModel('smplx', syn.make_model('smplx', 0), gmm=syn.make_gmm(0), device='cuda')
hbm, _ = bench.peaks()
for B in [int(a) for a in sys.argv[1:]] or [1250, 10000]:
    wl = bench.build_workload(pm, 10000, 100, 0, B)
    s = FitSession(pm, B, 8, 100, graph=False)
    T = lambda k: torch.from_numpy(wl[k]).cuda()
    s.load_inputs(T('kp'), torch.from_numpy(pack_cameras(wl['c2ws'], wl['Ks'])).cuda(), T('init_pose'), T('init_betas'))
    s.run()
    k = bench.kernel_breakdown(pm, s, B, hbm)
    print(json.dumps({'B': B, 'iter_ms_sum': sum(x['ms'] * x['launches_per_iteration'] for x in k),
                      'kernels': {x['kernel']: round(x['ms'] * 1e3, 1) for x in k}}), flush=True)
    del s
    torch.cuda.empty_cache()
