"""Grid kernels on BASELINE config 5 sizes: 100k-vertex scan (200k faces), 10,475 closest-point queries, 100k inside / ray
queries.  Times bf_grid_nearest / bf_grid_inside / bf_grid_intersects_any and the grid build, and -- when oracle/_ref was built -- the reference's own
mesh_grid kernel compiled for sm_100a on the same GPU (the on-box bar for this path)."""
import glob, importlib.util, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.utils.mesh_grid_searcher import MeshGridSearcher

def ev_time(fn, reps=20):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))

v, f = syn.make_template(100000, 9)
v = (v * 0.6).astype(np.float32); f = f.astype(np.int32)
body, _ = syn.make_template(10475, 0)
rng = np.random.RandomState(0)
q = (body * 0.6 * 1.02 + rng.randn(*body.shape) * 0.004).astype(np.float32)
MeshGridSearcher(v[:3000], f[:10])                       # warm-up (library load, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter(); s = MeshGridSearcher(v, f); torch.cuda.synchronize(); build_ms = 1e3 * (time.perf_counter() - t0)
qd = torch.from_numpy(q).cuda()
ours = ev_time(lambda: s.nearest_points(qd))
pts, faces, d2 = s.nearest_points(qd, return_dist2=True)
out = {'scan_verts': len(v), 'scan_faces': len(f), 'queries': len(q), 'cells': s.num[3], 'grid_build_ms_incl_alloc': build_ms,
       'ours_ms': ours, 'ours_mqueries_per_s': len(q) / ours / 1e3}
so = glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'mesh_grid*.so'))
if so:
    spec = importlib.util.spec_from_file_location('mesh_grid', so[0]); mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    verts, fa = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
    num = torch.tensor(s.num, dtype=torch.int32).cuda(); minmax = torch.from_numpy(np.asarray(s.minmax, np.float32)).cuda()
    tri_num = torch.zeros(s.num[3], dtype=torch.int32).cuda(); tri_idx = torch.zeros(1, dtype=torch.int32).cuda()
    t0 = time.perf_counter(); mg.insert_grid_surface(verts, fa, minmax, num, s.step, tri_num, tri_idx); torch.cuda.synchronize()
    out['reference_grid_build_ms'] = 1e3 * (time.perf_counter() - t0)
    nf = torch.zeros(len(q), dtype=torch.int32).cuda(); co = torch.zeros(len(q), 3).cuda(); npts = torch.zeros(len(q), 3).cuda()
    ref = ev_time(lambda: mg.search_nearest_point(qd, verts, fa, tri_num, tri_idx, num, minmax, s.step, nf, npts, co))
    dref = torch.norm(npts - qd, dim=1)
    out.update(reference_ms=ref, speedup_vs_reference_kernel=ref / ours,
               max_abs_dist_diff=float((dref - d2.sqrt()).abs().max()), ours_never_worse=bool((d2.sqrt() <= dref + 1e-6).all()))
# inside test / ray queries on the same grid: 100k queries in and around the scan
n = 100000
lo, hi = v.min(0), v.max(0)
qi = torch.from_numpy((rng.rand(n, 3) * (hi - lo) * 1.3 + lo - 0.15 * (hi - lo)).astype(np.float32)).cuda()
ro = torch.from_numpy((rng.rand(n, 3) * (hi - lo) * 2.0 + lo - 0.5 * (hi - lo)).astype(np.float32)).cuda()
rd = torch.from_numpy(rng.randn(n, 3).astype(np.float32)).cuda()
out['inside_queries'] = n
out['inside_ms'] = ev_time(lambda: s.inside_mesh(qi))
out['rays_ms'] = ev_time(lambda: s.intersects_any(ro, rd))
if so:
    signs = torch.zeros(n).cuda(); hit = torch.zeros(n, dtype=torch.bool).cuda()
    out['reference_inside_ms'] = ev_time(lambda: mg.search_inside_mesh(qi, verts, fa, tri_num, tri_idx, num, minmax, s.step, signs), reps=5)
    out['reference_rays_ms'] = ev_time(lambda: mg.search_intersect(ro, rd, verts, fa, tri_num, tri_idx, num, minmax, s.step, hit), reps=5)
    out['inside_agreement'] = float((s.inside_mesh(qi) == signs).float().mean())
    out['rays_agreement'] = float((s.intersects_any(ro, rd) == hit).float().mean())
print(json.dumps(out))
