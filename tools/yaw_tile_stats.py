"""Potential of sorting the frames of a batch by contour (yaw) row -- CPU only.

The blend GEMMs of the fit compute all active vertices for every frame, but a frame only uses the static ones and the 17 x 3
contour vertices of its own yaw row (model.py::_build_live_tables).  This script counts, for the benchmark workload's poses,
how many different rows (and how many distinct live vertices) a 128-frame GEMM tile holds with the frames in their given
order and with the frames permuted by row ("frames sorted by contour row").
    python tools/yaw_tile_stats.py [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from scipy.spatial.transform import Rotation as Rot
from bodyfitting_b200 import synthetic as syn
from bodyfitting_b200.model import PreparedModel

F = int(sys.argv[1]) if len(sys.argv) > 1 else 10000


def yaw_rows(go, bp):
    """Row of the dynamic contour table from the head yaw, as k_pose_fwd computes it (bf_pose.cuh, smplx
    find_dynamic_lmk_idx_and_bcoords): rel = R0 R3 R6 R9 R12, yaw = atan2(-rel[2][0], |rel[:2, 0]|), 1-degree rows."""
    R = lambda aa: Rot.from_rotvec(aa).as_matrix()
    rel = R(bp[:, 33:36])
    for j in (9, 6, 3):
        rel = R(bp[:, (j - 1) * 3:(j - 1) * 3 + 3]) @ rel
    rel = R(go) @ rel
    ang = np.arctan2(-rel[:, 2, 0], np.sqrt(rel[:, 0, 0] ** 2 + rel[:, 1, 0] ** 2))
    y = np.rint(np.minimum(-ang * 180 / np.pi, 39.0)).astype(int)
    return np.where(y < 0, np.where(y < -39, 78, 39 - y), y)


pm = PreparedModel('smplx', syn.make_model('smplx', 0), gmm=syn.make_gmm(0), device='cpu')
lv_n = pm._dev['act_lv_n'].numpy()
lv_vid = pm._dev['act_lv_vid'].numpy().reshape(len(lv_n), -1)
live = [set(lv_vid[a, :lv_n[a]].tolist()) for a in range(len(lv_n))]
gt, init = syn.make_params('smplx', F, seed=100)
for name, p in (('initial poses', init), ('ground-truth poses', gt)):
    y = yaw_rows(p['global_orient'].astype(np.float64), p['body_pose'].astype(np.float64))
    for order, yy in (('given order', y), ('sorted by row', np.sort(y))):
        rows = [sorted(set(yy[i:i + 128].tolist())) for i in range(0, F, 128)]
        nrows = np.mean([len(r) for r in rows])
        nvert = np.mean([len(set().union(*[live[a] for a in r])) for r in rows])
        print('%-19s %-14s: %.1f rows and %.0f of %d active vertices per 128-frame tile (%.0f %% of the GEMM columns)'
              % (name, order, nrows, nvert, pm.n_act, 100.0 * nvert / pm.n_act))

# structure of the contour candidates: how well a row's vertices cluster when the contour vertices are ordered by the first row
# that uses them (the column order a per-tile block list would want)
static = set.intersection(*live)
contour = [l - static for l in live]
allc = set().union(*contour)
first = {}
for a, c in enumerate(contour):
    for v in c:
        first.setdefault(v, a)
pos = {v: i for i, v in enumerate(sorted(allc, key=lambda v: (first[v], v)))}
blocks = [len({pos[v] // 16 for v in c}) for c in contour]
overlap = [len(contour[a] & contour[a + 1]) for a in range(len(contour) - 1)]
print('%d static + %d contour vertices; adjacent rows share %.1f of %.1f contour vertices; in first-use order a row touches '
      '%.1f (max %d) of %d 16-vertex blocks' % (len(static), len(allc), np.mean(overlap), np.mean([len(c) for c in contour]),
                                                 np.mean(blocks), max(blocks), (len(allc) + 15) // 16))
