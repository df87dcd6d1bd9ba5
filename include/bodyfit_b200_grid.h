/*
 * bodyfit_b200_grid.h -- uniform-grid closest point on a triangle mesh and the SMPL+D displacement
 * step.  Replaces the reference's native module `mesh_grid` (thirdparty/mesh_grid/mesh_grid.cpp:129-136:
 * insert_grid_surface, search_nearest_point) as used by utils/mesh_grid_searcher.py:56-84,
 * smplify/loss.py:233-242,260-288 and smplify/smplify.py:146-156,228-247.
 * Conventions as in bodyfit_b200.h (extern "C", device pointers, caller's stream, 0 / negative code,
 * no allocation: the caller sizes `cell_tris` from cell_start[ncell] after bf_grid_count).
 */
#ifndef BODYFIT_B200_GRID_H
#define BODYFIT_B200_GRID_H
#include <stdint.h>
#include "bodyfit_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct BfGrid {
    const float*   verts;       /* [Ns,3] scan vertices */
    const int32_t* faces;       /* [Fs,3] */
    int32_t*       cell_start;  /* [ncell+1] CSR offsets into cell_tris */
    int32_t*       cell_tris;   /* [cell_start[ncell]] face ids, ascending inside each cell */
    float          min[3];      /* grid origin (mesh_grid_searcher.py:69) */
    float          step;        /* cell size   (mesh_grid_searcher.py:65) */
    int32_t        dim[3];      /* cells per axis (:68) */
    int32_t        ncell, Ns, Fs;
} BfGrid;

/* pass 1: triangles per cell -> counts[ncell] (zeroed here), exclusive scan -> g->cell_start[ncell+1] */
int bf_grid_count(const BfGrid* g, int32_t* counts, void* stream);
/* pass 2: fill g->cell_tris (sized by the caller from cell_start[ncell]) and sort every cell's list; cursor[ncell] scratch */
int bf_grid_fill(const BfGrid* g, int32_t* cursor, void* stream);
/* closest point of every query: near_pts[Q,3], near_faces[Q] (-1 if the mesh is empty), dist2[Q] (optional) */
int bf_grid_nearest(const BfGrid* g, const float* points, int Q, float* near_pts, int32_t* near_faces, float* dist2, void* stream);

/* barycentric coefficients [Q,3] of the closest point on the selected face (the reference's `coeff` output,
 * mesh_grid.cpp:54-72, mesh_grid_kernel.cu:405-410): near_pt = c0 v0 + c1 v1 + c2 v2; zeros where near_faces < 0 */
int bf_grid_barycentric(const BfGrid* g, const float* points, const int32_t* near_faces, int Q, float* coeff, void* stream);
/* grad[Q,3,3,3]: grad[q][i][j][k] = d near_pt[q][j] / d verts[faces[near_faces[q]][i]][k] (replaces mesh_grid.cpp:119-127
 * search_nearest_point_backward, whose kernel is unfinished in the reference and never called) */
int bf_grid_nearest_backward(const BfGrid* g, const float* points, const int32_t* near_faces, int Q, float* grad, void* stream);

/* inside test (utils/mesh_grid_searcher.py:86-91, native search_inside_mesh): signs[Q] = +1 inside the closed mesh, -1 outside
 * (crossing parity of an axis ray towards the nearest grid border; points outside the grid box are outside) */
int bf_grid_inside(const BfGrid* g, const float* points, int Q, float* signs, void* stream);
/* ray queries (utils/mesh_grid_searcher.py:93-99, native search_intersect): hit[Q] = 1 iff the ray origin + t * direction,
 * t >= 0, meets any triangle; a zero direction never hits */
int bf_grid_intersects_any(const BfGrid* g, const float* origins, const float* directions, int Q, uint8_t* hit, void* stream);

/* SMPL+D displacement step (smplify/smplify.py:236-245) for one subject, body mesh with V vertices / F faces:
 *   P = base + disp; normals; closest points (computed ONCE, the reference searches twice: loss.py:239,267);
 *   loss = |P - C|_F + (mean(1 - <scan_fn[closest], n_v>) + laplacian(n)) * reg_scale; backward; Adam on disp.
 * totals[4] = icp, normal term, smoothness, loss. */
typedef struct BfSmpld {
    const float*   base;        /* [V,3] fitted body vertices (detached) */
    float*         disp;        /* [V,3] in/out */
    float*         adam_m;      /* [V,3] */
    float*         adam_v;      /* [V,3] */
    const int32_t* faces;       /* [F,3] body faces */
    const int32_t* vf_ptr;      /* [V+1] CSR vertex -> incident faces */
    const int32_t* vf_face;
    const float*   scan_fn;     /* [Fs,3] un-normalised scan face normals (smplify.py:149) */
    float*         P;           /* [V,3] scratch: deformed vertices */
    float*         C;           /* [V,3] scratch: closest points */
    int32_t*       near_faces;  /* [V] */
    float*         nhat;        /* [F,3] */
    float*         nlen;        /* [F] */
    float*         m;           /* [V,3] vertex normals */
    float*         Nlen;        /* [V] */
    float*         dN;          /* [V,3] */
    float*         dcorner;     /* [F,9] */
    float*         partial;     /* [64,3] */
    float*         totals;      /* [4] */
    float*         grad;        /* [V,3] optional: d loss / d disp of this step */
    float*         trace;       /* [n_iters,4] optional */
    double         lr, beta1, beta2, eps;
    float          reg_scale;   /* constant_scale * 0.1 */
    int32_t        V, F, iter;
} BfSmpld;
int bf_smpld_step(const BfGrid* g, const BfSmpld* s, void* stream);
int bf_smpld_run(const BfGrid* g, const BfSmpld* s, int n_iters, void* stream);

/* point-to-scan term of the main loop (smplify/loss.py:233-242, smplify/smplify.py:205-206,210) for the B frames of
 * `f` (all-vertex buffers, model-space f->verts): Pw = (verts + transl) * scale * constant_scale; closest points;
 * pc = scale * |Pw - C|_F per frame; f->loss[b] += weight * pc; its gradient is ADDED to f->dverts (model space)
 * and f->grad[:, 0:4] (transl, scale).  Pw, near_pts [B,V,3] and near_faces [B,V] are scratch / outputs. */
int bf_pc_loss(const BfGrid* g, const BfModel* m, const BfFrames* f, float scale, float weight, float* Pw,
               float* near_pts, int32_t* near_faces, float* pc_loss, void* stream);

#ifdef __cplusplus
}
#endif
#endif
