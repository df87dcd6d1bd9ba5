/*
 * bodyfit_b200_ops.h -- stand-alone operators of the reference's loss / prior surface, for callers
 * that compose the objective themselves (the fused fit loop in bodyfit_b200.h does not use them).
 * Same conventions as bodyfit_b200.h: extern "C", device pointers, caller's stream, 0 / negative code.
 * Every forward has its hand-written backward.  fp32, row-major, contiguous.
 *
 *   bf_op_project(+_backward)   smplify/loss.py:22-43   perspective_projection(points[B,N,3], rotation[nb,3,3],
 *                               translation[nb,3], K[3,3]) -> [B,N,2], nb = 1 (broadcast) or B
 *   bf_op_gmof(+_backward)      smplify/loss.py:45-51   sigma^2 x^2 / (sigma^2 + x^2), elementwise
 *   bf_op_reprojection          smplify/loss.py:132-136 sum_j w_j sum_c gmof((gt - cord)/coef): value + d/dcord;
 *                               w_j = conf_j^2 (body) or the group's sum of conf^2 ([N,1] confidences, :168-179)
 *   bf_op_keypoints_world       smplify/loss.py:156-203 data term on world joints [B,K,3] against packed
 *                               detections [B,K,Nv,3] = (x, y, weight) (joint-major) and cams [Nv,12] = K [R|t]:
 *                               per (frame, joint) loss / Nv and d/d joints
 *   bf_op_angle_prior           smplify/loss.py:54-61   exp(sign * pose[:, [52,55,9,12]])^2 and its derivative
 *   bf_op_gmm_pose              smplify/prior.py:181-196 weight * min_m(0.5 d^T P_m d - log nll_w_m) on pose[B, ld]
 *                               (first nvalid <= 69 columns used, rest 0) and its gradient [B,69]
 */
#ifndef BODYFIT_B200_OPS_H
#define BODYFIT_B200_OPS_H
#include "bodyfit_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int bf_op_project(const float* pts, const float* R, const float* t, const float* K, float* uv, int B, int N, int nb, void* stream);
int bf_op_project_backward(const float* pts, const float* R, const float* t, const float* K, const float* duv, float* dpts,
                           int B, int N, int nb, void* stream);
int bf_op_gmof(const float* x, float* y, float sigma, int64_t n, void* stream);
int bf_op_gmof_backward(const float* x, const float* dy, float* dx, float sigma, int64_t n, void* stream);
int bf_op_reprojection(const float* cord, const float* gt, const float* w, float coef, float sigma, int N, float* out,
                       float* dcord, void* stream);
int bf_op_keypoints_world(const float* joints, const float* kp, const float* cams, int B, int K, int Nv, float coef,
                          float sigma, float* loss_bk, float* dJ, void* stream);
int bf_op_angle_prior(const float* pose, int B, int D, float* out, float* dout, void* stream);
int bf_op_gmm_pose(const BfModel* m, const float* pose, int ld, int nvalid, int B, float weight, float* grad, float* loss,
                   void* stream);

/* scan / normal / silhouette terms (the SMPL+D loop and the dense loop run fused versions: bf_smpld_step, bf_pc_loss,
 * bf_mask_loss).  Every forward also returns the gradient w.r.t. its differentiable input.
 *   bf_op_pc_loss            smplify/loss.py:233-242  out[0] = |points - closest|_F (n floats), dpoints = (points - closest)/out
 *   bf_op_normal_loss        smplify/loss.py:260-271  out[0] = mean(1 - <face_norm[near_faces[v]], point_norm[v]>), d/d point_norm
 *   bf_op_laplacian          smplify/loss.py:273-288  out[0] = mean over faces of the pairwise squared normal differences, d/d norms
 *                            (vf_ptr / vf_face: CSR vertex -> incident faces, for the atomics-free gather)
 *   bf_op_vertex_normals     utils/io_utils.py:405-428 compute_normal_torch: unit face normals summed per vertex, renormalised
 *   bf_op_vertex_normals_backward   its gradient w.r.t. the vertex positions
 *   bf_op_mask_loss          smplify/loss.py:85-130 multview_mask_loss on WORLD vertices [B,V,3]: loss[b] and d/d verts
 *                            (scratch: 8 B floats; BfMask from bodyfit_b200_mask.h)
 *   bf_op_regress_joints     smplx vertices2joints / models/smpl.py:85-87 get_joints_h36m: out[B,R,3] = W[R,N] @ points[B,N,3]
 *                            with W in CSR form (ptr [R+1], idx, w); its backward is the same call with the CSR of W^T */
struct BfMask;
int bf_op_regress_joints(const float* points, const int32_t* ptr, const int32_t* idx, const float* w, int B, int N, int R,
                         float* out, void* stream);
int bf_op_pc_loss(const float* points, const float* closest, int64_t n, float* out, float* dpoints, void* stream);
int bf_op_normal_loss(const int32_t* near_faces, const float* face_norm, const float* point_norm, int V, float* out,
                      float* dpoint_norm, void* stream);
int bf_op_laplacian(const float* norms, const int32_t* faces, const int32_t* vf_ptr, const int32_t* vf_face, int V, int F,
                    float* out, float* dnorms, void* stream);
int bf_op_vertex_normals(const float* verts, const int32_t* faces, const int32_t* vf_ptr, const int32_t* vf_face, int V, int F,
                         float* nhat, float* nlen, float* normals, float* Nlen, void* stream);
int bf_op_vertex_normals_backward(const float* verts, const int32_t* faces, const int32_t* vf_ptr, const int32_t* vf_face, int V,
                                  int F, const float* nhat, const float* nlen, const float* normals, const float* Nlen,
                                  const float* dnormals, float* dm_scratch, float* dcorner_scratch, float* dverts, void* stream);
int bf_op_mask_loss(const float* verts_world, int B, int V, const struct BfMask* k, float* scratch, float* loss, float* dverts,
                    void* stream);
#ifdef __cplusplus
}
#endif
#endif
