// Stand-alone operators of the reference's loss / prior surface (smplify/loss.py, smplify/prior.py),
// each with its hand-written backward, for callers that compose the objective themselves instead of
// using the fused fit loop:
//   k_project_fwd/bwd   perspective_projection          smplify/loss.py:22-43
//   k_gmof_fwd/bwd      gmof                            smplify/loss.py:45-51
//   k_reproj            reprojection_loss (+ gradient)  smplify/loss.py:132-136
//   k_kp_world          data term of multiview_keypoint_loss on given world joints   loss.py:156-203
//   k_angle_prior       angle_prior (+ gradient)        smplify/loss.py:54-61
//   k_gmm_pose          MaxMixturePrior.forward (+ gradient) on a [B,69] pose   smplify/prior.py:181-196
#pragma once
#include "bf_common.cuh"
#include "bf_gmm.cuh"

// uv[b,n,:] = (K (R x + t))_xy / (.)_z ; R [nb,3,3], t [nb,3] with nb = 1 (broadcast) or B; K [3,3]
__global__ void k_project_fwd(const float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ tr,
                              const float* __restrict__ K, float* __restrict__ uv, int B, int N, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N;
    const float* Rb = R + (nb > 1 ? b * 9 : 0);
    const float* tb = tr + (nb > 1 ? b * 3 : 0);
    const float x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
    const float c0 = Rb[0] * x + Rb[1] * y + Rb[2] * z + tb[0];
    const float c1 = Rb[3] * x + Rb[4] * y + Rb[5] * z + tb[1];
    const float c2 = Rb[6] * x + Rb[7] * y + Rb[8] * z + tb[2];
    const float p0 = K[0] * c0 + K[1] * c1 + K[2] * c2;
    const float p1 = K[3] * c0 + K[4] * c1 + K[5] * c2;
    const float p2 = K[6] * c0 + K[7] * c1 + K[8] * c2;
    uv[i * 2] = p0 / p2;
    uv[i * 2 + 1] = p1 / p2;
}

__global__ void k_project_bwd(const float* __restrict__ pts, const float* __restrict__ R, const float* __restrict__ tr,
                              const float* __restrict__ K, const float* __restrict__ duv, float* __restrict__ dpts,
                              int B, int N, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N;
    const float* Rb = R + (nb > 1 ? b * 9 : 0);
    const float* tb = tr + (nb > 1 ? b * 3 : 0);
    const float x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
    const float c0 = Rb[0] * x + Rb[1] * y + Rb[2] * z + tb[0];
    const float c1 = Rb[3] * x + Rb[4] * y + Rb[5] * z + tb[1];
    const float c2 = Rb[6] * x + Rb[7] * y + Rb[8] * z + tb[2];
    const float p0 = K[0] * c0 + K[1] * c1 + K[2] * c2;
    const float p1 = K[3] * c0 + K[4] * c1 + K[5] * c2;
    const float p2 = K[6] * c0 + K[7] * c1 + K[8] * c2;
    const float iz = 1.0f / p2;
    const float du = duv[i * 2], dv = duv[i * 2 + 1];
    const float dp0 = du * iz, dp1 = dv * iz, dp2 = -(du * p0 + dv * p1) * iz * iz;
    const float dc0 = K[0] * dp0 + K[3] * dp1 + K[6] * dp2;
    const float dc1 = K[1] * dp0 + K[4] * dp1 + K[7] * dp2;
    const float dc2 = K[2] * dp0 + K[5] * dp1 + K[8] * dp2;
    dpts[i * 3] = Rb[0] * dc0 + Rb[3] * dc1 + Rb[6] * dc2;
    dpts[i * 3 + 1] = Rb[1] * dc0 + Rb[4] * dc1 + Rb[7] * dc2;
    dpts[i * 3 + 2] = Rb[2] * dc0 + Rb[5] * dc1 + Rb[8] * dc2;
}

// y = s^2 x^2 / (s^2 + x^2);  dy/dx = 2 s^4 x / (s^2 + x^2)^2
__global__ void k_gmof_fwd(const float* __restrict__ x, float* __restrict__ y, float sigma, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x2 = x[i] * x[i], s2 = sigma * sigma;
    y[i] = (s2 * x2) / (s2 + x2);
}
__global__ void k_gmof_bwd(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, float sigma, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s2 = sigma * sigma, d = s2 + x[i] * x[i];
    dx[i] = dy[i] * 2.0f * s2 * s2 * x[i] / (d * d);
}

// reprojection_loss: out = sum_j w_j * sum_c gmof((gt - cord)/coef)_jc with per-joint weights w (conf^2, or
// the group sum when the caller passes an [N,1] confidence as the reference does for hands / face);
// dcord written as well (gradient of out w.r.t. cord).  One block.
__global__ void k_reproj(const float* __restrict__ cord, const float* __restrict__ gt, const float* __restrict__ w,
                         float coef, float sigma, int N, float* __restrict__ out, float* __restrict__ dcord) {
    __shared__ float red[32];
    const float s2 = sigma * sigma;
    float acc = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float wj = w[j];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float r = (gt[j * 2 + c] - cord[j * 2 + c]) / coef;
            const float d = s2 + r * r;
            acc += wj * (s2 * r * r) / d;
            dcord[j * 2 + c] = wj * (2.0f * s2 * s2 * r / (d * d)) * (-1.0f / coef);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
        out[0] = s;
    }
}

// data term on given WORLD joints [B,K,3]: loss[b] = (1/Nv) sum_v sum_k w[b,v,k] (rho_x + rho_y), and dJ [B,K,3].
// One warp per (frame, joint) would waste lanes; one thread per (frame, joint), views in a loop.
__global__ void k_kp_world(const float* __restrict__ joints, const float* __restrict__ kp, const float* __restrict__ cams,
                           int B, int K, int Nv, float coef, float sigma, float* __restrict__ loss_bk,
                           float* __restrict__ dJ) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * K) return;
    const int b = i / K, k = i % K;
    const float X = joints[i * 3], Y = joints[i * 3 + 1], Z = joints[i * 3 + 2];
    const float s2 = sigma * sigma;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, ls = 0.f;
    for (int v = 0; v < Nv; ++v) {
        const float* M = cams + v * 12;
        const float p0 = M[0] * X + M[1] * Y + M[2] * Z + M[3];
        const float p1 = M[4] * X + M[5] * Y + M[6] * Z + M[7];
        const float p2 = M[8] * X + M[9] * Y + M[10] * Z + M[11];
        const float iz = 1.0f / p2;
        const float u = p0 * iz, w_ = p1 * iz;
        const float* q = kp + (((size_t)b * K + k) * Nv + v) * 3;
        const float wgt = q[2];
        const float rx = (q[0] - u) / coef, ry = (q[1] - w_) / coef;
        const float dx = s2 + rx * rx, dy = s2 + ry * ry;
        ls += wgt * ((s2 * rx * rx) / dx + (s2 * ry * ry) / dy);
        const float du = wgt * (2.0f * s2 * s2 * rx / (dx * dx)) * (-1.0f / coef);
        const float dw = wgt * (2.0f * s2 * s2 * ry / (dy * dy)) * (-1.0f / coef);
        const float dp0 = du * iz, dp1 = dw * iz, dp2 = -(du * u + dw * w_) * iz;
        g0 += M[0] * dp0 + M[4] * dp1 + M[8] * dp2;
        g1 += M[1] * dp0 + M[5] * dp1 + M[9] * dp2;
        g2 += M[2] * dp0 + M[6] * dp1 + M[10] * dp2;
    }
    const float inv = 1.0f / (float)Nv;
    loss_bk[i] = ls * inv;
    dJ[i * 3] = g0 * inv; dJ[i * 3 + 1] = g1 * inv; dJ[i * 3 + 2] = g2 * inv;
}

// angle prior: out[b, q] = exp(sign_q * pose[b, idx_q])^2 (q = 0..3), dpose gets 2 * out * sign at idx_q
__global__ void k_angle_prior(const float* __restrict__ pose, int B, int D, float* __restrict__ out, float* __restrict__ dout_dpose) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 4) return;
    const int b = i / 4, q = i % 4;
    const int idx = (q == 0) ? 52 : (q == 1) ? 55 : (q == 2) ? 9 : 12;
    const float sg = (q == 0) ? 1.0f : -1.0f;
    const float e = expf(pose[(size_t)b * D + idx] * sg);
    out[i] = e * e;
    dout_dpose[i] = 2.0f * e * e * sg;
}


// ---- scan / normal terms as stand-alone operators (smplify/loss.py:233-242,260-288, utils/io_utils.py:410-428) --------
// Each forward also produces the gradient w.r.t. its differentiable input (the objective is a scalar), like the operators
// above; single-block fixed-order reductions (deterministic).

// out[0] = |P - C|_F over n floats (loss.py:240-241: one Frobenius norm, the mean that follows is of a scalar);
// dP = (P - C) / out[0]
__global__ void __launch_bounds__(1024) k_op_pc_loss(const float* __restrict__ P, const float* __restrict__ C, int n,
                                                      float* __restrict__ out, float* __restrict__ dP) {
    __shared__ float red[32];
    __shared__ float nrm;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float acc = 0.f;
    for (int i = t; i < n; i += 1024) { const float d = P[i] - C[i]; acc += d * d; }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) { float s = 0.f; for (int w = 0; w < 32; ++w) s += red[w]; nrm = sqrtf(s); out[0] = nrm; }
    __syncthreads();
    const float inv = nrm > 0.f ? 1.0f / nrm : 0.f;
    for (int i = t; i < n; i += 1024) dP[i] = (P[i] - C[i]) * inv;
}

// out[0] = mean_v (1 - <fn[near_faces[v]], N_v>) (loss.py:267-269); dN_v = -fn[near_faces[v]] / V
__global__ void __launch_bounds__(1024) k_op_normal_loss(const int32_t* __restrict__ near_faces, const float* __restrict__ fn,
                                                          const float* __restrict__ N, int V, float* __restrict__ out,
                                                          float* __restrict__ dN) {
    __shared__ float red[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float acc = 0.f;
    const float iv = 1.0f / (float)V;
    for (int v = t; v < V; v += 1024) {
        const float* f3 = fn + 3 * (size_t)near_faces[v];
        acc += 1.0f - (f3[0] * N[3 * v] + f3[1] * N[3 * v + 1] + f3[2] * N[3 * v + 2]);
        dN[3 * v] = -f3[0] * iv; dN[3 * v + 1] = -f3[1] * iv; dN[3 * v + 2] = -f3[2] * iv;
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) { float s = 0.f; for (int w = 0; w < 32; ++w) s += red[w]; out[0] = s * iv; }
}

// out[0] = mean_f (|na - nb|^2 + |nc - na|^2 + |nb - nc|^2) (loss.py:273-288);
// dN_v = (2 / F) sum over incident face corners (2 N_v - N_o1 - N_o2), gathered per vertex through the CSR (no atomics)
__global__ void __launch_bounds__(1024) k_op_laplacian_value(const float* __restrict__ N, const int32_t* __restrict__ faces, int F,
                                                              float* __restrict__ out) {
    __shared__ float red[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float acc = 0.f;
    for (int i = t; i < F; i += 1024) {
        const float* na = N + 3 * (size_t)faces[3 * i];
        const float* nb = N + 3 * (size_t)faces[3 * i + 1];
        const float* nc = N + 3 * (size_t)faces[3 * i + 2];
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float ab = na[d] - nb[d], ca = nc[d] - na[d], bc = nb[d] - nc[d];
            s += ab * ab + ca * ca + bc * bc;
        }
        acc += s;
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (t == 0) { float s = 0.f; for (int w = 0; w < 32; ++w) s += red[w]; out[0] = s / (float)F; }
}
__global__ void k_op_laplacian_grad(const float* __restrict__ N, const int32_t* __restrict__ faces,
                                    const int32_t* __restrict__ vf_ptr, const int32_t* __restrict__ vf_face, int V, int F,
                                    float* __restrict__ dN) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float mv[3] = {N[3 * v], N[3 * v + 1], N[3 * v + 2]};
    const float k = 2.0f / (float)F;
    float g[3] = {0.f, 0.f, 0.f};
    for (int e = vf_ptr[v]; e < vf_ptr[v + 1]; ++e) {
        const int f = vf_face[e];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int o = faces[3 * f + c];
            if (o == v) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] += k * (mv[d] - N[3 * (size_t)o + d]);
        }
    }
    dN[3 * v] = g[0]; dN[3 * v + 1] = g[1]; dN[3 * v + 2] = g[2];
}

// vertex normals, backward: g = dL/dN_v (unit normals) -> through N = m / (|m| + eps) -> dm
__global__ void k_op_vnormal_bwd(const float* __restrict__ g, const float* __restrict__ N, const float* __restrict__ Nlen, int V,
                                 float* __restrict__ dm) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float len = Nlen[v], le = len + 1e-8f;
    const float nv[3] = {N[3 * v], N[3 * v + 1], N[3 * v + 2]};
    const float dot = nv[0] * g[3 * v] + nv[1] * g[3 * v + 1] + nv[2] * g[3 * v + 2];
    const float c = (len > 0.f) ? dot / len : 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) dm[3 * v + d] = g[3 * v + d] / le - nv[d] * c;
}
// per-corner gradients [F,3,3] -> per vertex (CSR gather, fixed order)
__global__ void k_op_corner_gather(const float* __restrict__ dcorner, const int32_t* __restrict__ faces,
                                   const int32_t* __restrict__ vf_ptr, const int32_t* __restrict__ vf_face, int V,
                                   float* __restrict__ dverts) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float g[3] = {0.f, 0.f, 0.f};
    for (int e = vf_ptr[v]; e < vf_ptr[v + 1]; ++e) {
        const int f = vf_face[e];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (faces[3 * f + c] == v) {
#pragma unroll
                for (int d = 0; d < 3; ++d) g[d] += dcorner[9 * (size_t)f + 3 * c + d];
            }
    }
    dverts[3 * v] = g[0]; dverts[3 * v + 1] = g[1]; dverts[3 * v + 2] = g[2];
}
__global__ void k_op_identity_theta(float* __restrict__ theta4, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { theta4[4 * b] = 0.f; theta4[4 * b + 1] = 0.f; theta4[4 * b + 2] = 0.f; theta4[4 * b + 3] = 1.0f; }
}

// out[b, r, :] = sum_e w[e] x[b, idx[e], :] over the CSR row r (ptr / idx / w) -- a sparse regressor applied to [B,N,3] points:
// joints = J_regressor @ vertices (smplx vertices2joints; models/smpl.py:85-87 get_joints_h36m) and, with the transposed
// matrix, its backward.  One warp per (frame, row), fixed butterfly -> deterministic.
__global__ void __launch_bounds__(256) k_op_spmm3(const float* __restrict__ x, const int32_t* __restrict__ ptr,
                                                  const int32_t* __restrict__ idx, const float* __restrict__ w, int B, int N, int R,
                                                  float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const size_t wi = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wi >= (size_t)B * R) return;
    const int b = (int)(wi / R), r = (int)(wi % R);
    const float* xb = x + (size_t)b * N * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) {
        const float we = __ldg(w + e);
        const float* p = xb + 3 * (size_t)__ldg(idx + e);
        a0 = fmaf(we, p[0], a0); a1 = fmaf(we, p[1], a1); a2 = fmaf(we, p[2], a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { float* o = out + 3 * wi; o[0] = a0; o[1] = a1; o[2] = a2; }
}
