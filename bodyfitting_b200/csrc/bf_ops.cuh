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
