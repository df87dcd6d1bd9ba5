// Output joints, multi-view projection, GMoF keypoint term and its hand-written backward.
// One CTA per frame.
//
//   k_keypoint_loss : model joints (chain joints, picked vertices, barycentric face landmarks,
//                     yaw-dependent contour landmarks) -> world = (x + transl) * scale * cs
//                     -> per view pixel = (K [R|t]) [X;1] -> rho = s^2 r^2 / (s^2 + r^2) per
//                     coordinate with r = (gt - uv) / (imsize/1024), weighted and summed / Nv;
//                     gradient w.r.t. transl, scale and every model joint; then a gather BY TARGET
//                     (chain joint or vertex) turns joint gradients into dJtr / dverts rows without
//                     atomics (deterministic).
//   k_joints_fwd / k_joints_bwd : the same joint table for the LBS operator surface.
//
// Replaces smplify/loss.py:22-51,132-136,156-203 (perspective_projection, gmof, reprojection_loss,
// the per-view loop of multiview_keypoint_loss), smplify/smplify.py:189-190, models/smpl.py:72-75,
// models/utils.py:25-29 and smplx vertices2landmarks / vertex_joint_selector.
#pragma once
#include "bf_common.cuh"

#define BF_MAXK 144      // output joints (49 SMPL wrapper / 135 mapped SMPL-X)
#define BF_MAXVIEWS 64

__device__ __forceinline__ void joint_pos(const BfVSet& vs, int k, int yaw, const float* __restrict__ Jtr_b,
                                          const float* __restrict__ verts_b, float* x) {
    const int kind = __ldg(vs.kj_kind + k);
    const int* src = vs.kj_src + k * 3;
    if (kind == 0) {
        const float* p = Jtr_b + __ldg(src) * 3;
        x[0] = p[0]; x[1] = p[1]; x[2] = p[2];
    } else if (kind == 1 || kind == 2) {
        const int* s = src;
        const float* w = vs.kj_w + k * 3;
        if (kind == 2) {
            const int slot = __ldg(src);
            s = vs.dyn_src + ((size_t)yaw * vs.n_dyn + slot) * 3;
            w = vs.dyn_w + ((size_t)yaw * vs.n_dyn + slot) * 3;
        }
        x[0] = x[1] = x[2] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float wi = __ldg(w + i);
            const float* p = verts_b + (size_t)__ldg(s + i) * 3;
            x[0] += p[0] * wi; x[1] += p[1] * wi; x[2] += p[2] * wi;
        }
    } else {
        const int row = __ldg(src);
        x[0] = x[1] = x[2] = 0.f;
        for (int e = __ldg(vs.xr_ptr + row); e < __ldg(vs.xr_ptr + row + 1); ++e) {
            const float wi = __ldg(vs.xr_w + e);
            const float* p = verts_b + (size_t)__ldg(vs.xr_vid + e) * 3;
            x[0] += p[0] * wi; x[1] += p[1] * wi; x[2] += p[2] * wi;
        }
    }
}

// joint gradients gx[k][3] (shared memory) -> dJtr[b] and dverts[b] rows, gathered per target.
__device__ __forceinline__ void scatter_by_target(const BfVSet& vs, int J, int Klim, int yaw,
                                                  const float* gx, float* __restrict__ dJtr_b,
                                                  float* __restrict__ dverts_b, int accumulate) {
    const int ntg = J + vs.n;
    for (int tg = threadIdx.x; tg < ntg; tg += blockDim.x) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int e0 = __ldg(vs.tg_ptr + tg), e1 = __ldg(vs.tg_ptr + tg + 1);
        if (accumulate && e0 == e1 && tg >= J) continue;     // nothing to add to this vertex: leave its row untouched
        for (int e = e0; e < e1; ++e) {
            const int k = __ldg(vs.tg_k + e);
            if (k < Klim) {
                const float w = __ldg(vs.tg_w + e);
                a0 += w * gx[k * 3]; a1 += w * gx[k * 3 + 1]; a2 += w * gx[k * 3 + 2];
            }
        }
        if (tg < J) {
            dJtr_b[tg * 3] = a0; dJtr_b[tg * 3 + 1] = a1; dJtr_b[tg * 3 + 2] = a2;
        } else {
            float* o = dverts_b + (size_t)(tg - J) * 3;
            if (accumulate) { o[0] += a0; o[1] += a1; o[2] += a2; }
            else { o[0] = a0; o[1] = a1; o[2] = a2; }
        }
    }
    // contour landmarks: their three source vertices depend on the frame's yaw row, so they are
    // added after the static gather, slot by slot in a fixed order by one warp (deterministic,
    // the three vertices of a face are distinct)
    if (vs.n_dyn > 0) {
        __syncthreads();
        if (threadIdx.x < 32) {
            for (int s = 0; s < vs.n_dyn; ++s) {
                const int k = __ldg(vs.dyn_k + s);
                if (k < Klim && threadIdx.x < 3) {
                    const size_t e = ((size_t)yaw * vs.n_dyn + s) * 3 + threadIdx.x;
                    const float w = __ldg(vs.dyn_w + e);
                    float* o = dverts_b + (size_t)__ldg(vs.dyn_src + e) * 3;
                    o[0] += w * gx[k * 3]; o[1] += w * gx[k * 3 + 1]; o[2] += w * gx[k * 3 + 2];
                }
                __syncwarp();
            }
        }
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {     // deterministic tree, all threads get the sum
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(256) k_keypoint_loss(BfModel m, BfVSet vs, BfFrames f) {
    __shared__ float gx[BF_MAXK * 3];
    __shared__ float cam[BF_MAXVIEWS * 12];
    __shared__ float red[8];
    const int b = blockIdx.x;
    const int K = m.K_used, Nv = f.Nv, J = m.J;
    for (int i = threadIdx.x; i < Nv * 12; i += blockDim.x) cam[i] = f.cams[i];
    __syncthreads();
    const float* th = f.theta + (size_t)b * m.NP;
    const float tx = th[0], ty = th[1], tz = th[2], sc = th[3];
    const float cs = f.constant_scale;
    const float coef = f.imsize / 1024.0f;
    const float s2 = f.sigma * f.sigma;
    const int yaw = f.yaw ? f.yaw[b] : 0;
    const float* Jtr_b = f.Jtr + (size_t)b * J * 3;
    const float* verts_b = f.verts + (size_t)b * f.ld_v;
    const float invNv = 1.0f / (float)Nv;

    float lsum = 0.f, gT0 = 0.f, gT1 = 0.f, gT2 = 0.f, gs = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float x[3];
        joint_pos(vs, k, yaw, Jtr_b, verts_b, x);
        const float qx = x[0] + tx, qy = x[1] + ty, qz = x[2] + tz;
        const float X = qx * sc * cs, Y = qy * sc * cs, Z = qz * sc * cs;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        for (int v = 0; v < Nv; ++v) {
            const float* M = cam + v * 12;
            const float p0 = M[0] * X + M[1] * Y + M[2] * Z + M[3];
            const float p1 = M[4] * X + M[5] * Y + M[6] * Z + M[7];
            const float p2 = M[8] * X + M[9] * Y + M[10] * Z + M[11];
            const float iz = 1.0f / p2;
            const float u = p0 * iz, w_ = p1 * iz;
            const float* kp = f.kp + (((size_t)(f.frame_index ? f.frame_index[b] : b) * K + k) * Nv + v) * 3;
            const float wgt = kp[2];
            const float rx = (kp[0] - u) / coef, ry = (kp[1] - w_) / coef;
            const float rx2 = rx * rx, ry2 = ry * ry;
            const float dx = s2 + rx2, dy = s2 + ry2;
            lsum += wgt * ((s2 * rx2) / dx + (s2 * ry2) / dy);
            // d rho / d r = 2 s^4 r / (s^2 + r^2)^2 ; d r / d u = -1 / coef
            const float du = wgt * (2.0f * s2 * s2 * rx / (dx * dx)) * (-1.0f / coef);
            const float dv = wgt * (2.0f * s2 * s2 * ry / (dy * dy)) * (-1.0f / coef);
            const float dp0 = du * iz, dp1 = dv * iz, dp2 = -(du * u + dv * w_) * iz;
            g0 += M[0] * dp0 + M[4] * dp1 + M[8] * dp2;
            g1 += M[1] * dp0 + M[5] * dp1 + M[9] * dp2;
            g2 += M[2] * dp0 + M[6] * dp1 + M[10] * dp2;
        }
        g0 *= invNv; g1 *= invNv; g2 *= invNv;
        // world = (x + T) * s * cs
        const float k0 = sc * cs;
        gx[k * 3] = g0 * k0; gx[k * 3 + 1] = g1 * k0; gx[k * 3 + 2] = g2 * k0;
        gT0 += g0 * k0; gT1 += g1 * k0; gT2 += g2 * k0;
        gs += (g0 * qx + g1 * qy + g2 * qz) * cs;
    }
    lsum = block_sum(lsum, red);
    gT0 = block_sum(gT0, red); gT1 = block_sum(gT1, red); gT2 = block_sum(gT2, red);
    gs = block_sum(gs, red);
    if (threadIdx.x == 0) {
        f.loss[b] = lsum * invNv;
        float* g = f.grad + (size_t)b * m.NP;
        g[0] = gT0; g[1] = gT1; g[2] = gT2; g[3] = gs;
    }
    __syncthreads();
    scatter_by_target(vs, J, K, yaw, gx, f.dJtr + (size_t)b * J * 3, f.dverts + (size_t)b * f.ld_v, 0);
}

__global__ void __launch_bounds__(256) k_joints_fwd(BfModel m, BfVSet vs, BfFrames f) {
    const int b = blockIdx.x;
    const int yaw = f.yaw ? f.yaw[b] : 0;
    const float* Jtr_b = f.Jtr + (size_t)b * m.J * 3;
    const float* verts_b = f.verts + (size_t)b * f.ld_v;
    for (int k = threadIdx.x; k < vs.K_out; k += blockDim.x) {
        float x[3];
        joint_pos(vs, k, yaw, Jtr_b, verts_b, x);
        float* o = f.joints + ((size_t)b * vs.K_out + k) * 3;
        if (f.flags & BF_F_WORLD) {       // verts are already in world space; chain joints are not
            const float* th = f.theta + (size_t)b * m.NP;
            if (__ldg(vs.kj_kind + k) == 0) {
                const float sc = th[3], cs = f.constant_scale;
                x[0] = (x[0] + th[0]) * sc * cs; x[1] = (x[1] + th[1]) * sc * cs; x[2] = (x[2] + th[2]) * sc * cs;
            }
        }
        o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
    }
}

__global__ void __launch_bounds__(256) k_joints_bwd(BfModel m, BfVSet vs, BfFrames f, int accumulate) {
    __shared__ float gx[BF_MAXK * 3];
    const int b = blockIdx.x;
    const int yaw = f.yaw ? f.yaw[b] : 0;
    for (int i = threadIdx.x; i < vs.K_out * 3; i += blockDim.x)
        gx[i] = f.djoints ? f.djoints[(size_t)b * vs.K_out * 3 + i] : 0.f;
    __syncthreads();
    scatter_by_target(vs, m.J, vs.K_out, yaw, gx, f.dJtr + (size_t)b * m.J * 3,
                      f.dverts + (size_t)b * f.ld_v, accumulate);
}
