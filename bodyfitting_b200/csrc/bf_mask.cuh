// Silhouette term of the fitting loop (use_mask=True): smplify/loss.py:85-130 multview_mask_loss, entering the
// objective with weight 5 from iteration N/3 + 1 on (smplify/smplify.py:196-199,210).
//
// Per frame and mask view: every 4th body vertex is projected; for every CONTOUR pixel of the view's mask the closest
// projected in-image vertex is found (the reference's cdist + min over the vertex axis, loss.py:111-112), the
// distance is charged once, or `epsilon` (10) times if that vertex's pixel lies outside the mask (:115-118); on top,
// epsilon x the bilinear sample of (1 - mask) at every projected vertex (grid_sample, zero padding, align_corners =
// False, :124-128).  Gradients flow through the selected vertex's projection and through the bilinear sample.
//
//   k_mask_project : (frame, view, sampled vertex) -> pixel
//   k_mask_nearest : one WARP per contour pixel: lanes stride over the projected vertices, fixed butterfly arg-min
//                    (ties -> lowest vertex index), then the inside/outside coefficient from the mask
//   k_mask_vertex  : one thread per (frame, sampled vertex): gathers the contour pixels that selected it (fixed order,
//                    no atomics), adds the bilinear term, chains both through the projection of every view
//   k_mask_finish  : one CTA per frame: loss value, d/d(model vertices), d/d(transl, scale)
// Distances are computed directly (|x - c|); torch.cdist switches to a matmul expansion for > 25 points whose fp32
// cancellation (~1e-2 px here) can flip arg-mins -- the oracle offers both (oracle/fit_port.py mask_objective).
#pragma once
#include "bf_common.cuh"
#include "../../include/bodyfit_b200_mask.h"

__device__ __forceinline__ void mask_world_point(const BfFrames& f, int NP, int b, int v, float* P) {
    const float* th = f.theta + (size_t)b * NP;
    const float* p = f.verts + (size_t)b * f.ld_v + 3 * (size_t)v;
    const float sc = th[3], cs = f.constant_scale;
    P[0] = (p[0] + th[0]) * sc * cs; P[1] = (p[1] + th[1]) * sc * cs; P[2] = (p[2] + th[2]) * sc * cs;
}

__global__ void __launch_bounds__(256) k_mask_project(BfFrames f, BfMask k, int NP) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)f.B * k.Nm * k.Nq;
    if (i >= n) return;
    const int q = (int)(i % k.Nq), m = (int)((i / k.Nq) % k.Nm), b = (int)(i / ((size_t)k.Nq * k.Nm));
    float P[3];
    mask_world_point(f, NP, b, q * k.stride, P);
    const float* M = k.cams + m * 12;
    const float p0 = M[0] * P[0] + M[1] * P[1] + M[2] * P[2] + M[3];
    const float p1 = M[4] * P[0] + M[5] * P[1] + M[6] * P[2] + M[7];
    const float p2 = M[8] * P[0] + M[9] * P[1] + M[10] * P[2] + M[11];
    k.uv[2 * i] = p0 / p2;
    k.uv[2 * i + 1] = p1 / p2;
}

__global__ void __launch_bounds__(256) k_mask_nearest(BfMask k, int total) {
    const int lane = threadIdx.x & 31;
    const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= total) return;
    const int bm = __ldg(k.cown + ci);                                  // frame * Nm + view
    const float cx = __ldg(k.contour + 2 * ci), cy = __ldg(k.contour + 2 * ci + 1);
    const float2* uv = reinterpret_cast<const float2*>(k.uv) + (size_t)bm * k.Nq;
    float best = 3.0e38f;
    int bq = 0x7fffffff;
    for (int q = lane; q < k.Nq; q += 32) {
        const float2 p = uv[q];
        if (!(p.x >= 0.f && p.x < k.imsize && p.y >= 0.f && p.y < k.imsize)) continue;     // in-image vertices only (:106-107)
        const float dx = p.x - cx, dy = p.y - cy;
        const float d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; bq = q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oq = __shfl_xor_sync(0xffffffffu, bq, o);
        if (ob < best || (ob == best && oq < bq)) { best = ob; bq = oq; }
    }
    if (lane == 0) {
        if (bq == 0x7fffffff) { k.near_q[ci] = -1; k.cdist[ci] = 0.f; k.cw[ci] = 0.f; return; }
        const float2 p = uv[bq];
        const int px = (int)p.x, py = (int)p.y;                          // .long() truncation of in-image coordinates (:115)
        const float mv = (px < k.W && py < k.H) ? __ldg(k.masks + ((size_t)bm * k.H + py) * k.W + px) : 0.f;
        k.near_q[ci] = bq;
        k.cdist[ci] = sqrtf(best);
        k.cw[ci] = mv < 0.1f ? k.epsilon : 1.0f;
    }
}

__global__ void __launch_bounds__(128) k_mask_vertex(BfFrames f, BfMask k, int NP) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)f.B * k.Nq) return;
    const int q = (int)(i % k.Nq), b = (int)(i / k.Nq);
    float P[3];
    mask_world_point(f, NP, b, q * k.stride, P);
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, bin = 0.f;
    for (int m = 0; m < k.Nm; ++m) {
        const int bm = b * k.Nm + m;
        const float2 p = reinterpret_cast<const float2*>(k.uv)[(size_t)bm * k.Nq + q];
        float du = 0.f, dv = 0.f;
        if (p.x >= 0.f && p.x < k.imsize && p.y >= 0.f && p.y < k.imsize) {          // only in-image vertices can be selected
            const int c0 = __ldg(k.cptr + bm), c1 = __ldg(k.cptr + bm + 1);
            for (int c = c0; c < c1; ++c) {
                if (__ldg(k.near_q + c) != q) continue;
                const float d = __ldg(k.cdist + c);
                if (d > 0.f) {
                    const float w = __ldg(k.cw + c) / d;
                    du += w * (p.x - __ldg(k.contour + 2 * c));
                    dv += w * (p.y - __ldg(k.contour + 2 * c + 1));
                }
            }
        }
        // epsilon * bilinear sample of (1 - mask), zero padding outside the image
        {
            const float sx = (float)k.W / k.imsize, sy = (float)k.H / k.imsize;
            const float ix = p.x * sx - 0.5f, iy = p.y * sy - 0.5f;
            const float fx = floorf(ix), fy = floorf(iy);
            const float ax = ix - fx, ay = iy - fy;
            const int x0 = (int)fx, y0 = (int)fy;
            const float* mk = k.masks + (size_t)bm * k.H * k.W;
            float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;          // v[y][x] of (1 - mask), 0 outside
            if (fx >= -1.f && fy >= -1.f && fx < (float)k.W && fy < (float)k.H) {
                const bool xa = x0 >= 0, xb = x0 + 1 < k.W, ya = y0 >= 0, yb = y0 + 1 < k.H;
                if (ya && xa) v00 = 1.0f - __ldg(mk + (size_t)y0 * k.W + x0);
                if (ya && xb) v01 = 1.0f - __ldg(mk + (size_t)y0 * k.W + x0 + 1);
                if (yb && xa) v10 = 1.0f - __ldg(mk + (size_t)(y0 + 1) * k.W + x0);
                if (yb && xb) v11 = 1.0f - __ldg(mk + (size_t)(y0 + 1) * k.W + x0 + 1);
            }
            bin += (v00 * (1.f - ax) + v01 * ax) * (1.f - ay) + (v10 * (1.f - ax) + v11 * ax) * ay;
            du += k.epsilon * sx * ((v01 - v00) * (1.f - ay) + (v11 - v10) * ay);
            dv += k.epsilon * sy * ((v10 - v00) * (1.f - ax) + (v11 - v01) * ax);
        }
        // chain through the projection of this view: uv = (p0, p1) / p2, p = M [P; 1]
        const float* M = k.cams + m * 12;
        const float p2 = M[8] * P[0] + M[9] * P[1] + M[10] * P[2] + M[11];
        const float iz = 1.0f / p2;
        const float dp0 = du * iz, dp1 = dv * iz, dp2 = -(du * p.x + dv * p.y) * iz;
        g0 += M[0] * dp0 + M[4] * dp1 + M[8] * dp2;
        g1 += M[1] * dp0 + M[5] * dp1 + M[9] * dp2;
        g2 += M[2] * dp0 + M[6] * dp1 + M[10] * dp2;
    }
    k.dPw[3 * i] = g0; k.dPw[3 * i + 1] = g1; k.dPw[3 * i + 2] = g2;
    k.part[i] = bin;
}

__global__ void __launch_bounds__(256) k_mask_finish(BfFrames f, BfMask k, int NP, float weight) {
    __shared__ float red[5][8];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float acc = 0.f;
    for (int c = __ldg(k.cptr + b * k.Nm) + t; c < __ldg(k.cptr + (b + 1) * k.Nm); c += 256) acc += k.cw[c] * k.cdist[c];
    float accb = 0.f;
    const float* th = f.theta + (size_t)b * NP;
    const float sc = th[3], cs = f.constant_scale;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gs = 0.f;
    for (int q = t; q < k.Nq; q += 256) {
        accb += k.part[(size_t)b * k.Nq + q];
        const float* g = k.dPw + 3 * ((size_t)b * k.Nq + q);
        const float w0 = weight * g[0], w1 = weight * g[1], w2 = weight * g[2];
        float* dv = f.dverts + (size_t)b * f.ld_v + 3 * (size_t)q * k.stride;
        const float* v = f.verts + (size_t)b * f.ld_v + 3 * (size_t)q * k.stride;
        dv[0] += w0 * sc * cs; dv[1] += w1 * sc * cs; dv[2] += w2 * sc * cs;
        g0 += w0 * sc * cs; g1 += w1 * sc * cs; g2 += w2 * sc * cs;
        gs += (w0 * (v[0] + th[0]) + w1 * (v[1] + th[1]) + w2 * (v[2] + th[2])) * cs;
    }
    float vals[5] = {acc + k.epsilon * accb, g0, g1, g2, gs};
#pragma unroll
    for (int i = 0; i < 5; ++i) { const float s = warp_sum(vals[i]); if (lane == 0) red[i][warp] = s; }
    __syncthreads();
    if (t < 5) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[t][w];
        if (t == 0) { f.loss[b] += weight * s; if (k.mask_loss) k.mask_loss[b] = s; }
        else f.grad[(size_t)b * NP + (t - 1)] += s;
    }
}
