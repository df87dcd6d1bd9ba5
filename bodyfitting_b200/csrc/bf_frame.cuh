// Fused per-frame kernel of the fitting loop on the active vertex set (one CTA per frame):
//
//   stage the frame's joint transforms A[b] and blended vertices v_posed (from the blend GEMM) in SHARED memory
//   -> LBS skinning of the frame's LIVE vertices (static picks / landmarks + the contour vertices of its yaw row)
//   -> output joints (chain joints, picked vertices, barycentric landmarks, contour landmarks)
//   -> keypoint term: world -> multi-view projection -> GMoF (smplify/loss.py:22-51,132-203), one thread per joint
//      walking its views, keypoints read as 16-byte vectors from the joint-major layout
//   -> gradient w.r.t. transl / scale / model joints -> gather by target into d(verts) (shared memory)
//   -> skinning backward: dvp = T_v^T dverts_v (written as the 3xTF32 operand split for the tensor-core blend
//      backward) and dA[j] = sum_v w_vj dverts_v (x) [vposed_v; 1] (per-vertex outer products staged once).
//
// Skinned vertices, d(verts), the blended transforms and the outer products never leave the chip; the unfused path
// (k_keypoint_loss, k_skin_rows, k_skin_bwd_dA) serves the all-vertex loop (silhouette / scan terms, operator surface).
#pragma once
#include "bf_common.cuh"
#include "bf_loss.cuh"

#define FR_THREADS 160      // launch shape of large batches (small ones: 224; BODYFIT_FRAME_THREADS=160 / 224 / 256 forces one)

// Dynamic shared-memory layout of k_frame_loss_bwd (float offsets, every region 16-byte aligned); host and device use the
// same function.  TMA = 1 adds the frame's keypoint row and two mbarriers.
struct FrameSmem {
    int gx, cam, red, As, dv, vp, outb, kp, bars, total;       // total in floats
};
__host__ __device__ __forceinline__ FrameSmem frame_smem_layout(int K, int Nv, int J, int ldn, int lmax, int tma) {
    FrameSmem L;
    const int K3 = (3 * K + 3) & ~3;
    L.gx = 0;                                   // [K*3] joint gradients
    L.cam = L.gx + K3;                          // [Nv*12]
    L.red = L.cam + Nv * 12;                    // [5*8]
    L.As = L.red + 64;                          // [J*12] this frame's joint transforms
    L.dv = L.As + ((J * 12 + 15) & ~15);        // [ldn] skinned vertices, later d(verts)
    L.vp = L.dv + ldn;                          // [ldn] v_posed
    L.outb = L.vp + ldn;                        // [lmax*12] blended transforms, later d(verts) (x) [v_posed; 1]
    L.kp = L.outb + 12 * lmax;                  // [K*Nv*3] keypoint row (TMA only)
    L.bars = L.kp + (tma == 1 ? ((K * Nv * 3 + 3) & ~3) : 0);       // tma == 2: transforms + v_posed by bulk copy, keypoints from global
    L.total = L.bars + (tma ? 4 : 0);
    return L;
}

// wait for a bulk copy: bounded spin, trap instead of hanging the GPU if the copy never lands
__device__ __forceinline__ void fr_wait(uint64_t* bar) {
    int spins = 0;
    while (!tc::mbar_try(bar, 0)) { if (++spins > (1 << 26)) __trap(); }
}

// 1 / x: hardware approximation (1 ulp) + one Newton step -> within 1 ulp of the correctly rounded value, 3 instructions
// and no slow path (__frcp_rn costs ~8 instructions and a call for special operands; the loss loop needs 3 per (joint, view))
__device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// T = sum_k w_k A[j_k] for vertex v (rows of A in shared memory); NC = 4: full 3x4 rows (forward), 3: rotation part (backward)
template <int NC>
__device__ __forceinline__ void blend_accum(const float* As, int j, float w, float (&T)[3 * NC]) {
    const float4* Aj = reinterpret_cast<const float4*>(As + j * 12);
    const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
    T[NC] = fmaf(w, r1.x, T[NC]); T[NC + 1] = fmaf(w, r1.y, T[NC + 1]); T[NC + 2] = fmaf(w, r1.z, T[NC + 2]);
    T[2 * NC] = fmaf(w, r2.x, T[2 * NC]); T[2 * NC + 1] = fmaf(w, r2.y, T[2 * NC + 1]); T[2 * NC + 2] = fmaf(w, r2.z, T[2 * NC + 2]);
    if (NC == 4) { T[3] = fmaf(w, r0.w, T[3]); T[7] = fmaf(w, r1.w, T[7]); T[11] = fmaf(w, r2.w, T[11]); }
}
template <int NC>
__device__ __forceinline__ void blend_transform(const BfVSet& vs, const float* As, int v, float (&T)[3 * NC]) {
    const int nnz = vs.nnz;
#pragma unroll
    for (int e = 0; e < 3 * NC; ++e) T[e] = 0.f;
    if (nnz == 4) {                                   // the usual skinning width: one 16-byte load each for joints and weights
        const int4 j4 = __ldg(reinterpret_cast<const int4*>(vs.ell_j) + v);
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(vs.ell_w) + v);
        blend_accum<NC>(As, j4.x, w4.x, T); blend_accum<NC>(As, j4.y, w4.y, T);
        blend_accum<NC>(As, j4.z, w4.z, T); blend_accum<NC>(As, j4.w, w4.w, T);
        return;
    }
    const int32_t* ej = vs.ell_j + (size_t)v * nnz;
    const float* ew = vs.ell_w + (size_t)v * nnz;
    for (int k = 0; k < nnz; ++k) blend_accum<NC>(As, __ldg(ej + k), __ldg(ew + k), T);
}

// Work is restricted to the frame's LIVE vertices: the static picks / landmarks plus the 17 x 3 contour vertices of
// the frame's yaw row (vs.lv_* / lt_* / lj_*, built per row on the host).  Every other vertex of the active set has
// an exactly-zero gradient for this frame, so it is neither skinned nor back-propagated; its dvp entries are zeroed.
// skin_here != 0: v_posed comes from the blend GEMM and the live vertices are skinned in this kernel (f.verts unused).
// TMA = 1: the frame's transforms, v_posed row and keypoint row (J*48 + 12*n + 12*K*Nv bytes, each one contiguous in
// HBM) are fetched by three bulk copies issued by one thread at kernel entry; the keypoints land while the vertices are
// being skinned, so the loss loop reads them from shared memory instead of waiting on HBM.
// Launch shape.  The kernel is bound by the latency of a frame's dependent phases, so what counts is how many frames an SM holds:
//   NT = 160, TMA = 2 (batches of >= 2048 frames): five warps per frame -- exactly the 135 joints of the loss loop, the kernel's longest phase; the live
//     vertices take two passes -- and the keypoint row is read from global memory instead of being staged (13 KB less shared memory):
//     27 KB per frame = EIGHT frames per SM at 48 registers.  Measured, 10,000 frames: 181.1 us; whole fit 48.1 ms.
//   NT = 224, TMA = 1 (smaller batches): seven warps, keypoints staged by a third bulk copy, five frames per SM, 56 registers:
//     189.0 us; fit 48.9 ms; at 1,250 frames 32.0 us against 37.8 us for the 160-thread shape (one wave of 1,184 + 66 frames).
//   NT = 256, TMA = 1: the round-1 shape, 48 registers: 191.8 us.
// Not kept (profiles/r2_frame_kernel_experiments.md): 192 threads (staged keypoints, five frames: 215.7 us; global keypoints, six frames
// at 56 registers: 208 us);
// six frames of 224 threads with the skinned vertices kept per live vertex (<= 40 registers, spills: 191 us); persistent CTAs with
// next-frame prefetch (203 us); a packed-fp32 (FFMA2) loss loop (197-232 us).
template <int TMA, int NT>
__global__ void __launch_bounds__(NT, NT == 160 ? 8 : 5) k_frame_loss_bwd(BfModel m, BfVSet vs, BfFrames f, int skin_here) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int K = m.K_used, Nv = f.Nv, J = m.J;
    const FrameSmem SL = frame_smem_layout(K, Nv, J, vs.ldn, vs.lmax, TMA);
    float* gx = sm + SL.gx;
    float* cam = sm + SL.cam;
    float* red = sm + SL.red;
    float* As = sm + SL.As;
    float* dv = sm + SL.dv;
    float* vp = sm + SL.vp;
    float4* outb = reinterpret_cast<float4*>(sm + SL.outb);
    const float* kps = sm + SL.kp;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SL.bars);
    const bool skin = skin_here != 0;
    const int n4 = (3 * vs.n + 3) / 4;                                  // float4s of a v_posed row; ld_v % 4 == 0 (checked on the host)
    if (TMA && t == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t a_bytes = (uint32_t)J * 48u, v_bytes = (uint32_t)n4 * 16u, k_bytes = (uint32_t)(K * Nv) * 12u;
        tc::mbar_expect_tx(&bars[0], a_bytes + v_bytes);
        tc::bulk_g2s(As, f.A + (size_t)b * J * 12, a_bytes, &bars[0]);
        tc::bulk_g2s(vp, f.vposed + (size_t)b * f.ld_v, v_bytes, &bars[0]);
        if (TMA == 1) {
            tc::mbar_expect_tx(&bars[1], k_bytes);
            tc::bulk_g2s(sm + SL.kp, f.kp + (size_t)(f.frame_index ? f.frame_index[b] : b) * K * Nv * 3, k_bytes, &bars[1]);
        }
    }
    if (TMA == 2 && t < K) {
        // keypoints are read from global memory in the loss loop: pull this thread's row (Nv x 12 bytes) towards L2 now (measured:
        // within noise either way -- eight resident frames hide the first touch)
        const char* kr = reinterpret_cast<const char*>(f.kp + ((size_t)(f.frame_index ? f.frame_index[b] : b) * K + t) * Nv * 3);
        for (int o = 0; o < Nv * 12; o += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(kr + o));
    }
    const int yaw = f.yaw ? f.yaw[b] : 0;
    const int row = vs.n_rows > 1 ? yaw : 0;
    const int L = __ldg(vs.lv_n + row);
    const int32_t* lv = vs.lv_vid + (size_t)row * vs.lmax;
    for (int i = t; i < Nv * 12; i += NT) cam[i] = f.cams[i];
    {
        if (!TMA) {
            const float4* src = reinterpret_cast<const float4*>(f.A + (size_t)b * J * 12);
            for (int i = t; i < J * 3; i += NT) reinterpret_cast<float4*>(As)[i] = src[i];
            const float4* vsrc = reinterpret_cast<const float4*>(f.vposed + (size_t)b * f.ld_v);
            for (int i = t; i < n4; i += NT) reinterpret_cast<float4*>(vp)[i] = vsrc[i];
        }
        if (!skin) {
            const float4* wsrc = reinterpret_cast<const float4*>(f.verts + (size_t)b * f.ld_v);
            for (int i = t; i < n4; i += NT) reinterpret_cast<float4*>(dv)[i] = wsrc[i];
        }
    }
    // clear this frame's d(v_posed) rows: only live vertices are written below (a vertex live on another yaw row in an
    // earlier iteration must read as zero in the blend backward GEMM).  With a block mask (BfFrames.blk_mask) the GEMM only
    // reads the 16-vertex blocks (12 float4 each) its 128-frame tile needs, so only those are cleared.
    const bool split = f.dvp_hi != nullptr;
    {
        const int ntile = (f.B + 127) >> 7;
        const uint32_t mask = f.blk_mask ? f.blk_mask[(size_t)(f.iter & 1) * ntile + (b >> 7)] : 0xFFFFFFFFu;
        if (f.blk_mask && t == 0 && (b & 127) == 0) f.blk_mask[(size_t)((f.iter + 1) & 1) * ntile + (b >> 7)] = 0u;   // next iteration's mask
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (split) {
            float4* oh = reinterpret_cast<float4*>(f.dvp_hi + (size_t)b * vs.ldn);
            float4* ol = reinterpret_cast<float4*>(f.dvp_lo + (size_t)b * vs.ldn);
            for (int i = t; i < n4; i += NT)
                if ((mask >> min(i / 12, 31)) & 1u) { oh[i] = z; ol[i] = z; }
        } else {
            float4* o = reinterpret_cast<float4*>(f.dvp + (size_t)b * f.ld_v);
            for (int i = t; i < n4; i += NT) o[i] = z;
        }
    }
    __syncthreads();
    if (TMA) fr_wait(&bars[0]);                      // transforms + v_posed have landed
    const float* th = f.theta + (size_t)b * m.NP;
    const float tx = th[0], ty = th[1], tz = th[2], sc = th[3];
    const float cs = f.constant_scale;
    const float icoef = 1024.0f / f.imsize;          // 1 / scale_coeff (exact for power-of-two image sizes)
    const float s2 = f.sigma * f.sigma;
    const float* Jtr_b = f.Jtr + (size_t)b * J * 3;
    const float invNv = 1.0f / (float)Nv;

    if (skin) {                                 // skin the live vertices: verts = (sum_k w_k A_jk) [v_posed; 1]
        for (int i = t; i < L; i += NT) {
            const int v = __ldg(lv + i);
            float T[12];
            blend_transform<4>(vs, As, v, T);
            const float px = vp[3 * v], py = vp[3 * v + 1], pz = vp[3 * v + 2];
            dv[3 * v] = T[0] * px + T[1] * py + T[2] * pz + T[3];
            dv[3 * v + 1] = T[4] * px + T[5] * py + T[6] * pz + T[7];
            dv[3 * v + 2] = T[8] * px + T[9] * py + T[10] * pz + T[11];
            // keep the blended transform for the backward (same thread, same slot; overwritten there by the outer product)
            outb[3 * i] = make_float4(T[0], T[1], T[2], T[3]);
            outb[3 * i + 1] = make_float4(T[4], T[5], T[6], T[7]);
            outb[3 * i + 2] = make_float4(T[8], T[9], T[10], T[11]);
        }
        __syncthreads();
    }
    // one thread per joint, its views in sequence: no cross-lane reduction, world coordinates computed once per joint,
    // and the joint-major keypoint rows ([B,K,Nv,3]) are read as 16-byte vectors, four views at a time
    float acc5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};               // loss, d/d transl (3), d/d scale
    const bool vec4 = (Nv & 3) == 0;
    if (TMA == 1) fr_wait(&bars[1]);                 // keypoint row (in flight since kernel entry)
    for (int k = t; k < K; k += NT) {
        float x[3];                                  // joint position (model space) + translation: q = x + T
        joint_pos(vs, k, yaw, Jtr_b, dv, x);
        const float qx = x[0] + tx, qy = x[1] + ty, qz = x[2] + tz;
        const float X = qx * sc * cs, Y = qy * sc * cs, Z = qz * sc * cs;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, ls = 0.f;
        const float* kpr = TMA == 1 ? kps + k * Nv * 3 : f.kp + ((size_t)(f.frame_index ? f.frame_index[b] : b) * K + k) * Nv * 3;
        for (int v0 = 0; v0 < Nv; v0 += 4) {
            float kv[12];
            if (vec4) {
                const float4* p4 = reinterpret_cast<const float4*>(kpr + v0 * 3);
                const float4 a = p4[0], bq = p4[1], c = p4[2];
                kv[0] = a.x; kv[1] = a.y; kv[2] = a.z; kv[3] = a.w; kv[4] = bq.x; kv[5] = bq.y; kv[6] = bq.z; kv[7] = bq.w;
                kv[8] = c.x; kv[9] = c.y; kv[10] = c.z; kv[11] = c.w;
            } else {
#pragma unroll
                for (int i = 0; i < 12; ++i) kv[i] = (v0 * 3 + i < Nv * 3) ? kpr[v0 * 3 + i] : 0.f;
            }
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
                const int v = v0 + u4;
                if (v >= Nv) break;
                const float kx = kv[3 * u4], ky = kv[3 * u4 + 1], wgt = kv[3 * u4 + 2];
                const float4* M4 = reinterpret_cast<const float4*>(cam + v * 12);
                const float4 m0 = M4[0], m1 = M4[1], m2 = M4[2];
                const float p0 = m0.x * X + m0.y * Y + m0.z * Z + m0.w;
                const float p1 = m1.x * X + m1.y * Y + m1.z * Z + m1.w;
                const float p2 = m2.x * X + m2.y * Y + m2.z * Z + m2.w;
                // three reciprocals instead of seven IEEE divisions: 1/p2, 1/(s^2+rx^2), 1/(s^2+ry^2)
                const float iz = rcp_nr(p2);
                const float u = p0 * iz, w_ = p1 * iz;
                const float rx = (kx - u) * icoef, ry = (ky - w_) * icoef;
                const float rx2 = rx * rx, ry2 = ry * ry;
                const float ix = rcp_nr(s2 + rx2), iy = rcp_nr(s2 + ry2);
                ls += wgt * (s2 * rx2 * ix + s2 * ry2 * iy);
                const float du = wgt * (2.0f * s2 * s2 * rx * ix * ix) * (-icoef);
                const float dw = wgt * (2.0f * s2 * s2 * ry * iy * iy) * (-icoef);
                const float dp0 = du * iz, dp1 = dw * iz, dp2 = -(du * u + dw * w_) * iz;
                g0 += m0.x * dp0 + m1.x * dp1 + m2.x * dp2;
                g1 += m0.y * dp0 + m1.y * dp1 + m2.y * dp2;
                g2 += m0.z * dp0 + m1.z * dp1 + m2.z * dp2;
            }
        }
        g0 *= invNv; g1 *= invNv; g2 *= invNv;
        const float k0s = sc * cs;
        gx[k * 3] = g0 * k0s; gx[k * 3 + 1] = g1 * k0s; gx[k * 3 + 2] = g2 * k0s;
        acc5[0] += ls;
        acc5[1] += g0 * k0s; acc5[2] += g1 * k0s; acc5[3] += g2 * k0s;
        acc5[4] += (g0 * qx + g1 * qy + g2 * qz) * cs;
    }
    const int nwk = (K + 31) >> 5;                   // warps that own joints; the others hold zeros
    if (warp < nwk) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float s = warp_sum(acc5[i]);
            if (lane == 0) red[i * 8 + warp] = s;
        }
    }
    __syncthreads();                               // gx + red complete
    if (t == 0) {
        float tot[5];
        for (int i = 0; i < 5; ++i) { float s = 0.f; for (int w = 0; w < nwk; ++w) s += red[i * 8 + w]; tot[i] = s; }
        f.loss[b] = tot[0] * invNv;
        float* g = f.grad + (size_t)b * m.NP;
        g[0] = tot[1]; g[1] = tot[2]; g[2] = tot[3]; g[3] = tot[4];
    }
    // joint gradients -> chain joints (static gather by target; the threads at the top of the block, idle below)
    float* dJtr_b = f.dJtr + (size_t)b * J * 3;
    for (int i = NT - 1 - t; i < J; i += NT) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int e0 = __ldg(vs.tg_ptr + i), e1 = __ldg(vs.tg_ptr + i + 1);
        for (int e = e0; e < e1; ++e) {
            const int k = __ldg(vs.tg_k + e);
            if (k < K) { const float w = __ldg(vs.tg_w + e); a0 += w * gx[k * 3]; a1 += w * gx[k * 3 + 1]; a2 += w * gx[k * 3 + 2]; }
        }
        dJtr_b[i * 3] = a0; dJtr_b[i * 3 + 1] = a1; dJtr_b[i * 3 + 2] = a2;
    }
    // live vertices: gather d(vertex) from the joints it feeds (per-row lists) and back-propagate it through the skinning
    // right away -- d(verts) stays in registers: dvp = (sum_k w_k A_jk)[:3,:3]^T dverts
    const int32_t* ltp = vs.lt_ptr + (size_t)row * (vs.lmax + 1);
    for (int i = t; i < L; i += NT) {
        float gx_ = 0.f, gy_ = 0.f, gz_ = 0.f;
        {
            const int e0 = __ldg(ltp + i), e1 = __ldg(ltp + i + 1);
            for (int e = e0; e < e1; ++e) {
                const int k = __ldg(vs.lt_k + e);
                if (k < K) { const float w = __ldg(vs.lt_w + e); gx_ += w * gx[k * 3]; gy_ += w * gx[k * 3 + 1]; gz_ += w * gx[k * 3 + 2]; }
            }
        }
        const int v = __ldg(lv + i);
        float T[9];
        if (skin) {                               // blended transform saved by the forward skinning above
            const float4 r0 = outb[3 * i], r1 = outb[3 * i + 1], r2 = outb[3 * i + 2];
            T[0] = r0.x; T[1] = r0.y; T[2] = r0.z; T[3] = r1.x; T[4] = r1.y; T[5] = r1.z; T[6] = r2.x; T[7] = r2.y; T[8] = r2.z;
        } else {
            blend_transform<3>(vs, As, v, T);
        }
        const float o0 = T[0] * gx_ + T[3] * gy_ + T[6] * gz_;
        const float o1 = T[1] * gx_ + T[4] * gy_ + T[7] * gz_;
        const float o2 = T[2] * gx_ + T[5] * gy_ + T[8] * gz_;
        if (split) {
            float* oh = f.dvp_hi + (size_t)b * vs.ldn + 3 * v;
            float* ol = f.dvp_lo + (size_t)b * vs.ldn + 3 * v;
            split_tf32(o0, oh[0], ol[0]); split_tf32(o1, oh[1], ol[1]); split_tf32(o2, oh[2], ol[2]);
        } else {
            float* o = f.dvp + (size_t)b * f.ld_v + 3 * v;
            o[0] = o0; o[1] = o1; o[2] = o2;
        }
        // the vertex's outer product, once; the joint side below only scales and adds it
        const float px = vp[3 * v], py = vp[3 * v + 1], pz = vp[3 * v + 2];
        outb[3 * i] = make_float4(gx_ * px, gx_ * py, gx_ * pz, gx_);
        outb[3 * i + 1] = make_float4(gy_ * px, gy_ * py, gy_ * pz, gy_);
        outb[3 * i + 2] = make_float4(gz_ * px, gz_ * py, gz_ * pz, gz_);
    }
    // skinning backward, joint side: dA[j] = sum_{live v} w_vj dverts_v (x) [vposed_v; 1]
    float* dAb = f.dA + (size_t)b * J * 12;
    if (vs.n_nz < J) {                                 // joints without live vertices keep zero rows (16-byte stores)
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = t; i < J * 3; i += NT) reinterpret_cast<float4*>(dAb)[i] = z;
    }
    __syncthreads();
    // 4 lanes per joint (the live skinning lists are short), 8 joints per warp at a time; the 12 (padded 16) partial sums
    // are reduced over the 4 lanes with a multi-value butterfly (8 + 4 shuffles), after which lane l of the group holds
    // elements 4l .. 4l+3 = one row of dA[j] -> one 16-byte store
    const int32_t* ljp = vs.lj_ptr + (size_t)row * (vs.n_nz + 1);
    for (int jn0 = 0; jn0 < vs.n_nz; jn0 += 8 * (NT / 32)) {
        if (jn0 + warp * 8 >= vs.n_nz) break;                  // warp-uniform
        const int jn = jn0 + warp * 8 + (lane >> 2);
        const bool jv_ = jn < vs.n_nz;
        const int j = jv_ ? __ldg(vs.jv_nz + jn) : 0;
        const int sl = lane & 3;
        float acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        const int e0 = jv_ ? __ldg(ljp + jn) : 0, e1 = jv_ ? __ldg(ljp + jn + 1) : 0;
        for (int e = e0 + sl; e < e1; e += 4) {
            const float4* o = outb + 3 * __ldg(vs.lj_vid + e);          // lj_vid: index into the row's live list
            const float w = __ldg(vs.lj_w + e);
            const float4 o0 = o[0], o1 = o[1], o2 = o[2];
            acc[0] = fmaf(w, o0.x, acc[0]); acc[1] = fmaf(w, o0.y, acc[1]); acc[2] = fmaf(w, o0.z, acc[2]); acc[3] = fmaf(w, o0.w, acc[3]);
            acc[4] = fmaf(w, o1.x, acc[4]); acc[5] = fmaf(w, o1.y, acc[5]); acc[6] = fmaf(w, o1.z, acc[6]); acc[7] = fmaf(w, o1.w, acc[7]);
            acc[8] = fmaf(w, o2.x, acc[8]); acc[9] = fmaf(w, o2.y, acc[9]); acc[10] = fmaf(w, o2.z, acc[10]); acc[11] = fmaf(w, o2.w, acc[11]);
        }
        float a8[8], a4[4];
        bool hi = lane & 2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float recv = __shfl_xor_sync(0xffffffffu, hi ? acc[i] : acc[i + 8], 2);
            a8[i] = (hi ? acc[i + 8] : acc[i]) + recv;
        }
        hi = lane & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float recv = __shfl_xor_sync(0xffffffffu, hi ? a8[i] : a8[i + 4], 1);
            a4[i] = (hi ? a8[i + 4] : a8[i]) + recv;
        }
        // this lane now holds elements 4*sl .. 4*sl+3 of joint j (sl = 3: the padding)
        if (jv_ && sl < 3) *reinterpret_cast<float4*>(dAb + j * 12 + 4 * sl) = make_float4(a4[0], a4[1], a4[2], a4[3]);
    }
}
