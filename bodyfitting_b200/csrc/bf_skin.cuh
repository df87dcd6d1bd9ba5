// Blend-shape contraction + linear blend skinning over a vertex set, forward and backward.
//
//   k_skin_fwd   : v_posed[b, :] = pf[b, :] @ Bm  (shape + pose blend shapes + template in ONE
//                  dense contraction, K = P + NS + 1), then in the epilogue, with the v_posed tile
//                  still in registers, verts = (sum_k w_k A[b, j_k]) [v_posed; 1].
//                  FP32 FFMA tile kernel (64 frames x 32 vertices per CTA, K chunks of 16 staged in
//                  shared memory, double buffered).  The tcgen05 / 3xTF32 variant lives in
//                  bf_blend_tc.cuh and shares the epilogue.
//   k_skin_bwd_dvp : dvp = T_v[:3,:3]^T dverts            (thread per vertex, FB frames each)
//   k_skin_bwd_dA  : dA[b,j] = sum_v w_vj dverts_v (x) [v_posed_v; 1]   (warp per (frame, joint),
//                    gather over a CSR-by-joint list -> deterministic, no atomics)
//   k_blend_bwd    : dpf = dvp @ Bm^T  (64x64 tiles, reduction over the vertex coordinates)
//
// Replaces smplx.lbs.lbs (blend_shapes, pose_offsets matmul, skinning matmul) and its autograd
// backward; reference call sites models/smpl.py:71, smplify/smplify.py:179-187.
#pragma once
#include <string.h>
#include "bf_common.cuh"
#include "bf_loss.cuh"

#define SK_TB 64      // frames per CTA
#define SK_TV 32      // vertices per CTA
#define SK_TN 96      // coordinates per CTA
#define SK_KC 16      // K chunk
#define SK_LDA 68     // padded frame stride of the transposed A tile (floats)

__global__ void __launch_bounds__(256) k_skin_fwd(BfVSet vs, int J, int Kp, const float* __restrict__ pf,
                                                  const float* __restrict__ A, float* __restrict__ verts,
                                                  float* __restrict__ vposed, int B, int ld_v,
                                                  const float* __restrict__ theta, int NP, float cs) {
    __shared__ __align__(16) float As[2][SK_KC][SK_LDA];
    __shared__ __align__(16) float Bs[2][SK_KC][SK_TN];
    const int t = threadIdx.x;
    const int tf = t >> 5, tv = t & 31;
    const int b0 = blockIdx.y * SK_TB;
    const int n0 = blockIdx.x * SK_TN;
    const int ldn = vs.ldn;

    // global -> register staging maps
    const int ar = t >> 2, akq = (t & 3) * 4;
    int arow = b0 + ar; if (arow >= B) arow = B - 1;
    const float* aptr = pf + (size_t)arow * Kp + akq;
    const int bk0 = t / 24, bc0 = (t % 24) * 4;
    const int t2 = t + 256;
    const int bk1 = t2 / 24, bc1 = (t2 % 24) * 4;
    const bool has2 = t2 < (SK_KC * SK_TN / 4);
    const float* bptr0 = vs.Bm + (size_t)bk0 * ldn + n0 + bc0;
    const float* bptr1 = vs.Bm + (size_t)bk1 * ldn + n0 + bc1;

    float acc[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; }

    float4 ra = *reinterpret_cast<const float4*>(aptr);
    float4 rb0 = __ldg(reinterpret_cast<const float4*>(bptr0));
    float4 rb1 = has2 ? __ldg(reinterpret_cast<const float4*>(bptr1)) : make_float4(0, 0, 0, 0);
    As[0][akq + 0][ar] = ra.x; As[0][akq + 1][ar] = ra.y; As[0][akq + 2][ar] = ra.z; As[0][akq + 3][ar] = ra.w;
    *reinterpret_cast<float4*>(&Bs[0][bk0][bc0]) = rb0;
    if (has2) *reinterpret_cast<float4*>(&Bs[0][bk1][bc1]) = rb1;
    __syncthreads();

    const int nk = Kp / SK_KC;
    for (int kc = 0; kc < nk; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nk) {
            const size_t ko = (size_t)(kc + 1) * SK_KC;
            ra = *reinterpret_cast<const float4*>(aptr + ko);
            rb0 = __ldg(reinterpret_cast<const float4*>(bptr0 + ko * ldn));
            if (has2) rb1 = __ldg(reinterpret_cast<const float4*>(bptr1 + ko * ldn));
        }
#pragma unroll
        for (int k = 0; k < SK_KC; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][tf * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][tf * 8 + 4]);
            const float bx = Bs[cur][k][tv * 3 + 0], by = Bs[cur][k][tv * 3 + 1], bz = Bs[cur][k][tv * 3 + 2];
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(av[i], bx, acc[i][0]);
                acc[i][1] = fmaf(av[i], by, acc[i][1]);
                acc[i][2] = fmaf(av[i], bz, acc[i][2]);
            }
        }
        if (kc + 1 < nk) {
            const int nxt = cur ^ 1;
            As[nxt][akq + 0][ar] = ra.x; As[nxt][akq + 1][ar] = ra.y; As[nxt][akq + 2][ar] = ra.z; As[nxt][akq + 3][ar] = ra.w;
            *reinterpret_cast<float4*>(&Bs[nxt][bk0][bc0]) = rb0;
            if (has2) *reinterpret_cast<float4*>(&Bs[nxt][bk1][bc1]) = rb1;
        }
        __syncthreads();
    }

    // epilogue: skinning with the v_posed micro-tile in registers
    const int v = blockIdx.x * SK_TV + tv;
    if (v >= vs.n) return;
    const int nnz = vs.nnz;
    const int32_t* ej = vs.ell_j + (size_t)v * nnz;
    const float* ew = vs.ell_w + (size_t)v * nnz;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int b = b0 + tf * 8 + i;
        if (b >= B) break;
        const float px = acc[i][0], py = acc[i][1], pz = acc[i][2];
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        const float* Ab = A + (size_t)b * J * 12;
        for (int k = 0; k < nnz; ++k) {
            const float w = __ldg(ew + k);
            const float4* Aj = reinterpret_cast<const float4*>(Ab + __ldg(ej + k) * 12);
            const float4 r0 = __ldg(Aj), r1 = __ldg(Aj + 1), r2 = __ldg(Aj + 2);
            T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
            T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
            T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
        }
        float* o = verts + (size_t)b * ld_v + 3 * v;
        float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
        float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
        float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
        if (theta) {                      // world = (x + transl) * scale * constant_scale
            const float* th = theta + (size_t)b * NP;
            const float sc = __ldg(th + 3);
            ox = (ox + __ldg(th + 0)) * sc * cs; oy = (oy + __ldg(th + 1)) * sc * cs; oz = (oz + __ldg(th + 2)) * sc * cs;
        }
        o[0] = ox; o[1] = oy; o[2] = oz;
        if (vposed) {
            float* q = vposed + (size_t)b * ld_v + 3 * v;
            q[0] = px; q[1] = py; q[2] = pz;
        }
    }
}

#define DV_FB 4
__global__ void __launch_bounds__(256) k_skin_bwd_dvp(BfVSet vs, int J, const float* __restrict__ A,
                                                      const float* __restrict__ dverts, float* __restrict__ dvp,
                                                      int B, int ld_v, float* __restrict__ dvp_hi, float* __restrict__ dvp_lo) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vs.n) return;
    const int nnz = vs.nnz;
    const int32_t* ej = vs.ell_j + (size_t)v * nnz;
    const float* ew = vs.ell_w + (size_t)v * nnz;
    for (int i = 0; i < DV_FB; ++i) {
        const int b = blockIdx.y * DV_FB + i;
        if (b >= B) break;
        const float* g = dverts + (size_t)b * ld_v + 3 * v;
        const float gx = g[0], gy = g[1], gz = g[2];
        float T[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) T[e] = 0.f;
        const float* Ab = A + (size_t)b * J * 12;
        for (int k = 0; k < nnz; ++k) {
            const float w = __ldg(ew + k);
            const float4* Aj = reinterpret_cast<const float4*>(Ab + __ldg(ej + k) * 12);
            const float4 r0 = __ldg(Aj), r1 = __ldg(Aj + 1), r2 = __ldg(Aj + 2);
            T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
            T[3] = fmaf(w, r1.x, T[3]); T[4] = fmaf(w, r1.y, T[4]); T[5] = fmaf(w, r1.z, T[5]);
            T[6] = fmaf(w, r2.x, T[6]); T[7] = fmaf(w, r2.y, T[7]); T[8] = fmaf(w, r2.z, T[8]);
        }
        float* o = dvp + (size_t)b * ld_v + 3 * v;
        o[0] = T[0] * gx + T[3] * gy + T[6] * gz;
        o[1] = T[1] * gx + T[4] * gy + T[7] * gz;
        o[2] = T[2] * gx + T[5] * gy + T[8] * gz;
        if (dvp_hi) {                      // 3xTF32 operand split for the tensor-core backward GEMM
            float* oh = dvp_hi + (size_t)b * vs.ldn + 3 * v;
            float* ol = dvp_lo + (size_t)b * vs.ldn + 3 * v;
            split_tf32(o[0], oh[0], ol[0]); split_tf32(o[1], oh[1], ol[1]); split_tf32(o[2], oh[2], ol[2]);
        }
    }
}

__global__ void __launch_bounds__(256) k_skin_bwd_dA(BfVSet vs, int J, const float* __restrict__ dverts,
                                                     const float* __restrict__ vposed, float* __restrict__ dA,
                                                     int B, int ld_v) {
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* g = dverts + (size_t)b * ld_v;
    const float* x = vposed + (size_t)b * ld_v;
    // joints without skinned vertices in this set (most of them for the active set): zero rows
    if (vs.n_nz < J) {
        for (int i = threadIdx.x; i < J * 12; i += blockDim.x) dA[(size_t)b * J * 12 + i] = 0.f;
        __syncthreads();
    }
    for (int jn = warp; jn < vs.n_nz; jn += nw) {
        const int j = __ldg(vs.jv_nz + jn);
        float acc[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) acc[e] = 0.f;
        const int e0 = __ldg(vs.jv_ptr + j), e1 = __ldg(vs.jv_ptr + j + 1);
        for (int e = e0 + lane; e < e1; e += 32) {
            const int v = __ldg(vs.jv_vid + e);
            const float w = __ldg(vs.jv_w + e);
            const float gx = w * g[3 * v], gy = w * g[3 * v + 1], gz = w * g[3 * v + 2];
            const float px = x[3 * v], py = x[3 * v + 1], pz = x[3 * v + 2];
            acc[0] = fmaf(gx, px, acc[0]); acc[1] = fmaf(gx, py, acc[1]); acc[2] = fmaf(gx, pz, acc[2]); acc[3] += gx;
            acc[4] = fmaf(gy, px, acc[4]); acc[5] = fmaf(gy, py, acc[5]); acc[6] = fmaf(gy, pz, acc[6]); acc[7] += gy;
            acc[8] = fmaf(gz, px, acc[8]); acc[9] = fmaf(gz, py, acc[9]); acc[10] = fmaf(gz, pz, acc[10]); acc[11] += gz;
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) acc[e] = warp_sum(acc[e]);
        if (lane < 12) {
            float val = acc[0];
#pragma unroll
            for (int e = 1; e < 12; ++e) if (lane == e) val = acc[e];
            dA[((size_t)b * J + j) * 12 + lane] = val;
        }
    }
}

#define GB_T 64
#define GB_KC 16
#define GB_LD 68
// dpf[b, k] = sum_n dvp[b, n] * Bm[k, n],  n < 3 * vs.n
__global__ void __launch_bounds__(256) k_blend_bwd(BfVSet vs, int Kp, const float* __restrict__ dvp,
                                                   float* __restrict__ dpf, int B, int ld_v) {
    __shared__ __align__(16) float Xs[2][GB_KC][GB_LD];
    __shared__ __align__(16) float Ys[2][GB_KC][GB_LD];
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int b0 = blockIdx.y * GB_T, k0 = blockIdx.x * GB_T;
    const int nvalid = 3 * vs.n;
    const int ldn = vs.ldn;
    const int r = t >> 2, nq = (t & 3) * 4;
    int xrow = b0 + r; if (xrow >= B) xrow = B - 1;
    int yrow = k0 + r; if (yrow >= Kp) yrow = Kp - 1;
    const float* xp = dvp + (size_t)xrow * ld_v + nq;
    const float* yp = vs.Bm + (size_t)yrow * ldn + nq;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nchunks = (nvalid + GB_KC - 1) / GB_KC;
    float rx[4];
    float4 ry;
    auto load = [&](int c) {
        const int n = c * GB_KC + nq;
#pragma unroll
        for (int i = 0; i < 4; ++i) rx[i] = (n + i < nvalid) ? xp[(size_t)c * GB_KC + i] : 0.f;
        ry = __ldg(reinterpret_cast<const float4*>(yp + (size_t)c * GB_KC));   // ldn >= padded extent, pad columns are 0
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) Xs[buf][nq + i][r] = rx[i];
        Ys[buf][nq + 0][r] = ry.x; Ys[buf][nq + 1][r] = ry.y; Ys[buf][nq + 2][r] = ry.z; Ys[buf][nq + 3][r] = ry.w;
    };
    load(0); store(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int cur = c & 1;
        if (c + 1 < nchunks) load(c + 1);
#pragma unroll
        for (int n = 0; n < GB_KC; ++n) {
            const float4 xv = *reinterpret_cast<const float4*>(&Xs[cur][n][ty * 4]);
            const float4 yv = *reinterpret_cast<const float4*>(&Ys[cur][n][tx * 4]);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            const float ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
        }
        if (c + 1 < nchunks) store(cur ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = b0 + ty * 4 + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < Kp) dpf[(size_t)b * Kp + k] = acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_skin_rows : verts = (sum_k w_k A[b, j_k]) [v_posed; 1] for 32 frames x a slab of vertices per CTA, with the
// joint transforms of the 32 frames resident in SHARED memory in frame-minor order As[(j*3 + r)][frame] (float4 rows).
//
// Why this shape: skinning gathers 4 x 48 B of transforms per vertex and frame -- 16x the 12 B it writes.  Gathered
// from L2 that traffic bounds the kernel far below the HBM roofline (measured: the GEMM-epilogue version of this step
// ran at ~37 us per 128x64 tile whatever the GEMM depth), so the transforms have to be on chip, and they are reused by
// every vertex of the sweep.  LANE = FRAME: the influences (joint, weight) of a vertex are warp-uniform and one
// LDS.128 of the warp reads 32 consecutive frames of one transform row -- conflict-free for any skinning pattern.
// v_posed / verts rows are transposed through a per-warp [48 coords][32 frames] tile so that global accesses are
// row-contiguous.  In place (verts == vposed) is allowed: a warp reads its 16-vertex segment completely before it
// writes it.
#define SR_WARPS 16
#define SR_LD 33
#define SR_TILE (48 * SR_LD + 128)     // per-warp floats: [48 coords][32 frames] tile + the chunk's ELL rows (16 x (4 joints | 4 weights))

// MODE 0: forward  verts = T [v_posed; 1]  (optionally -> world space)
// MODE 1: backward dvp   = T[:3,:3]^T dverts, written as the 3xTF32 split (out_hi / out_lo, row stride ld_out) and / or plain
template <int MODE>
__global__ void __launch_bounds__(32 * SR_WARPS, 1)
k_skin_rows(BfVSet vs, int J, const float* __restrict__ A, const float* in, float* out, float* out_hi, float* out_lo,
            int B, int ld_v, int ld_out, const float* __restrict__ theta, int NP, float cs, int chunks_per_slab) {
    extern __shared__ __align__(16) float sm_sr[];
    float4* As = reinterpret_cast<float4*>(sm_sr);                          // [3J][32]
    float* st = sm_sr + (size_t)3 * J * 32 * 4 + (threadIdx.x >> 5) * SR_TILE;
    float* ell = st + 48 * SR_LD;                                           // 16B-aligned: 48*33*4 = 6336 = 16 * 396
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * 32;
    const int nrows = min(32, B - b0);
    {   // A[b][j][r] (float4) -> As[j*3 + r][frame]: consecutive threads take consecutive frames (conflict-free stores)
        const float4* src = reinterpret_cast<const float4*>(A);
        const int n = 3 * J;
        for (int i = threadIdx.x; i < 32 * n; i += blockDim.x) {
            const int jr = i >> 5, fl = i & 31;
            const int b = min(b0 + fl, B - 1);
            As[i] = __ldg(src + (size_t)b * n + jr);
        }
    }
    const int bl = min(b0 + lane, B - 1);
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, sc = 1.f;
    if (MODE == 0 && theta) {                                               // world = (x + transl) * scale * cs
        const float* th = theta + (size_t)bl * NP;
        t0 = __ldg(th); t1 = __ldg(th + 1); t2 = __ldg(th + 2); sc = __ldg(th + 3);
    }
    __syncthreads();
    const int nnz = vs.nnz;
    const int n_chunks = (vs.n + 15) >> 4;
    const int c_end = min(n_chunks, (int)(blockIdx.y + 1) * chunks_per_slab);
    // this lane's three (row-of-the-pair, coord) slots of a 2-row x 48-coord step: global / tile offsets
    const int hi1 = lane >= 16, c1 = lane + 32 - 48 * hi1;
    const size_t g_off0 = (size_t)lane, g_off1 = (size_t)hi1 * ld_v + c1, g_off2 = (size_t)ld_v + lane + 16;
    const size_t h_off0 = (size_t)lane, h_off1 = (size_t)hi1 * ld_out + c1, h_off2 = (size_t)ld_out + lane + 16;
    const int s_off0 = lane * SR_LD, s_off1 = c1 * SR_LD + hi1, s_off2 = (lane + 16) * SR_LD + 1;
    for (int ch = blockIdx.y * chunks_per_slab + warp; ch < c_end; ch += SR_WARPS) {
        const int vbase = ch << 4;
        const int ncols = 3 * min(16, vs.n - vbase);
        const float* src = in + (size_t)b0 * ld_v + 3 * vbase;
        if (nnz == 4) {                                                      // the chunk's influences -> shared (one 16-byte load per lane)
            const int vi = lane & 15;
            const int v = min(vbase + vi, vs.n - 1);
            if (lane < 16) reinterpret_cast<int4*>(ell)[vi] = __ldg(reinterpret_cast<const int4*>(vs.ell_j) + v);
            else reinterpret_cast<float4*>(ell)[16 + vi] = __ldg(reinterpret_cast<const float4*>(vs.ell_w) + v);
        }
        // rows -> tile: 2 rows x 48 coords per three warp-wide loads; the lane's three (row, coord) slots are fixed, so a
        // full chunk is just pointer increments
        const bool full_chunk = nrows == 32 && ncols == 48;
        if (full_chunk) {
            const float* p = src;
#pragma unroll 8
            for (int rp = 0; rp < 16; ++rp, p += 2 * (size_t)ld_v) {
                const float x0 = p[g_off0], x1 = p[g_off1], x2 = p[g_off2];
                st[s_off0 + 2 * rp] = x0; st[s_off1 + 2 * rp] = x1; st[s_off2 + 2 * rp] = x2;
            }
        } else {
#pragma unroll 4
            for (int rp = 0; rp < 16; ++rp) {
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int idx = lane + 32 * s, hi = idx >= 48;
                    const int fr = 2 * rp + hi, c = idx - 48 * hi;
                    if (fr < nrows && c < ncols) st[c * SR_LD + fr] = src[(size_t)fr * ld_v + c];
                }
            }
        }
        __syncwarp();
#pragma unroll 2
        for (int vi = 0; vi < 16; ++vi) {
            constexpr int NC = MODE == 0 ? 4 : 3;
            float T[3 * NC];
#pragma unroll
            for (int e = 0; e < 3 * NC; ++e) T[e] = 0.f;
            if (nnz == 4) {
                const int4 j4 = reinterpret_cast<const int4*>(ell)[vi];
                const float4 w4 = reinterpret_cast<const float4*>(ell)[16 + vi];
                const int jj[4] = {j4.x, j4.y, j4.z, j4.w};
                const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float w = ww[k];
                    const float4* Aj = As + jj[k] * 96 + lane;
                    const float4 r0 = Aj[0], r1 = Aj[32], r2 = Aj[64];
                    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
                    T[NC] = fmaf(w, r1.x, T[NC]); T[NC + 1] = fmaf(w, r1.y, T[NC + 1]); T[NC + 2] = fmaf(w, r1.z, T[NC + 2]);
                    T[2 * NC] = fmaf(w, r2.x, T[2 * NC]); T[2 * NC + 1] = fmaf(w, r2.y, T[2 * NC + 1]); T[2 * NC + 2] = fmaf(w, r2.z, T[2 * NC + 2]);
                    if (NC == 4) { T[3] = fmaf(w, r0.w, T[3]); T[7] = fmaf(w, r1.w, T[7]); T[11] = fmaf(w, r2.w, T[11]); }
                }
            } else {
                const int v = min(vbase + vi, vs.n - 1);                     // warp-uniform; clamped pad vertices are not stored
                const int32_t* ej = vs.ell_j + (size_t)v * nnz;
                const float* ew = vs.ell_w + (size_t)v * nnz;
                for (int k = 0; k < nnz; ++k) {
                    const float w = __ldg(ew + k);
                    const float4* Aj = As + __ldg(ej + k) * 96 + lane;
                    const float4 r0 = Aj[0], r1 = Aj[32], r2 = Aj[64];
                    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
                    T[NC] = fmaf(w, r1.x, T[NC]); T[NC + 1] = fmaf(w, r1.y, T[NC + 1]); T[NC + 2] = fmaf(w, r1.z, T[NC + 2]);
                    T[2 * NC] = fmaf(w, r2.x, T[2 * NC]); T[2 * NC + 1] = fmaf(w, r2.y, T[2 * NC + 1]); T[2 * NC + 2] = fmaf(w, r2.z, T[2 * NC + 2]);
                    if (NC == 4) { T[3] = fmaf(w, r0.w, T[3]); T[7] = fmaf(w, r1.w, T[7]); T[11] = fmaf(w, r2.w, T[11]); }
                }
            }
            float* q = st + (3 * vi) * SR_LD + lane;
            const float px = q[0], py = q[SR_LD], pz = q[2 * SR_LD];
            float ox, oy, oz;
            if (MODE == 0) {
                ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
                oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
                oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
                if (theta) { ox = (ox + t0) * sc * cs; oy = (oy + t1) * sc * cs; oz = (oz + t2) * sc * cs; }
            } else {
                ox = T[0] * px + T[3] * py + T[6] * pz;
                oy = T[1] * px + T[4] * py + T[7] * pz;
                oz = T[2] * px + T[5] * py + T[8] * pz;
            }
            q[0] = ox; q[SR_LD] = oy; q[2 * SR_LD] = oz;
        }
        __syncwarp();
        if (full_chunk) {
            if (out) {
                float* p = out + (size_t)b0 * ld_v + 3 * vbase;
#pragma unroll 8
                for (int rp = 0; rp < 16; ++rp, p += 2 * (size_t)ld_v) {
                    p[g_off0] = st[s_off0 + 2 * rp]; p[g_off1] = st[s_off1 + 2 * rp]; p[g_off2] = st[s_off2 + 2 * rp];
                }
            }
            if (MODE == 1 && out_hi) {
                const size_t base = (size_t)b0 * ld_out + 3 * vbase;
                float* ph = out_hi + base;
                float* pl = out_lo + base;
#pragma unroll 8
                for (int rp = 0; rp < 16; ++rp, ph += 2 * (size_t)ld_out, pl += 2 * (size_t)ld_out) {
                    float h_, l_;
                    split_tf32(st[s_off0 + 2 * rp], h_, l_); ph[h_off0] = h_; pl[h_off0] = l_;
                    split_tf32(st[s_off1 + 2 * rp], h_, l_); ph[h_off1] = h_; pl[h_off1] = l_;
                    split_tf32(st[s_off2 + 2 * rp], h_, l_); ph[h_off2] = h_; pl[h_off2] = l_;
                }
            }
        } else {
#pragma unroll 4
            for (int rp = 0; rp < 16; ++rp) {
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int idx = lane + 32 * s, hi = idx >= 48;
                    const int fr = 2 * rp + hi, c = idx - 48 * hi;
                    if (fr < nrows && c < ncols) {
                        const float val = st[c * SR_LD + fr];
                        if (out) out[(size_t)(b0 + fr) * ld_v + 3 * vbase + c] = val;
                        if (MODE == 1 && out_hi) {
                            float h_, l_;
                            split_tf32(val, h_, l_);
                            const size_t o = (size_t)(b0 + fr) * ld_out + 3 * vbase + c;
                            out_hi[o] = h_; out_lo[o] = l_;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

static int bf_launch_skin_rows(int mode, const BfVSet* vs, int J, const float* A, const float* in, float* out, float* out_hi,
                               float* out_lo, int B, int ld_v, int ld_out, const float* theta, int NP, float cs, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)3 * J * 32 * 4 + (size_t)SR_WARPS * SR_TILE);
    static size_t attr[2][BF_MAXDEV] = {{0}};
    {
        const int rc = mode == 0 ? bf_ensure_smem(k_skin_rows<0>, smem, attr[0], "k_skin_rows<0>")
                                 : bf_ensure_smem(k_skin_rows<1>, smem, attr[1], "k_skin_rows<1>");
        if (rc) return rc;
    }
    const int num_sms = bf_num_sms();
    // 32 frames per CTA (one CTA per SM); the vertex range is cut into slabs only while that is needed to fill the SMs:
    // at most two full rounds of CTAs, each slab at least one 16-vertex chunk per warp
    const int groups = (B + 31) / 32, n_chunks = (vs->n + 15) / 16;
    int slabs = (2 * num_sms) / groups;
    const int max_slabs = (n_chunks + SR_WARPS - 1) / SR_WARPS;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs < 1) slabs = 1;
    const int cps = (n_chunks + slabs - 1) / slabs;
    const dim3 grid(groups, (n_chunks + cps - 1) / cps);
    if (mode == 0) k_skin_rows<0><<<grid, 32 * SR_WARPS, smem, s>>>(*vs, J, A, in, out, out_hi, out_lo, B, ld_v, ld_out, theta, NP, cs, cps);
    else k_skin_rows<1><<<grid, 32 * SR_WARPS, smem, s>>>(*vs, J, A, in, out, out_hi, out_lo, B, ld_v, ld_out, theta, NP, cs, cps);
    BF_LAUNCH_CHECK();
    return BF_OK;
}


// ---------------------------------------------------------------------------------------------
// k_skin_frame : the same skinning with FOUR FRAMES per CTA and LANE = VERTEX -- the all-vertex operator's default.
//
// k_skin_rows (above) keeps 32 frames' transforms in shared memory and maps lane = frame, which makes every transform read
// a conflict-free broadcast but forces each row segment through a shared-memory transpose (42 % of its instructions) and
// limits it to one 145 KB CTA per SM.  Here a CTA stages only its four frames' J x 12 transforms (4.6 - 10.6 KB), a warp takes 32
// consecutive vertices = 96 consecutive floats of the row (three fully coalesced 128-byte accesses, de-interleaved through a
// 384-byte per-warp stage: stride-3 reads are conflict-free), and each lane blends its own vertex' four transforms with
// LDS.128 (neighbouring vertices mostly share joints -> mostly broadcasts); the vertex' four (joint, weight) pairs are loaded once
// and reused for the CTA's frames (one frame per CTA re-read the 32-byte table entry per frame: 226 MB of L2 traffic for 170 MB
// of vertices), and the next frame's floats are requested before the current frame is computed.  ~6 CTAs per SM, no transposes.
// With one slab per frame group (gridDim.y == 1) the output joints are produced by the same CTA right after its vertices.
//   MODE 0: verts = T [v_posed; 1] (optionally -> world space);   MODE 1: dvp = T[:3,:3]^T dverts (plain and / or 3xTF32 split)
#define SF_THREADS 256
#define SF_FPB 4                      // frames per CTA: the skinning table of a vertex group is read once for all of them
template <int MODE>
__global__ void __launch_bounds__(SF_THREADS) k_skin_frame(BfModel m, BfVSet vs, BfFrames f, const float* in, float* out,
                                                           float* out_hi, float* out_lo, int ld_out, int world, int fuse_joints) {
    __shared__ __align__(16) float As[SF_FPB][BF_MAXJ * 12];
    __shared__ float stage[SF_THREADS / 32][96];
    const int b0 = blockIdx.x * SF_FPB, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nfr = min(SF_FPB, f.B - b0);
    const int J = m.J, n = vs.n, ld_v = f.ld_v;
    for (int fr = 0; fr < nfr; ++fr) {
        const float4* src = reinterpret_cast<const float4*>(f.A + (size_t)(b0 + fr) * J * 12);
        for (int i = t; i < J * 3; i += SF_THREADS) reinterpret_cast<float4*>(As[fr])[i] = src[i];
    }
    const float cs = f.constant_scale;
    __syncthreads();
    float* st = stage[warp];
    const int groups = (n + 31) >> 5;                                 // 32-vertex groups of the row
    const int g_per = (groups + gridDim.y - 1) / gridDim.y;
    const int g_end = min(groups, (int)(blockIdx.y + 1) * g_per);
    const int nnz = vs.nnz;
    constexpr int NC = MODE == 0 ? 4 : 3;
    for (int g = blockIdx.y * g_per + warp; g < g_end; g += SF_THREADS / 32) {
        const int v0 = g << 5, nf = 3 * min(32, n - v0);              // floats of this group
        const int v = min(v0 + lane, n - 1);
        int jj[4] = {0, 0, 0, 0};
        float ww[4] = {0.f, 0.f, 0.f, 0.f};
        if (nnz == 4) {                                               // this vertex' influences, once for all frames of the CTA
            const int4 j4 = __ldg(reinterpret_cast<const int4*>(vs.ell_j) + v);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(vs.ell_w) + v);
            jj[0] = j4.x; jj[1] = j4.y; jj[2] = j4.z; jj[3] = j4.w;
            ww[0] = w4.x; ww[1] = w4.y; ww[2] = w4.z; ww[3] = w4.w;
        }
        // software pipeline over the frames: the next frame's 96 floats are requested before this frame is computed
        float nx[3];
        {
            const float* p = in + (size_t)b0 * ld_v + 3 * v0;
#pragma unroll
            for (int s = 0; s < 3; ++s) { const int i = lane + 32 * s; nx[s] = i < nf ? p[i] : 0.f; }
        }
        for (int fr = 0; fr < nfr; ++fr) {
            const int b = b0 + fr;
#pragma unroll
            for (int s = 0; s < 3; ++s) st[lane + 32 * s] = nx[s];
            if (fr + 1 < nfr) {
                const float* p = in + (size_t)(b + 1) * ld_v + 3 * v0;
#pragma unroll
                for (int s = 0; s < 3; ++s) { const int i = lane + 32 * s; nx[s] = i < nf ? p[i] : 0.f; }
            }
            __syncwarp();
            const float px = st[3 * lane], py = st[3 * lane + 1], pz = st[3 * lane + 2];
            float T[3 * NC];
#pragma unroll
            for (int e = 0; e < 3 * NC; ++e) T[e] = 0.f;
            const float* Ab = As[fr];
            if (nnz == 4) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float w = ww[k];
                    const float4* Aj = reinterpret_cast<const float4*>(Ab + jj[k] * 12);
                    const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
                    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
                    T[NC] = fmaf(w, r1.x, T[NC]); T[NC + 1] = fmaf(w, r1.y, T[NC + 1]); T[NC + 2] = fmaf(w, r1.z, T[NC + 2]);
                    T[2 * NC] = fmaf(w, r2.x, T[2 * NC]); T[2 * NC + 1] = fmaf(w, r2.y, T[2 * NC + 1]); T[2 * NC + 2] = fmaf(w, r2.z, T[2 * NC + 2]);
                    if (NC == 4) { T[3] = fmaf(w, r0.w, T[3]); T[7] = fmaf(w, r1.w, T[7]); T[11] = fmaf(w, r2.w, T[11]); }
                }
            } else {
                const int32_t* ej = vs.ell_j + (size_t)v * nnz;
                const float* ew = vs.ell_w + (size_t)v * nnz;
                for (int k = 0; k < nnz; ++k) {
                    const float w = __ldg(ew + k);
                    const float4* Aj = reinterpret_cast<const float4*>(Ab + __ldg(ej + k) * 12);
                    const float4 r0 = Aj[0], r1 = Aj[1], r2 = Aj[2];
                    T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]);
                    T[NC] = fmaf(w, r1.x, T[NC]); T[NC + 1] = fmaf(w, r1.y, T[NC + 1]); T[NC + 2] = fmaf(w, r1.z, T[NC + 2]);
                    T[2 * NC] = fmaf(w, r2.x, T[2 * NC]); T[2 * NC + 1] = fmaf(w, r2.y, T[2 * NC + 1]); T[2 * NC + 2] = fmaf(w, r2.z, T[2 * NC + 2]);
                    if (NC == 4) { T[3] = fmaf(w, r0.w, T[3]); T[7] = fmaf(w, r1.w, T[7]); T[11] = fmaf(w, r2.w, T[11]); }
                }
            }
            float ox, oy, oz;
            if (MODE == 0) {
                ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
                oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
                oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
                if (world) {
                    const float* th = f.theta + (size_t)b * m.NP;
                    const float sc = __ldg(th + 3);
                    ox = (ox + __ldg(th)) * sc * cs; oy = (oy + __ldg(th + 1)) * sc * cs; oz = (oz + __ldg(th + 2)) * sc * cs;
                }
            } else {
                ox = T[0] * px + T[3] * py + T[6] * pz;
                oy = T[1] * px + T[4] * py + T[7] * pz;
                oz = T[2] * px + T[5] * py + T[8] * pz;
            }
            __syncwarp();
            st[3 * lane] = ox; st[3 * lane + 1] = oy; st[3 * lane + 2] = oz;
            __syncwarp();
            if (out) {
                float* q = out + (size_t)b * ld_v + 3 * v0;
#pragma unroll
                for (int s = 0; s < 3; ++s) { const int i = lane + 32 * s; if (i < nf) q[i] = st[i]; }
            }
            if (MODE == 1 && out_hi) {
                float* qh = out_hi + (size_t)b * ld_out + 3 * v0;
                float* ql = out_lo + (size_t)b * ld_out + 3 * v0;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const int i = lane + 32 * s;
                    if (i < nf) { float h_, l_; split_tf32(st[i], h_, l_); qh[i] = h_; ql[i] = l_; }
                }
            }
            __syncwarp();
        }
    }
    if (MODE == 0 && fuse_joints) {
        __syncthreads();                                   // this CTA's vertex stores are visible to its own threads
        for (int fr = 0; fr < nfr; ++fr) {
            const int b = b0 + fr;
            const int yaw = f.yaw ? f.yaw[b] : 0;
            const float* Jtr_b = f.Jtr + (size_t)b * J * 3;
            const float* verts_b = out + (size_t)b * ld_v;
            const float* th = f.theta + (size_t)b * m.NP;
            for (int k = t; k < vs.K_out; k += SF_THREADS) {
                float x[3];
                joint_pos(vs, k, yaw, Jtr_b, verts_b, x);
                if (world && __ldg(vs.kj_kind + k) == 0) { // vertices are in world space already; chain joints are not
                    const float sc = __ldg(th + 3);
                    x[0] = (x[0] + __ldg(th)) * sc * cs; x[1] = (x[1] + __ldg(th + 1)) * sc * cs; x[2] = (x[2] + __ldg(th + 2)) * sc * cs;
                }
                float* o = f.joints + ((size_t)b * vs.K_out + k) * 3;
                o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
            }
        }
    }
}

// BODYFIT_SKIN=rows selects k_skin_rows (A/B timing); returns 1 when the per-frame kernel is not selected
static int bf_launch_skin_frame(int mode, const BfModel* m, const BfVSet* vs, const BfFrames* f, const float* in, float* out,
                                float* out_hi, float* out_lo, int ld_out, int world, int fuse_joints, cudaStream_t s) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("BODYFIT_SKIN"); env = (e && strcmp(e, "rows") == 0) ? 0 : 1; }
    if (!env) return 1;
    // slabs only while the frames alone do not fill the SMs a few times over
    const int num_sms = bf_num_sms();
    const int fgroups = (f->B + SF_FPB - 1) / SF_FPB;
    int slabs = (6 * num_sms + fgroups - 1) / fgroups;                 // ~6 CTAs of 8 warps per SM keep enough loads in flight
    const int groups = (vs->n + 31) / 32;
    if (slabs > (groups + 7) / 8) slabs = (groups + 7) / 8;
    if (slabs < 1) slabs = 1;
    if (fuse_joints && slabs > 1) fuse_joints = 0;
    const dim3 grid(fgroups, slabs);
    if (mode == 0) k_skin_frame<0><<<grid, SF_THREADS, 0, s>>>(*m, *vs, *f, in, out, out_hi, out_lo, ld_out, world, fuse_joints);
    else k_skin_frame<1><<<grid, SF_THREADS, 0, s>>>(*m, *vs, *f, in, out, out_hi, out_lo, ld_out, world, fuse_joints);
    BF_LAUNCH_CHECK();
    return fuse_joints ? 2 : BF_OK;                        // 2: joints were produced as well
}
