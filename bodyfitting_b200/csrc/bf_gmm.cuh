// GMM pose prior (smplify/prior.py:181-196, MaxMixturePrior.merged_log_likelihood) for all
// frames: min over 8 components of 0.5 d^T P_m d - log(nll_w_m), d = pose69 - mu_m, and the
// gradient of the winning component.
//
// 32 frames per CTA, lane = frame, 4 warps = 4 row ranges of the 69x69 precision matrix.  The
// symmetrised precision P_sym = (P + P^T)/2 of the current component sits in shared memory
// (double buffered, next component prefetched while computing) and every lane of a warp walks
// the same rows in the same order, so each LDS.128 is a warp-wide broadcast feeding 4 FFMAs per
// thread; the per-frame vector d lives in registers.  With P_sym, y = P_sym d is at once the
// gradient (0.5 (P + P^T) d, what autograd of the reference produces) and gives the quadratic
// form d.y = d^T P d.  The gradient of the best component so far is kept in a per-frame
// shared-memory column (two buffers, swapped when a better component is found): no second
// pass, no divergence.
#pragma once
#include "bf_common.cuh"

#define GM_F 32                 // frames per CTA (lane = frame)
#define GM_PARTS 4              // warps per CTA, each owns a range of rows
#define GM_D BF_GMM_D           // 69
#define GM_LD 72                // padded row length of P_sym (float4 rows, pad = 0)
#define GM_ROWS 18              // rows per part (4 * 18 >= 69)

struct GmmSmem {
    float P[2][GM_D * GM_LD];   // double-buffered P_sym of one component
    float g[2][GM_D][GM_F];     // gradient candidates, column per frame
    float x[GM_D][GM_F];        // pose vector, column per frame
    float mu[2][GM_LD];
    float qp[GM_PARTS][GM_F];   // partial quadratic forms
};

// pose: row b at pose + b*ld, first nvalid entries used (the rest of the 69 are 0, smplify/loss.py:206-207)
__global__ void __launch_bounds__(GM_F * GM_PARTS) k_gmm_prior(BfModel m, const float* __restrict__ pose, int ld, int nvalid,
                                                                int B, float wp, float* __restrict__ gmm_grad,
                                                                float* __restrict__ gmm_loss) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GmmSmem& S = *reinterpret_cast<GmmSmem*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, part = t >> 5;
    const int b = blockIdx.x * GM_F + lane;
    const bool valid = b < B;
    const float* th = pose + (size_t)(valid ? b : B - 1) * ld;
    for (int j = part; j < GM_D; j += GM_PARTS) S.x[j][lane] = (j < nvalid) ? th[j] : 0.f;

    auto load_comp = [&](int c, int buf) {
        const float4* src = reinterpret_cast<const float4*>(m.gmm_psym + (size_t)c * GM_D * GM_LD);
        float4* dst = reinterpret_cast<float4*>(S.P[buf]);
        for (int i = t; i < GM_D * GM_LD / 4; i += GM_F * GM_PARTS) dst[i] = __ldg(src + i);
        if (t < GM_LD) S.mu[buf][t] = (t < GM_D) ? __ldg(m.gmm_mean + c * GM_D + t) : 0.f;
    };
    load_comp(0, 0);
    __syncthreads();

    const int i0 = part * GM_ROWS;
    const int i1 = min(GM_D, i0 + GM_ROWS);
    float best = 0.f;
    int bi = 0;                                      // buffer holding the gradient of the best component
    for (int c = 0; c < m.n_gmm; ++c) {
        const int cur = c & 1;
        if (c + 1 < m.n_gmm) load_comp(c + 1, cur ^ 1);      // prefetch into the other buffer
        float d[GM_LD];
#pragma unroll
        for (int j = 0; j < GM_LD; ++j) d[j] = (j < GM_D) ? S.x[j][lane] - S.mu[cur][j] : 0.f;
        const int cand = bi ^ 1;
        float q = 0.f;
        const float* Pc = S.P[cur];
#pragma unroll 1
        for (int i = i0; i < i1; ++i) {
            const float4* row = reinterpret_cast<const float4*>(Pc + i * GM_LD);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < GM_LD / 4; ++j4) {
                const float4 p = row[j4];
                a0 = fmaf(p.x, d[4 * j4 + 0], a0);
                a1 = fmaf(p.y, d[4 * j4 + 1], a1);
                a2 = fmaf(p.z, d[4 * j4 + 2], a2);
                a3 = fmaf(p.w, d[4 * j4 + 3], a3);
            }
            const float y = (a0 + a1) + (a2 + a3);
            S.g[cand][i][lane] = y;
            q = fmaf(y, S.x[i][lane] - S.mu[cur][i], q);
        }
        S.qp[part][lane] = q;
        __syncthreads();
        const float qs = (S.qp[0][lane] + S.qp[1][lane]) + (S.qp[2][lane] + S.qp[3][lane]);
        const float ll = 0.5f * qs - __ldg(m.gmm_logw + c);
        if (c == 0 || ll < best) { best = ll; bi = cand; }      // identical decision in all four parts
        __syncthreads();                             // qp / P[cur] may be overwritten from here on
    }
    if (valid) {
        if (part == 0) gmm_loss[b] = wp * best;
        float* o = gmm_grad + (size_t)b * GM_D;
        for (int j = part; j < GM_D; j += GM_PARTS) o[j] = wp * S.g[bi][j][lane];
    }
}


// ---- tensor-core form (BF_F_TC): y_m = P_sym,m (x - mu_m) for all components is ONE GEMM -------------------------
//   [pose69 | 1 | 0 ...] (K = 80)  @  [P_sym,m | -P_sym,m mu_m | 0]  ->  Y [B, n_gmm * 72]
// run by the blend GEMM kernel (k_blend_fwd_tc, 3xTF32), between a pack kernel (operand rows, hi/lo split) and a select
// kernel (quadratic forms d.y, arg-min, gradient of the winner).  39.7k FFMA per frame move to the tensor pipe, which is
// idle while the per-frame loss kernel of the same iteration runs on the main stream.
#define GM_KG 80

__global__ void __launch_bounds__(256) k_gmm_pack(const float* __restrict__ pose, int ld, int nvalid, int B,
                                                  float* __restrict__ x_hi, float* __restrict__ x_lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * GM_KG) return;
    const int b = (int)(i / GM_KG), j = (int)(i % GM_KG);
    const float v = j < nvalid ? pose[(size_t)b * ld + j] : (j == GM_D ? 1.0f : 0.0f);
    float hi, lo;
    split_tf32(v, hi, lo);
    x_hi[i] = hi; x_lo[i] = lo;
}

// one warp per frame: q_m = (x - mu_m) . y_m, ll_m = q_m / 2 - log w_m, first minimum wins (prior.py:188-196)
__global__ void __launch_bounds__(128) k_gmm_select(BfModel m, const float* __restrict__ pose, int ld, int nvalid, int B,
                                                    const float* __restrict__ Y, int ldy, float wp,
                                                    float* __restrict__ gmm_grad, float* __restrict__ gmm_loss) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    float x[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = lane + 32 * r;
        x[r] = i < nvalid ? pose[(size_t)b * ld + i] : 0.f;
    }
    const float* y = Y + (size_t)b * ldy;
    float best = 0.f;
    int bc = 0;
    for (int c = 0; c < m.n_gmm; ++c) {
        float q = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int i = lane + 32 * r;
            if (i < GM_D) q = fmaf(x[r] - __ldg(m.gmm_mean + c * GM_D + i), y[c * GM_LD + i], q);
        }
        q = warp_sum(q);
        const float ll = 0.5f * q - __ldg(m.gmm_logw + c);
        if (c == 0 || ll < best) { best = ll; bc = c; }
    }
    if (lane == 0) gmm_loss[b] = wp * best;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int i = lane + 32 * r;
        if (i < GM_D) gmm_grad[(size_t)b * GM_D + i] = wp * y[bc * GM_LD + i];
    }
}
