// CTA-pair (cta_group::2) variant of the forward blend GEMM: v_posed tile [256 frames x 192 coords] = pf @ Bm per
// CLUSTER of two CTAs on the two SMs of a TPC.
//
// Why: the single-CTA kernel (bf_blend_tc.cuh) feeds every 128 x 192 x 8 TF32 MMA with 4 KB of A and 6 KB of B from
// shared memory -- 107 B/clk of the SM's 128 B/clk -- and streams (128 + 192) operand rows per tile from L2; it runs at
// 55-66 % tensor-pipe active.  With M = 256 across the pair each SM holds its own 128 frame rows of A but only HALF of the B
// rows (the hardware reads the other half from the peer's shared memory): 7 KB per MMA step and SM (73 B/clk), and
// (256 + 192) operand rows from L2 per 256-frame tile instead of 2 x (128 + 192).
//
// Protocol (rank = %cluster_ctarank; rank 0 = leader):
//   * both CTAs run a TMA producer (warp 0) for their own operand halves; every TMA load signals the LEADER's full[s]
//     barrier (.cta_group::2 TMA, barrier address mapped to rank 0 with mapa); the leader's producer arms it with the bytes
//     of both CTAs;
//   * only the leader's warp 1 issues tcgen05.mma.cta_group::2; tcgen05.commit ... .multicast::cluster frees the smem
//     stage (empty[s]) and publishes the accumulator (tmem_full[buf]) in BOTH CTAs;
//   * the epilogue warps of both CTAs drain their own TMEM (lanes = their own 128 frames) and arrive on the LEADER's
//     tmem_empty[buf] (count = 2 x epilogue warps);
//   * cluster barriers after the mbarrier initialisation and before the TMEM deallocation.
#pragma once
#include "bf_blend_tc.cuh"

namespace tc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
}  // namespace tc

#ifndef TC2_STAGES
#define TC2_STAGES 6       // 28 KB per stage and CTA (A hi/lo 2 x 8 KB, half of B hi/lo 2 x 6 KB)
#endif

__global__ void __launch_bounds__(64 + 32 * TC_EPI_WARPS, 1)
k_blend_fwd_tc2(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
                const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo,
                int n_verts, int Kp, float* __restrict__ vposed, int B, int ld_v, int n_tiles_m, int n_tiles_n, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = TC_BM * TC_ROWB, b_bytes = (TC_BN1 / 2) * TC_ROWB;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    float* St = reinterpret_cast<float*>(base + TC2_STAGES * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(St + TC_EPI_WARPS * TC_ST_FLOATS);
    uint64_t* full = bars;                               // [TC2_STAGES]  (used in the leader only)
    uint64_t* empty = bars + TC2_STAGES;                 // [TC2_STAGES]  (one per CTA, multicast commit)
    uint64_t* tmem_full = bars + 2 * TC2_STAGES;         // [2]           (one per CTA, multicast commit)
    uint64_t* tmem_empty = bars + 2 * TC2_STAGES + 2;    // [2]           (leader only, 2 x TC_EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int num_k = Kp / TC_BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 2 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();                              // the peer's barriers are initialised, both allocations are done
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int t = cluster_id; t < n_tiles; t += n_clusters) {
                int fm, vn;
                tc_tile_coords(t, n_tiles_m, n_tiles_n, fm, vn);
                const int b0 = fm * (2 * TC_BM) + (int)rank * TC_BM;          // this CTA's 128 frames of the 256-frame tile
                const int n0 = vn * TC_BN1 + (int)rank * (TC_BN1 / 2);        // this CTA's half of the tile's Bm rows
                for (int kc = 0; kc < num_k; ++kc, ++it) {
                    const int s = it % TC2_STAGES;
                    const uint32_t ph = (it / TC2_STAGES) & 1;
                    tc::mbar_wait(&empty[s], ph ^ 1);
                    if (rank == 0) tc::mbar_expect_tx(&full[s], 2 * stage_bytes);
                    const uint32_t fb = tc::map_to_rank(&full[s], 0);
                    uint8_t* st = base + (size_t)s * stage_bytes;
                    tc::tma_load_2d_pair(st, &mA_hi, fb, kc * TC_BK, b0);
                    tc::tma_load_2d_pair(st + a_bytes, &mA_lo, fb, kc * TC_BK, b0);
                    tc::tma_load_2d_pair(st + 2 * a_bytes, &mB_hi, fb, kc * TC_BK, n0);
                    tc::tma_load_2d_pair(st + 2 * a_bytes + b_bytes, &mB_lo, fb, kc * TC_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = tc::make_idesc_tf32(2 * TC_BM, TC_BN1);
            int it = 0, i = 0;
            for (int t = cluster_id; t < n_tiles; t += n_clusters, ++i) {
                const int buf = i & 1;
                tc::mbar_wait(&tmem_empty[buf], ((i >> 1) & 1) ^ 1);          // both CTAs have drained this accumulator
                tc::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TC_BN1);
                for (int kc = 0; kc < num_k; ++kc, ++it) {
                    const int s = it % TC2_STAGES;
                    const uint32_t ph = (it / TC2_STAGES) & 1;
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    uint8_t* st = base + (size_t)s * stage_bytes;
                    const uint64_t a_hi = tc::make_desc(st), a_lo = tc::make_desc(st + a_bytes);
                    const uint64_t b_hi = tc::make_desc(st + 2 * a_bytes), b_lo = tc::make_desc(st + 2 * a_bytes + b_bytes);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t o = (uint64_t)(2 * k);
                        tc::umma_tf32_pair(tmem_d, a_hi + o, b_hi + o, idesc, (kc | k) ? 1u : 0u);
                        tc::umma_tf32_pair(tmem_d, a_lo + o, b_hi + o, idesc, 1u);
                        tc::umma_tf32_pair(tmem_d, a_hi + o, b_lo + o, idesc, 1u);
                    }
                    tc::umma_commit_pair(&empty[s]);
                }
                tc::umma_commit_pair(&tmem_full[buf]);
            }
        }
    } else {
        const int q = warp & 3;
        const int h = (warp - 2) >> 2;
        float* st = St + (warp - 2) * TC_ST_FLOATS;
        const uint32_t te[2] = {tc::map_to_rank(&tmem_empty[0], 0), tc::map_to_rank(&tmem_empty[1], 0)};
        int i = 0;
        for (int t = cluster_id; t < n_tiles; t += n_clusters, ++i) {
            const int buf = i & 1;
            int fm, vn;
            tc_tile_coords(t, n_tiles_m, n_tiles_n, fm, vn);
            const int b0 = fm * (2 * TC_BM) + (int)rank * TC_BM + 32 * q, n0 = vn * TC_BN1;
            const int nrows = min(32, B - b0);
            tc::mbar_wait(&tmem_full[buf], (i >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t tmem_d = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TC_BN1 + h * 96);
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                uint32_t r[48];
                tc::tmem_ld48(tmem_d + (uint32_t)(pass * 48), r);
                if (pass == 1) {
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive_cluster(te[buf]);
                }
                const int vbase = n0 / 3 + h * 32 + pass * 16;
                if (vbase >= n_verts || nrows <= 0) continue;
#pragma unroll
                for (int c = 0; c < 48; ++c) st[c * TC_ST_LD + lane] = __uint_as_float(r[c]);
                __syncwarp();
                tc::store_rows48(st, vposed + (size_t)b0 * ld_v + 3 * vbase, (size_t)ld_v, 3 * min(16, n_verts - vbase), nrows, lane);
                __syncwarp();
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();                              // no CTA frees tensor memory the pair may still be using
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// returns 1 when the pair kernel is not applicable (caller uses the single-CTA kernel)
static int bf_gemm_forward_tc2(const float* a_hi_p, const float* a_lo_p, const float* bt_hi_p, const float* bt_lo_p, int B, int Kp,
                               int ldn, int n_verts, float* dst, int ld_dst, cudaStream_t s) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("BODYFIT_TC2"); env = e ? atoi(e) : 0; }
    if (!env || B < 2 * TC_BM) return 1;
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    int rc;
    if (Kp % TC_BK != 0 || ldn % 3 != 0) return 1;
    if ((rc = bf_make_map(&a_hi, a_hi_p, B, Kp, Kp, TC_BM))) return rc;
    if ((rc = bf_make_map(&a_lo, a_lo_p, B, Kp, Kp, TC_BM))) return rc;
    if ((rc = bf_make_map(&b_hi, bt_hi_p, ldn, Kp, Kp, TC_BN1 / 2))) return rc;
    if ((rc = bf_make_map(&b_lo, bt_lo_p, ldn, Kp, Kp, TC_BN1 / 2))) return rc;
    const size_t smem = 1024 + (size_t)TC2_STAGES * (2 * TC_BM * TC_ROWB + 2 * (TC_BN1 / 2) * TC_ROWB) + TC_EPI_WARPS * TC_ST_FLOATS * 4 + 256;
    static size_t attr[BF_MAXDEV] = {0};
    if ((rc = bf_ensure_smem(k_blend_fwd_tc2, smem, attr, "k_blend_fwd_tc2"))) return rc;
    const int num_sms = bf_num_sms();
    const int tn = (ldn + TC_BN1 - 1) / TC_BN1, tm = (B + 2 * TC_BM - 1) / (2 * TC_BM);
    const int tiles = tn * tm;
    int clusters = num_sms / 2;
    if (clusters > tiles) clusters = tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(64 + 32 * TC_EPI_WARPS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_blend_fwd_tc2, a_hi, a_lo, b_hi, b_lo, n_verts, Kp, dst, B, ld_dst, tm, tn, tiles);
    if (e != cudaSuccess) { bf_set_error("k_blend_fwd_tc2 launch failed: %s", cudaGetErrorString(e)); return BF_ECUDA; }
    return BF_OK;
}
