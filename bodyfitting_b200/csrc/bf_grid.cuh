// Uniform-grid triangle index + closest point on a triangle mesh, and the SMPL+D displacement
// objective (vertex normals, point-to-scan, normal and Laplacian terms) with hand-written backward.
//
// Replaces thirdparty/mesh_grid/mesh_grid_kernel.cu:110-157 (insert_grid_surface_kernel, two passes +
// host cumsum), :239-353 (search_nearest_point_kenerel) and :12-109 (search_nearest_proj), the wrapper
// utils/mesh_grid_searcher.py:56-84, smplify/loss.py:233-242,260-288 and utils/io_utils.py:410-428.
// Differences by design:
//   * the cell lists are made DETERMINISTIC (sorted per cell) -- the reference claims slots with
//     atomicCAS in arrival order (:150-155);
//   * the point/triangle solve is EXACT (Voronoi-region classification); the reference's 4x4 KKT solve
//     with a single-edge fallback is approximate near some edges/vertices (:74-101), so parity is
//     checked on distances against an fp64 brute force, with the reference kernel as a second referee;
//   * one WARP per query: lanes take the triangles of the shell's cells, then a fixed-order butterfly
//     picks the minimum (ties -> smallest face id); the reference runs one thread per query;
//   * errors are returned, not printed (:210-212).
#pragma once
#include "bf_common.cuh"
#include "../../include/bodyfit_b200_grid.h"

__device__ __forceinline__ void grid_cell_range(const BfGrid& g, const float* a, const float* b, const float* c,
                                                int* lo, int* hi) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float mn = fminf(a[d], fminf(b[d], c[d])), mx = fmaxf(a[d], fmaxf(b[d], c[d]));
        float x = (mn - g.min[d]) / g.step;
        lo[d] = x < 0.f ? 0 : (x >= (float)g.dim[d] ? g.dim[d] - 1 : (int)floorf(x));
        x = (mx - g.min[d]) / g.step;
        hi[d] = (x < 0.f ? 0 : (x >= (float)g.dim[d] ? g.dim[d] - 1 : (int)floorf(x))) + 1;
    }
}

// pass 1 (cell_tris == NULL): count triangles per overlapped cell; pass 2: claim slots
__global__ void k_grid_insert(BfGrid g, int32_t* __restrict__ counts, const int32_t* __restrict__ cell_start,
                              int32_t* __restrict__ cell_tris) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= g.Fs) return;
    const int32_t* tri = g.faces + 3 * f;
    const float* a = g.verts + 3 * tri[0];
    const float* b = g.verts + 3 * tri[1];
    const float* c = g.verts + 3 * tri[2];
    int lo[3], hi[3];
    grid_cell_range(g, a, b, c, lo, hi);
    for (int x = lo[0]; x < hi[0]; ++x)
        for (int y = lo[1]; y < hi[1]; ++y)
            for (int z = lo[2]; z < hi[2]; ++z) {
                const int cell = (x * g.dim[1] + y) * g.dim[2] + z;
                const int slot = atomicAdd(counts + cell, 1);
                if (cell_tris) cell_tris[cell_start[cell] + slot] = f;
            }
}

// exclusive scan of counts[n] -> start[n+1], one block (n ~ number of scan vertices)
__global__ void __launch_bounds__(1024) k_grid_scan(const int32_t* __restrict__ counts, int32_t* __restrict__ start, int n) {
    __shared__ int32_t wsum[32];
    __shared__ int32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + t;
        int v = i < n ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            wsum[lane] = wi - w;
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < n) start[i] = carry + wsum[warp] + incl - v;
        __syncthreads();
        if (t == 1023) carry_s = carry + wsum[warp] + incl;
        __syncthreads();
    }
    if (t == 0) start[n] = carry_s;
}

// ascending face id inside every cell -> deterministic lists
__global__ void k_grid_sort(const int32_t* __restrict__ cell_start, int32_t* __restrict__ cell_tris, int ncell) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int s = cell_start[c], e = cell_start[c + 1];
    for (int i = s + 1; i < e; ++i) {
        const int v = cell_tris[i];
        int j = i - 1;
        while (j >= s && cell_tris[j] > v) { cell_tris[j + 1] = cell_tris[j]; --j; }
        cell_tris[j + 1] = v;
    }
}

// exact closest point on triangle (a,b,c) to p (Voronoi regions): q = a + u (b - a) + v (c - a), i.e. barycentric
// coefficients (1 - u - v, u, v).  T = float, or a forward-mode dual number (Dual9 below) carrying d/d(a, b, c).
template <typename T>
__device__ __forceinline__ void closest_uv(const float* p, const T* a, const T* b, const T* c, T& u, T& v) {
    const T ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    const T ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const T ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    const T d1 = ab[0] * ap[0] + ab[1] * ap[1] + ab[2] * ap[2];
    const T d2 = ac[0] * ap[0] + ac[1] * ap[1] + ac[2] * ap[2];
    const T zero = T(0.f), one = T(1.f);
    u = zero; v = zero;
    if (d1 <= 0.f && d2 <= 0.f) return;
    const T bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
    const T d3 = ab[0] * bp[0] + ab[1] * bp[1] + ab[2] * bp[2];
    const T d4 = ac[0] * bp[0] + ac[1] * bp[1] + ac[2] * bp[2];
    if (d3 >= 0.f && d4 <= d3) { u = one; return; }
    const T vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { u = d1 / (d1 - d3); return; }
    const T cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    const T d5 = ab[0] * cp[0] + ab[1] * cp[1] + ab[2] * cp[2];
    const T d6 = ac[0] * cp[0] + ac[1] * cp[1] + ac[2] * cp[2];
    if (d6 >= 0.f && d5 <= d6) { v = one; return; }
    const T vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { v = d2 / (d2 - d6); return; }
    const T va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const T w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        u = one - w; v = w;
        return;
    }
    const T den = one / (va + vb + vc);
    u = vb * den; v = vc * den;
}

// returns squared distance, writes the point
__device__ __forceinline__ float closest_on_triangle(const float* p, const float* a, const float* b, const float* c, float* q) {
    float u, v;
    closest_uv<float>(p, a, b, c, u, v);
    q[0] = a[0] + u * (b[0] - a[0]) + v * (c[0] - a[0]);
    q[1] = a[1] + u * (b[1] - a[1]) + v * (c[1] - a[1]);
    q[2] = a[2] + u * (b[2] - a[2]) + v * (c[2] - a[2]);
    const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return dx * dx + dy * dy + dz * dz;
}

// forward-mode dual number: value + the 9 partial derivatives w.r.t. the triangle's vertex coordinates (a, b, c)
struct Dual9 {
    float x, d[9];
    __device__ Dual9() {}
    __device__ explicit Dual9(float v) : x(v) {
#pragma unroll
        for (int i = 0; i < 9; ++i) d[i] = 0.f;
    }
};
__device__ __forceinline__ Dual9 operator+(const Dual9& a, const Dual9& b) { Dual9 r; r.x = a.x + b.x; for (int i = 0; i < 9; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ Dual9 operator-(const Dual9& a, const Dual9& b) { Dual9 r; r.x = a.x - b.x; for (int i = 0; i < 9; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ Dual9 operator-(float a, const Dual9& b) { Dual9 r; r.x = a - b.x; for (int i = 0; i < 9; ++i) r.d[i] = -b.d[i]; return r; }
__device__ __forceinline__ Dual9 operator*(const Dual9& a, const Dual9& b) { Dual9 r; r.x = a.x * b.x; for (int i = 0; i < 9; ++i) r.d[i] = a.d[i] * b.x + a.x * b.d[i]; return r; }
__device__ __forceinline__ Dual9 operator/(const Dual9& a, const Dual9& b) {
    Dual9 r; const float ib = 1.0f / b.x; r.x = a.x * ib;
    for (int i = 0; i < 9; ++i) r.d[i] = (a.d[i] - r.x * b.d[i]) * ib;
    return r;
}
__device__ __forceinline__ bool operator<=(const Dual9& a, float b) { return a.x <= b; }
__device__ __forceinline__ bool operator>=(const Dual9& a, float b) { return a.x >= b; }
__device__ __forceinline__ bool operator<=(const Dual9& a, const Dual9& b) { return a.x <= b.x; }

// barycentric coefficients of the closest point on the triangle the search selected (the reference's `coeff` output,
// mesh_grid_kernel.cu:405-410,:12-109): near_pt = c0 v0 + c1 v1 + c2 v2
__global__ void k_grid_bary(BfGrid g, const float* __restrict__ points, const int32_t* __restrict__ near_faces, int Q,
                            float* __restrict__ coeff) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const int f = near_faces[qi];
    float u = 0.f, v = 0.f;
    if (f >= 0) {
        const int32_t* tri = g.faces + 3 * f;
        const float p[3] = {points[3 * qi], points[3 * qi + 1], points[3 * qi + 2]};
        closest_uv<float>(p, g.verts + 3 * __ldg(tri), g.verts + 3 * __ldg(tri + 1), g.verts + 3 * __ldg(tri + 2), u, v);
    }
    coeff[3 * qi] = f >= 0 ? 1.0f - u - v : 0.f; coeff[3 * qi + 1] = u; coeff[3 * qi + 2] = v;
}

// grad[q][i][j][k] = d near_pt[q][j] / d verts[tri[i]][k] for the face the search selected (piecewise exact: the derivative of
// the closest-point map inside the Voronoi region the query falls in).  The reference's kernel of this name is unfinished
// (mesh_grid_kernel.cu:354-382: "TODO: calculate inverse matrix", every thread writes the first 27 entries) and never called.
__global__ void k_grid_nearest_bwd(BfGrid g, const float* __restrict__ points, const int32_t* __restrict__ near_faces, int Q,
                                   float* __restrict__ grad) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    float* o = grad + (size_t)qi * 27;
    const int f = near_faces[qi];
    if (f < 0) { for (int e = 0; e < 27; ++e) o[e] = 0.f; return; }
    const int32_t* tri = g.faces + 3 * f;
    const float p[3] = {points[3 * qi], points[3 * qi + 1], points[3 * qi + 2]};
    Dual9 vtx[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vtx[i * 3 + k] = Dual9(g.verts[3 * __ldg(tri + i) + k]);
            vtx[i * 3 + k].d[i * 3 + k] = 1.0f;
        }
    Dual9 u, v;
    closest_uv<Dual9>(p, vtx, vtx + 3, vtx + 6, u, v);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const Dual9 q = vtx[j] + u * (vtx[3 + j] - vtx[j]) + v * (vtx[6 + j] - vtx[j]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) o[(i * 3 + j) * 3 + k] = q.d[i * 3 + k];
    }
}

// one warp per query point: shells of cells of growing L-inf radius around the query's cell
__global__ void __launch_bounds__(256) k_grid_nearest(BfGrid g, const float* __restrict__ points, int Q,
                                                      float* __restrict__ near_pts, int32_t* __restrict__ near_faces,
                                                      float* __restrict__ dist2_out) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= Q) return;
    const float p[3] = {points[3 * qi], points[3 * qi + 1], points[3 * qi + 2]};
    int c0[3];
    int maxL = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float x = (p[d] - g.min[d]) / g.step;
        x = x < 0.f ? 0.f : (x >= (float)g.dim[d] ? (float)(g.dim[d] - 1) : floorf(x));
        c0[d] = (int)x;
        maxL = max(maxL, max(c0[d], g.dim[d] - 1 - c0[d]));
    }
    float best = 3.0e38f;
    int bestf = 0x7fffffff;
    float bq[3] = {0.f, 0.f, 0.f};
    for (int L = 0; L <= maxL; ++L) {
        const int side = 2 * L + 1;
        const int ncand = side * side * side;
        for (int idx = lane; idx < ncand; idx += 32) {
            const int ox = idx / (side * side) - L, oy = (idx / side) % side - L, oz = idx % side - L;
            if (max(abs(ox), max(abs(oy), abs(oz))) != L) continue;          // shell only
            const int cx = c0[0] + ox, cy = c0[1] + oy, cz = c0[2] + oz;
            if (cx < 0 || cy < 0 || cz < 0 || cx >= g.dim[0] || cy >= g.dim[1] || cz >= g.dim[2]) continue;
            // lower bound of the distance to this cell's box
            float lb = 0.f;
            {
                const int cc[3] = {cx, cy, cz};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float lo = g.min[d] + g.step * cc[d], hi = lo + g.step;
                    const float e = p[d] < lo ? lo - p[d] : (p[d] > hi ? p[d] - hi : 0.f);
                    lb += e * e;
                }
            }
            if (lb > best) continue;
            const int cell = (cx * g.dim[1] + cy) * g.dim[2] + cz;
            const int s = __ldg(g.cell_start + cell), e = __ldg(g.cell_start + cell + 1);
            for (int i = s; i < e; ++i) {
                const int f = __ldg(g.cell_tris + i);
                const int32_t* tri = g.faces + 3 * f;
                float q[3];
                const float d2 = closest_on_triangle(p, g.verts + 3 * __ldg(tri), g.verts + 3 * __ldg(tri + 1),
                                                     g.verts + 3 * __ldg(tri + 2), q);
                if (d2 < best || (d2 == best && f < bestf)) { best = d2; bestf = f; bq[0] = q[0]; bq[1] = q[1]; bq[2] = q[2]; }
            }
        }
        // warp minimum (ties: smallest face id), fixed butterfly
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int of = __shfl_xor_sync(0xffffffffu, bestf, o);
            const float q0 = __shfl_xor_sync(0xffffffffu, bq[0], o), q1 = __shfl_xor_sync(0xffffffffu, bq[1], o),
                        q2 = __shfl_xor_sync(0xffffffffu, bq[2], o);
            if (ob < best || (ob == best && of < bestf)) { best = ob; bestf = of; bq[0] = q0; bq[1] = q1; bq[2] = q2; }
        }
        // everything outside the searched cube is at least L*step away (mesh_grid_kernel.cu:349)
        const float r = (float)L * g.step;
        if (bestf != 0x7fffffff && best < r * r) break;
    }
    if (lane == 0) {
        near_faces[qi] = bestf == 0x7fffffff ? -1 : bestf;
        near_pts[3 * qi] = bq[0]; near_pts[3 * qi + 1] = bq[1]; near_pts[3 * qi + 2] = bq[2];
        if (dist2_out) dist2_out[qi] = best;
    }
}

// ---- SMPL+D displacement objective (smplify/smplify.py:228-247) -------------------------------------------
// face normals of the deformed body mesh: n = (vb - va) x (vc - va), nhat = n / (|n| + 1e-8)   (io_utils.py:405-415)
__global__ void k_face_normals(const float* __restrict__ v, const int32_t* __restrict__ faces, int F,
                               float* __restrict__ nhat, float* __restrict__ nlen) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float* a = v + 3 * faces[3 * f];
    const float* b = v + 3 * faces[3 * f + 1];
    const float* c = v + 3 * faces[3 * f + 2];
    const float u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const float inv = 1.0f / (len + 1e-8f);
    nhat[3 * f] = n[0] * inv; nhat[3 * f + 1] = n[1] * inv; nhat[3 * f + 2] = n[2] * inv;
    nlen[f] = len;
}

// vertex normals: N_v = sum of nhat over incident faces (CSR, fixed order), m_v = N_v / (|N_v| + 1e-8)
__global__ void k_vertex_normals(const float* __restrict__ nhat, const int32_t* __restrict__ vf_ptr,
                                 const int32_t* __restrict__ vf_face, int V, float* __restrict__ m, float* __restrict__ Nlen) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int e = vf_ptr[v]; e < vf_ptr[v + 1]; ++e) {
        const int f = vf_face[e];
        s0 += nhat[3 * f]; s1 += nhat[3 * f + 1]; s2 += nhat[3 * f + 2];
    }
    const float len = sqrtf(s0 * s0 + s1 * s1 + s2 * s2);
    const float inv = 1.0f / (len + 1e-8f);
    m[3 * v] = s0 * inv; m[3 * v + 1] = s1 * inv; m[3 * v + 2] = s2 * inv;
    Nlen[v] = len;
}

// per-block partial sums (fixed order) of: |P - C|^2 ; 1 - <scan_face_normal[closest], m_v> ; smoothness per face
__global__ void __launch_bounds__(256) k_smpld_partials(const float* __restrict__ P, const float* __restrict__ C,
                                                        const int32_t* __restrict__ near_faces,
                                                        const float* __restrict__ scan_fn, const float* __restrict__ m,
                                                        const int32_t* __restrict__ faces, int V, int F,
                                                        float* __restrict__ partial /* [gridDim.x,3] */) {
    __shared__ float red[3][8];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(V, F); i += gridDim.x * blockDim.x) {
        if (i < V) {
            const float dx = P[3 * i] - C[3 * i], dy = P[3 * i + 1] - C[3 * i + 1], dz = P[3 * i + 2] - C[3 * i + 2];
            a0 += dx * dx + dy * dy + dz * dz;
            const float* fn = scan_fn + 3 * near_faces[i];
            a1 += 1.0f - (fn[0] * m[3 * i] + fn[1] * m[3 * i + 1] + fn[2] * m[3 * i + 2]);
        }
        if (i < F) {
            const float* na = m + 3 * faces[3 * i];
            const float* nb = m + 3 * faces[3 * i + 1];
            const float* nc = m + 3 * faces[3 * i + 2];
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float ab = na[d] - nb[d], ca = nc[d] - na[d], bc = nb[d] - nc[d];
                s += ab * ab + ca * ca + bc * bc;
            }
            a2 += s;
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        partial[blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// totals[0] = icp = sqrt(sum |P-C|^2), totals[1] = mean normal term, totals[2] = mean smoothness, totals[3] = loss
__global__ void k_smpld_finish(const float* __restrict__ partial, int nblocks, int V, int F, float reg_scale,
                               float* __restrict__ totals) {
    if (threadIdx.x != 0) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < nblocks; ++i) { s0 += partial[3 * i]; s1 += partial[3 * i + 1]; s2 += partial[3 * i + 2]; }
    const float icp = sqrtf(s0), nl = s1 / (float)V, sm = s2 / (float)F;
    totals[0] = icp; totals[1] = nl; totals[2] = sm;
    totals[3] = icp + (nl + sm) * reg_scale;
}

// d loss / d m_v (vertex normal), both regularisers: -scan_fn[closest]/V * reg, and the Laplacian term gathered
// over incident faces (CSR with the vertex's corner index) 2 * (2 m_self - m_other1 - m_other2) / F * reg
__global__ void k_smpld_dm(const int32_t* __restrict__ near_faces, const float* __restrict__ scan_fn,
                           const float* __restrict__ m, const int32_t* __restrict__ faces,
                           const int32_t* __restrict__ vf_ptr, const int32_t* __restrict__ vf_face, int V, int F,
                           float reg_scale, const float* __restrict__ Nlen, float* __restrict__ dN) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float* fn = scan_fn + 3 * near_faces[v];
    float g[3] = {-fn[0] / (float)V, -fn[1] / (float)V, -fn[2] / (float)V};
    const float mv[3] = {m[3 * v], m[3 * v + 1], m[3 * v + 2]};
    const float k = 2.0f / (float)F;
    for (int e = vf_ptr[v]; e < vf_ptr[v + 1]; ++e) {
        const int f = vf_face[e];
#pragma unroll
        for (int cnr = 0; cnr < 3; ++cnr) {
            const int o = faces[3 * f + cnr];
            if (o == v) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) g[d] += k * (mv[d] - m[3 * o + d]);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] *= reg_scale;
    // through m = N / (|N| + eps):  dN = g / (len + eps) - N (N . g) / (len (len + eps)^2)
    const float len = Nlen[v], le = len + 1e-8f;
    const float dot = mv[0] * g[0] + mv[1] * g[1] + mv[2] * g[2];       // m . g ; N = m * le
    const float c = (len > 0.f) ? dot * le / (len * le) : 0.f;          // (N.g)/(len (len+eps)^2) * le = (m.g)/len ... times m*le
#pragma unroll
    for (int d = 0; d < 3; ++d) dN[3 * v + d] = g[d] / le - mv[d] * c;
}

// d loss / d nhat_f = sum of dN over the face's 3 vertices, through nhat = n/(|n|+eps), then to the 3 corner
// positions; stored per face-corner [F,3,3] and gathered per vertex by k_smpld_dv (no atomics)
__global__ void k_smpld_dface(const float* __restrict__ v, const int32_t* __restrict__ faces, const float* __restrict__ nhat,
                              const float* __restrict__ nlen, const float* __restrict__ dN, int F, float* __restrict__ dcorner) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
    float g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] = dN[3 * ia + d] + dN[3 * ib + d] + dN[3 * ic + d];
    const float len = nlen[f], le = len + 1e-8f;
    const float h[3] = {nhat[3 * f], nhat[3 * f + 1], nhat[3 * f + 2]};
    const float dot = h[0] * g[0] + h[1] * g[1] + h[2] * g[2];
    const float c = (len > 0.f) ? dot / len : 0.f;
    float dn[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) dn[d] = g[d] / le - h[d] * c;
    const float* a = v + 3 * ia; const float* b = v + 3 * ib; const float* cc = v + 3 * ic;
    const float u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, w[3] = {cc[0] - a[0], cc[1] - a[1], cc[2] - a[2]};
    // n = u x w :  dL/du = w x dn ,  dL/dw = dn x u
    const float du[3] = {w[1] * dn[2] - w[2] * dn[1], w[2] * dn[0] - w[0] * dn[2], w[0] * dn[1] - w[1] * dn[0]};
    const float dw[3] = {dn[1] * u[2] - dn[2] * u[1], dn[2] * u[0] - dn[0] * u[2], dn[0] * u[1] - dn[1] * u[0]};
    float* o = dcorner + 9 * f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { o[d] = -du[d] - dw[d]; o[3 + d] = du[d]; o[6 + d] = dw[d]; }
}

// gradient w.r.t. the displaced vertex = icp term (P - C)/icp + gathered corner gradients; then Adam on disp
__global__ void k_smpld_step(const float* __restrict__ P, const float* __restrict__ C, const float* __restrict__ totals,
                             const float* __restrict__ dcorner, const int32_t* __restrict__ faces,
                             const int32_t* __restrict__ vf_ptr, const int32_t* __restrict__ vf_face, int V,
                             float* __restrict__ disp, float* __restrict__ am, float* __restrict__ av, float* __restrict__ grad_out,
                             float step, float bc2_sqrt, float beta2, float om_b1, float om_b2, float eps) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float icp = totals[0];
    float g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) g[d] = (P[3 * v + d] - C[3 * v + d]) / icp;
    for (int e = vf_ptr[v]; e < vf_ptr[v + 1]; ++e) {
        const int f = vf_face[e];
#pragma unroll
        for (int cnr = 0; cnr < 3; ++cnr)
            if (faces[3 * f + cnr] == v) {
#pragma unroll
                for (int d = 0; d < 3; ++d) g[d] += dcorner[9 * f + 3 * cnr + d];
            }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int i = 3 * v + d;
        if (grad_out) grad_out[i] = g[d];
        float mi = am[i], vi = av[i];
        mi = mi + (g[d] - mi) * om_b1;
        vi = vi * beta2 + om_b2 * g[d] * g[d];
        disp[i] = disp[i] + (-step) * mi / (sqrtf(vi) / bc2_sqrt + eps);
        am[i] = mi; av[i] = vi;
    }
}

__global__ void k_add3(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i] + b[i];
}

// ---- inside / ray queries on the same grid (utils/mesh_grid_searcher.py:86-99) ----------------------------
// 2-D orientation of p against the directed edge a -> b in the plane (u, w); the edge is evaluated in a canonical
// direction (lower vertex id first) so that the two triangles sharing it see exactly the same number -> no ray
// can slip between them or be counted twice (watertight crossing parity).
__device__ __forceinline__ float edge_fn(const float* va, const float* vb, int ia, int ib, float pu, float pw, int u, int w) {
    const bool sw = ia > ib;
    const float* a = sw ? vb : va;
    const float* b = sw ? va : vb;
    const float e = (b[u] - a[u]) * (pw - a[w]) - (b[w] - a[w]) * (pu - a[u]);
    return sw ? -e : e;
}

// Inside test by crossing parity along an axis ray (the reference marches the same kind of ray through its grid and
// counts distinct triangles, mesh_grid_kernel.cu:568-650): the axis / direction with the fewest cells to the grid
// border is taken, every triangle listed in the cells on the way is tested in the projection plane, and a crossing
// is charged to the ONE cell that contains it (so a triangle listed in several cells of the column is counted once;
// the reference keeps a 16-entry visited list instead).  +1 inside, -1 outside (also outside the grid box).
__global__ void __launch_bounds__(128) k_grid_inside(BfGrid g, const float* __restrict__ points, int Q, float* __restrict__ signs) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const float p[3] = {points[3 * qi], points[3 * qi + 1], points[3 * qi + 2]};
    int c[3];
    int best_d = 0, best_n = 0x7fffffff;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float x = (p[d] - g.min[d]) / g.step;
        if (!(x >= 0.f) || x >= (float)g.dim[d]) { signs[qi] = -1.0f; return; }
        c[d] = (int)x;
        if (c[d] < best_n) { best_n = c[d]; best_d = 2 * d; }                                  // towards the lower border
        if (g.dim[d] - 1 - c[d] < best_n) { best_n = g.dim[d] - 1 - c[d]; best_d = 2 * d + 1; } // towards the upper border
    }
    const int ax = best_d >> 1, u = (ax + 1) % 3, w = (ax + 2) % 3;
    const int dir = (best_d & 1) ? 1 : -1;
    int crossings = 0;
    int cc[3] = {c[0], c[1], c[2]};
    for (int s = 0; s <= best_n; ++s, cc[ax] += dir) {
        const int cell = (cc[0] * g.dim[1] + cc[1]) * g.dim[2] + cc[2];
        const int e0 = __ldg(g.cell_start + cell), e1 = __ldg(g.cell_start + cell + 1);
        for (int i = e0; i < e1; ++i) {
            const int f = __ldg(g.cell_tris + i);
            const int ia = __ldg(g.faces + 3 * f), ib = __ldg(g.faces + 3 * f + 1), ic = __ldg(g.faces + 3 * f + 2);
            const float* a = g.verts + 3 * ia;
            const float* b = g.verts + 3 * ib;
            const float* cv = g.verts + 3 * ic;
            const float e_ab = edge_fn(a, b, ia, ib, p[u], p[w], u, w);
            const float e_bc = edge_fn(b, cv, ib, ic, p[u], p[w], u, w);
            const float e_ca = edge_fn(cv, a, ic, ia, p[u], p[w], u, w);
            // inside the projected triangle (either winding); an exactly-zero edge value belongs to the positive side
            const bool pos = e_ab >= 0.f && e_bc >= 0.f && e_ca >= 0.f;
            const bool neg = e_ab < 0.f && e_bc < 0.f && e_ca < 0.f;
            if (!pos && !neg) continue;
            const float area = e_ab + e_bc + e_ca;
            if (area == 0.f) continue;                                         // degenerate in projection
            // crossing coordinate along the axis (barycentric interpolation: weights e_bc -> a, e_ca -> b, e_ab -> c)
            const float t = (e_bc * a[ax] + e_ca * b[ax] + e_ab * cv[ax]) / area;
            if (dir > 0 ? !(t > p[ax]) : !(t < p[ax])) continue;
            int ct = (int)floorf((t - g.min[ax]) / g.step);
            ct = min(max(ct, 0), g.dim[ax] - 1);
            if (dir > 0) ct = max(ct, c[ax]); else ct = min(ct, c[ax]);       // never before the query's own cell
            if (ct == cc[ax]) ++crossings;
        }
    }
    signs[qi] = (crossings & 1) ? 1.0f : -1.0f;
}

// Ray / triangle (Moeller-Trumbore), hit iff barycentrics >= -eps and t >= -eps (one-directional ray that includes
// its origin, as mesh_grid_kernel.cu:742-781 with both_direction = false)
__device__ __forceinline__ bool ray_hits_triangle(const float* o, const float* d, const float* a, const float* b, const float* c) {
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float pv[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const float det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (fabsf(det) <= 1e-20f) return false;                                   // parallel to the triangle's plane
    const float inv = 1.0f / det;
    const float tv[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
    const float bu = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv;
    const float qv[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
    const float bv = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv;
    const float t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv;
    const float eps = 1e-6f;
    return bu >= -eps && bv >= -eps && bu + bv <= 1.0f + eps && t >= -eps;
}

// intersects_any: 3-D DDA through the grid from the ray's entry point; the first cell holding a hit ends the walk
__global__ void __launch_bounds__(128) k_grid_ray_any(BfGrid g, const float* __restrict__ origins, const float* __restrict__ dirs,
                                                      int Q, uint8_t* __restrict__ hit) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const float o[3] = {origins[3 * qi], origins[3 * qi + 1], origins[3 * qi + 2]};
    const float d[3] = {dirs[3 * qi], dirs[3 * qi + 1], dirs[3 * qi + 2]};
    hit[qi] = 0;
    if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < 1e-9f) return;               // degenerate direction (:1052-1055)
    // clip the ray to the grid box: t in [t0, t1]
    float t0 = 0.f, t1 = 3.0e38f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = g.min[k], hi = g.min[k] + g.step * (float)g.dim[k];
        if (d[k] == 0.f) { if (o[k] < lo || o[k] > hi) return; continue; }
        const float inv = 1.0f / d[k];
        float ta = (lo - o[k]) * inv, tb = (hi - o[k]) * inv;
        if (ta > tb) { const float s = ta; ta = tb; tb = s; }
        t0 = fmaxf(t0, ta); t1 = fminf(t1, tb);
    }
    if (t0 > t1) return;
    int c[3], stp[3];
    float tmax[3], tdel[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = (o[k] + t0 * d[k] - g.min[k]) / g.step;
        c[k] = min(max((int)floorf(x), 0), g.dim[k] - 1);
        if (d[k] > 0.f) { stp[k] = 1; tdel[k] = g.step / d[k]; tmax[k] = (g.min[k] + g.step * (float)(c[k] + 1) - o[k]) / d[k]; }
        else if (d[k] < 0.f) { stp[k] = -1; tdel[k] = -g.step / d[k]; tmax[k] = (g.min[k] + g.step * (float)c[k] - o[k]) / d[k]; }
        else { stp[k] = 0; tdel[k] = 3.0e38f; tmax[k] = 3.0e38f; }
    }
    for (int guard = g.dim[0] + g.dim[1] + g.dim[2] + 3; guard > 0; --guard) {
        const int cell = (c[0] * g.dim[1] + c[1]) * g.dim[2] + c[2];
        const int e0 = __ldg(g.cell_start + cell), e1 = __ldg(g.cell_start + cell + 1);
        for (int i = e0; i < e1; ++i) {
            const int f = __ldg(g.cell_tris + i);
            if (ray_hits_triangle(o, d, g.verts + 3 * __ldg(g.faces + 3 * f), g.verts + 3 * __ldg(g.faces + 3 * f + 1),
                                  g.verts + 3 * __ldg(g.faces + 3 * f + 2))) { hit[qi] = 1; return; }
        }
        const int k = tmax[0] <= tmax[1] ? (tmax[0] <= tmax[2] ? 0 : 2) : (tmax[1] <= tmax[2] ? 1 : 2);
        c[k] += stp[k];
        if (stp[k] == 0 || c[k] < 0 || c[k] >= g.dim[k]) return;
        tmax[k] += tdel[k];
    }
}
