// dmc.cuh -- dual marching cubes phase 2 (dual vertices, quads) and backward stage A.
//
// A DMC "quad" is a crossing grid edge (same enumeration / order as the MC vertices, so the
// edge records E are shared with MC); a DMC "vertex" is a patch of a used cell.  The patch
// records P[k] = {base, lo, hi, flip} give, per 32-cell chunk, the id of the first dual vertex
// and per cell its patch count - 1 (two bit planes) and whether its case index is complemented
// (ambiguity resolution, cudualmc.cu:815-839).
#pragma once
#include "mc.cuh"

namespace diso {

// Everything a consumer needs about one cell: id of its first dual vertex + packed case entry.
struct CellInfo {
    unsigned first;  // id of the cell's first dual vertex
    unsigned ce;     // T_DMC_CASE[good code]  (0 for unused cells)
    unsigned code;   // good (possibly complemented) case index
};

// Cell info for cell `j` of chunk `k`, computed from the sign words and the patch record.
__device__ __forceinline__ CellInfo dmc_cell_info(const unsigned *__restrict__ S, const uint4 *__restrict__ P,
                                                  const Geo &g, const unsigned *__restrict__ s_case, int k, int j)
{
    const CellWords w = load_cell_words(S, g, k);
    const unsigned used = used_mask(w);
    const uint4 p = P[k];
    const unsigned lt = lanemask_lt(j);
    CellInfo ci;
    ci.first = p.x + __popc(used & lt) + __popc(p.y & lt) + 2 * __popc(p.z & lt);
    unsigned code = cell_code<DISO_ALG_DMC>(w, j);
    if (bit(p.w, j)) code ^= 0xffu;
    ci.code = code;
    ci.ce = bit(used, j) ? s_case[code] : 0u;
    return ci;
}

__device__ __forceinline__ unsigned dual_id(const CellInfo &c, int eid) { return c.first + ((c.ce >> (2 * eid)) & 3u); }

// ------------------------------------------------------------------------------------------
// K3d: dual vertices.  Replaces create_dmc_verts_kernel (cudualmc.cu:907-955) + epilogue
// (diso/__init__.py:110-114).  Lane == cell.  The 8 corner values / deformations are fetched
// once per cell; each crossing edge's vertex is formed exactly like computeMcVert and added
// to its patch accumulator in ascending edge id, which is the order of the reference's patch
// table (asserted in tools/extract_tables.py), then scaled by 1/len.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EMIT_WARPS * 32) dmc_emit_verts_kernel(const T *__restrict__ sdf,
                                                                       const T *__restrict__ deform, Geo g, T iso,
                                                                       T padv, Epilogue<T> epi,
                                                                       const unsigned *__restrict__ S,
                                                                       const uint4 *__restrict__ P,
                                                                       unsigned short *__restrict__ C,
                                                                       T *__restrict__ verts)
{
    __shared__ unsigned s_case[256];
    __shared__ unsigned s_plen[256];
    __shared__ T s_stage[EMIT_WARPS][128 * 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_case[threadIdx.x] = T_DMC_CASE[threadIdx.x];
    s_plen[threadIdx.x] = T_DMC_PATCHLEN[threadIdx.x];
    __syncthreads();
    const int group = blockIdx.x * EMIT_WARPS + wid;
    const int k0 = group * 32;
    if (k0 >= g.NCH) return;
    const bool has_def = deform != nullptr;
    T *stage = s_stage[wid];

    unsigned p_lo = 0, p_hi = 0;
    if (k0 + lane < g.NCH) { p_lo = P[k0 + lane].x; p_hi = P[k0 + lane + 1].x; }
    unsigned active = __ballot_sync(FULL, p_hi != p_lo);
    while (active) {
        const int i = __ffs(active) - 1;
        active &= active - 1;
        const int k = k0 + i;
        const unsigned vbase = __shfl_sync(FULL, p_lo, i);
        const unsigned nvert = __shfl_sync(FULL, p_hi, i) - vbase;
        const ChunkPos cp = chunk_pos(g, k);
        const int xp = cp.xp, yp = cp.yp, zp = 32 * cp.c + lane;
        const CellInfo ci = dmc_cell_info(S, P, g, s_case, k, lane);
        C[(size_t)k * 32 + lane] = ci.ce ? (unsigned short)(ci.code | ((ci.first - vbase) << 8)) : (unsigned short)0;
        if (ci.ce) {
            const unsigned np = (ci.ce >> 24) & 7u;
            const unsigned plen = s_plen[ci.code];
            const unsigned cross = plen >> 16;
            T d[8];
            Vec3<T> f[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                d[c] = fetch_padded(sdf, g, xp + (c & 1), yp + ((c >> 1) & 1), zp + (c >> 2), padv);
                f[c] = Vec3<T>{T(0), T(0), T(0)};
                if (has_def) f[c] = fetch_deform(deform, g, xp + (c & 1), yp + ((c >> 1) & 1), zp + (c >> 2));
            }
            Vec3<T> acc[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[p] = Vec3<T>{T(0), T(0), T(0)};
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                if ((cross >> e) & 1u) {
                    const int c0 = edge_dx(e) | (edge_dy(e) << 1) | (edge_dz(e) << 2);
                    const int ax = edge_axis(e);
                    const int c1 = c0 + (1 << ax);
                    Vec3<T> v;
                    const int ex = xp + edge_dx(e), ey = yp + edge_dy(e), ez = zp + edge_dz(e);
                    if (ax == 0) v = edge_vertex<T, 0>(d[c0], d[c1], iso, ex, ey, ez, has_def, f[c0], f[c1]);
                    else if (ax == 1) v = edge_vertex<T, 1>(d[c0], d[c1], iso, ex, ey, ez, has_def, f[c0], f[c1]);
                    else v = edge_vertex<T, 2>(d[c0], d[c1], iso, ex, ey, ez, has_def, f[c0], f[c1]);
                    const unsigned p = (ci.ce >> (2 * e)) & 3u;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (p == (unsigned)q) { acc[q].x = acc[q].x + v.x; acc[q].y = acc[q].y + v.y; acc[q].z = acc[q].z + v.z; }
                }
            }
            const unsigned slot0 = ci.first - vbase;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((unsigned)q < np) {
                    const T inv = T(1) / T((plen >> (4 * q)) & 0xfu);
                    Vec3<T> v{acc[q].x * inv, acc[q].y * inv, acc[q].z * inv};
                    v = epi.apply(v);
                    stage[3 * (slot0 + q)] = v.x; stage[3 * (slot0 + q) + 1] = v.y; stage[3 * (slot0 + q) + 2] = v.z;
                }
            }
        }
        __syncwarp();
        T *dst = verts + (size_t)vbase * 3;
        for (unsigned q = lane; q < 3 * nvert; q += 32) st_stream(dst + q, stage[q]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K4d / K5d-A: per crossing edge, visit the four cells around it.
//   MODE 0: write the quad (4 dual-vertex ids, int64).  Replaces index_cell_mc_verts +
//           create_quads (cudualmc.cu:753-794, 1019-1056) + int64 widening.
//   MODE 1: gather dL/d(edge vertex) = sum_i adj_dual[id_i] / len_i        (exact adjoint)
//   MODE 2: same, but every patch reads the adjoint of its cell's FIRST dual vertex --
//           bug-compatible with cudualmc.cu:975,990 where `first` is never advanced.
// The four (cell offset, local edge id) pairs per axis are those of the reference's dmcQuad
// table (cudualmc.cu:504-513); for an "exiting" edge (start point inside, d0 >= iso > d1)
// corners 1 and 3 are swapped (reversed winding) -- asserted in tools/extract_tables.py.
// ------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(EMIT_WARPS * 32) dmc_edges_kernel(Geo g, const unsigned *__restrict__ S,
                                                                  const uint4 *__restrict__ E,
                                                                  const uint4 *__restrict__ P, Epilogue<T> epi,
                                                                  const T *__restrict__ adj_dual,
                                                                  long long *__restrict__ quads,
                                                                  T *__restrict__ gedge)
{
    __shared__ unsigned s_case[256];
    __shared__ unsigned s_plen[256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_case[threadIdx.x] = T_DMC_CASE[threadIdx.x];
    s_plen[threadIdx.x] = T_DMC_PATCHLEN[threadIdx.x];
    __syncthreads();
    const int group = blockIdx.x * EMIT_WARPS + wid;
    const int k0 = group * 32;
    if (k0 >= g.NCH) return;

    uint4 mine = make_uint4(0, 0, 0, 0);
    if (k0 + lane < g.NCH) mine = E[k0 + lane];
    unsigned active = __ballot_sync(FULL, (mine.y | mine.z | mine.w) != 0u);
    while (active) {
        const int i = __ffs(active) - 1;
        active &= active - 1;
        const int k = k0 + i;
        uint4 e;
        e.x = __shfl_sync(FULL, mine.x, i); e.y = __shfl_sync(FULL, mine.y, i);
        e.z = __shfl_sync(FULL, mine.z, i); e.w = __shfl_sync(FULL, mine.w, i);
        const ChunkPos cp = chunk_pos(g, k);
        const bool bx = bit(e.y, lane), by = bit(e.z, lane), bz = bit(e.w, lane);
        const bool inside = bit(S[k], lane);  // start point value >= iso  -> crossing edges are "exiting"

        // cell info for the four rows at my z, then the z-1 variants by shuffle
        CellInfo c00 = dmc_cell_info(S, P, g, s_case, k, lane);
        CellInfo c0m{0, 0, 0}, cm0{0, 0, 0}, cmm{0, 0, 0};
        if (cp.yp > 0) c0m = dmc_cell_info(S, P, g, s_case, k - g.sY, lane);
        if (cp.xp > 0) cm0 = dmc_cell_info(S, P, g, s_case, k - g.sX, lane);
        if (cp.xp > 0 && cp.yp > 0) cmm = dmc_cell_info(S, P, g, s_case, k - g.sX - g.sY, lane);
        CellInfo c00z, c0mz, cm0z;
        c00z.first = __shfl_up_sync(FULL, c00.first, 1); c00z.ce = __shfl_up_sync(FULL, c00.ce, 1); c00z.code = __shfl_up_sync(FULL, c00.code, 1);
        c0mz.first = __shfl_up_sync(FULL, c0m.first, 1); c0mz.ce = __shfl_up_sync(FULL, c0m.ce, 1); c0mz.code = __shfl_up_sync(FULL, c0m.code, 1);
        cm0z.first = __shfl_up_sync(FULL, cm0.first, 1); cm0z.ce = __shfl_up_sync(FULL, cm0.ce, 1); cm0z.code = __shfl_up_sync(FULL, cm0.code, 1);
        if (lane == 0 && (bx | by)) {
            // z-1 lives in the previous chunk of the same row (c > 0 is guaranteed: the pad
            // column zp = 0 owns no crossing x/y edge)
            c00z = dmc_cell_info(S, P, g, s_case, k - 1, 31);
            if (bx) c0mz = dmc_cell_info(S, P, g, s_case, k - g.sY - 1, 31);
            if (by) cm0z = dmc_cell_info(S, P, g, s_case, k - g.sX - 1, 31);
        }

        const RowRank r = row_rank(e, lane);
        auto emit = [&](unsigned rank, const CellInfo &q0, int e0, const CellInfo &q1, int e1, const CellInfo &q2,
                        int e2, const CellInfo &q3, int e3) {
            // corners in "entering" order; exiting swaps 1 <-> 3
            if (MODE == 0) {
                const long long i0 = dual_id(q0, e0), i2 = dual_id(q2, e2);
                long long i1 = dual_id(q1, e1), i3 = dual_id(q3, e3);
                if (inside) { long long t = i1; i1 = i3; i3 = t; }
                longlong2 *dst = reinterpret_cast<longlong2 *>(quads + (size_t)rank * 4);
                __stcs(dst, make_longlong2(i0, i1));
                __stcs(dst + 1, make_longlong2(i2, i3));
            } else {
                Vec3<T> acc{T(0), T(0), T(0)};
                auto add = [&](const CellInfo &q, int eid) {
                    const unsigned ord = (q.ce >> (2 * eid)) & 3u;
                    const unsigned src = (MODE == 1) ? q.first + ord : q.first;
                    const T inv = T(1) / T((s_plen[q.code] >> (4 * ord)) & 0xfu);
                    const T *p = adj_dual + (size_t)src * 3;
                    Vec3<T> a = epi.adjoint(Vec3<T>{__ldg(p), __ldg(p + 1), __ldg(p + 2)});
                    acc.x = acc.x + a.x * inv; acc.y = acc.y + a.y * inv; acc.z = acc.z + a.z * inv;
                };
                add(q0, e0); add(q1, e1); add(q2, e2); add(q3, e3);
                T *dst = gedge + (size_t)rank * 3;
                dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z;
            }
        };
        if (bx) emit(r.start, c00, 0, c0m, 4, c0mz, 6, c00z, 2);
        if (by) emit(r.start + r.bx, c00, 8, c00z, 11, cm0z, 10, cm0, 9);
        if (bz) emit(r.start + r.bx + r.by, c00, 3, cm0, 1, cmm, 5, c0m, 7);
    }
}

// ------------------------------------------------------------------------------------------
// Diagnostics: dense per-cell case index (see diso_b200_debug_cell_codes in the header).
// ------------------------------------------------------------------------------------------
template <int ALG>
__global__ void debug_codes_kernel(Geo g, const unsigned *__restrict__ S, const uint4 *__restrict__ P,
                                   unsigned char *__restrict__ codes)
{
    const long long n = (long long)g.PX * g.PY * g.PZ;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int zp = (int)(idx % g.PZ);
    const long long r = idx / g.PZ;
    const int k = (int)(r * g.NC + (zp >> 5)), j = zp & 31;
    const CellWords w = load_cell_words(S, g, k);
    unsigned code = cell_code<ALG>(w, j);
    if (ALG == DISO_ALG_DMC && bit(P[k].w, j)) code ^= 0xffu;
    if (!bit(used_mask(w), j)) code = (code == 0u || code == 0xffu) ? code : 0u;
    codes[idx] = (unsigned char)code;
}

}  // namespace diso
