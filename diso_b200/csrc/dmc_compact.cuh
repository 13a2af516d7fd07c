// dmc_compact.cuh -- edge-parallel DMC kernels on the compaction pattern of compact.cuh.
//
// K4d  quads            : replaces index_cell_mc_verts + create_quads (cudualmc.cu:753-794,
//                         1019-1056) + the int64 widening (diso/__init__.py:116).
// K5dA edge adjoints    : stage A of the backward (adj_create_dmc_verts, cudualmc.cu:957-1005):
//                         dL/d(edge crossing) = sum over the 4 cells around the edge of
//                         adj_dual[patch of the edge in that cell] / len(patch).
// Both visit, per crossing edge, the four cells around it.  A cell's data (possibly complemented
// case index + id of its first dual vertex) is read from the per-cell array C written once by
// classify_scan, instead of being re-derived from 8 sign words 7 times per grid point.
#pragma once
#include "compact.cuh"
#include "tables.cuh"

namespace diso {

// MODE 0: quads.  MODE 1: exact adjoint.  MODE 2: reference-compatible adjoint (every patch of a
// cell reads the adjoint of the cell's FIRST dual vertex, cudualmc.cu:975,990).
template <typename T, int MODE>
__global__ void __launch_bounds__(CT_THREADS) dmc_edges2_kernel(Geo g, const unsigned *__restrict__ S,
                                                              const uint4 *__restrict__ E, const uint4 *__restrict__ P,
                                                              const unsigned short *__restrict__ C,
                                                              const unsigned *__restrict__ alist, int n_active, T ix, T iy, T iz,
                                                              const T *__restrict__ adj_dual,
                                                              long long *__restrict__ quads, T *__restrict__ gedge)
{
    __shared__ unsigned short s_list[CT_MAX_EDGES];
    __shared__ unsigned s_case[256];
    __shared__ unsigned s_plen[256];
    __shared__ unsigned s_quad[8];
    __shared__ T s_inv[8];
    __shared__ int s_k[CT_CHUNKS];
    s_case[threadIdx.x] = T_DMC_CASE[threadIdx.x];
    if (MODE != 0) s_plen[threadIdx.x] = T_DMC_PATCHLEN[threadIdx.x];
    if (threadIdx.x < 6) s_quad[threadIdx.x] = T_DMC_QUAD[threadIdx.x];
    if (threadIdx.x < 8) s_inv[threadIdx.x] = threadIdx.x ? T(1) / T((int)threadIdx.x) : T(0);
    const int count = load_tile_chunks(alist, n_active, s_k);
    unsigned tile_base;
    const unsigned n = build_edge_list(g, E, s_k, count, s_list, nullptr, tile_base, S);
    if (n == 0) return;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const int axis = d & 3, j = (d >> 2) & 31, cl = (d >> 7) & 63, inside = (d >> 13) & 1;
        const int k = s_k[cl];
        const unsigned q4 = s_quad[inside * 3 + axis];  // reference dmcQuad[type], type = (exiting ? 3 : 0) + axis
        long long id[4];
        Vec3<T> acc{T(0), T(0), T(0)};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned b = (q4 >> (8 * c)) & 0xffu;
            int kk = k - (int)(b & 1u) * g.sX - (int)((b >> 1) & 1u) * g.sY;
            int jj = j - (int)((b >> 2) & 1u);
            if (jj < 0) { kk -= 1; jj = 31; }
            const unsigned info = C[(size_t)kk * 32 + jj];
            const unsigned first = P[kk].x + (info >> 8);
            const unsigned code = info & 0xffu;
            const unsigned eid = b >> 4;
            const unsigned ord = (s_case[code] >> (2 * eid)) & 3u;
            if (MODE == 0) {
                id[c] = (long long)(first + ord);
            } else {
                const unsigned src = (MODE == 1) ? first + ord : first;
                const T inv = s_inv[(s_plen[code] >> (4 * ord)) & 7u];
                const T *p = adj_dual + (size_t)src * 3;
                acc.x = fma_rn(__ldg(p), inv, acc.x);
                acc.y = fma_rn(__ldg(p + 1), inv, acc.y);
                acc.z = fma_rn(__ldg(p + 2), inv, acc.z);
            }
        }
        const size_t rank = (size_t)tile_base + i;
        if (MODE == 0) {
            longlong2 *dst = reinterpret_cast<longlong2 *>(quads + rank * 4);
            __stcs(dst, make_longlong2(id[0], id[1]));
            __stcs(dst + 1, make_longlong2(id[2], id[3]));
        } else {
            T *dst = gedge + rank * 3;
            dst[0] = acc.x * ix; dst[1] = acc.y * iy; dst[2] = acc.z * iz;
        }
    }
}

}  // namespace diso

namespace diso {

constexpr int CT_MAX_PATCHES = CT_CHUNKS * 128;

// ------------------------------------------------------------------------------------------
// K3d (v2): dual vertices, patch-parallel.  Replaces create_dmc_verts_kernel
// (cudualmc.cu:907-955) + epilogue (diso/__init__.py:110-114).
// The reference recomputes every edge crossing on the fly in each of the 4 cells around the
// edge (cudualmc.cu:946-948).  Here the crossings are evaluated ONCE by edge_verts_kernel in
// the raw padded frame into `mcv` (caller scratch, [n_edges,3]) and this kernel only averages:
//   phase A  lane == cell: the per-cell word written by classify_scan ((possibly complemented)
//            case index | offset of the cell's first dual vertex) -> one descriptor per patch in
//            the shared list at slot (dual vertex id - first id of the tile).
//   phase B  thread == dual vertex: gather the crossings of the patch's member edges by rank in
//            ascending edge id (== the reference's table order, asserted in
//            tools/extract_tables.py), sum, scale by 1/len, apply the epilogue, store at
//            consecutive output ranks.  Same values in the same order => bit-identical.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CT_THREADS) dmc_dual_verts_kernel(const T *__restrict__ mcv, Geo g, EpilogueC<T> epi,
                                                                  const uint4 *__restrict__ E,
                                                                  const uint4 *__restrict__ P,
                                                                  const unsigned short *__restrict__ C,
                                                                  const unsigned *__restrict__ alist, int n_active,
                                                                  T *__restrict__ verts)
{
    __shared__ unsigned short s_list[CT_MAX_PATCHES];
    __shared__ unsigned short s_cell[CT_CHUNKS * 32];
    __shared__ unsigned s_case[256];
    __shared__ unsigned s_plen[256];
    __shared__ unsigned long long s_members[256];
    __shared__ uint4 s_E[CT_RECS];
    __shared__ T s_inv[8];
    __shared__ int s_k[CT_CHUNKS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int count = load_tile_chunks(alist, n_active, s_k);
    const unsigned tile_base = P[s_k[0]].x;
    const unsigned n = P[s_k[count - 1] + 1].x - tile_base;
    if (n == 0) return;
    s_case[threadIdx.x] = T_DMC_CASE[threadIdx.x];
    s_plen[threadIdx.x] = T_DMC_PATCHLEN[threadIdx.x];
    s_members[threadIdx.x] = T_DMC_MEMBERS[threadIdx.x];
    const RecCache rc = load_record_cache(g, E, s_k, count, s_E);
    if (threadIdx.x < 8) s_inv[threadIdx.x] = threadIdx.x ? T(1) / T((int)threadIdx.x) : T(0);
    __syncthreads();

    // ---- phase A ---------------------------------------------------------------------------------
    constexpr int PER_WARP = CT_CHUNKS / CT_WARPS;
    {
        const int emine = wid * PER_WARP + lane;
        const int kmine = (lane < PER_WARP && emine < count) ? s_k[emine] : -1;
        uint4 pr = make_uint4(0, 0, 0, 0);
        if (kmine >= 0) pr = P[kmine];
        unsigned active = __ballot_sync(FULL, pr.y != 0u);
        while (active) {
            const int i = __ffs(active) - 1;
            active &= active - 1;
            const int cl = wid * PER_WARP + i;
            const int k = __shfl_sync(FULL, kmine, i);
            const unsigned pb = __shfl_sync(FULL, pr.x, i) - tile_base;
            const unsigned used = __shfl_sync(FULL, pr.y, i);
            const unsigned info = bit(used, lane) ? C[(size_t)k * 32 + lane] : 0u;
            s_cell[cl * 32 + lane] = (unsigned short)info;
            const unsigned np = bit(used, lane) ? (s_case[info & 0xffu] >> 24) & 7u : 0u;
            const unsigned slot = pb + (info >> 8);
#pragma unroll
            for (unsigned q = 0; q < 4; ++q)
                if (q < np) s_list[slot + q] = (unsigned short)((cl << 7) | (lane << 2) | q);
        }
    }
    __syncthreads();

    // ---- phase B ---------------------------------------------------------------------------------
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const unsigned q = d & 3u;
        const int j = (d >> 2) & 31, cl = (d >> 7) & 63;
        const unsigned code = s_cell[cl * 32 + j] & 0xffu;
        const unsigned plen = s_plen[code];
        unsigned mm = (unsigned)(s_members[code] >> (12 * q)) & 0xfffu;  // member edges of this patch
        // pass 1: ranks of the (<= 7) member edges, ascending edge id
        unsigned rank[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            rank[t] = 0xffffffffu;
            if (mm) {
                const int e = __ffs(mm) - 1;
                mm &= mm - 1;
                const unsigned r = edge_rank(rc, cl, j, e);
                rank[t] = r;
            }
        }
        // pass 2: gather + sum in the same (ascending edge id) order
        T vx[7], vy[7], vz[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            vx[t] = vy[t] = vz[t] = T(0);
            if (rank[t] != 0xffffffffu) {
                const T *pv = mcv + (size_t)rank[t] * 3;
                vx[t] = __ldg(pv); vy[t] = __ldg(pv + 1); vz[t] = __ldg(pv + 2);
            }
        }
        Vec3<T> acc{T(0), T(0), T(0)};
#pragma unroll
        for (int t = 0; t < 7; ++t)
            if (rank[t] != 0xffffffffu) { acc.x = acc.x + vx[t]; acc.y = acc.y + vy[t]; acc.z = acc.z + vz[t]; }
        const T inv = s_inv[(plen >> (4 * q)) & 7u];
        Vec3<T> v{acc.x * inv, acc.y * inv, acc.z * inv};
        v = epi.apply(v);
        T *dst = verts + (size_t)(tile_base + i) * 3;
        st_stream(dst, v.x); st_stream(dst + 1, v.y); st_stream(dst + 2, v.z);
    }
}

}  // namespace diso
