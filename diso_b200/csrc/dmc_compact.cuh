// dmc_compact.cuh -- edge-parallel DMC kernels on the compaction pattern of compact.cuh.
//
// K4d  quads            : replaces index_cell_mc_verts + create_quads (cudualmc.cu:753-794,
//                         1019-1056) + the int64 widening (diso/__init__.py:116).
// K5dA edge adjoints    : stage A of the backward (adj_create_dmc_verts, cudualmc.cu:957-1005):
//                         dL/d(edge crossing) = sum over the 4 cells around the edge of
//                         adj_dual[patch of the edge in that cell] / len(patch).
// Both visit, per crossing edge, the four cells around it.  A cell's data (possibly complemented
// case index + id of its first dual vertex) is read from the per-cell array C written once by the
// dual-vertex kernel, instead of being re-derived from 8 sign words 7 times per grid point.
#pragma once
#include "compact.cuh"
#include "tables.cuh"

namespace diso {

// MODE 0: quads.  MODE 1: exact adjoint.  MODE 2: reference-compatible adjoint (every patch of a
// cell reads the adjoint of the cell's FIRST dual vertex, cudualmc.cu:975,990).
template <typename T, int MODE>
__global__ void __launch_bounds__(CT_THREADS) dmc_edges2_kernel(Geo g, const unsigned *__restrict__ S,
                                                              const uint4 *__restrict__ E, const uint4 *__restrict__ P,
                                                              const unsigned short *__restrict__ C, T ix, T iy, T iz,
                                                              const T *__restrict__ adj_dual,
                                                              long long *__restrict__ quads, T *__restrict__ gedge)
{
    __shared__ unsigned short s_list[CT_MAX_EDGES];
    __shared__ TilePos s_pos[CT_CHUNKS];
    __shared__ unsigned s_case[256];
    __shared__ unsigned s_plen[256];
    __shared__ unsigned s_quad[8];
    __shared__ T s_inv[8];
    s_case[threadIdx.x] = T_DMC_CASE[threadIdx.x];
    if (MODE != 0) s_plen[threadIdx.x] = T_DMC_PATCHLEN[threadIdx.x];
    if (threadIdx.x < 6) s_quad[threadIdx.x] = T_DMC_QUAD[threadIdx.x];
    if (threadIdx.x < 8) s_inv[threadIdx.x] = threadIdx.x ? T(1) / T((int)threadIdx.x) : T(0);
    const int k0 = blockIdx.x * CT_CHUNKS;
    unsigned tile_base;
    const unsigned n = build_edge_list(g, E, k0, s_list, s_pos, tile_base, S);
    if (n == 0) return;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const int axis = d & 3, j = (d >> 2) & 31, cl = (d >> 7) & 63, inside = (d >> 13) & 1;
        const int k = k0 + cl;
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
