// mc.cuh -- marching cubes phase 2 (emit) and backward.
//
// All three kernels walk the chunk list in the reference's order (ascending padded linear
// index); one warp owns a chunk at a time, lane j == padded point / cell j of the chunk.
// A warp first inspects 32 chunk records with one coalesced load and a ballot, then visits
// only the chunks that have work, so sparse surfaces cost ~one record read per 32 points.
#pragma once
#include "classify.cuh"
#include "edge_math.cuh"

namespace diso {

constexpr int EMIT_WARPS = 8;  // warps per CTA in the emit / backward kernels

struct ChunkPos { int xp, yp, c; };
__device__ __forceinline__ ChunkPos chunk_pos(const Geo &g, int k)
{
    ChunkPos p;
    int r = k / g.NC;
    p.c = k - r * g.NC;
    p.xp = r / g.PY;
    p.yp = r - p.xp * g.PY;
    return p;
}

// ------------------------------------------------------------------------------------------
// K3: vertices.  Replaces create_cell_mc_verts_kernel (cumc.cu:370-410) and the "-1" /
// normalise epilogue (diso/__init__.py:56-60).  Output order: point-major, axis-minor ==
// the reference's (used cell, x/y/z edge) order.  Vertices of a chunk are staged in shared
// memory and written with fully coalesced stores.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EMIT_WARPS * 32) mc_emit_verts_kernel(const T *__restrict__ sdf,
                                                                      const T *__restrict__ deform, Geo g, T iso,
                                                                      T padv, Epilogue<T> epi,
                                                                      const uint4 *__restrict__ E,
                                                                      T *__restrict__ verts)
{
    __shared__ T s_stage[EMIT_WARPS][96 * 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int group = blockIdx.x * EMIT_WARPS + wid;
    const int k0 = group * 32;
    if (k0 >= g.NCH) return;
    const bool has_def = deform != nullptr;
    T *stage = s_stage[wid];

    uint4 mine = make_uint4(0, 0, 0, 0);
    if (k0 + lane < g.NCH) mine = E[k0 + lane];
    unsigned active = __ballot_sync(FULL, (mine.y | mine.z | mine.w) != 0u);
    while (active) {
        const int i = __ffs(active) - 1;
        active &= active - 1;
        const int k = k0 + i;
        const unsigned base = __shfl_sync(FULL, mine.x, i);
        const unsigned mx = __shfl_sync(FULL, mine.y, i);
        const unsigned my = __shfl_sync(FULL, mine.z, i);
        const unsigned mz = __shfl_sync(FULL, mine.w, i);
        const ChunkPos cp = chunk_pos(g, k);
        const int xp = cp.xp, yp = cp.yp, zp = 32 * cp.c + lane;
        const unsigned lt = lanemask_lt(lane);
        int slot = __popc(mx & lt) + __popc(my & lt) + __popc(mz & lt);
        const bool bx = bit(mx, lane), by = bit(my, lane), bz = bit(mz, lane);

        const T d0 = fetch_padded(sdf, g, xp, yp, zp, padv);
        T dzv = __shfl_down_sync(FULL, d0, 1);
        if (lane == 31 && bz) dzv = fetch_padded(sdf, g, xp, yp, zp + 1, padv);
        Vec3<T> f0{T(0), T(0), T(0)};
        if (has_def && (bx | by | bz)) f0 = fetch_deform(deform, g, xp, yp, zp);
        if (bx) {
            const T d1 = fetch_padded(sdf, g, xp + 1, yp, zp, padv);
            Vec3<T> f1{T(0), T(0), T(0)};
            if (has_def) f1 = fetch_deform(deform, g, xp + 1, yp, zp);
            Vec3<T> p = epi.apply(edge_vertex<T, 0>(d0, d1, iso, xp, yp, zp, has_def, f0, f1));
            stage[3 * slot] = p.x; stage[3 * slot + 1] = p.y; stage[3 * slot + 2] = p.z;
            ++slot;
        }
        if (by) {
            const T d1 = fetch_padded(sdf, g, xp, yp + 1, zp, padv);
            Vec3<T> f1{T(0), T(0), T(0)};
            if (has_def) f1 = fetch_deform(deform, g, xp, yp + 1, zp);
            Vec3<T> p = epi.apply(edge_vertex<T, 1>(d0, d1, iso, xp, yp, zp, has_def, f0, f1));
            stage[3 * slot] = p.x; stage[3 * slot + 1] = p.y; stage[3 * slot + 2] = p.z;
            ++slot;
        }
        if (bz) {
            Vec3<T> f1{T(0), T(0), T(0)};
            if (has_def) f1 = fetch_deform(deform, g, xp, yp, zp + 1);
            Vec3<T> p = epi.apply(edge_vertex<T, 2>(d0, dzv, iso, xp, yp, zp, has_def, f0, f1));
            stage[3 * slot] = p.x; stage[3 * slot + 1] = p.y; stage[3 * slot + 2] = p.z;
        }
        __syncwarp();
        const int n3 = 3 * (__popc(mx) + __popc(my) + __popc(mz));
        T *dst = verts + (size_t)base * 3;
        for (int q = lane; q < n3; q += 32) st_stream(dst + q, stage[q]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Vertex id ("rank") of every edge of every cell of a chunk, from four edge records.
// r[e] for the 12 local edges (numbering of the reference diagram, cumc.cu:43-58; owner
// offsets / axes per mcEdgeLocations, cumc.cu:109-122).
// ------------------------------------------------------------------------------------------
struct RowRank { unsigned start, bx, by, n; };
__device__ __forceinline__ RowRank row_rank(const uint4 &e, int lane)
{
    const unsigned lt = lanemask_lt(lane);
    RowRank r;
    r.start = e.x + __popc(e.y & lt) + __popc(e.z & lt) + __popc(e.w & lt);
    r.bx = bit(e.y, lane);
    r.by = bit(e.z, lane);
    r.n = r.bx + r.by + bit(e.w, lane);
    return r;
}

// ------------------------------------------------------------------------------------------
// K4: triangles.  Replaces count_cell_mc_tris / create_cell_mc_tris (cumc.cu:540-612) and the
// int64 widening (diso/__init__.py:61).  The owner-cell lookup + linear search of the
// reference (cumc.cu:589-607) becomes popcount arithmetic on the edge records.  Indices are
// produced index-parallel: lane q of a round emits output element 32*round + q, so stores are
// dense 8-byte coalesced.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EMIT_WARPS * 32) mc_emit_tris_kernel(Geo g, const unsigned *__restrict__ S,
                                                                     const uint4 *__restrict__ E,
                                                                     const unsigned *__restrict__ F,
                                                                     long long *__restrict__ tris)
{
    __shared__ unsigned long long s_case[256];
    __shared__ unsigned s_rank[EMIT_WARPS][12][32];
    __shared__ unsigned s_tri[EMIT_WARPS][160];  // per triangle: cell lane | 3 edge ids << 8
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_case[threadIdx.x] = T_MC_CASE[threadIdx.x];
    __syncthreads();
    const int group = blockIdx.x * EMIT_WARPS + wid;
    const int k0 = group * 32;
    if (k0 >= g.NCH) return;

    unsigned f_lo = 0, f_hi = 0;
    if (k0 + lane < g.NCH) { f_lo = F[k0 + lane]; f_hi = F[k0 + lane + 1]; }
    unsigned active = __ballot_sync(FULL, f_hi != f_lo);
    while (active) {
        const int i = __ffs(active) - 1;
        active &= active - 1;
        const int k = k0 + i;
        const unsigned tbase = __shfl_sync(FULL, f_lo, i);
        const unsigned ntot = __shfl_sync(FULL, f_hi, i) - tbase;

        const CellWords w = load_cell_words(S, g, k);
        const unsigned used = used_mask(w);
        const unsigned code = bit(used, lane) ? cell_code<DISO_ALG_MC>(w, lane) : 0u;
        const unsigned long long entry = s_case[code];
        const unsigned nt = (unsigned)(entry >> 60);
        unsigned incl = nt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        unsigned excl = incl - nt;
#pragma unroll
        for (unsigned q = 0; q < 5; ++q)
            if (q < nt) s_tri[wid][excl + q] = (unsigned)lane | (((unsigned)(entry >> (12 * q)) & 0xfffu) << 8);

        // vertex ids of the 12 edges of my cell
        const uint4 e00 = E[k], e01 = E[k + g.sY], e10 = E[k + g.sX], e11 = E[k + g.sX + g.sY];
        const unsigned mx00n = E[k + 1].y, mx10n = E[k + g.sX + 1].y;
        const RowRank r00 = row_rank(e00, lane), r01 = row_rank(e01, lane);
        const RowRank r10 = row_rank(e10, lane), r11 = row_rank(e11, lane);
        unsigned(*rk)[32] = s_rank[wid];
        rk[0][lane] = r00.start;
        rk[8][lane] = r00.start + r00.bx;
        rk[3][lane] = r00.start + r00.bx + r00.by;
        rk[2][lane] = r00.start + r00.n;
        rk[11][lane] = r00.start + r00.n + bit(shift_in(e00.y, mx00n), lane);
        rk[4][lane] = r01.start;
        rk[7][lane] = r01.start + r01.bx + r01.by;
        rk[6][lane] = r01.start + r01.n;
        rk[9][lane] = r10.start + r10.bx;
        rk[1][lane] = r10.start + r10.bx + r10.by;
        rk[10][lane] = r10.start + r10.n + bit(shift_in(e10.y, mx10n), lane);
        rk[5][lane] = r11.start + r11.bx + r11.by;
        __syncwarp();

        long long *dst = tris + (size_t)tbase * 3;
        const unsigned n3 = 3 * ntot;
        for (unsigned q = lane; q < n3; q += 32) {
            const unsigned t = q / 3, which = q - 3 * t;
            const unsigned rec = s_tri[wid][t];
            const unsigned e = (rec >> (8 + 4 * which)) & 0xfu;
            st_stream(dst + q, (long long)rk[e][rec & 31u]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K5: backward as an atomic-free gather.  Replaces adj_create_cell_mc_verts_kernel
// (cumc.cu:474-512), the dense zero fills (diso/__init__.py:33,40) and the pad-backward
// slices.  The thread of real grid point P sums the adjoint contributions of its <= 6
// incident crossing edges in the fixed order (+x,+y,+z owned edges, then the edges arriving
// from -x,-y,-z) and writes adj_sdf[P] / adj_deform[P] exactly once, zeros included.
// `gsrc` is dL/dvertex per crossing edge, [n_edges,3]; `epi` carries the 1/(dim-1) chain rule
// (identity for the DMC second stage).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EMIT_WARPS * 32) mc_backward_kernel(const T *__restrict__ sdf,
                                                                    const T *__restrict__ deform, Geo g, T iso,
                                                                    T padv, Epilogue<T> epi,
                                                                    const uint4 *__restrict__ E,
                                                                    const T *__restrict__ gsrc,
                                                                    T *__restrict__ adj_sdf,
                                                                    T *__restrict__ adj_deform)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k = blockIdx.x * EMIT_WARPS + wid;  // one chunk per warp
    if (k >= g.NCH) return;
    const ChunkPos cp = chunk_pos(g, k);
    const int xp = cp.xp, yp = cp.yp, zp = 32 * cp.c + lane;
    if (xp < 1 || xp > g.X || yp < 1 || yp > g.Y) return;  // pad rows receive no gradient
    const bool real = zp >= 1 && zp <= g.Z;
    const bool has_def = deform != nullptr;

    const uint4 e = E[k];
    const uint4 ex = E[k - g.sX];              // row (xp-1, yp)
    const uint4 ey = E[k - g.sY];              // row (xp, yp-1)
    const uint4 ezp = E[k > 0 ? k - 1 : 0];    // previous chunk of this row (only its bit 31 matters)
    // incoming -z edge of lane j is the +z edge of lane j-1
    const unsigned mz_in = (e.w << 1) | ((cp.c > 0) ? (ezp.w >> 31) : 0u);

    const bool ox = bit(e.y, lane), oy = bit(e.z, lane), oz = bit(e.w, lane);
    const bool ix = bit(ex.y, lane), iy = bit(ey.z, lane), iz = bit(mz_in, lane);
    const bool any = real && (ox | oy | oz | ix | iy | iz);

    T acc_d = T(0);
    Vec3<T> acc_f{T(0), T(0), T(0)};
    if (__any_sync(FULL, any)) {
        const T dme = real ? fetch_padded(sdf, g, xp, yp, zp, padv) : padv;
        T d_zp = __shfl_down_sync(FULL, dme, 1);
        T d_zm = __shfl_up_sync(FULL, dme, 1);
        if (any) {
            if (lane == 31 && oz) d_zp = fetch_padded(sdf, g, xp, yp, zp + 1, padv);
            if (lane == 0 && iz) d_zm = fetch_padded(sdf, g, xp, yp, zp - 1, padv);
            Vec3<T> fme{T(0), T(0), T(0)};
            if (has_def) fme = fetch_deform(deform, g, xp, yp, zp);
            const RowRank r = row_rank(e, lane);
            auto load_g = [&](unsigned id) {
                const T *p = gsrc + (size_t)id * 3;
                Vec3<T> v{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
                return epi.adjoint(v);
            };
            const Vec3<T> zero{T(0), T(0), T(0)};
            if (ox) {
                const T d1 = fetch_padded(sdf, g, xp + 1, yp, zp, padv);
                const Vec3<T> f1 = has_def ? fetch_deform(deform, g, xp + 1, yp, zp) : zero;
                edge_adjoint<T, 0>(dme, d1, iso, xp, yp, zp, has_def, fme, f1, load_g(r.start), 0, acc_d, acc_f);
            }
            if (oy) {
                const T d1 = fetch_padded(sdf, g, xp, yp + 1, zp, padv);
                const Vec3<T> f1 = has_def ? fetch_deform(deform, g, xp, yp + 1, zp) : zero;
                edge_adjoint<T, 1>(dme, d1, iso, xp, yp, zp, has_def, fme, f1, load_g(r.start + r.bx), 0, acc_d, acc_f);
            }
            if (oz) {
                const Vec3<T> f1 = has_def ? fetch_deform(deform, g, xp, yp, zp + 1) : zero;
                edge_adjoint<T, 2>(dme, d_zp, iso, xp, yp, zp, has_def, fme, f1, load_g(r.start + r.bx + r.by), 0, acc_d, acc_f);
            }
            if (ix) {
                const T d0 = fetch_padded(sdf, g, xp - 1, yp, zp, padv);
                const Vec3<T> f0 = has_def ? fetch_deform(deform, g, xp - 1, yp, zp) : zero;
                const RowRank rr = row_rank(ex, lane);
                edge_adjoint<T, 0>(d0, dme, iso, xp - 1, yp, zp, has_def, f0, fme, load_g(rr.start), 1, acc_d, acc_f);
            }
            if (iy) {
                const T d0 = fetch_padded(sdf, g, xp, yp - 1, zp, padv);
                const Vec3<T> f0 = has_def ? fetch_deform(deform, g, xp, yp - 1, zp) : zero;
                const RowRank rr = row_rank(ey, lane);
                edge_adjoint<T, 1>(d0, dme, iso, xp, yp - 1, zp, has_def, f0, fme, load_g(rr.start + rr.bx), 1, acc_d, acc_f);
            }
            if (iz) {
                const Vec3<T> f0 = has_def ? fetch_deform(deform, g, xp, yp, zp - 1) : zero;
                // id of the +z edge owned by point zp-1
                unsigned id;
                if (lane > 0) { const RowRank rr = row_rank(e, lane - 1); id = rr.start + rr.bx + rr.by; }
                else          { const RowRank rr = row_rank(ezp, 31);     id = rr.start + rr.bx + rr.by; }
                edge_adjoint<T, 2>(d_zm, dme, iso, xp, yp, zp - 1, has_def, f0, fme, load_g(id), 1, acc_d, acc_f);
            }
        }
    }
    if (real) {
        const size_t o = ((size_t)(xp - 1) * g.Y + (yp - 1)) * g.Z + (zp - 1);
        st_stream(adj_sdf + o, acc_d);
        if (adj_deform) {
            st_stream(adj_deform + 3 * o, acc_f.x);
            st_stream(adj_deform + 3 * o + 1, acc_f.y);
            st_stream(adj_deform + 3 * o + 2, acc_f.z);
        }
    }
}

}  // namespace diso
