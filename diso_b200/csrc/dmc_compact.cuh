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
#include "quad_split.cuh"
#include "tables.cuh"

namespace diso {

// MODE 0: quads.  MODE 1: exact adjoint.  MODE 2: reference-compatible adjoint (every patch of a
// cell reads the adjoint of the cell's FIRST dual vertex, cudualmc.cu:975,990).
// OFFSET (MODE 0 only): add id_offset to every index (slab -> global ids).
// DIAG (MODE 0 only): also decide the quad's diagonal for the quad -> triangle split (diso/__init__.py:118-147) while its
// four dual-vertex ids are in registers: gathers the four (final, API-frame) vertices written by dmc_dual_verts and
// stores one flag byte per quad; DiffDMC's default path then needs only a scan over those bytes (quad_diag<PRECOMPUTED>)
// instead of re-reading the 32-byte quads and chasing ids -> vertices (2.15 ms at 512^3, latency-bound).
template <typename T, bool LISTED> struct QuadSmem {
    unsigned short list[CT_MAX_EDGES];
    unsigned long long edge5[256];   // T_DMC_EDGE5: per cell edge {length of its patch : 3, index of the patch in the cell : 2}
    unsigned quad[8];
    T inv[8];
    int k[LISTED ? CT_CHUNKS : 1];
};

template <typename T, int MODE, bool LISTED, bool OFFSET = false, bool DIAG = false>
__device__ __forceinline__ void dmc_edges2_tile(QuadSmem<T, LISTED> &sm, int tile, const Geo &g, const unsigned *__restrict__ S,
                                                const uint4 *__restrict__ E, const uint4 *__restrict__ P,
                                                const unsigned short *__restrict__ C,
                                                const unsigned *__restrict__ alist, int n_active, T ix, T iy, T iz,
                                                const T *__restrict__ adj_dual, long long id_offset,
                                                long long *__restrict__ quads, T *__restrict__ gedge, int gedge_soa,
                                                const T *__restrict__ verts, unsigned char *__restrict__ qflags,
                                                T *__restrict__ rec)
{
    // gedge_soa: MODE 1/2 = layout of gedge (1: blocked SoA); MODE 0 = components per saved edge record (6 with deform, 3 without:
    // the meta word is the last one)
    unsigned short *s_list = sm.list;
    unsigned long long *s_edge5 = sm.edge5;
    unsigned *s_quad = sm.quad;
    T *s_inv = sm.inv;
    int *s_k = sm.k;
    s_edge5[threadIdx.x] = T_DMC_EDGE5[threadIdx.x];
    if (threadIdx.x < 6) s_quad[threadIdx.x] = T_DMC_QUAD[threadIdx.x];
    if (threadIdx.x < 8) s_inv[threadIdx.x] = threadIdx.x ? T(1) / T((int)threadIdx.x) : T(0);
    const TileRange<LISTED> tr(alist, n_active, tile);
    unsigned tile_base;
    const unsigned n = build_edge_list<LISTED, true>(g, E, tr, S, s_list, nullptr, s_k, tile_base);
    if (n == 0) return;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const int axis = d & 3, j = (d >> 2) & 31, cl = (d >> 7) & 63, inside = (d >> 13) & 1;
        const int k = LISTED ? s_k[cl] : tr.e0 + cl;
        // reference dmcQuad[type], type = (exiting ? 3 : 0) + axis: per corner, the cell at (-dx,-dy,-dz) of the
        // edge's start point (bits 0..2) and the id of the edge inside that cell (bits 4..7).  Decoded with
        // ALU ops on purpose: these kernels are bound by L1 / shared-memory wavefronts (a per-corner table in
        // shared memory measured slower)
        const unsigned q4 = s_quad[inside * 3 + axis];
        long long id[4];
        unsigned meta = 0;   // MODE 0: per corner {patch length : 3, index of the patch inside its cell : 2}
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
            const unsigned f5 = (unsigned)(s_edge5[code] >> (5 * eid)) & 31u;   // {patch length : 3, patch index : 2}
            const unsigned ord = f5 >> 3;
            if (MODE == 0) {
                id[c] = (long long)(first + ord);
                meta |= f5 << (5 * c);
            } else {
                const unsigned src = (MODE == 1) ? first + ord : first;
                const T inv = s_inv[f5 & 7u];
                const T *p = adj_dual + (size_t)src * 3;
                acc.x = fma_rn(__ldg(p), inv, acc.x);
                acc.y = fma_rn(__ldg(p + 1), inv, acc.y);
                acc.z = fma_rn(__ldg(p + 2), inv, acc.z);
            }
        }
        const size_t rank = (size_t)tile_base + i;
        if (MODE == 0) {
            // saved for the backward (6th record component, mc_backward_v2.cuh): with the quad's four ids the adjoint of
            // the dual-vertex averaging needs no cell / patch lookups at all
            if (rec) reinterpret_cast<unsigned *>(rec + (rank >> 5) * (size_t)(32 * gedge_soa) + 32 * (gedge_soa - 1) + (rank & 31))[0] = meta;
            if (DIAG) {   // local ids index the local vertex array (the slab offset is added below)
                Vec3<T> v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const T *pv = verts + id[c] * 3;
                    v[c] = Vec3<T>{__ldg(pv), __ldg(pv + 1), __ldg(pv + 2)};
                }
                qflags[rank] = quad_first_diagonal(v[0], v[1], v[2], v[3]) ? 1 : 0;
            }
            if (OFFSET) { id[0] += id_offset; id[1] += id_offset; id[2] += id_offset; id[3] += id_offset; }
            longlong2 *dst = reinterpret_cast<longlong2 *>(quads + rank * 4);
            __stcs(dst, make_longlong2(id[0], id[1]));
            __stcs(dst + 1, make_longlong2(id[2], id[3]));
        } else {
            if (gedge_soa) {   // blocked SoA (mc_backward_v2.cuh:blk_index): coalesced here and in stage B
                T *dst = gedge + (rank >> 5) * 96 + (rank & 31);
                dst[0] = acc.x * ix; dst[32] = acc.y * iy; dst[64] = acc.z * iz;
            } else {
                T *dst = gedge + rank * 3;
                dst[0] = acc.x * ix; dst[1] = acc.y * iy; dst[2] = acc.z * iz;
            }
        }
    }
}

template <typename T, int MODE, bool LISTED, bool OFFSET = false, bool DIAG = false>
__global__ void __launch_bounds__(CT_THREADS) dmc_edges2_kernel(Geo g, const unsigned *__restrict__ S,
                                                              const uint4 *__restrict__ E, const uint4 *__restrict__ P,
                                                              const unsigned short *__restrict__ C,
                                                              const unsigned *__restrict__ alist, int n_active, T ix, T iy, T iz,
                                                              const T *__restrict__ adj_dual, long long id_offset,
                                                              long long *__restrict__ quads, T *__restrict__ gedge, int gedge_soa,
                                                              const T *__restrict__ verts = nullptr, unsigned char *__restrict__ qflags = nullptr,
                                                              T *__restrict__ rec = nullptr)
{
    __shared__ QuadSmem<T, LISTED> sm;
    dmc_edges2_tile<T, MODE, LISTED, OFFSET, DIAG>(sm, blockIdx.x, g, S, E, P, C, alist, n_active, ix, iy, iz, adj_dual, id_offset, quads, gedge,
                                                   gedge_soa, verts, qflags, rec);
}

// Edge crossings + quads in one launch (dense tile flavour, return_quads path): even CTAs evaluate the crossings of tile b/2
// (DRAM-bound: 4.5 GB per launch), odd CTAs emit the quads of the same tile (LSU-bound); see mc_emit_fused_kernel.
template <typename T, bool OFFSET>
__global__ void __launch_bounds__(CT_THREADS) dmc_emit_fused_kernel(const T *__restrict__ sdf, const T *__restrict__ deform,
                                                                  Geo g, T iso, T padv, EpilogueC<T> raw, const unsigned *__restrict__ S,
                                                                  const uint4 *__restrict__ E, const uint4 *__restrict__ P,
                                                                  const unsigned short *__restrict__ C, long long id_offset,
                                                                  T *__restrict__ scratch, T *__restrict__ rec,
                                                                  long long *__restrict__ quads)
{
    __shared__ union U { EvSmem ev; QuadSmem<T, false> quad; __device__ U() {} } sm;
    const int tile = blockIdx.x >> 1;
    if (blockIdx.x & 1)
        dmc_edges2_tile<T, 0, false, OFFSET, false>(sm.quad, tile, g, S, E, P, C, nullptr, g.NCH, T(1), T(1), T(1), nullptr, id_offset, quads, nullptr,
                                                   deform ? 6 : 3, nullptr, nullptr, rec);
    else
        edge_verts_tile<T, false>(sm.ev, tile, sdf, deform, g, iso, padv, raw, E, nullptr, g.NCH, scratch, rec, deform ? 6 : 3);
}

}  // namespace diso

namespace diso {

constexpr int CT_MAX_PATCHES = CT_CHUNKS * 128;

// ------------------------------------------------------------------------------------------
// K3d (v3): dual vertices, patch-parallel.  Replaces create_dmc_verts_kernel
// (cudualmc.cu:907-955) + epilogue (diso/__init__.py:110-114).
// The reference recomputes every edge crossing on the fly in each of the 4 cells around the
// edge (cudualmc.cu:946-948).  Here the crossings are evaluated ONCE by edge_verts_kernel in
// the raw padded frame into `mcv` (caller scratch, [n_edges,3]) and this kernel only averages:
//   phase A  thread == 8 cells of a chunk: the per-cell words written by classify_scan ((possibly
//            complemented) case index | offset of the cell's first dual vertex), one 128-bit load
//            -> one descriptor per patch in the shared list at slot (dual vertex id - first id of
//            the tile); the patch counts come from the record's bit planes (no table lookup).
//   phase B  thread == dual vertex: gather the crossings of the patch's member edges by rank in
//            ascending edge id (== the reference's table order, asserted in
//            tools/extract_tables.py), sum, scale by 1/len, apply the epilogue, store at
//            consecutive output ranks.  Same values in the same order => bit-identical.
// ------------------------------------------------------------------------------------------
template <typename T, bool LISTED>
// fp32: 32 registers for 8 CTAs/SM (34 registers / 7 CTAs by default: 1.18 -> 1.11 ms at 512^3, no spills); fp64: 64 registers
// (the compiler's own choice without a bound; a bound of 1-2 CTAs lets it take 110 registers: 1.74 -> 2.35 ms)
__global__ void __launch_bounds__(CT_THREADS, sizeof(T) == 4 ? 8 : 4) dmc_dual_verts_kernel(const T *__restrict__ mcv, Geo g, EpilogueC<T> epi,
                                                                  const uint4 *__restrict__ E,
                                                                  const uint4 *__restrict__ P,
                                                                  const unsigned short *__restrict__ C,
                                                                  const unsigned *__restrict__ alist, int n_active,
                                                                  T *__restrict__ verts)
{
    __shared__ unsigned short s_list[CT_MAX_PATCHES];
    __shared__ __align__(16) unsigned short s_cell[CT_CHUNKS * 32];
    __shared__ unsigned long long s_members[256];
    __shared__ uint4 s_E[CT_RECS];
    __shared__ T s_inv[8];
    __shared__ int s_k[LISTED ? CT_CHUNKS : 1];
    const TileRange<LISTED> tr(alist, n_active, blockIdx.x);
    const unsigned tile_base = P[tr.kfirst].x;
    const unsigned n = P[tr.klast + 1].x - tile_base;
    if (n == 0) return;
    s_members[threadIdx.x] = T_DMC_MEMBERS[threadIdx.x];
    if (threadIdx.x < 8) s_inv[threadIdx.x] = threadIdx.x ? T(1) / T((int)threadIdx.x) : T(0);
    if (LISTED) {
        if ((int)threadIdx.x < tr.count) s_k[threadIdx.x] = tr.chunk(threadIdx.x);
    }
    __syncthreads();
    const RecCache<LISTED> rc = load_record_cache<LISTED>(g, E, tr, s_k, s_E);

    // ---- phase A ---------------------------------------------------------------------------------
    {
        const int cl = threadIdx.x >> 2, q4 = threadIdx.x & 3;
        if (cl < tr.count) {
            const int k = LISTED ? s_k[cl] : tr.e0 + cl;
            const uint4 pr = __ldg(P + k);
            const unsigned ub = (pr.y >> (8 * q4)) & 0xffu;
            if (ub) {
                const uint4 cw = __ldg(reinterpret_cast<const uint4 *>(C + (size_t)k * 32 + 8 * q4));
                *reinterpret_cast<uint4 *>(s_cell + cl * 32 + 8 * q4) = cw;
                const unsigned w[4] = {cw.x, cw.y, cw.z, cw.w};
                const unsigned lo = pr.z >> (8 * q4), hi = pr.w >> (8 * q4);
                const unsigned pb = pr.x - tile_base;
                const unsigned dbase = ((unsigned)cl << 7) | ((unsigned)q4 << 5);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if ((ub >> j) & 1u) {
                        const unsigned info = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                        const unsigned np = 1u + ((lo >> j) & 1u) + 2u * ((hi >> j) & 1u);
                        const unsigned slot = pb + (info >> 8);
                        const unsigned dd = dbase | ((unsigned)j << 2);
#pragma unroll
                        for (unsigned t = 0; t < 4; ++t)
                            if (t < np) s_list[slot + t] = (unsigned short)(dd | t);
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B ---------------------------------------------------------------------------------
    // Vertex ids of all 12 cell edges from 4 records (one per row of the cell): id = B[row][dz] (+ the
    // lower-axis bits of the owning point), with B[row][1] = B[row][0] + the three bits of point j.  The
    // member edges are then visited in ascending edge id with compile-time row / axis selectors
    // (the table-driven loop with a per-edge record decode cost ~2x the instructions, ncu r1_v5).
    const bool fast = !LISTED || rc.contig;
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const unsigned q = d & 3u;
        const int j = (d >> 2) & 31, cl = (d >> 7) & 63;
        const unsigned code = s_cell[cl * 32 + j] & 0xffu;
        const unsigned long long mw = s_members[code];                // four 12-bit member masks + four 4-bit patch lengths
        const unsigned plen = (unsigned)(mw >> 48);
        const unsigned mm = (unsigned)(mw >> (12 * q)) & 0xfffu;      // member edges of this patch
        Vec3<T> acc{T(0), T(0), T(0)};
        if (fast) {
            const unsigned l = lanemask_lt(j);
            unsigned B0[4], B1[4], yb[4], zb[4];
#pragma unroll
            for (int rs = 0; rs < 4; ++rs) {
                const uint4 rec = s_E[rs * (CT_CHUNKS + 1) + cl];
                B0[rs] = rec.x + __popc(rec.y & l) + __popc(rec.z & l) + __popc(rec.w & l);
                yb[rs] = (rec.y >> j) & 1u;
                zb[rs] = (rec.z >> j) & 1u;
                B1[rs] = B0[rs] + yb[rs] + zb[rs] + ((rec.w >> j) & 1u);
            }
            // y-bit of point j+1 in rows 0 and 2 (edges 11 and 10): lane 0 of the next chunk when j == 31
            unsigned y01, y21;
            {
                const int nx = (j + 1) >> 5, jj = (j + 1) & 31;
                y01 = (s_E[cl + nx].y >> jj) & 1u;
                y21 = (s_E[2 * (CT_CHUNKS + 1) + cl + nx].y >> jj) & 1u;
            }
            unsigned r[12];
            r[0] = B0[0];                 r[1] = B0[2] + yb[2] + zb[2];  r[2] = B1[0];
            r[3] = B0[0] + yb[0] + zb[0]; r[4] = B0[1];                  r[5] = B0[3] + yb[3] + zb[3];
            r[6] = B1[1];                 r[7] = B0[1] + yb[1] + zb[1];  r[8] = B0[0] + yb[0];
            r[9] = B0[2] + yb[2];         r[10] = B1[2] + y21;           r[11] = B1[0] + y01;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                T vx[6], vy[6], vz[6];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int e = 6 * h + t;
                    vx[t] = vy[t] = vz[t] = T(0);
                    if ((mm >> e) & 1u) {
                        const T *pv = mcv + (size_t)r[e] * 3;
                        vx[t] = __ldg(pv); vy[t] = __ldg(pv + 1); vz[t] = __ldg(pv + 2);
                    }
                }
                // unconditional: a non-member slot holds +0, and acc (started at +0, never -0) + 0 is acc bit for bit
#pragma unroll
                for (int t = 0; t < 6; ++t) { acc.x = acc.x + vx[t]; acc.y = acc.y + vy[t]; acc.z = acc.z + vz[t]; }
            }
        } else {
            // scattered tile of a sparse surface: the generic per-edge decode (next-chunk records from global)
            for (unsigned m2 = mm; m2; m2 &= m2 - 1) {
                const int e = __ffs(m2) - 1;
                const T *pv = mcv + (size_t)edge_rank<LISTED>(rc, cl, j, e) * 3;
                acc.x = acc.x + __ldg(pv); acc.y = acc.y + __ldg(pv + 1); acc.z = acc.z + __ldg(pv + 2);
            }
        }
        const T inv = s_inv[(plen >> (4 * q)) & 7u];
        Vec3<T> v{acc.x * inv, acc.y * inv, acc.z * inv};
        v = epi.apply(v);
        T *dst = verts + (size_t)(tile_base + i) * 3;
        st_stream(dst, v.x); st_stream(dst + 1, v.y); st_stream(dst + 2, v.z);
    }
}

}  // namespace diso
