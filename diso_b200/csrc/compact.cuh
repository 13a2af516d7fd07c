// compact.cuh -- CTA-level work compaction shared by the emit kernels.
//
// Only ~18 % of the (grid point, axis) slots of a random-init SDF carry a crossing edge (and
// far fewer on smooth surfaces), so a lane-per-point kernel idles most of its lanes: the ncu
// baseline (profiles/r1_base_summary.md) shows every such kernel issue-bound with ~15 of 32 threads
// active.  Instead each CTA owns a tile of CT_CHUNKS consecutive entries of the ACTIVE-chunk list
// that classify_scan builds in ascending chunk order (inactive chunks own nothing, so the items
// of a tile still occupy one contiguous rank range, and a sparse surface costs time proportional
// to the surface instead of the volume).  A tile is processed in two phases:
//   phase A  thread == one byte (8 points) of a chunk's lane masks: walk the set bits and drop a
//            16-bit descriptor {chunk-in-tile, lane, axis} for every crossing edge into a shared
//            list.  Because the records carry GLOBAL exclusive prefix sums, the list slot of an edge
//            is simply rank - rank_of_first_edge_of_tile: no scan, and list order == output order.
//   phase B  thread == list entry: fully populated warps evaluate one edge each and write the
//            result at consecutive output ranks (dense, coalesced stores).
#pragma once
#include "classify.cuh"
#include "edge_math.cuh"
#include "tables.cuh"

namespace diso {

constexpr int CT_CHUNKS = 64;    // chunks per CTA tile
constexpr int CT_THREADS = 256;  // threads per CTA: four threads per chunk in phase A
constexpr int CT_WARPS = CT_THREADS / 32;
constexpr int CT_MAX_EDGES = CT_CHUNKS * 96;
static_assert(CT_THREADS == 4 * CT_CHUNKS, "phase A maps four threads (one byte of the lane masks each) to a chunk");

// Tiles come in two flavours, selected on the host from the active-chunk counts:
//   LISTED = false  "dense":  tile b = chunks [64 b, 64 b + 64); no indirection, the +1 neighbour of entry
//                             cl is entry cl + 1.  Used when most chunks are active (random fields).
//   LISTED = true   "sparse": tile b = entries [64 b, 64 b + 64) of the ordered active-chunk list.
template <bool LISTED> struct TileRange {
    int e0, count, kfirst, klast;
    const unsigned *alist;
    __device__ __forceinline__ TileRange(const unsigned *__restrict__ al, int n_active, int tile) : alist(al)
    {
        e0 = tile * CT_CHUNKS;
        count = min(CT_CHUNKS, n_active - e0);
        kfirst = LISTED ? (int)__ldg(al + e0) : e0;
        klast = LISTED ? (int)__ldg(al + e0 + count - 1) : e0 + count - 1;
    }
    __device__ __forceinline__ int chunk(int cl) const { return LISTED ? (int)__ldg(alist + e0 + cl) : e0 + cl; }
};

// Position of a chunk for the value fetches of phase B: 32-bit index of the element "lane 0" of the
// chunk in the UNPADDED arrays (may lie outside: validity comes from the flags and the z range).
struct TilePos {
    int rowbase;     // ((xp-1) Y + (yp-1)) Z + 32 c - 1
    int xp, yp;
    unsigned zf;     // 32 c (28 bits; PZ < 2^28 by the per-call index budget) | flags << 28
                     // flags bit0: 1 <= xp <= X   bit1: 1 <= yp <= Y   bit2: xp + 1 <= X   bit3: yp + 1 <= Y
};

__device__ __forceinline__ TilePos make_tile_pos(const Geo &g, int k)
{
    const int r = k / g.NC, c = k - r * g.NC;
    const int xp = r / g.PY, yp = r - xp * g.PY;
    TilePos tp;
    tp.rowbase = ((xp - 1) * g.Y + (yp - 1)) * g.Z + 32 * c - 1;
    tp.xp = xp; tp.yp = yp;
    const unsigned flags = (unsigned)(xp >= 1 && xp <= g.X) | ((unsigned)(yp >= 1 && yp <= g.Y) << 1) |
                           ((unsigned)(xp + 1 <= g.X) << 2) | ((unsigned)(yp + 1 <= g.Y) << 3);
    tp.zf = (unsigned)(32 * c) | (flags << 28);
    return tp;
}

// Phase A for edge lists: thread t owns byte (t & 3) of the three crossing masks of tile entry t >> 2
// and walks its set bits serially (a few per thread), so list building costs ~15 warp instructions
// per chunk instead of the ~55 of a lane-per-point formulation (ncu, profiles/r1_v4_summary.md vs r1_final_summary.md).
// Fills s_list[rank - tile_base], s_pos / s_k (when given); returns the number of edges of the tile
// (uniform).  Must be called by all CT_THREADS threads; the caller synchronises afterwards.
// SIGN: bit 13 of each descriptor tells whether the edge's start point is inside (value >= iso), i.e.
// whether the crossing is "exiting" in the DMC sense (cudualmc.cu:782-788).
template <bool LISTED, bool SIGN>
__device__ __forceinline__ unsigned build_edge_list(const Geo &g, const uint4 *__restrict__ E, const TileRange<LISTED> &tr,
                                                    const unsigned *__restrict__ S, unsigned short *s_list, TilePos *s_pos,
                                                    int *s_k, unsigned &tile_base)
{
    tile_base = E[tr.kfirst].x;
    const unsigned n = E[tr.klast + 1].x - tile_base;  // E[k+1].base = E[k].base + edges of chunk k
    if (n == 0) return 0;
    const int cl = threadIdx.x >> 2, sh = (threadIdx.x & 3) * 8;
    if (cl < tr.count) {
        const int k = tr.chunk(cl);
        const uint4 rec = __ldg(E + k);
        if (sh == 0) {
            if (LISTED && s_k) s_k[cl] = k;
            if (s_pos) s_pos[cl] = make_tile_pos(g, k);
        }
        const unsigned bx = (rec.y >> sh) & 0xffu, by = (rec.z >> sh) & 0xffu, bz = (rec.w >> sh) & 0xffu;
        unsigned m = bx | by | bz;
        if (m) {
            const unsigned lt = (1u << sh) - 1u;
            unsigned slot = rec.x - tile_base + __popc(rec.y & lt) + __popc(rec.z & lt) + __popc(rec.w & lt);
            const unsigned sb = SIGN ? (__ldg(S + k) >> sh) & 0xffu : 0u;
            const unsigned dbase = ((unsigned)cl << 7) | ((unsigned)sh << 2);
            do {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                unsigned d = dbase | ((unsigned)j << 2);
                if (SIGN) d |= ((sb >> j) & 1u) << 13;
                if ((bx >> j) & 1u) s_list[slot++] = (unsigned short)d;
                if ((by >> j) & 1u) s_list[slot++] = (unsigned short)(d | 1u);
                if ((bz >> j) & 1u) s_list[slot++] = (unsigned short)(d | 2u);
            } while (m);
        }
    }
    return n;
}

// IEEE-correct x / d for a divisor d whose correctly rounded reciprocal r = RN(1/d) is known
// (Markstein's sequence: q = RN(x r); rem = x - q d exactly by FMA; result = RN(q + rem r)).
// Correctly rounded for normal operands unless d's significand is all ones; our divisors are
// small integers (dims - 1).  Verified bit-for-bit against the division in tests.
template <typename T> __device__ __forceinline__ T div_by_const(T x, T d, T r)
{
    const T q = x * r;
    const T rem = fma_rn(-q, d, x);
    return fma_rn(rem, r, q);
}

template <typename T> struct EpilogueC {
    T dx, dy, dz;  // (T)dim - 1
    T rx, ry, rz;  // RN(1 / (dim - 1))
    int normalize; // 0: raw padded frame (no shift), 1: (p-1), 2: (p-1)/(dim-1) via div_by_const,
                   // 3: (p-1)/(dim-1) by plain division (some dim == 1, i.e. a zero divisor)
    int x0;        // slab frame: global index of the local grid's first x layer (added to the integer
                   // x coordinate BEFORE the deformation, so a slab reproduces the whole grid's bits)
    __device__ __forceinline__ Vec3<T> apply(Vec3<T> p) const
    {
        if (normalize == 0) return p;
        p.x = p.x - T(1); p.y = p.y - T(1); p.z = p.z - T(1);
        if (normalize == 2) {
            p.x = div_by_const(p.x, dx, rx); p.y = div_by_const(p.y, dy, ry); p.z = div_by_const(p.z, dz, rz);
        } else if (normalize == 3) {
            p.x = p.x / dx; p.y = p.y / dy; p.z = p.z / dz;
        }
        return p;
    }
};

// ------------------------------------------------------------------------------------------
// K3 (v2): edge vertices, edge-parallel.  Replaces create_cell_mc_verts_kernel
// (cumc.cu:370-410) + the "-1"/normalise epilogue (diso/__init__.py:56-60).  With
// epi.normalize == 0 it produces the raw padded-frame crossings that the DMC dual-vertex
// kernel averages (computeMcVert of cudualmc.cu:683-708, each edge evaluated once, not 4x).
// ------------------------------------------------------------------------------------------
// The tile bodies are device functions over an explicit shared-memory struct so that two of them can share one launch
// (mc_emit_fused_kernel / dmc_emit_fused_kernel below: CTAs of a DRAM-bound and of an LSU-bound pass co-resident on every SM).
struct EvSmem {
    unsigned short list[CT_MAX_EDGES];
    TilePos pos[CT_CHUNKS];
};

template <typename T, bool LISTED>
__device__ __forceinline__ void edge_verts_tile(EvSmem &sm, int tile, const T *__restrict__ sdf, const T *__restrict__ deform,
                                                const Geo &g, T iso, T padv, const EpilogueC<T> &epi,
                                                const uint4 *__restrict__ E,
                                                const unsigned *__restrict__ alist, int n_active,
                                                T *__restrict__ verts, T *__restrict__ rec, int rec_ncomp)
{
    unsigned short *s_list = sm.list;
    TilePos *s_pos = sm.pos;
    const TileRange<LISTED> tr(alist, n_active, tile);
    unsigned tile_base;
    const unsigned n = build_edge_list<LISTED, false>(g, E, tr, nullptr, s_list, s_pos, nullptr, tile_base);
    if (n == 0) return;
    __syncthreads();
    const bool has_def = deform != nullptr;
    const int sZ = g.Z, sYZ = g.Y * g.Z;
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const int axis = d & 3, j = (d >> 2) & 31;
        const TilePos tp = s_pos[(d >> 7) & 63];
        const unsigned flags = tp.zf >> 28;
        const int zp = (int)(tp.zf & 0x0fffffffu) + j, zq = zp + (axis == 2);
        const unsigned need1 = axis == 0 ? 6u : (axis == 1 ? 9u : 3u);
        const bool v0 = (flags & 3u) == 3u && (unsigned)(zp - 1) < (unsigned)g.Z;
        const bool v1 = (flags & need1) == need1 && (unsigned)(zq - 1) < (unsigned)g.Z;
        const int i0 = tp.rowbase + j;
        const int i1 = i0 + (axis == 0 ? sYZ : (axis == 1 ? sZ : 1));
        const T d0 = v0 ? __ldg(sdf + i0) : padv;
        const T d1 = v1 ? __ldg(sdf + i1) : padv;
        const T t = edge_t(d0, d1, iso);
        const int xg = tp.xp + epi.x0;
        Vec3<T> p0{T(xg), T(tp.yp), T(zp)};
        Vec3<T> p1{T(xg + (axis == 0)), T(tp.yp + (axis == 1)), T(zq)};
        if (has_def) {
            // the pad layer carries zero deformation (diso/__init__.py:54)
            if (v0) {
                const T *f = deform + (size_t)i0 * 3;
                p0.x = p0.x + __ldg(f); p0.y = p0.y + __ldg(f + 1); p0.z = p0.z + __ldg(f + 2);
            }
            if (v1) {
                const T *f = deform + (size_t)i1 * 3;
                p1.x = p1.x + __ldg(f); p1.y = p1.y + __ldg(f + 1); p1.z = p1.z + __ldg(f + 2);
            }
        }
        const Vec3<T> dp{p1.x - p0.x, p1.y - p0.y, p1.z - p0.z};
        Vec3<T> p;
        p.x = fma_rn(dp.x, t, p0.x);
        p.y = fma_rn(dp.y, t, p0.y);
        p.z = fma_rn(dp.z, t, p0.z);
        p = epi.apply(p);
        const size_t rank = (size_t)tile_base + i;
        T *dst = verts + rank * 3;
        st_stream(dst, p.x); st_stream(dst + 1, p.y); st_stream(dst + 2, p.z);
        if (rec) {
            // saved for the backward (mc_backward_v2.cuh): everything adjComputeMcVert (cumc.cu:412-453) needs of this
            // edge, as five arrays indexed by rank -- coalesced here and there, no sdf / deform gathers in the backward
            // (groups of 32 edges, component-major inside a group: mc_backward_v2.cuh:blk_index)
            // rec_ncomp = (deform ? 5 : 2) + (DMC: quad meta ? 1 : 0); without deform p1 - p0 is the edge's unit axis vector
            // and only {d0, d1} are kept
            T *r = rec + (rank >> 5) * (size_t)(32 * rec_ncomp) + (rank & 31);
            // streaming stores: the records are not read again before the backward (no measurable difference to plain stores)
            if (has_def) { st_stream(r, dp.x); st_stream(r + 32, dp.y); st_stream(r + 64, dp.z); r += 96; }
            st_stream(r, d0); st_stream(r + 32, d1);
        }
    }
}

template <typename T, bool LISTED>
__global__ void __launch_bounds__(CT_THREADS) edge_verts_kernel(const T *__restrict__ sdf, const T *__restrict__ deform,
                                                              Geo g, T iso, T padv, EpilogueC<T> epi,
                                                              const uint4 *__restrict__ E,
                                                              const unsigned *__restrict__ alist, int n_active,
                                                              T *__restrict__ verts, T *__restrict__ rec, int rec_ncomp)
{
    __shared__ EvSmem sm;
    edge_verts_tile<T, LISTED>(sm, blockIdx.x, sdf, deform, g, iso, padv, epi, E, alist, n_active, verts, rec, rec_ncomp);
}

// ------------------------------------------------------------------------------------------
// Edge-record cache: the records of the four rows a cell's 12 edges are owned by
// (rowset = 2*dx + dy of mcEdgeLocations, cumc.cu:109-122), for every chunk of the tile and the
// chunk following it (dz = 1 from lane 31).  Filled with 16-byte loads in phase A so that
// phase B turns "edge id -> vertex id" into one LDS.128 + three popcounts -- the reference does
// an owner-cell lookup, two offset loads and a linear search per index (cumc.cu:589-607).
// ------------------------------------------------------------------------------------------
// Layout [rowset][CT_CHUNKS + 1].  In a contiguous tile (dense surfaces: chunks k0 .. k0+count-1) the
// "next chunk" of entry cl is simply entry cl + 1 (one extra record per rowset).  In a scattered tile
// (sparse surfaces) the next chunk is not in the cache; that lookup (only lane 31 with dz = 1) goes to
// global memory instead -- keeping the cache at 4 KB matters more for occupancy than the rare load.
constexpr int CT_RECS = 4 * (CT_CHUNKS + 1);

template <bool LISTED> struct RecCache { const uint4 *s_E; const uint4 *E; const int *s_k; int sX, sY; bool contig; };

// s_k (LISTED only) must already hold the tile's chunk ids.
template <bool LISTED>
__device__ __forceinline__ RecCache<LISTED> load_record_cache(const Geo &g, const uint4 *__restrict__ E, const TileRange<LISTED> &tr,
                                                              const int *s_k, uint4 *s_E)
{
    const bool contig = !LISTED || (tr.klast - tr.kfirst == tr.count - 1);
    const int per = contig ? tr.count + 1 : tr.count;
    for (int i = threadIdx.x; i < 4 * per; i += CT_THREADS) {
        const int rs = i / per, cl = i - rs * per;
        const int kk = (contig ? tr.kfirst + cl : s_k[cl]) + (rs >> 1) * g.sX + (rs & 1) * g.sY;   // E has a zero-filled tail
        s_E[rs * (CT_CHUNKS + 1) + cl] = __ldg(E + kk);
    }
    return RecCache<LISTED>{s_E, E, s_k, g.sX, g.sY, contig};
}

// vertex id of local edge e of cell (tile entry cl, lane j)
template <bool LISTED>
__device__ __forceinline__ unsigned edge_rank(const RecCache<LISTED> &rc, int cl, int j, int e)
{
    const int ax = (EDGE_AX >> (2 * e)) & 3;
    const int rs = (((EDGE_DX >> e) & 1) << 1) | ((EDGE_DY >> e) & 1);
    int jj = j + ((EDGE_DZ >> e) & 1);
    const int nxt = jj >> 5;     // dz = 1 from lane 31: first point of the following chunk
    jj &= 31;
    uint4 rec;
    if (LISTED && nxt && !rc.contig) rec = __ldg(rc.E + rc.s_k[cl] + 1 + (rs >> 1) * rc.sX + (rs & 1) * rc.sY);
    else rec = rc.s_E[rs * (CT_CHUNKS + 1) + cl + nxt];
    const unsigned l = lanemask_lt(jj);
    unsigned r = rec.x + __popc(rec.y & l) + __popc(rec.z & l) + __popc(rec.w & l);
    if (ax >= 1) r += bit(rec.y, jj);
    if (ax == 2) r += bit(rec.z, jj);
    return r;
}

// the same for a pre-decoded corner code f = {row set : 2, dz : 1, axis >= 1 : 1, axis == 2 : 1} (T_MC_TRI5, tools/extract_tables.py):
// the axis bits extend the "points before lane jj" mask by lane jj itself for the x- (and y-) crossing planes, so the rank is
// base + three popcounts with no per-axis selects.  The triangle kernel is bound by the integer ALU pipe (83 % busy,
// profiles/r2_emit_pipes.md); this form needs ~25 instead of ~36 instructions per corner.
template <bool LISTED>
__device__ __forceinline__ unsigned corner_rank(const RecCache<LISTED> &rc, int cl, int j, unsigned f)
{
    const int rs = f & 3u;
    int jj = j + (int)((f >> 2) & 1u);
    const int nxt = jj >> 5;     // dz = 1 from lane 31: first point of the following chunk
    jj &= 31;
    uint4 rec;
    if (LISTED && nxt && !rc.contig) rec = __ldg(rc.E + rc.s_k[cl] + 1 + (rs >> 1) * rc.sX + (rs & 1) * rc.sY);
    else rec = rc.s_E[rs * (CT_CHUNKS + 1) + cl + nxt];
    const unsigned l = lanemask_lt(jj);
    const unsigned lx = l | (((f >> 3) & 1u) << jj), ly = l | ((f >> 4) << jj);
    return rec.x + __popc(rec.y & lx) + __popc(rec.z & ly) + __popc(rec.w & l);
}

// ------------------------------------------------------------------------------------------
// K4 (v3): triangles, triangle-parallel.  Replaces count_cell_mc_tris / create_cell_mc_tris
// (cumc.cu:540-612) + the int64 widening (diso/__init__.py:61).
//   phase A  thread == 8 cells of a chunk: the per-cell words written by classify_scan (case index |
//            offset of the cell's first triangle), fetched as one 128-bit load -> one descriptor
//            {chunk-in-tile, lane, k-th triangle} per triangle at slot (triangle id - first id of tile).
//   phase B  thread == triangle: three edge ids from the case table -> three vertex ids via the
//            record cache -> 24 contiguous bytes at the triangle's output rank.
// ------------------------------------------------------------------------------------------
constexpr int CT_MAX_TRIS = CT_CHUNKS * 160;

// OFFSET: add id_offset to every index (slab -> global ids); a separate instantiation keeps the 64-bit adds
// out of the standalone kernel (they cost 2.5 % there).
template <bool LISTED> struct TriSmem {
    unsigned tri5[1024];     // T_MC_TRI5: per case {tris 0|1, tris 2|3, tri 4, #tris}
    uint4 recs[CT_RECS];
    unsigned short list[CT_MAX_TRIS];
    __align__(8) unsigned char code[CT_CHUNKS * 32];
    int k[LISTED ? CT_CHUNKS : 1];
};

template <bool LISTED, bool OFFSET>
__device__ __forceinline__ void mc_tris_tile(TriSmem<LISTED> &sm, int tile, const Geo &g, const uint4 *__restrict__ E,
                                             const uint2 *__restrict__ F, const unsigned short *__restrict__ C,
                                             const unsigned *__restrict__ alist, int n_active,
                                             long long id_offset, long long *__restrict__ tris)
{
    unsigned *s_tri = sm.tri5;
    uint4 *s_E = sm.recs;
    unsigned short *s_list = sm.list;
    unsigned char *s_code = sm.code;
    int *s_k = sm.k;
    const TileRange<LISTED> tr(alist, n_active, tile);
    const unsigned tile_base = F[tr.kfirst].x;
    const unsigned n = F[tr.klast + 1].x - tile_base;
    if (n == 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) s_tri[i * CT_THREADS + threadIdx.x] = T_MC_TRI5[i * CT_THREADS + threadIdx.x];
    if (LISTED) {
        if ((int)threadIdx.x < tr.count) s_k[threadIdx.x] = tr.chunk(threadIdx.x);
    }
    __syncthreads();
    const RecCache<LISTED> rc = load_record_cache<LISTED>(g, E, tr, s_k, s_E);
    {
        const int cl = threadIdx.x >> 2, q4 = threadIdx.x & 3;
        if (cl < tr.count) {
            const int k = LISTED ? s_k[cl] : tr.e0 + cl;
            const uint2 f = __ldg(F + k);
            const unsigned ub = (f.y >> (8 * q4)) & 0xffu;
            unsigned codes_lo = 0, codes_hi = 0;
            if (ub) {
                // 8 per-cell words = 16 bytes, 16-byte aligned ((32 k + 8 q) * 2)
                const uint4 cw = __ldg(reinterpret_cast<const uint4 *>(C + (size_t)k * 32 + 8 * q4));
                const unsigned w[4] = {cw.x, cw.y, cw.z, cw.w};
                const unsigned tb = f.x - tile_base;
                const unsigned dbase = ((unsigned)cl << 8) | ((unsigned)q4 << 6);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if ((ub >> j) & 1u) {
                        const unsigned info = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                        const unsigned code = info & 0xffu;
                        if (j < 4) codes_lo |= code << (8 * j); else codes_hi |= code << (8 * (j - 4));
                        const unsigned nt = s_tri[4 * code + 3];
                        const unsigned slot = tb + (info >> 8);
                        const unsigned dd = dbase | ((unsigned)j << 3);
#pragma unroll
                        for (unsigned t = 0; t < 5; ++t)
                            if (t < nt) s_list[slot + t] = (unsigned short)(dd | t);
                    }
                }
            }
            *reinterpret_cast<uint2 *>(s_code + cl * 32 + 8 * q4) = make_uint2(codes_lo, codes_hi);
        }
    }
    __syncthreads();

    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const unsigned q = d & 7u;
        const int j = (d >> 3) & 31, cl = (d >> 8) & 63;
        const unsigned tri = (s_tri[4 * s_code[cl * 32 + j] + (q >> 1)] >> (15 * (q & 1u))) & 0x7fffu;
        long long a = corner_rank<LISTED>(rc, cl, j, tri & 31u);
        long long b = corner_rank<LISTED>(rc, cl, j, (tri >> 5) & 31u);
        long long c = corner_rank<LISTED>(rc, cl, j, tri >> 10);
        if (OFFSET) { a += id_offset; b += id_offset; c += id_offset; }
        long long *dst = tris + (size_t)(tile_base + i) * 3;
        st_stream(dst, a); st_stream(dst + 1, b); st_stream(dst + 2, c);
    }
}

template <bool LISTED, bool OFFSET>
__global__ void __launch_bounds__(CT_THREADS) mc_tris_kernel(Geo g, const uint4 *__restrict__ E, const uint2 *__restrict__ F,
                                                           const unsigned short *__restrict__ C,
                                                           const unsigned *__restrict__ alist, int n_active,
                                                           long long id_offset, long long *__restrict__ tris)
{
    __shared__ TriSmem<LISTED> sm;
    mc_tris_tile<LISTED, OFFSET>(sm, blockIdx.x, g, E, F, C, alist, n_active, id_offset, tris);
}

// ------------------------------------------------------------------------------------------
// K3 + K4 in one launch (dense tile flavour): even CTAs run the edge pass of tile b/2, odd CTAs the triangle pass of the
// same tile.  The edge pass is DRAM-bound (4.5 GB per launch with the saved records), the triangle pass is bound by
// LSU wavefronts / issue slots with DRAM at 37 %: launched back to back each leaves the other's resource idle; co-resident
// on every SM they overlap (0.82 + 0.90 -> see profiles/r2_backward.md).  The two passes share no data and no barrier.
// ------------------------------------------------------------------------------------------
template <typename T, bool OFFSET>
__global__ void __launch_bounds__(CT_THREADS) mc_emit_fused_kernel(const T *__restrict__ sdf, const T *__restrict__ deform,
                                                                 Geo g, T iso, T padv, EpilogueC<T> epi,
                                                                 const uint4 *__restrict__ E, const uint2 *__restrict__ F,
                                                                 const unsigned short *__restrict__ C, long long id_offset,
                                                                 T *__restrict__ verts, T *__restrict__ rec,
                                                                 long long *__restrict__ tris)
{
    __shared__ union U { EvSmem ev; TriSmem<false> tri; __device__ U() {} } sm;
    const int tile = blockIdx.x >> 1;
    if (blockIdx.x & 1) mc_tris_tile<false, OFFSET>(sm.tri, tile, g, E, F, C, nullptr, g.NCH, id_offset, tris);
    else edge_verts_tile<T, false>(sm.ev, tile, sdf, deform, g, iso, padv, epi, E, nullptr, g.NCH, verts, rec, deform ? 5 : 2);
}

}  // namespace diso
