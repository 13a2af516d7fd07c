// compact.cuh -- CTA-level work compaction shared by the emit kernels.
//
// Only ~18 % of the (grid point, axis) slots of a random-init SDF carry a crossing edge (and
// far fewer on smooth surfaces), so a lane-per-point kernel idles most of its lanes: the ncu
// baseline (profiles/r1_base_summary.md) shows every such kernel issue-bound with ~15 of 32 threads
// active.  Instead each CTA owns a tile of CT_CHUNKS consecutive entries of the ACTIVE-chunk list
// that classify_scan builds in ascending chunk order (inactive chunks own nothing, so the items
// of a tile still occupy one contiguous rank range, and a sparse surface costs time proportional
// to the surface instead of the volume).  A tile is processed in two phases:
//   phase A  lane == grid point: decode the chunk's edge record and drop a 16-bit descriptor
//            {chunk-in-tile, lane, axis} for every crossing edge into a shared list.  Because the
//            records carry GLOBAL exclusive prefix sums, the list slot of an edge is simply
//            rank - rank_of_first_edge_of_tile: no scan, and list order == output order.
//   phase B  thread == list entry: fully populated warps evaluate one edge each and write the
//            result at consecutive output ranks (dense, coalesced stores).
#pragma once
#include "classify.cuh"
#include "edge_math.cuh"
#include "tables.cuh"

namespace diso {

constexpr int CT_CHUNKS = 64;    // chunks per CTA tile
constexpr int CT_THREADS = 256;  // threads per CTA
constexpr int CT_WARPS = CT_THREADS / 32;
constexpr int CT_MAX_EDGES = CT_CHUNKS * 96;

struct TilePos { short xp, yp, c, pad; };

__device__ __forceinline__ unsigned short edge_desc(int chunk_local, int lane, int axis)
{
    return (unsigned short)((chunk_local << 7) | (lane << 2) | axis);
}

// Chunk ids of this CTA's tile: entries [64 b, 64 b + 64) of the active list (`alist == nullptr`:
// every chunk).  Returns the number of valid entries; s_k[i] = chunk id.  Contains a barrier.
__device__ __forceinline__ int load_tile_chunks(const unsigned *__restrict__ alist, int n_active, int *s_k)
{
    const int e0 = blockIdx.x * CT_CHUNKS;
    const int count = min(CT_CHUNKS, n_active - e0);
    if (threadIdx.x < CT_CHUNKS)
        s_k[threadIdx.x] = (int)threadIdx.x < count ? (alist ? (int)alist[e0 + threadIdx.x] : e0 + (int)threadIdx.x) : 0;
    __syncthreads();
    return count;
}

// Phase A for edge lists.  Fills s_list[rank - tile_base] and (if s_pos != nullptr) the padded
// coordinates of every chunk of the tile; returns the number of edges of the tile (uniform).
// Must be called by all CT_THREADS threads.
// If S != nullptr, bit 13 of each descriptor tells whether the edge's start point is inside
// (value >= iso), i.e. whether the crossing is "exiting" in the DMC sense (cudualmc.cu:782-788).
__device__ __forceinline__ unsigned build_edge_list(const Geo &g, const uint4 *__restrict__ E, const int *s_k, int count,
                                                    unsigned short *s_list, TilePos *s_pos, unsigned &tile_base,
                                                    const unsigned *__restrict__ S = nullptr)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    tile_base = E[s_k[0]].x;
    const unsigned n = E[s_k[count - 1] + 1].x - tile_base;  // E[k+1].base = E[k].base + edges of chunk k
    if (n == 0) return 0;
    // each warp: entries wid*8 .. +7 ; lanes 0..7 fetch the records, then broadcast
    constexpr int PER_WARP = CT_CHUNKS / CT_WARPS;
    uint4 mine = make_uint4(0, 0, 0, 0);
    const int emine = wid * PER_WARP + lane;
    const int kmine = (lane < PER_WARP && emine < count) ? s_k[emine] : -1;
    if (kmine >= 0) mine = E[kmine];
    unsigned active = __ballot_sync(FULL, (mine.y | mine.z | mine.w) != 0u);
    if (s_pos && kmine >= 0) {
        const int r = kmine / g.NC;
        TilePos tp;
        tp.c = (short)(kmine - r * g.NC);
        tp.xp = (short)(r / g.PY);
        tp.yp = (short)(r - (r / g.PY) * g.PY);
        tp.pad = 0;
        s_pos[emine] = tp;
    }
    const unsigned lt = lanemask_lt(lane);
    while (active) {
        const int i = __ffs(active) - 1;
        active &= active - 1;
        const unsigned base = __shfl_sync(FULL, mine.x, i), mx = __shfl_sync(FULL, mine.y, i);
        const unsigned my = __shfl_sync(FULL, mine.z, i), mz = __shfl_sync(FULL, mine.w, i);
        const int cl = wid * PER_WARP + i;
        unsigned slot = base - tile_base + __popc(mx & lt) + __popc(my & lt) + __popc(mz & lt);
        unsigned short in13 = 0;
        if (S) in13 = (unsigned short)(bit(S[__shfl_sync(FULL, kmine, i)], lane) << 13);
        if (bit(mx, lane)) s_list[slot++] = edge_desc(cl, lane, 0) | in13;
        if (bit(my, lane)) s_list[slot++] = edge_desc(cl, lane, 1) | in13;
        if (bit(mz, lane)) s_list[slot] = edge_desc(cl, lane, 2) | in13;
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
template <typename T>
__global__ void __launch_bounds__(CT_THREADS) edge_verts_kernel(const T *__restrict__ sdf, const T *__restrict__ deform,
                                                              Geo g, T iso, T padv, EpilogueC<T> epi,
                                                              const uint4 *__restrict__ E,
                                                              const unsigned *__restrict__ alist, int n_active,
                                                              T *__restrict__ verts)
{
    __shared__ unsigned short s_list[CT_MAX_EDGES];
    __shared__ TilePos s_pos[CT_CHUNKS];
    __shared__ int s_k[CT_CHUNKS];
    const int count = load_tile_chunks(alist, n_active, s_k);
    unsigned tile_base;
    const unsigned n = build_edge_list(g, E, s_k, count, s_list, s_pos, tile_base);
    if (n == 0) return;
    __syncthreads();
    const bool has_def = deform != nullptr;
    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const int axis = d & 3, j = (d >> 2) & 31;
        const TilePos tp = s_pos[(d >> 7) & 63];
        const int xp = tp.xp, yp = tp.yp, zp = 32 * tp.c + j;
        const int xq = xp + (axis == 0), yq = yp + (axis == 1), zq = zp + (axis == 2);
        const T d0 = fetch_padded(sdf, g, xp, yp, zp, padv);
        const T d1 = fetch_padded(sdf, g, xq, yq, zq, padv);
        const T t = edge_t(d0, d1, iso);
        Vec3<T> p0{T(xp), T(yp), T(zp)}, p1{T(xq), T(yq), T(zq)};
        if (has_def) {
            const Vec3<T> f0 = fetch_deform(deform, g, xp, yp, zp), f1 = fetch_deform(deform, g, xq, yq, zq);
            p0.x = p0.x + f0.x; p0.y = p0.y + f0.y; p0.z = p0.z + f0.z;
            p1.x = p1.x + f1.x; p1.y = p1.y + f1.y; p1.z = p1.z + f1.z;
        }
        Vec3<T> p;
        p.x = fma_rn(p1.x - p0.x, t, p0.x);
        p.y = fma_rn(p1.y - p0.y, t, p0.y);
        p.z = fma_rn(p1.z - p0.z, t, p0.z);
        p = epi.apply(p);
        T *dst = verts + (size_t)(tile_base + i) * 3;
        st_stream(dst, p.x); st_stream(dst + 1, p.y); st_stream(dst + 2, p.z);
    }
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

struct RecCache { const uint4 *s_E; const uint4 *E; const int *s_k; int sX, sY; bool contig; };

__device__ __forceinline__ RecCache load_record_cache(const Geo &g, const uint4 *__restrict__ E, const int *s_k, int count, uint4 *s_E)
{
    const bool contig = s_k[count - 1] - s_k[0] == count - 1;
    const int per = contig ? count + 1 : count;
    for (int i = threadIdx.x; i < 4 * per; i += CT_THREADS) {
        const int rs = i / per, cl = i - rs * per;
        const int kk = (contig ? s_k[0] + cl : s_k[cl]) + (rs >> 1) * g.sX + (rs & 1) * g.sY;   // E has a zero-filled tail
        s_E[rs * (CT_CHUNKS + 1) + cl] = __ldg(E + kk);
    }
    return RecCache{s_E, E, s_k, g.sX, g.sY, contig};
}

// vertex id of local edge e of cell (tile entry cl, lane j)
__device__ __forceinline__ unsigned edge_rank(const RecCache &rc, int cl, int j, int e)
{
    const int ax = (EDGE_AX >> (2 * e)) & 3;
    const int rs = (((EDGE_DX >> e) & 1) << 1) | ((EDGE_DY >> e) & 1);
    int jj = j + ((EDGE_DZ >> e) & 1);
    const int nxt = jj >> 5;     // dz = 1 from lane 31: first point of the following chunk
    jj &= 31;
    uint4 rec;
    if (nxt && !rc.contig) rec = __ldg(rc.E + rc.s_k[cl] + 1 + (rs >> 1) * rc.sX + (rs & 1) * rc.sY);
    else rec = rc.s_E[rs * (CT_CHUNKS + 1) + cl + nxt];
    const unsigned l = lanemask_lt(jj);
    unsigned r = rec.x + __popc(rec.y & l) + __popc(rec.z & l) + __popc(rec.w & l);
    if (ax >= 1) r += bit(rec.y, jj);
    if (ax == 2) r += bit(rec.z, jj);
    return r;
}

// ------------------------------------------------------------------------------------------
// K4 (v2): triangles, triangle-parallel.  Replaces count_cell_mc_tris / create_cell_mc_tris
// (cumc.cu:540-612) + the int64 widening (diso/__init__.py:61).
//   phase A  lane == cell: the per-cell word written by classify_scan (case index | offset of the
//            cell's first triangle) -> one descriptor {chunk-in-tile, lane, k-th triangle} per
//            triangle at slot (triangle id - first triangle id of the tile).
//   phase B  thread == triangle: three edge ids from the case table -> three vertex ids via the
//            record cache -> 24 contiguous bytes at the triangle's output rank.
// ------------------------------------------------------------------------------------------
constexpr int CT_MAX_TRIS = CT_CHUNKS * 160;

__global__ void __launch_bounds__(CT_THREADS) mc_tris_kernel(Geo g, const uint4 *__restrict__ E, const uint2 *__restrict__ F,
                                                           const unsigned short *__restrict__ C,
                                                           const unsigned *__restrict__ alist, int n_active,
                                                           long long *__restrict__ tris)
{
    __shared__ unsigned long long s_case[256];
    __shared__ uint4 s_E[CT_RECS];
    __shared__ unsigned short s_list[CT_MAX_TRIS];
    __shared__ unsigned char s_code[CT_CHUNKS * 32];
    __shared__ int s_k[CT_CHUNKS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int count = load_tile_chunks(alist, n_active, s_k);
    const unsigned tile_base = F[s_k[0]].x;
    const unsigned n = F[s_k[count - 1] + 1].x - tile_base;
    if (n == 0) return;
    s_case[threadIdx.x] = T_MC_CASE[threadIdx.x];
    const RecCache rc = load_record_cache(g, E, s_k, count, s_E);
    __syncthreads();

    constexpr int PER_WARP = CT_CHUNKS / CT_WARPS;
    {
        const int emine = wid * PER_WARP + lane;
        const int kmine = (lane < PER_WARP && emine < count) ? s_k[emine] : -1;
        uint2 f = make_uint2(0, 0);
        if (kmine >= 0) f = F[kmine];
        unsigned active = __ballot_sync(FULL, f.y != 0u);
        while (active) {
            const int i = __ffs(active) - 1;
            active &= active - 1;
            const int cl = wid * PER_WARP + i;
            const int k = __shfl_sync(FULL, kmine, i);
            const unsigned tb = __shfl_sync(FULL, f.x, i) - tile_base;
            const unsigned used = __shfl_sync(FULL, f.y, i);
            // per-cell word written by classify_scan: case index | offset of the cell's first triangle << 8
            const unsigned info = bit(used, lane) ? C[(size_t)k * 32 + lane] : 0u;
            const unsigned code = info & 0xffu;
            s_code[cl * 32 + lane] = (unsigned char)code;
            const unsigned nt = (unsigned)(s_case[code] >> 60);
            const unsigned slot = tb + (info >> 8);
#pragma unroll
            for (unsigned q = 0; q < 5; ++q)
                if (q < nt) s_list[slot + q] = (unsigned short)((cl << 8) | (lane << 3) | q);
        }
    }
    __syncthreads();

    for (unsigned i = threadIdx.x; i < n; i += CT_THREADS) {
        const unsigned d = s_list[i];
        const unsigned q = d & 7u;
        const int j = (d >> 3) & 31, cl = (d >> 8) & 63;
        const unsigned tri = (unsigned)(s_case[s_code[cl * 32 + j]] >> (12 * q)) & 0xfffu;
        const long long a = edge_rank(rc, cl, j, tri & 15u);
        const long long b = edge_rank(rc, cl, j, (tri >> 4) & 15u);
        const long long c = edge_rank(rc, cl, j, tri >> 8);
        long long *dst = tris + (size_t)(tile_base + i) * 3;
        st_stream(dst, a); st_stream(dst + 1, b); st_stream(dst + 2, c);
    }
}

}  // namespace diso
