// mc_backward_v2.cuh -- backward of the edge crossings from SAVED edge records.
//
// Replaces adj_create_cell_mc_verts_kernel (cumc.cu:474-512: one thread per used cell, 2+6 atomicAdds per
// vertex, sdf / deform re-gathered), the dense zero fills (diso/__init__.py:33,40) and the pad-backward slices;
// with GSRC = 2 / 3 it is the WHOLE DMC backward (adj_create_dmc_verts, cudualmc.cu:957-1005), one kernel as in the reference.
//
// Why a second design (round 2).  ncu on the first one (mc_backward_compact.cuh, kept for callers without saved
// records) at 512^3: 290 warp instructions and 104 LSU wavefronts per 32-point chunk, issue slots 67 % and the LSU
// data pipe 95 % busy -- both walls at once.  52 of the wavefronts and ~2/3 of the instructions are the eight scalar
// gathers per edge (sdf x2, deform x6) with their bounds logic and 64-bit addressing, issued from per-axis list segments
// whose 32 consecutive entries span ~5.5 rows; another 29 wavefronts are scalar read-modify-writes of the accumulators,
// 19 of them bank conflicts (rows collide at equal lane).  Here
//   * the forward's edge pass saves, per crossing edge, the five numbers its adjoint needs -- p1 - p0 (3), d0, d1 -- in
//     blocked SoA form indexed by the edge's global rank (edge_verts_kernel, compact.cuh; blk_index below).  The adjoint
//     reads NO sdf / deform, and because every ordering is ascending in the linear point index, the edges of one
//     row-chunk are ONE contiguous rank range;
//   * the block's edge list is in RANK order (row-major, point-major, axis-minor) instead of per-axis segments: a warp's
//     32 entries cover ~2 row-chunks, so the loads of the adjoints and the records are (nearly) coalesced and the
//     accumulator updates of one instruction fall into distinct banks;
//   * accumulators are {d sdf, d deform.xyz} quads: one LDS.128 + two packed FADD2 (Blackwell f32x2) + one STS.128 per
//     endpoint instead of four scalar read-modify-writes.
// The summation order is fixed (six conflict-free phases: axis x {start point, end point}), so results are
// deterministic without atomics, as before.
// (Also tried this round: "store once, gather once" -- entries stored to shared memory, one thread per output point
// summing its <= 6 incident entries found by popcount arithmetic, no read-modify-write at all.  Fewer LSU wavefronts,
// but ~300 instructions per chunk (150 of them the per-point gather): issue-bound at 1.80 ms, slower than v1.)
#pragma once
#include "mc_backward_compact.cuh"   // rcp_fast, bwd_mark_kernel

namespace diso {

constexpr int B2_THREADS = 256;
constexpr int B2_WARPS = B2_THREADS / 32;

// Saved edge records and SoA adjoints are stored in groups of 32 edges, component-major inside a group:
// element (rank, comp) of an NCOMP-component array lives at (rank / 32) (32 NCOMP) + 32 comp + rank % 32.  A warp
// touching 32 consecutive ranks reads / writes full 128-byte lines per component, and a thread needs ONE address for
// all of its components (the others are immediate offsets).
// experiment knob (build variants): how the read-once record / quad loads are issued.  Measured at 512^3 (mc / dmc backward):
// ld.global.nc 1.442 / 2.247 ms, ld.global.cs 1.467 / 2.304, ld.global.lu 1.457 / 2.290 -- the plain read-only path stays.
#ifndef DISO_REC_LD
#define DISO_REC_LD 0
#endif
template <typename T> __device__ __forceinline__ T ld_once(const T *p)
{
#if DISO_REC_LD == 1
    return __ldcs(p);
#elif DISO_REC_LD == 2
    return __ldlu(p);
#else
    return __ldg(p);
#endif
}
template <int NCOMP> __device__ __forceinline__ size_t blk_index(size_t rank) { return (rank >> 5) * (32 * NCOMP) + (rank & 31); }

template <typename T> struct alignas(4 * sizeof(T)) Quad { T d, x, y, z; };
template <typename T, bool HAS_DEF> struct EntOf { using type = Quad<T>; };
template <typename T> struct EntOf<T, false> { using type = T; };

// acc += e.  fp32: two packed FADD2 (Blackwell f32x2 pipe), each half rounded like a scalar add.
__device__ __forceinline__ void ent_add(Quad<float> &a, const Quad<float> &b)
{
    float2 *pa = reinterpret_cast<float2 *>(&a);
    const float2 *pb = reinterpret_cast<const float2 *>(&b);
    pa[0] = __fadd2_rn(pa[0], pb[0]);
    pa[1] = __fadd2_rn(pa[1], pb[1]);
}
__device__ __forceinline__ void ent_add(Quad<double> &a, const Quad<double> &b) { a.d = a.d + b.d; a.x = a.x + b.x; a.y = a.y + b.y; a.z = a.z + b.z; }
__device__ __forceinline__ void ent_add(float &a, const float &b) { a = a + b; }
__device__ __forceinline__ void ent_add(double &a, const double &b) { a = a + b; }
template <typename T> __device__ __forceinline__ T ent_d(const Quad<T> &a) { return a.d; }
__device__ __forceinline__ float ent_d(const float &a) { return a; }
__device__ __forceinline__ double ent_d(const double &a) { return a; }

template <typename T, bool HAS_DEF, bool DMC, int BX, int BY> struct Bwd2Layout {
    static constexpr int ROWS = (BX + 1) * (BY + 1);          // candidate rows incl. the -x / -y halo
    static constexpr int PTS = BX * BY * 32;                  // output points per block
    using Ent = typename EntOf<T, HAS_DEF>::type;
    // worst-case list length: own rows 96 + 1 entries, halo rows 32 (their one relevant axis)
    static constexpr int CAP = BX * BY * 97 + (BX + BY) * 32;
    static constexpr size_t off_acc = 0;                                   // Ent [PTS]
    static constexpr size_t off_rec = off_acc + PTS * sizeof(Ent);         // uint4 [ROWS]
    static constexpr size_t off_off = off_rec + ROWS * 16;                 // u32 [ROWS + 1]   exclusive prefix of the entries
    static constexpr size_t off_zin = off_off + (ROWS + 1) * 4;            // u32 [ROWS]
    static constexpr size_t off_delta = off_zin + ROWS * 4;                // u32 [ROWS]       own rows: rank - list slot
    static constexpr size_t off_info = (off_delta + ROWS * 4 + 15) / 16 * 16;  // int4 [ROWS]  accumulator bases {own, +x target, +y target, -}
    static constexpr size_t off_rowb = off_info + ROWS * 16;               // i64 [ROWS]       output element of lane 0 (own rows)
    static constexpr size_t off_tab = (off_rowb + ROWS * 8 + 15) / 16 * 16;    // DMC: 1 / patch length, T [8]
    // the edge list is dead once the last entry has been evaluated, the deform write-out stage (T [WARPS][96]) lives after that:
    // they share one region (3 KB less per CTA: 7 instead of 6 CTAs fit the 164 KB carve-out)
    static constexpr size_t off_list = (off_tab + (DMC ? 8 * sizeof(T) : 0) + 15) / 16 * 16;   // u16 [CAP]
    static constexpr size_t off_stage = off_list;
    static constexpr size_t stage_bytes = HAS_DEF ? B2_WARPS * 96 * sizeof(T) : 0;
    static constexpr size_t bytes = off_list + ((size_t)CAP * 2 > stage_bytes ? (size_t)CAP * 2 : stage_bytes) + 16;
    static_assert(ROWS <= 64 && 4 * ROWS <= B2_THREADS, "four threads per row build the list; row index fits the descriptor");
};

// Where dL/d(edge crossing) comes from.
//   GSRC == 0: gsrc is [n, 3] AoS (autograd's dL/dverts of DiffMC);
//   GSRC == 1: gsrc is blocked SoA (stage A of an unfused DMC backward);
//   GSRC == 2 / 3: DMC, evaluated HERE (exact / reference-compatible adjoint of the dual-vertex averaging,
//                  adj_create_dmc_verts, cudualmc.cu:957-1005, which is one kernel in the reference too): the edge's
//                  adjoint is the sum over the 4 cells around it of adj_dual[patch of the edge in that cell] / len(patch).
//                  The four dual-vertex ids ARE the edge's quad (the forward's own output, saved by autograd for free) and
//                  the patch lengths / patch indices were saved by the quad kernel as a 6th record component, so this needs
//                  no cell words, patch bases or case tables: two 128-bit loads + one word per edge, all by rank.
struct DmcSrc {
    const long long *quads;     // [n_quads, 4] the forward's quads (ids shifted by id_offset in a slab frame)
    long long id_offset;
};

template <typename T, bool HAS_DEF, int GSRC, int BX, int BY, bool OUTPUTS_ZEROED>
__device__ __forceinline__ void mc_backward2_block(const Geo &g, T iso, T ix, T iy, T iz, const uint4 *__restrict__ E,
                                                   const T *__restrict__ gsrc, const DmcSrc dmc, const T *__restrict__ rec,
                                                   T *__restrict__ adj_sdf, T *__restrict__ adj_deform,
                                                   int tx, int ty, int c)
{
    constexpr bool DMC = GSRC >= 2;
    using L = Bwd2Layout<T, HAS_DEF, (GSRC >= 2), BX, BY>;
    using Ent = typename L::Ent;
    constexpr int ROWS = L::ROWS, PTS = L::PTS;
    extern __shared__ __align__(32) unsigned char smem2_raw[];
    Ent *s_acc = reinterpret_cast<Ent *>(smem2_raw + L::off_acc);
    uint4 *s_rec = reinterpret_cast<uint4 *>(smem2_raw + L::off_rec);
    unsigned *s_off = reinterpret_cast<unsigned *>(smem2_raw + L::off_off);
    unsigned *s_zin = reinterpret_cast<unsigned *>(smem2_raw + L::off_zin);
    unsigned *s_delta = reinterpret_cast<unsigned *>(smem2_raw + L::off_delta);
    int4 *s_info = reinterpret_cast<int4 *>(smem2_raw + L::off_info);
    long long *s_rowb = reinterpret_cast<long long *>(smem2_raw + L::off_rowb);
    T *s_inv = reinterpret_cast<T *>(smem2_raw + L::off_tab);
    T *s_stage = reinterpret_cast<T *>(smem2_raw + L::off_stage);
    unsigned short *s_list = reinterpret_cast<unsigned short *>(smem2_raw + L::off_list);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int xp0 = 1 + tx * BX, yp0 = 1 + ty * BY;  // padded coords of the first output row
    const bool zfull = c > 0 && 32 * c + 31 <= g.Z;   // every point of the chunk is a real grid point

    // ---- 1. edge records of the candidate rows, entries per row --------------------------------------------------
    bool mine_any = false;
    if (tid < ROWS) {
        const int dxr = tid / (BY + 1), dyr = tid - dxr * (BY + 1);
        const int xp = xp0 - 1 + dxr, yp = yp0 - 1 + dyr;
        uint4 r4 = make_uint4(0, 0, 0, 0);
        unsigned zin = 0;
        if (xp <= g.X + 1 && yp <= g.Y + 1) {
            const int k = (xp * g.PY + yp) * g.NC + c;
            r4 = E[k];
            if (c > 0 && dxr >= 1 && dyr >= 1) zin = E[k - 1].w >> 31;
        }
        unsigned cnt;
        if (dxr >= 1 && dyr >= 1) cnt = __popc(r4.y) + __popc(r4.z) + __popc(r4.w) + zin;   // own row: every edge
        else if (dxr == 0 && dyr == 0) cnt = 0;              // corner: touches nothing
        else if (dxr == 0) cnt = __popc(r4.y);               // -x halo row: its +x edges end in the block
        else cnt = __popc(r4.z);                             // -y halo row: its +y edges
        s_rec[tid] = r4;
        s_zin[tid] = zin;
        s_off[tid] = cnt;
        mine_any = cnt != 0u;
        // accumulator slots of lane 0 of: this row (own rows), the row its +x edges end in, the row its +y edges end in
        // (negative: outside the block's output region).  One LDS.128 per entry instead of a dozen index instructions.
        constexpr int OUT = -(1 << 20);
        const int ox = dxr - 1, oy = dyr - 1;
        const bool own = ox >= 0 && oy >= 0;
        s_info[tid] = make_int4(own ? (ox * BY + oy) * 32 : OUT, (oy >= 0 && ox + 1 < BX) ? ((ox + 1) * BY + oy) * 32 : OUT,
                                (ox >= 0 && oy + 1 < BY) ? (ox * BY + oy + 1) * 32 : OUT, 0);
        s_rowb[tid] = (own && xp <= g.X && yp <= g.Y) ? ((long long)(xp - 1) * g.Y + (yp - 1)) * g.Z + (32 * c - 1) : -(1ll << 62);
    }
    if (!__syncthreads_or(mine_any)) {
        // no crossing edge touches this block (the common case on smooth surfaces): zeros, straight from registers
        for (int o = wid; !OUTPUTS_ZEROED && o < BX * BY; o += B2_WARPS) {
            const long long rowb = s_rowb[(o / BY + 1) * (BY + 1) + o % BY + 1];
            if (rowb < -1) continue;
            const int zp = 32 * c + lane;
            if (adj_sdf && zp >= 1 && zp <= g.Z) st_stream(adj_sdf + rowb + lane, T(0));
            if (HAS_DEF && adj_deform) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int e = lane + 32 * q;
                    const int zz = 32 * c + e / 3;
                    if (zz >= 1 && zz <= g.Z) st_stream(adj_deform + 3 * rowb + e, T(0));
                }
            }
        }
        return;
    }
    if (DMC && tid < 8) s_inv[tid] = tid ? T(1) / T(tid) : T(0);   // (read after the barrier below)
    if (wid == 0) {  // exclusive scan of the ROWS counts (<= two per lane)
        constexpr int PER = (ROWS + 31) / 32;
        unsigned v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { const int idx = lane * PER + i; v[i] = idx < ROWS ? s_off[idx] : 0u; sum += v[i]; }
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
        unsigned run = inc - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int idx = lane * PER + i;
            if (idx < ROWS) { s_off[idx] = run; s_delta[idx] = s_rec[idx].x - (s_zin[idx] & 1u) - run; }
            run += v[i];
        }
        if (lane == 31) s_off[ROWS] = inc;
    } else {
        // zero the accumulators meanwhile (128-bit stores)
        uint4 *z = reinterpret_cast<uint4 *>(s_acc);
        constexpr int NZ = (int)(PTS * sizeof(Ent) / 16);
        for (int i = tid - 32; i < NZ; i += B2_THREADS - 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const unsigned n = s_off[ROWS];

    // ---- 2. descriptors {axis:2 | row:7 | lane+1:6} in RANK order: thread == one byte of a row's masks -----------
    if (tid < 4 * ROWS) {
        const int r = tid >> 2, sh = (tid & 3) * 8;
        const uint4 r4 = s_rec[r];
        const int dxr = r / (BY + 1), dyr = r - dxr * (BY + 1);
        const unsigned lt = (1u << sh) - 1u;
        const unsigned dbase = ((unsigned)r << 6) | (unsigned)(sh + 1);
        const unsigned o = s_off[r];
        if (dxr >= 1 && dyr >= 1) {
            const unsigned zin = s_zin[r] & 1u;
            if (zin && sh == 0) s_list[o] = (unsigned short)(((unsigned)r << 6) | (2u << 13));   // lane -1: +z edge of the previous chunk's last point
            const unsigned bx = (r4.y >> sh) & 0xffu, by = (r4.z >> sh) & 0xffu, bz = (r4.w >> sh) & 0xffu;
            unsigned slot = o + zin + __popc(r4.y & lt) + __popc(r4.z & lt) + __popc(r4.w & lt);
            for (unsigned m = bx | by | bz; m; m &= m - 1) {
                const unsigned j = (unsigned)(__ffs(m) - 1);
                const unsigned d = dbase + j;
                if ((bx >> j) & 1u) s_list[slot++] = (unsigned short)d;
                if ((by >> j) & 1u) s_list[slot++] = (unsigned short)(d | (1u << 13));
                if ((bz >> j) & 1u) s_list[slot++] = (unsigned short)(d | (2u << 13));
            }
        } else if (dxr == 0 && dyr >= 1) {
            unsigned slot = o + __popc(r4.y & lt);
            for (unsigned m = (r4.y >> sh) & 0xffu; m; m &= m - 1) s_list[slot++] = (unsigned short)(dbase + (unsigned)(__ffs(m) - 1));
        } else if (dyr == 0 && dxr >= 1) {
            unsigned slot = o + __popc(r4.z & lt);
            for (unsigned m = (r4.z >> sh) & 0xffu; m; m &= m - 1) s_list[slot++] = (unsigned short)((dbase + (unsigned)(__ffs(m) - 1)) | (1u << 13));
        }
    }
    __syncthreads();

    // ---- 3. evaluate each edge once (thread == entry), accumulate in six conflict-free phases --------------------
    for (unsigned lo = 0; lo < n; lo += B2_THREADS) {
        const unsigned i = lo + tid;
        int axis = -1, p0 = -1, p1 = -1;
        Ent c0, c1;
        if (i < n) {
            const unsigned d = s_list[i];
            axis = d >> 13;
            const int r = (d >> 6) & 127, j = (int)(d & 63u) - 1;
            const int4 inf = s_info[r];
            unsigned rank;
            if (inf.x >= 0) {
                rank = i + s_delta[r];   // own rows hold ALL their edges in rank order: contiguous
            } else {
                const uint4 r4 = s_rec[r];   // halo rows list one axis only
                const unsigned l = lanemask_lt(j);
                rank = r4.x + __popc(r4.y & l) + __popc(r4.z & l) + __popc(r4.w & l);
                if (axis == 1) rank += bit(r4.y, j);
            }
            T gx, gy, gz;
            // record components per edge: {p1 - p0 (x, y, z), d0, d1} with deform, {d0, d1} without (p1 - p0 = the unit axis
            // vector); DMC extractions: + the quad meta word
            constexpr int NBASE = HAS_DEF ? 5 : 2;
            constexpr int NREC = NBASE + (GSRC >= 1 ? 1 : 0);
            const T *rp = rec + blk_index<NREC>(rank);
            if constexpr (DMC) {
                // stage A of adj_create_dmc_verts (cudualmc.cu:957-1005): same operations and order as dmc_edges2_kernel<1|2>
                const longlong2 *qp = reinterpret_cast<const longlong2 *>(dmc.quads + (size_t)rank * 4);
                const longlong2 qa = ld_once(qp), qb = ld_once(qp + 1);
                const unsigned meta = ld_once(reinterpret_cast<const unsigned *>(rp + 32 * NBASE));
                const long long id[4] = {qa.x, qa.y, qb.x, qb.y};
                Vec3<T> acc{T(0), T(0), T(0)};
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const unsigned mm = meta >> (5 * cc);
                    // reference mode: the cell's FIRST dual vertex (cudualmc.cu:975,990) = this patch's id - its index in the cell
                    const long long src = id[cc] - dmc.id_offset - (GSRC == 3 ? (long long)((mm >> 3) & 3u) : 0ll);
                    const T inv = s_inv[mm & 7u];
                    const T *pa = gsrc + (size_t)src * 3;
                    acc.x = fma_rn(__ldg(pa), inv, acc.x);
                    acc.y = fma_rn(__ldg(pa + 1), inv, acc.y);
                    acc.z = fma_rn(__ldg(pa + 2), inv, acc.z);
                }
                gx = acc.x * ix; gy = acc.y * iy; gz = acc.z * iz;
            } else {
                if (GSRC == 1) { const T *gp = gsrc + blk_index<3>(rank); gx = __ldg(gp); gy = __ldg(gp + 32); gz = __ldg(gp + 64); }
                else { const T *gp = gsrc + (size_t)rank * 3; gx = __ldg(gp); gy = __ldg(gp + 1); gz = __ldg(gp + 2); }
                gx = gx * ix; gy = gy * iy; gz = gz * iz;
            }
            const T d0 = ld_once(rp + 32 * (NBASE - 2)), d1 = ld_once(rp + 32 * (NBASE - 1));
            // adjComputeMcVert (cumc.cu:412-453) with one reciprocal: (iso - d1) / (d1 - d0)^2 * adj_t etc.
            // without deform p1 - p0 is the unit axis vector (not stored); the full dot product is kept so that a non-finite
            // gradient scale (normalize with a 1-point dimension: 1 / (dims - 1) = inf) propagates as in the reference
            T dpx = T(axis == 0), dpy = T(axis == 1), dpz = T(axis == 2);
            if constexpr (HAS_DEF) { dpx = ld_once(rp); dpy = ld_once(rp + 32); dpz = ld_once(rp + 64); }
            T adj_t = dpx * gx;
            adj_t = fma_rn(dpy, gy, adj_t);
            adj_t = fma_rn(dpz, gz, adj_t);
            const T rr = rcp_fast(d1 - d0);
            const T sc = adj_t * rr * rr;
            const T c0d = (iso - d1) * sc, c1d = (d0 - iso) * sc;
            if constexpr (HAS_DEF) {
                const T t = clamp01((iso - d0) * rr);
                const T w0 = T(1) - t;
                c0 = Quad<T>{c0d, w0 * gx, w0 * gy, w0 * gz};
                c1 = Quad<T>{c1d, t * gx, t * gy, t * gz};
            } else {
                c0 = c0d;
                c1 = c1d;
            }
            // accumulator slots of the two endpoints (negative: outside this block's output region)
            p0 = j >= 0 ? inf.x + j : -1;
            p1 = axis == 2 ? (j < 31 ? inf.x + j + 1 : -1) : (axis == 0 ? inf.y : inf.z) + j;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (axis == a && p0 >= 0) { Ent v = s_acc[p0]; ent_add(v, c0); s_acc[p0] = v; }
            __syncthreads();
            if (axis == a && p1 >= 0) { Ent v = s_acc[p1]; ent_add(v, c1); s_acc[p1] = v; }
            __syncthreads();
        }
    }

    // ---- 4. dense write-out: one warp per output row-chunk, lane == point ------------------------------------------
    // Warp w owns the RPW consecutive rows o = w * RPW + it: they share one x row of the block when BY % RPW == 0, so the row's
    // slot in s_rowb is a per-warp base + it (the strided assignment o = w + it * B2_WARPS cost a division per row: 12 of the ~55
    // instructions of a row's write-out).
    constexpr bool RPW_OK = (BX * BY) % B2_WARPS == 0 && BY % ((BX * BY) / B2_WARPS ? (BX * BY) / B2_WARPS : 1) == 0;
    constexpr int RPW = RPW_OK ? (BX * BY) / B2_WARPS : 1;
    const int o0 = RPW_OK ? wid * RPW : wid;
    const int slot0 = (o0 / BY + 1) * (BY + 1) + o0 % BY + 1;
#pragma unroll
    for (int it = 0; it < (RPW_OK ? RPW : (BX * BY + B2_WARPS - 1) / B2_WARPS); ++it) {
        const int o = RPW_OK ? o0 + it : wid + it * B2_WARPS;
        if (!RPW_OK && o >= BX * BY) break;
        const long long rowb = s_rowb[RPW_OK ? slot0 + it : (o / BY + 1) * (BY + 1) + o % BY + 1];   // element of lane 0
        if (rowb < -1) continue;                                                // row outside the grid
        Ent acc;
        if constexpr (HAS_DEF && sizeof(T) == 4) {
            // one LDS.128 (the compiler splits a struct load into two 64-bit loads, which conflict 2-way at this stride)
            float4 q;
            const unsigned a = (unsigned)__cvta_generic_to_shared(s_acc + o * 32 + lane);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(a));
            acc = Quad<float>{q.x, q.y, q.z, q.w};
        } else {
            acc = s_acc[o * 32 + lane];
        }
        if (adj_sdf) {
            const int zp = 32 * c + lane;
            if (zfull || (zp >= 1 && zp <= g.Z)) st_stream(adj_sdf + rowb + lane, ent_d(acc));
        }
        if constexpr (HAS_DEF) {
            if (adj_deform) {
                T *st = s_stage + wid * 96;
                st[3 * lane] = acc.x; st[3 * lane + 1] = acc.y; st[3 * lane + 2] = acc.z;   // stride 3: conflict-free
                __syncwarp();
                T *of = adj_deform + 3 * rowb + lane;
                if (zfull) {
                    st_stream(of, st[lane]); st_stream(of + 32, st[lane + 32]); st_stream(of + 64, st[lane + 64]);
                } else {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int zz = 32 * c + (lane + 32 * q) / 3;
                        if (zz >= 1 && zz <= g.Z) st_stream(of + 32 * q, st[lane + 32 * q]);
                    }
                }
                __syncwarp();
            }
        }
    }
}

// Dense launch: one CTA per block of the (chunk, y tile, x tile) grid.
template <typename T, bool HAS_DEF, int GSRC, int BX, int BY>
__global__ void __launch_bounds__(B2_THREADS) mc_backward2_kernel(Geo g, T iso, T ix, T iy, T iz, const uint4 *__restrict__ E,
                                                                const T *__restrict__ gsrc, DmcSrc dmc, const T *__restrict__ rec,
                                                                T *__restrict__ adj_sdf, T *__restrict__ adj_deform, int ntx,
                                                                int nty, int flat)
{
    int c, ty, tx;
    if (flat) {   // degenerate shapes whose tile counts exceed the y / z grid limits
        int b = blockIdx.x;
        c = b % g.NC; b /= g.NC;
        ty = b % nty; tx = b / nty;
    } else {
        c = blockIdx.x; ty = blockIdx.y; tx = blockIdx.z;
    }
    mc_backward2_block<T, HAS_DEF, GSRC, BX, BY, false>(g, iso, ix, iy, iz, E, gsrc, dmc, rec, adj_sdf, adj_deform, tx, ty, c);
}

// Sparse surfaces: outputs zero-filled by cudaMemsetAsync, bwd_mark_kernel (mc_backward_compact.cuh) lists the
// touched blocks, a persistent grid pulls them from that list.
template <typename T, bool HAS_DEF, int GSRC, int BX, int BY>
__global__ void __launch_bounds__(B2_THREADS) mc_backward2_queue_kernel(Geo g, T iso, T ix, T iy, T iz, const uint4 *__restrict__ E,
                                                                      const T *__restrict__ gsrc, DmcSrc dmc, const T *__restrict__ rec,
                                                                      T *__restrict__ adj_sdf, T *__restrict__ adj_deform, int nty,
                                                                      unsigned *__restrict__ work)
{
    __shared__ unsigned s_next;
    const unsigned count = work[0];
    while (true) {
        if (threadIdx.x == 0) s_next = atomicAdd(&work[1], 1u);
        __syncthreads();
        const unsigned i = s_next;
        if (i >= count) break;
        unsigned b = work[16 + i];
        const int c = (int)(b % (unsigned)g.NC); b /= (unsigned)g.NC;
        const int ty = (int)(b % (unsigned)nty), tx = (int)(b / (unsigned)nty);
        mc_backward2_block<T, HAS_DEF, GSRC, BX, BY, true>(g, iso, ix, iy, iz, E, gsrc, dmc, rec, adj_sdf, adj_deform, tx, ty, c);
        __syncthreads();   // shared memory (and s_next) is reused by the next block
    }
}

}  // namespace diso
