// mc_backward_compact.cuh -- backward of the edge crossings, edge-parallel and atomic-free.
//
// Replaces adj_create_cell_mc_verts_kernel (cumc.cu:474-512: one thread per used cell, 2+6
// atomicAdds per vertex), the dense zero fills (diso/__init__.py:33,40) and the pad-backward
// slices; it is also stage B of the DMC backward (cudualmc.cu:957-1005).
//
// One CTA owns a BX x BY x 32 block of real grid points (one 32-point chunk in z):
//   1. the edge records of the (BX+1) x (BY+1) rows that can touch the block (own rows plus the
//      -x / -y halo rows) are loaded; per (row, axis) crossing-edge counts are scanned;
//   2. every incident crossing edge gets a 16-bit descriptor in ONE shared list (x-edges, then
//      y-edges, then z-edges; the -z halo edge of each row is the "lane -1" entry);
//   3. thread == edge: each edge is evaluated exactly once per block by fully populated warps
//      (ncu on the lane-per-point version: 15 of 32 threads active, 540 instructions per chunk);
//   4. the two endpoint contributions go to a shared-memory accumulator tile in six phases
//      (axis x {start point, end point}).  Within a phase every accumulator is touched by at
//      most one thread, so plain read-modify-writes suffice and the summation order is fixed:
//      deterministic results without atomics;
//   5. the tile is written out densely (zeros included): adj_sdf, and adj_deform as contiguous
//      96-element row-chunks.
// A block that no crossing edge touches skips 1-4 and only writes zeros.
#pragma once
#include "compact.cuh"

namespace diso {

constexpr int BC_THREADS = 256;

__device__ __forceinline__ float rcp_fast(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }

template <typename T, bool HAS_DEF, int BC_X, int BC_Y, bool OUTPUTS_ZEROED>
__device__ __forceinline__ void mc_backward_block(const T *__restrict__ sdf, const T *__restrict__ deform, const Geo &g, T iso,
                                                  T padv, T ix, T iy, T iz, const uint4 *__restrict__ E,
                                                  const T *__restrict__ gsrc, T *__restrict__ adj_sdf,
                                                  T *__restrict__ adj_deform, int tx, int ty, int c)
{
    constexpr int BC_ROWS = (BC_X + 1) * (BC_Y + 1);     // candidate rows incl. the -x / -y halo
    constexpr int BC_PTS = BC_X * BC_Y * 32;             // output points per block
    constexpr int BC_ACC = BC_PTS * (HAS_DEF ? 4 : 1);   // accumulators: [BC_PTS] sdf, then [BC_PTS * 3] deform
    static_assert(BC_ROWS <= 127 && 4 * BC_ROWS <= BC_THREADS, "row index must fit the 7-bit descriptor field; 4 threads per row in step 2");
    static_assert((BC_ACC * sizeof(T)) % 16 == 0, "accumulators are zeroed with 128-bit stores");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_accd = reinterpret_cast<T *>(smem_raw);                       // [BC_PTS]
    T *s_accf = s_accd + BC_PTS;                                       // [BC_PTS * 3] (HAS_DEF only)
    uint4 *s_rec = reinterpret_cast<uint4 *>(s_accd + BC_ACC);         // [BC_ROWS]
    int4 *s_row = reinterpret_cast<int4 *>(s_rec + BC_ROWS);           // [BC_ROWS] {rowbase, flags, xs, ys}
    unsigned *s_off = reinterpret_cast<unsigned *>(s_row + BC_ROWS);   // [3 * BC_ROWS + 1]
    unsigned *s_zin = s_off + 3 * BC_ROWS + 1;                         // [BC_ROWS] -z halo edge present
    unsigned short *s_list = reinterpret_cast<unsigned short *>(s_zin + BC_ROWS);  // [BC_CAP]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int xp0 = 1 + tx * BC_X, yp0 = 1 + ty * BC_Y;  // padded coords of the first output row

    // ---- 1. records + counts ---------------------------------------------------------------------
    if (tid < BC_ROWS) {
        const int dxr = tid / (BC_Y + 1), dyr = tid - dxr * (BC_Y + 1);
        const int xp = xp0 - 1 + dxr, yp = yp0 - 1 + dyr;
        uint4 rec = make_uint4(0, 0, 0, 0);
        unsigned zin = 0;
        if (xp <= g.X + 1 && yp <= g.Y + 1) {
            const int k = (xp * g.PY + yp) * g.NC + c;
            rec = E[k];
            if (c > 0 && dxr >= 1 && dyr >= 1) zin = E[k - 1].w >> 31;
        }
        s_rec[tid] = rec;
        s_zin[tid] = zin;
        const unsigned fl = (unsigned)(xp >= 1 && xp <= g.X) | ((unsigned)(yp >= 1 && yp <= g.Y) << 1) |
                            ((unsigned)(xp + 1 <= g.X) << 2) | ((unsigned)(yp + 1 <= g.Y) << 3);
        s_row[tid] = make_int4(((xp - 1) * g.Y + (yp - 1)) * g.Z + 32 * c - 1, (int)fl, xp, yp);
        s_off[tid] = dyr >= 1 ? __popc(rec.y) : 0;                                   // x-edges: rows with dy >= 0
        s_off[BC_ROWS + tid] = dxr >= 1 ? __popc(rec.z) : 0;                         // y-edges: rows with dx >= 0
        s_off[2 * BC_ROWS + tid] = (dxr >= 1 && dyr >= 1) ? __popc(rec.w) + zin : 0;  // z-edges: output rows
    }
    bool mine_any = false;
    if (tid < BC_ROWS) mine_any = (s_off[tid] | s_off[BC_ROWS + tid] | s_off[2 * BC_ROWS + tid]) != 0u;
    if (!__syncthreads_or(mine_any)) {
        // no crossing edge touches this block (the common case on smooth surfaces): zeros, straight from registers
        for (int r = wid; !OUTPUTS_ZEROED && r < BC_X * BC_Y; r += BC_THREADS / 32) {
            const int ox = r / BC_Y, oy = r - ox * BC_Y;
            const int xp = xp0 + ox, yp = yp0 + oy;
            if (xp > g.X || yp > g.Y) continue;
            const long long rowb = ((long long)(xp - 1) * g.Y + (yp - 1)) * g.Z + (32 * c - 1);
            const int zp = 32 * c + lane;
            if (zp >= 1 && zp <= g.Z) st_stream(adj_sdf + rowb + lane, T(0));
            if (HAS_DEF) {
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
    if (wid == 0) {  // exclusive scan of 3*BC_ROWS counts (a few per lane)
        constexpr int N = 3 * BC_ROWS, PER = (N + 31) / 32;
        unsigned v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { const int idx = lane * PER + i; v[i] = idx < N ? s_off[idx] : 0u; sum += v[i]; }
        unsigned inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
        unsigned run = inc - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { const int idx = lane * PER + i; if (idx < N) s_off[idx] = run; run += v[i]; }
        if (lane == 31) s_off[N] = inc;
    }
    // zero the accumulators meanwhile (128-bit stores)
    {
        uint4 *z = reinterpret_cast<uint4 *>(s_accd);
        constexpr int NZ = (int)(BC_ACC * sizeof(T) / 16);
        for (int i = tid; i < NZ; i += BC_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const unsigned n = s_off[3 * BC_ROWS];

    if (n) {
        // ---- 2. descriptors {axis:2 | row:7 | lane+1:6}: thread == one byte of a row's masks -----------
        if (tid < 4 * BC_ROWS) {
            const int r = tid >> 2, sh = (tid & 3) * 8;
            const uint4 rec = s_rec[r];
            const int dxr = r / (BC_Y + 1), dyr = r - dxr * (BC_Y + 1);
            const unsigned lt = (1u << sh) - 1u;
            const unsigned dbase = ((unsigned)r << 6) | (unsigned)(sh + 1);
            if (dyr >= 1) {
                unsigned m = (rec.y >> sh) & 0xffu;
                unsigned slot = s_off[r] + __popc(rec.y & lt);
                for (; m; m &= m - 1) s_list[slot++] = (unsigned short)(dbase + (unsigned)(__ffs(m) - 1));
            }
            if (dxr >= 1) {
                unsigned m = (rec.z >> sh) & 0xffu;
                unsigned slot = s_off[BC_ROWS + r] + __popc(rec.z & lt);
                for (; m; m &= m - 1) s_list[slot++] = (unsigned short)((dbase + (unsigned)(__ffs(m) - 1)) | (1u << 13));
                if (dyr >= 1) {
                    const unsigned zin = s_zin[r];
                    const unsigned o = s_off[2 * BC_ROWS + r];
                    if (zin && sh == 0) s_list[o] = (unsigned short)(((unsigned)r << 6) | (2u << 13));  // lane -1
                    m = (rec.w >> sh) & 0xffu;
                    slot = o + zin + __popc(rec.w & lt);
                    for (; m; m &= m - 1) s_list[slot++] = (unsigned short)((dbase + (unsigned)(__ffs(m) - 1)) | (2u << 13));
                }
            }
        }
        __syncthreads();

        // ---- 3 + 4. evaluate each edge once, accumulate in conflict-free phases ----------------------
        const unsigned seg1 = s_off[BC_ROWS], seg2 = s_off[2 * BC_ROWS];  // list = [x | y | z]
        const int sZ = g.Z, sYZ = g.Y * g.Z;
        for (unsigned lo = 0; lo < n; lo += BC_THREADS) {
            const unsigned i = lo + tid;
            int axis = -1, p0 = -1, p1 = -1;
            T c0d = T(0), c1d = T(0);
            Vec3<T> c0f{T(0), T(0), T(0)}, c1f{T(0), T(0), T(0)};
            if (i < n) {
                const unsigned d = s_list[i];
                axis = d >> 13;
                const int r = (d >> 6) & 127, j = (int)(d & 63u) - 1;
                const int4 ri = s_row[r];
                const int xs = ri.z, ys = ri.w, zs = 32 * c + j;
                const int ze = zs + (axis == 2);
                const uint4 rec = s_rec[r];
                unsigned rank;
                if (j >= 0) {
                    const unsigned l = lanemask_lt(j);
                    rank = rec.x + __popc(rec.y & l) + __popc(rec.z & l) + __popc(rec.w & l);
                    if (axis >= 1) rank += bit(rec.y, j);
                    if (axis == 2) rank += bit(rec.z, j);
                } else {
                    rank = rec.x - 1u;  // +z edge of the previous chunk's last point
                }
                const T *gp = gsrc + (size_t)rank * 3;
#ifdef DISO_GSRC_CS
                const T gx = __ldcs(gp) * ix, gy = __ldcs(gp + 1) * iy, gz = __ldcs(gp + 2) * iz;
#else
                const T gx = __ldg(gp) * ix, gy = __ldg(gp + 1) * iy, gz = __ldg(gp + 2) * iz;
#endif
                const unsigned fl = (unsigned)ri.y;
                const unsigned need1 = axis == 0 ? 6u : (axis == 1 ? 9u : 3u);
                const bool v0 = (fl & 3u) == 3u && (unsigned)(zs - 1) < (unsigned)g.Z;
                const bool v1 = (fl & need1) == need1 && (unsigned)(ze - 1) < (unsigned)g.Z;
                const int i0 = ri.x + j;
                const int i1 = i0 + (axis == 0 ? sYZ : (axis == 1 ? sZ : 1));
                const T d0 = v0 ? __ldg(sdf + i0) : padv;
                const T d1 = v1 ? __ldg(sdf + i1) : padv;
                // p1 - p0 = unit vector of the axis (+ deformation difference; the pad layer carries none)
                T dpx = T(axis == 0 ? 1 : 0), dpy = T(axis == 1 ? 1 : 0), dpz = T(axis == 2 ? 1 : 0);
                if (HAS_DEF) {
                    T p0x = T(xs), p0y = T(ys), p0z = T(zs);
                    T p1x = T(xs + (axis == 0)), p1y = T(ys + (axis == 1)), p1z = T(ze);
                    if (v0) {
                        const T *f = deform + (size_t)i0 * 3;
                        p0x = p0x + __ldg(f); p0y = p0y + __ldg(f + 1); p0z = p0z + __ldg(f + 2);
                    }
                    if (v1) {
                        const T *f = deform + (size_t)i1 * 3;
                        p1x = p1x + __ldg(f); p1y = p1y + __ldg(f + 1); p1z = p1z + __ldg(f + 2);
                    }
                    dpx = p1x - p0x; dpy = p1y - p0y; dpz = p1z - p0z;
                }
                const T rr = rcp_fast(d1 - d0);
                T adj_t = dpx * gx;
                adj_t = fma_rn(dpy, gy, adj_t);
                adj_t = fma_rn(dpz, gz, adj_t);
                const T s = adj_t * rr * rr;
                c0d = (iso - d1) * s;
                c1d = (d0 - iso) * s;
                if (HAS_DEF) {
                    const T t = clamp01((iso - d0) * rr);
                    const T w0 = T(1) - t;
                    c0f = Vec3<T>{w0 * gx, w0 * gy, w0 * gz};
                    c1f = Vec3<T>{t * gx, t * gy, t * gz};
                }
                // accumulator slots of the two endpoints (-1: outside this block's output region)
                const int ox = xs - xp0, oy = ys - yp0;
                if (ox >= 0 && oy >= 0 && j >= 0) p0 = (ox * BC_Y + oy) * 32 + j;
                const int ex = ox + (axis == 0), ey = oy + (axis == 1), ej = j + (axis == 2);
                if (ex >= 0 && ex < BC_X && ey >= 0 && ey < BC_Y && ej < 32) p1 = (ex * BC_Y + ey) * 32 + ej;
            }
            const unsigned hi = min(lo + BC_THREADS, n);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const unsigned s_lo = a == 0 ? 0u : (a == 1 ? seg1 : seg2), s_hi = a == 0 ? seg1 : (a == 1 ? seg2 : n);
                if (s_lo >= hi || s_hi <= lo) continue;  // uniform: this round holds no edge of axis a
                if (axis == a && p0 >= 0) {
                    s_accd[p0] = s_accd[p0] + c0d;
                    if (HAS_DEF) { T *q = s_accf + 3 * p0; q[0] = q[0] + c0f.x; q[1] = q[1] + c0f.y; q[2] = q[2] + c0f.z; }
                }
                __syncthreads();
                if (axis == a && p1 >= 0) {
                    s_accd[p1] = s_accd[p1] + c1d;
                    if (HAS_DEF) { T *q = s_accf + 3 * p1; q[0] = q[0] + c1f.x; q[1] = q[1] + c1f.y; q[2] = q[2] + c1f.z; }
                }
                __syncthreads();
            }
        }
    }

    // ---- 5. dense write-out ----------------------------------------------------------------------------
    const bool zfull = c > 0 && 32 * c + 31 <= g.Z;   // every point of the chunk is a real grid point
    for (int r = wid; r < BC_X * BC_Y; r += BC_THREADS / 32) {
        const int ox = r / BC_Y, oy = r - ox * BC_Y;
        const int xp = xp0 + ox, yp = yp0 + oy;
        if (xp > g.X || yp > g.Y) continue;
        const long long rowb = ((long long)(xp - 1) * g.Y + (yp - 1)) * g.Z + (32 * c - 1);  // element of lane 0
        T *od = adj_sdf + rowb;
        T *of = adj_deform + 3 * rowb;
        if (zfull) {
            st_stream(od + lane, s_accd[r * 32 + lane]);
            if (HAS_DEF) {
                st_stream(of + lane, s_accf[r * 96 + lane]);
                st_stream(of + lane + 32, s_accf[r * 96 + lane + 32]);
                st_stream(of + lane + 64, s_accf[r * 96 + lane + 64]);
            }
        } else {
            const int zp = 32 * c + lane;
            if (zp >= 1 && zp <= g.Z) st_stream(od + lane, s_accd[r * 32 + lane]);
            if (HAS_DEF) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int e = lane + 32 * q;
                    const int zz = 32 * c + e / 3;
                    if (zz >= 1 && zz <= g.Z) st_stream(of + e, s_accf[r * 96 + e]);
                }
            }
        }
    }
}

// Dense launch: one CTA per block of the (chunk, y tile, x tile) grid.
template <typename T, bool HAS_DEF, int BC_X, int BC_Y>
__global__ void __launch_bounds__(BC_THREADS) mc_backward_compact_kernel(const T *__restrict__ sdf,
                                                                       const T *__restrict__ deform, Geo g, T iso,
                                                                       T padv, T ix, T iy, T iz,
                                                                       const uint4 *__restrict__ E,
                                                                       const T *__restrict__ gsrc,
                                                                       T *__restrict__ adj_sdf,
                                                                       T *__restrict__ adj_deform, int ntx, int nty, int flat)
{
    int c, ty, tx;
    if (flat) {   // degenerate shapes whose tile counts exceed the y / z grid limits
        int b = blockIdx.x;
        c = b % g.NC; b /= g.NC;
        ty = b % nty; tx = b / nty;
    } else {
        c = blockIdx.x; ty = blockIdx.y; tx = blockIdx.z;
    }
    mc_backward_block<T, HAS_DEF, BC_X, BC_Y, false>(sdf, deform, g, iso, padv, ix, iy, iz, E, gsrc, adj_sdf, adj_deform, tx, ty, c);
}

// ---- sparse surfaces ------------------------------------------------------------------------------
// On a smooth surface only a few percent of the blocks are touched by a crossing edge, and a grid of
// CTAs that each fetch their records only to find nothing is bound by that fetch's latency (sphere 512^3:
// 0.55 ms = 3.9 TB/s of zeros, where a plain fill reaches 7.4 TB/s).  Sparse path: the outputs are
// zero-filled by cudaMemsetAsync, bwd_mark lists the touched blocks (one thread per block, coalesced record
// reads, one atomic per warp), and a persistent grid pulls them from that list.
// work: u32 {count, cursor, pad...} in the first 64 bytes, then the block ids (flat: (tx nty + ty) NC + c).
template <int BC_X, int BC_Y>
__global__ void __launch_bounds__(256) bwd_mark_kernel(Geo g, const uint4 *__restrict__ E, int ntx, int nty,
                                                       unsigned *__restrict__ work)
{
    const long long nblk = (long long)ntx * nty * g.NC;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = false;
    if (b < nblk) {
        const int c = (int)(b % g.NC);
        const long long t = b / g.NC;
        const int ty = (int)(t % nty), tx = (int)(t / nty);
        const int xp0 = 1 + tx * BC_X, yp0 = 1 + ty * BC_Y;
        unsigned any = 0;
        for (int dxr = 0; dxr <= BC_X; ++dxr) {
            const int xp = xp0 - 1 + dxr;
            if (xp > g.X + 1) break;
            for (int dyr = 0; dyr <= BC_Y; ++dyr) {
                const int yp = yp0 - 1 + dyr;
                if (yp > g.Y + 1) break;
                const int k = (xp * g.PY + yp) * g.NC + c;
                const uint4 rec = __ldg(E + k);
                if (dyr >= 1) any |= rec.y;
                if (dxr >= 1) any |= rec.z;
                if (dxr >= 1 && dyr >= 1) {
                    any |= rec.w;
                    if (c > 0) any |= __ldg(&E[k - 1].w) >> 31;
                }
            }
        }
        active = any != 0u;
    }
    const unsigned m = __ballot_sync(FULL, active);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&work[0], (unsigned)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (active) work[16 + base + __popc(m & lanemask_lt(lane))] = (unsigned)b;
    }
}

template <typename T, bool HAS_DEF, int BC_X, int BC_Y>
__global__ void __launch_bounds__(BC_THREADS) mc_backward_queue_kernel(const T *__restrict__ sdf,
                                                                     const T *__restrict__ deform, Geo g, T iso,
                                                                     T padv, T ix, T iy, T iz,
                                                                     const uint4 *__restrict__ E,
                                                                     const T *__restrict__ gsrc,
                                                                     T *__restrict__ adj_sdf,
                                                                     T *__restrict__ adj_deform, int nty,
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
        mc_backward_block<T, HAS_DEF, BC_X, BC_Y, true>(sdf, deform, g, iso, padv, ix, iy, iz, E, gsrc, adj_sdf, adj_deform, tx, ty, c);
        __syncthreads();   // shared memory (and s_next) is reused by the next block
    }
}

template <typename T, bool HAS_DEF, int BC_X, int BC_Y> constexpr size_t bwd_compact_smem()
{
    constexpr int BC_ROWS = (BC_X + 1) * (BC_Y + 1), BC_PTS = BC_X * BC_Y * 32;
    constexpr int BC_CAP = (BC_X + 1) * BC_Y * 32 + BC_X * (BC_Y + 1) * 32 + BC_X * BC_Y * 33;  // worst-case list length
    return (size_t)BC_PTS * (HAS_DEF ? 4 : 1) * sizeof(T) + BC_ROWS * (sizeof(uint4) + sizeof(int4)) + (3 * BC_ROWS + 1 + BC_ROWS) * 4 +
           (size_t)BC_CAP * 2 + 16;
}

}  // namespace diso
