// common.cuh -- geometry, state layout and warp/scan primitives shared by all kernels.
//
// Data layout in HBM:
//   * The reference pads the grid by one layer of "iso+1" (diso/__init__.py:52) and works on
//     the padded (X+2,Y+2,Z+2) lattice with linear index z + PZ*(y + PY*x) (cumc.h:101-105).
//     We keep that lattice as the index space (so every ordering is identical to the
//     reference's) but never materialise it: the pad is applied virtually.
//   * Every padded z-row (xp,yp) is cut into NC = ceil(PZ/32) "chunks" of 32 consecutive
//     points; chunk id k = (xp*PY + yp)*NC + c.  All per-chunk metadata is a handful of
//     32-bit lane masks, so a warp (lane == point) or a single thread (bit-parallel) can
//     process a chunk.
//   * sign word  S[k]   : bit j = (value(point j of chunk k) >= iso); pad / beyond-row bits = 1.
//   * edge record E[k]  : {base, mx, my, mz}: mx/my/mz bit j = the +x/+y/+z edge owned by
//                         point j crosses the iso level; base = number of crossing edges
//                         owned by all points before chunk k (== id of the first MC vertex /
//                         DMC quad of the chunk, reference order: point-major, axis-minor).
//   * MC  : F[k]        : {base, used}: id of the first triangle emitted by the cells of chunk
//                         k; used bit j = cell j is a used cell.
//   * DMC : P[k]        : {base, used, lo, hi}: base = id of the first dual vertex of the
//                         chunk; (lo,hi) bit j = (patch count - 1) of cell j.
//   * C[32k + j]        : per USED cell, 16 bits: case index (DMC: after the ambiguity flip,
//                         cudualmc.cu:815-839) | (id of the cell's first triangle / dual vertex
//                         - base of its chunk) << 8.  Entries of unused cells are never written
//                         nor read (consumers mask with `used`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/diso_b200.h"

namespace diso {

constexpr unsigned FULL = 0xffffffffu;

struct Geo {
    int X, Y, Z;     // unpadded dims
    int PX, PY, PZ;  // padded dims
    int NC;          // chunks per padded row
    int NR;          // padded rows = PX*PY
    int NCH;         // chunks = NR*NC
    int sY, sX;      // chunk-id stride of +1 in yp (= NC) and in xp (= PY*NC)
};

inline Geo make_geo(int X, int Y, int Z)
{
    Geo g;
    g.X = X; g.Y = Y; g.Z = Z;
    g.PX = X + 2; g.PY = Y + 2; g.PZ = Z + 2;
    g.NC = (g.PZ + 31) / 32;
    g.NR = g.PX * g.PY;
    g.NCH = g.NR * g.NC;
    g.sY = g.NC;
    g.sX = g.PY * g.NC;
    return g;
}

// One descriptor per scan tile (decoupled look-back, see classify.cuh).
// Four quantities are scanned together: a = crossing edges, b = triangles (MC) / dual vertices (DMC),
// c = chunks owning >= 1 crossing edge, d = chunks with >= 1 face (the last two give the ordered
// active-chunk lists the emit kernels iterate over).
struct __align__(64) TileDesc {
    unsigned flag;            // 0 = empty, 1 = aggregate available, 2 = inclusive prefix available
    unsigned pad0;
    unsigned long long agg;   // the tile's four aggregates, 16 bits each (a | b<<16 | c<<32 | d<<48)
    unsigned long long incl[4];
};

#ifndef DISO_SCAN_TILE
#define DISO_SCAN_TILE 256
#endif
constexpr int SCAN_TILE = DISO_SCAN_TILE;  // chunks per scan tile == threads per classify CTA
#ifndef DISO_BWD_BX
#define DISO_BWD_BX 4
#define DISO_BWD_BY 6
#endif
constexpr int BWD_BX = DISO_BWD_BX, BWD_BY = DISO_BWD_BY;  // rows per backward block (x, y); one 32-point chunk in z
#ifndef DISO_BWD2_BX
#define DISO_BWD2_BX 4
#define DISO_BWD2_BY 6
#endif
constexpr int BWD2_BX = DISO_BWD2_BX, BWD2_BY = DISO_BWD2_BY;  // same for the saved-record backward (mc_backward_v2.cuh)

// Byte offsets of the arrays inside the caller-owned state buffer.
struct StateLayout {
    size_t off_counts;   // int64[DISO_COUNT_SLOTS]
    size_t off_ticket;   // u32 ticket + padding (64 B)
    size_t off_desc;     // TileDesc[n_tiles]
    size_t off_sign;     // u32[NCH + sign_tail]
    size_t off_erec;     // uint4[NCH + rec_tail]
    size_t off_aux;      // MC: uint2 F[NCH + 8] ; DMC: uint4 P[NCH + rec_tail]
    size_t off_cell;     // u16 C[(NCH + 8) * 32] per-cell {case index | offset of first triangle / dual vertex << 8}
    size_t off_active;   // u32 active-chunk lists (ascending chunk id): [0, NCH) chunks owning crossing
                         // edges, [NCH, 2 NCH) chunks with faces
    size_t total;
    int n_tiles;
    int sign_tail;
    int rec_tail;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline StateLayout make_layout(int alg, const Geo &g)
{
    StateLayout L;
    L.n_tiles = (g.NCH + 1 + SCAN_TILE - 1) / SCAN_TILE;  // +1: slot NCH carries the totals
    L.sign_tail = g.sX + g.sY + 8;
    L.rec_tail = g.sX + g.sY + 8;
    size_t o = 0;
    L.off_counts = o; o += 64;
    L.off_ticket = o; o += 64;
    L.off_desc = o;   o += (size_t)L.n_tiles * sizeof(TileDesc);
    o = align_up(o, 256);
    L.off_sign = o;   o += ((size_t)g.NCH + L.sign_tail) * 4;
    o = align_up(o, 256);
    L.off_erec = o;   o += ((size_t)g.NCH + L.rec_tail) * 16;
    o = align_up(o, 256);
    L.off_aux = o;
    if (alg == DISO_ALG_MC) o += ((size_t)g.NCH + 8) * 8;
    else                    o += ((size_t)g.NCH + L.rec_tail) * 16;
    o = align_up(o, 256);
    L.off_cell = o;
    o += ((size_t)g.NCH + 8) * 32 * 2;
    o = align_up(o, 256);
    L.off_active = o;
    o += (size_t)g.NCH * 2 * 4;
    o = align_up(o, 256);
    L.total = align_up(o, 256);
    return L;
}

// ---- memory-ordering helpers (tile descriptors) -------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- small bit helpers -----------------------------------------------------------------
__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }
__device__ __forceinline__ int bit(unsigned w, int j) { return (int)((w >> j) & 1u); }
// word shifted down by one position, pulling bit 0 of the next chunk into bit 31
__device__ __forceinline__ unsigned shift_in(unsigned w, unsigned next) { return __funnelshift_r(w, next, 1); }

// streaming (evict-first) stores for write-once outputs
template <typename T> __device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }

// virtual-pad aware value fetch at padded coordinates (xp,yp,zp)
template <typename T>
__device__ __forceinline__ T fetch_padded(const T *__restrict__ sdf, const Geo &g, int xp, int yp, int zp, T padv)
{
    int x = xp - 1, y = yp - 1, z = zp - 1;
    bool in = (unsigned)x < (unsigned)g.X && (unsigned)y < (unsigned)g.Y && (unsigned)z < (unsigned)g.Z;
    return in ? __ldg(sdf + ((size_t)x * g.Y + y) * g.Z + z) : padv;
}

}  // namespace diso
