// quad_split.cuh -- DMC quad -> triangle split (diso/__init__.py:118-147) as two kernels.
//
// The reference does this with ~60 PyTorch ops and dozens of [Q,3] temporaries.  Here:
//   Q1 quad_diag_kernel : one thread per quad gathers its 4 vertices, evaluates both diagonals
//                         (max cosine over the 2x3 triangle angles, cos via x / max(|x|, 1e-12)
//                         like F.normalize) and records flag = (angles1 < angles2); the per-tile
//                         counts are scanned in the same pass (decoupled look-back);
//   Q2 quad_emit_kernel : one thread per quad writes its two triangles at the position the
//                         reference's boolean-mask + cat produces: config-1 quads first
//                         ([0,1,3],[1,2,3]), then config-2 quads ([0,1,2],[0,2,3]), each group
//                         in quad order.
// Ties (angles1 == angles2) go to config 2, as in the reference (`<`, then `>=`).
#pragma once
#include "common.cuh"
#include "edge_math.cuh"

namespace diso {

constexpr int QS_THREADS = 256;
constexpr int QS_PER = 1;                       // quads per thread
constexpr int QS_TILE = QS_THREADS * QS_PER;    // quads per tile; quad (i, tid) of tile t is t*QS_TILE + i*QS_THREADS + tid

__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ double sqrt_rn(double x) { return __dsqrt_rn(x); }

// Summation order of torch's CUDA reductions over a row of 3 (measured on B200 with tools/probe_torch_reduce.py,
// torch 2.11: the reduce kernel gives the row to two threads -- elements {0, 2} and {1} -- and combines them with
// one shuffle): vector_norm = sqrt((x^2 + z^2) + y^2) and sum = (p0 + p2) + p1, every square / product rounded on its
// own (no FMA).  Reproducing that order makes every cosine of the reference bit-identical, so the diagonal choice
// agrees on ALL quads, ties included (35 494 of 72.5 M quads flipped at 512^3 with the x,y,z FMA chain used before).
template <typename T> __device__ __forceinline__ Vec3<T> unit(const Vec3<T> &a, const Vec3<T> &b)
{
    // F.normalize(a - b): v / max(||v||_2, 1e-12)   (diso/__init__.py:126-128)
    Vec3<T> v{a.x - b.x, a.y - b.y, a.z - b.z};
    const T n2 = (v.x * v.x + v.z * v.z) + v.y * v.y;
    T n = sqrt_rn(n2);
    n = n > T(1e-12) ? n : T(1e-12);
    return Vec3<T>{v.x / n, v.y / n, v.z / n};
}
template <typename T> __device__ __forceinline__ T dot3(const Vec3<T> &a, const Vec3<T> &b)
{
    return (a.x * b.x + a.z * b.z) + a.y * b.y;
}
template <typename T> __device__ __forceinline__ T max3(T a, T b, T c)
{
    const T m = a > b ? a : b;   // the reference's max over the three angles, same select order
    return m > c ? m : c;
}

// One descriptor per tile for the decoupled look-back over the per-tile counts of config-1 quads
// (same protocol as classify_scan: ticketed tile order, acquire / release flags).
struct __align__(16) QuadTileDesc {
    unsigned flag;   // 0 = empty, 1 = aggregate available, 2 = inclusive prefix available
    unsigned agg;
    unsigned long long incl;
};

// The reference's choice for one quad (diso/__init__.py:118-147): true = config 1 ([0,1,3],[1,2,3]), i.e. angles1 < angles2.
template <typename T>
__device__ __forceinline__ bool quad_first_diagonal(const Vec3<T> &v0, const Vec3<T> &v1, const Vec3<T> &v2, const Vec3<T> &v3)
{
    const Vec3<T> S0 = unit(v1, v0), S1 = unit(v2, v1), S2 = unit(v3, v2), S3 = unit(v0, v3);
    const Vec3<T> D0 = unit(v2, v0), D1 = unit(v3, v1);
    const T t13a = max3(-dot3(S0, S3), -dot3(D1, S0), -dot3(S3, D1));
    const T t13b = max3(dot3(S1, D1), -dot3(S2, S1), dot3(D1, S2));
    const T t02a = max3(dot3(S0, D0), -dot3(S1, S0), dot3(D0, S1));
    const T t02b = max3(-dot3(D0, S3), -dot3(S2, D0), -dot3(S3, S2));
    const T a1 = t13a > t13b ? t13a : t13b;
    const T a2 = t02a > t02b ? t02a : t02b;
    return a1 < a2;
}

// Q1: diagonal choice + exclusive prefix of the config-1 counts in one pass.
// (When dmc_emit_quads already wrote the flags -- DiffDMC's default path -- quad_scan_kernel below replaces this kernel.)
//
// The reference evaluates 4 triangles x 3 angles, each from two freshly normalised edge vectors
// (24 normalisations: 24 square roots, 72 divisions per quad).  Only six distinct directions exist -- the
// four sides S0 = 0->1, S1 = 1->2, S2 = 2->3, S3 = 3->0 and the diagonals D0 = 0->2, D1 = 1->3 -- and
// IEEE arithmetic is sign-symmetric (fl(b-a) = -fl(a-b), x/n and the products of the dot negate exactly),
// so every cosine of the reference is +-dot(.,.) of two of those six unit vectors, bit for bit:
//   tri(0,1,3): -S0.S3, -D1.S0, -S3.D1      tri(1,2,3):  S1.D1, -S2.S1,  D1.S2
//   tri(0,1,2):  S0.D0, -S1.S0,  D0.S1      tri(0,2,3): -D0.S3, -S2.D0, -S3.S2
// The kernel is bound by the latency of its two dependent loads (quad ids -> vertex gathers; ncu: 40 % issue
// activity, nothing else above 35 %), so resident warps are what counts: 8 CTAs/SM (32 registers, a few spilled
// bytes) 2.15 ms at 512^3 vs 2.44 ms at the compiler's 40 registers; several quads per thread cost registers
// and were slower (2.75 / 3.15 ms for 2 / 4).
template <typename T>
__global__ void __launch_bounds__(QS_THREADS, sizeof(T) == 4 ? 8 : 3) quad_diag_kernel(const T *__restrict__ verts, const long long *__restrict__ quads,
                                                             long long nq, unsigned char *__restrict__ flags,
                                                             unsigned *__restrict__ tile_off, QuadTileDesc *__restrict__ desc,
                                                             unsigned *__restrict__ ticket, unsigned long long *__restrict__ total)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) { s_tile = atomicAdd(ticket, 1u); s_cnt = 0u; }
    __syncthreads();
    const unsigned tile = s_tile;
    const long long q0 = (long long)tile * QS_TILE + threadIdx.x;
    // all index loads first, then all vertex gathers: QS_PER independent dependent-load chains per thread
    // (with one quad per thread the kernel was bound by the latency of that chain: ncu showed 40 % issue
    // activity and nothing else above 35 %)
    longlong2 qa[QS_PER], qb[QS_PER];
#pragma unroll
    for (int i = 0; i < QS_PER; ++i) {
        const long long q = q0 + (long long)i * QS_THREADS;
        qa[i] = qb[i] = make_longlong2(0, 0);
        if (q < nq) {
            qa[i] = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q));
            qb[i] = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q) + 1);
        }
    }
    Vec3<T> v[QS_PER][4];
#pragma unroll
    for (int i = 0; i < QS_PER; ++i) {
        const long long id[4] = {qa[i].x, qa[i].y, qb[i].x, qb[i].y};   // (out-of-range quads read vertex 0)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const T *p = verts + id[c] * 3;
            v[i][c] = Vec3<T>{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        }
    }
    unsigned mine = 0;
#pragma unroll
    for (int i = 0; i < QS_PER; ++i) {
        const long long q = q0 + (long long)i * QS_THREADS;
        const bool f = q < nq && quad_first_diagonal(v[i][0], v[i][1], v[i][2], v[i][3]);
        if (q < nq) flags[q] = f ? 1 : 0;
        mine += f ? 1u : 0u;
    }
    mine = __reduce_add_sync(FULL, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const unsigned cnt = s_cnt;
    // ---- decoupled look-back over the tile counts (warp 0) ------------------------------------
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) { st_relaxed_u64(&desc[0].incl, cnt); st_release_u32(&desc[0].flag, 2u); }
        } else {
            if (lane == 0) { desc[tile].agg = cnt; st_release_u32(&desc[tile].flag, 1u); }
            int base = (int)tile - 1;
            while (true) {
                const int t = base - lane;
                unsigned fl = 2u;
                unsigned long long x = 0;
                if (t >= 0) {
                    do { fl = ld_acquire_u32(&desc[t].flag); } while (fl == 0u);
                    x = fl == 2u ? ld_relaxed_u64(&desc[t].incl) : (unsigned long long)*reinterpret_cast<volatile unsigned *>(&desc[t].agg);
                }
                const unsigned m = __ballot_sync(FULL, fl == 2u);
                const int stop = m ? (__ffs(m) - 1) : 32;   // nearest predecessor with an inclusive prefix
                if (lane > stop) x = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                excl += x;
                if (m) break;
                base -= 32;
            }
            if (lane == 0) { st_relaxed_u64(&desc[tile].incl, excl + cnt); st_release_u32(&desc[tile].flag, 2u); }
        }
        if (lane == 0) {
            tile_off[tile] = (unsigned)excl;   // < 2^32: callers cap n_quads
            if ((long long)(tile + 1) * QS_TILE >= nq) *total = excl + cnt;   // last tile: number of config-1 quads
        }
    }
}

// Q1': the scan alone, for flags that dmc_emit_quads already wrote (DiffDMC's default path).  A CTA owns 4096 quads
// (one 128-bit load of 16 flag bytes per thread) = 16 of quad_emit's 256-quad tiles, so the ticket / look-back
// machinery runs 16x less often than in quad_diag (283 k tiles at 512^3 cost 1.18 ms when nothing hides them).
constexpr int QSC_PER = 16;
constexpr int QSC_TILE = QS_THREADS * QSC_PER;
static_assert(QS_TILE == 16 * QSC_PER && QS_PER == 1, "16 threads of the scan cover one emit tile");

__global__ void __launch_bounds__(QS_THREADS) quad_scan_kernel(const unsigned char *__restrict__ flags, long long nq,
                                                             unsigned *__restrict__ tile_off, QuadTileDesc *__restrict__ desc,
                                                             unsigned *__restrict__ ticket, unsigned long long *__restrict__ total)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned s_warp[QS_THREADS / 32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const long long q0 = (long long)tile * QSC_TILE + (long long)tid * QSC_PER;
    unsigned c = 0;
    if (q0 + QSC_PER <= nq) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(flags + q0));
        c = __popc(w.x & 0x01010101u) + __popc(w.y & 0x01010101u) + __popc(w.z & 0x01010101u) + __popc(w.w & 0x01010101u);
    } else {
        for (long long q = q0; q < nq; ++q) c += flags[q] != 0;
    }
    unsigned inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned before = inc - c, cnt = 0;
#pragma unroll
    for (int w = 0; w < QS_THREADS / 32; ++w) { const unsigned t = s_warp[w]; if (w < wid) before += t; cnt += t; }
    // ---- decoupled look-back over the 4096-quad tiles (warp 0), same protocol as quad_diag ----------------------
    if (tid < 32) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) { st_relaxed_u64(&desc[0].incl, cnt); st_release_u32(&desc[0].flag, 2u); }
        } else {
            if (lane == 0) { desc[tile].agg = cnt; st_release_u32(&desc[tile].flag, 1u); }
            int base = (int)tile - 1;
            while (true) {
                const int t = base - lane;
                unsigned fl = 2u;
                unsigned long long x = 0;
                if (t >= 0) {
                    do { fl = ld_acquire_u32(&desc[t].flag); } while (fl == 0u);
                    x = fl == 2u ? ld_relaxed_u64(&desc[t].incl) : (unsigned long long)*reinterpret_cast<volatile unsigned *>(&desc[t].agg);
                }
                const unsigned m = __ballot_sync(FULL, fl == 2u);
                const int stop = m ? (__ffs(m) - 1) : 32;
                if (lane > stop) x = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(FULL, x, d);
                excl += x;
                if (m) break;
                base -= 32;
            }
            if (lane == 0) { st_relaxed_u64(&desc[tile].incl, excl + cnt); st_release_u32(&desc[tile].flag, 2u); }
        }
        if (lane == 0) {
            s_excl = excl;
            if ((long long)(tile + 1) * QSC_TILE >= nq) *total = excl + cnt;   // last tile: number of config-1 quads
        }
    }
    __syncthreads();
    // thread 16 s starts emit tile 16 tile + s
    if ((tid & 15) == 0) {
        const long long et = (long long)tile * 16 + (tid >> 4);
        if (et * QS_TILE < nq) tile_off[et] = (unsigned)(s_excl + before);
    }
}

__global__ void __launch_bounds__(QS_THREADS) quad_emit_kernel(const long long *__restrict__ quads, long long nq,
                                                             const unsigned char *__restrict__ flags,
                                                             const unsigned *__restrict__ tile_off,
                                                             const unsigned long long *__restrict__ total,
                                                             long long *__restrict__ faces)
{
    __shared__ unsigned s_w[QS_PER][QS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long q0 = (long long)blockIdx.x * QS_TILE + tid;
    bool f[QS_PER];
    unsigned bal[QS_PER];
#pragma unroll
    for (int i = 0; i < QS_PER; ++i) {
        const long long q = q0 + (long long)i * QS_THREADS;
        f[i] = q < nq && flags[q];
        bal[i] = __ballot_sync(FULL, f[i]);
        if (lane == 0) s_w[i][wid] = __popc(bal[i]);
    }
    __syncthreads();
    const long long n1 = (long long)*total;
    unsigned run = tile_off[blockIdx.x];   // config-1 quads before this tile; quads are ranked in (i, tid) order == quad order
#pragma unroll
    for (int i = 0; i < QS_PER; ++i) {
        unsigned before = run + __popc(bal[i] & lanemask_lt(lane));
#pragma unroll
        for (int w = 0; w < QS_THREADS / 32; ++w) {
            const unsigned c = s_w[i][w];
            if (w < wid) before += c;
            run += c;
        }
        const long long q = q0 + (long long)i * QS_THREADS;
        if (q >= nq) continue;
        const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q));
        const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q) + 1);
        const long long pos = f[i] ? (long long)before : n1 + (q - (long long)before);
        longlong2 *dst = reinterpret_cast<longlong2 *>(faces + pos * 6);
        if (f[i]) {  // [0,1,3] [1,2,3]
            __stcs(dst, make_longlong2(a.x, a.y));
            __stcs(dst + 1, make_longlong2(b.y, a.y));
            __stcs(dst + 2, make_longlong2(b.x, b.y));
        } else {  // [0,1,2] [0,2,3]
            __stcs(dst, make_longlong2(a.x, a.y));
            __stcs(dst + 1, make_longlong2(b.x, a.x));
            __stcs(dst + 2, make_longlong2(b.x, b.y));
        }
    }
}

}  // namespace diso
