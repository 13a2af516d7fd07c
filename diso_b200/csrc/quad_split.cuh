// quad_split.cuh -- DMC quad -> triangle split (diso/__init__.py:118-147) as three small kernels.
//
// The reference does this with ~60 PyTorch ops and dozens of [Q,3] temporaries.  Here:
//   Q1 quad_diag_kernel : one thread per quad gathers its 4 vertices, evaluates both diagonals
//                         (max cosine over the 2x3 triangle angles, cos via x / max(|x|, 1e-12)
//                         like F.normalize) and records flag = (angles1 < angles2) plus a
//                         per-tile population count;
//   Q2 tile_scan_kernel : one CTA turns the tile counts into exclusive offsets (+ total n1);
//   Q3 quad_emit_kernel : one thread per quad writes its two triangles at the position the
//                         reference's boolean-mask + cat produces: config-1 quads first
//                         ([0,1,3],[1,2,3]), then config-2 quads ([0,1,2],[0,2,3]), each group
//                         in quad order.
// Ties (angles1 == angles2) go to config 2, as in the reference (`<`, then `>=`).
#pragma once
#include "common.cuh"
#include "edge_math.cuh"

namespace diso {

constexpr int QS_TILE = 256;

__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }
__device__ __forceinline__ double sqrt_rn(double x) { return __dsqrt_rn(x); }

template <typename T> __device__ __forceinline__ Vec3<T> unit(const Vec3<T> &a, const Vec3<T> &b)
{
    // F.normalize(a - b): v / max(||v||_2, 1e-12)
    Vec3<T> v{a.x - b.x, a.y - b.y, a.z - b.z};
    T n2 = v.x * v.x;
    n2 = fma_rn(v.y, v.y, n2);
    n2 = fma_rn(v.z, v.z, n2);
    T n = sqrt_rn(n2);
    n = n > T(1e-12) ? n : T(1e-12);
    return Vec3<T>{v.x / n, v.y / n, v.z / n};
}
template <typename T> __device__ __forceinline__ T dot3(const Vec3<T> &a, const Vec3<T> &b)
{
    T s = a.x * b.x;
    s = s + a.y * b.y;
    s = s + a.z * b.z;
    return s;
}
template <typename T> __device__ __forceinline__ T tri_max_cos(const Vec3<T> &v0, const Vec3<T> &v1, const Vec3<T> &v2)
{
    const T c1 = dot3(unit(v1, v0), unit(v2, v0));
    const T c2 = dot3(unit(v2, v1), unit(v0, v1));
    const T c3 = dot3(unit(v0, v2), unit(v1, v2));
    T m = c1 > c2 ? c1 : c2;
    return m > c3 ? m : c3;
}

template <typename T>
__global__ void __launch_bounds__(QS_TILE) quad_diag_kernel(const T *__restrict__ verts, const long long *__restrict__ quads,
                                                          long long nq, unsigned char *__restrict__ flags,
                                                          unsigned *__restrict__ tile_cnt)
{
    const long long q = (long long)blockIdx.x * QS_TILE + threadIdx.x;
    bool f = false;
    if (q < nq) {
        const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q));
        const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q) + 1);
        Vec3<T> v[4];
        const long long id[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const T *p = verts + id[i] * 3;
            v[i] = Vec3<T>{__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        }
        const T t13a = tri_max_cos(v[0], v[1], v[3]), t13b = tri_max_cos(v[1], v[2], v[3]);
        const T t02a = tri_max_cos(v[0], v[1], v[2]), t02b = tri_max_cos(v[0], v[2], v[3]);
        const T a1 = t13a > t13b ? t13a : t13b;
        const T a2 = t02a > t02b ? t02a : t02b;
        f = a1 < a2;
        flags[q] = f ? 1 : 0;
    }
    const int c = __syncthreads_count(f);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = (unsigned)c;
}

// exclusive scan of n u32 values in place by ONE CTA; total -> *total_out (u64)
__global__ void __launch_bounds__(1024) tile_scan_kernel(unsigned *__restrict__ v, int n, unsigned long long *__restrict__ total_out)
{
    __shared__ unsigned long long s_part[1024];
    const int tid = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    unsigned long long s = 0;
    for (int i = lo; i < hi; ++i) s += v[i];
    s_part[tid] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long t = tid >= d ? s_part[tid - d] : 0ull;
        __syncthreads();
        s_part[tid] += t;
        __syncthreads();
    }
    unsigned long long run = s_part[tid] - s;
    for (int i = lo; i < hi; ++i) {
        const unsigned c = v[i];
        v[i] = (unsigned)run;  // < 2^32: callers cap n_quads
        run += c;
    }
    if (tid == 1023) *total_out = s_part[1023];
}

__global__ void __launch_bounds__(QS_TILE) quad_emit_kernel(const long long *__restrict__ quads, long long nq,
                                                          const unsigned char *__restrict__ flags,
                                                          const unsigned *__restrict__ tile_off,
                                                          const unsigned long long *__restrict__ total,
                                                          long long *__restrict__ faces)
{
    __shared__ unsigned s_w[QS_TILE / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long q = (long long)blockIdx.x * QS_TILE + tid;
    const bool f = q < nq && flags[q];
    const unsigned bal = __ballot_sync(FULL, f);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    unsigned before = tile_off[blockIdx.x] + __popc(bal & lanemask_lt(lane));
    for (int i = 0; i < wid; ++i) before += s_w[i];
    if (q >= nq) return;
    const long long n1 = (long long)*total;
    const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q));
    const longlong2 b = __ldg(reinterpret_cast<const longlong2 *>(quads + 4 * q) + 1);
    const long long pos = f ? (long long)before : n1 + (q - (long long)before);
    longlong2 *dst = reinterpret_cast<longlong2 *>(faces + pos * 6);
    if (f) {  // [0,1,3] [1,2,3]
        __stcs(dst, make_longlong2(a.x, a.y));
        __stcs(dst + 1, make_longlong2(b.y, a.y));
        __stcs(dst + 2, make_longlong2(b.x, b.y));
    } else {  // [0,1,2] [0,2,3]
        __stcs(dst, make_longlong2(a.x, a.y));
        __stcs(dst + 1, make_longlong2(b.x, a.x));
        __stcs(dst + 2, make_longlong2(b.x, b.y));
    }
}

}  // namespace diso
