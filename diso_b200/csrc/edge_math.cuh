// edge_math.cuh -- the floating-point core: iso-crossing position on a grid edge and its adjoint.
//
// Arithmetic order follows the reference expression by expression so that vertices are
// bit-identical (the library is compiled with -fmad=false; the single multiply-add that nvcc's
// default contraction fuses in the reference, p0 + (p1-p0)*t at cumc.cu:367, is an explicit FMA):
//   forward  computeMcVert     cumc.cu:343-368   (== cudualmc.cu:683-708)
//   adjoint  adjComputeMcVert  cumc.cu:412-453   (== cudualmc.cu:710-751)
#pragma once
#include "common.cuh"

namespace diso {

template <typename T> struct Vec3 { T x, y, z; };

__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

// clamp(x, 0, 1) with the reference's select order (cumc.cu:82-91): min(max(0, x), 1)
template <typename T> __device__ __forceinline__ T clamp01(T x)
{
    T m = (T(0) > x) ? T(0) : x;
    return (m < T(1)) ? m : T(1);
}

template <typename T> __device__ __forceinline__ T edge_t(T d0, T d1, T iso)
{
    return (d1 != d0) ? clamp01((iso - d0) / (d1 - d0)) : T(0.5);
}

// AoS xyz fetch of the deformation at padded coords; the pad layer carries zero deformation
// (diso/__init__.py:54).
template <typename T>
__device__ __forceinline__ Vec3<T> fetch_deform(const T *__restrict__ deform, const Geo &g, int xp, int yp, int zp)
{
    int x = xp - 1, y = yp - 1, z = zp - 1;
    Vec3<T> r{T(0), T(0), T(0)};
    if ((unsigned)x < (unsigned)g.X && (unsigned)y < (unsigned)g.Y && (unsigned)z < (unsigned)g.Z) {
        const T *p = deform + (((size_t)x * g.Y + y) * g.Z + z) * 3;
        r.x = __ldg(p); r.y = __ldg(p + 1); r.z = __ldg(p + 2);
    }
    return r;
}

// Crossing point of the edge from padded point (xp,yp,zp) along +AXIS, in the PADDED frame.
template <typename T, int AXIS>
__device__ __forceinline__ Vec3<T> edge_vertex(T d0, T d1, T iso, int xp, int yp, int zp, bool has_def,
                                               const Vec3<T> &f0, const Vec3<T> &f1)
{
    const T t = edge_t(d0, d1, iso);
    Vec3<T> p0{T(xp), T(yp), T(zp)};
    Vec3<T> p1{T(xp + (AXIS == 0)), T(yp + (AXIS == 1)), T(zp + (AXIS == 2))};
    if (has_def) {
        p0.x = p0.x + f0.x; p0.y = p0.y + f0.y; p0.z = p0.z + f0.z;
        p1.x = p1.x + f1.x; p1.y = p1.y + f1.y; p1.z = p1.z + f1.z;
    }
    Vec3<T> r;
    r.x = fma_rn(p1.x - p0.x, t, p0.x);
    r.y = fma_rn(p1.y - p0.y, t, p0.y);
    r.z = fma_rn(p1.z - p0.z, t, p0.z);
    return r;
}

// API-frame epilogue of diso/__init__.py:56-60 fused into the emit: (p - 1) [/ (dim - 1)]
template <typename T> struct Epilogue {
    T dx, dy, dz;  // (T)dim - 1
    bool normalize;
    __device__ __forceinline__ Vec3<T> apply(Vec3<T> p) const
    {
        p.x = p.x - T(1); p.y = p.y - T(1); p.z = p.z - T(1);
        if (normalize) { p.x = p.x / dx; p.y = p.y / dy; p.z = p.z / dz; }
        return p;
    }
    // chain rule for an incoming adjoint (autograd of `verts / d` is `grad / d`)
    __device__ __forceinline__ Vec3<T> adjoint(Vec3<T> a) const
    {
        if (normalize) { a.x = a.x / dx; a.y = a.y / dy; a.z = a.z / dz; }
        return a;
    }
};

// Adjoint of edge_vertex w.r.t. ONE endpoint of the edge (gather formulation: the thread that
// owns grid point `which` (0 = start point, 1 = end point) calls this for every incident
// crossing edge and sums the results in a fixed order -> no atomics, deterministic).
//   adj_t  = (p1 - p0) . g                       cumc.cu:436
//   adj_d0 = (iso - d1) / (d1 - d0)^2 * adj_t    cumc.cu:448
//   adj_d1 = (d0 - iso) / (d1 - d0)^2 * adj_t    cumc.cu:449
//   adj_p0 = (1 - t) g ; adj_p1 = t g            cumc.cu:434-435
template <typename T, int AXIS>
__device__ __forceinline__ void edge_adjoint(T d0, T d1, T iso, int xp, int yp, int zp, bool has_def,
                                             const Vec3<T> &f0, const Vec3<T> &f1, const Vec3<T> &gv, int which,
                                             T &acc_d, Vec3<T> &acc_f)
{
    const T t = edge_t(d0, d1, iso);
    Vec3<T> p0{T(xp), T(yp), T(zp)};
    Vec3<T> p1{T(xp + (AXIS == 0)), T(yp + (AXIS == 1)), T(zp + (AXIS == 2))};
    if (has_def) {
        p0.x = p0.x + f0.x; p0.y = p0.y + f0.y; p0.z = p0.z + f0.z;
        p1.x = p1.x + f1.x; p1.y = p1.y + f1.y; p1.z = p1.z + f1.z;
    }
    T adj_t = (p1.x - p0.x) * gv.x;
    adj_t = adj_t + (p1.y - p0.y) * gv.y;
    adj_t = adj_t + (p1.z - p0.z) * gv.z;
    const T den = (d1 - d0) * (d1 - d0);
    if (which == 0) {
        acc_d = acc_d + (iso - d1) / den * adj_t;
        const T w = T(1) - t;
        acc_f.x = acc_f.x + w * gv.x; acc_f.y = acc_f.y + w * gv.y; acc_f.z = acc_f.z + w * gv.z;
    } else {
        acc_d = acc_d + (d0 - iso) / den * adj_t;
        acc_f.x = acc_f.x + t * gv.x; acc_f.y = acc_f.y + t * gv.y; acc_f.z = acc_f.z + t * gv.z;
    }
}

}  // namespace diso
