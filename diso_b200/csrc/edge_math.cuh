// edge_math.cuh -- floating-point building blocks of the iso-crossing on a grid edge.
//
// Arithmetic order follows the reference expression by expression so that vertices are
// bit-identical (the library is compiled with -fmad=false; the single multiply-add that nvcc's
// default contraction fuses in the reference, p0 + (p1-p0)*t at cumc.cu:367, is an explicit FMA
// in compact.cuh:edge_verts_kernel):
//   forward  computeMcVert     cumc.cu:343-368   (== cudualmc.cu:683-708)  -> edge_verts_kernel
//   adjoint  adjComputeMcVert  cumc.cu:412-453   (== cudualmc.cu:710-751)  -> mc_backward_compact_kernel
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

}  // namespace diso
