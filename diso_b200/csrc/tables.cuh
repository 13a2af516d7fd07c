// tables.cuh -- device copies of the packed case tables (generated data, see
// tools/extract_tables.py for the encodings and the reference file:line of each source table).
// Kernels copy what they need into shared memory at CTA start; the tables are never indexed
// with divergent addresses through the constant cache.
#pragma once
#define DISO_TABLE_QUAL static __device__ const
#include "case_tables.inc"
#undef DISO_TABLE_QUAL

namespace diso {
// Owner-point offset and axis of local cell edge e (reference mcEdgeLocations, cumc.cu:109-122),
// as compile-time functions: bit-packed {dx,dy,dz} and axis per edge id.
//   e : 0  1  2  3  4  5  6  7  8  9 10 11
//  dx : 0  1  0  0  0  1  0  0  0  1  1  0
//  dy : 0  0  0  0  1  1  1  1  0  0  0  0
//  dz : 0  0  1  0  0  0  1  0  0  0  1  1
//  ax : 0  2  0  2  0  2  0  2  1  1  1  1
constexpr unsigned EDGE_DX = 0x622u;   // bits e: 1,5,9,10
constexpr unsigned EDGE_DY = 0x0f0u;   // bits e: 4,5,6,7
constexpr unsigned EDGE_DZ = 0xc44u;   // bits e: 2,6,10,11
constexpr unsigned EDGE_AX = 0x558888u;  // 2 bits per edge: axis
__host__ __device__ constexpr int edge_dx(int e) { return (EDGE_DX >> e) & 1; }
__host__ __device__ constexpr int edge_dy(int e) { return (EDGE_DY >> e) & 1; }
__host__ __device__ constexpr int edge_dz(int e) { return (EDGE_DZ >> e) & 1; }
__host__ __device__ constexpr int edge_axis(int e) { return (EDGE_AX >> (2 * e)) & 3; }
}  // namespace diso
