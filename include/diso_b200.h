/* diso_b200.h -- C ABI of libdiso_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for the reference's hot path.  It replaces the pybind11 module
 * `diso._C` of the reference (/root/reference/src/pybind.cpp:419-447: classes CUMCFloat,
 * CUMCDouble, CUDMCFloat, CUDMCDouble with forward(...)/backward(...)) and everything below it
 * (/root/reference/src/cumc.cu, cudualmc.cu).  Differences by design:
 *   - plain C, no torch / pybind types: raw device pointers, sizes, a cudaStream_t as void*;
 *   - stateless: all memory (outputs, saved state, scratch) is owned by the caller, so one
 *     process may run any number of extractions concurrently on any streams / devices
 *     (the reference keeps mutable scratch inside the extractor object, pybind.cpp:16-39);
 *   - inputs are the UNPADDED grid [X,Y,Z] (+ deform [X,Y,Z,3], AoS xyz); the iso+1 / zero
 *     padding of diso/__init__.py:52-54 is applied virtually inside the kernels, and the
 *     "-1" shift, the optional division by (dims-1) (diso/__init__.py:56-60) and the int64
 *     widening of faces (diso/__init__.py:61,116) are fused into the emit kernels;
 *   - two-phase forward (count -> caller allocates exact outputs -> emit): exactly one host
 *     synchronisation per forward instead of the reference's five;
 *   - every function returns 0 on success or a negative DISO_E_* code and never prints/aborts
 *     (the reference prints CUDA errors and continues, cumc.cu:68-78).
 *
 * Conventions: dtype 0 = float32, 1 = float64.  alg 0 = marching cubes, 1 = dual marching cubes.
 * All pointers are device pointers on the current CUDA device unless stated otherwise.  All
 * work is enqueued on `stream` (a cudaStream_t); nothing here synchronises the host.
 * Grid memory order is C-contiguous [X][Y][Z] (z fastest), like the reference (cumc.h:101-105).
 */
#ifndef DISO_B200_H
#define DISO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISO_B200_ABI_VERSION 3

#define DISO_ALG_MC 0
#define DISO_ALG_DMC 1
#define DISO_F32 0
#define DISO_F64 1

#define DISO_GRAD_REFERENCE 0 /* bug-compatible DMC adjoint (cudualmc.cu:975,990) */
#define DISO_GRAD_EXACT 1     /* true adjoint of the forward */

#define DISO_OK 0
#define DISO_E_INVALID (-1)  /* bad argument (null pointer, dims < 1, unknown dtype/alg) */
#define DISO_E_STATE (-2)    /* state buffer too small / not produced by diso_b200_count */
#define DISO_E_CUDA (-3)     /* a CUDA runtime call or launch failed; see diso_b200_last_error */
#define DISO_E_TOOLARGE (-4) /* grid exceeds the 32-bit index budget of one call */

/* Number of int64 slots at the start of the state buffer that diso_b200_count fills. */
#define DISO_COUNT_SLOTS 8
#define DISO_CNT_VERTS 0    /* MC: #vertices        DMC: #dual vertices                    */
#define DISO_CNT_FACES 1    /* MC: #triangles       DMC: #quads                            */
#define DISO_CNT_ANY_GT 2   /* 1 iff some sdf value > iso  (max <= iso  <=>  0)            */
#define DISO_CNT_EDGES 3    /* #crossing edges (== MC verts == DMC quads)                  */
#define DISO_CNT_USED 4     /* #used cells (cells whose 8 corners are not all on one side) */
#define DISO_CNT_EDGE_CHUNKS 5 /* #32-point chunks owning >= 1 crossing edge (the emit kernels visit only those) */
#define DISO_CNT_CELL_CHUNKS 6 /* #chunks with >= 1 triangle / dual vertex */

/* Frame of a slab inside a larger grid (slab sharding of one grid across GPUs along dim 0,
 * SURVEY.md section 8e; the reference has no counterpart).  Passed by HOST pointer to emit /
 * backward; NULL = the grid stands alone.  With a frame, a rank that extracts layers
 * [x_origin, x_origin + X) of a grid of X_global layers writes vertices in the GLOBAL frame
 * (bit-identical to a single extraction of the whole grid: the integer x coordinate is offset
 * before the deformation is added and the normalisation divides by X_global - 1) and face indices
 * shifted by id_offset (local id -> global id), so no post-processing pass touches the outputs. */
typedef struct diso_b200_frame {
    int32_t x_origin;  /* global index of the local grid's first x layer */
    int32_t X_global;  /* X of the whole grid (<= 0: use the local X) */
    int64_t id_offset; /* added to every vertex id written to tris / quads (may be negative) */
} diso_b200_frame;

int diso_b200_abi_version(void);

/* Thread-local, NUL-terminated description of the last error returned on this thread. */
const char *diso_b200_last_error(void);

/* Bytes of caller-owned state needed for one extraction of an X*Y*Z grid.  The first
 * DISO_COUNT_SLOTS*8 bytes receive the counts; the rest is the compact rank structure
 * (sign bitmask + per-32-point records) that emit/backward consume.  Returns 0 on bad args. */
size_t diso_b200_state_bytes(int alg, int X, int Y, int Z);

/* Byte offsets of the arrays inside `state`, for hosts that want to
 * read the per-chunk prefix sums (slab sharding, diso_b200/parallel.py):
 * out[0..3] = offsets of S (u32), E (uint4), F (uint2, MC) | P (uint4, DMC), C (u16 per cell);
 * out[4] = NC (chunks per padded row), out[5] = NCH (chunks), out[6] = total bytes,
 * out[7] = chunks per padded x-layer.  Entry NCH of E / F / P holds the grand totals. */
int diso_b200_state_layout(int alg, int X, int Y, int Z, int64_t *out);

/* Phase 1 (replaces count_used_cells / index_used_cells / count_cell_mc_verts /
 * count_cell_mc_tris|count_cell_patches and the three cub scans, cumc.cu:661-723,
 * cudualmc.cu:1068-1122): classify every cell, count vertices / faces, build the rank
 * structure.  Afterwards the first DISO_COUNT_SLOTS int64 of `state` hold the counts; the
 * caller copies them to the host (its single sync), allocates outputs and calls *_emit. */
int diso_b200_count(int alg, const void *sdf, int dtype, int X, int Y, int Z, double iso,
                    void *state, size_t state_bytes, void *stream);

/* The forward's single host synchronisation: copies the DISO_COUNT_SLOTS int64 at the start of `state` into
 * counts_host (HOST memory; pinned memory avoids a staging copy) on `stream` and waits for that stream.
 * (The reference synchronises five times per forward: cumc.cu:674,694,723 and the two reductions of
 * diso/__init__.py:49.) */
int diso_b200_read_counts(const void *state, int64_t *counts_host, void *stream);

/* Phase 2, marching cubes (replaces create_cell_mc_verts / create_cell_mc_tris,
 * cumc.cu:370-410, 564-612, and the epilogue diso/__init__.py:56-61).
 * verts: [n_verts,3] dtype; tris: [n_tris,3] int64.  deform may be NULL.
 * counts_host: HOST pointer to the DISO_COUNT_SLOTS int64 the caller read back after
 * diso_b200_count (the active-chunk counts size the launches; sparse surfaces then cost time
 * proportional to the surface, not the volume).  NULL = visit every chunk.
 *
 * edge_rec (ABI v3, may be NULL): caller-owned, NCOMP * 32 * ceil(edge_rec_stride / 32) elements of dtype with
 * edge_rec_stride >= #crossing edges (groups of 32 edges, component-major inside a group); NCOMP = 5 with deform, 2 when
 * deform is NULL, plus 1 for diso_b200_dmc_emit (the last word per edge holds the quad's patch lengths).  When given,
 * the edge pass also SAVES, per crossing edge and indexed by its rank (== its MC vertex id == its DMC quad id), what the
 * adjoint of computeMcVert needs (adjComputeMcVert, cumc.cu:412-453): p1 - p0 (x, y, z), d0, d1 -- without deform p1 - p0
 * is the edge's unit axis vector and only d0, d1 are kept.  Passing the same buffer (and the same deform / NULL) to
 * *_backward selects the saved-record backward, which reads neither sdf nor deform; the reference instead re-runs the
 * whole forward inside backward (diso/__init__.py:32,86).  20 / 8 bytes (fp32; fp64: twice that) per crossing edge;
 * callers that never differentiate pass NULL. */
int diso_b200_mc_emit(const void *sdf, const void *deform, int dtype, int X, int Y, int Z,
                      double iso, const void *state, const int64_t *counts_host, int normalize,
                      const diso_b200_frame *frame, void *verts, int64_t *tris, void *edge_rec,
                      int64_t edge_rec_stride, void *stream);

/* Phase 2, dual marching cubes (replaces create_dmc_verts / create_quads,
 * cudualmc.cu:907-955, 1027-1056, and diso/__init__.py:110-116).
 * verts: [n_verts,3] dtype; quads: [n_quads,4] int64.
 * scratch: caller-owned, n_quads*3 elements of dtype (edge crossings, each evaluated once).
 * edge_rec / edge_rec_stride: as for diso_b200_mc_emit (the crossing edges are the quads).
 * quad_flags (may be NULL): [n_quads] bytes; when given, the quad kernel also decides each quad's diagonal for the
 * quad -> triangle split (1 = first diagonal, [0,1,3][1,2,3]) while the four ids are in registers; pass the buffer to
 * diso_b200_quad_split, which then only scans the bytes. */
int diso_b200_dmc_emit(const void *sdf, const void *deform, int dtype, int X, int Y, int Z,
                       double iso, const void *state, const int64_t *counts_host, int normalize,
                       const diso_b200_frame *frame, void *scratch, void *verts, int64_t *quads,
                       void *edge_rec, int64_t edge_rec_stride, uint8_t *quad_flags, void *stream);

/* Backward, marching cubes (replaces adj_create_cell_mc_verts, cumc.cu:474-512, the dense
 * zero-fills of diso/__init__.py:33,40 and the pad-backward slices).  adj_verts is dL/dverts in
 * the API frame (after -1 / normalisation), [n_verts,3] contiguous.  adj_sdf [X,Y,Z] and
 * adj_deform [X,Y,Z,3] (NULL iff deform is NULL) are FULLY written (zeros included);
 * accumulation is an atomic-free gather in a fixed order, so results are deterministic.
 * counts_host: as for emit (NULL allowed); on sparse surfaces it selects the zero-fill + touched-block
 * path (its block list is a stream-ordered allocation private to the call, so `state` is only read and
 * several backward passes may share it -- retained graphs, several streams).
 * edge_rec (ABI v3): the buffer the emit call filled, or NULL.  With it the adjoint runs from the saved records
 * (mc_backward_v2.cuh) and either of adj_sdf / adj_deform may be NULL when that gradient is not wanted; without it
 * the kernel re-gathers sdf / deform (mc_backward_compact.cuh) and both outputs are required. */
int diso_b200_mc_backward(const void *sdf, const void *deform, int dtype, int X, int Y, int Z,
                          double iso, void *state, const int64_t *counts_host, const void *adj_verts,
                          int normalize, const diso_b200_frame *frame, const void *edge_rec,
                          int64_t edge_rec_stride, void *adj_sdf, void *adj_deform, void *stream);

/* Backward, dual marching cubes (replaces adj_create_dmc_verts, cudualmc.cu:957-1005).
 * scratch: caller-owned per-edge adjoints, n_quads*3 elements of dtype; may be NULL when edge_rec AND quads are given.
 * quads (may be NULL): the [n_quads,4] int64 output of diso_b200_dmc_emit.  With edge_rec and quads the whole adjoint is
 * ONE kernel: the four dual vertices around an edge are its quad, the patch lengths were saved next to the records, so
 * neither cell words nor case tables are consulted and no per-edge adjoint is materialised.
 * edge_rec / edge_rec_stride: as for diso_b200_mc_backward. */
int diso_b200_dmc_backward(const void *sdf, const void *deform, int dtype, int X, int Y, int Z,
                           double iso, void *state, const int64_t *counts_host,
                           const void *adj_verts, int normalize, const diso_b200_frame *frame,
                           int grad_mode, const void *edge_rec, int64_t edge_rec_stride,
                           const int64_t *quads, void *scratch, void *adj_sdf, void *adj_deform, void *stream);

/* Quad -> triangle split of diso/__init__.py:118-147 as two kernels (no PyTorch
 * temporaries).  verts [n_verts,3] dtype (API frame), quads [n_quads,4] int64, faces
 * [2*n_quads,3] int64.  scratch: caller-owned, diso_b200_quad_split_scratch_bytes(n_quads).
 * Output order == the reference's mask + cat: all quads whose first diagonal wins
 * (angles1 < angles2 -> [0,1,3],[1,2,3]) in quad order, then the rest ([0,1,2],[0,2,3]).
 * quad_flags: NULL, or the per-quad flags diso_b200_dmc_emit wrote (then verts may be NULL: nothing is re-read). */
size_t diso_b200_quad_split_scratch_bytes(int64_t n_quads);
int diso_b200_quad_split(const void *verts, int dtype, const int64_t *quads, int64_t n_quads,
                         const uint8_t *quad_flags, void *scratch, int64_t *faces, void *stream);

/* Tracing (the reference has none).  diso_b200_launch_count: number of
 * kernels this library has launched in the process.  diso_b200_profile_enable(1) makes every
 * subsequent launch (process-wide: autograd issues backward from its own thread) record a
 * CUDA-event pair on its stream;
 * diso_b200_profile_dump waits for them and writes one "kernel_name milliseconds" line per
 * launch into buf (NUL-terminated), then clears the list.  Used by bench.py for the live
 * per-kernel roofline figure; off by default (no events, no overhead). */
long long diso_b200_launch_count(void);
int diso_b200_profile_enable(int on);
int diso_b200_profile_dump(char *buf, size_t cap);

/* Test / diagnostics hook: expands the rank structure into the reference's intermediate
 * "case index per cell" so parity tests can compare the active-cell set and the 8-bit case
 * index bit-exactly (reference: used_cell_index / used_cell_code, cumc.cu:299-311,540-562).
 * codes [PX*PY*PZ] uint8 (PX=X+2, ...) receives, for every padded cell in linear order
 * z + PZ*(y + PY*x), its case index (DMC: after the ambiguity flip); 0 / 255 == unused. */
int diso_b200_debug_cell_codes(int alg, int X, int Y, int Z, const void *state, uint8_t *codes,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DISO_B200_H */
