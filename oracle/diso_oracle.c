/* ==========================================================================================
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement (plain C, single thread, sequential loops) of the algorithm of the
 * reference's hot path: differentiable Marching Cubes and Dual Marching Cubes,
 * forward + backward, as implemented by
 *     /root/reference/src/cumc.cu      (CuMC::forward :651-732, ::backward :734-743)
 *     /root/reference/src/cudualmc.cu  (CUDualMC::forward :1058-1128, ::backward :1130-1138)
 * Each function in diso_oracle_impl.inc cites the reference lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the reported CPU baseline.  The product
 * (diso_b200/) never imports, links or falls back to it.
 *
 * Parity pin: the reference ships no golden vectors.  This oracle is
 * pinned against outputs of the UNMODIFIED reference CUDA build (baseline/_ref) run on a B200
 * via tests/golden/make_golden.py; the resulting fixtures live in tests/golden/ (npz files) and are
 * checked by tests/test_oracle_golden.py.
 * ========================================================================================== */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "diso_tables.h"

typedef struct oracle_mesh {
    int64_t n_used;      /* number of used (surface-crossing) cells */
    int64_t n_verts;     /* MC: edge vertices; DMC: dual vertices */
    int64_t n_faces;     /* MC: triangles; DMC: quads */
    int64_t n_quads;     /* DMC only (== n_faces) */
    void *verts;         /* [n_verts,3] scalar, PADDED frame, before the "-1" shift */
    int32_t *faces;      /* MC [n_faces,3]; DMC [n_faces,4] */
    int32_t *used_index; /* [n_used] linear padded cell index, ascending */
    uint8_t *used_code;  /* [n_used] 8-bit case index (DMC: after the ambiguity flip) */
    int scalar_size;
} oracle_mesh;

void oracle_mesh_free(oracle_mesh *m)
{
    if (!m)
        return;
    free(m->verts);
    free(m->faces);
    free(m->used_index);
    free(m->used_code);
    free(m);
}

int oracle_abi_version(void) { return 1; }

#define SCALAR float
#define NAME(x) x##_f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#include "diso_oracle_impl.inc"
#undef SCALAR
#undef NAME
#undef FMA

#define SCALAR double
#define NAME(x) x##_f64
#define FMA(a, b, c) fma((a), (b), (c))
#include "diso_oracle_impl.inc"
#undef SCALAR
#undef NAME
#undef FMA
