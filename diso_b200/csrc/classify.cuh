// classify.cuh -- phase 1 of the forward: sign packing and the fused count + single-pass scan.
//
// Replaces (reference): count_used_cells_kernel + cub::ExclusiveSum + index_used_cells_kernel
// + count_cell_mc_verts_kernel + cub::ExclusiveSum + count_cell_mc_tris_kernel /
// count_cell_patches_kernel + cub::ExclusiveSum  (cumc.cu:265-341,540-562,661-723;
// cudualmc.cu:605-681,815-863,1068-1122) and the two full-grid min/max reductions of
// diso/__init__.py:49.
//
//   K1 sign_pack      : streams the SDF once (the only compulsory dense read of phase 1),
//                       emits 1 bit per padded point.  HBM-bound: G*s bytes in, G/8 out.
//   K2 classify_scan  : one THREAD per 32-point chunk, bit-parallel over lane masks that
//                       live in L2 (G/8 bytes); counts crossing edges / triangles / patches
//                       and scans them with a decoupled look-back across 256-chunk tiles.
#pragma once
#include "common.cuh"
#include "tables.cuh"

namespace diso {

// Tails of the state arrays that no kernel computes (sign words beyond the last chunk read as "inside", records
// beyond it as "no edges"): filled by the sign pass itself instead of three cudaMemsetAsync calls per extraction
// (host latency is what a 64^3 call is made of).  classify_scan, the first reader, runs after this kernel.
struct TailFill {
    unsigned *s_tail; int n_s;      // sign words  := ~0
    uint4 *e_tail; int n_e;         // edge records := 0
    uint2 *aux_tail; int n_aux;     // F / P records (8-byte units) := 0
};
__device__ __forceinline__ void fill_tails(const TailFill &tf)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = t; i < tf.n_s; i += nt) tf.s_tail[i] = FULL;
    for (int i = t; i < tf.n_e; i += nt) tf.e_tail[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = t; i < tf.n_aux; i += nt) tf.aux_tail[i] = make_uint2(0u, 0u);
}

// ------------------------------------------------------------------------------------------
// K1: one warp per padded z-row.  Lane j of iteration c looks at padded point zp = 32c + j.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) sign_pack_kernel(const T *__restrict__ sdf, Geo g, T iso,
                                                        unsigned *__restrict__ S,
                                                        long long *__restrict__ counts, TailFill tf)
{
    fill_tails(tf);
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int row = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    if (row >= g.NR) return;
    const int xp = row / g.PY, yp = row - xp * g.PY;
    unsigned *out = S + (size_t)row * g.NC;
    const bool real_row = xp >= 1 && xp <= g.X && yp >= 1 && yp <= g.Y;
    if (!real_row) {
        for (int c = lane; c < g.NC; c += 32) out[c] = FULL;
        return;
    }
    const T *rowp = sdf + ((size_t)(xp - 1) * g.Y + (yp - 1)) * g.Z;
    bool any_gt = false;
    unsigned mine = FULL;
    for (int c0 = 0; c0 < g.NC; c0 += 32) {
        const int cend = min(32, g.NC - c0);
#pragma unroll 4
        for (int cc = 0; cc < cend; ++cc) {
            const int z = 32 * (c0 + cc) + lane - 1;
            const bool in = (unsigned)z < (unsigned)g.Z;
            bool b = true;
            if (in) {
                T v = __ldcs(rowp + z);  // streamed: this pass never re-reads a value
                b = v >= iso;
                any_gt |= v > iso;
            }
            unsigned w = __ballot_sync(FULL, b);
            if (lane == cc) mine = w;
        }
        if (lane < cend) out[c0 + lane] = mine;
    }
    if (__any_sync(FULL, any_gt) && lane == 0) counts[DISO_CNT_ANY_GT] = 1;
}

// Vectorised K1 for rows whose length is a multiple of the vector width and 16-byte aligned: each lane
// loads 128-bit vectors (float4 / double2), i.e. a warp consumes 512 contiguous bytes per instruction,
// four instructions per batch.  The grid is persistent (a few CTAs per SM, warps stride over the
// rows) and the loads of the NEXT batch are issued before the current one is digested, so a
// warp always has 2 KB in flight (the one-row-per-CTA version was bound by CTA turnover: ncu
// showed 34 % issue activity and 14 % DRAM throughput).
// The per-lane bit groups (4 bits for fp32, 2 for fp64) are merged into aligned 32-bit words inside
// 8- / 16-lane groups with shuffle-OR steps, staged in shared memory, then shifted by the one-point
// pad offset.
template <typename T> struct Vec128;
template <> struct Vec128<float> {
    using type = float4;
    static constexpr int N = 4;
    static __device__ __forceinline__ float4 fill(float v) { return make_float4(v, v, v, v); }
    static __device__ __forceinline__ unsigned ge(const float4 &q, float iso) { return (q.x >= iso ? 1u : 0u) | (q.y >= iso ? 2u : 0u) | (q.z >= iso ? 4u : 0u) | (q.w >= iso ? 8u : 0u); }
    static __device__ __forceinline__ bool gt(const float4 &q, float iso) { return (q.x > iso) | (q.y > iso) | (q.z > iso) | (q.w > iso); }
};
template <> struct Vec128<double> {
    using type = double2;
    static constexpr int N = 2;
    static __device__ __forceinline__ double2 fill(double v) { return make_double2(v, v); }
    static __device__ __forceinline__ unsigned ge(const double2 &q, double iso) { return (q.x >= iso ? 1u : 0u) | (q.y >= iso ? 2u : 0u); }
    static __device__ __forceinline__ bool gt(const double2 &q, double iso) { return (q.x > iso) | (q.y > iso); }
};

template <typename T>
__global__ void __launch_bounds__(256) sign_pack_vec_kernel(const T *__restrict__ sdf, Geo g, T iso, unsigned *__restrict__ S,
                                                            long long *__restrict__ counts, TailFill tf)
{
    fill_tails(tf);
    using VT = Vec128<T>;
    using V = typename VT::type;
    constexpr int N = VT::N;        // values per lane per load
    constexpr int LPW = 32 / N;     // lanes that share one 32-bit sign word
    extern __shared__ unsigned sm_words[];  // per warp: NA+2 aligned words
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int warps_per_cta = blockDim.x >> 5;
    const int nwarps = gridDim.x * warps_per_cta;
    const int NA = (g.Z + 31) / 32;  // aligned words covering z = 0..Z-1
    unsigned *aw = sm_words + wid * (NA + 2);
    const int nvec = g.Z / N;
    const int nb = (nvec + 127) / 128;  // batches of 4 x 32 vectors per row
    bool any_gt = false;

    auto row_ptr = [&](int row) -> const V * {
        const int xp = row / g.PY, yp = row - xp * g.PY;
        const bool real = xp >= 1 && xp <= g.X && yp >= 1 && yp <= g.Y;
        return real ? reinterpret_cast<const V *>(sdf + ((size_t)(xp - 1) * g.Y + (yp - 1)) * g.Z) : nullptr;
    };
    auto load_batch = [&](int row, int b, V (&q)[4]) {
        const V *rp = row_ptr(row);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int v = b * 128 + 32 * u + lane;
            q[u] = VT::fill(iso);  // pad rows / beyond the row: "inside"
            if (rp && v < nvec) q[u] = __ldcs(rp + v);
        }
    };

    // A warp owns whole rows (row = first, first + nwarps, ...) and walks their batches in order,
    // because the aligned words of a row are staged in the warp's private shared-memory slice.
    int row = blockIdx.x * warps_per_cta + wid, b = 0;
    V q[4], qn[4];
    if (row < g.NR) load_batch(row, 0, q);
    while (row < g.NR) {
        int nrow = row, nbat = b + 1;
        if (nbat == nb) { nrow = row + nwarps; nbat = 0; }
        if (nrow < g.NR) load_batch(nrow, nbat, qn);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            any_gt |= VT::gt(q[u], iso);
            unsigned w = VT::ge(q[u], iso) << (N * (lane & (LPW - 1)));
#pragma unroll
            for (int d = 1; d < LPW; d <<= 1) w |= __shfl_xor_sync(FULL, w, d);
            const int widx = (b * 128 + 32 * u) / LPW + lane / LPW;  // aligned word index: LPW vectors per word
            if ((lane & (LPW - 1)) == 0 && widx < NA) aw[widx] = w;
        }
        if (b == nb - 1) {  // row complete: shift by the pad offset and store
            __syncwarp();
            unsigned *out = S + (size_t)row * g.NC;
            // padded word c covers z = 32c-1 .. 32c+30  ->  (aligned[c] << 1) | (aligned[c-1] >> 31)
            for (int c = lane; c < g.NC; c += 32) {
                const unsigned cur = c < NA ? aw[c] : FULL;
                const unsigned prev = c > 0 ? aw[c - 1] : FULL;
                out[c] = (cur << 1) | (prev >> 31);
            }
            __syncwarp();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = qn[u];
        row = nrow;
        b = nbat;
    }
    if (__any_sync(FULL, any_gt) && lane == 0) counts[DISO_CNT_ANY_GT] = 1;
}

// ------------------------------------------------------------------------------------------
// Cell case index from lane masks.  Words: A=(x,y) B=(x+1,y) C=(x+1,y+1) D=(x,y+1); the *1
// variants are the same rows one point further in z.
//   MC  corner order (cumc.cu:107)       : 000,100,110,010,001,101,111,011 -> A B C D A1 B1 C1 D1
//   DMC corner order (cudualmc.cu:95-104): 000,100,010,110,001,101,011,111 -> A B D C A1 B1 D1 C1
// ------------------------------------------------------------------------------------------
struct CellWords {
    unsigned A, B, C, D, A1, B1, C1, D1;
};

__device__ __forceinline__ CellWords load_cell_words(const unsigned *__restrict__ S, const Geo &g, int k)
{
    CellWords w;
    unsigned a = S[k], an = S[k + 1];
    unsigned d = S[k + g.sY], dn = S[k + g.sY + 1];
    unsigned b = S[k + g.sX], bn = S[k + g.sX + 1];
    unsigned c = S[k + g.sX + g.sY], cn = S[k + g.sX + g.sY + 1];
    w.A = a; w.B = b; w.C = c; w.D = d;
    w.A1 = shift_in(a, an); w.B1 = shift_in(b, bn); w.C1 = shift_in(c, cn); w.D1 = shift_in(d, dn);
    return w;
}

__device__ __forceinline__ unsigned used_mask(const CellWords &w)
{
    unsigned all = w.A & w.B & w.C & w.D & w.A1 & w.B1 & w.C1 & w.D1;
    unsigned any = w.A | w.B | w.C | w.D | w.A1 | w.B1 | w.C1 | w.D1;
    return any & ~all;
}

template <int ALG>
__device__ __forceinline__ unsigned cell_code(const CellWords &w, int j)
{
    unsigned c = bit(w.A, j) | (bit(w.B, j) << 1) | (bit(w.A1, j) << 4) | (bit(w.B1, j) << 5);
    if (ALG == DISO_ALG_MC)
        c |= (bit(w.C, j) << 2) | (bit(w.D, j) << 3) | (bit(w.C1, j) << 6) | (bit(w.D1, j) << 7);
    else
        c |= (bit(w.D, j) << 2) | (bit(w.C, j) << 3) | (bit(w.D1, j) << 6) | (bit(w.C1, j) << 7);
    return c;
}

// Case indices of all 32 cells of a chunk at once: the 8 (corner) x 32 (cell) bit matrix is transposed in
// four 8x8 blocks (three masked swap rounds each).  codes[i] holds cells 4i .. 4i+3, one byte per cell.
// ~110 instructions per chunk instead of ~23 per used cell for the bit-by-bit gather of cell_code().
template <int ALG>
__device__ __forceinline__ void cell_codes32(const CellWords &w, unsigned (&codes)[8])
{
    unsigned W[8];
    W[0] = w.A; W[1] = w.B; W[4] = w.A1; W[5] = w.B1;
    if (ALG == DISO_ALG_MC) { W[2] = w.C; W[3] = w.D; W[6] = w.C1; W[7] = w.D1; }
    else                    { W[2] = w.D; W[3] = w.C; W[6] = w.D1; W[7] = w.C1; }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const unsigned sel = (unsigned)g | ((unsigned)(4 + g) << 4);   // byte g of the first, byte g of the second operand
        const unsigned lo = __byte_perm(__byte_perm(W[0], W[1], sel), __byte_perm(W[2], W[3], sel), 0x5410);
        const unsigned hi = __byte_perm(__byte_perm(W[4], W[5], sel), __byte_perm(W[6], W[7], sel), 0x5410);
        unsigned long long x = ((unsigned long long)hi << 32) | lo;   // row i (byte i) = corner i, bit c = cell 8g + c
        unsigned long long t;
        t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
        t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
        t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
        codes[2 * g] = (unsigned)x;                                    // byte c = case index of cell 8g + c
        codes[2 * g + 1] = (unsigned)(x >> 32);
    }
}

// DMC case index of the cell at (chunk k, bit j) computed from scratch (used for the
// ambiguity test's neighbour, cudualmc.cu:828-829).  j may be -1 or 32 (previous/next chunk).
__device__ __forceinline__ unsigned dmc_code_at(const unsigned *__restrict__ S, const Geo &g, int k, int j)
{
    if (j < 0) { k -= 1; j = 31; }
    if (j > 31) { k += 1; j = 0; }
    CellWords w = load_cell_words(S, g, k);
    return cell_code<DISO_ALG_DMC>(w, j);
}

// cudualmc.cu:815-839: does the cell at (xp,yp,c,j) with raw case `code` get complemented?
// `ctab` = T_DMC_CASE (bit 31 problematic, bits 28..30 = 2*axis + (dir>0)).
__device__ __forceinline__ bool dmc_flip(const unsigned *__restrict__ S, const Geo &g,
                                         const unsigned *__restrict__ ctab, int k, int xp, int yp,
                                         int c, int j, unsigned code)
{
    unsigned e = ctab[code];
    if (!(e >> 31)) return false;
    int dir = (e >> 28) & 7;
    int comp = dir >> 1, delta = (dir & 1) ? 1 : -1;
    int nk = k, nj = j;
    // neighbour must be a valid padded cell: 0 <= n < dim-1 (cudualmc.cu:826).  The upper bound
    // needs no test: cells beyond it read only pad bits (all ones) -> case 255 -> not problematic.
    if (comp == 0) { if (xp + delta < 0) return false; nk += delta * g.sX; }
    else if (comp == 1) { if (yp + delta < 0) return false; nk += delta * g.sY; }
    else { if (32 * c + j + delta < 0) return false; nj += delta; }
    unsigned ncode = dmc_code_at(S, g, nk, nj);
    return (ctab[ncode] >> 31) != 0;
}

// ------------------------------------------------------------------------------------------
// K2: fused count + decoupled look-back scan.  Thread t of tile T owns chunk k = 256*T + t.
// ------------------------------------------------------------------------------------------
template <int ALG>
#ifndef DISO_CLASSIFY_MINB
#define DISO_CLASSIFY_MINB 8
#endif
__global__ void __launch_bounds__(SCAN_TILE, DISO_CLASSIFY_MINB)   // 32 registers: 8 CTAs/SM hide the look-back latency (DMC 0.44 -> 0.42 ms)
 classify_scan_kernel(Geo g, const unsigned *__restrict__ S,
                                                                  uint4 *__restrict__ E, void *__restrict__ aux,
                                                                  unsigned short *__restrict__ C,
                                                                  unsigned *__restrict__ active,
                                                                  TileDesc *__restrict__ desc,
                                                                  unsigned *__restrict__ ticket,
                                                                  long long *__restrict__ counts)
{
    __shared__ unsigned s_tab[256];  // MC: triangle count per case; DMC: T_DMC_CASE
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_warp_tot[SCAN_TILE / 32];
    __shared__ unsigned long long s_excl[4];
    __shared__ unsigned s_used_tot;
    // per-cell words of the tile, staged so they leave the SM as coalesced stores: row stride of 17
    // words keeps the per-thread 2-byte writes of one warp in 32 distinct banks
    constexpr int CW_STRIDE = 17;
    __shared__ unsigned s_cw[SCAN_TILE * CW_STRIDE];
    __shared__ unsigned s_any_cells;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < CW_STRIDE; ++i) s_cw[i * SCAN_TILE + tid] = 0u;
    if (tid == 0) s_any_cells = 0u;
    for (int i = tid; i < 256; i += SCAN_TILE) {
        if (ALG == DISO_ALG_MC) s_tab[i] = (unsigned)(T_MC_CASE[i] >> 60);
        else                    s_tab[i] = T_DMC_CASE[i];
    }
    if (tid == 0) { s_tile = atomicAdd(ticket, 1u); s_used_tot = 0; }
    __syncthreads();
    const int tile = (int)s_tile;
    const int k = tile * SCAN_TILE + tid;

    unsigned mx = 0, my = 0, mz = 0, lo = 0, hi = 0, used = 0;
    unsigned na = 0, nb = 0, nused = 0;
    CellWords w{FULL, FULL, FULL, FULL, FULL, FULL, FULL, FULL};
    if (k < g.NCH) {
        w = load_cell_words(S, g, k);
        mx = w.A ^ w.B;
        my = w.A ^ w.D;
        mz = w.A ^ w.A1;
        na = __popc(mx) + __popc(my) + __popc(mz);
        used = used_mask(w);
        nused = __popc(used);
    }
    // Dense warps (many used cells per chunk: random fields) take the transposed, fully predicated path
    // below: all 32 cells in straight-line code, no per-lane loop trip counts to diverge on.  Warps that
    // only graze a smooth surface keep the short per-cell loops.
    const bool dense_warp = __reduce_max_sync(FULL, nused) > 8u;
    if (__any_sync(FULL, nused != 0u) && lane == 0) s_any_cells = 1u;
    {
        if (used) {
            int r = k / g.NC, c = k - r * g.NC;
            int xp = r / g.PY, yp = r - xp * g.PY;
            unsigned short *crow = reinterpret_cast<unsigned short *>(s_cw + tid * CW_STRIDE);
            if (ALG == DISO_ALG_MC) {
                if (dense_warp) {
                    unsigned codes[8];
                    cell_codes32<ALG>(w, codes);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {   // ascending j: nb is the cell's offset inside the chunk
                        if ((used >> j) & 1u) {
                            const unsigned code = (codes[j >> 2] >> (8 * (j & 3))) & 0xffu;
                            crow[j] = (unsigned short)(code | (nb << 8));
                            nb += s_tab[code];
                        }
                    }
                } else {
                    unsigned u = used;
                    while (u) {
                        const int j = __ffs(u) - 1;
                        u &= u - 1;
                        const unsigned code = cell_code<ALG>(w, j);
                        crow[j] = (unsigned short)(code | (nb << 8));
                        nb += s_tab[code];
                    }
                }
            } else {
                // Three passes keep the warp converged: only ~13 % of the used cells of a random field
                // are "problematic", but inside one per-cell loop nearly every warp iteration would
                // have SOME lane on the long neighbour-lookup path.
                unsigned prob = 0;
                if (dense_warp) {                                      // 1: raw case index, who needs the test
                    unsigned codes[8];
                    cell_codes32<ALG>(w, codes);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((used >> j) & 1u) {
                            const unsigned code = (codes[j >> 2] >> (8 * (j & 3))) & 0xffu;
                            crow[j] = (unsigned short)code;
                            prob |= (s_tab[code] >> 31) << j;
                        }
                    }
                } else {
                    for (unsigned u = used; u; u &= u - 1) {
                        const int j = __ffs(u) - 1;
                        const unsigned code = cell_code<ALG>(w, j);
                        crow[j] = (unsigned short)code;
                        prob |= (s_tab[code] >> 31) << j;
                    }
                }
                for (unsigned u = prob; u; u &= u - 1) {           // 2: ambiguity test (cudualmc.cu:815-839)
                    const int j = __ffs(u) - 1;
                    const unsigned code = crow[j];
                    if (dmc_flip(S, g, s_tab, k, xp, yp, c, j, code)) crow[j] = (unsigned short)(code ^ 0xffu);
                }
                if (dense_warp) {                                      // 3: patch counts + offsets, ascending j
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((used >> j) & 1u) {
                            const unsigned code = crow[j];
                            crow[j] = (unsigned short)(code | (nb << 8));
                            const unsigned np = (s_tab[code] >> 24) & 7u;  // 1..4
                            nb += np;
                            lo |= ((np - 1u) & 1u) << j;
                            hi |= ((np - 1u) >> 1) << j;
                        }
                    }
                } else {
                    for (unsigned u = used; u; u &= u - 1) {
                        const int j = __ffs(u) - 1;
                        const unsigned code = crow[j];
                        crow[j] = (unsigned short)(code | (nb << 8));
                        const unsigned np = (s_tab[code] >> 24) & 7u;  // 1..4
                        nb += np;
                        lo |= ((np - 1u) & 1u) << j;
                        hi |= ((np - 1u) >> 1) << j;
                    }
                }
            }
        }
    }


    // ---- tile-local exclusive scan of the packed quadruple (a | b<<16 | c<<32 | d<<48) -------
    const unsigned long long v = (unsigned long long)na | ((unsigned long long)nb << 16) |
                                 ((unsigned long long)(na != 0u) << 32) | ((unsigned long long)(nb != 0u) << 48);
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp_tot[wid] = inc;
    unsigned ucnt = __reduce_add_sync(FULL, nused);
    if (lane == 0 && ucnt) atomicAdd(&s_used_tot, ucnt);
    __syncthreads();
    unsigned long long warp_off = 0, tile_tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_TILE / 32; ++i) {
        unsigned long long t = s_warp_tot[i];
        if (i < wid) warp_off += t;
        tile_tot += t;
    }
    const unsigned long long excl_local = warp_off + inc - v;

    // ---- decoupled look-back across tiles (warp 0) ----------------------------------------
    if (wid == 0) {
        unsigned long long e[4] = {0, 0, 0, 0};
        if (tile == 0) {
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) st_relaxed_u64(&desc[0].incl[q], (tile_tot >> (16 * q)) & 0xffffull);
                st_release_u32(&desc[0].flag, 2u);
            }
        } else {
            if (lane == 0) {
                st_relaxed_u64(&desc[tile].agg, tile_tot);
                st_release_u32(&desc[tile].flag, 1u);
            }
            int base = tile - 1;
            while (true) {
                const int t = base - lane;
                unsigned f = 2u;
                unsigned long long x[4] = {0, 0, 0, 0};
                if (t >= 0) {
                    do { f = ld_acquire_u32(&desc[t].flag); } while (f == 0u);
                    if (f == 2u) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) x[q] = ld_relaxed_u64(&desc[t].incl[q]);
                    } else {
                        const unsigned long long pk = ld_relaxed_u64(&desc[t].agg);
#pragma unroll
                        for (int q = 0; q < 4; ++q) x[q] = (pk >> (16 * q)) & 0xffffull;
                    }
                }
                const unsigned m = __ballot_sync(FULL, f == 2u);
                const int stop = m ? (__ffs(m) - 1) : 32;  // nearest predecessor with an inclusive prefix
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (lane > stop) x[q] = 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) x[q] += __shfl_xor_sync(FULL, x[q], d);
                    e[q] += x[q];
                }
                if (m) break;
                base -= 32;
            }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) st_relaxed_u64(&desc[tile].incl[q], e[q] + ((tile_tot >> (16 * q)) & 0xffffull));
                st_release_u32(&desc[tile].flag, 2u);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) s_excl[q] = e[q];
            if (s_used_tot) atomicAdd((unsigned long long *)&counts[DISO_CNT_USED], (unsigned long long)s_used_tot);
        }
    } else if (s_any_cells) {
        // Meanwhile warps 1..7 flush the staged per-cell words (they do not depend on the prefix): 16 words (32 cells)
        // per chunk, coalesced.  A scan tile without used cells is skipped (its words are never read), so sparse
        // surfaces write almost nothing.  (Before, all eight warps waited at the barrier below for warp 0's look-back:
        // ncu showed 19 % of the issue slots stalled on it.)
        unsigned *cout = reinterpret_cast<unsigned *>(C) + (size_t)tile * SCAN_TILE * 16;
        const int rows = min(SCAN_TILE, g.NCH - tile * SCAN_TILE);
        for (int i = tid - 32; i < rows * 16; i += SCAN_TILE - 32) cout[i] = s_cw[(i >> 4) * CW_STRIDE + (i & 15)];
    }
    __syncthreads();
    const unsigned long long base_a = s_excl[0] + (excl_local & 0xffffull);
    const unsigned long long base_b = s_excl[1] + ((excl_local >> 16) & 0xffffull);
    const unsigned long long base_c = s_excl[2] + ((excl_local >> 32) & 0xffffull);
    const unsigned long long base_d = s_excl[3] + (excl_local >> 48);

    // ordered active-chunk lists (ascending chunk id): what the emit kernels iterate over
    if (k < g.NCH) {
        if (na) active[base_c] = (unsigned)k;
        if (nb) active[(size_t)g.NCH + base_d] = (unsigned)k;
    }
    if (k < g.NCH) {
        E[k] = make_uint4((unsigned)base_a, mx, my, mz);
        if (ALG == DISO_ALG_MC) reinterpret_cast<uint2 *>(aux)[k] = make_uint2((unsigned)base_b, used);
        else reinterpret_cast<uint4 *>(aux)[k] = make_uint4((unsigned)base_b, used, lo, hi);
    } else if (k == g.NCH) {
        // slot NCH: totals (also the sentinel "next chunk" of the last real chunk)
        E[k] = make_uint4((unsigned)base_a, 0u, 0u, 0u);
        if (ALG == DISO_ALG_MC) reinterpret_cast<uint2 *>(aux)[k] = make_uint2((unsigned)base_b, 0u);
        else reinterpret_cast<uint4 *>(aux)[k] = make_uint4((unsigned)base_b, 0u, 0u, 0u);
        counts[DISO_CNT_EDGES] = (long long)base_a;
        counts[DISO_CNT_EDGE_CHUNKS] = (long long)base_c;
        counts[DISO_CNT_CELL_CHUNKS] = (long long)base_d;
        if (ALG == DISO_ALG_MC) { counts[DISO_CNT_VERTS] = (long long)base_a; counts[DISO_CNT_FACES] = (long long)base_b; }
        else                    { counts[DISO_CNT_VERTS] = (long long)base_b; counts[DISO_CNT_FACES] = (long long)base_a; }
    }
}

// ------------------------------------------------------------------------------------------
// Diagnostics: dense per-cell case index (see diso_b200_debug_cell_codes in the header).
// ------------------------------------------------------------------------------------------
template <int ALG>
__global__ void debug_codes_kernel(Geo g, const unsigned *__restrict__ S, const unsigned short *__restrict__ C,
                                   unsigned char *__restrict__ codes)
{
    const long long n = (long long)g.PX * g.PY * g.PZ;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int zp = (int)(idx % g.PZ);
    const long long r = idx / g.PZ;
    const int k = (int)(r * g.NC + (zp >> 5)), j = zp & 31;
    const CellWords w = load_cell_words(S, g, k);
    unsigned code = cell_code<ALG>(w, j);                       // unused cells: 0 or 255
    if (bit(used_mask(w), j)) code = C[(size_t)k * 32 + j] & 0xffu;  // used cells: the stored (DMC: flipped) case
    codes[idx] = (unsigned char)code;
}

}  // namespace diso
