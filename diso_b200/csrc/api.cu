// api.cu -- C ABI of libdiso_b200.so: argument checking, state layout, kernel launches.
// See include/diso_b200.h for the contract and the reference interfaces each entry replaces.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <map>
#include <string>
#include <vector>

#include "classify.cuh"
#include "compact.cuh"
#include "dmc_compact.cuh"
#include "mc_backward_compact.cuh"
#include "mc_backward_v2.cuh"
#include "quad_split.cuh"

using namespace diso;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) return fail(DISO_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define CU_LAUNCH_CHECK(name)                                                                        \
    do {                                                                                             \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess) return fail(DISO_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

int check_dims(int alg, int dtype, int X, int Y, int Z)
{
    if (alg != DISO_ALG_MC && alg != DISO_ALG_DMC) return fail(DISO_E_INVALID, "unknown alg %d", alg);
    if (dtype != DISO_F32 && dtype != DISO_F64) return fail(DISO_E_INVALID, "unknown dtype %d", dtype);
    if (X < 1 || Y < 1 || Z < 1) return fail(DISO_E_INVALID, "grid dims must be >= 1 (got %d,%d,%d)", X, Y, Z);
    // 32-bit budget: padded point count and worst-case crossing-edge count must fit in int32/uint32
    const double pts = (double)(X + 2) * (Y + 2) * (Z + 2);
    if (pts >= 1.4e9 || Z + 2 >= (1 << 28)) return fail(DISO_E_TOOLARGE, "grid %dx%dx%d exceeds the per-call index budget; shard it into slabs", X, Y, Z);
    return DISO_OK;
}

struct StatePtrs {
    long long *counts;
    unsigned *ticket;
    TileDesc *desc;
    unsigned *S;
    uint4 *E;
    void *aux;
    unsigned short *C;
    unsigned *active;     // [0, NCH): chunks owning crossing edges, [NCH, 2 NCH): chunks with faces (ascending)
};

StatePtrs state_ptrs(void *state, const StateLayout &L)
{
    char *b = static_cast<char *>(state);
    StatePtrs p;
    p.counts = reinterpret_cast<long long *>(b + L.off_counts);
    p.ticket = reinterpret_cast<unsigned *>(b + L.off_ticket);
    p.desc = reinterpret_cast<TileDesc *>(b + L.off_desc);
    p.S = reinterpret_cast<unsigned *>(b + L.off_sign);
    p.E = reinterpret_cast<uint4 *>(b + L.off_erec);
    p.aux = b + L.off_aux;
    p.C = reinterpret_cast<unsigned short *>(b + L.off_cell);
    p.active = reinterpret_cast<unsigned *>(b + L.off_active);
    return p;
}

// Frame of a slab inside a larger grid (diso_b200_frame; standalone: x_origin 0, X_global X, id_offset 0)
struct Frame { int x_origin; int X_global; long long id_offset; };
inline Frame make_frame(const diso_b200_frame *f, int X)
{
    Frame r{0, X, 0};
    if (f) { r.x_origin = f->x_origin; r.X_global = f->X_global > 0 ? f->X_global : X; r.id_offset = (long long)f->id_offset; }
    return r;
}

template <typename T> EpilogueC<T> make_epilogue_c(const Geo &g, const Frame &fr, int normalize, bool shift = true)
{
    EpilogueC<T> e;
    e.dx = T(fr.X_global) - T(1); e.dy = T(g.Y) - T(1); e.dz = T(g.Z) - T(1);
    e.rx = T(1) / e.dx; e.ry = T(1) / e.dy; e.rz = T(1) / e.dz;
    const bool zero_div = fr.X_global == 1 || g.Y == 1 || g.Z == 1;
    e.normalize = !shift ? 0 : (normalize ? (zero_div ? 3 : 2) : 1);
    e.x0 = fr.x_origin;
    return e;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// launch geometry of the compacted emit kernels: over the ordered active-chunk list when the caller
// passed the counts it read back and the surface is sparse, else over every chunk ("dense" tiles of 64
// consecutive chunks: no indirection; list == nullptr)
struct TileGrid { int ctas; int n_active; const unsigned *list; };
inline TileGrid tile_grid(const StatePtrs &p, const Geo &g, const int64_t *counts_host, int which)
{
    const TileGrid dense{(g.NCH + CT_CHUNKS - 1) / CT_CHUNKS, g.NCH, nullptr};
    if (!counts_host) return dense;
    const int n = (int)std::min<long long>(std::max<long long>(counts_host[which ? DISO_CNT_CELL_CHUNKS : DISO_CNT_EDGE_CHUNKS], 0), g.NCH);
    if (n == 0) return TileGrid{0, 0, nullptr};
    if ((long long)n * 4 >= (long long)g.NCH * 3) return dense;   // >= 75 % of the chunks are active
    return TileGrid{(n + CT_CHUNKS - 1) / CT_CHUNKS, n, p.active + (size_t)which * g.NCH};
}

inline int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;  // B200
    return n;
}

// ---- per-kernel function attributes, applied once per (device, kernel) ----------------------------
// Preferred shared-memory carve-out in percent of the 228 KB: it decides how much of the SM's 256 KB is
// left as L1 for the gathers.  The defaults come from sweeps on B200; DISO_CARVEOUT_<NAME> overrides
// them for experiments (-1 = leave the driver's choice).
inline int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

inline void kernel_attrs(const void *fn, const char *env, int carveout_pct, size_t dyn_smem = 0)
{
    // (device, kernel) -> largest dynamic shared-memory size opted into so far.  The opt-in is re-applied whenever a
    // call needs more than any earlier one did (a long-row grid after a short-row one), so the outcome no longer
    // depends on the call history; the carve-out preference is set on first use only.
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(std::make_pair(dev, fn));
    const bool first = it == done.end();
    if (!first && dyn_smem <= it->second) return;
    if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (first) {
        const int pct = env_int(env, carveout_pct);
        if (pct >= 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, std::min(pct, 100));
        done[std::make_pair(dev, fn)] = dyn_smem;
    } else {
        it->second = dyn_smem;
    }
    cudaGetLastError();
}

constexpr size_t MAX_OPTIN_SMEM = 227 * 1024;   // sm_100: 227 KB per CTA

// Stream-ordered scratch for lists that are private to one call (the sparse backward's touched-block list): a memory pool
// per device that KEEPS its memory across synchronisations.  (The device's default pool releases everything whenever the
// stream synchronises -- release threshold 0 -- so every call would pay a fresh cudaMalloc: measured 1.14 -> 3.84 ms for the
// sphere 512^3 step.)  Returns nullptr on failure with the error recorded.
inline cudaMemPool_t call_pool()
{
    static std::mutex mu;
    static std::map<int, cudaMemPool_t> pools;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto it = pools.find(dev);
    if (it != pools.end()) return it->second;
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = pool;
    return pool;
}

inline cudaError_t call_scratch(void **ptr, size_t bytes, cudaStream_t st)
{
    cudaMemPool_t pool = call_pool();
    if (!pool) return cudaMallocAsync(ptr, bytes, st);
    return cudaMallocFromPoolAsync(ptr, bytes, pool, st);
}

// ---- tracing: launch counter + optional per-kernel CUDA-event timing (per host thread) --------
std::atomic<long long> g_launches{0};
struct ProfRec { const char *name; cudaEvent_t a, b; };
std::atomic<bool> g_prof_on{false};  // process-wide: autograd runs backward on its own thread
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;

template <typename F> int launch(const char *name, cudaStream_t st, F &&f)
{
    ProfRec r{name, nullptr, nullptr};
    const bool prof = g_prof_on.load(std::memory_order_relaxed);
    if (prof) {
        CU_TRY(cudaEventCreate(&r.a));
        CU_TRY(cudaEventCreate(&r.b));
        CU_TRY(cudaEventRecord(r.a, st));
    }
    f();
    if (prof) {
        CU_TRY(cudaEventRecord(r.b, st));
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(r);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU_LAUNCH_CHECK(name);
    return DISO_OK;
}
#define LAUNCH(name, st, ...)                                   \
    do {                                                        \
        int rc_ = launch(name, st, [&]() { __VA_ARGS__; });     \
        if (rc_) return rc_;                                    \
    } while (0)

// ---- phase 1 ---------------------------------------------------------------------------------
template <typename T>
int count_impl(int alg, const T *sdf, const Geo &g, double iso, const StateLayout &L, const StatePtrs &p, cudaStream_t st)
{
    // header (counts, ticket, tile descriptors) = 0: the one memset of the count phase.  The array tails (sign tail =
    // ones, record tails = 0) are written by the sign pass (classify.cuh:fill_tails) -- three launches fewer per call.
    CU_TRY(cudaMemsetAsync(p.counts, 0, L.off_sign - L.off_counts, st));
    TailFill tf;
    tf.s_tail = p.S + g.NCH; tf.n_s = L.sign_tail;
    tf.e_tail = p.E + g.NCH; tf.n_e = L.rec_tail;
    if (alg == DISO_ALG_MC) { tf.aux_tail = reinterpret_cast<uint2 *>(p.aux) + g.NCH; tf.n_aux = 8; }
    else { tf.aux_tail = reinterpret_cast<uint2 *>(reinterpret_cast<uint4 *>(p.aux) + g.NCH); tf.n_aux = 2 * L.rec_tail; }

    const T isoT = (T)iso;
    const int warps = 8;
    constexpr int VN = 16 / (int)sizeof(T);   // values per 128-bit load
    const int NA = (g.Z + 31) / 32;
    const size_t smem = (size_t)warps * (NA + 2) * 4;   // per-warp staging of one row's aligned sign words
    const bool vec = (g.Z % VN == 0) && ((reinterpret_cast<uintptr_t>(sdf) & 15) == 0) && smem <= MAX_OPTIN_SMEM;
    if (vec) {
        const int ctas = std::min(cdiv(g.NR, warps), sm_count() * 6);  // persistent: warps stride over the rows
        kernel_attrs(reinterpret_cast<const void *>(sign_pack_vec_kernel<T>), "DISO_CARVEOUT_SIGN", -1, smem);
        LAUNCH(sizeof(T) == 4 ? "sign_pack_f32x4" : "sign_pack_f64x2", st, sign_pack_vec_kernel<T><<<ctas, warps * 32, smem, st>>>(sdf, g, isoT, p.S, p.counts, tf));
    } else {
        LAUNCH("sign_pack", st, sign_pack_kernel<T><<<cdiv(g.NR, warps), warps * 32, 0, st>>>(sdf, g, isoT, p.S, p.counts, tf));
    }
    if (alg == DISO_ALG_MC)
        LAUNCH("classify_scan_mc", st, classify_scan_kernel<DISO_ALG_MC><<<L.n_tiles, SCAN_TILE, 0, st>>>(g, p.S, p.E, p.aux, p.C, p.active, p.desc, p.ticket, p.counts));
    else
        LAUNCH("classify_scan_dmc", st, classify_scan_kernel<DISO_ALG_DMC><<<L.n_tiles, SCAN_TILE, 0, st>>>(g, p.S, p.E, p.aux, p.C, p.active, p.desc, p.ticket, p.counts));
    return DISO_OK;
}

template <typename T>
int mc_emit_impl(const T *sdf, const T *deform, const Geo &g, double iso, const StatePtrs &p, const int64_t *counts_host,
                 int normalize, const Frame &fr, T *verts, long long *tris, T *rec, cudaStream_t st)
{
    const T isoT = (T)iso, padv = (T)(iso + 1.0);
    const EpilogueC<T> epi = make_epilogue_c<T>(g, fr, normalize);
    const TileGrid te = tile_grid(p, g, counts_host, 0), tc = tile_grid(p, g, counts_host, 1);
    // Experiment (opt-in, DISO_EMIT_FUSION=1): ONE launch with edge-pass and triangle-pass CTAs interleaved, so that the
    // DRAM-bound pass and the LSU-bound pass share every SM (compact.cuh: mc_emit_fused_kernel).  Measured at 512^3: 2.01 ms
    // against 0.82 + 0.90 ms back to back -- the union of the two shared-memory footprints (28.8 KB x 7 CTAs) leaves ~20 KB of L1
    // and the edge pass's sdf / deform gathers live on L1 capacity.  Off by default.
    static const int fuse = env_int("DISO_EMIT_FUSION", 0);
    if (fuse && te.ctas && tc.ctas && !te.list && !tc.list && te.ctas == tc.ctas) {
        const uint2 *F = reinterpret_cast<const uint2 *>(p.aux);
        if (fr.id_offset != 0) LAUNCH("mc_emit_fused", st, (mc_emit_fused_kernel<T, true><<<2 * te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, epi, p.E, F, p.C, fr.id_offset, verts, rec, tris)));
        else LAUNCH("mc_emit_fused", st, (mc_emit_fused_kernel<T, false><<<2 * te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, epi, p.E, F, p.C, fr.id_offset, verts, rec, tris)));
        return DISO_OK;
    }
    kernel_attrs(reinterpret_cast<const void *>(edge_verts_kernel<T, true>), "DISO_CARVEOUT_EV", -1);
    kernel_attrs(reinterpret_cast<const void *>(edge_verts_kernel<T, false>), "DISO_CARVEOUT_EV", -1);
    if (te.ctas) {
        if (te.list) LAUNCH("mc_emit_verts", st, (edge_verts_kernel<T, true><<<te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, epi, p.E, te.list, te.n_active, verts, rec, deform ? 5 : 2)));
        else LAUNCH("mc_emit_verts", st, (edge_verts_kernel<T, false><<<te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, epi, p.E, te.list, te.n_active, verts, rec, deform ? 5 : 2)));
    }
    if (tc.ctas) {
        const uint2 *F = reinterpret_cast<const uint2 *>(p.aux);
        kernel_attrs(reinterpret_cast<const void *>(mc_tris_kernel<false, false>), "DISO_CARVEOUT_TRIS", -1);
#define DISO_TRIS(LISTED, OFFSET) LAUNCH("mc_emit_tris", st, (mc_tris_kernel<LISTED, OFFSET><<<tc.ctas, CT_THREADS, 0, st>>>(g, p.E, F, p.C, tc.list, tc.n_active, fr.id_offset, tris)))
        if (fr.id_offset != 0) { if (tc.list) DISO_TRIS(true, true); else DISO_TRIS(false, true); }
        else                   { if (tc.list) DISO_TRIS(true, false); else DISO_TRIS(false, false); }
#undef DISO_TRIS
    }
    return DISO_OK;
}

template <typename T>
int dmc_emit_impl(const T *sdf, const T *deform, const Geo &g, double iso, const StatePtrs &p, const int64_t *counts_host,
                  int normalize, const Frame &fr, T *scratch, T *verts, long long *quads, T *rec, unsigned char *qflags, cudaStream_t st)
{
    const T isoT = (T)iso, padv = (T)(iso + 1.0);
    const uint4 *P = reinterpret_cast<const uint4 *>(p.aux);
    const EpilogueC<T> raw = make_epilogue_c<T>(g, fr, 0, false), epic = make_epilogue_c<T>(g, fr, normalize);
    const TileGrid te = tile_grid(p, g, counts_host, 0), tc = tile_grid(p, g, counts_host, 1);
    kernel_attrs(reinterpret_cast<const void *>(edge_verts_kernel<T, true>), "DISO_CARVEOUT_EV", -1);
    kernel_attrs(reinterpret_cast<const void *>(edge_verts_kernel<T, false>), "DISO_CARVEOUT_EV", -1);
    kernel_attrs(reinterpret_cast<const void *>(dmc_dual_verts_kernel<T, true>), "DISO_CARVEOUT_DUAL", -1);
    kernel_attrs(reinterpret_cast<const void *>(dmc_dual_verts_kernel<T, false>), "DISO_CARVEOUT_DUAL", -1);
    kernel_attrs(reinterpret_cast<const void *>(dmc_edges2_kernel<T, 0, true>), "DISO_CARVEOUT_QUAD", -1);
    kernel_attrs(reinterpret_cast<const void *>(dmc_edges2_kernel<T, 0, false>), "DISO_CARVEOUT_QUAD", -1);
    // Experiment (opt-in, DISO_EMIT_FUSION=1): crossings + quads in one launch (dmc_emit_fused_kernel): 1.55 ms against
    // 0.83 + 0.77 ms -- within noise of the two launches, so the separate kernels (and their per-kernel accounting) stay.
    static const int fuse = env_int("DISO_EMIT_FUSION", 0);
    const bool fuse_cq = fuse && !qflags && te.ctas && !te.list;     // dense flavour, quads requested as quads
    if (fuse_cq) {
        if (fr.id_offset != 0) LAUNCH("dmc_emit_cross_quads", st, (dmc_emit_fused_kernel<T, true><<<2 * te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, raw, p.S, p.E, P, p.C, fr.id_offset, scratch, rec, quads)));
        else LAUNCH("dmc_emit_cross_quads", st, (dmc_emit_fused_kernel<T, false><<<2 * te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, raw, p.S, p.E, P, p.C, fr.id_offset, scratch, rec, quads)));
    } else if (te.ctas) {
        if (te.list) LAUNCH("dmc_edge_crossings", st, (edge_verts_kernel<T, true><<<te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, raw, p.E, te.list, te.n_active, scratch, rec, deform ? 6 : 3)));
        else LAUNCH("dmc_edge_crossings", st, (edge_verts_kernel<T, false><<<te.ctas, CT_THREADS, 0, st>>>(sdf, deform, g, isoT, padv, raw, p.E, te.list, te.n_active, scratch, rec, deform ? 6 : 3)));
    }
    if (tc.ctas) {
        if (tc.list) LAUNCH("dmc_emit_verts", st, (dmc_dual_verts_kernel<T, true><<<tc.ctas, CT_THREADS, 0, st>>>(scratch, g, epic, p.E, P, p.C, tc.list, tc.n_active, verts)));
        else LAUNCH("dmc_emit_verts", st, (dmc_dual_verts_kernel<T, false><<<tc.ctas, CT_THREADS, 0, st>>>(scratch, g, epic, p.E, P, p.C, tc.list, tc.n_active, verts)));
    }
    if (te.ctas && !fuse_cq) {
#define DISO_QUADS(LISTED, OFFSET, DIAG) LAUNCH("dmc_emit_quads", st, (dmc_edges2_kernel<T, 0, LISTED, OFFSET, DIAG><<<te.ctas, CT_THREADS, 0, st>>>(g, p.S, p.E, P, p.C, te.list, te.n_active, T(1), T(1), T(1), nullptr, fr.id_offset, quads, nullptr, deform ? 6 : 3, verts, qflags, rec)))
        if (qflags) {
            if (fr.id_offset != 0) { if (te.list) DISO_QUADS(true, true, true); else DISO_QUADS(false, true, true); }
            else                   { if (te.list) DISO_QUADS(true, false, true); else DISO_QUADS(false, false, true); }
        } else {
            if (fr.id_offset != 0) { if (te.list) DISO_QUADS(true, true, false); else DISO_QUADS(false, true, false); }
            else                   { if (te.list) DISO_QUADS(true, false, false); else DISO_QUADS(false, false, false); }
        }
#undef DISO_QUADS
    }
    return DISO_OK;
}

template <typename T, bool HAS_DEF, int BX, int BY>
int launch_bwd_compact(const T *sdf, const T *deform, const Geo &g, T isoT, T padv, T ix, T iy, T iz, const uint4 *E,
                       const T *gsrc, T *adj_sdf, T *adj_deform, bool sparse, cudaStream_t st)
{
    static const int smem_pad = env_int("DISO_BWD_SMEM_PAD", 0);   // experiment knob: caps the CTAs per SM
    const size_t smem = bwd_compact_smem<T, HAS_DEF, BX, BY>() + (size_t)smem_pad;
    const int ntx = cdiv(g.X, BX), nty = cdiv(g.Y, BY);
    const long long nblk = (long long)ntx * nty * g.NC;
    // L1 capacity matters more than occupancy here (the gathers of sdf / deform / adjoints hit L1 ~63 %): with the
    // driver's default carve-out (8 CTAs/SM, ~28 KB L1) the fp32 kernel takes 2.48 ms at 512^3, with 132 KB of
    // shared memory (~124 KB L1) 1.70 ms; fp64: 2.91 ms at 58-65 % vs 4.53 ms at 86 % (sweeps in DESIGN.md)
    const int carve = 58;
    if (sparse) {
        // zero fill + list of the touched blocks + persistent grid over that list (mc_backward_compact.cuh)
        auto kern = mc_backward_queue_kernel<T, HAS_DEF, BX, BY>;
        kernel_attrs(reinterpret_cast<const void *>(kern), "DISO_CARVEOUT_BWD", carve, smem);
        const size_t G = (size_t)g.X * g.Y * g.Z;
        CU_TRY(cudaMemsetAsync(adj_sdf, 0, G * sizeof(T), st));
        if (HAS_DEF) CU_TRY(cudaMemsetAsync(adj_deform, 0, G * 3 * sizeof(T), st));
        unsigned *work = nullptr;   // private to this call, see launch_bwd2
        CU_TRY(call_scratch(reinterpret_cast<void **>(&work), ((size_t)nblk + 16) * sizeof(unsigned), st));
        CU_TRY(cudaMemsetAsync(work, 0, 64, st));
        LAUNCH("mc_backward_mark", st, (bwd_mark_kernel<BX, BY><<<cdiv(nblk, 256), 256, 0, st>>>(g, E, ntx, nty, work)));
        const int ctas = (int)std::min<long long>(nblk, (long long)sm_count() * 6);
        LAUNCH("mc_backward", st, kern<<<ctas, BC_THREADS, smem, st>>>(sdf, deform, g, isoT, padv, ix, iy, iz, E, gsrc, adj_sdf, adj_deform, nty, work));
        CU_TRY(cudaFreeAsync(work, st));
        return DISO_OK;
    }
    auto kern = mc_backward_compact_kernel<T, HAS_DEF, BX, BY>;
    kernel_attrs(reinterpret_cast<const void *>(kern), "DISO_CARVEOUT_BWD", carve, smem);
    // 3-D grid (chunk, y tile, x tile): no index divisions in the kernel; shapes whose tile counts exceed
    // the 65535 limit of grid.y / grid.z fall back to a flat grid decoded with divisions
    const bool flat = nty > 65535 || ntx > 65535;
    const dim3 grid = flat ? dim3((unsigned)nblk, 1, 1) : dim3((unsigned)g.NC, (unsigned)nty, (unsigned)ntx);
    LAUNCH("mc_backward", st, kern<<<grid, BC_THREADS, smem, st>>>(sdf, deform, g, isoT, padv, ix, iy, iz, E, gsrc, adj_sdf,
                                                                 adj_deform, ntx, nty, flat ? 1 : 0));
    return DISO_OK;
}

// v2 backward from saved edge records (mc_backward_v2.cuh).
template <typename T, bool HAS_DEF, int GSRC, int BX, int BY>
int launch_bwd2(const Geo &g, T isoT, T ix, T iy, T iz, const uint4 *E, const T *gsrc, const DmcSrc &dmc, const T *rec,
                T *adj_sdf, T *adj_deform, bool sparse, cudaStream_t st)
{
    const char *name = GSRC >= 2 ? "dmc_backward" : "mc_backward";
    // Shared-memory carve-out: the edge pass keeps 8 (MC) / 15 (DMC, with the dual-vertex adjoint) loads per thread in flight and
    // every pending line occupies L1, so L1 capacity bounds the memory-level parallelism: with the driver's default (all 228 KB
    // shared for 8 CTAs/SM, 28 KB L1) the fp32 MC kernel takes 1.88 ms at 512^3, with 132 KB shared (6 CTAs of 20.6 KB, 96 KB L1)
    // 1.44 ms; DMC: 100-116 KB shared 2.25 ms, 132 KB 2.28, 164 KB 2.43 (sweeps in profiles/r2_backward.md)
    const char *carve_env = GSRC >= 2 ? "DISO_CARVEOUT_DBWD" : "DISO_CARVEOUT_BWD2";
    const int carve = GSRC >= 2 ? 50 : (sizeof(T) == 4 ? 58 : 72);   // fp64 MC: 72 % 2.88 ms, 58 % 2.97 ms
    using L = Bwd2Layout<T, HAS_DEF, (GSRC >= 2), BX, BY>;
    const size_t smem = L::bytes;
    const int ntx = cdiv(g.X, BX), nty = cdiv(g.Y, BY);
    const long long nblk = (long long)ntx * nty * g.NC;
    if (sparse) {
        auto kern = mc_backward2_queue_kernel<T, HAS_DEF, GSRC, BX, BY>;
        kernel_attrs(reinterpret_cast<const void *>(kern), carve_env, carve, smem);
        const size_t G = (size_t)g.X * g.Y * g.Z;
        if (adj_sdf) CU_TRY(cudaMemsetAsync(adj_sdf, 0, G * sizeof(T), st));
        if (HAS_DEF && adj_deform) CU_TRY(cudaMemsetAsync(adj_deform, 0, G * 3 * sizeof(T), st));
        // list of the touched blocks: private to this call (stream-ordered allocation), so backward passes that share one
        // saved state -- retained graphs, several streams -- never race on it
        unsigned *work = nullptr;
        CU_TRY(call_scratch(reinterpret_cast<void **>(&work), ((size_t)nblk + 16) * sizeof(unsigned), st));
        CU_TRY(cudaMemsetAsync(work, 0, 64, st));
        LAUNCH("mc_backward_mark", st, (bwd_mark_kernel<BX, BY><<<cdiv(nblk, 256), 256, 0, st>>>(g, E, ntx, nty, work)));
        const int ctas = (int)std::min<long long>(nblk, (long long)sm_count() * 6);
        LAUNCH(name, st, kern<<<ctas, B2_THREADS, smem, st>>>(g, isoT, ix, iy, iz, E, gsrc, dmc, rec, adj_sdf, adj_deform, nty, work));
        CU_TRY(cudaFreeAsync(work, st));
        return DISO_OK;
    }
    auto kern = mc_backward2_kernel<T, HAS_DEF, GSRC, BX, BY>;
    kernel_attrs(reinterpret_cast<const void *>(kern), carve_env, carve, smem);
    const bool flat = nty > 65535 || ntx > 65535;
    const dim3 grid = flat ? dim3((unsigned)nblk, 1, 1) : dim3((unsigned)g.NC, (unsigned)nty, (unsigned)ntx);
    LAUNCH(name, st, kern<<<grid, B2_THREADS, smem, st>>>(g, isoT, ix, iy, iz, E, gsrc, dmc, rec, adj_sdf, adj_deform,
                                                          ntx, nty, flat ? 1 : 0));
    return DISO_OK;
}

// gsrc_kind (mc_backward_v2.cuh GSRC): 0 per-edge adjoints [n,3], 1 blocked SoA, 2 / 3 DMC fused (gsrc = dL/d dual vertices,
// exact / reference-compatible).  rec != NULL selects the v2 kernel.
template <typename T>
int mc_backward_impl(const T *sdf, const T *deform, const Geo &g, double iso, const StatePtrs &p, const int64_t *counts_host,
                     const T *gsrc, int gsrc_kind, const T *rec, int normalize, int X_global, T *adj_sdf, T *adj_deform,
                     cudaStream_t st, const long long *quads = nullptr, long long id_offset = 0)
{
    const T isoT = (T)iso, padv = (T)(iso + 1.0);
    // chain rule of verts / (dims - 1): multiply by the reciprocal (gradients carry a 1e-5 bar, not bit parity)
    const T ix = normalize ? T(1) / (T(X_global) - T(1)) : T(1), iy = normalize ? T(1) / (T(g.Y) - T(1)) : T(1),
            iz = normalize ? T(1) / (T(g.Z) - T(1)) : T(1);
    // sparse surface (known from the counts the caller read back in the forward): fewer than 1/8 of the chunks own a
    // crossing edge -> zero fill + touched-block list instead of one CTA per block
    static const int force = env_int("DISO_BWD_SPARSE", -1);   // experiment knob: 0 / 1 force the path
    // (and the grid is large enough for its four extra launches to pay: below ~64 k chunks the dense grid is a few waves)
    bool sparse = counts_host && counts_host[DISO_CNT_EDGE_CHUNKS] * 8 < (long long)g.NCH && g.NCH >= 65536;
    if (force >= 0) sparse = force != 0;
    if (rec) {
        DmcSrc dmc{quads, id_offset};
#define DISO_B2(HD, K) launch_bwd2<T, HD, K, BWD2_BX, BWD2_BY>(g, isoT, ix, iy, iz, p.E, gsrc, dmc, rec, adj_sdf, HD ? adj_deform : nullptr, sparse, st)
        if (deform) {
            switch (gsrc_kind) { case 1: return DISO_B2(true, 1); case 2: return DISO_B2(true, 2); case 3: return DISO_B2(true, 3); default: return DISO_B2(true, 0); }
        }
        switch (gsrc_kind) { case 1: return DISO_B2(false, 1); case 2: return DISO_B2(false, 2); case 3: return DISO_B2(false, 3); default: return DISO_B2(false, 0); }
#undef DISO_B2
    }
    // no saved records (callers of the bare ABI, the diso._C shim): v1, which re-gathers sdf / deform
    if (gsrc_kind) return fail(DISO_E_INVALID, "internal: this adjoint source needs the saved edge records");
    if (!adj_sdf || (deform && !adj_deform)) return fail(DISO_E_INVALID, "adj_sdf / adj_deform may only be NULL when edge_rec is given");
    // block shape (common.cuh: BWD_BX x BWD_BY) from sweeps on B200 (512^3 rand-flexi): 4x6 1.70 ms, 3x8 1.71, 4x7 / 4x8 1.73,
    // 8x4 1.76, 6x8 1.82, 2x8 1.90, 4x4 1.92
    if (deform) return launch_bwd_compact<T, true, BWD_BX, BWD_BY>(sdf, deform, g, isoT, padv, ix, iy, iz, p.E, gsrc, adj_sdf, adj_deform, sparse, st);
    return launch_bwd_compact<T, false, BWD_BX, BWD_BY>(sdf, deform, g, isoT, padv, ix, iy, iz, p.E, gsrc, adj_sdf, adj_deform, sparse, st);
}

template <typename T>
int dmc_backward_impl(const T *sdf, const T *deform, const Geo &g, double iso, const StatePtrs &p, const int64_t *counts_host,
                      const T *adj_verts, const T *rec, const long long *quads, long long id_offset, int normalize, int X_global, int grad_mode,
                      T *scratch, T *adj_sdf, T *adj_deform, cudaStream_t st)
{
    const uint4 *P = reinterpret_cast<const uint4 *>(p.aux);
    const T ix = normalize ? T(1) / (T(X_global) - T(1)) : T(1), iy = normalize ? T(1) / (T(g.Y) - T(1)) : T(1),
            iz = normalize ? T(1) / (T(g.Z) - T(1)) : T(1);
    if (rec) {
        // saved records: ONE kernel, like the reference's adj_create_dmc_verts (cudualmc.cu:957-1005): the per-edge adjoint
        // is evaluated inside the edge pass of mc_backward2 (no per-edge array, one edge list instead of two)
        static const int unfused = env_int("DISO_DMC_BWD_UNFUSED", 0);   // experiment knob: stage A as its own kernel
        if (!unfused && quads)
            return mc_backward_impl<T>(sdf, deform, g, iso, p, counts_host, adj_verts, grad_mode == DISO_GRAD_EXACT ? 2 : 3, rec, normalize,
                                       X_global, adj_sdf, adj_deform, st, quads, id_offset);
        if (!scratch) return fail(DISO_E_INVALID, "scratch required");
    }
    const TileGrid te = tile_grid(p, g, counts_host, 0);
    // stage A writes the per-edge adjoints in blocked SoA form when stage B is the v2 kernel (coalesced on both sides)
    const int g_soa = rec ? 1 : 0;
    if (te.ctas) {
#define DISO_ADJ(MODE, LISTED) { kernel_attrs(reinterpret_cast<const void *>(dmc_edges2_kernel<T, MODE, LISTED>), "DISO_CARVEOUT_ADJ", -1); LAUNCH("dmc_edge_adjoint", st, (dmc_edges2_kernel<T, MODE, LISTED><<<te.ctas, CT_THREADS, 0, st>>>(g, p.S, p.E, P, p.C, te.list, te.n_active, ix, iy, iz, adj_verts, 0ll, nullptr, scratch, g_soa))); }
        if (grad_mode == DISO_GRAD_EXACT) { if (te.list) DISO_ADJ(1, true) else DISO_ADJ(1, false) }
        else                              { if (te.list) DISO_ADJ(2, true) else DISO_ADJ(2, false) }
#undef DISO_ADJ
    }
    return mc_backward_impl<T>(sdf, deform, g, iso, p, counts_host, scratch, g_soa, rec, 0, g.X, adj_sdf, adj_deform, st);
}

}  // namespace

extern "C" {

int diso_b200_abi_version(void) { return DISO_B200_ABI_VERSION; }

const char *diso_b200_last_error(void) { return g_err; }

long long diso_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int diso_b200_profile_enable(int on)
{
    g_prof_on.store(on != 0);
    if (!on) {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (auto &r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        g_prof.clear();
    }
    return DISO_OK;
}

int diso_b200_profile_dump(char *buf, size_t cap)
{
    if (!buf || cap == 0) return fail(DISO_E_INVALID, "null buffer");
    std::string out;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof) {
        CU_TRY(cudaEventSynchronize(r.b));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        char line[128];
        snprintf(line, sizeof(line), "%s %.6f\n", r.name, ms);
        out += line;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    if (out.size() + 1 > cap) return fail(DISO_E_INVALID, "profile buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return DISO_OK;
}

size_t diso_b200_state_bytes(int alg, int X, int Y, int Z)
{
    if (check_dims(alg, DISO_F32, X, Y, Z) != DISO_OK) return 0;
    return make_layout(alg, make_geo(X, Y, Z)).total;
}

int diso_b200_state_layout(int alg, int X, int Y, int Z, int64_t *out)
{
    int rc = check_dims(alg, DISO_F32, X, Y, Z);
    if (rc) return rc;
    if (!out) return fail(DISO_E_INVALID, "null pointer");
    const Geo g = make_geo(X, Y, Z);
    const StateLayout L = make_layout(alg, g);
    out[0] = (int64_t)L.off_sign; out[1] = (int64_t)L.off_erec; out[2] = (int64_t)L.off_aux; out[3] = (int64_t)L.off_cell;
    out[4] = g.NC; out[5] = g.NCH; out[6] = (int64_t)L.total; out[7] = g.sX;
    return DISO_OK;
}

int diso_b200_count(int alg, const void *sdf, int dtype, int X, int Y, int Z, double iso, void *state,
                    size_t state_bytes, void *stream)
{
    int rc = check_dims(alg, dtype, X, Y, Z);
    if (rc) return rc;
    if (!sdf || !state) return fail(DISO_E_INVALID, "null pointer");
    const Geo g = make_geo(X, Y, Z);
    const StateLayout L = make_layout(alg, g);
    if (state_bytes < L.total) return fail(DISO_E_STATE, "state buffer too small: %zu < %zu", state_bytes, L.total);
    if (reinterpret_cast<uintptr_t>(state) & 255) return fail(DISO_E_STATE, "state buffer must be 256-byte aligned");
    const StatePtrs p = state_ptrs(state, L);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DISO_F32) return count_impl<float>(alg, static_cast<const float *>(sdf), g, iso, L, p, st);
    return count_impl<double>(alg, static_cast<const double *>(sdf), g, iso, L, p, st);
}

int diso_b200_read_counts(const void *state, int64_t *counts_host, void *stream)
{
    if (!state || !counts_host) return fail(DISO_E_INVALID, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(cudaMemcpyAsync(counts_host, state, DISO_COUNT_SLOTS * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return DISO_OK;
}

int diso_b200_mc_emit(const void *sdf, const void *deform, int dtype, int X, int Y, int Z, double iso,
                      const void *state, const int64_t *counts_host, int normalize, const diso_b200_frame *frame,
                      void *verts, int64_t *tris, void *edge_rec, int64_t edge_rec_stride, void *stream)
{
    int rc = check_dims(DISO_ALG_MC, dtype, X, Y, Z);
    if (rc) return rc;
    if (!sdf || !state || !verts || !tris) return fail(DISO_E_INVALID, "null pointer");
    if (edge_rec && edge_rec_stride < 1) return fail(DISO_E_INVALID, "edge_rec_stride must be >= the number of crossing edges");
    const Geo g = make_geo(X, Y, Z);
    const Frame fr = make_frame(frame, X);
    const StatePtrs p = state_ptrs(const_cast<void *>(state), make_layout(DISO_ALG_MC, g));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DISO_F32)
        return mc_emit_impl<float>(static_cast<const float *>(sdf), static_cast<const float *>(deform), g, iso, p, counts_host, normalize, fr,
                                   static_cast<float *>(verts), reinterpret_cast<long long *>(tris), static_cast<float *>(edge_rec), st);
    return mc_emit_impl<double>(static_cast<const double *>(sdf), static_cast<const double *>(deform), g, iso, p, counts_host, normalize, fr,
                                static_cast<double *>(verts), reinterpret_cast<long long *>(tris), static_cast<double *>(edge_rec), st);
}

int diso_b200_dmc_emit(const void *sdf, const void *deform, int dtype, int X, int Y, int Z, double iso,
                       const void *state, const int64_t *counts_host, int normalize, const diso_b200_frame *frame,
                       void *scratch, void *verts, int64_t *quads, void *edge_rec, int64_t edge_rec_stride, uint8_t *quad_flags,
                       void *stream)
{
    int rc = check_dims(DISO_ALG_DMC, dtype, X, Y, Z);
    if (rc) return rc;
    if (!sdf || !state || !scratch || !verts || !quads) return fail(DISO_E_INVALID, "null pointer");
    if (edge_rec && edge_rec_stride < 1) return fail(DISO_E_INVALID, "edge_rec_stride must be >= the number of crossing edges");
    const Geo g = make_geo(X, Y, Z);
    const Frame fr = make_frame(frame, X);
    const StatePtrs p = state_ptrs(const_cast<void *>(state), make_layout(DISO_ALG_DMC, g));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DISO_F32)
        return dmc_emit_impl<float>(static_cast<const float *>(sdf), static_cast<const float *>(deform), g, iso, p, counts_host, normalize, fr,
                                    static_cast<float *>(scratch), static_cast<float *>(verts), reinterpret_cast<long long *>(quads), static_cast<float *>(edge_rec), quad_flags, st);
    return dmc_emit_impl<double>(static_cast<const double *>(sdf), static_cast<const double *>(deform), g, iso, p, counts_host, normalize, fr,
                                 static_cast<double *>(scratch), static_cast<double *>(verts), reinterpret_cast<long long *>(quads), static_cast<double *>(edge_rec), quad_flags, st);
}

int diso_b200_mc_backward(const void *sdf, const void *deform, int dtype, int X, int Y, int Z, double iso,
                          void *state, const int64_t *counts_host, const void *adj_verts, int normalize,
                          const diso_b200_frame *frame, const void *edge_rec, int64_t edge_rec_stride, void *adj_sdf,
                          void *adj_deform, void *stream)
{
    int rc = check_dims(DISO_ALG_MC, dtype, X, Y, Z);
    if (rc) return rc;
    if (!sdf || !state || !adj_verts) return fail(DISO_E_INVALID, "null pointer");
    if (!deform && adj_deform) return fail(DISO_E_INVALID, "adj_deform given without deform");
    if (!edge_rec && (!adj_sdf || (deform != nullptr) != (adj_deform != nullptr)))
        return fail(DISO_E_INVALID, "without edge_rec, adj_sdf is required and adj_deform must be given iff deform is");
    if (edge_rec && edge_rec_stride < 1) return fail(DISO_E_INVALID, "edge_rec_stride must be >= the number of crossing edges");
    if (!adj_sdf && !adj_deform) return DISO_OK;   // nothing to compute
    const Geo g = make_geo(X, Y, Z);
    const Frame fr = make_frame(frame, X);
    const StatePtrs p = state_ptrs(state, make_layout(DISO_ALG_MC, g));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DISO_F32)
        return mc_backward_impl<float>(static_cast<const float *>(sdf), static_cast<const float *>(deform), g, iso, p, counts_host,
                                       static_cast<const float *>(adj_verts), 0, static_cast<const float *>(edge_rec),
                                       normalize, fr.X_global, static_cast<float *>(adj_sdf), static_cast<float *>(adj_deform), st);
    return mc_backward_impl<double>(static_cast<const double *>(sdf), static_cast<const double *>(deform), g, iso, p, counts_host,
                                    static_cast<const double *>(adj_verts), 0, static_cast<const double *>(edge_rec),
                                    normalize, fr.X_global, static_cast<double *>(adj_sdf), static_cast<double *>(adj_deform), st);
}

int diso_b200_dmc_backward(const void *sdf, const void *deform, int dtype, int X, int Y, int Z, double iso,
                           void *state, const int64_t *counts_host, const void *adj_verts, int normalize,
                           const diso_b200_frame *frame, int grad_mode, const void *edge_rec, int64_t edge_rec_stride,
                           const int64_t *quads, void *scratch, void *adj_sdf, void *adj_deform, void *stream)
{
    int rc = check_dims(DISO_ALG_DMC, dtype, X, Y, Z);
    if (rc) return rc;
    if (!sdf || !state || !adj_verts || (!scratch && !(edge_rec && quads))) return fail(DISO_E_INVALID, "null pointer");
    if (!deform && adj_deform) return fail(DISO_E_INVALID, "adj_deform given without deform");
    if (!edge_rec && (!adj_sdf || (deform != nullptr) != (adj_deform != nullptr)))
        return fail(DISO_E_INVALID, "without edge_rec, adj_sdf is required and adj_deform must be given iff deform is");
    if (edge_rec && edge_rec_stride < 1) return fail(DISO_E_INVALID, "edge_rec_stride must be >= the number of crossing edges");
    if (grad_mode != DISO_GRAD_REFERENCE && grad_mode != DISO_GRAD_EXACT) return fail(DISO_E_INVALID, "unknown grad_mode %d", grad_mode);
    if (!adj_sdf && !adj_deform) return DISO_OK;
    const Geo g = make_geo(X, Y, Z);
    const Frame fr = make_frame(frame, X);
    const StatePtrs p = state_ptrs(state, make_layout(DISO_ALG_DMC, g));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == DISO_F32)
        return dmc_backward_impl<float>(static_cast<const float *>(sdf), static_cast<const float *>(deform), g, iso, p, counts_host,
                                        static_cast<const float *>(adj_verts), static_cast<const float *>(edge_rec),
                                        reinterpret_cast<const long long *>(quads), fr.id_offset,
                                        normalize, fr.X_global, grad_mode, static_cast<float *>(scratch),
                                        static_cast<float *>(adj_sdf), static_cast<float *>(adj_deform), st);
    return dmc_backward_impl<double>(static_cast<const double *>(sdf), static_cast<const double *>(deform), g, iso, p, counts_host,
                                     static_cast<const double *>(adj_verts), static_cast<const double *>(edge_rec),
                                     reinterpret_cast<const long long *>(quads), fr.id_offset,
                                     normalize, fr.X_global, grad_mode, static_cast<double *>(scratch),
                                     static_cast<double *>(adj_sdf), static_cast<double *>(adj_deform), st);
}

// scratch layout: [0,256) u64 total (number of config-1 quads), u32 ticket at byte 8 ; QuadTileDesc[tiles] ;
// tile offsets (u32 each) ; flags (1 byte per quad); every part 256-byte aligned
struct QuadScratch { size_t off_desc, off_tile, off_flags, total; };
static QuadScratch quad_scratch_layout(int64_t n_quads)
{
    const size_t tiles = (size_t)((n_quads + QS_TILE - 1) / QS_TILE);
    QuadScratch q;
    q.off_desc = 256;
    q.off_tile = q.off_desc + align_up(tiles * sizeof(QuadTileDesc) + 16, 256);
    q.off_flags = q.off_tile + align_up(tiles * 4 + 4, 256);
    q.total = q.off_flags + align_up((size_t)n_quads + 1, 256);
    return q;
}

size_t diso_b200_quad_split_scratch_bytes(int64_t n_quads)
{
    if (n_quads < 0) return 0;
    return quad_scratch_layout(n_quads).total;
}

int diso_b200_quad_split(const void *verts, int dtype, const int64_t *quads, int64_t n_quads, const uint8_t *quad_flags, void *scratch,
                         int64_t *faces, void *stream)
{
    if (dtype != DISO_F32 && dtype != DISO_F64) return fail(DISO_E_INVALID, "unknown dtype %d", dtype);
    if (n_quads < 0) return fail(DISO_E_INVALID, "negative n_quads");
    if (n_quads == 0) return DISO_OK;
    if (n_quads >= (1ll << 32) - 1) return fail(DISO_E_TOOLARGE, "too many quads for one call");
    if ((!verts && !quad_flags) || !quads || !scratch || !faces) return fail(DISO_E_INVALID, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles = cdiv(n_quads, QS_TILE);
    const QuadScratch L = quad_scratch_layout(n_quads);
    char *b = static_cast<char *>(scratch);
    unsigned long long *total = reinterpret_cast<unsigned long long *>(b);
    unsigned *ticket = reinterpret_cast<unsigned *>(b + 8);
    QuadTileDesc *desc = reinterpret_cast<QuadTileDesc *>(b + L.off_desc);
    unsigned *tile_off = reinterpret_cast<unsigned *>(b + L.off_tile);
    unsigned char *flags = reinterpret_cast<unsigned char *>(b + L.off_flags);
    const long long *qd = reinterpret_cast<const long long *>(quads);
    CU_TRY(cudaMemsetAsync(b, 0, L.off_tile, st));   // total, ticket, tile descriptors
    if (quad_flags) {
        // the diagonal of every quad was decided by dmc_emit_quads: only the scan of the flag bytes is left
        flags = const_cast<unsigned char *>(quad_flags);
        LAUNCH("quad_scan", st, (quad_scan_kernel<<<cdiv(n_quads, QSC_TILE), QS_THREADS, 0, st>>>(flags, n_quads, tile_off, desc, ticket, total)));
    } else if (dtype == DISO_F32) {
        LAUNCH("quad_diag", st, (quad_diag_kernel<float><<<tiles, QS_THREADS, 0, st>>>(static_cast<const float *>(verts), qd, n_quads, flags, tile_off, desc, ticket, total)));
    } else {
        LAUNCH("quad_diag", st, (quad_diag_kernel<double><<<tiles, QS_THREADS, 0, st>>>(static_cast<const double *>(verts), qd, n_quads, flags, tile_off, desc, ticket, total)));
    }
    LAUNCH("quad_emit", st, quad_emit_kernel<<<tiles, QS_THREADS, 0, st>>>(qd, n_quads, flags, tile_off, total, reinterpret_cast<long long *>(faces)));
    return DISO_OK;
}

int diso_b200_debug_cell_codes(int alg, int X, int Y, int Z, const void *state, uint8_t *codes, void *stream)
{
    int rc = check_dims(alg, DISO_F32, X, Y, Z);
    if (rc) return rc;
    if (!state || !codes) return fail(DISO_E_INVALID, "null pointer");
    const Geo g = make_geo(X, Y, Z);
    const StatePtrs p = state_ptrs(const_cast<void *>(state), make_layout(alg, g));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = (long long)g.PX * g.PY * g.PZ;
    if (alg == DISO_ALG_MC)
        LAUNCH("debug_codes", st, debug_codes_kernel<DISO_ALG_MC><<<cdiv(n, 256), 256, 0, st>>>(g, p.S, p.C, codes));
    else
        LAUNCH("debug_codes", st, debug_codes_kernel<DISO_ALG_DMC><<<cdiv(n, 256), 256, 0, st>>>(g, p.S, p.C, codes));
    return DISO_OK;
}

}  // extern "C"
