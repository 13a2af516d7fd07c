#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json: Gvoxel/s, forward+backward,
DiffMC and DiffDMC, 512^3 random-init SDF with learnable deform, fp32; % of HBM roofline).

One "step" = DiffMC forward+backward AND DiffDMC (return_quads=True) forward+backward on one
512^3 grid, with the harness of BASELINE.md section 2:
    verts, faces = m(sdf, deform); (verts * w).sum().backward()        (w fixed, random)
`value` counts G = X*Y*Z voxels per extractor pass (2*G per step) over the device-timed region
with inputs resident in HBM; `e2e` is the same step starting from pinned HOST buffers (H2D of sdf
and deform every step, D2H of the two losses).  Multi-GPU: one independent 512^3 shape per rank
(config C5a, seeds differ per rank), no data-path collective -> weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cpu] [--size 512] [--kind flexi]

`--impl reference`: the UNMODIFIED reference (baseline/_ref, SarahWeiii/diso v0.1.4 built for sm_100
by __graft_entry__.build()) through its own public API (diso.DiffMC / diso.DiffDMC), same inputs,
same harness, same timing code as our arm -- the reference is a CUDA library with NO CPU path
(SURVEY.md 8c), so its own implementation of the path runs on the GPU.  Its line also carries a
`cpu_baseline` (the CPU oracle port, oracle/, on a bounded sample).  If baseline/_ref cannot be
loaded the arm falls back to `--impl reference-cpu`: the oracle port on all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gvoxels/s fwd+bwd DiffMC+DiffDMC 512^3 rand SDF"
UNIT = "Gvoxel/s"


def load_reference():
    """The UNMODIFIED reference build (baseline/_ref, installed by __graft_entry__.build()) under the alias
    ``diso_ref``; None when it is not in the snapshot."""
    import importlib.util
    if "diso_ref" in sys.modules:
        return sys.modules["diso_ref"]
    pkg = os.path.join(ROOT, "baseline", "_ref", "diso")
    if not os.path.exists(os.path.join(pkg, "_C.so")):
        return None
    try:
        spec = importlib.util.spec_from_file_location("diso_ref", os.path.join(pkg, "__init__.py"), submodule_search_locations=[pkg])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["diso_ref"] = mod
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        sys.modules.pop("diso_ref", None)
        return None


def bind_to_gpu_numa_node(local_rank, world):
    """Pin this process to the host cores of its GPU's NUMA node BEFORE pinned buffers are allocated (first touch
    places them on that node).  Falls back to an even split of the visible cores.  Returns a description."""
    try:
        import torch
        cpus_all = sorted(os.sched_getaffinity(0))
        node = None
        try:
            prop = torch.cuda.get_device_properties(local_rank)
            bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
            with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
                node = int(f.read().strip())
        except Exception:
            node = None
        cpus = None
        if node is not None and node >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus = sorted(cpus & set(cpus_all)) or None
        how = "numa node %s" % node
        if cpus is None:
            if world <= 1:
                return "unbound (single rank)"
            per = max(1, len(cpus_all) // world)
            cpus = cpus_all[local_rank * per: (local_rank + 1) * per] or cpus_all
            how = "even split (no NUMA information)"
        os.sched_setaffinity(0, cpus)
        return "%s: %d cores" % (how, len(cpus))
    except Exception as ex:
        return "unbound (%s)" % type(ex).__name__


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--kind", default="flexi", choices=["flexi", "sparse", "dense"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference CUDA build beside ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line"), runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self, t0=None, t1=None):
        """Median SM clock / throttle reasons of the samples taken inside [t0, t1] (the timed region)."""
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1)]
        if not rows:  # region shorter than the sampling period: take the samples closest to it
            rows = [r for (t, r) in sorted(self.rows, key=lambda tr: abs(tr[0] - (t0 or 0)))[:3]]
        sm, mx, reasons = [], 0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# algorithmic bytes (BASELINE.md section 3 / DESIGN.md section 5)
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(G, s, V, F, k, E, deform, w=8):
    fwd = G * s + 3 * V * s + k * F * w + (3 * min(E, G) * s if deform else 0)
    bwd = 3 * V * s + min(E, G) * s + G * s + ((3 * min(E, G) * s + 3 * G * s) if deform else 0)
    return fwd, bwd


def kernel_bytes(name, G, s, c_mc, c_dmc, deform):
    """Share of the algorithmic bytes each kernel is responsible for."""
    Vm, Fm, Em = c_mc["verts"], c_mc["faces"], c_mc["endpoints"]
    Vd, Qd = c_dmc["verts"], c_dmc["faces"]
    d3 = 3 if deform else 0
    table = {
        "sign_pack_f32x4": G * s, "sign_pack_f64x2": G * s, "sign_pack": G * s,
        "classify_scan_mc": 0, "classify_scan_dmc": 0,
        "mc_emit_verts": 3 * Vm * s + d3 * min(Em, G) * s,
        "mc_emit_tris": 3 * Fm * 8,
        # one launch: edge pass + triangle pass (CTAs interleaved) / edge crossings + quads
        "mc_emit_fused": 3 * Vm * s + d3 * min(Em, G) * s + 3 * Fm * 8,
        "dmc_emit_cross_quads": d3 * min(Em, G) * s + 4 * Qd * 8,
        "mc_backward": 3 * Vm * s + min(Em, G) * s + G * s + d3 * (min(Em, G) + G) * s,
        "dmc_edge_crossings": d3 * min(Em, G) * s,   # internal pass: reads the deform endpoints
        "dmc_emit_verts": 3 * Vd * s,
        "dmc_emit_quads": 4 * Qd * 8,
        "dmc_edge_adjoint": 3 * Vd * s,
        # fused DMC backward (per-edge adjoint evaluated inside the edge pass): dL/d dual vertices in, dense adjoints out
        "dmc_backward": 3 * Vd * s + min(Em, G) * s + G * s + d3 * (min(Em, G) + G) * s,
    }
    return table.get(name, 0)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def oracle_pass(sdf, deform):
    """MC + DMC forward+backward on one numpy sample with the CPU oracle. Returns voxels processed."""
    import numpy as np
    from oracle import diso_oracle as O
    for alg in ("mc", "dmc"):
        v, _ = O.forward(alg, sdf, deform, 0.0, True)
        w = np.ones_like(v)
        O.backward(alg, sdf, deform, 0.0, True, w, "reference")
    return 2 * sdf.size


def cpu_sample(kind, n, seed, dtype):
    from diso_b200 import synthetic as syn
    import torch
    dt = torch.float32 if dtype == "f32" else torch.float64
    return syn.random_sdf(n, kind, seed, dt).numpy(), syn.random_deform(n, seed + 1, dt).numpy()


def run_reference_cpu_arm(args, rank, world, why=None):
    """CPU oracle port on all host cores; bounded sample per step (see module docstring)."""
    if rank != 0:
        return
    from oracle import diso_oracle as O
    O.lib()
    cores = os.cpu_count() or 1
    n = 64
    samples = [cpu_sample(args.kind, n, 100 + i, args.dtype) for i in range(cores)]

    def step():
        out = [0] * cores

        def work(i):
            out[i] = oracle_pass(*samples[i])  # ctypes releases the GIL inside the C oracle
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        [t.start() for t in th]
        [t.join() for t in th]
        return sum(out)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    vox = 0
    for _ in range(args.steps):
        vox += step()
    dt = time.perf_counter() - t0
    value = vox / dt / 1e9
    sample = "%d concurrent %d^3 crops of the rand-%s workload per step (one per host thread), MC+DMC fwd+bwd each" % (cores, n, args.kind)
    line = {"impl": "reference", "reference_kind": "cpu-oracle-port", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C4: random-init %d^3 SDF (rand-%s) + learnable deform, DiffMC and DiffDMC fwd+bwd" % (args.size, args.kind),
                       "note": "reference has no CPU path; timed: CPU oracle port on a bounded sample" + ((" (fallback: %s)" % why) if why else "")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# C5b: ONE large grid in slabs along dim 0 (diso_b200/parallel.py), timed inside the N > 1 runs
# ------------------------------------------------------------------------------------------------
def _slab_layers(dev, n, xa, xb):
    """Layers [xa, xb) of the n^3 rand-flexi grid + deform, generated on the device from per-layer seeds (no rank ever
    holds the whole grid; the unsharded comparison on rank 0 generates all layers the same way)."""
    import torch

    def layer(x, shape):
        g = torch.Generator(device=dev).manual_seed(1000003 * x + 17)
        return torch.rand(shape, generator=g, device=dev)
    sdf = torch.stack([layer(x, (n, n)) - 0.1 for x in range(xa, xb)])
    deform = torch.stack([0.5 * torch.tanh(layer(x + n, (n, n, 3))) for x in range(xa, xb)])
    return sdf, deform


def run_slab_leg(rank, world, dev, n=1024, steps=3, warmup=2):
    """DiffMC + DiffDMC fwd+bwd of one n^3 grid sharded over the ranks: halo P2P + count all_gather over NCCL, ids and
    vertices stitched by the kernels (frame).  Then rank 0 alone runs the SAME grid unsharded (single call per extractor).
    Returns (on rank 0) the dict attached to the bench line as "slab"."""
    import torch
    import torch.distributed as dist
    import diso_b200
    from diso_b200 import parallel
    # pre-flight, agreed on by all ranks (a rank that failed alone inside the sharded step would leave the others waiting in
    # a collective): enough free memory for the slab step here, and for the unsharded comparison on rank 0
    free, _ = torch.cuda.mem_get_info()
    need = (100e9 * (n / 1024.0) ** 3) / world + 8e9
    if rank == 0:
        need = max(need, 100e9 * (n / 1024.0) ** 3)
    ok = torch.tensor([1 if free >= need else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        return {"skipped": "not enough free device memory for the %d^3 slab leg (rank %d: %.0f GB free)" % (n, rank, free / 1e9)}
    xa, xb = parallel.plan_slabs(n, world)[rank]
    sdf, deform = _slab_layers(dev, n, xa, xb)
    sf, df = parallel.SlabField(sdf, rank, world), parallel.SlabField(deform, rank, world)
    del sdf, deform
    sf.ext.requires_grad_(True)
    df.ext.requires_grad_(True)
    mesh = {}

    def step():
        for alg in ("mc", "dmc"):
            sf.ext.grad = df.ext.grad = None
            # halos are refreshed once per step (the fields do not change between the two extractions of a step)
            verts, faces, info = parallel.extract_slab_ext(alg, sf, df, (xa, xb), n, 0.0, True, refresh=(alg == "mc"))
            (verts * 0.5).sum().backward()
            mesh[alg] = dict(verts=info["n_verts_total"], faces=info["n_faces_total"])
            del verts, faces

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    for _ in range(warmup):
        step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    del sf, df
    torch.cuda.empty_cache()
    out = {"workload": "C5b: one %d^3 grid (rand-flexi + deform, fp32) in %d slabs along dim 0, DiffMC + DiffDMC(quads) fwd+bwd per step; "
                       "2-layer halo P2P + count all_gather over NCCL, global ids / global-frame vertices written by the kernels" % (n, world),
           "n_gpus": world, "ms_per_step": ms, "value": 2 * n ** 3 / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps, "warmup": warmup, "mesh": mesh}
    # the same grid, unsharded, on rank 0 alone (the other ranks wait at the barrier)
    if rank == 0:
        try:
            sdf, deform = _slab_layers(dev, n, 0, n)
            sdf.requires_grad_(True)
            deform.requires_grad_(True)
            mods = ((diso_b200.DiffMC(), {}), (diso_b200.DiffDMC(), dict(return_quads=True)))

            def ustep():
                for m, kw in mods:
                    sdf.grad = deform.grad = None
                    v, f = m(sdf, deform, **kw)
                    (v * 0.5).sum().backward()
                    del v, f
            ustep()
            torch.cuda.synchronize()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            for _ in range(2):
                ustep()
            u1.record()
            torch.cuda.synchronize()
            ums = u0.elapsed_time(u1) / 2
            out.update(unsharded_ms_per_step=ums, speedup_vs_unsharded=ums / ms, strong_eff_vs_unsharded=ums / ms / world)
            del sdf, deform
        except Exception as ex:   # e.g. not enough memory for the whole grid on one GPU
            out["unsharded_ms_per_step"] = None
            out["unsharded_error"] = "%s: %s" % (type(ex).__name__, str(ex)[:200])
        torch.cuda.empty_cache()
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference-cpu":
        run_reference_cpu_arm(args, rank, world)
        return
    is_ref = args.impl == "reference"
    ref_mod = None
    if is_ref:
        why = None
        try:
            import torch
            ref_mod = load_reference()
            if ref_mod is None:
                why = "baseline/_ref not loadable"
            elif not torch.cuda.is_available():
                why = "no CUDA device"
        except Exception as ex:
            why = "%s: %s" % (type(ex).__name__, ex)
        if why:
            run_reference_cpu_arm(args, rank, world, why)
            return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")        # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    import torch
    import torch.distributed as dist
    import diso_b200
    from diso_b200 import _lib, synthetic as syn

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    s_bytes = 4 if args.dtype == "f32" else 8
    n = args.size
    G = n ** 3

    # C5a: one independent shape per rank (seeds differ), generated on the CPU, pinned for the e2e leg
    sdf_h = syn.random_sdf(n, args.kind, seed=rank, dtype=dt).pin_memory()
    def_h = syn.random_deform(n, seed=1000 + rank, dtype=dt).pin_memory()
    sdf_d = sdf_h.to(dev).requires_grad_(True)
    def_d = def_h.to(dev).requires_grad_(True)
    impl_pkg = ref_mod if is_ref else diso_b200
    mods = {"mc": (impl_pkg.DiffMC(dt), {}), "dmc": (impl_pkg.DiffDMC(dt), dict(return_quads=True))}

    # fixed random dL/dverts per extractor (sizes are deterministic for a fixed input)
    wts, counts = {}, {}
    with torch.no_grad():
        for key, (m, kw) in mods.items():
            v, f = m(sdf_d, def_d, **kw)
            gen = torch.Generator(device="cpu").manual_seed(7)
            wts[key] = torch.rand(v.shape, generator=gen, dtype=torch.float32).to(dt).to(dev)
            # crossing edges == MC vertices == DMC quads
            counts[key] = dict(verts=v.shape[0], faces=f.shape[0], edges=v.shape[0] if key == "mc" else f.shape[0])
            del v, f
    # E = grid points incident to >= 1 crossing edge (for the algorithmic-bytes formula)
    with torch.no_grad():
        b = sdf_d.detach() >= 0
        bp = torch.nn.functional.pad(b, (1, 1, 1, 1, 1, 1), value=True)
        inc = torch.zeros_like(bp)
        for ax in range(3):
            a0 = bp.narrow(ax, 0, bp.shape[ax] - 1)
            a1 = bp.narrow(ax, 1, bp.shape[ax] - 1)
            cr = a0 != a1
            inc.narrow(ax, 0, bp.shape[ax] - 1).logical_or_(cr)
            inc.narrow(ax, 1, bp.shape[ax] - 1).logical_or_(cr)
        endpoints = int(inc[1:-1, 1:-1, 1:-1].sum().item())
        del b, bp, inc, a0, a1, cr
    for key in counts:
        counts[key]["endpoints"] = endpoints
    torch.cuda.empty_cache()

    def step(s, d):
        losses = []
        for key, (m, kw) in mods.items():
            s.grad = None
            d.grad = None
            v, f = m(s, d, **kw)
            loss = (v * wts[key]).sum()
            loss.backward()
            losses.append(loss)
        return losses

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    clk = ClockSampler(local_rank)
    clk.__enter__()
    for _ in range(args.warmup):
        step(sdf_d, def_d)
    sync_all()
    import contextlib
    L0 = 0 if is_ref else _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with (contextlib.nullcontext() if is_ref else _lib.kernel_profile()) as prof:
        sync_all()
        t_wall0 = time.time()
        e0.record()
        for _ in range(args.steps):
            step(sdf_d, def_d)
        e1.record()
        sync_all()
        t_wall1 = time.time()
    clk.__exit__()
    ms = e0.elapsed_time(e1)
    launches = None if is_ref else _lib.launch_count() - L0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- end-to-end timing: pinned host inputs -> device every step, loss read back -------------
    # Every step uploads ITS inputs (sdf + deform, 2.15 GB) from pinned host memory and reads its two
    # losses back.  Uploads are double-buffered on a copy stream, so step i+1's H2D overlaps step i's
    # kernels (what a training loop with a prefetching loader does); all K copies are inside the timed region.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(sdf_d.detach()), torch.empty_like(def_d.detach())) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[i % 2])
            bufs[i % 2][0].copy_(sdf_h, non_blocking=True)
            bufs[i % 2][1].copy_(def_h, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(k):
        out = []
        for b in range(2):
            freed[b].record()
        upload(0)
        for i in range(k):
            if i + 1 < k:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            s = bufs[i % 2][0].requires_grad_(True)
            d = bufs[i % 2][1].requires_grad_(True)
            losses = step(s, d)
            out.append([float(x.item()) for x in losses])   # D2H read of the step's result
            s.requires_grad_(False); d.requires_grad_(False)
            s.grad = None; d.grad = None
            freed[i % 2].record()
        return out
    e2e_loop(2)
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 6))
    e2.record()
    e2e_loop(e2e_steps)
    e3.record()
    sync_all()
    t = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    del bufs

    # ---- DiffDMC's DEFAULT path (return_quads=False: the quad -> triangle split of diso/__init__.py:117-147) ------
    m_dmc = mods["dmc"][0]

    def dmc_default_step():
        sdf_d.grad = None
        def_d.grad = None
        v, f = m_dmc(sdf_d, def_d)
        (v * wts["dmc"]).sum().backward()
    for _ in range(2):
        dmc_default_step()
    sync_all()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nd = max(3, min(args.steps, 5))
    d0.record()
    for _ in range(nd):
        dmc_default_step()
    d1.record()
    sync_all()
    t = torch.tensor([d0.elapsed_time(d1) / nd], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dmc_default_ms = float(t.item())

    # ---- C5b inside the multi-GPU runs: one 1024^3 grid in slabs (our arm only: the reference has no multi-GPU path) --
    slab = None
    if world > 1 and not is_ref and args.size >= 512 and not os.environ.get("DISO_BENCH_NO_SLAB"):
        del sdf_d, def_d
        wts.clear()
        torch.cuda.empty_cache()
        try:
            slab = run_slab_leg(rank, world, dev)
        except Exception as ex:
            slab = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fwd_mc, bwd_mc = algorithmic_bytes(G, s_bytes, counts["mc"]["verts"], counts["mc"]["faces"], 3, endpoints, True)
    fwd_d, bwd_d = algorithmic_bytes(G, s_bytes, counts["dmc"]["verts"], counts["dmc"]["faces"], 4, endpoints, True)
    step_bytes = fwd_mc + bwd_mc + fwd_d + bwd_d
    ms_step = ms_max / args.steps
    peak, peak_src = peak_hbm()
    roofline, kernels = None, {}
    if not is_ref:
        # ---- per-kernel roofline (dominant kernel, live CUDA-event durations of the timed region) ----
        per_kernel = {k: sum(v) / len(v) for k, v in prof.times.items()}
        total_k = {k: sum(v) / args.steps for k, v in prof.times.items()}
        dom = max(total_k, key=total_k.get)
        kb = kernel_bytes(dom, G, s_bytes, counts["mc"], counts["dmc"], True)
        achieved = kb / (per_kernel[dom] * 1e-3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("%s@%d^3-%s-%s" % (dom, n, args.kind, args.dtype))
        except Exception:
            pass
        for k in sorted(total_k, key=total_k.get, reverse=True):
            b_ = kernel_bytes(k, G, s_bytes, counts["mc"], counts["dmc"], True)
            kernels[k] = {"ms": round(per_kernel[k], 4), "launches_per_step": round(len(prof.times[k]) / args.steps, 2),
                          "share_of_step": round(total_k[k] / ms_step, 4),
                          "alg_GBps": round(b_ / (per_kernel[k] * 1e-3) / 1e9, 1) if b_ else None}
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel_ms": per_kernel[dom], "kernel_alg_bytes": kb,
                    # the same figure for every kernel with algorithmic bytes (dominant = largest total time per step)
                    "frac_by_kernel": {k: round(v["alg_GBps"] / peak, 4) for k, v in kernels.items() if v["alg_GBps"]},
                    "note": "dmc_backward is ONE kernel since round 2 (per-edge adjoint of the dual vertices evaluated inside the edge pass); "
                            "round 1 ran the same work as dmc_edge_adjoint (1.06 ms) + mc_backward (1.70 ms)"}

    line = {
        "metric": METRIC, "value": world * 2 * G / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "C4: random-init %d^3 SDF (rand-%s, seed=rank) + learnable deform, DiffMC and DiffDMC(return_quads) fwd+bwd per step" % (n, args.kind),
                   "grid": [n, n, n], "l2": "inputs (%.2f GB) exceed the 126 MB L2" % ((G * s_bytes * 4) / 1e9),
                   "parallelism": "one independent shape per GPU (C5a), no collective"},
        "clocks": clk.summary(t_wall0, t_wall1),
        "e2e": {"value": world * 2 * G / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(sdf_h.numel() * s_bytes + def_h.numel() * s_bytes), "d2h_bytes_per_step": 2 * s_bytes},
        "gpu_launches": launches,
        "dmc_default": {"what": "DiffDMC fwd+bwd through the DEFAULT call (return_quads=False: triangles)", "ms_per_step": dmc_default_ms,
                        "value": world * G / (dmc_default_ms * 1e-3) / 1e9, "unit": UNIT},
        "slab": slab,
        "host": {"affinity": affinity, "h2d_GBps_per_rank": (sdf_h.numel() + def_h.numel()) * s_bytes / (e2e_ms * 1e-3) / 1e9},
        "roofline": roofline,
        "step_roofline": {"alg_bytes_per_step": step_bytes, "achieved_GBps": step_bytes / (ms_step * 1e-3) / 1e9,
                          "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak},
        "kernels": kernels,
        "mesh": {"mc": counts["mc"], "dmc": counts["dmc"]},
    }

    if is_ref:
        line["impl"] = "reference"
        line["reference_kind"] = "unmodified SarahWeiii/diso v0.1.4 CUDA build (baseline/_ref, sm_100) through diso.DiffMC / diso.DiffDMC"
        line.pop("gpu_launches")      # not ours to count
        line.pop("kernels")
    # ---- CPU baseline (rank 0, N=1): single-threaded oracle port on a bounded sample --------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            ncpu = 320   # ~10-15 s of single-threaded CPU work
            smp = cpu_sample(args.kind, ncpu, 0, args.dtype)
            t0 = time.perf_counter()
            vox = oracle_pass(*smp)
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": vox / dtc / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                                    "host_cores": os.cpu_count(),
                                    "sample": "%d^3 crop of the rand-%s workload with deform, MC+DMC fwd+bwd, %.1f s" % (ncpu, args.kind, dtc)}
        except Exception as ex:  # the baseline is a reported number, never a dependency
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "failed: %s" % ex}

    # ---- reference CUDA build on the same GPU, same harness (reported beside ours; N=1 only) ------
    if world == 1 and not args.no_ref_cuda and not is_ref:
        try:
            ref = load_reference()
            if ref is None:
                line["ref_cuda"] = {"unavailable": "baseline/_ref not present"}
            else:
                rmods = {"mc": (ref.DiffMC(dt), {}), "dmc": (ref.DiffDMC(dt), dict(return_quads=True))}

                def rstep():
                    for key, (m, kw) in rmods.items():
                        sdf_d.grad = None
                        def_d.grad = None
                        v, f = m(sdf_d, def_d, **kw)
                        (v * wts[key]).sum().backward()
                for _ in range(2):
                    rstep()
                torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                nref = 3
                r0.record()
                for _ in range(nref):
                    rstep()
                r1.record()
                torch.cuda.synchronize()
                rms = r0.elapsed_time(r1) / nref
                rd = rmods["dmc"][0]

                def rdefault():
                    sdf_d.grad = None
                    def_d.grad = None
                    v, f = rd(sdf_d, def_d)
                    (v * wts["dmc"]).sum().backward()
                rdefault()
                torch.cuda.synchronize()
                r2, r3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r2.record()
                for _ in range(2):
                    rdefault()
                r3.record()
                torch.cuda.synchronize()
                rdms = r2.elapsed_time(r3) / 2
                line["dmc_default"]["reference_ms_per_step"] = rdms
                line["dmc_default"]["speedup_device"] = rdms / dmc_default_ms
                line["ref_cuda"] = {"ms_per_step": rms, "value": 2 * G / (rms * 1e-3) / 1e9, "unit": UNIT,
                                    "speedup_device": rms / ms_step,
                                    "what": "unmodified SarahWeiii/diso v0.1.4 built for sm_100 (baseline/_ref), same inputs/harness"}
        except Exception as ex:
            line["ref_cuda"] = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}

    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
