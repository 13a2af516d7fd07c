#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED reference (SarahWeiii/diso v0.1.4, installed
from /root/reference into baseline/_ref by __graft_entry__.build()) on a CUDA GPU.

The reference ships no golden data and has no CPU path, so its outputs on
a B200 are the only way to pin the CPU oracle.  Run on the GPU box:

    python tests/golden/make_golden.py --out gpurun_out/golden

then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  Inputs are regenerated
deterministically from tests/cases.py (CPU generators); a sha256 of the input bytes is stored
in each fixture so drift is detected.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402

GOLDEN_CASES = ["sphere32", "sphere64", "roundcube32_def", "rand_flexi_24", "rand_dense_19", "rand_sparse_36",
                "ragged_5x9x70", "ragged_31x2x30", "ragged_3x4x62", "tiny_1x1x1", "tiny_2x2x2", "thin_1x7x33",
                "iso_0p37", "iso_neg", "ties_int", "plane_x", "plane_z_tie", "all_inside_but_one",
                "boundary_negative"]
F64_CASES = {"sphere32", "rand_flexi_24", "ragged_5x9x70", "ties_int", "iso_0p37"}
UNNORMALISED_CASES = {"sphere32", "ragged_5x9x70", "thin_1x7x33"}


def load_reference():
    pkg = os.path.join(ROOT, "baseline", "_ref", "diso")
    spec = importlib.util.spec_from_file_location("diso_ref", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["diso_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


def weights(n, dtype, device):
    """Fixed, non-trivial dL/dverts: loss = sum(verts * w)."""
    i = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
    return torch.cos(i * 0.6180339887 + 0.25).to(dtype).to(device)


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def run_one(module, sdf, deform, iso, normalize, **kw):
    s = sdf.clone().requires_grad_(True)
    d = deform.clone().requires_grad_(True) if deform is not None else None
    verts, faces = module(s, d, isovalue=iso, normalize=normalize, **kw)
    out = dict(verts=verts.detach().cpu().numpy(), faces=faces.cpu().numpy())
    if verts.shape[0] > 0 and verts.requires_grad:
        (verts * weights(verts.shape[0], verts.dtype, verts.device)).sum().backward()
        out["gsdf"] = s.grad.cpu().numpy()
        if d is not None:
            out["gdef"] = d.grad.cpu().numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    ref = load_reference()
    dev = "cuda:0"
    summary = {}
    for name in GOLDEN_CASES:
        for dtype in ([torch.float32, torch.float64] if name in F64_CASES else [torch.float32]):
            tag = "f32" if dtype == torch.float32 else "f64"
            sdf, deform, iso = cases.make(name, dtype)
            arrays = {}
            meta = dict(case=name, dtype=tag, iso=iso, shape=list(sdf.shape), has_deform=deform is not None,
                        sdf_sha256=sha(sdf.numpy()), deform_sha256=sha(deform.numpy()) if deform is not None else None,
                        reference="SarahWeiii/diso v0.1.4 (unmodified, baseline/_ref), torch %s, %s"
                                  % (torch.__version__, torch.cuda.get_device_name(0)))
            sd = sdf.to(dev)
            df = deform.to(dev) if deform is not None else None
            mc, dmc = ref.DiffMC(dtype=dtype), ref.DiffDMC(dtype=dtype)
            variants = [("mc", mc, {}, True), ("dmcq", dmc, dict(return_quads=True), True)]
            if tag == "f32":
                variants.append(("dmct", dmc, dict(return_quads=False), True))
            if name in UNNORMALISED_CASES:
                variants += [("mc", mc, {}, False), ("dmcq", dmc, dict(return_quads=True), False)]
            for key, mod, kw, norm in variants:
                r = run_one(mod, sd, df, iso, norm, **kw)
                pre = "%s_%s" % (key, "n" if norm else "u")
                arrays[pre + "_verts"] = r["verts"]
                arrays[pre + "_faces"] = r["faces"].astype(np.int32) if r["faces"].dtype == np.int64 else r["faces"]
                meta[pre + "_faces_dtype"] = str(r["faces"].dtype)
                if "gsdf" in r:
                    arrays[pre + "_gsdf"] = r["gsdf"]
                if "gdef" in r:
                    arrays[pre + "_gdef"] = r["gdef"]
            arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
            path = os.path.join(args.out, "%s_%s.npz" % (name, tag))
            np.savez_compressed(path, **arrays)
            summary["%s_%s" % (name, tag)] = dict(mc_verts=int(arrays["mc_n_verts"].shape[0]), mc_tris=int(arrays["mc_n_faces"].shape[0]),
                                                  dmc_verts=int(arrays["dmcq_n_verts"].shape[0]), dmc_quads=int(arrays["dmcq_n_faces"].shape[0]),
                                                  bytes=os.path.getsize(path))
            print(name, tag, summary["%s_%s" % (name, tag)], flush=True)
    with open(os.path.join(args.out, "SUMMARY.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
