"""Per-layer error budget of the precision modes against an fp64 evaluation of the same plan (run on the B200 box).

For every materialised plan tensor of a forward pass: max |engine - fp64| next to the error of the fp32 torch-CPU
evaluation of the same layers (what ONNX Runtime's fp32 arithmetic would show).  Writes gpurun_out/x3_error_budget.txt.

    python tests/x3_error_budget.py [--tiles 2] [--modes fp32,fp32_simt,fp16]
"""

from __future__ import annotations

import argparse
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import plan_exec  # noqa: E402
from floodsr_b200 import graph as G  # noqa: E402
from floodsr_b200.engine import EngineB200  # noqa: E402
from floodsr_b200.h1 import write_h1_model  # noqa: E402
from floodsr_b200.synth import synth_dem, synth_depth  # noqa: E402
from oracle import preprocessing_np as pp  # noqa: E402  (checker only)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=2)
    ap.add_argument("--modes", default="fp32,fp32+hr_simt,fp32_simt,fp16")
    ap.add_argument("--out", default=str(REPO / "gpurun_out" / "x3_error_budget.txt"))
    args = ap.parse_args()
    lines: list[str] = []

    def say(s: str) -> None:
        print(s, flush=True)
        lines.append(s)

    with tempfile.TemporaryDirectory() as td:
        fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
        lm = G.lower_onnx(fp)
        b = args.tiles
        dn = np.stack([pp.scale_depth_log1p(synth_depth(32, 32, seed=20 + i), 5.0) for i in range(b)])
        en = np.stack([pp.normalize_dem(synth_dem(512, 512, seed=20 + i))[0] for i in range(b)])
        ref64 = plan_exec.run_plan(lm, dn, en, dtype=torch.float64, return_all=True)
        ref32 = plan_exec.run_plan(lm, dn, en, dtype=torch.float32, return_all=True)
        for mode in args.modes.split(","):
            prec, _, flag = mode.partition("+")
            if flag == "hr_simt":
                os.environ["FSR_X3_HR_SIMT"] = "1"
            try:
                eng = EngineB200(fp, precision=prec)
            finally:
                os.environ.pop("FSR_X3_HR_SIMT", None)
            out = eng.stage_forward(dn, en)
            say(f"== mode {mode}: output max |err| vs fp64 {np.abs(out - ref64[lm.out_tensor][..., 0]).max():.3e} "
                f"(torch fp32 {np.abs(ref32[lm.out_tensor] - ref64[lm.out_tensor]).max():.3e}), "
                f"vs torch fp32 {np.abs(out - ref32[lm.out_tensor][..., 0]).max():.3e}")
            for op in lm.ops:
                t = op.dst
                if t == lm.out_tensor:
                    continue
                try:
                    got = eng.debug_tensor(t, b)
                except Exception as exc:  # not materialised in this mode (fused away)
                    say(f"   {G._OP_NAMES[op.kind]:8s} t{t:<3d} {str(lm.tensors[t]):18s} not materialised ({str(exc)[:40]})")
                    continue
                r = ref64[t]
                say(f"   {G._OP_NAMES[op.kind]:8s} t{t:<3d} {str(lm.tensors[t]):18s} |ref|max {np.abs(r).max():9.3e}  "
                    f"err {np.abs(got - r).max():9.3e}  torch-fp32 err {np.abs(ref32[t] - r).max():9.3e}")
            eng.close()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
