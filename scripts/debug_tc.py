"""Layer-by-layer comparison of the bf16 tensor-core backend against the fp32 backend (run on a B200)."""
import sys, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from floodsr_b200.engine import EngineB200
from floodsr_b200.h1 import write_h1_model
from floodsr_b200.synth import synth_tile
from floodsr_b200.graph import _OP_NAMES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
td = tempfile.mkdtemp()
fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
e32 = EngineB200(fp, precision="fp32")
e16 = EngineB200(fp, precision=sys.argv[2] if len(sys.argv) > 2 else "bf16")
tiles = [synth_tile(s) for s in range(n)]
depth = np.stack([t[0] for t in tiles]); dem = np.stack([t[1] for t in tiles])
nrm = e32.stage_normalize(depth, dem)
p32 = e32.stage_forward(nrm["depth_norm"], nrm["dem_norm"])
p16 = e16.stage_forward(nrm["depth_norm"], nrm["dem_norm"])
lm = e32.lowered
for i, op in enumerate(lm.ops):
    t = op.dst
    if t == lm.out_tensor:
        continue
    h, w, c = lm.tensors[t]
    nt = min(n, 4) if h * w * c > 512 * 512 else n
    try:
        a = e32.debug_tensor(t, nt); b = e16.debug_tensor(t, nt)
    except ValueError as exc:  # tensors the fused high-resolution kernel never materialises
        print(f"op {i:2d} {_OP_NAMES[op.kind]:8s} t{t:<3d} skipped: {exc}")
        continue
    err = np.abs(a - b); scale = np.abs(a).max() + 1e-9
    print(f"op {i:2d} {_OP_NAMES[op.kind]:8s} t{t:<3d} {str(lm.tensors[t]):18s} max|a|={scale:9.4f} maxerr={err.max():9.5f} rel={err.max()/scale:8.5f} meanerr={err.mean():9.6f}")
    if err.max() / scale > 0.05:
        bad = np.argwhere(err > 0.05 * scale)
        print("   first bad idx:", bad[:5].tolist(), " a:", a[tuple(bad[0])], " b:", b[tuple(bad[0])])
print("final pred_norm: max err", np.abs(p32 - p16).max(), "mean", np.abs(p32 - p16).mean(), "range", p32.min(), p32.max())
m32 = e32.stage_invert(p32); 
r16 = e16.run_tiles(depth, dem)
print("run_tiles bf16 vs fp32 metres: max err", np.abs(r16["prediction_m"] - m32).max(), "norm err", np.abs(r16["prediction_norm"] - p32).max())
wet32 = m32 > 0.01; wet16 = r16["prediction_m"] > 0.01
print("wet/dry mismatches:", int((wet32 != wet16).sum()), "of", wet32.size)
