"""Quick check of the tensor-core path against the fp32 CUDA path (run on a B200): errors + per-stage timing."""
import sys, tempfile, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from floodsr_b200.engine import EngineB200
from floodsr_b200.h1 import write_h1_model
from floodsr_b200.synth import synth_tile
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
td = tempfile.mkdtemp(); fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
e32 = EngineB200(fp, precision="fp32")
e = EngineB200(fp, precision=prec)
tiles = [synth_tile(s) for s in range(n)]
d = np.stack([t[0] for t in tiles]); m = np.stack([t[1] for t in tiles])
want = e32.run_tiles(d, m)
for i in range(2):
    try:
        r = e.run_tiles(d, m)
        err = np.abs(r["prediction_m"] - want["prediction_m"])
        print("ok", prec, "max err m", err.max(), "norm err", np.abs(r["prediction_norm"] - want["prediction_norm"]).max(),
              "flips", int(((r["prediction_m"] > 0.01) != (want["prediction_m"] > 0.01)).sum()))
        if err.max() > 1e-2:
            bad = np.argwhere(err > 1e-2)
            print("  bad count", len(bad), "first", bad[:5].tolist(), "rows hist", np.bincount(bad[:, 1] % 64, minlength=64).tolist())
    except Exception as ex:
        print("ERR", str(ex)[:300])
        break
