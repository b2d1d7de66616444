"""Small workload for ncu: one batch of tiles through EngineB200.run_tiles (run on a B200)."""
import sys, tempfile
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from floodsr_b200.engine import EngineB200
from floodsr_b200.h1 import write_h1_model
from floodsr_b200.synth import synth_tile

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
td = tempfile.mkdtemp()
fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
eng = EngineB200(fp, precision=prec)
tiles = [synth_tile(s % 4) for s in range(n)]
depth = np.stack([t[0] for t in tiles]); dem = np.stack([t[1] for t in tiles])
for _ in range(reps):
    r = eng.run_tiles(depth, dem, want_norm=False)
print("ok", float(r["prediction_m"].mean()), eng.launch_count())
