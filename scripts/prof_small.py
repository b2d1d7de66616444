"""Small workload for ncu / pipeline stats: batches of tiles through EngineB200.run_tiles (run on a B200)."""
import sys, tempfile, time
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
tiles = [synth_tile(s % 4) for s in range(min(n, 4))]
depth = np.stack([tiles[i % len(tiles)][0] for i in range(n)]); dem = np.stack([tiles[i % len(tiles)][1] for i in range(n)])
eng.run_tiles(depth, dem, want_norm=False)
eng.profile(True)
for _ in range(reps):
    r = eng.run_tiles(depth, dem, want_norm=False)
prof = eng.profile_fetch()
print("ok", float(r["prediction_m"].mean()), eng.launch_count(), {k: round(v[0] / reps / n * 1e3, 2) for k, v in prof.items()}, "us/tile")
