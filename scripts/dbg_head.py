import sys, tempfile, ctypes
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from floodsr_b200.engine import EngineB200
from floodsr_b200.h1 import write_h1_model
from floodsr_b200.synth import synth_tile
td = tempfile.mkdtemp(); fp = write_h1_model(Path(td) / "model_infer.onnx", seed=0)
e = EngineB200(fp, precision="fp16")
d, m = synth_tile(0)
for i in range(2):
    try:
        r = e.run_tiles(d[None], m[None]); print("ok", r["prediction_m"].mean())
    except Exception as ex:
        print("ERR", str(ex)[:200])
