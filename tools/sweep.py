"""HarmonyBench-shape sweep (BASELINE config 5, 1 GPU slice): res x images -> denoise-steps/s, ms/step, TFLOP/s.
Writes gpurun_out/sweep.json; summarised in profiles/r1_sweep.md."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from imagharmony_b200.config import SDXL_BASE as cfg  # noqa: E402
from imagharmony_b200.denoise import DenoiseEngine  # noqa: E402


def main():
    torch.cuda.set_device(0)
    unet = bench.build_native(cfg, torch.device("cuda", 0))
    eng = DenoiseEngine(unet)
    rows = []
    K = 20
    for res in (512, 768, 1024):
        for n in (1, 4, 8):
            lat = res // 8
            ins = [t.pin_memory() for t in bench.synth_inputs(cfg, n, lat, K, 0)]
            eng.run(*ins, K, stop_after=3)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            eng.run(*ins, K)
            e.record()
            torch.cuda.synchronize()
            sec = s.elapsed_time(e) * 1e-3
            tf = bench.TFLOP_PER_PAIR.get(lat, 13.524 * (lat / 128.0) ** 2) * n
            rows.append({"res": res, "images": n, "unet_batch": 2 * n, "steps": K, "ms_per_step": sec / K * 1e3,
                         "denoise_steps_per_s": n * K / sec, "tflops": tf * K / sec})
            print(json.dumps(rows[-1]), flush=True)
            eng._graphs.clear()
            eng._static.clear()
            torch.cuda.empty_cache()
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
