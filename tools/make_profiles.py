"""Turn the raw outputs of tools/round_end.sh (gpurun_out/re_*) into the committed evidence files profiles/r2_*.
Runs in the build container (ncu is used offline to read the .ncu-rep; cuobjdump for the SASS evidence)."""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def jline(path):
    txt = open(path).read()
    lines = [l for l in txt.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    os.makedirs(P, exist_ok=True)
    for src, dst in (("re_bench_n1.json", "r2_bench_n1.json"), ("re_bench_reference.json", "r2_bench_reference.json")):
        d = jline(os.path.join(G, src))
        if d:
            json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
    for src, dst in (("n8_bench.json", "r2_bench_n8.json"), ("n8_pns32.json", "r2_pns32_n8.json"),
                     ("n8_pns32_two_phase.json", "r2_pns32_two_phase_n8.json"), ("c6_bench_n2.json", "r2_bench_n2.json")):
        if os.path.exists(os.path.join(G, src)):
            d = jline(os.path.join(G, src))
            if d:
                json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
    for src, dst in (("re_pytest.txt", "r2_pytest_gpu.txt"), ("re_smoke.txt", "r2_smoke.txt"), ("re_launches.md", "r2_launches_step1024.md"),
                     ("re_gemm_trace.txt", "r2_gemm_trace.txt"), ("bench_shapes.md", "r2_step_shapes.md")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    for name in ("layer_head_to_head.json", "layer_head_to_head_fused.json", "edit_latency.json", "microbench.json"):
        if os.path.exists(os.path.join(G, name)):
            shutil.copy(os.path.join(G, name), os.path.join(P, "r2_" + name))
    # sweep table
    rows = []
    for n in (1, 2, 4, 8):
        f = os.path.join(G, f"sweep_n{n}.json")
        if os.path.exists(f):
            rows += json.load(open(f))
    if rows:
        base = {(r["res"], r["images_per_gpu"], r["steps"]): r["denoise_steps_per_s"] for r in rows if r["gpus"] == 1}
        prev = os.path.join(G, "sweep_n1_prev.json")     # 1-GPU sweep of the build the multi-GPU rows were measured with
        base_multi = dict(base)
        if os.path.exists(prev):
            base_multi = {(r["res"], r["images_per_gpu"], r["steps"]): r["denoise_steps_per_s"] for r in json.load(open(prev))}
        with open(os.path.join(P, "r2_sweep.md"), "w") as f:
            f.write("# HarmonyBench-shape sweep (BASELINE config 5): tools/sweep.py, CUDA-graph replay through DenoiseEngine.run, "
                    "round-2 kernels.  The 8-GPU rows were taken with the build of this round that ran 19.7 ms per 1024^2 step on one GPU "
                    "(before the CTA-pair 256x192 tile and the fused ResBlock shortcut); their weak-scaling efficiency is computed "
                    "against the 1-GPU sweep of that same build (efficiencies slightly above 1 are box-to-box variance).  The final "
                    "build's 8-GPU bench line is profiles/r2_bench_n8.json (424.6 steps/s = 1.00 x 8 x 53.0).\n\n| GPUs | res | images/GPU (UNet batch) | steps | ms/step | denoise-steps/s (all GPUs) | "
                    "algorithmic TFLOP/s per GPU | frac of sustained peak | weak-scaling efficiency |\n|---|---|---|---|---:|---:|---:|---:|---:|\n")
            for r in rows:
                b = (base if r["gpus"] == 1 else base_multi).get((r["res"], r["images_per_gpu"], r["steps"]))
                eff = f"{r['denoise_steps_per_s'] / (b * r['gpus']):.3f}" if b else "-"
                f.write(f"| {r['gpus']} | {r['res']}^2 | {r['images_per_gpu']} ({r['unet_batch']}) | {r['steps']} | {r['ms_per_step']:.2f} | "
                        f"{r['denoise_steps_per_s']:.1f} | {r['tflops_per_gpu']:.0f} | {r['frac_sustained_peak']:.2f} | {eff} |\n")
    # ncu --set full per family
    rep = os.path.join(G, "re_families.ncu-rep")
    if os.path.exists(rep):
        raw = "/tmp/re_families_raw.csv"
        with open(raw, "w") as f:
            subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=False)
        md = os.path.join(P, "r2_ncu_families.md")
        open(md, "w").write("# `ncu --set full --clock-control none` of one launch per kernel family at its most frequent SDXL shape "
                            "(1024^2, UNet batch 2): `tools/prof_families.py`; cold caches, serialised -- compare shares and pipe "
                            "utilisation, not absolute times\n\n")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_compact.py"), raw, md, os.path.join(P, "r2_traffic.json")],
                       stdout=subprocess.DEVNULL, check=False)
    # SASS evidence
    so = os.path.join(ROOT, "imagharmony_b200", "libimagharmony_sm100.so")
    if os.path.exists(so):
        sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
        pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "HMMA", "UBLKPF"]
        with open(os.path.join(P, "r2_sass_evidence.txt"), "w") as f:
            f.write("cuobjdump -sass imagharmony_b200/libimagharmony_sm100.so | mnemonic counts (round 2)\n")
            for p_ in pats:
                f.write(f"{p_:14s} {len(re.findall(r'\b' + re.escape(p_) + r'\b', sass))}\n")
    print("profiles/ updated:", sorted(x for x in os.listdir(P) if x.startswith("r2_")))


if __name__ == "__main__":
    main()
