"""Compact per-kernel table of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv > raw.csv`, then
`python tools/ncu_compact.py raw.csv [out.md] [traffic.json]`.  The markdown goes to profiles/, the JSON (DRAM bytes per launch
per kernel family) is what bench.py quotes as `roofline.traffic`."""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem/CTA"),
        ("smsp__inst_executed.sum", "warp insts")]
FAMILY = [("gemm_f16_kernel<256, 4, 1", "gemm_geglu"), ("gemm_f16_kernel", "gemm"), ("attnx_f16_kernel", "attn_cross"),
          ("attn2_f16_kernel", "attn_self"), ("gn_", "groupnorm")]


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path, out_md=None, out_json=None):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | grid | " + " | ".join(n for _, n in COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    traffic = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ih::", "")
        cells = []
        for key, _ in COLS:
            if key in ix and r[ix[key]] not in ("", "n/a"):
                v, u = r[ix[key]], units[ix[key]]
                cells.append(f"{float(v):.2f} {u}".replace(".00 ", " ") if u not in ("%",) else f"{float(v):.1f}")
            else:
                cells.append("-")
        lines.append(f"| `{name}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |")
        fam = next((f for pat, f in FAMILY if pat in r[ix["Kernel Name"]]), None)
        if fam and fam not in traffic:
            rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
            wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            traffic[fam] = {"dram_bytes_per_launch": rd + wr, "kernel": name, "grid": r[ix["Grid Size"]]}
    text = "\n".join(lines)
    print(text)
    if out_md:
        open(out_md, "a").write(text + "\n")
    if out_json:
        for v in traffic.values():
            v["source"] = f"dram__bytes_read.sum + dram__bytes_write.sum of one cold-cache launch, ncu --set full ({out_md or path})"
        json.dump(traffic, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
