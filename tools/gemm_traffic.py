#!/usr/bin/env python
"""profiles/gemm_traffic.json from an `ncu --set full` capture of four consecutive vision-layer GEMM launches
(qkv, attn_out, mlp_up, mlp_down) of one Base x 64-clip bench step:

    ncu --set full --clock-control none -k regex:gemm_tcgen05_kernel -s 476 -c 4 -o gpurun_out/prof_gemm \\
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs          # on the GPU box
    ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > /tmp/gemm_raw.csv            # here
    python tools/gemm_traffic.py /tmp/gemm_raw.csv <tag>

(-s 476 = 3 warm-up steps x 153 GEMM launches + the patch embed + 4 layers x 4: the launches are those of layer 4.)
Writes the DRAM bytes per launch, the tensor-pipe activity and the commit the capture was taken at; bench.py reads
`dram_bytes_per_launch_avg` for `roofline.traffic`."""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
    hdr = rows[0]
    names = ["qkv", "attn_out", "mlp_up", "mlp_down"]
    out = {"source": f"ncu --set full, 4 consecutive vision-layer GEMM launches (layer 4) of one Base x64 bench step, capture {tag}",
           "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip(),
           "dram_bytes_per_launch": {}, "launches": []}
    unit = dict(zip(hdr, rows[1]))

    def to_bytes(v, u):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v) * mult

    for name, r in zip(names, rows[2:6]):
        d = dict(zip(hdr, r))
        rd = to_bytes(d["dram__bytes_read.sum"], unit["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], unit["dram__bytes_write.sum"])
        out["dram_bytes_per_launch"][name] = rd + wr
        out["launches"].append({"gemm": name, "kernel": d.get("Kernel Name", "")[:90], "duration_us": float(d["gpu__time_duration.sum"]),
                                "tensor_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
                                "dram_read_bytes": rd, "dram_write_bytes": wr,
                                "sm_ghz": float(d["sm__cycles_elapsed.avg.per_second"])})
    vals = list(out["dram_bytes_per_launch"].values())
    out["dram_bytes_per_launch_avg"] = sum(vals) / len(vals)
    (ROOT / "profiles" / "gemm_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
