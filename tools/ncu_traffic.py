#!/usr/bin/env python
"""Writes profiles/dominant_traffic.json from a kept `ncu --set full` report: DRAM bytes read / written by ONE launch of the
bench's dominant kernel (the FFN-up GEMM with the bias + GELU epilogue, gemm_kernel<256, 0, 0, 1, 1>) at the bench batch.
bench.py reports their sum as `roofline.traffic` when the batch matches.
    python tools/ncu_traffic.py gpurun_out/r2_prof_x2.ncu-rep --workload visual_bert --batch 166 --kernel "gemm_kernel<256, 0, 0, 1, 1>"
"""
import argparse
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", default="visual_bert")
    ap.add_argument("--batch", type=int, default=166)
    ap.add_argument("--kernel", default="gemm_kernel<256, 0, 0, 1, 1>")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    ri, wi, ti = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("gpu__time_duration.sum")
    hit = [r for r in rows[2:] if args.kernel in r[ki]]
    if not hit:
        raise SystemExit("kernel %r not in %s" % (args.kernel, args.report))
    r = hit[0]
    path = os.path.join(ROOT, "profiles", "dominant_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
    except Exception:
        d = {}
    d[args.workload] = {"batch": args.batch, "kernel": r[ki][:120],
                        "dram_bytes_read": float(r[ri]) * UNIT[units[ri]], "dram_bytes_write": float(r[wi]) * UNIT[units[wi]],
                        "duration_us_under_ncu": float(r[ti]), "source": "ncu --set full, " + os.path.basename(args.report)}
    with open(path, "w") as fh:
        json.dump(d, fh, indent=1)
    print(json.dumps(d[args.workload], indent=1))


if __name__ == "__main__":
    main()
