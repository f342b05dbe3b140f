"""Reduce an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum --csv) of tools/prof_chol.py to
profiles/ncu_traffic.json: average DRAM bytes per launch of the dominant solve kernel over whole iterations.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'cholesky_half|cg_rows' \\
        --csv --log-file gpurun_out/traffic_chol.csv python tools/prof_chol.py
    python tools/ncu_traffic.py cholesky gpurun_out/traffic_chol.csv [halves=6]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    kernel, path = sys.argv[1], sys.argv[2]
    with open(path) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    per_launch = {}
    for row in csv.DictReader(lines):
        if "dram__bytes" not in row["Metric Name"]:
            continue
        v = float(row["Metric Value"].replace(",", "")) * UNIT[row["Metric Unit"]]
        per_launch[row["ID"]] = per_launch.get(row["ID"], 0.0) + v
    vals = [per_launch[k] for k in sorted(per_launch, key=int)]
    # a half = the main launch plus (when the shard has giant rows) its chunk-finish launches; bench.py's
    # `achieved` folds those into the main launch the same way
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 6  # halves run by tools/prof_chol.py
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    out[kernel] = {"bytes_per_launch": sum(vals) / n, "launches": n, "per_kernel_launch": vals,
                   "source": os.path.basename(path), "workload": "C2" if kernel == "cholesky" else "C3"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(out[kernel])


if __name__ == "__main__":
    main()
