"""profiles/traffic.json from an ncu metrics CSV of the default bench command.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"simTensorKernel|etkdgKernel|bfgsKernel|verifyCandidates" --csv --log-file gpurun_out/traffic.csv \
        python bench.py --steps 1 --warmup 3 --cross-n 0
    python tools/traffic_from_ncu.py gpurun_out/traffic.csv "<the command, for the record>"

For each kernel the LARGEST launch is kept (the timed step's; warm-up launches of the conformer legs are smaller pools):
bytes = dram__bytes_read.sum + dram__bytes_write.sum of that launch. bench.py reads the file for the roofline "traffic" key.
"""
import collections
import csv
import json
import os
import sys

KEYS = {"simTensorKernel<0": "simTensorKernel<count>", "etkdgKernel": "etkdgKernel", "bfgsKernel<ff::Mmff>": "bfgsKernel<Mmff>",
        "verifyCandidatesKernel": "verifyCandidatesKernel"}


def main(path: str, source: str) -> None:
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    launches = collections.defaultdict(dict)  # id -> {name, metric: value}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        rec = launches[d["ID"]]
        rec["name"] = d["Kernel Name"]
        rec[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    out = {}
    for rec in launches.values():
        key = next((v for k, v in KEYS.items() if k in rec["name"]), None)
        if key is None or "dram__bytes_read.sum" not in rec:
            continue
        total = rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
        if key not in out or total > out[key]["bytes"]:
            out[key] = {"bytes": total, "read": rec["dram__bytes_read.sum"], "write": rec["dram__bytes_write.sum"],
                        "ncu_duration_ms": rec.get("gpu__time_duration.sum", 0.0) / 1e6, "source": source}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]))
