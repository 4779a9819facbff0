"""Per-kernel-family DRAM traffic of one train step from an ncu metrics CSV.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none --nvtx --nvtx-include "profiled_step/" --csv --log-file gpurun_out/traffic.csv python tools/one_step.py 3
    python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r02_ncu_traffic.json

bench.py reads the JSON (it never runs under a profiler itself) and reports `roofline.traffic` = measured DRAM bytes per launch
of the dominant family (full-width tcgen05 conv + wgrad kernels) next to the algorithmic figure.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3,
        "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "%": 1.0}


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)          # launch id -> {metric: value, "name": ...}
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", "")) * UNIT.get(row["Metric Unit"], 1.0)
        except (ValueError, KeyError):
            continue
        d = per[int(row["ID"])]
        d[row["Metric Name"]] = v
        d["name"] = re.sub(r"\(.*", "", row["Kernel Name"])
        d["grid"] = row.get("Grid Size", "")
    fam = collections.defaultdict(lambda: dict(launches=0, dram_read=0.0, dram_write=0.0, seconds=0.0, tensor_pct_time=0.0))
    for d in per.values():
        n = d["name"]
        key = ("thin" if "pixgemm" in n or "persistent" in n else
               "conv_tc" if ("tapgemm_tc_kernel" in n or "tapwgrad_tc_kernel" in n) else
               "batchnorm" if "bn_" in n else
               "adam" if "adam" in n else "other")
        a = fam[key]
        t = d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
        a["seconds"] += t
        a["tensor_pct_time"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * t
    res = {"source": path, "families": {}}
    for k, a in fam.items():
        tot = a["dram_read"] + a["dram_write"]
        res["families"][k] = {"launches": a["launches"], "dram_bytes_read": a["dram_read"], "dram_bytes_write": a["dram_write"],
                              "dram_bytes_per_launch": tot / max(a["launches"], 1), "kernel_seconds_cold_serialised": a["seconds"],
                              "dram_GBps": tot / a["seconds"] / 1e9 if a["seconds"] > 0 else 0.0,
                              "tensor_pipe_active_pct_time_weighted": a["tensor_pct_time"] / a["seconds"] if a["seconds"] > 0 else 0.0}
    c = res["families"].get("conv_tc")
    if c:
        res["dominant_family_bytes_per_launch"] = c["dram_bytes_per_launch"]
        res["note"] = (f"mean over the {c['launches']} tapgemm_tc / tapwgrad_tc launches of one train step (B=16, 256x256), ncu "
                       f"dram__bytes_read.sum + dram__bytes_write.sum = {(c['dram_bytes_read'] + c['dram_bytes_write']) / 1e9:.3f} GB per step "
                       f"(cold-cache, serialised replays); tensor pipe active, time-weighted: {c['tensor_pipe_active_pct_time_weighted']:.1f} %")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
