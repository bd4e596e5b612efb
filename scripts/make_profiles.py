#!/usr/bin/env python
"""Turn the ncu outputs of scripts/gpu_profiles.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
alias = {"k_qv_code<1>": "k_qv_emit(file)", "k_qv_code<2>": "k_qv_emit", "k_qv_code<0>": "k_qv_size", "k_qv_hist<0>": "k_qv_hist_plain",
         "k_pred_slots<0>": "k_pred_slots", "k_pred_slots<2>": "k_pred_slots(candidates)"}

def rows_of(path):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

def kname(r):
    n = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
    return alias.get(n, n)

# launch list
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(G, "launches.csv")],
                     capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_launches_2GB.txt"), "w").write(
    "# ncu --metrics gpu__time_duration.sum --clock-control none on: python bench.py --size-gb 2 --steps 2 --warmup 3 --no-extras --no-cpu\n" + out)
# traffic
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows_of(os.path.join(G, "traffic.csv")):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3,
          "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
    agg[kname(r)][r["Metric Name"]].append(v)
bench = json.load(open(os.path.join(G, "bench.json")))
U = bench["config"]["uncompressed_bytes_per_gpu"]
tr = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on "
                "`python bench.py --size-gb 2 --steps 2 --warmup 3 --no-extras --no-cpu` (B200), averages per launch",
      "text_bytes": U, "kernels": {}}
for k, m in sorted(agg.items()):
    rd = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"])
    wr = sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
    t = sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"])
    tr["kernels"][k] = {"launches": len(m["dram__bytes_read.sum"]), "dram_read_bytes": round(rd), "dram_write_bytes": round(wr),
                        "dram_bytes_per_text_byte": round((rd + wr) / U, 4), "avg_us": round(t, 1)}
json.dump(tr, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
# full capture summaries
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_metrics.py"), os.path.join(G, "prof_full.ncu-rep")],
                     capture_output=True, text=True).stdout
stall = []
for k in ("k_qv_decode5", "k_qv_code", "k_qv_hist", "k_pred_slots"):
    for skip in ("0", "1"):
        o = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot.py"), os.path.join(G, "prof_full.ncu-rep"), k, skip, "0"],
                           capture_output=True, text=True).stdout.splitlines()[:3]
        if len(o) == 3 and o not in stall:
            stall.append(o)
open(os.path.join(P, f"{tag}_ncu_full_2GB.txt"), "w").write(
    "# ncu --set full --clock-control none --import-source on, same command; key metrics per captured launch\n" + out +
    "\n# warp-stall sample shares (source page)\n" + "\n".join("\n".join(x) for x in stall) + "\n")
if os.path.exists(os.path.join(G, "prof_pack.ncu-rep")):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_metrics.py"), os.path.join(G, "prof_pack.ncu-rep")],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_ncu_full_pack_1GB.txt"), "w").write(
        "# ncu --set full --clock-control none on: python scripts/ncu_once.py 0.05 1.0 (1 GB .fasta: dexta then undexta)\n" + out)
print(open(os.path.join(P, f"{tag}_launches_2GB.txt")).read()[:1500])
print({k: (v["dram_bytes_per_text_byte"], v["avg_us"]) for k, v in tr["kernels"].items()})
