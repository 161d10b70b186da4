"""Copies the judged artefacts of one GPU profiling run (tools/gpu_profile.sh TAG) from gpurun_out/
into profiles/: bench lines, the ncu launch list with a per-kernel share table, selected metrics of
the full captures, and profiles/traffic.json (DRAM bytes per uncompressed byte, read by bench.py)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for f in (f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_launches.csv",
          f"{tag}_bench_4mz.json", f"{tag}_bench_4mz_reference.json", f"{tag}_launches_4mz.csv"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))

# ---- launch lists -> per-kernel shares
for suffix, cmd in (("", "python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"),
                    ("_4mz", "python bench.py --codec 4mz --steps 1 --warmup 1 --no-e2e --no-cpu")):
    lp = os.path.join(G, f"{tag}_launches{suffix}.csv")
    if not os.path.exists(lp):
        continue
    rows = list(csv.reader(open(lp)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows:
        if len(r) != len(hdr) or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1
    with open(os.path.join(P, f"{tag}_launches{suffix}_summary.md"), "w") as f:
        f.write(f"# ncu launch list summary ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none {cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |\n")

# ---- full captures -> selected metrics + traffic
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
# uncompressed bytes one captured launch covers: the whole 16 GiB batch, except the zstd entropy
# stage, which runs once per group of 512 blocks (capi.cu zgroup_blocks)
LAUNCH_BYTES = {"zstd_entropy_kernel": 512 * 4 * 2 ** 20}
traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep in sorted(os.listdir(G)):
    if not (rep.startswith(tag + "_") and rep.endswith(".ncu-rep")):
        continue
    kernel = rep[len(tag) + 1:-len(".ncu-rep")]
    out = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    with open(os.path.join(P, f"{tag}_{kernel}_metrics.txt"), "w") as f:
        f.write(f"# {kernel}: ncu --set full --clock-control none (one launch of `bench.py --steps 1 --total-gib 16`, "
                f"{LAUNCH_BYTES.get(kernel, 16 * 2 ** 30) / 2 ** 30:g} GiB per launch)\n")
        for w in WANT:
            if w in got:
                f.write(f"{w:80s} {got[w][0]} {got[w][1]}\n")
    try:
        def num(k):
            v, u = got[k]
            return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        traffic[kernel] = {"dram_bytes_per_uncompressed_byte": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / LAUNCH_BYTES.get(kernel, 16 * 2 ** 30),
                           "source": f"profiles/{tag}_{kernel}_metrics.txt"}
    except (KeyError, ValueError):
        pass
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("profiles/ updated for", tag)
