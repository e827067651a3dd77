"""Turns the ncu outputs of scripts/gpu_round_capture.sh (gpurun_out/) into the tracked summaries under profiles/.
Usage: python scripts/summarize_profiles.py <tag>   (tag e.g. r01_final)"""
import csv, collections, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, f"{tag}_ncu_launches.csv"))
bench = json.load(open(os.path.join(G, "bench_final.json")))

# ---- launch list of the timed region
rows = list(csv.reader(open(os.path.join(G, "launches_final.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    k = r[ik].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_ncu_launches.md"), "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 2 --no-cpu-baseline` inside the timed region ({tag})\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -s <launches before the timed region> -c 1200`; the same "
            "command had exited 0 without ncu immediately before (scripts/gpu_round_capture.sh).  Per-launch times are cold-cache "
            "and serialised: compare SHARES.  Kernels inside the CUDA graph of the preconditioner are listed as graph nodes.\n\n"
            "| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f} % |\n")
    fam = sum(t for k, (n, t) in agg.items() if k.startswith("spmv"))
    asm = sum(t for k, (n, t) in agg.items() if k.startswith(("rows_kernel", "facet_kernel", "gate_kernel")))
    gs = sum(t for k, (n, t) in agg.items() if k.startswith(("multi_dot", "multi_axpy", "reduce_rows", "axpby", "update_x")))
    f.write(f"\nTotal {tot:.0f} us over {sum(n for n, _ in agg.values())} launches.  SpMV family (A, AMG levels, transfers, mass matrix) "
            f"{100 * fam / tot:.1f} %; Gram-Schmidt {100 * gs / tot:.1f} %; assembly {100 * asm / tot:.1f} %.\n"
            f"bench.py's own share estimate for the A SpMV: {bench['roofline']['share_of_step']:.3f} of the step.\n")

# ---- full capture
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
def col(d, k):
    return d.get(k, "")
traffic = {}
with open(os.path.join(P, f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none --import-source on` on scripts/profile_probe.py 2048 ({tag})\n\n"
            "One gate step, one assembly, one SpMV on A, one preconditioner application on the bench workload (C3).\n\n"
            "| kernel | time | DRAM read | DRAM write | DRAM GB/s | warps active % | issue active % | warp insts | grid | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        def val(k):
            v, u = float(d[k]), units[hdr.index(k)]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
            return v * scale.get(u, 1.0)
        t, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        f.write(f"| {name} | {t * 1e3:.3f} ms | {rd / 1e9:.3f} GB | {wr / 1e9:.3f} GB | {(rd + wr) / t / 1e9:.0f} | "
                f"{float(d['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} | "
                f"{float(col(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active') or 0):.1f} | {float(d['smsp__inst_executed.sum']):.3g} | "
                f"{d['launch__grid_size']} | {d['launch__registers_per_thread']} |\n")
        traffic.setdefault(name, rd + wr)
    f.write("\nFirst `spmv_stream_kernel<0, 2>` row = y = A x on the system matrix; the following ones are the levels of the two AMG "
            "hierarchies in cycle order.\n")
tj = {f"spmv_stream_kernel<0>@N=2048": traffic.get("spmv_stream_kernel<0, 2>"), "rows_kernel<2,0>@N=2048": traffic.get("rows_kernel<2, 0>")}
json.dump(tj, open(os.path.join(P, "traffic.json"), "w"))
print(open(os.path.join(P, f"{tag}_ncu_launches.md")).read())
print(open(os.path.join(P, f"{tag}_ncu_full.md")).read())
print(tj)
