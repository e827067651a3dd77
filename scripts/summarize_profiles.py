"""Turns the text outputs of scripts/gpu_round_capture.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
Usage: python scripts/summarize_profiles.py <tag>   (tag e.g. r02)"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
g = lambda name: os.path.join(G, f"{tag}_{name}")
shutil.copy(g("bench.json"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(g("launches.csv"), os.path.join(P, f"{tag}_ncu_launches.csv"))
bench = json.loads(open(g("bench.json")).read().strip().splitlines()[-1])
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9, "usecond": 1e-6,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def short(name):
    name = name.replace("void ", "").replace("knp::", "")
    m = re.match(r"([A-Za-z_0-9]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def metric_rows(path):
    """rows of an `ncu --csv --metrics ...` log: (id, kernel, metric, unit, value)"""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    for r in rows[hi + 1:]:
        if len(r) > iv and r[0].isdigit():
            yield int(r[0]), short(r[ik]), r[im], r[iu], float(r[iv].replace(",", ""))


# ---- launch list of the timed region
agg = collections.OrderedDict()
for _, k, m, u, v in metric_rows(g("launches.csv")):
    if m != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v * SCALE.get(u, 1e-9) * 1e6
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_ncu_launches.md"), "w") as f:
    f.write(f"# ncu launch list of `python bench.py --steps 2 --no-cpu-baseline --skip-c4 --skip-parity` inside the timed region ({tag})\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -s <launches before the timed region> -c 1000`; the same "
            "command had exited 0 without ncu immediately before (scripts/gpu_round_capture.sh).  Per-launch times are cold-cache "
            "and serialised: compare SHARES.  Kernels inside the CUDA graph of the preconditioner are listed as graph nodes.\n\n"
            "| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f} % |\n")
    fam = sum(t for k, (n, t) in agg.items() if k.startswith(("spmv", "dense_gemv", "scale_dinv", "schur", "amg_tail")))
    spa = sum(t for k, (n, t) in agg.items() if k.startswith("spmv_stream_kernel<0, 2"))
    asm = sum(t for k, (n, t) in agg.items() if k.startswith(("rows_", "facet_kernel", "gate_kernel")))
    gs = sum(t for k, (n, t) in agg.items() if k.startswith(("multi_dot", "multi_axpy", "reduce_rows", "axpby", "update_x", "range_")))
    f.write(f"\nTotal {tot:.0f} us over {sum(n for n, _ in agg.values())} launches.  SpMV family (A, AMG levels, transfers, mass matrix, "
            f"coarse solves, Schur glue) {100 * fam / tot:.1f} %; Gram-Schmidt and nullspace projection {100 * gs / tot:.1f} %; "
            f"assembly {100 * asm / tot:.1f} %.\nbench.py's own share estimates of the same step: "
            + "; ".join(f"{k.split(' (')[0]} {v['share_of_step']:.3f}" for k, v in bench["kernels"].items()) + ".\n")

# ---- full capture (raw page of the first kernels of one probe repetition)
rows = list(csv.reader(open(g("full_raw.csv"))))
hdr, units = rows[0], rows[1]
with open(os.path.join(P, f"{tag}_ncu_full.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none --import-source on` on scripts/profile_probe.py c3 2048 ({tag})\n\n"
            "One gate step, one assembly (facet + row kernel), one SpMV on A and the first kernels of one preconditioner application "
            "on the bench workload (C3); the .ncu-rep stays on the GPU box, this table is its raw page.\n\n"
            "| kernel | time | DRAM read | DRAM write | DRAM GB/s | warps active % | issue active % | warp insts | grid | regs |\n"
            "|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))

        def val(k):
            return float(d[k]) * SCALE.get(units[hdr.index(k)], 1.0)
        t, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        f.write(f"| {short(d['Kernel Name'])} | {t * 1e3:.3f} ms | {rd / 1e9:.3f} GB | {wr / 1e9:.3f} GB | {(rd + wr) / t / 1e9:.0f} | "
                f"{float(d['sm__warps_active.avg.pct_of_peak_sustained_active']):.1f} | "
                f"{float(d.get('smsp__issue_active.avg.pct_of_peak_sustained_active') or 0):.1f} | "
                f"{float(d['smsp__inst_executed.sum']):.3g} | {d['launch__grid_size']} | {d['launch__registers_per_thread']} |\n")

# ---- DRAM traffic of one assembly / SpMV(A) / preconditioner application (per launch group)
per = collections.OrderedDict()
for i, k, m, u, v in metric_rows(g("traffic.csv")):
    per.setdefault(i, {"kernel": k})[m] = v * SCALE.get(u, 1.0)
order = [per[i] for i in sorted(per)]
wl, n = "c3", 2048
asm_b = sum(o.get("dram__bytes_read.sum", 0) + o.get("dram__bytes_write.sum", 0) for o in order
            if o["kernel"].startswith(("facet_kernel", "rows_", "gate_kernel")))
ia = next(i for i, o in enumerate(order) if o["kernel"].startswith("spmv_stream_kernel"))      # first SpMV = y = A x
spmv_b = order[ia].get("dram__bytes_read.sum", 0) + order[ia].get("dram__bytes_write.sum", 0)
pc = order[ia + 1:]
pc_b = sum(o.get("dram__bytes_read.sum", 0) + o.get("dram__bytes_write.sum", 0) for o in pc)
pc_t = sum(o.get("gpu__time_duration.sum", 0) for o in pc)
traffic = {f"assembly@{wl}:N={n}": asm_b, f"spmv_stream_kernel<0>@{wl}:N={n}": spmv_b, f"pc_apply@{wl}:N={n}": pc_b}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
with open(os.path.join(P, f"{tag}_ncu_traffic.md"), "w") as f:
    f.write(f"# DRAM traffic per launch group ({tag}): `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` "
            "on one probe repetition (scripts/profile_probe.py c3 2048, plain launches)\n\n"
            "| group | launches | DRAM bytes (read + write) | algorithmic bytes (bench.py) | ratio |\n|---|---|---|---|---|\n")
    kk = bench["kernels"]
    names = list(kk)
    for label, nl, b, key in (("assembly (gate + facet + rows)", 3, asm_b, names[0]), ("y = A x", 1, spmv_b, names[1]),
                              ("preconditioner application", len(pc), pc_b, names[2])):
        alg = kk[key]["algorithmic_bytes"]
        f.write(f"| {label} | {nl} | {b / 1e9:.3f} GB | {alg / 1e9:.3f} GB | {b / alg:.2f} |\n")
    f.write(f"\nSerialised, cold-cache time of the {len(pc)} launches of the preconditioner application: {pc_t * 1e3:.3f} ms "
            f"(graph replay in the bench: {kk[names[2]]['ms']:.3f} ms).\n")
print(open(os.path.join(P, f"{tag}_ncu_launches.md")).read())
print(open(os.path.join(P, f"{tag}_ncu_full.md")).read())
print(open(os.path.join(P, f"{tag}_ncu_traffic.md")).read())
