#!/usr/bin/env python
"""Turn the raw artifacts of profiles/r02_final.sh (gpurun_out/r02/) into the committed round-2 evidence:
    profiles/r02_*.json|csv|txt      copies of the measurement files
    profiles/r02_ncu_<name>.csv      selected metrics of every kernel of each `ncu --set full` capture (from the .ncu-rep, read
                                     here with `ncu -i ... --page raw --csv`; the reports themselves stay in gpurun_out/)
    profiles/r02_tensor_pipe.json    tensor-pipe activity of the Unet_3D conv kernels (read by bench.py's roofline block)
    profiles/r02_sass_counts.txt     tcgen05 / TMA / bulk-copy mnemonics in the shipped library (cuobjdump, no GPU needed)
    profiles/r02_summary.md          the numbers DESIGN.md quotes, with their sources
Run from the repo root after the GPU pass:  python profiles/make_r02_summary.py [gpurun_out/r02]"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = os.path.join(REPO, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02")

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def load_json(name):
    p = os.path.join(SRC, name)
    if not os.path.exists(p):
        return None
    txt = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(txt[-1]) if txt else None


def copy(name, dst=None):
    p = os.path.join(SRC, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(HERE, "r02_" + (dst or name)))
        return True
    return False


def ncu_rows(rep):
    """rows (dicts) of one capture, metric name -> (value, unit); reads the raw CSV made on the GPU box (`ncu -i ... --page raw
    --csv`, next to where the .ncu-rep was) or, if the report itself is here, converts it now"""
    p = os.path.join(SRC, rep)
    raw = p[:-len(".ncu-rep")] + ".rawcsv"
    if os.path.exists(raw):
        out = open(raw).read()
    elif os.path.exists(p):
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    else:
        return []
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    return [{h: (r[i], units[i]) for i, h in enumerate(hdr) if i < len(r)} for r in rows[2:]]


def write_ncu_csv(rep, dst):
    rows = ncu_rows(rep)
    if not rows:
        return []
    with open(os.path.join(HERE, dst), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + METRICS)
        for r in rows:
            w.writerow([r.get("Kernel Name", ("?", ""))[0][:110]] + [" ".join(r[m]).strip() if m in r else "" for m in METRICS])
    return rows


def launch_table(csv_name, top=14):
    p = os.path.join(SRC, csv_name)
    if not os.path.exists(p):
        return "(missing %s)\n" % csv_name, 0.0
    lines = [l for l in open(p) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row.get("Metric Unit", "ns").startswith("n") else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    s = "kernel | launches | total us | share\n---|---|---|---\n"
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        s += "%s | %d | %.1f | %.1f%%\n" % (n, c, t, 100 * t / max(tot, 1e-9))
    return s, tot


def tensor_pipe(csv_name):
    p = os.path.join(SRC, csv_name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (row["ID"], re.sub(r"\(.*", "", row["Kernel Name"])[:90])
        per.setdefault(key, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    rows = []
    for (_, name), m in per.items():
        if not any(k in name for k in ("convt3d", "convflat", "col2im")):
            continue
        t = m.get("gpu__time_duration.sum", 0.0)
        rows.append({"kernel": name, "us": t / 1e3 if t > 1e4 else t, "tensor_pipe_active_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")})
    tot = sum(r["us"] for r in rows)
    w = sum(r["us"] * (r["tensor_pipe_active_pct"] or 0) for r in rows) / max(tot, 1e-9)
    return {"kernels": rows, "conv_kernel_us_total": tot, "time_weighted_tensor_pipe_active_pct": w,
            "max_tensor_pipe_active_pct": max([r["tensor_pipe_active_pct"] or 0 for r in rows] or [0])}


def main():
    md = ["# Round 2 — measured on a B200 through `gpurun` (`profiles/r02_final.sh`); raw files: `profiles/r02_*`\n"]
    for n in ("pytest_gpu.txt", "bench_n1.json", "bench_reference_arm.json", "unet_breakdown_exact.json", "unet_breakdown_f16.json",
              "microbench_ops.json", "microbench_render.json", "microbench_cam_bp.json", "genre_breakdown.json", "train_unet_exact.json",
              "train_unet_f16.json", "train_unet_launches_f16.csv", "launches_genre_step.csv", "unet_tensor_pipe_exact.csv",
              "unet_tensor_pipe_f16.csv", "bench_n2.json", "bench_n8.json", "bench_n4.json", "train_shapehd_launches.csv",
              "train_wgan_launches.csv", "train_genre_launches.csv"):
        copy(n)
    b, r = load_json("bench_n1.json"), load_json("bench_reference_arm.json")
    if b:
        rf = b["roofline"] or {}
        md.append("## bench.py, 1 GPU (`r02_bench_n1.json`)\n")
        md.append("- metric: %s\n- workload: %s\n- conv mode: %s" % (b["metric"], b["config"]["workload"], b["config"].get("conv_mode")))
        md.append("- value: **%.0f shapes/s** (%.2f ms per batch-16 forward, %s); e2e %.0f shapes/s (%.2f ms/step; %s)"
                  % (b["value"], b["ms_per_step"], b["launch_mode"], b["e2e"]["value"], b["e2e"]["ms_per_step"], b["e2e"]["pipeline"]))
        if b.get("cpu_baseline"):
            md.append("- cpu_baseline: %.2f shapes/s (%s)" % (b["cpu_baseline"]["value"], b["cpu_baseline"]["sample"][:160]))
        if r:
            md.append("- reference arm (`r02_bench_reference_arm.json`): %.2f shapes/s, %d cores" % (r["value"], r["cpu_baseline"]["cores"]))
        if rf.get("clauses"):
            md.append("- roofline (measured HBM peak %.0f GB/s): cam_bp whole op %.1f us = **%.3f**;" % (rf["peak"], rf["op_us"], rf["frac"]))
            for k, v in rf["clauses"].items():
                md.append("  - %s: %s = %.3f of its roofline (%s)" % (k, ("%.1f us" % v["us"]) if "us" in v else ("%.2f ms" % v["ms"]), v["frac"], v["bound"]))
        sec = b.get("secondary") or {}
        for k, v in sec.items():
            md.append("- secondary %s: %s" % (k, json.dumps(v)[:260]))
        ddp = b.get("secondary_ddp") or {}
        for k in ("shapehd", "wgangp_critic", "genre_finetune"):
            if k in ddp:
                md.append("- training step %s (1 GPU): %.2f ms, %.0f shapes/s" % (k, ddp[k]["step_ms"], ddp[k]["shapes_per_s"]))
        md.append("- clocks: %s\n" % json.dumps(b.get("clocks")))
    for n in (2, 4, 8):
        bn = load_json("bench_n%d.json" % n)
        if bn and b:
            line = "- %d GPUs: %.0f shapes/s (%.2fx), e2e %.0f (%.2fx)" % (n, bn["value"], bn["value"] / b["value"], bn["e2e"]["value"], bn["e2e"]["value"] / b["e2e"]["value"])
            ddp = bn.get("secondary_ddp") or {}
            for k in ("shapehd", "wgangp_critic", "genre_finetune"):
                if k in ddp and "step_ms" in ddp[k]:
                    line += "; %s %.2f ms (%.0f shapes/s, exposed all-reduce %.2f ms = %.1f%%)" % (
                        k, ddp[k]["step_ms"], ddp[k]["shapes_per_s"], ddp[k].get("exposed_allreduce_ms", 0), 100 * ddp[k].get("exposed_allreduce_frac", 0))
            md.append(line)
    md.append("\n## launch list of the headline step (`r02_launches_genre_step.csv`, eager launches, `ncu --metrics gpu__time_duration.sum`)\n")
    md.append("Cold-cache, serialised replays: compare SHARES, not absolutes.\n")
    t, _ = launch_table("launches_genre_step.csv", 22)
    md.append(t)
    md.append("\n## Unet_3D training step B=4, opt-in f16 mode (`r02_train_unet_launches_f16.csv`)\n")
    t, tot = launch_table("train_unet_launches_f16.csv", 16)
    md.append(t)
    for w, title in (("shapehd", "ShapeHD fine-tune step, B=8"), ("wgan", "WGAN-GP critic step, B=8"), ("genre", "GenRe joint fine-tune + Chamfer step, B=4")):
        md.append("\n## %s, default (exact) conv mode (`r02_train_%s_launches.csv`)\n" % (title, w))
        t, _ = launch_table("train_%s_launches.csv" % w, 12)
        md.append(t)
    tp = {}
    for mode in ("exact", "f16"):
        x = tensor_pipe("unet_tensor_pipe_%s.csv" % mode)
        if x:
            tp[mode] = x
    if tp:
        json.dump(tp, open(os.path.join(HERE, "r02_tensor_pipe.json"), "w"), indent=1)
        md.append("\n## tensor-pipe activity of the conv kernels of one Unet_3D forward, B=16 (`r02_tensor_pipe.json`)\n")
        for mode, x in tp.items():
            md.append("- %s mode: %.0f us of conv kernels, time-weighted `sm__pipe_tensor_cycles_active` **%.1f %%**, max %.1f %%"
                      % (mode, x["conv_kernel_us_total"], x["time_weighted_tensor_pipe_active_pct"], x["max_tensor_pipe_active_pct"]))
            for k in x["kernels"]:
                md.append("  - %.0f us, %.1f %% : %s" % (k["us"], k["tensor_pipe_active_pct"] or 0, k["kernel"][:80]))
    md.append("\n## ncu --set full captures (selected metrics in `r02_ncu_*.csv`)\n")
    for rep, dst in [("prof_unet_convs.ncu-rep", "r02_ncu_unet_convs.csv"), ("prof_render.ncu-rep", "r02_ncu_render.csv")] + [
            ("prof_op_%s.ncu-rep" % k, "r02_ncu_op_%s.csv" % k) for k in ("nnd_forward", "calc_prob_forward", "calc_prob_backward", "sph_project",
                                                                          "vox_splat", "cam_project", "sph_bp_backward", "cam_bp_backward")]:
        rows = write_ncu_csv(rep, dst)
        for rr in rows[:16]:
            g = lambda m: " ".join(rr[m]).strip() if m in rr else "?"
            md.append("- `%s`: %s; dram r/w %s / %s; issue-active %s; warps-active %s; L1 hit %s; tensor pipe %s; regs %s"
                      % (rr.get("Kernel Name", ("?",))[0][:70], g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum"),
                         g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"),
                         g("l1tex__t_sector_hit_rate.pct"), g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         g("launch__registers_per_thread")))
    # dram traffic of the splat kernel per launch (bench.py's roofline.traffic)
    rows = ncu_rows("prof_op_vox_splat.ncu-rep")
    if rows:
        def num(x):
            v, u = x
            return float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        rr = rows[0]
        json.dump({"kernel": rr["Kernel Name"][0][:80], "dram_bytes_per_launch_b16": num(rr["dram__bytes_read.sum"]) + num(rr["dram__bytes_write.sum"]),
                   "note": "ncu --set full, one launch of the batch-16 splat WITH the count volume (microbench_ops.py): writes tdf + cnt; "
                           "the write-back L2 still holds dirty lines when the window closes"},
                  open(os.path.join(HERE, "splat_traffic.json"), "w"))
    lib = os.path.join(REPO, "genre_shapehd_b200", "lib", "libgenre_b200.so")
    if os.path.exists(lib):
        sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        cnt = collections.Counter(m for m in re.findall(r"\b(UTCHMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|UTCBAR|LDGSTS|FFMA2|FADD2|FMUL2|FMNMX3|ATOMS|MATCH)\b", sass))
        with open(os.path.join(HERE, "r02_sass_counts.txt"), "w") as f:
            f.write("cuobjdump -sass genre_shapehd_b200/lib/libgenre_b200.so | grep -c <mnemonic>\n")
            for k, v in sorted(cnt.items()):
                f.write("%-8s %d\n" % (k, v))
        md.append("\n## SASS mnemonics in the shipped library (`r02_sass_counts.txt`)\n")
        md.append(", ".join("%s %d" % kv for kv in sorted(cnt.items())))
    open(os.path.join(HERE, "r02_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md)[:3000])


if __name__ == "__main__":
    main()
