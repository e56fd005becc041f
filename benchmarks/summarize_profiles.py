#!/usr/bin/env python
"""Turn the raw ncu outputs of benchmarks/r2_profiles.sh (gpurun_out/) into the small text / json summaries kept under profiles/."""
import csv, json, os, sys, collections

G, P = "gpurun_out", "profiles"


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
    return dict(zip(hdr, vals))


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def launches(path, out_csv, out_txt, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.OrderedDict()
    n = collections.Counter()
    with open(out_csv, "w") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "gpu__time_duration_ns"])
        i = 0
        for r in rows[1:]:
            if r[mi] != "gpu__time_duration.sum":
                continue
            name = r[ki].split("(")[0][:90]
            v = num(r[vi])
            w.writerow([i, name, v])
            tot[name] = tot.get(name, 0.0) + v
            n[name] += 1
            i += 1
    s = sum(tot.values())
    with open(out_txt, "w") as f:
        f.write(title + "\n")
        f.write(f"{i} launches profiled, {s / 1e6:.1f} ms of kernel time (per-launch times are cold-cache and serialised: read the SHARES)\n\n")
        for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:25]:
            f.write(f"{100 * v / s:6.2f} %  {v / 1e6:10.3f} ms  {n[name]:5d} x  {name}\n")


def main():
    if os.path.exists(f"{G}/r2_bench_launches.csv"):
        launches(f"{G}/r2_bench_launches.csv", f"{P}/r2_bench_launches.csv", f"{P}/r2_bench_launch_shares.txt",
                 "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500, command: python bench.py --steps 2 --warmup 1")
    if os.path.exists(f"{G}/r2_train_launches.csv"):
        launches(f"{G}/r2_train_launches.csv", f"{P}/r2_train_launches.csv", f"{P}/r2_train_launch_shares.txt",
                 "ncu launch list of benchmarks/train_launches.py 1024 8192 (3 eager training steps per batch size, C4 model)")
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
    for tag, what in (("sampler", "sampler_pair_kernel<512>, bench launch configuration (8 bases x 1e6 shots, T = 100): python bench.py --steps 1 --warmup 1 --no-extras"),
                      ("fused", "train_fused_kernel<512>, C4 model, batch 8192 (32 CTA pairs): python benchmarks/train_launches.py 8192")):
        path = f"{G}/r2_{tag}_raw.csv"
        if not os.path.exists(path):
            continue
        m = raw_metrics(path)
        with open(f"{P}/r2_{tag}_ncu_summary.txt", "w") as f:
            f.write(f"ncu --set full --clock-control none --import-source on -- {what}\n\n")
            for k in keys:
                hits = [h for h in m if h == k or h.endswith(k)]
                for h in hits[:1]:
                    f.write(f"{k:100s} {m[h]}\n")
        if tag == "sampler":
            rd = num(m.get("dram__bytes_read.sum")); wr = num(m.get("dram__bytes_write.sum"))
            units = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
            # the raw page prints a unit row (rows[1]); re-read it
            rows = list(csv.reader(open(path)))
            unit = dict(zip(rows[0], rows[1]))
            scale = lambda k: units.get(unit.get(k, "byte"), 1.0)
            total = rd * scale("dram__bytes_read.sum") + wr * scale("dram__bytes_write.sum")
            json.dump({"bases_per_step": 8, "shots": 1000000, "dram_bytes_per_launch": int(total),
                       "dram_bytes_read": int(rd * scale("dram__bytes_read.sum")), "dram_bytes_write": int(wr * scale("dram__bytes_write.sum")),
                       "source": "profiles/r2_sampler_ncu_summary.txt (ncu --set full of sampler_pair_kernel<512> at the bench launch configuration)"},
                      open(f"{P}/r2_sampler_traffic.json", "w"), indent=1)
    for name in ("r2_hbm_kernels.json",):
        if os.path.exists(f"{G}/{name}"):
            open(f"{P}/{name}", "w").write(open(f"{G}/{name}").read())
    for name in ("r2_recon_breakdown.log", "r2_fused_stamps.log"):
        if os.path.exists(f"{G}/{name}"):
            open(f"{P}/{name.replace('.log', '.txt')}", "w").write(open(f"{G}/{name}").read())


if __name__ == "__main__":
    main()
