"""Turn ncu artefacts in gpurun_out/ into the small text summaries committed under profiles/.
usage: python scripts/summarize_ncu.py <tag> [report names ...]     (reports: gpurun_out/<name>_<tag>.ncu-rep)
Also writes profiles/<tag>_traffic.json (DRAM bytes per algorithmic byte of the fused QDQ + stats kernel) when the
report prof_qdq_stats_<tag> is present: bench.py's roofline.traffic reads it."""
import collections
import csv
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
names = sys.argv[2:] or ["prof_qdq_stats", "prof_gptq_tc", "prof_gptq_simt", "prof_gptq_lowbit"]


def launch_list():
    rows = [r for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list of bench.py's timed region ({tag}): `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`",
           "# (per-launch times are cold-cache and serialised: compare SHARES, not absolutes)", "",
           f"{'launches':>8} {'total_us':>12} {'share':>7}  kernel"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{n:8d} {t / 1e3:12.1f} {100 * t / tot:6.1f}%  {k}")
    out.append(f"{sum(v[0] for v in agg.values()):8d} {tot / 1e3:12.1f}  100.0%  TOTAL")
    return "\n".join(out)


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
           "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_alu.sum",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def raw_page(name):
    """Raw-metric page of a report as CSV text: from gpurun_out/<name>_<tag>.ncu-rep, or -- when the report itself was
    too large to bring back from the GPU box -- from the gpurun_out/<name>_<tag>.csv exported there with the same command."""
    import os

    rep = f"gpurun_out/{name}_{tag}.ncu-rep"
    if os.path.exists(rep):
        return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return open(f"gpurun_out/{name}_{tag}.csv").read()


def report(name):
    rows = list(csv.reader(raw_page(name).splitlines()))
    while rows and "Kernel Name" not in rows[0]:  # ncu banner lines in front of the header of an exported file
        rows.pop(0)
    if len(rows) < 3:
        return f"# {name}: no data"
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --clock-control none, report gpurun_out/{name}_{tag} (selected raw metrics per captured launch, in launch order)"]
    tens = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_active" in h and ".avg" in h]
    for r in rows[2:]:
        out.append("")
        out.append("kernel: " + r[hdr.index("Kernel Name")][:150])
        for m in METRICS + [t for t in tens if t not in METRICS]:
            if m in hdr:
                out.append(f"  {m:88s} {r[hdr.index(m)]:>18s} {units[hdr.index(m)]}")
    return "\n".join(out)


try:
    open(f"profiles/{tag}_bench_launch_list.txt", "w").write(launch_list() + "\n")
except Exception as e:
    print("launch list unchanged:", e)
for name in names:
    try:
        open(f"profiles/{tag}_{name}.txt", "w").write(report(name) + "\n")
        print("wrote", f"profiles/{tag}_{name}.txt")
    except Exception as e:
        print("skip", name, e)


def traffic():
    """dram__bytes_read + dram__bytes_write of the captured stream_kernel<TENSOR, STATS> launches over their
    algorithmic bytes (8 B/elem; grid-stride kernel: elements = the bench's first activation sites)."""
    rows = list(csv.reader(raw_page("prof_qdq_stats").splitlines()))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = sum(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]] for r in rows[2:])
    sites = [256 * 3 * 224 * 224, 256 * 64 * 56 * 56, 256 * 64 * 56 * 56][: len(rows) - 2]  # bench.py site order
    ratio = total / (8.0 * sum(sites))
    json.dump({"qdq_stats_pertensor": {"dram_bytes_per_algorithmic_byte": ratio, "launches": len(rows) - 2,
                                       "note": f"dram__bytes_read.sum + dram__bytes_write.sum over {len(rows) - 2} captured launches "
                                               f"(ncu --set full, {tag}) / 8 B per element"}},
              open(f"profiles/{tag}_traffic.json", "w"), indent=1)
    print("wrote", f"profiles/{tag}_traffic.json", ratio)


try:
    traffic()
except Exception as e:
    print("traffic summary skipped:", e)


def sweep_table(name="prof_all_kernels", log="gpurun_out/ncu_all.log"):
    """One line per captured launch of the all-kernel sweep (scripts/exp/run_all_kernels.py), labelled with the operation
    that launched it (the '## ...' markers of the run's log): duration, DRAM bytes and rate, issue-slot use, the two
    largest stall reasons.  -> profiles/<tag>_prof_all_kernels_table.txt"""
    import re

    rows = list(csv.reader(raw_page(name).splitlines()))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    labels, cur = [], ""
    for line in open(log, errors="replace"):
        if line.startswith("## "):
            cur = line[3:].strip()
        elif line.startswith("==PROF== Profiling"):
            labels.append(cur)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3,
             "byte/second": 1.0, "Kbyte/second": 1e3, "Mbyte/second": 1e6, "Gbyte/second": 1e9, "Tbyte/second": 1e12}

    def val(r, m, default=0.0):
        if m not in hdr:
            return default
        i = hdr.index(m)
        try:
            return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        except ValueError:
            return default

    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = [f"# all-kernel ncu sweep ({tag}): `ncu --metrics <scripts/ncu_metrics.txt> --clock-control none -k regex:sb200` over scripts/exp/run_all_kernels.py",
           "# one line per captured launch, in launch order; times are single cold launches under the profiler (compare rates, not absolutes);",
           "# DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration; pct = gpu__dram_throughput pct of peak; the full metric",
           f"# list per launch is in {tag}_prof_all_kernels.txt", "",
           f"{'operation':44s} {'kernel':40s} {'grid':>7s} {'regs':>4s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>7s} {'dram%':>6s} {'issue%':>6s} {'tensor%':>7s}  top stalls"]
    for i, r in enumerate(rows[2:]):
        kn = re.sub(r"^void |sb200::", "", r[hdr.index("Kernel Name")])
        kn = re.sub(r"\((int|bool)\)", "", kn.split("(const")[0].split("(unsigned")[0].split("(sb200")[0].split("(DecBatch")[0])[:40]
        us = val(r, "gpu__time_duration.sum")
        mb = (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / 1e6
        st = sorted(((val(r, s), s.split("stalled_")[1].split("_per_issue")[0]) for s in stalls), reverse=True)[:2]
        out.append(f"{(labels[i] if i < len(labels) else '')[:44]:44s} {kn:40s} {int(val(r, 'launch__grid_size')):7d} {int(val(r, 'launch__registers_per_thread')):4d} "
                   f"{us:9.1f} {mb:9.2f} {mb / us * 1e3 if us else 0:7.0f} {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
                   f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f}  "
                   + ", ".join(f"{n} {v:.1f}" for v, n in st))
    open(f"profiles/{tag}_prof_all_kernels_table.txt", "w").write("\n".join(out) + "\n")
    print("wrote", f"profiles/{tag}_prof_all_kernels_table.txt", len(rows) - 2, "launches")


if "prof_all_kernels" in names:
    try:
        sweep_table()
    except Exception as e:
        print("sweep table skipped:", e)
