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
