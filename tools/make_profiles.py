"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/."""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(PR, exist_ok=True)

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_dynamic", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__cycles_elapsed.avg.per_second", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
traffic = {}
tj = os.path.join(PR, "traffic.json")
if os.path.exists(tj):
    traffic = json.load(open(tj))
lines = ["# %s: ncu --set full of the dominant kernel (k_step_march), 512^3 crystal, one launch = one time step" % tag, ""]
for d in ("f64", "f32"):
    rep = os.path.join(GO, "%s_march_%s.ncu-rep" % (tag, d))
    if not os.path.exists(rep):
        continue
    v, u = raw(rep)
    lines += ["## %s  (%s)" % (d, v.get("Kernel Name", "")[:100]), "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in v:
            lines.append("| %s | %s | %s |" % (k, v[k], u[k]))
    rd = float(v["dram__bytes_read.sum"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1, "Kbyte": 1e3}[u["dram__bytes_read.sum"]]
    wr = float(v["dram__bytes_write.sum"].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "byte": 1, "Kbyte": 1e3}[u["dram__bytes_write.sum"]]
    cells = 512 ** 3
    alg = cells * (9 * (8 if d == "f64" else 4) + 1)
    lines += ["", "DRAM traffic per launch: %.3f GB read + %.3f GB written = %.3f GB; algorithmic bytes %.3f GB -> traffic/algorithmic = %.3f "
              "(%.1f B per cell-update vs %d algorithmic)" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, alg / 1e9, (rd + wr) / alg, (rd + wr) / cells, alg // cells), ""]
    traffic["%s_512" % d] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "algorithmic": alg, "source": "%s_march_%s.ncu-rep" % (tag, d)}
    # SASS evidence
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) > iex and r[iex].isdigit():
            op = r[isrc].strip().split()
            op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
            ops[op.split(".")[0]] += int(r[iex])
    tot = sum(ops.values())
    lines += ["Executed SASS mix (warp instructions, top 14 of %d): " % tot + ", ".join("%s %.1f%%" % (k, 100.0 * n / tot) for k, n in ops.most_common(14)),
              "TMA evidence: UTMALDG executed %d times, SYNCS (mbarrier) %d, LDG %d, STG %d" % (ops.get("UTMALDG", 0), ops.get("SYNCS", 0), ops.get("LDG", 0), ops.get("STG", 0)), ""]
open(os.path.join(PR, "%s_march_ncu_summary.md" % tag), "w").write("\n".join(lines))
json.dump(traffic, open(tj, "w"), indent=1)
# launch list -> per-kernel shares
ll = os.path.join(GO, "%s_launches.csv" % tag)
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[14]) / 1e3
    step = [r for r in rows if "k_step" in r[4] or "k_abc" in r[4] or "k_source" in r[4]]
    tot = sum(float(r[14]) for r in step) / 1e3
    out = ["# %s: ncu launch list of `python bench.py --steps 5 --warmup 3` (gpu__time_duration.sum, --clock-control none)" % tag, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total us | share of step kernels |", "|---|---|---|---|"]
    for k, (n, us) in agg.items():
        share = "%.2f%%" % (100 * us / tot) if ("k_step" in k or "k_abc" in k or "k_source" in k) else "(init)"
        out.append("| %s | %d | %.1f | %s |" % (k[:70], n, us, share))
    open(os.path.join(PR, "%s_launch_list.md" % tag), "w").write("\n".join(out) + "\n")
    import shutil; shutil.copy(ll, os.path.join(PR, "%s_launches.csv" % tag))
print(open(os.path.join(PR, "%s_march_ncu_summary.md" % tag)).read()[:3000])
