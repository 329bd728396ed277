"""Round 2: turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (a step = TWO stencil launches:
the face-owning z-tile and the other z-tiles; traffic and time are reported per launch and summed per step)."""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
traffic = json.load(open(os.path.join(PR, "traffic.json"))) if os.path.exists(os.path.join(PR, "traffic.json")) else {}
lines = ["# %s: ncu --set full --clock-control none of the stencil launches of ONE step (k_step_march), 512^3 crystal" % tag, "",
         "A step is two launches side by side (DESIGN.md section 4, split step): the z-tile that owns the z = -1 face with the `ZF`",
         "instantiation (last template argument 1) and all other z-tiles with the plain one (0).  ncu serialises them; in the run",
         "they share the same waves of blocks -- so the DRAM traffic below is an UPPER bound for the real step: serialised, the",
         "face tile's launch re-reads from DRAM the y / z halo data it shares through L2 with its neighbour tiles when they run",
         "together (round 1, one launch for all tiles: traffic / algorithmic = 1.005).", ""]
for d in ("f64", "f32"):
    rep = os.path.join(GO, "%s_march_%s.ncu-rep" % (tag, d))
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    tot_rd = tot_wr = tot_t = 0.0
    for r in rows[2:]:
        v, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        name = re.sub(r"phb::", "", v["Kernel Name"])[:110]
        lines += ["## %s  `%s`" % (d, name), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in v:
                lines.append("| %s | %s | %s |" % (k, v[k], u[k]))
        rd = float(v["dram__bytes_read.sum"].replace(",", "")) * UNIT[u["dram__bytes_read.sum"]]
        wr = float(v["dram__bytes_write.sum"].replace(",", "")) * UNIT[u["dram__bytes_write.sum"]]
        t = float(v["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u["gpu__time_duration.sum"], 1e-3)
        tot_rd += rd; tot_wr += wr; tot_t += t
        lines.append("")
    cells = 512 ** 3
    alg = cells * (9 * (8 if d == "f64" else 4) + 1)
    lines += ["**%s, per step (both launches):** DRAM %.3f GB read + %.3f GB written = %.3f GB; algorithmic %.3f GB -> traffic / algorithmic = %.3f "
              "(%.1f B per cell-update vs %d algorithmic); serialised kernel time %.4f ms" % (
                  d, tot_rd / 1e9, tot_wr / 1e9, (tot_rd + tot_wr) / 1e9, alg / 1e9, (tot_rd + tot_wr) / alg, (tot_rd + tot_wr) / cells, alg // cells, tot_t * 1e3), ""]
    traffic["%s_512" % d] = {"dram_bytes_per_launch": tot_rd + tot_wr, "read": tot_rd, "write": tot_wr, "algorithmic": alg,
                             "source": "%s_march_%s.ncu-rep (sum of the two stencil launches of one step)" % (tag, d)}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for r in csv.reader(io.StringIO(src)):
        if len(r) > 6 and r[5].isdigit() and not r[0].startswith(("Kernel", "Address")):
            op = r[1].strip().split()
            op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
            ops[op.split(".")[0]] += int(r[5])
    tot = sum(ops.values()) or 1
    lines += ["Executed SASS mix of the step (warp instructions, top 14 of %d): " % tot + ", ".join("%s %.1f%%" % (k, 100.0 * n / tot) for k, n in ops.most_common(14)),
              "TMA evidence: UTMALDG executed %d times, SYNCS (mbarrier) %d, LDG %d, STG %d" % (ops.get("UTMALDG", 0), ops.get("SYNCS", 0), ops.get("LDG", 0), ops.get("STG", 0)), ""]
open(os.path.join(PR, "%s_march_ncu_summary.md" % tag), "w").write("\n".join(lines))
json.dump(traffic, open(os.path.join(PR, "traffic.json"), "w"), indent=1)
ll = os.path.join(GO, "%s_launches.csv" % tag)
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"phb::", "", r[ik]).split("(")[0].replace("void ", "")[:80]
        agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[iv].replace(",", "")) / 1e3
    step = {k: v for k, v in agg.items() if any(s in k for s in ("k_step", "k_abc", "k_source", "k_faces"))}
    tot = sum(v[1] for v in step.values())
    out = ["# %s: ncu launch list of `python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu` (gpu__time_duration.sum, --clock-control none)" % tag, "",
           "Per-launch times under ncu are cold-cache and serialised (the two stencil launches of a step run side by side in the real",
           "run): compare SHARES, not absolutes.  Stencil share = both k_step_march rows.", "",
           "| kernel | launches | total us | mean us | share of step kernels |", "|---|---|---|---|---|"]
    for k, (n, us) in agg.items():
        share = "%.2f%%" % (100 * us / tot) if k in step else "(init)"
        out.append("| %s | %d | %.1f | %.1f | %s |" % (k, n, us, us / n, share))
    ks = sum(v[1] for k, v in step.items() if "k_step" in k)
    out += ["", "stencil launches: %.2f%% of the step kernels' time" % (100 * ks / tot)]
    open(os.path.join(PR, "%s_launch_list.md" % tag), "w").write("\n".join(out) + "\n")
    import shutil; shutil.copy(ll, os.path.join(PR, "%s_launches.csv" % tag))
print(open(os.path.join(PR, "%s_march_ncu_summary.md" % tag)).read()[-2500:])
