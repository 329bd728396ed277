"""Copy the round-2 bench lines from gpurun_out/ to profiles/ and print the DESIGN.md tables."""
import json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
def load(name):
    p = os.path.join(GO, name + ".json")
    if not os.path.exists(p): return None
    d = json.loads(open(p).read().strip().splitlines()[-1])
    shutil.copy(p, os.path.join(PR, name + ".json"))
    return d
for n in (2, 4, 8):
    p = os.path.join(GO, "r2_multicheck_n%d.log" % n)
    if os.path.exists(p):
        open(os.path.join(PR, "r2_multicheck_n%d.log" % n), "w").write("".join(l for l in open(p) if l.startswith("{")))
peak = 6451.2
print("single GPU")
for name in ("r2_bench_n1", "r2_bench_n1_f32", "r2_bench_n1_f64_3c", "r2_bench_n1_f32_3c", "r2_long_n1", "r2_bench_reference_arm"):
    d = load(name)
    if not d: continue
    if d.get("impl") == "reference":
        print(name, d["value"], d["cpu_baseline"]["kind"], d["host"], {k: {s: round(v["value"], 5) for s, v in r.items() if isinstance(v, dict) and "value" in v} for k, r in d["extra"]["runs"].items()},
              d["extra"].get("numpy_port_6_threads", {}).get("value"), d["extra"].get("c_openmp_port", {}).get("value"))
        continue
    e = d.get("e2e", {})
    print("%-20s %s value %.2f ms %.4f step_frac %.4f kernel_ms %.4f frac %.4f e2e %.2f (%.3f of value) init %.2fs disk %s clocks %s %s launches/step %.1f" % (
        name, d["dtype"], d["value"], d["ms_per_step"], d["roofline"]["step_frac"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], e.get("value", 0),
        e.get("value", 0) / d["value"], e.get("init_s", 0), (e.get("disk") or {}).get("value"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["gpu_launches"] / d["steps"]))
    if "extra" in d: print("   extra:", {k: (round(v["gcells_per_s"], 2), round(v["ms_per_step"], 3)) for k, v in d["extra"].items()})
    if "cpu_baseline" in d: print("   cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
for kind in ("weak", "strong", "long"):
    print(kind)
    base = None
    for n in (1, 2, 4, 8):
        d = load("r2_%s_n%d" % (kind, n)) if not (kind == "weak" and n == 1) else load("r2_bench_n1")
        if not d: continue
        if base is None: base = d["value"] / (1 if kind != "strong" else 1)
        eff = d["value"] / (base * n) if kind != "strong" else d["value"] / (base * n)
        e = d.get("e2e", {})
        print("  N=%d value %.2f ms %.4f eff %.3f e2e %.2f parity %s clocks %s %s power %.0f" % (n, d["value"], d["ms_per_step"], eff, e.get("value", 0),
              (d.get("parity") or {}).get("slabs_bit_identical"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["clocks"].get("power_w_max") or 0))
