"""Prints the headline fields of bench.py JSON lines (files given on the command line)."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][0])
    except Exception as e:
        print(path, "unreadable:", e); continue
    s = d.get("sampler") or {}
    e = d.get("e2e") or {}
    p = d.get("parity") or {}
    print("%s: N=%d value %.3g %s  step %.1f us (steady %.1f)  fp64 frac %.2f hbm frac %.3f  parity lp %.1e g %.1e" % (
        path, d["n_gpus"], d["value"], d["unit"], d["ms_per_step"] * 1e3, d["steady_state"]["ms_per_step"] * 1e3,
        d["fp64_roofline"]["frac"] or 0, d["roofline"]["frac"], p.get("logp_rel", -1), p.get("grad_rel", -1)))
    print("   e2e %.3g (%.1f us/step, floor %.1f us)" % (e.get("value", 0), e.get("ms_per_step", 0) * 1e3, e.get("pcie_floor_ms", 0) * 1e3))
    if s:
        print("   sampler: %.1f us/tick  %.3g chain-leapfrogs/s  ESS/s %.0f  accept %.3f depth %.2f  [%s]" % (
            s["us_per_tick"], s["chain_leapfrogs_per_sec"], s["ess_per_sec"], s["mean_tree_accept"], s["mean_depth"], s["engine"]))
