"""Print the headline fields of a bench.py JSON line (developer convenience)."""
import json, sys
txt = open(sys.argv[1]).read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("value %.4g %s | ms/step %.3f | LM it/s %.0f | e2e %.4g (%.3f ms) | launches %d" % (
  d["value"], d["unit"], d["ms_per_step"], d.get("lm_iters_per_sec", 0), d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches", 0)))
print("clocks", d.get("clocks"))
r = d.get("roofline", {})
print("roofline frac %.4f achieved %.1f GB/s launch_ms %.4f traffic %s" % (r.get("frac", 0), r.get("achieved", 0), r.get("launch_ms", 0), r.get("traffic")))
print("cpu_baseline", d.get("cpu_baseline", {}).get("value"))
