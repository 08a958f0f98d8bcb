"""Pretty-prints the key fields of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("value %.1f %s | %.2f ms/step | e2e %.1f | launches %s | clocks %s" % (
    d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks")))
print("dominant kernel: %.1f TFLOP/s = %.3f of %s (%.1f us) | step %.1f TFLOP/s = %.3f of sustained" % (
    r.get("achieved", 0), r.get("frac", 0), r.get("peak"), (r.get("kernel_ms") or 0) * 1e3, r.get("step_tflops", 0),
    r.get("step_frac_of_sustained", 0)))
if d.get("cpu_baseline"):
    print("cpu baseline:", d["cpu_baseline"])
