"""Print the key numbers of a bench.py JSON line read from stdin (helper for gpurun one-liners)."""
import json
import sys

tag = " ".join(sys.argv[1:])
for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print("%s | %s = %.4g %s | ms/step %.4f | roofline frac %.4f (kernel %.4f ms) | e2e %.4g | launches %s" % (
        tag, d["metric"], d["value"], d["unit"], d["ms_per_step"], r.get("frac", float("nan")),
        r.get("kernel_ms", float("nan")), d["e2e"]["value"], d.get("gpu_launches")))
    if "icp" in d:
        i = d["icp"]
        print("    icp %.4g pairs/s, ms/step %.4f, e2e %.4g" % (i["value"], i["ms_per_step"], i["e2e"]["value"]))
    if "grid" in d:
        g = d["grid"]
        print("    grid %.4g beams/s, frac %.4f, e2e %.4g" % (g["value"], g["roofline"]["frac"], g["e2e"]["value"]))
