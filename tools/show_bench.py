"""Print the headline fields of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    e = d["e2e"]
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "| e2e", round(e["value"]), "ms", round(e["ms_per_step"], 1),
          "lib wall", round(e.get("library_wall_ms_per_step", 0.0), 1), "| launches", d["gpu_launches"], "roofline frac", round(d["roofline"]["frac"], 3),
          "| stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items() if v > 0.3})
