"""One-line digest of bench.py output files: python tools/show_bench.py a.json [b.json ...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as exc:
        print(f, "unreadable:", exc)
        continue
    e = d.get("e2e", {})
    r = d.get("roofline", {}) or {}
    pg = d.get("e2e_pageable") or {}
    print(f, "| value", round(d["value"]), "genes/s, ms", round(d["ms_per_step"], 2), "(events", round(d.get("device_event_ms_per_step", 0), 2), ")",
          "| e2e", round(e.get("value", 0)), "ms", round(e.get("ms_per_step", 0), 1), "| pageable", round(pg.get("value", 0)) if pg else None,
          "| launches", d.get("gpu_launches"), "| roofline", r.get("bound"), round(r["frac"], 3) if r.get("frac") else None,
          "| stages", {k: round(v, 2) for k, v in d.get("stage_ms_per_step", {}).items() if v > 0.3},
          "| cpu", round(d["cpu_baseline"]["value"], 2) if "cpu_baseline" in d else None)
