"""Print a one-line summary of bench.py JSON lines: python tools/brief.py file..."""
import json, sys
for p in sys.argv[1:]:
    for line in open(p):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        print(p, "value=%.0f e2e=%.0f ms/step=%.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]),
              "dom=%s frac=%.3f" % (r.get("kernel"), r.get("frac", 0)),
              {k: round(v, 3) for k, v in (r.get("stage_ms_per_step") or {}).items()})
