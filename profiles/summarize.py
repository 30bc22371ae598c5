"""print one-line summaries of bench.py JSON lines: python profiles/summarize.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        line = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("impl") == "reference":
        print(f, "| REF value", round(d["value"], 3), d["unit"], "|", d["config"]["workload"][:150]); continue
    r = d["roofline"]
    print(f, "| ms/step", round(d["ms_per_step"], 3), "| e2e ms", round(d["e2e"]["ms_per_step"], 3), "| trsv ms", round(r["ms"], 3), "GB/s", round(r["achieved"]), "frac", round(r["frac"], 3),
          "| factor GB", round(d["config"]["factor_gb"], 2), "levels", d["config"]["levels"], "fronts", d["config"]["fronts"], "| numfact", d["config"]["numfact_s"],
          "| launches", d["gpu_launches"], "|", d.get("krylov", ""), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
