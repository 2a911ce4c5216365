import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: v for k, v in d["e2e"].items() if k not in ("api", "unit", "value")})
if "e2e_records" in d: print("   e2e_records", round(d["e2e_records"]["value"]))
for k, v in d["roofline_kernels"].items(): print("  ", k, round(v["avg_launch_ms"], 4), round(v["frac"], 3))
print("   cpu", d.get("cpu_baseline", {}) and round(d["cpu_baseline"]["value"]), d["clocks"])
