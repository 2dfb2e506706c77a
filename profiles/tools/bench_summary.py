import sys, json
d = json.loads(sys.stdin.read())
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "k1m ms", d["roofline"]["kernel_ms"], "fused", d["fused_fixup_dp"]["kernel_ms"])
print("e2e", d["e2e"]["ms_each"])
for s in d.get("secondary", []):
    print(s["workload"][:3], "ms", s["ms_per_step"], "k1", s["kernel_ms"], "frac", s["roofline_frac"], "e2e", s["e2e"]["ms_each"])
