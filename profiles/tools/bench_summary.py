import sys, json
d = json.loads(sys.stdin.read())
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "k1m ms", d["roofline"]["kernel_ms"], "fused", d["fused_fixup_dp"]["kernel_ms"])
print("e2e", d["e2e"]["ms_each"])
for s in d.get("secondary", []):
    print(s["workload"][:3], "ms", s["ms_per_step"], "k1", s["kernel_ms"], "frac", s["roofline_frac"], "e2e", s["e2e"]["ms_each"])
print("sha", d.get("result_sha256_24"), "shard_check", d.get("shard_check"), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]) if d["e2e"].get("value") == d["e2e"].get("value") else None)
