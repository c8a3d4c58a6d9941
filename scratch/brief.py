import json, sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
def brief(n, r):
    if "error" in r: print(n, "ERROR", r["error"]); return
    rl=r.get("roofline",{})
    print(n, "ms/step", round(r["ms_per_step"],4), "value", round(r["value"]), "frac", round(rl.get("frac",0),3), "parity", {k:v for k,v in (r.get("parity") or {}).items() if k in ("steps_checked","occupancy_equal","semantic_equal","all_ranks_equal")}, "e2e", round(r["e2e"]["value"]) if "e2e" in r else None, "cpu", round(r["cpu_baseline"]["value"],1) if "cpu_baseline" in r else None, r.get("tour") or "", r.get("load_s") or "")
brief(j["config"]["workload"].split(":")[0], j)
for n,r in j.get("configs",{}).items(): brief(n,r)
print("clocks", j.get("clocks")); print("e2e", j.get("e2e")); print("cpu", j.get("cpu_baseline"))
