import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        e = d["e2e"]
        print("value", d["value"] / 1e9, "frac", d["roofline"]["frac"], "verified", d["verified_state_digest"])
        print("e2e", e["value"] / 1e9, "slab", e["slab_steps"], e["by_slab_steps"], "int32", e.get("int32_actions", {}).get("value"),
              "full", e.get("full_d2h", {}).get("value"), e.get("host_affinity"))
        print("slab_launches", d.get("slab_launches"))
