"""Short view of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_n1.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "roofline", d.get("roofline", {}).get("frac"))
print("e2e", d.get("e2e"), "h2d_ms", (d.get("h2d_probe") or {}).get("ms"))
print("stage_ms", d.get("stage_ms"), "clocks", d.get("clocks"))
if d.get("tsqr_mode"):
    print("tsqr", {k: d["tsqr_mode"][k]["ms_per_step"] for k in ("tsqr", "tsqr_full")})
if d.get("c5"):
    print("c5", {m: (d["c5"][m]["wall_s"], d["c5"][m]["frac"], d["c5"][m]["status"]) for m in ("gram2", "tsqr", "tsqr_full")}, d["c5"].get("parity"))
if d.get("parity"):
    print("parity", d["parity"]["status"])
