import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/selftest.json"))
keep = ("n_L", "n_L_ref", "angle", "sig_err_kept", "sig_err_all", "ms_eig", "ms_pass2", "ms_total", "ms_gram", "wall_s",
        "rank_mismatch", "gram_tflops", "ms_project", "ms_podsig", "ms_basis")
for sec, v in d.items():
    if not isinstance(v, dict):
        continue
    print("==", sec)
    if "error" in v:
        print(v["error"], v.get("trace", "")[-800:]); continue
    for k, r in v.items():
        if isinstance(r, dict):
            print(f"  {k:34s}", {a: (f"{b:.2e}" if isinstance(b, float) else b) for a, b in r.items() if a in keep})
        else:
            print(f"  {k:34s} {r}")
