import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value %.3f it/s" % d["value"], "ms/step %.1f" % d["ms_per_step"], "newton/step", d.get("newton_iterations_per_step"), "lin/newton", d.get("linear_iterations_per_newton"), d.get("linear_iterations"), "conv", d.get("converged"))
    if "kernels" in d:
        for k,v in d["kernels"].items(): print("  ", k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
        e=d["e2e"]; print("  e2e %.3f it/s ms/step %.1f" % (e["value"], e["ms_per_step"]), "h2d", e["h2d_bytes_per_step"], "d2h", e["d2h_bytes_per_step"])
        print("  cpu", d["cpu_baseline"]); print("  clocks", d["clocks"], "launches", d["gpu_launches"])
    else:
        print("  cpu", d["cpu_baseline"])
