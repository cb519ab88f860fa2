import json, sys
for f in sys.argv[1:]:
    txt = open(f).read()
    lines = [l for l in txt.splitlines() if l.startswith("{")]
    if not lines: print(f, "no json"); continue
    d = json.loads(lines[-1])
    print(f, "value %.3f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "newton", d["newton_iterations_per_step"], "lin", d["linear_iterations"], "part_s %.1f" % d["config"].get("partition_seconds", 0), "ilu", d["config"].get("ilu"))
    print("   sizes", d["config"].get("owned_ghost_per_rank"))
    for k, v in d["kernels"].items(): print("   ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
    print("   e2e %.3f ms/step %.1f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "launches", d["gpu_launches"], d["clocks"])
