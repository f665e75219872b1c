import json,sys
# usage: show_bench.py FILE   (or JSON lines on stdin)
for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    line=line.strip()
    if not line.startswith('{'): 
        print(line); continue
    d=json.loads(line)
    print("value %.3e pts/s  ms/step %.4f  e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    r=d["roofline"]; print("kernels_ms", {k:round(v,4) for k,v in r["kernels_ms"].items()}); print("kernels_gbs", {k:round(v) for k,v in r["kernels_gbs"].items()}, "step_frac %.3f"%r["step_frac"], "dominant", r["kernel"], "frac %.3f"%r["frac"])
    for k,v in d.get("extra",{}).items():
        if isinstance(v,dict) and "points_per_s" in v and "ms" in v:
            print(k, "%.3e pts/s  %.4f ms  hbm_frac %.3f"%(v["points_per_s"], v["ms"], v.get("hbm_frac",0)), "fp32_frac", v.get("fp32_frac",""), v.get("with_dz",""))
        else: print(k,v)
    print("clocks", d["clocks"], "cpu", d.get("cpu_baseline"))
