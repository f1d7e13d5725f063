import json, sys, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, round(d["ms_per_step"], 2), "ms/fit", round(d["value"]), "frames/s | gemm avg ms", round(r["avg_launch_ms"], 4), "launches", r["launches"],
              "| sm_mhz", d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
