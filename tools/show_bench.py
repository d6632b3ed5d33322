import json, sys
d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f}  launches {d['gpu_launches']}  clocks {d['clocks']}")
print("roofline", d["roofline"])
print("whole_net", {k: round(v, 4) for k, v in d["whole_net"].items()})
for k, v in d["kernels"].items():
    print(f"{k:22s} share {v['share']:.3f} ms/launch {v['ms_per_launch']:.3f} n/step {v['launches_per_step']:.0f} "
          f"TF {v['tflops']:.1f} GB/s {v['gbs']:.0f}  by block {v.get('ms_per_launch_by_block', {})}")
if d.get("align"):
    a = d["align"]
    print("align", round(a["value"]), a["unit"], "ms/step", round(a["ms_per_step"], 3), "e2e", round(a["e2e"]["value"]),
          "cpu", a.get("cpu_baseline"))
print("cpu_baseline", d.get("cpu_baseline"))
