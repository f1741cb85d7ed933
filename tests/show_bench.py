import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["phase_ms"]); w=d["work"]; print('render Gs/s', w["render_samples_shaded"]/d["kernel_breakdown_ms"]["mve_render_rays"]["ms"]/1e6, w["render_samples_shaded"])
print({k:v["ms"] for k,v in d["kernel_breakdown_ms"].items() if v["ms"]>5}); print(d["per_call_ms"])
