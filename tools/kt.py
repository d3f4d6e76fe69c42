import json,sys
d=json.load(open(sys.argv[1]));print(f"value {d['value']:.4g} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:.4g}")
print(" ".join(f"{k}={v['ms_per_call']*1000:.1f}" for k,v in d["kernels"].items()))
