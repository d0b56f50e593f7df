import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
rs=d["e2e_ref_signature"]
for k,v in rs.items():
    if isinstance(v,dict) and "value" in v: print(" ", k, round(v["value"]), "ms/call", round(v["ms_per_call"],3), "merge", round(v["merge_tile_ms_per_call"],3), "dev", round(v["device_ms_per_call"],3), "passes", v.get("merge_passes"), "calls/wavefront", round(v.get("calls_per_wavefront", 0), 2))
