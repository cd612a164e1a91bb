import json,sys
d=json.load(open(sys.argv[1]))
print("value",d["value"],"kernel_only",d.get("value_kernel_only"),"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"] if d.get("e2e") else None, d.get("sharded_result_check"))
for k in ("c3_shard","c5_shard","related_branch","c4_resolutions","bed_writer","reference_selection","bgen_ingest","haps_ingest","cpu_baseline"):
    v=d.get(k)
    if v is None: print(k,None); continue
    if "error" in v: print(k,v); continue
    if k=="c4_resolutions": print(k,[(l["mean_group_size"],round(l["kernel_ms"],2),round(l["roofline"]["frac"],3),round(l["set_groups_host_ms"],1)) for l in v["levels"]]); continue
    print(k,{kk:(vv if not isinstance(vv,dict) else {a:b for a,b in vv.items() if a in ("frac","kernel_ms","ms","hap_snps_per_s","sm_mhz","reasons")}) for kk,vv in v.items() if kk not in ("workload","note","sample")})
