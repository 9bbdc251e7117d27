import json,sys
for f in sys.argv[1:]:
    try:
        l=json.loads(open(f).read().strip().splitlines()[-1])
        r=l["roofline"]
        print(f.split('/')[-1], "%.3fM img/s"%(l["value"]/1e6), "step %.1f us"%(l["ms_per_step"]*1e3), "serial %.1f"%(r.get("serial_ms_per_step",0)*1e3), {k:round(v*1e3,1) for k,v in r["stage_ms_per_step"].items()}, "K3 frac %.3f"%r["frac"], "pipe %.3f"%r["pipeline_frac"], "e2e %.0f"%l["e2e"]["value"], "hum %.1f"%l["config"]["humans_per_image"])
    except Exception as e:
        print(f, "ERR", e); print(open(f).read()[-600:])
