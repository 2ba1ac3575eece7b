import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        j=json.loads(l)
        rc=j.get("roofline_chain") or {}
        print("days/s %.1f  ms/step %.2f  e2e %.1f  chain_frac %.3f  chain_ms/day %.4f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], rc.get("frac",0), rc.get("chain_kernel_ms_per_day",0)))
        print({k: round(v/ (j["config"]["days"]-1)*1000,1) for k,v in rc.get("kernel_ms",{}).items()})
    else: print(l.rstrip())
