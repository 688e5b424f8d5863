import json,csv,sys
def bench(files):
    for f in files:
        try: lines=open(f).read().splitlines()
        except Exception as e: print(f, e); continue
        for l in lines:
            if l.startswith("{"):
                d=json.loads(l)
                print(f, d["metric"][:2], d["config"]["points_per_gpu"], "c",d["config"]["window_bits"], "value %.1fM ms %.2f | e2e %.1fM ms %.2f | acc %.2f frac %.2f"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["e2e"]["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"]))
            elif "Error" in l or "rror:" in l: print(f, l[:200])
def launches(f,a,b):
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    data=[(r[ki].split("(")[0].replace("void ","").replace("b200::","")[:40], float(r[vi])/1e3) for r in rows[1:]]
    print(f, len(data))
    for i,(k,v) in enumerate(data):
        if a<=i<b: print(i, "  %-40s %9.1f us"%(k,v))
if sys.argv[1]=="bench": bench(sys.argv[2:])
else: launches(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
