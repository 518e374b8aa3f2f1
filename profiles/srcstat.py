#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample shares of one kernel from an .ncu-rep captured with
--import-source on (kernels are built with -lineinfo).  usage: python profiles/srcstat.py x.ncu-rep [top_n]"""
import csv,sys,subprocess,io
rep=sys.argv[1]
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
hdr=None; out=[]
for r in rows:
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)>8 and r[0].isdigit():
        d=dict(zip(hdr,r))
        try: out.append((int(d["Instructions Executed"]),int(d["# Samples"]),int(r[0]),r[1][:110]))
        except: pass
tot=sum(o[0] for o in out); ts=sum(o[1] for o in out)
print("total inst",tot,"samples",ts)
for o in sorted(out,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print("%5.1f%% inst %5.1f%% smp  L%-5d %s"%(100*o[0]/tot,100*o[1]/max(ts,1),o[2],o[3]))
