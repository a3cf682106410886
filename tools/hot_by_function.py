#!/usr/bin/env python
"""Hot SASS instructions of one captured kernel by source function: tools/hot_by_function.py REP TRIPS MIN_PER_TRIP
(companion of tools/hot_footprint.py; rows of the cuda,sass view, so an instruction inlined under several lines counts once per line)."""
import sys, csv, io, subprocess, collections, os
sys.path.insert(0,'tools')
from ncu_regions import function_map
rep, trips, f = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
fmaps={}; hdr=None; cur_file='?'; cur_fn='?'
fn=collections.defaultdict(lambda: [0,0,0.0,set()])
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0]=="File Path":
        cur_file=r[1]
        if cur_file not in fmaps: fmaps[cur_file]=function_map(cur_file.replace('/tmp/code/AMReX-Combustion__PeleLM/repo','/root/repo'))
        continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No":
        hdr={c:i for i,c in enumerate(r)}; continue
    if hdr is None or len(r)<len(hdr): continue
    if r[0]!="":
        fm=fmaps.get(cur_file); base=os.path.basename(cur_file)
        cur_fn = fm.get(int(r[0]),base) if (fm and "pelelm_b200" in cur_file) else base
        continue
    try: n=int(r[hdr["Instructions Executed"]])
    except ValueError: continue
    addr=int(r[hdr["Address"]],16) if r[hdr["Address"]].startswith('0x') else int(r[hdr["Address"]])
    e=fn[cur_fn]; e[0]+=1
    if n>=f*trips:
        e[1]+=1; e[3].add(addr//128)
    e[2]+=n/trips
tot=sum(e[1] for e in fn.values())
print("function  sass  hot_sass  hotKB  lines  exec/step")
for k,e in sorted(fn.items(), key=lambda kv:-kv[1][1])[:45]:
    print("%-28s %5d %5d %6.1f %4d %8.1f"%(k[:28],e[0],e[1],e[1]*16/1024,len(e[3]),e[2]))
print("total hot", tot, tot*16/1024)
