import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
# multiple kernels concatenated: take the first kernel block
hdr=None; data=[]; nk=0
for r in rows:
    if r and r[0]=="Kernel Name":
        nk+=1
        if nk>int(sys.argv[2]) if len(sys.argv)>2 else nk>1: break
        data=[]; continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr and len(r)==len(hdr): data.append(dict(zip(hdr,r)))
tot=sum(int(d["# Samples"]) for d in data); ex=sum(int(d["Instructions Executed"]) for d in data)
print("instrs static",len(data),"samples",tot,"warp-instr executed",ex)
# segment by cumulative: print regions of 40 instructions with their sample share and executed count
seg=int(sys.argv[3]) if len(sys.argv)>3 else 50
for i in range(0,len(data),seg):
    blk=data[i:i+seg]
    s=sum(int(d["# Samples"]) for d in blk); e=sum(int(d["Instructions Executed"]) for d in blk)
    ops={}
    for d in blk:
        op=d["Source"].split()[0] if not d["Source"].strip().startswith("@") else d["Source"].split()[1]
        op=op.split(".")[0]; ops[op]=ops.get(op,0)+1
    top=sorted(ops.items(), key=lambda kv:-kv[1])[:4]
    # dominant stall
    st={}
    for d in blk:
        for k,v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v not in ("","0"): st[k]=st.get(k,0)+int(v)
    st=sorted(st.items(), key=lambda kv:-kv[1])[:3]
    print(f"{i:5d}-{i+len(blk):5d}  samples {100*s/tot:5.1f}%  exec {100*e/ex:5.1f}%  {top}  {[(k[6:],round(100*v/max(s,1))) for k,v in st]}")
