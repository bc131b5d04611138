"""Opcode histogram of the loop bodies of one kernel of the shipped library (static; no GPU needed):

    cuobjdump -sass poreseq_b200/libporeseq_b200.so > /tmp/all.sass
    python scripts/sass_loops.py _ZN5psdev6k_fillILi160ELi3EEEvNS_5BatchEi 8

A loop body = the instructions between a backward BRA and its target; the largest loops are printed first."""
import re, sys, collections
fn = sys.argv[1]
lines = open("/tmp/all.sass").read().split("\n")
start = next(i for i,l in enumerate(lines) if "Function : "+fn in l)
end = next((i for i in range(start+1,len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[start:end]:
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1),16), m.group(2)))
addr_idx = {a:i for i,(a,_) in enumerate(ins)}
loops = []
for i,(a,t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1),16)
        if tgt <= a and tgt in addr_idx: loops.append((addr_idx[tgt], i))
print(fn, "instructions", len(ins), "backward branches", len(loops))
def op(t):
    t = re.sub(r"^@!?U?P\w+\s+", "", t)
    return t.split()[0].split(".")[0]
for (s,e) in sorted(loops, key=lambda x: x[0]-x[1])[:int(sys.argv[2]) if len(sys.argv)>2 else 6]:
    c = collections.Counter(op(t) for _,t in ins[s:e+1])
    n = e-s+1
    f64 = sum(v for k,v in c.items() if k in ("DADD","DMUL","DFMA","DSETP"))
    print("loop %05x..%05x  n=%d  f64=%d  " % (ins[s][0], ins[e][0], n, f64), c.most_common(24))
