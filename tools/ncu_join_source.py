import csv,re,collections,sys
src_csv=sys.argv[1]
# address -> source line from nvdisasm
lines=open('/tmp/enc_lines.sass').read().split('\n')
infn=False; cur=None; addr2line={}
for l in lines:
    if '.text.' in l and ('.section' in l or l.startswith('.text.')): infn='encoder_kernel' in l
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2)))
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if infn and m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(src_csv)))
hdr=rows[1]
ia=hdr.index('Address'); isamp=hdr.index('# Samples'); iinst=hdr.index('Instructions Executed'); isrc=hdr.index('Source')
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base=None
agg=collections.defaultdict(lambda:[0,0,collections.Counter()])
tot=0
for r in rows[2:]:
    if len(r)<len(hdr): continue
    a=int(r[ia],16) if r[ia].startswith('0x') else int(r[ia])
    if base is None: base=a
    off=a-base
    ln=addr2line.get(off)
    s=int(r[isamp] or 0); n=int(r[iinst] or 0)
    agg[ln][0]+=s; agg[ln][1]+=n; tot+=s
    for i in stall_cols:
        v=int(r[i] or 0)
        if v: agg[ln][2][hdr[i]]+=v
print("total samples",tot)
for ln,(s,n,st) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:40]:
    print(ln, s, f"{s/tot:5.1%}", "inst",n, dict(st.most_common(3)))
print("---- by region")
regions=[(100,112,'ex2/rcp/gelu_fast math'),(113,160,'helpers ld/st shared, ldc'),(161,196,'MMA issue helpers'),(197,222,'epi_qkv'),(223,270,'softmax load/exp/store'),(271,286,'epi_o'),(287,346,'epi_ln'),(347,380,'act load/store'),(381,395,'Phase/params'),(396,420,'kernel prologue'),(421,448,'producer'),(449,548,'MMA issuer'),(549,670,'compute schedule')]
reg=collections.Counter(); regi=collections.Counter()
for ln,(s,n,st) in agg.items():
    if ln is None: reg['(none)']+=s; continue
    f,l=ln
    if f=='tc_ptx.cuh':
        key='ptx: mbar wait/arrive' if 28<=l<=75 else ('ptx: tmem ld/st/mma' if l>=134 else 'ptx: other')
    elif f=='tc_layout.cuh': key='pack_bf16x2'
    else:
        key=next((name for a,b,name in regions if a<=l<=b), f'other {l}')
    reg[key]+=s; regi[key]+=n
for k,v in reg.most_common(): print(f"{k:32s} {v:8d} {v/tot:6.1%}  inst {regi[k]}")
