import sys
ev=[tuple(map(int,l.split()[1:])) for l in open(sys.argv[1]) if l.startswith('TL')]
ev.sort(key=lambda e:e[1])
t0=ev[0][1]
names={100:'M QKV issue start',110:'M QKV issued',120:'M QKV_READY seen',130:'M S_LOADED seen',140:'M P_READY seen/PV start',150:'M PV issued',160:'M out_proj start',161:'M out_proj issued',170:'M X1_READY seen',180:'M F1 issued',199:'M F2(3) issued',
200:'C wait QKV_DONE',210:'C QKV_DONE seen',220:'C QKV epi done',230:'C S_DONE seen',240:'C S loaded',250:'C exps done',260:'C PV_DONE(t-1) seen',270:'C P stored',280:'C wait OUT_DONE',281:'C OUT_DONE seen',282:'C LN1 done',300:'C F1_DONE seen',320:'C GELU stored',340:'C F2 complete seen',341:'C LN2 done'}
def nm(i):
    for base in sorted(names,reverse=True):
        if i>=base and i-base<20 and (base>=300 or base==180 or i-base<10): return f"{names[base]} [{i-base}]"
    return str(i)
prev=t0
for i,t in ev:
    print(f"{t-t0:8d} (+{t-prev:6d})  {nm(i)}")
    prev=t
