"""Phase profile of encoder_kernel from an ncu source-page CSV: samples split at every synchronisation instruction
(mbarrier try_wait loop / bar.sync); wait loops reported separately from the code that follows them.
usage: ncu -i rep --page source --csv > x.csv; python tools/ncu_phases.py x.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index('Address'); isamp = hdr.index('# Samples'); iinst = hdr.index('Instructions Executed')
per = []; base = None; tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    a = int(r[ia], 16)
    if base is None: base = a
    s = int(r[isamp] or 0); tot += s
    per.append((a - base, s, r[1].strip(), int(r[iinst] or 0)))
names = {0:'X_FULL',8:'X_FREE',16:'ATTN_DONE',24:'W_FULL0',32:'W_FULL1',40:'W_FULL2',48:'W_FULL3',56:'W_EMPTY0',64:'W_EMPTY1+',88:'QKV_DONE',96:'QKV_READY',104:'S_DONE',112:'S_LOADED',120:'P_READY',128:'PV_DONE',136:'O_FREE',144:'O_FREE1',152:'OUT_DONE',160:'X1_READY',168:'F1_DONE',176:'F1_FREE',184:'F2_DONE0',192:'F2_DONE1',200:'F2_DONE2',208:'HID_READY',216:'HID_READY1',224:'X2_READY',232:'VEC_FULL',240:'BIAS_FULL'}
segs = []; cur = ['start', 0, 0, 0, 0]   # label, start, wait samples, code samples, code inst
i = 0
while i < len(per):
    off, s, ins, n = per[i]
    if 'TRYWAIT' in ins or 'BAR.SYNC' in ins:
        segs.append(cur)
        m = re.search(r'\+0x([0-9a-f]+)\]', ins)
        bar = int(m.group(1), 16) - 0x37b00 if m else None
        if bar is not None and bar < 0:   # second barrier block (OFF_BAR2 = 0x11c00): per-row-tile hand-offs
            b2 = bar + 0x37b00 - 0x11c00
            bar = ("OUT_DONE[%d]" % (b2 // 8)) if b2 < 24 else ("X1_READY[%d]" % ((b2 - 24) // 8)) if b2 < 48 else ("X2_READY[%d]" % ((b2 - 48) // 8))
        label = ('wait ' + str(names.get(bar, bar))) if 'TRYWAIT' in ins else 'bar.sync'
        w = s
        # the spin loop: following instructions up to and including the backward branch
        j = i + 1
        while 'TRYWAIT' in ins and j < len(per) and j < i + 12:
            w += per[j][1]
            if 'BRA' in per[j][2]: j += 1; break
            j += 1
        cur = [label, off, w, 0, 0]
        i = j if 'TRYWAIT' in ins else i + 1
        continue
    cur[3] += s; cur[4] += n
    i += 1
segs.append(cur)
print("total samples", tot)
for label, off, w, c, n in segs:
    if (w + c) / tot > 0.003:
        print(f"{off:6x} wait {w / tot:6.2%} code {c / tot:6.2%} (inst {n:>11}) [{label}]")
