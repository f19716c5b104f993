"""Phase profile of encoder3_kernel from an ncu source-page CSV (ncu -i rep --page source --csv): warp samples split at
every mbarrier wait; the wait loop is reported separately from the code that follows it.  Barrier names from the
shared-memory offsets of tc_encoder3.cu (MISC_BARS = 0x37b00; per-stream blocks are addressed through a register)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index('Address'); isamp = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
R = [r for r in rows[2:] if len(r) >= len(hdr)]
base = int(R[0][ia], 16)
per = [(int(r[ia], 16) - base, int(r[isamp] or 0), r[1].strip(), int(r[ii] or 0)) for r in R]
tot = sum(p[1] for p in per)
nsl = float(sys.argv[2]) if len(sys.argv) > 2 else 98304.0   # sequence-layers per launch
shared = {0: 'X_FULL', 8: 'X_DONE', 16: 'ATTN_DONE', 24: 'QKV_READY', 32: 'QKV_FREE', 40: 'VEC_FULL', 48: 'BIAS_FULL0', 56: 'BIAS_FULL1',
          64: 'W_FULL0', 72: 'W_FULL1', 80: 'W_FULL2', 88: 'W_FULL_IN', 96: 'W_EMPTY0', 104: 'W_EMPTY1', 112: 'W_EMPTY2', 120: 'W_EMPTY_IN'}
stream = {0: 'QKV_DONE', 8: 'S_DONE0', 16: 'S_DONE1', 24: 'P_READY0', 32: 'P_READY1', 40: 'PV_DONE', 48: 'O_READY', 56: 'OUT_DONE', 64: 'X1_READY',
          72: 'F1_DONE', 80: 'F1_FREE', 88: 'HID_READY0', 96: 'HID_READY1', 104: 'F2_DONE0', 112: 'F2_DONE1', 120: 'X2_READY'}
def name(ins):
    m = re.search(r'\[(U?R\d+)(?:\+URZ)?(?:\+0x([0-9a-f]+))?\]', ins)
    if not m: return ins[-30:]
    off = int(m.group(2), 16) if m.group(2) else None
    if off is None: return '[%s] (computed)' % m.group(1)
    b = off - 0x37b00
    if b < 0: return '[%s+0x%x]' % (m.group(1), off)
    if m.group(1).startswith('UR'): return shared.get(b, str(b))
    return ('s.' + stream.get(b - 128, str(b))) if b >= 128 else shared.get(b, str(b))
segs = []; cur = ['start', 0, 0, 0, 0]; i = 0
while i < len(per):
    off, s, ins, n = per[i]
    if 'TRYWAIT' in ins:
        segs.append(cur)
        w = s; j = i + 1
        while j < len(per) and j < i + 12:
            w += per[j][1]
            if 'BRA' in per[j][2]: j += 1; break
            j += 1
        cur = [name(ins), off, w, 0, 0]; i = j; continue
    cur[3] += s; cur[4] += n; i += 1
segs.append(cur)
print("total samples", tot)
for lab, off, w, c, n in segs:
    if (w + c) / tot > 0.003: print(f"{off:6x} wait {w / tot:6.2%} code {c / tot:6.2%} inst/seq-layer {n / nsl:8.0f}  after [{lab}]")
