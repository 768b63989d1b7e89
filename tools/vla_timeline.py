"""Summarise gpurun_out/vla_times.csv (ROBOVLN_VLA_TIMES): per-phase SM-clock intervals of the fused block kernel."""
import sys
import numpy as np
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/vla_times.csv"
lines = [l for l in open(path) if l.strip()]
launches = []
for l in lines:
    if l.startswith('#'):
        launches.append([]); continue
    launches[-1].append([int(x) for x in l.strip().split(',')])
t = np.array(launches[-1])[:, 1:].astype(np.int64)
names = {0: 'mma:start', 1: 'mma:inputs landed', 2: 'mma:S issued', 3: 'mma:P ready', 4: 'mma:ctx ready', 5: 'mma:fc_o issued', 6: 'mma:X ready', 7: 'mma:all issued',
         32: 'epi:start', 33: 'epi:S done', 34: 'epi:softmax done', 35: 'epi:O done', 36: 'epi:ctx written', 37: 'epi:fc_o done', 38: 'epi:ffn epilogues done', 39: 'epi:Y done', 56: 'epi:end', 57: 'epi:LN2 written', 58: 'epi:LN2 barrier', 59: 'epi:LN2 pass1 done', 60: 'epi:LN2 stats exchanged', 61: 'epi:LN1 pass1 done', 62: 'epi:LN1 stats exchanged'}
for s in range(16): names[8 + s] = f'mma:ffn step{s} begin'
for c in range(8): names[40 + 2 * c] = f'epi:chunk{c} wait'; names[41 + 2 * c] = f'epi:chunk{c} hfull'
ctas = [int(a) for a in sys.argv[2:]] or [0]
for cta in ctas:
    r = t[cta]; t0 = r[0] if r[0] else r[32]
    print('CTA', cta)
    prev = 0
    for dt, i in sorted([(r[i] - t0, i) for i in range(64) if r[i] != 0]):
        print(f'   {dt:8d} (+{dt - prev:6d}) {names.get(i, i)}'); prev = dt
tot = t[:, 56] - t[:, 32]
print('per-CTA total cycles: min %d median %d max %d' % (tot.min(), np.median(tot), tot.max()))
