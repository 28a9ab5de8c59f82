"""Top SASS instructions of an `ncu --page source --csv` dump by executed count and by stall samples -- optimisation aid.
    python scripts/ncu_source_top.py <source.csv> [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_samp = sum(int(r[ix['# Samples']]) for r in data)
print(f'total warp instructions {tot_inst}, samples {tot_samp}')
ops = collections.Counter()
for r in data:
    op = r[ix['Source']].split()[0] if not r[ix['Source']].strip().startswith('@') else r[ix['Source']].split()[1]
    ops[op.split('.')[0]] += int(r[ix['Instructions Executed']])
print('by opcode:', ', '.join(f'{k} {v / tot_inst:.1%}' for k, v in ops.most_common(18)))
print('--- top by stall samples')
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:N]:
    st = {k[6:]: int(r[ix[k]]) for k in hdr if k.startswith('stall_') and '(' not in k and int(r[ix[k]]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{int(r[ix['# Samples']]):7d} {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:70]:70s} {top}")
