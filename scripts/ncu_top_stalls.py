"""Reads an `ncu --page source --csv` dump: stall totals, instruction mix per warp-step and the instructions with the most samples.
usage: python scripts/ncu_top_stalls.py dump.csv [n_top] [warp_steps]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wsteps = float(sys.argv[3]) if len(sys.argv) > 3 else 524288.0
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
ia, isamp, ie = col['Source'], col['# Samples'], col['Instructions Executed']
body = rows[2:]
tot = sum(int(r[isamp]) for r in body)
print('total samples', tot, 'instructions', len(body), 'executed/warp-step', sum(int(r[ie]) for r in body) / wsteps)
for name in ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_branch_resolving', 'stall_lg', 'stall_barrier', 'stall_no_inst', 'stall_membar']:
    print(f"  {name:24s} {sum(int(r[col[name]]) for r in body):7d} {100.0 * sum(int(r[col[name]]) for r in body) / tot:5.1f} %")
mix = collections.Counter()
for r in body:
    s = re.sub(r'^@!?U?P\d+\s+', '', r[ia].strip())
    mix[s.split()[0]] += int(r[ie])
print('mix per warp-step:', ', '.join(f"{op} {n / wsteps:.1f}" for op, n in mix.most_common(24)))
top = sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:ntop]
for i in sorted(top):
    r = body[i]
    print(f"{i:5d} {int(r[isamp]):6d} L{int(r[col['stall_long_sb']]):6d} S{int(r[col['stall_short_sb']]):5d} W{int(r[col['stall_wait']]):5d} M{int(r[col['stall_math']]):4d} x{int(r[ie]):8d}  {r[ia].strip()[:100]}")
