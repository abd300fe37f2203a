#!/usr/bin/env python
"""profiles/r02_sass_mnemonics.md: SASS mnemonic counts per kernel of libmdsf.so (cuobjdump -sass), the evidence for
TMA bulk copies (UBLKCP), mbarriers (SYNCS), cp.async (LDGSTS) and the absence of atomics in the splat.  No GPU needed."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else 'md-structure-factor_b200/libmdsf.so'
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
fn, per = None, collections.defaultdict(collections.Counter)
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m: fn = m.group(1); continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn: per[fn][m.group(1).split('.')[0]] += 1
want = ['UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'LDGSTS', 'ATOMS', 'ATOMG', 'RED', 'DFMA', 'DADD', 'DMUL', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'WARPSYNC', 'SHFL']
names = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.split('\n')
dem = {k: n.split('(')[0][:70] for k, n in zip(per, names)}
out = ['# SASS mnemonic counts per kernel of libmdsf.so (cuobjdump -sass, sm_100a); UBLKCP = cp.async.bulk (TMA copy engine), SYNCS = mbarrier, LDGSTS = cp.async', '',
       '| kernel | ' + ' | '.join(want) + ' |', '|---|' + '---|' * len(want)]
for k, c in sorted(per.items(), key=lambda kv: dem[kv[0]]):
    if 'cub::' in dem[k] or 'DeviceScan' in dem[k]: continue
    out.append('| `%s` | ' % dem[k] + ' | '.join(str(c.get(w, 0)) for w in want) + ' |')
open('profiles/r02_sass_mnemonics.md', 'w').write('\n'.join(out) + '\n')
print(len(per), 'kernels')
