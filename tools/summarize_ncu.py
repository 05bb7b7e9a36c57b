"""Summarises an `ncu --page raw --csv` export into profiles/<name>.md (+ profiles/traffic.json).
    python tools/summarize_ncu.py gpurun_out/prof_steady_full_raw.csv profiles/r01_ncu_steady_application_v16 [tag]
`tag` (e.g. the precision variant): traffic.json keeps its other entries and gains `<kernel>/<tag>` ones, plus
`conv_igemm_pair_kernel/<tag>` = the mean over every CTA-pair launch (the dominant kernel of bench.py's roofline)."""
import csv
import json
import sys

COLS = [('gpu__time_duration.sum', 'us'), ('sm__cycles_elapsed.avg', 'SM cycles'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('dram__bytes_read.sum', 'DRAM read MB'), ('dram__bytes_write.sum', 'DRAM write MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
        ('lts__t_bytes.sum', 'L2 bytes'), ('launch__grid_size', 'grid'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem')]


def main(src, dst, tag=None):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(hdr.index(c), c, label) for c, label in COLS if c in hdr]
    kn = hdr.index('Kernel Name')
    lines = ['| # | kernel | ' + ' | '.join('%s [%s]' % (label, units[i]) for i, c, label in idx) + ' |',
             '|---|---|' + '---|' * len(idx)]
    traffic = {}
    for n, r in enumerate(data):
        name = r[kn].replace('void ', '').split('(')[0].replace('iiseg::', '')
        lines.append('| %d | %s | ' % (n, name) + ' | '.join(r[i] for i, c, label in idx) + ' |')
        mb = float(r[hdr.index('dram__bytes_read.sum')]) + float(r[hdr.index('dram__bytes_write.sum')])
        scale = {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Gbyte': 1e9}[units[hdr.index('dram__bytes_read.sum')]]
        traffic.setdefault(name, []).append(mb * scale)
    with open(dst + '.md', 'w') as fh:
        fh.write('ncu --set full --clock-control none, one steady-state DAE application (batch 10, 360x480), launches in order.\n'
                 'Per-launch times are cold-cache and serialised (ncu replays each kernel); shares, not absolutes, compare with bench.py.\n\n')
        fh.write('\n'.join(lines) + '\n')
    out = {k: {'dram_bytes_per_launch': sum(v) / len(v), 'launches': len(v)} for k, v in traffic.items()}
    if tag:
        import os
        old = json.load(open('profiles/traffic.json')) if os.path.exists('profiles/traffic.json') else {}
        pair = [b for k, v in traffic.items() if k.startswith('conv_igemm_pair_kernel') for b in v]
        old.update({'%s/%s' % (k, tag): v for k, v in out.items()})
        if pair:
            old['conv_igemm_pair_kernel/' + tag] = {'dram_bytes_per_launch': sum(pair) / len(pair), 'launches': len(pair)}
        out = old
    with open('profiles/traffic.json', 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('\n'.join(lines))


if __name__ == '__main__':
    main(*sys.argv[1:4])
