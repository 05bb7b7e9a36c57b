"""Shared by tests/test_oracle.py and tests/test_path_gpu.py: the fixtures under tests/golden/ref_*.npz hold what the
REFERENCE's own drivers computed (tests/golden/make_reference_golden.py ran iterative_inference.py:inference and
iterative_inference_valid.py:inference through oracle/refrun); this module rebuilds each case's seeded inputs and
checkpoints and parses the reference's stdout."""
import json
import os
import re

import numpy as np

from tests.golden import make_reference_golden as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NCLS = G.NCLS
LOOP_CASES = [n for n, c in G.CASES.items() if c['script'] in ('inference', 'valid')]
_NUM = r'([-+0-9.eE]+|nan|inf)'
_ITER_LINE = re.compile(r'^%s %s %s$' % (_NUM, _NUM, _NUM))


def load(name):
    with np.load(os.path.join(GOLD, name + '.npz')) as f:
        fx = {k: f[k] for k in f.files}
    case = json.loads(str(fx['case']))
    assert case == json.loads(json.dumps(G.CASES[name])), 'fixture %s was generated from a different case description' % name
    return fx, case


def dae_kwargs(case):
    d = case['dae']
    return dict(concat_h=tuple(d['concat_h']), additional_pool=d['additional_pool'], unpool_type=d['unpool_type'], bn=bool(d['bn']),
                skip=bool(d['skip']), conv_before_pool=d['conv_before_pool'])


def parse_stdout(text):
    """-> (per_iteration, blocks): per_iteration = [[(rec, acc, jaccard), ...] per image, in order]
    (iterative_inference.py:282 `print rec_iter, acc_iter, np.nanmean(...)`, images separated by the dashed line of :260);
    blocks = [(title, loss, acc, jaccard)] of every helpers.py:172-177 print_results call."""
    per_iter, blocks = [], []
    lines = text.split('\n')
    for i, line in enumerate(lines):
        if line == '-----------------------':
            per_iter.append([])
        m = _ITER_LINE.match(line.strip())
        if m and per_iter:
            per_iter[-1].append(tuple(float(v) for v in m.groups()))
        if line.startswith('>>>>> ') and i + 3 < len(lines) and lines[i + 1].startswith('    Loss: '):
            blocks.append((line[6:].rstrip(':'), float(lines[i + 1].split(': ')[1]), float(lines[i + 2].split(': ')[1]),
                           float(lines[i + 3].split(': ')[1])))
    return per_iter, blocks
