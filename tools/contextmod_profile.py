"""One steady-state application of the context-module DAE (batch 10, 360x480): the launch list for `ncu`
(profiles/r02_ncu_contextmod.md) and the event-timed application.
    python tools/contextmod_profile.py            # event timing
    ncu --set full -k regex:ctx_conv --launch-skip 8 ... python tools/contextmod_profile.py once"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as S                      # noqa: E402
from iterative_inference_segm_b200.models.contextmod_dae import ContextModNet  # noqa: E402

B, H, W, C = 10, 360, 480, 11
net = ContextModNet(C, 3, S.synthetic_contextmod_params(C, 3, seed=3))
X = torch.rand((B, 3, H, W), device='cuda')
y = torch.softmax(torch.randn((B, C, H, W), device='cuda'), 1)
net.logits(X, None, full_down=True, y_f32=y)
torch.cuda.synchronize()
once = len(sys.argv) > 1 and sys.argv[1] == 'once'
net.logits(X, None, full_down=False, y_f32=y)
torch.cuda.synchronize()
if not once:
    n = 20
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        net.logits(X, None, full_down=False, y_f32=y)
    e.record()
    torch.cuda.synchronize()
    print('%.3f ms per application' % (s.elapsed_time(e) / n))
