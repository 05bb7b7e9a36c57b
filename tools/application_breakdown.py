"""Per-launch CUDA-event times of one steady-state DAE application: `python tools/application_breakdown.py PADDING NB_H`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import synthetic as weights, _kernels as K
from iterative_inference_segm_b200.models.DAE_h import buildDAE
from iterative_inference_segm_b200.profiling import KernelTimer

pad = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nbh = int(sys.argv[2]) if len(sys.argv) > 2 else 512
prec = sys.argv[3] if len(sys.argv) > 3 else 'bf16'
B, H, W = 10, 360, 480
dae = buildDAE([None], None, 11, nb_features_to_concat=nbh, padding=pad, concat_h=['pool4'], noise=0.0, n_filters=64,
               additional_pool=2, skip=True, unpool_type='trackind', params=weights.synthetic_dae_params(11, nbh, seed=1, out_gain=0.1), precision=prec)
net = dae.net
hs = net.h_spatial(H, W)
h = K.pack_nchw(torch.relu(torch.randn(B, nbh, hs[0], hs[1], device='cuda')), net.h_pad, split=net.split)
yf = torch.softmax(torch.randn(B, 11, H, W, device='cuda'), 1)
y = K.pack_nchw(yf, net.y_cpad, split=net.split)
upd = dict(y=yf, active=torch.ones(B, dtype=torch.int32, device='cuda'), norm_acc=torch.zeros(B, dtype=torch.int64, device='cuda'), step=0.05) if not net.split else None
net.logits(h, y, full_down=True, update=upd)
for _ in range(2):
    net.logits(h, y, full_down=False, update=upd)
timer = KernelTimer()
with timer.recording():
    for _ in range(3):
        net.logits(h, y, full_down=False, update=upd)
torch.cuda.synchronize()
tot = 0.0
seen = {}
for name, tag, s, e in timer.records:
    seen.setdefault((name, tag), []).append(s.elapsed_time(e))
for (name, tag), v in seen.items():
    ms = sum(v) / 3
    tot += ms
    print('%-10s %.4f ms  %s' % (name, ms, tag))
print('padding %d: %.3f ms per application (sum of event-timed launches)' % (pad, tot))
