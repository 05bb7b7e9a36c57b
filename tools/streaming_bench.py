"""Achieved HBM bandwidth of the stand-alone memory-bound kernels at config-2 sizes (batch 10, 360x480 -> the DAE's
padded level sizes), CUDA-event timed over inputs larger than L2, algorithmic bytes / time vs MEASURED_PEAKS.json."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import _kernels as K

dev = torch.device('cuda:0')
peak = 6550.7
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk)).get('hbm_gbs', peak)

def ev(fn, n=10):
    """Device time per launch: n launches captured in a CUDA graph (the ctypes call costs the host 20-40 us, more
    than some of these kernels run), replayed three times, last replay timed."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

rows = []
def report(name, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append((name, nbytes / 1e6, ms * 1e3, gbs, gbs / peak))
    print('%-58s %8.1f MB %8.1f us %7.0f GB/s  %.2f of %.0f' % (name, nbytes / 1e6, ms * 1e3, gbs, gbs / peak, peak))

B, C = 10, 11
torch.manual_seed(0)
# pool + tie mask / unpool at the two largest DAE levels (558x678x64 and 279x339x128, bf16 NHWC)
for (H, W, Cc) in [(558, 678, 64), (279, 339, 128)]:
    x = torch.randn(B, H, W, Cc, device=dev).to(torch.bfloat16)
    pooled, mask = K.maxpool2(x, True)
    nb = x.numel() * 2 + pooled.numel() * 2 + mask.numel() * 4
    report('maxpool2_mask %dx%dx%d' % (H, W, Cc), nb, ev(lambda: K.maxpool2(x, True, pooled, mask)))
    out = torch.empty_like(x)
    nb = pooled.numel() * 2 + mask.numel() * 4 + out.numel() * 2
    report('unpool2_mask  %dx%dx%d' % (H, W, Cc), nb, ev(lambda: K.unpool2(pooled, mask, H, W, out=out)))
# stand-alone softmax + update (the unfused loop): logits fp32 NHWC16 + y fp32 NCHW in, y + y_bf16 out
H, W = 360, 480
logits = torch.randn(B, H, W, 16, device=dev)
y = torch.softmax(torch.randn(B, C, H, W, device=dev), 1)
y_bf16 = torch.empty(B, H, W, 16, dtype=torch.bfloat16, device=dev)
active = torch.ones(B, dtype=torch.int32, device=dev)
lib = K._lib.load()
part = torch.zeros(B, lib.iiseg_update_blocks(H, W), dtype=torch.float32, device=dev)
nb = logits.numel() * 4 + 2 * y.numel() * 4 + y_bf16.numel() * 2
report('softmax_update 10x11x360x480', nb, ev(lambda: K.softmax_update(logits, y, y_bf16, active, part, 0.05)))
# metrics: y fp32 + int32 labels in
labels = torch.randint(0, C + 1, (B, H, W), device=dev, dtype=torch.int32)
cm = torch.zeros(B, C * C, dtype=torch.int64, device=dev); counts = torch.zeros(B, 2, dtype=torch.int64, device=dev)
sq = torch.zeros(B, 2, dtype=torch.float64, device=dev)
nb = y.numel() * 4 + labels.numel() * 4
report('metrics_accumulate 10x11x360x480', nb, ev(lambda: K.metrics_accumulate(y, cm, counts, sq, labels=labels, void_label=C)))
onehot = torch.zeros(B, C + 1, H, W, device=dev)
nb = onehot.numel() * 4 + labels.numel() * 4
report('onehot_to_labels 10x12x360x480', nb, ev(lambda: K.onehot_to_labels(onehot, labels)))
nb = y.numel() * 4 + y_bf16.numel() * 2
report('pack_nchw (y fp32 NCHW -> bf16 NHWC16)', nb, ev(lambda: K.pack_nchw(y, 16, out=y_bf16)))
json.dump([dict(kernel=r[0], mbytes=r[1], us=r[2], gbs=r[3], frac=r[4]) for r in rows], open('gpurun_out/streaming_bench.json', 'w'), indent=1)
