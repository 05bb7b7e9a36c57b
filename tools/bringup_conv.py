"""GPU bring-up for the tcgen05 conv kernel: runs each case in its own process (a
trapped kernel poisons the CUDA context) with a timeout, compares against torch's
fp32 conv on the same bf16-rounded operands, and prints where mismatches sit.

    python tools/bringup_conv.py            # all cases
    python tools/bringup_conv.py --case 3   # one case, in-process
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name, N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32
CASES = [
    ('1x1 one tile one kblock', 1, 8, 16, 64, 0, 64, 1, 0, 0, 0, None, 0),
    ('1x1 K=128', 1, 8, 16, 128, 0, 64, 1, 0, 0, 0, None, 0),
    ('1x1 Cout=128', 1, 8, 16, 64, 0, 128, 1, 0, 0, 0, None, 0),
    ('1x1 Cout=256', 1, 8, 16, 64, 0, 256, 1, 0, 0, 0, None, 0),
    ('1x1 Cout=16 f32', 1, 8, 16, 64, 0, 16, 1, 0, 0, 0, None, 1),
    ('3x3 pad1 one tile', 1, 8, 16, 64, 0, 64, 3, 1, 1, 0, None, 0),
    ('3x3 pad1 ragged 37x45', 2, 37, 45, 64, 0, 128, 3, 1, 1, 0, None, 0),
    ('3x3 pad1 17x21 Cout=512 K=9x256', 2, 17, 21, 256, 0, 512, 3, 1, 1, 0, None, 0),
    ('3x3 dual source', 2, 17, 21, 128, 64, 256, 3, 1, 1, 0, None, 0),
    ('3x3 addend linear', 2, 33, 29, 128, 0, 64, 3, 1, 0, 1, None, 0),
    ('3x3 pad100', 1, 24, 32, 64, 0, 64, 3, 100, 1, 0, None, 0),
    ('3x3 window f32 Cout16', 1, 40, 56, 64, 0, 16, 3, 1, 0, 0, (5, 7, 24, 32), 1),
    ('7x7 valid', 1, 17, 21, 128, 0, 256, 7, 0, 1, 0, None, 0),
    ('3x3 many tiles persistent', 4, 139, 169, 64, 0, 128, 3, 1, 1, 0, None, 0),
]


def run_case(idx):
    import torch
    import torch.nn.functional as F
    from iterative_inference_segm_b200 import _kernels as K, _lib
    name, N, H, W, C0, C1, Cout, R, pad, relu, addend, window, out_f32 = CASES[idx]
    torch.manual_seed(idx)
    dev = 'cuda'
    Cin = C0 + C1
    x = torch.randn(N, Cin, H, W, device=dev).to(torch.bfloat16)
    Wt = (torch.randn(Cout, Cin, R, R, device=dev) / (Cin * R * R) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    src0 = x_nhwc[..., :C0].contiguous()
    src1 = x_nhwc[..., C0:].contiguous() if C1 else None
    Wk = Wt.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()   # [Cout][R*S][Cin]
    fOH, fOW = H + 2 * pad - R + 1, W + 2 * pad - R + 1
    oh0, ow0, OH, OW = window if window else (0, 0, fOH, fOW)
    add = None
    if addend:
        add = torch.randn(N, OH, OW, Cout, device=dev).to(torch.bfloat16)
    out = K.conv2d(src0, Wk, b, R, R, pad, relu, src1=src1, addend=add, window=window, out_f32=bool(out_f32))
    try:
        torch.cuda.synchronize()
    except Exception as e:   # noqa
        print('  SYNC FAILED:', str(e).splitlines()[0], 'diag', _lib.read_diag())
        return 2
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = F.conv2d(x.float(), Wt.float(), b, padding=pad)[:, :, oh0:oh0 + OH, ow0:ow0 + OW]
    if add is not None:
        ref = ref + add.float().permute(0, 3, 1, 2)
    if relu:
        ref = torch.relu(ref)
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 2e-2 * ref.abs().clamp(min=1.0) if not out_f32 else 1e-3 * ref.abs().clamp(min=1.0)
    bad = err > tol
    nbad = int(bad.sum())
    print('  max err %.4g  mean err %.4g  ref absmax %.3g  bad %d / %d' % (
        float(err.max()), float(err.mean()), float(ref.abs().max()), nbad, bad.numel()))
    if nbad:
        idxs = bad.nonzero()[:8].tolist()
        for (n, c, h, w) in idxs:
            print('    n=%d c=%d h=%d w=%d got %.4f ref %.4f' % (n, c, h, w, float(got[n, c, h, w]), float(ref[n, c, h, w])))
        print('    bad by channel%%64 (first 16 nonzero):', [(i, int(v)) for i, v in enumerate(
            bad.sum((0, 2, 3)).view(-1, min(64, Cout)).sum(0).tolist()) if v][:16])
        print('    bad by row h (first 16 nonzero):', [(i, int(v)) for i, v in enumerate(bad.sum((0, 1, 3)).tolist()) if v][:16])
        print('    bad by col w (first 16 nonzero):', [(i, int(v)) for i, v in enumerate(bad.sum((0, 1, 2)).tolist()) if v][:16])
        return 1
    return 0


def main():
    if '--case' in sys.argv:
        idx = int(sys.argv[sys.argv.index('--case') + 1])
        print('[case %d] %s' % (idx, CASES[idx][0]), flush=True)
        rc = run_case(idx)
        print('  ->', 'PASS' if rc == 0 else 'FAIL', flush=True)
        sys.exit(rc)
    results = []
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', str(i)], timeout=120,
                               capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode not in (0, 1, 2):
                sys.stdout.write('  stderr tail: ' + r.stderr[-600:] + '\n')
            results.append(r.returncode)
        except subprocess.TimeoutExpired:
            print('[case %d] %s\n  -> TIMEOUT' % (i, CASES[i][0]))
            results.append(-9)
        sys.stdout.flush()
    print('SUMMARY', results)
    sys.exit(0 if all(r == 0 for r in results) else 1)


if __name__ == '__main__':
    main()
