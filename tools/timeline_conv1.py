"""Per-tile timeline of the steady-state conv1_1 launch (IISEG_CONV_DBG=4): where does a tile's time go?"""
import ctypes as C, os, sys
os.environ['IISEG_CONV_DBG'] = '4'
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iterative_inference_segm_b200 import _kernels as K, _lib

def run(name, fn):
    fn(); fn(); torch.cuda.synchronize()
    buf = (C.c_longlong * 512)()
    _lib.load().iiseg_debug_read_timeline(buf, 512)
    t = [[buf[i * 16 + s] for s in range(16)] for i in range(32)]
    base = t[2][0]
    print(name)
    print(' tile | mma: wait_tmem_empty  got_it  a_full  issued | epi(group 0): wait_full got_full end')
    for i in range(2, 26):
        print(' %3d  | %s' % (i, ' '.join('%7d' % (t[i][s] - base) if t[i][s] else '      -' for s in range(7))))

B = 10
y = torch.randn(B, 360, 480, 16, device='cuda').to(torch.bfloat16)
W1 = (torch.randn(64, 9 * 16, device='cuda') * 0.02).to(torch.bfloat16)
b1 = torch.zeros(64, device='cuda')
pooled = torch.empty(B, 279, 339, 64, dtype=torch.bfloat16, device='cuda')
mask = torch.empty(B, 279, 339, 8, dtype=torch.int32, device='cuda')
run('conv1_1 steady window', lambda: K.conv2d(y, W1, b1, 3, 3, 100, relu=True, window=(98, 98, 362, 482), pooled=pooled, pool_mask=mask))
up = torch.randn(B, 362, 482, 64, device='cuda').to(torch.bfloat16)
W2 = (torch.randn(16, 9 * 64, device='cuda') * 0.02).to(torch.bfloat16)
b2 = torch.zeros(16, device='cuda')
out = torch.empty(B, 360, 480, 16, dtype=torch.float32, device='cuda')
run('up_conv1', lambda: K.conv2d(up, W2, b2, 3, 3, 1, relu=False, window=(1, 1, 360, 480), out=out, out_f32=True))
v2 = torch.randn(B, 183, 243, 128, device='cuda').to(torch.bfloat16)
W3 = (torch.randn(64, 9 * 128, device='cuda') * 0.02).to(torch.bfloat16)
b3 = torch.zeros(64, device='cuda')
add = torch.randn(B, 279, 339, 64, device='cuda').to(torch.bfloat16)
out3 = torch.empty(B, 181, 241, 64, dtype=torch.bfloat16, device='cuda')
run('up_conv2', lambda: K.conv2d(v2, W3, b3, 3, 3, 1, relu=False, window=(1, 1, 181, 241), out=out3, addend=add, addend_off=(49, 49)))
