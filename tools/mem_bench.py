import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thinktwice_b200.engine import Engine, FMap
e = Engine('cuda:0')
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3
n = 4 * 112 * 224 * 256
x = torch.randn(n, device='cuda'); y = torch.empty_like(x)
print('torch copy 103MB r+w', t(lambda: y.copy_(x)), 'us')
print('torch fill 103MB w  ', t(lambda: y.fill_(1.0)), 'us')
print('tt_fill 103MB w     ', t(lambda: e.fill(y, 1.0)), 'us')
xs = FMap(x.view(4 * 112 * 224, 1, 1, 256), 4 * 112 * 224, 1, 1, 256)
ys = FMap(y.view(4 * 112 * 224, 1, 1, 256), 4 * 112 * 224, 1, 1, 256)
print('tt_copy2d 103MB r+w ', t(lambda: e.copy_cols(xs, ys)), 'us')
z = torch.empty(4, 112, 224, 256, device='cuda')
print('relu torch          ', t(lambda: torch.relu_(z)), 'us')
