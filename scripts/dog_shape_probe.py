import sys, json
sys.path.insert(0, '.')
import numpy as np, torch
from arcadia_microscopy_tools_b200 import _gpu, _lib as L
lib = L.load()
for (h, w) in [(1200, 1920), (2160, 2560), (2048, 2048)]:
    planes = 16
    img = torch.randint(0, 32767, (planes, h, w), dtype=torch.int16, device='cuda')
    res = {}
    for generic in (1, 0):
        lib.amt_tune(b'dog_generic', generic)
        for _ in range(3): _gpu.dog2d(img, 1/65535.0, 0.6, 16.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out, mm = _gpu.dog2d(img, 1/65535.0, 0.6, 16.0)
        e1.record(); torch.cuda.synchronize()
        res['generic' if generic else 'strip'] = e0.elapsed_time(e1) / 5
    res['gpix_s_strip'] = planes*h*w/res['strip']/1e6
    print((h, w), res)
lib.amt_tune(b'dog_generic', 0)
