import sys
sys.path.insert(0, '.')
from __graft_entry__ import load_package
vb = load_package(); L = vb.lib.lib()
print('DFMA', L.vela_b200_fp64_peak_tflops(0, 0), 'DMMA full', L.vela_b200_fp64_peak_tflops(0, 1))
for w in (4, 8):
    print('DMMA with', w, 'warps/SM:', L.vela_b200_fp64_peak_tflops(0, 100 + w))
for mix, name in ((0, 'pure (8 warps/SM)'), (1, '+1 DMUL per 16 DMMA'), (2, '+4 DMUL per 16 DMMA'), (3, '+8 LDS +4 DMUL per 16 DMMA')):
    print('DMMA mix', name, L.vela_b200_fp64_peak_tflops(0, 200 + mix))
