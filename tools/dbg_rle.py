import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PRODUCT_SO, QB3Lib, golden_cases, golden_image, oracle
P = QB3Lib(PRODUCT_SO, 256)
def ref_derle(p):
    out=bytearray(); i=0; n=len(p)
    while i<n:
        b=p[i]; i+=1
        if b==0xff and i+1<n and p[i]==0xff:
            c=p[i+1]; i+=2
            out += b'\xff\xff' if c==0xff else bytes(4+c)
        else: out.append(b)
    return bytes(out)
for name in ('c_uint16_fewvals_7', 'c_int16_fewvals_7', 'c_uint8_fewvals_7'):
    case=[c for c in golden_cases() if c['name']==name][0]
    s=bytes.fromhex(case['stream']); img=golden_image(case)
    hdr=oracle().info(s)['data_offset']
    x=bytearray(s[:hdr]+ref_derle(s[hdr:])); x[10]=s[10]-2
    want=oracle().decode(bytes(x))
    print(name, 'mode byte', s[10], 'oracle on expanded ok', want is not None and np.array_equal(want, img))
    for label, st in (('rle stream', s), ('expanded', bytes(x))):
        d=P.decode(st)
        print('  device', label, None if d is None else np.array_equal(d, img), None if d is None else np.argwhere(d!=img)[:4].tolist())
