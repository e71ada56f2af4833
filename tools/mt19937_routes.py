import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, "a-watermark-for-diffusion-models_b200")
import gswm
from gswm import codec
dev = torch.device("cuda:0")
key, nonce = bytes.fromhex(gswm.DEFAULT_KEY_HEX), bytes.fromhex(gswm.DEFAULT_NONCE_HEX)
km = codec.KeyMaterial.make(key, nonce, codec.pad_message("lthero", 32), 256)
def med(f, n=100):
    for _ in range(10): f()
    t=[]
    for _ in range(n):
        torch.cuda.synchronize(); t0=time.perf_counter(); f(); torch.cuda.synchronize(); t.append(time.perf_counter()-t0)
    return round(1e6*float(np.median(t)),1)
out={}
for B, shape in ((1,(4,64,64)),(8,(4,64,64)),(1,(4,128,128)),(64,(4,64,64))):
    n=int(np.prod(shape))
    a = med(lambda: codec.embed_batch_mt19937(42, B, shape, km, torch.float32, dev))
    b = med(lambda: codec.embed_batch_injected(codec.mt19937_uniform(42, n, B, dev), shape, km, B, torch.float32))
    g = med(lambda: codec.mt19937_uniform(42, n, B, dev))
    za = codec.embed_batch_mt19937(42, B, shape, km, torch.float32, dev); zb = codec.embed_batch_injected(codec.mt19937_uniform(42, n, B, dev), shape, km, B, torch.float32)
    out[f"B={B},n={n}"]={"fused_us":a,"two_kernels_us":b,"generator_alone_us":g,"identical":bool(torch.equal(za,zb))}
print(json.dumps(out,indent=1))
