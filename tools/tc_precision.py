"""CPU emulation of the tensor-core operand splits of usip_layer_fwd (precision 1): error of 3xTF32 and of
TF32 + 2 BF16 cross terms against float64, next to a plain fp32 sgemm, on post-ReLU activations x N(0,1/K) weights."""
import numpy as np
rng=np.random.default_rng(0)
def rna_tf32(x):
    b=x.astype(np.float32).view(np.uint32)
    b=((b.astype(np.uint64)+0x1000)&0xffffe000).astype(np.uint32)
    return b.view(np.float32)
def rn_bf16(x):
    b=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r=((b+0x7fff+((b>>16)&1))&0xffff0000).astype(np.uint32)
    return r.view(np.float32)
for K in (64,256,512):
    a=np.maximum(rng.normal(size=(2048,K)).astype(np.float32)*1.3+0.2,0)
    w=(rng.normal(size=(K,256))/np.sqrt(K)).astype(np.float32)
    ref=a.astype(np.float64)@w.astype(np.float64)
    f32=(a@w)
    ah=rna_tf32(a); al=rna_tf32(a-ah); wh=rna_tf32(w); wl=rna_tf32(w-wh)
    x3=(ah.astype(np.float64)@wh+al.astype(np.float64)@wh+ah.astype(np.float64)@wl)
    alb=rn_bf16(a-ah); ahb=rn_bf16(ah); whb=rn_bf16(wh); wlb=rn_bf16(w-wh)
    xb=(ah.astype(np.float64)@wh+alb.astype(np.float64)@whb+ahb.astype(np.float64)@wlb)
    x1=ah.astype(np.float64)@wh
    s=np.abs(ref).max()
    print(K,"fp32 sgemm err",np.abs(f32-ref).max()/s,"3xtf32",np.abs(x3-ref).max()/s,"tf32+2bf16",np.abs(xb-ref).max()/s,"1xtf32",np.abs(x1-ref).max()/s,
          "rms:",np.sqrt(((x3-ref)**2).mean())/s, np.sqrt(((xb-ref)**2).mean())/s)
