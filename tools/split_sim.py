import numpy as np
rs=np.random.RandomState(0)
def bf16(x):
    u=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r=((u+0x7fff+((u>>16)&1))>>16)<<16
    return r.astype(np.uint32).view(np.float32)
def tf32(x):
    u=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r=((u+0xfff+((u>>13)&1))>>13)<<13
    return r.astype(np.uint32).view(np.float32)
def split2(x,f):
    hi=f(x); lo=f(x-hi); return hi,lo
def mm(a,b,mode):
    a=a.astype(np.float32); b=b.astype(np.float32)
    if mode=='f32': return a@b
    if mode=='f64': return (a.astype(np.float64)@b.astype(np.float64))
    if mode=='bf16': return bf16(a)@bf16(b)
    if mode=='tf32': return tf32(a)@tf32(b)
    f=bf16 if mode.startswith('bf16') else tf32
    ah,al=split2(a,f); bh,bl=split2(b,f)
    if mode.endswith('x3'): return ah@bh + (ah@bl + al@bh)
    if mode.endswith('x4'): return ah@bh + (ah@bl + al@bh) + al@bl
def sig(x): return 1/(1+np.exp(-x))
M,H,T=256,256,64
def glorot(s): 
    l=np.sqrt(6/(s[0]+s[1])); return rs.uniform(-l,l,s).astype(np.float32)
for stress in (1.0,2.0):
    Wg=glorot((384,512))[128:]*stress; Wc=glorot((384,256))[128:]*stress
    Xg=(rs.normal(0,0.6*stress,(T,M,512))+1).astype(np.float32); Xc=rs.normal(0,0.6*stress,(T,M,256)).astype(np.float32)
    s=rs.uniform(0,1,(T,M,1)).astype(np.float32)
    def run(mode,dt=np.float32):
        h=np.zeros((M,H),dt)
        for t in range(T):
            g=sig(mm(h,Wg,mode).astype(dt)+Xg[t]); r,u=g[:,:H],g[:,H:]
            c=np.tanh(mm(r*h,Wc,mode).astype(dt)+Xc[t])
            u=(1-s[t])*u; h=(u*h+(1-u)*c).astype(dt)
        return h
    ref=run('f64',np.float64)
    rms=np.sqrt((ref**2).mean())
    for mode in ('f32','bf16x3','bf16x4','tf32x3','tf32','bf16'):
        h=run(mode); e=np.abs(h-ref)
        print('stress',stress,mode,'max abs err %.3g  max err/rms %.3g  mean err/rms %.3g'%(e.max(), e.max()/rms, e.mean()/rms))
