import sys, time, numpy as np, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from gomel_b200 import _lib
from util import synth_clip
ctx=_lib.Context(0)
cfg0=_lib.make_config(gl_iters=0); ctx.set_mel_tables(cfg0,0.0,16000.0)
frames=342; ola=4096+341*1280
mel=np.random.default_rng(0).uniform(-9,2,(frames*192,2))
init=np.random.default_rng(1).random(ola)
def t(f,n=20):
    for _ in range(3): f()
    t0=time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter()-t0)/n*1e3
for it in (0,1,2,4,8,32):
    c=_lib.make_config(gl_iters=it)
    print("iters",it, "ms", round(t(lambda: ctx.from_mel(c,mel,init=init)),4), "kernels", ctx.last_lead_kernel_ms(), ctx.last_hot_kernel_ms() if it>4 else None)
# raw copies
d=ctx.dev_malloc(ola*8)
out=np.empty(ola)
print("H2D 3.5MB pageable ms", round(t(lambda: ctx.h2d(d,init)),4))
print("D2H 3.5MB pageable ms", round(t(lambda: ctx.d2h(out,d)),4))
pin,own=ctx.pinned_array((ola,),np.float64); pin[:]=init
print("H2D 3.5MB pinned ms", round(t(lambda: ctx.h2d(d,pin)),4))
print("D2H 3.5MB pinned ms", round(t(lambda: ctx.d2h(pin,d)),4))
print("np.empty+ascontig overhead ms", round(t(lambda: (np.ascontiguousarray(mel,np.float64), np.empty(ola))),4))
print("host memcpy 3.5MB ms", round(t(lambda: np.copyto(out,init)),4))
