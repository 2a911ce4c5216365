import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from oracle import oracle as O
from p64_b200 import y4m, _lib
from p64_b200.encoder import DeviceContext
it=y4m.IT_QCIF; w,h=y4m.DIMS[it]
rng=np.random.default_rng(7)
ref=rng.integers(0,256,(h,w)).astype(np.uint8); cur=rng.integers(0,256,(h,w)).astype(np.uint8)
ctx=DeviceContext(it,1)
r=torch.from_numpy(ref).cuda(); c=torch.from_numpy(cur).cuda()
out=torch.zeros(99*8,dtype=torch.int32,device='cuda'); surf=torch.zeros(99*961,dtype=torch.int32,device='cuda')
torch.cuda.synchronize()
_lib.check(ctx.L.p64b_ctx_sad_surface_dev(ctx.h,C.c_void_p(r.data_ptr()),C.c_void_p(c.data_ptr()),1,C.c_void_p(out.data_ptr()),C.c_void_p(surf.data_ptr())))
torch.cuda.synchronize()
sf=surf.cpu().numpy().reshape(99,31,31)
for mb in (0,1,50,98):
    want=O.sad_surface(ref,cur,mb%11,mb//11)
    bad=np.argwhere(sf[mb]!=want)
    print(mb,len(bad),bad[:12].tolist())
    if len(bad):
        y,x=bad[0]; print(' got',sf[mb][y,x],'want',want[y,x])
        print(' bad dx set', sorted(set(bad[:,1].tolist())), 'bad dy set', sorted(set(bad[:,0].tolist())))
