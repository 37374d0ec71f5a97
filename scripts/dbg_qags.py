import sys; sys.path.insert(0,'.')
import numpy as np, torch, ctypes
from tests.configs import oracle_model
from oracle import cg_oracle as O
from cluster_generator_b200 import _native as N
om=oracle_model("hernquist"); ov=om.dm_virial
ee=np.ascontiguousarray(ov.ee); t,c,_=ov.f._eval_args; t=np.ascontiguousarray(t); c=np.ascontiguousarray(c[:ee.size])
lib=O.clib(); ip=ctypes.POINTER(ctypes.c_int)
td,cd=N.dev_f64(t),N.dev_f64(c)
for b in (0,6,7,12,100,500,900):
    res=ctypes.c_double(); err=ctypes.c_double(); last=ctypes.c_int(); ier=ctypes.c_int()
    al=np.zeros(64); bl=np.zeros(64); rl=np.zeros(64); el=np.zeros(64)
    lib.cgo_qags(O._dptr(t),t.size,O._dptr(c),ee[b],0.0,ee[b],1.49e-8,1,ctypes.byref(res),ctypes.byref(err),ctypes.byref(last),ctypes.byref(ier),O._dptr(al),O._dptr(bl),O._dptr(rl),O._dptr(el))
    out=torch.zeros(204,dtype=torch.float64,device='cuda')
    N.check(N.lib().cgb_qags_trace(N.ptr(td),N.ptr(cd),ee.size,ee[b],0.0,ee[b],1.49e-8,1.49e-8,1,N.ptr(out),None))
    o=out.cpu().numpy(); L=int(o[2]); Lo=last.value
    print('bin',b,'dev last',L,'ier',o[3],'res',o[0],'err',o[1],'| oracle last',Lo,'ier',ier.value,'res',res.value,'err',err.value)
    m=min(L,Lo)
    print(' alist eq',np.array_equal(o[4:4+m],al[:m]),'rlist eq',np.array_equal(o[104:104+m],rl[:m]),'elist eq',np.array_equal(o[154:154+m],el[:m]))
    d=o[154:154+m]/el[:m]-1
    print(' elist max rel diff', np.abs(d).max(), 'first idx differing lists', next((i for i in range(m) if o[4+i]!=al[i]), None))
