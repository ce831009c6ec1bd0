import sys, time; sys.path.insert(0,'/root/repo')
import torch, ctypes as C
from puffer_phc_b200 import _ffi, synth
from puffer_phc_b200.envs import common
lib=_ffi.load()
dev='cuda:0'
N=4096
bs=torch.randn(N,24,13,device=dev); ref=torch.randn(N,24,3,device=dev)
prog=torch.zeros(N,dtype=torch.int16,device=dev); pt=torch.zeros(N,dtype=torch.bool,device=dev); td=torch.full((24,),0.25,device=dev); rb=torch.ones(N,dtype=torch.bool,device=dev)
def t(f,n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); a=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-a)/n*1e6
pos=bs[...,0:3]
print('full reset call', t(lambda: common.compute_humanoid_im_reset(rb,prog,None,None,pos,ref,pt,True,td,False)))
print('torch.empty x2', t(lambda: (torch.empty(N,dtype=torch.bool,device=dev), torch.empty(N,dtype=torch.bool,device=dev))))
print('stream_ptr', t(lambda: _ffi.stream_ptr()))
print('view3 x2', t(lambda: (_ffi.view3(pos), _ffi.view3(ref))))
print('as_view_tensor x2 + require', t(lambda: common._prep(pos, ref)))
print('ptr x5', t(lambda: [_ffi.ptr(x) for x in (prog,pt,td,rb,rb)]))
print('.to/.contiguous x3', t(lambda: (prog.to(torch.int16).contiguous(), pt.to(torch.bool).contiguous(), td.to(torch.float32).reshape(-1).contiguous())))
print('on_device ctx', t(lambda: _ffi.on_device(pos.device).__enter__()))
r1=torch.empty(N,dtype=torch.bool,device=dev); r2=torch.empty(N,dtype=torch.bool,device=dev)
v1,v2=_ffi.view3(pos),_ffi.view3(ref); sp=_ffi.stream_ptr()
print('raw ctypes call', t(lambda: lib.phc_im_reset(_ffi.ptr(prog), v1, v2, _ffi.ptr(pt), 1, _ffi.ptr(td), 0, N, 24, _ffi.ptr(r1), _ffi.ptr(r2), 1, sp)))
a1=(prog.data_ptr(), v1, v2, pt.data_ptr(), 1, td.data_ptr(), 0, N, 24, r1.data_ptr(), r2.data_ptr(), 1, sp)
print('raw ctypes call prebuilt args', t(lambda: lib.phc_im_reset(*a1)))
print('empty launch torch op (add_)', t(lambda: r1.logical_not_()))
