import torch, ctypes
dev=torch.device('cuda:0')
x=torch.empty((16,128,800,800),dtype=torch.float32,device=dev)
def t(fn,n=20):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b)/n
ms=t(lambda: x.zero_())
print('zero_ (fill kernel) %.3f ms  %.0f GB/s'%(ms, x.numel()*4/ms/1e6))
rt=ctypes.CDLL('libcudart.so')
rt.cudaMemsetAsync.argtypes=[ctypes.c_void_p,ctypes.c_int,ctypes.c_size_t,ctypes.c_void_p]
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ms=t(lambda: rt.cudaMemsetAsync(ctypes.c_void_p(x.data_ptr()),0,x.numel()*4,st))
print('cudaMemsetAsync %.3f ms  %.0f GB/s'%(ms, x.numel()*4/ms/1e6))
y=torch.empty_like(x[:8]); z=torch.empty_like(x[:8])
ms=t(lambda: y.copy_(z))
print('copy 2.6GB->2.6GB %.3f ms  %.0f GB/s (r+w)'%(ms, 2*y.numel()*4/ms/1e6))
