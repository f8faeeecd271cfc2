"""Does on-demand (managed) memory behave as the sparse capacity growth needs?  One GPU."""
import ctypes as C, time, sys, os
import torch
torch.cuda.init()
x0 = torch.zeros(1, device='cuda')
rt = C.CDLL('libcudart.so.12')
rt.cudaMallocManaged.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaMemAdvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
rt.cudaFree.argtypes = [C.c_void_p]

class Holder:
    def __init__(self, nbytes, shape, typestr):
        p = C.c_void_p()
        rc = rt.cudaMallocManaged(C.byref(p), nbytes, 1)
        if rc != 0:
            raise RuntimeError('cudaMallocManaged(%d) -> %d' % (nbytes, rc))
        self.ptr = p.value
        self.__cuda_array_interface__ = {'shape': shape, 'typestr': typestr, 'data': (self.ptr, False), 'version': 2}

def report(tag):
    free, total = torch.cuda.mem_get_info()
    print('%-40s free %.1f GB of %.1f' % (tag, free / 2**30, total / 2**30), flush=True)

report('start')
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 420
S, F, K, Cn = 64, 8, GB * 2**30 // (64 * 8 * 200 * 8), 200
h = Holder(S * F * K * Cn * 8, (S, F, K, Cn), '<f8')
print('managed alloc of %.1f GB ok, K = %d' % (S * F * K * Cn * 8 / 2**30, K))
rc = rt.cudaMemAdvise(h.ptr, S * F * K * Cn * 8, 3, 0)     # cudaMemAdviseSetPreferredLocation = 3, device 0
print('advise preferred location rc', rc)
t = torch.as_tensor(h, device='cuda')
print('tensor', t.shape, t.dtype, t.device, t.is_contiguous())
report('after wrap')
old = torch.randn(S, F, 128, Cn, dtype=torch.float64, device='cuda')
torch.cuda.synchronize(); t0 = time.time()
t[:, :, :127] = old[:, :, :127]
t[:, :, K - 1] = old[:, :, 127]
torch.cuda.synchronize(); print('sparse copy-in %.3f s' % (time.time() - t0))
report('after sparse copy')
print('far page zero:', float(t[3, 2, K // 2, 5].item()), float(t[63, 7, K - 2].abs().sum().item()), 'copied ok:', bool(torch.equal(t[:, :, :127], old[:, :, :127])))
# one chain grows to the full K in one column (the giant view): touch every slot of chain 5
torch.cuda.synchronize(); t0 = time.time()
t[5, :, :, 7] = 1.0
torch.cuda.synchronize(); print('first touch of one chain, all slots: %.3f s' % (time.time() - t0))
report('after one giant chain')
torch.cuda.synchronize(); t0 = time.time()
t[5, :, :, 7] += 1.0
torch.cuda.synchronize(); t1 = time.time() - t0
ref = torch.zeros(F, K, Cn, dtype=torch.float64, device='cuda')
torch.cuda.synchronize(); t0 = time.time()
ref[:, :, 7] += 1.0
torch.cuda.synchronize(); t2 = time.time() - t0
print('second touch %.4f s vs ordinary device memory %.4f s' % (t1, t2))
s = float(t[5, :, :, 7].sum().item())
print('sum', s, 'expected', 2.0 * F * K)
