import torch, ctypes
p = torch.cuda.get_device_properties(0)
print(p)
cudart = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int()
for name, attr in (("maxPersistingL2CacheSize", 108), ("maxAccessPolicyWindowSize", 109), ("l2CacheSize", 38)):
    cudart.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value)
