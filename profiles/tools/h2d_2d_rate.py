"""Host->device rate of strided 2-D copies (g2g_upload_block = cudaMemcpy2DAsync) by row width, pinned source."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from genes2genes_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
n_rows, ld = 40000, 20000
host = torch.empty((n_rows, ld), dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
st = torch.cuda.current_stream()
for width in (ld, 10000, 2500, 1372, 343):
    dst = torch.empty((n_rows, width), dtype=torch.float32, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(lib.g2g_upload_block(ctypes.c_void_p(host.data_ptr()), ld, 0, n_rows, 0, width, ctypes.c_void_p(dst.data_ptr()), width,
                                        ctypes.c_void_p(st.cuda_stream)), "upload")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("width %6d floats (%7d B rows): %.1f GB/s" % (width, width * 4, n_rows * width * 4 / dt / 1e9))
flat = torch.empty(n_rows * 2500, dtype=torch.float32, device=dev)
src = host.view(-1)[:n_rows * 2500]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); flat.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("contiguous 400 MB: %.1f GB/s" % (flat.numel() * 4 / dt / 1e9))
