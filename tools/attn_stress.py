"""Stress: the attention kernel must be bit-reproducible run to run (any race in the mbarrier hand-offs would show up as
differing outputs or a trapped wait) for several token counts, including ragged and odd-tile-count ones."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsformerplusplus_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
bad = 0
for N in (64, 128, 129, 255, 257, 384, 1000, 4097, 27648):
    torch.manual_seed(N)
    qkv = torch.randn(N, 192, device=dev)
    ws = torch.empty((N + 128) * 224 + 16, device=dev)
    ref = None
    reps = 300 if N < 5000 else 60
    for i in range(reps):
        out = torch.empty(N, 64, device=dev)
        _lib.check(L.mvsf_attention_forward(qkv.data_ptr(), out.data_ptr(), ws.data_ptr(), ctypes.c_size_t(ws.numel() * 4), N,
                                            ctypes.c_float(0.31), None), "attn")
        if ref is None:
            ref = out.clone()
            q, k, v = [qkv[:, j * 64:(j + 1) * 64].reshape(N, 4, 16).permute(1, 0, 2).double() for j in range(3)]
            want = torch.softmax(q @ k.transpose(1, 2) * 0.31, -1) @ v
            err = float((out.double() - want.permute(1, 0, 2).reshape(N, 64)).abs().max())
        elif not torch.equal(out, ref):
            bad += 1
    torch.cuda.synchronize()
    print(f"N={N:6d}: {reps} runs, max err vs fp64 {err:.2e}, mismatching runs so far {bad}", flush=True)
print("OK" if bad == 0 else "NOT REPRODUCIBLE")
