"""Debug: per-tile phase timestamps of the attention kernel (library built with -DMVSF_FA_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsformerplusplus_b200 import _lib
L = _lib.lib()
N = 27648
dev = torch.device("cuda:0")
qkv = torch.randn(N, 192, device=dev)
out = torch.empty(N, 64, device=dev)
ws = torch.empty((N + 128) * 192 + 16, device=dev)
for _ in range(2):
    _lib.check(L.mvsf_attention_forward(qkv.data_ptr(), out.data_ptr(), ws.data_ptr(), ctypes.c_size_t(ws.numel() * 4), N, ctypes.c_float(0.3), None), "attn")
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
L.mvsf_debug_fa_trace.argtypes = [ctypes.c_void_p]
L.mvsf_debug_fa_trace(buf)
t = [[buf[j * 8 + k] for k in range(8)] for j in range(64)]
names = ["loop top", "S ready", "pre sync1", "post sync1", "pre fold", "post fold", "pre sync2", "post sync2"]
print("per-tile deltas (clk), tiles 110..125:")
for j in range(10, 26):
    row = t[j]
    d = [row[k + 1] - row[k] for k in range(7)] + [t[j + 1][0] - row[7]]
    print(j + 100, " ".join(f"{x:6d}" for x in d), "| total", t[j + 1][0] - row[0])
import statistics
tot = [t[j + 1][0] - t[j][0] for j in range(5, 60)]
print("median clk per tile:", statistics.median(tot))
for k in range(8):
    ds = [(t[j][k + 1] - t[j][k]) if k < 7 else (t[j + 1][0] - t[j][7]) for j in range(5, 60)]
    print(f"{names[k]:>12} -> next: median {statistics.median(ds):8.0f}")
