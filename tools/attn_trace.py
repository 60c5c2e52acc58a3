"""Phase timeline of the attention kernel's softmax loop (CTA 0), in SM clocks, from hoigen_debug_attention_trace."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi
dev = torch.device("cuda:0")
_cabi.init(dev)
B = 64
qkv = (torch.randn(B * 197, 2304, device=dev) * 1.5).bfloat16()
out = torch.zeros(B * 197, 768, device=dev, dtype=torch.bfloat16)
tr = torch.zeros(16 * 8, device=dev, dtype=torch.int64)
for _ in range(3):
    _cabi.call("hoigen_debug_attention_trace", qkv.data_ptr(), out.data_ptr(), B, tr.data_ptr())
torch.cuda.synchronize()
t = tr.cpu().view(16, 8)
names = ["top", "S ready", "load+max", "bar", "O ready", "epilogue", "exp+P"]
print("item " + " ".join(f"{n:>9}" for n in names[1:]) + "     total")
for i in range(1, 10):
    row = t[i]
    d = [int(row[k] - row[k - 1]) for k in range(1, 7)]
    print(f"{i:4d} " + " ".join(f"{v:9d}" for v in d) + f" {int(t[i][6] - t[i - 1][6]):9d}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    _cabi.call("hoigen_attention", qkv.data_ptr(), out.data_ptr(), B)
e1.record(); torch.cuda.synchronize()
print(f"attention: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch")
