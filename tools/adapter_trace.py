"""Phase timeline (SM clocks) of CTA 0 of the adapter block kernel, via hoigen_debug_adapter_trace."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi, synthetic as S
from hoigen_b200.encoder import VisionTransformer
from oracle import hoi_forward_ref as O

dev = torch.device("cuda:0")
B = 64
enc = S.make_encoder_state(0)
vt = VisionTransformer()
vt.load_state_dict({k[len(O.ENC):]: v for k, v in enc.items()}, strict=False)
vt = vt.to(dev).eval()
head = S.make_head_state(117, 256)
props = S.make_region_props(B)
prior, mask = O.prior_tokens(props, (224, 224), head.tensors, head.attrs["object_embedding"])
imgs = S.make_images(B, seed=1).to(dev)
tr = torch.zeros(16, dtype=torch.int64, device=dev)
lib = _cabi.load()
vt(imgs, (prior.to(dev), mask.to(dev)))
lib.hoigen_debug_adapter_trace(tr.data_ptr())
vt(imgs, (prior.to(dev), mask.to(dev)))
torch.cuda.synchronize()
lib.hoigen_debug_adapter_trace(None)
t = tr.cpu().tolist()
names = ["setup", "phase0 (down-proj ring)", "D epilogue + KV stage", "MMA1 q", "cross-attention", "MMA2 out-proj", "LN2",
         "MMA3 linear1", "relu + store hidden", "MMA4 linear2", "LN3 + A0 store", "up chunk0", "up chunk1", "up chunk2", "-", "tail"]
for k in range(1, 16):
    if k == 14:
        continue
    prev = t[k - 1] if k != 15 else t[13]
    print(f"{names[k]:28s} {t[k] - prev:7d} clk")
print(f"total (stamp 0 -> 15)        {t[15] - t[0]:7d} clk = {(t[15] - t[0]) / 1.965e3:.1f} us at 1965 MHz")
