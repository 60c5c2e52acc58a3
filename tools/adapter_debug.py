import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hoigen_b200 import _cabi, synthetic as S
from oracle import hoi_forward_ref as O
dev = torch.device("cuda:0")
_cabi.init(dev)
sd = S.make_encoder_state(0)
B, layer = 4, 3
n_list = [int(v) for v in os.environ.get('NLIST', '16,9,1,13').split(',')]
g = torch.Generator().manual_seed(5)
n_max = max(n_list)
prior = torch.randn(B, n_max, 64, generator=g)
mask = torch.ones(B, n_max, dtype=torch.bool)
for b, n in enumerate(n_list):
    mask[b, :n] = False
blk = f"{O.ENC}transformer.resblocks.{layer}.adaptermlp.mhsa_layers.0."
ad = blk.replace("mhsa_layers.0.", "")
names = ["multihead_attn.in_proj_weight", "multihead_attn.in_proj_bias", "multihead_attn.out_proj.weight",
         "multihead_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
         "norm2.weight", "norm2.bias", "norm3.weight", "norm3.bias"]
w = [sd[blk + n].to(dev).contiguous() for n in names]
kv = torch.empty(1, B * n_max, 128, device=dev)
pr = prior.to(dev).contiguous()
_cabi.call("hoigen_adapter_kv", pr.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), kv.data_ptr(), B * n_max, 1)
wd, bd = sd[ad + "down_proj.weight"], sd[ad + "down_proj.bias"]
wu = sd[ad + "up_proj.weight"]
keep = [wd.bfloat16().contiguous().to(dev), bd.contiguous().to(dev), w[0][:64].bfloat16().contiguous(), w[2].bfloat16().contiguous(),
        w[4].bfloat16().contiguous(), w[6].bfloat16().contiguous(), wu.bfloat16().contiguous().to(dev)]
aw = _cabi.AdapterWeights()
for f, t in zip(("wd", "down_b", "wq", "wo", "w1", "w2", "in_proj_b", "out_proj_b", "linear1_b", "linear2_b", "norm2_w", "norm2_b",
                 "norm3_w", "norm3_b", "wup"), (*keep[:6], w[1], w[3], w[5], w[7], w[8], w[9], w[10], w[11], keep[6])):
    setattr(aw, f, t.data_ptr())
x0 = torch.randn(B, 197, 768, generator=torch.Generator().manual_seed(17))
x = x0.bfloat16().view(B * 197, 768).to(dev).contiguous()
m8 = mask.to(dev).view(torch.uint8).contiguous()
out = torch.zeros(B * 197, 64, device=dev, dtype=torch.bfloat16)
dout = torch.zeros(B * 197, 768, device=dev, dtype=torch.bfloat16)
_cabi.call("hoigen_adapter_block", x.data_ptr(), None, kv.data_ptr(), m8.data_ptr(), C.byref(aw), out.data_ptr(), dout.data_ptr(), B, n_max)
torch.cuda.synchronize()
F = torch.nn.functional
xr = x0.bfloat16().float()
d = torch.relu(F.linear(xr, wd, bd))
def rest(t2):
    t = O._ln(d + t2, sd[blk + "norm2.weight"], sd[blk + "norm2.bias"])
    f = F.linear(torch.relu(F.linear(t, sd[blk + "linear1.weight"], sd[blk + "linear1.bias"])), sd[blk + "linear2.weight"], sd[blk + "linear2.bias"])
    return O._ln(t + f, sd[blk + "norm3.weight"], sd[blk + "norm3.bias"]).view(B * 197, 64)
Wi, bi = sd[blk + "multihead_attn.in_proj_weight"], sd[blk + "multihead_attn.in_proj_bias"]
Wo, bo = sd[blk + "multihead_attn.out_proj.weight"], sd[blk + "multihead_attn.out_proj.bias"]
t2 = O._mha(d, prior, prior, Wi, bi, Wo, bo, 2, mask)
got = out.float().cpu()
print("vs full reference      ", (got - rest(t2)).abs().max().item())
print("vs attention = bias only", (got - rest(bo.expand_as(t2))).abs().max().item())
# uniform attention over unmasked keys
v = F.linear(prior, Wi[128:], bi[128:])
um = torch.stack([v[b, :n].mean(0) for b, n in enumerate(n_list)])
print("vs uniform attention   ", (got - rest(F.linear(um, Wo, bo)[:, None, :].expand(B, 197, 64))).abs().max().item())
# per image error
e = (got - rest(t2)).abs().view(B, 197, 64)
print("per-image max err", e.amax(dim=(1, 2)).tolist())
print("per-head max err (cols 0-31 / 32-63) — after LN so mixed:", e[..., :32].max().item(), e[..., 32:].max().item())
print("rows of image 0 with err > .1:", (e[0].amax(-1) > .1).nonzero().flatten().tolist()[:20])
# hypotheses for images >= 1
def mha_with(pr_used, mask_used):
    return O._mha(d, pr_used, pr_used, Wi, bi, Wo, bo, 2, mask_used)
h1 = rest(mha_with(prior[0:1].expand(B, -1, -1).contiguous(), mask[0:1].expand(B, -1).contiguous()))
print("h1 (all rows use image 0's keys): per-image", (got - h1).abs().view(B, 197, 64).amax(dim=(1, 2)).tolist())
roll = rest(mha_with(torch.roll(prior, 1, 0), torch.roll(mask, 1, 0)))
print("h2 (keys of image b-1): per-image", (got - roll).abs().view(B, 197, 64).amax(dim=(1, 2)).tolist())
