"""Layer-wise error growth of the converted model vs the fp32 oracle (diagnostic, GPU box)."""
import copy, sys
import torch
sys.path.insert(0, ".")
from oracle import ref_models
import medsegpretrainimagenet_b200 as b
from medsegpretrainimagenet_b200 import converter as cv, functional as Fn

def stats(a, r):
    d = (a - r).double(); r = r.double()
    return f"rms_rel={(d.pow(2).mean().sqrt() / (r.pow(2).mean().sqrt() + 1e-30)).item():.4f} max_rel={(d.abs().max() / (r.abs().max() + 1e-30)).item():.4f}"

x = torch.rand((4, 3, 192, 192), generator=torch.Generator().manual_seed(2))
for sdr in (0.0, 0.1):
    torch.manual_seed(0)
    ref = ref_models.kaiming_init_(ref_models.resnet50_attention_unet(out_ch=4, final_activation="softmax", stochastic_depth_rate=sdr))
    gpu = b.convert(copy.deepcopy(ref).cuda())
    ref.train(); gpu.train()
    torch.manual_seed(5); y_ref, s_ref = ref.encoder(x, return_skip_vals=True)
    torch.manual_seed(5); y, s = b.convert(gpu.encoder)(x.cuda(), return_skip_vals=True)
    print(f"R50 encoder sdr={sdr}")
    for i, (a, r) in enumerate(zip(s + [y], s_ref + [y_ref])):
        print("  level", i, tuple(r.shape), stats(a.detach().cpu(), r.detach()))
    # decoder alone, teacher-forced with the oracle's encoder outputs
    ctx = cv.ExecContext()
    to = lambda t: Fn.to_nhwc(t.detach().cuda())
    out_ref = ref.final_act(ref.decoder(y_ref, list(s_ref)))
    out = cv.run_unet_decoder(ctx, gpu.decoder, to(y_ref), [to(t) for t in s_ref], gpu.final_act)
    print("  decoder (teacher-forced)", stats(out.detach().cpu(), out_ref.detach()))
    torch.manual_seed(5); full_ref = ref(x)
    torch.manual_seed(5); full = gpu(x.cuda())
    print("  full", stats(full.detach().cpu(), full_ref.detach()), " logit-free prob range", full_ref.min().item(), full_ref.max().item())
    # trace decoder internals
    hooks, rec = [], {}
    for name, m in ref.decoder.named_modules():
        if type(m).__name__ in ("ConvBlock", "UpConvBlock", "AttentionBlock"):
            hooks.append(m.register_forward_hook(lambda mod, a, o, name=name: rec.__setitem__(name, o.detach())))
    ref.decoder(y_ref, list(s_ref))
    for h in hooks: h.remove()
    # same on gpu via monkeypatching run_module to record
    got = {}
    names = {id(m): n for n, m in gpu.decoder.named_modules()}
    orig_rm, orig_att = cv.run_module, cv.run_attention_block
    def rm(ctx_, m, xx, **kw):
        o = orig_rm(ctx_, m, xx, **kw)
        n = names.get(id(cv._unwrap(m)))
        if n in rec: got[n] = o
        return o
    def ra(ctx_, m, *a):
        o = orig_att(ctx_, m, *a)
        got[names[id(m)]] = o
        return o
    cv.run_module, cv.run_attention_block = rm, ra
    cv.run_unet_decoder(ctx, gpu.decoder, to(y_ref), [to(t) for t in s_ref], gpu.final_act)
    cv.run_module, cv.run_attention_block = orig_rm, orig_att
    for n in rec:
        if n in got:
            print("   ", n, stats(Fn.to_nchw(got[n]).detach().cpu(), rec[n]))
