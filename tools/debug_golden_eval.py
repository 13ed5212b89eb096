import copy, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_models, bf16_emulation as be
from oracle.seeded_weights import fill_state_
import medsegpretrainimagenet_b200 as b
def rms(a, r): return ((a - r).double().pow(2).mean().sqrt() / (r.double().pow(2).mean().sqrt() + 1e-30)).item()
cases = torch.load("tests/golden/models.pt", weights_only=False)
for case in cases[:3]:
    name = case["name"]
    if name == "basic_unet_binary":
        m = ref_models.basic_unet(out_ch=1, final_activation="sigmoid", in_channels=case["x"].shape[1])
    else:
        m = ref_models.resnet50_attention_unet(out_ch=4 if "4class" in name else 1, final_activation="softmax" if "4class" in name else "sigmoid")
    fill_state_(m, 100)
    x = case["x"]
    for mode in ("eval", "train"):
        ref = copy.deepcopy(m); emu = copy.deepcopy(m); be.emulate_bf16_storage(emu)
        emu_w = be.round_weights_(copy.deepcopy(m)); be.emulate_bf16_storage(emu_w)
        gpu = b.convert(copy.deepcopy(m).cuda())
        for mm in (ref, emu, emu_w, gpu): mm.train(mode == "train")
        outs = {}
        with torch.no_grad():
            for k, mm in (("ref", ref), ("emu", emu), ("emu_w", emu_w)):
                torch.manual_seed(3); outs[k] = mm(x)
            torch.manual_seed(3); outs["gpu"] = gpu(x.cuda()).cpu()
        gold = case["y_eval"] if mode == "eval" else case["y_train"]
        print(f"{name} {mode}: ref-vs-golden {rms(outs['ref'], gold):.2e}  gpu-vs-golden {rms(outs['gpu'], gold):.4f}  emu(act only)-vs-golden {rms(outs['emu'], gold):.4f}  emu(act+weights)-vs-golden {rms(outs['emu_w'], gold):.4f}  gpu-vs-emu_w {rms(outs['gpu'], outs['emu_w']):.4f}  |y| max {gold.abs().max():.3f}")
        if "r50" in name and mode == "eval":
            with torch.no_grad():
                yr, sr = ref.encoder(x, return_skip_vals=True)
                ye, se = emu_w.encoder(x, return_skip_vals=True)
                yg, sg = b.convert(gpu.encoder)(x.cuda(), return_skip_vals=True)
            for i, (a, e, r) in enumerate(zip(sg + [yg], se + [ye], sr + [yr])):
                print(f"   encoder level {i}: gpu-vs-ref {rms(a.cpu(), r):.4f} emu_w-vs-ref {rms(e, r):.4f} gpu-vs-emu_w {rms(a.cpu(), e):.4f}  rms|act| {r.pow(2).mean().sqrt():.3g}")
