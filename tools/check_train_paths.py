"""Referee for the two training paths: gradients of the fused training layers (F) and of HF's layer bodies over the K1/K3
module slots (H), run in any order on ONE model in ONE process (argv[1], e.g. FHF), against fp32 autograd of the CPU oracle.
Order matters for what this guards: torch's cuDNN attention backward keeps one graph per q/k/v layout, so both paths must hand
it dO in the same layout (whisper_train._to_heads).  POISON=1 fills freed device memory with NaN first."""
import os, sys
sys.path.insert(0, "/root/repo")
os.environ.setdefault("SAR_RANDOM_INIT", "1")
import torch
import speech_adapter_routing_b200 as sar
from speech_adapter_routing_b200 import whisper_train as wt
from oracle import whisper as owhisper, lora as olora
dev = torch.device("cuda")
def build():
    torch.manual_seed(11)
    w = sar.WhisperLoRA("whisper-tiny", lora_r=16, lora_alpha=32, lora_dropout=0.0, device="cuda", use_gradient_checkpointing=False)
    w.train()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in sar.lora_modules(w.model).values():
            m.lora_B["default"].weight.copy_((torch.randn(m.out_features, 16, generator=g) * 0.02).to(dev))
    return w
def inputs(cfg):
    x = owhisper.make_input_features(2, cfg.num_mel_bins, [0, 0], 1, seed=8)
    _, labels = owhisper.make_decoder_inputs(2, 6, cfg.vocab_size, cfg.decoder_start_token_id)
    return x, labels
def run(w, fused):
    x, labels = inputs(w.model.config)
    for p in w.model.parameters(): p.grad = None
    wt.ENABLED = fused
    loss = w(input_features=x.to(dev).to(torch.bfloat16), labels=labels.to(dev)).loss
    loss.backward()
    wt.ENABLED = True
    torch.cuda.synchronize()
    return {n.replace("base_model.model.", "").replace(".default.weight", ""): p.grad.float().cpu().clone() for n, p in w.model.named_parameters() if p.grad is not None}
def cpu_truth(w):
    hf = w.model.base_model.model
    ref = owhisper.build_whisper("tiny")
    ref.load_state_dict({k: v.float().cpu() for k, v in hf.state_dict().items() if ".lora_" not in k and ".base_layer." not in k}, strict=False)
    params = {}
    for p, m in sar.lora_modules(hf).items():
        lin = ref.get_submodule(p)
        with torch.no_grad():
            lin.weight.copy_(m.base_layer.weight.float().cpu()); lin.bias.copy_(m.base_layer.bias.float().cpu())
        A = m.lora_A["default"].weight.detach().float().cpu().clone().requires_grad_(True)
        B = m.lora_B["default"].weight.detach().float().cpu().clone().requires_grad_(True)
        params[p] = (A, B)
        def fwd(x, lin=lin, A=A, B=B):
            return olora.lora_linear(x, lin.weight, lin.bias, A, B, 2.0)
        lin.forward = fwd
    x, labels = inputs(hf.config)
    ref.train(False)
    ref(input_features=x, labels=labels).loss.backward()
    out = {}
    for p, (A, B) in params.items():
        out[p + ".lora_A"] = A.grad.clone(); out[p + ".lora_B"] = B.grad.clone()
    return out
def rel(a, b): return max(((a[n] - b[n]).abs().max() / b[n].abs().max().clamp_min(1e-12)).item() for n in b)
order = sys.argv[1] if len(sys.argv) > 1 else "FH"
w = build()
truth = cpu_truth(w)
def poison():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    blocks = [torch.full((256 << 20,), float("nan"), dtype=torch.bfloat16, device=dev) for _ in range(8)]   # 4 GB of NaN
    small = [torch.full((n,), float("nan"), dtype=torch.bfloat16, device=dev) for n in (1 << 8, 1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20) for _ in range(32)]
    del blocks, small
    torch.cuda.synchronize()
for ch in order:
    if os.environ.get("POISON"):
        poison()
    g = run(w, ch == "F")
    nan = [n for n in g if not torch.isfinite(g[n]).all()]
    if nan:
        print(f"{ch}: {len(nan)} of {len(g)} parameter gradients are non-finite, e.g. {nan[:4]}")
        continue
    print(f"{ch}: vs CPU fp32 oracle {rel(g, truth):.3e}")
