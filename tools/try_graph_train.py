"""Eager vs captured training step (whisper-small r16, 16 clips): timing and gradient agreement.  argv[1] = "ckpt" turns gradient
checkpointing on; env: DROPOUT=p (lora_dropout), OPT=1 (AdamW + clip inside the eager loop), OVERLAP=1 (chunked all-reduce
hooks), SAR_TRAIN_SDPA=cudnn|flash, SAR_TRAIN_OWN_LN=0, SAR_TRAIN_FUSED_GELU_BWD=0 (A/B switches of whisper_train.py)."""
import os, sys, time
sys.path.insert(0, "/root/repo")
os.environ.setdefault("SAR_RANDOM_INIT", "1")
import torch
import speech_adapter_routing_b200 as sar
dev = torch.device("cuda")
for ckpt in ((False,) if len(sys.argv) < 2 else (True,)):
    w = sar.WhisperLoRA("whisper-small", lora_r=16, lora_alpha=32, lora_dropout=float(os.environ.get("DROPOUT", "0")), device="cuda", use_gradient_checkpointing=ckpt)
    w.train()
    cfg = w.model.config
    params = [p for p in w.model.parameters() if p.requires_grad]
    bucket = sar.FlatGradBucket(params)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, cfg.num_mel_bins, 3000, generator=g).to(dev).to(torch.bfloat16)
    labels = torch.randint(5, cfg.vocab_size, (16, 128), generator=g).to(dev)
    opt = torch.optim.AdamW(params, lr=1e-4) if os.environ.get("OPT") else None
    if os.environ.get("OVERLAP"): bucket.enable_overlap(n_chunks=4)
    def eager():
        bucket.zero_(); l = w(input_features=x, labels=labels).loss; l.backward()
        if os.environ.get("OVERLAP"): bucket.finish_overlap()
        if opt is not None:
            bucket.clip_grad_norm_(1.0); opt.step()
        return l.detach()
    for _ in range(3): eager()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): l = eager()
    torch.cuda.synchronize(); te = (time.time() - t0) / 5
    ge = bucket.buffer.clone()
    if ckpt:
        hfm = w.model.base_model.model
        hfm.gradient_checkpointing_disable(); eager(); gp = bucket.buffer.clone(); hfm.gradient_checkpointing_enable()
        eager(); ge2 = bucket.buffer.clone()
        print(f"eager ckpt vs eager plain rel diff {float((ge - gp).abs().max() / gp.abs().max()):.3e}; ckpt again {float((ge2 - gp).abs().max() / gp.abs().max()):.3e}")
    from speech_adapter_routing_b200 import whisper_train as wt
    print('fused layer calls', wt.CALLS, 'refused', wt.REFUSED)
    try:
        step = sar.GraphedTrainStep(w, bucket, x, labels)
        for _ in range(2): step(x, labels)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5): lg = step(x, labels)
        torch.cuda.synchronize(); tg = (time.time() - t0) / 5
        err = ((bucket.buffer - ge).abs().max() / ge.abs().max()).item()
        if err > 5e-2:
            names = {id(p): n for n, p in w.model.named_parameters()}
            for i_, p_ in enumerate(bucket.params):
                a = bucket.buffer[bucket._offsets[i_]:bucket._offsets[i_] + p_.numel()]; b = ge[bucket._offsets[i_]:bucket._offsets[i_] + p_.numel()]
                print(f"   {names[id(p_)][-60:]:60s} graph {float(a.abs().max()):.3e} eager {float(b.abs().max()):.3e}")
        print(f"ckpt={ckpt}: eager {te*1e3:.1f} ms  graphed {tg*1e3:.1f} ms  loss {float(l):.4f} / {float(lg):.4f}  grad rel diff {err:.2e}")
    except Exception as e:
        import traceback; traceback.print_exc(); print(f"ckpt={ckpt}: eager {te*1e3:.1f} ms  graph capture FAILED: {type(e).__name__}: {str(e)[:300]}")
    del w, bucket
    torch.cuda.empty_cache()
