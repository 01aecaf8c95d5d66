#!/usr/bin/env python
"""bench.py — routed multi-LoRA Whisper forward, 30 s clips per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path (libsar kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle) on the host cores

One "step" = one routed forward over one batch of synthetic clips, exactly the reference's AdapterRouter.forward
with strategy="hard" (src/models/adapter_router.py:568-625): LID encoder pass on base weights → router head →
routed encoder+decoder pass (teacher-forced, T_dec=128) with the per-utterance adapter on every q_proj / v_proj.
Workload at every N: BASELINE.json configs[1] — whisper-small geometry, 4 language adapters r16, mixed-language
batch 64 *per GPU* (weak scaling: utterances are independent, batch sharded, adapters replicated, no data-path
collective).  Weights are random-init (no hub access), inputs synthetic.

Beside the headline the line carries bounded extra legs under ``"extra"`` (``--extras none`` skips them):
  medium / large_v3   BASELINE configs[2] / [3] per-GPU shards (whisper-medium 4 x r32, whisper-large-v3 8 x r64, 64 clips
                      per GPU), a few steps each, with their own q|k|v+LoRA roofline;
  train               BASELINE configs[4]: whisper-small LoRA r16 training step (fused forward + backward layers, LoRA-only
                      gradients through K3, bf16 base, gradient checkpointing as in WhisperLoRA's default, clip + AdamW),
                      16 clips per GPU; `value` = the step replayed as one CUDA graph (GraphedTrainStep), `eager_value` =
                      the reference trainer's loop as written; NCCL all-reduce of the flat fp32 adapter-gradient bucket
                      at N > 1 with its own time and bytes;
  gpu_eager_baseline  (N = 1) the reference's per-utterance loop — HF Whisper + eager PEFT-formula LoRA — in bf16 on the
                      SAME B200 (oracle port moved to the GPU), on a bounded sample: what the kernels buy over torch eager.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "routed_multi_lora_whisper_fwd_clips_per_sec"
UNIT = "clips/s"
T_DEC = 128
ALL_LANGUAGES = ["hindi", "italian", "punjabi", "telugu", "english", "german", "french", "spanish"]


class Workload:
    """One routed-forward configuration (BASELINE.json configs[1..3])."""

    def __init__(self, model: str, adapters: int, rank: int, batch: int):
        self.model, self.adapters, self.rank, self.batch = model, adapters, rank, batch
        self.languages = ALL_LANGUAGES[:adapters]
        self.name = (f"{model} routed fwd: {adapters} adapters r{rank} on q_proj/v_proj, LID pass + router + routed "
                     f"enc/dec pass, T_dec={T_DEC}")


HEADLINE = Workload("whisper-small", 4, 16, 64)            # configs[1]: the configuration the metric is quoted on
EXTRA_WORKLOADS = {"medium": Workload("whisper-medium", 4, 32, 64),       # configs[2]: 256 clips over 4 GPUs
                   "large_v3": Workload("whisper-large-v3", 8, 64, 64)}   # configs[3]: 512 clips over 8 GPUs
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel (k1v2<256,AUG>: the dense q|k|v launch
# with the low-rank term as one extra K block, M = 96000, 768 -> 2304, r = 16), from the `ncu --set full` capture of THIS
# build summarised in profiles/r02_qkv_lora_ncu_full_summary.csv (tools/profile_all.sh regenerates it).  Algorithmic bytes
# of that launch: x 147.5 + y 442.4 + W 3.5 + U 6.1 + adapters 0.1 = 599.6 MB.  None until the capture exists.
K1_DRAM_TRAFFIC = {"bytes": None, "source": None}
_traffic_file = ROOT / "profiles" / "r02_qkv_lora_traffic.json"
if _traffic_file.exists():
    try:
        K1_DRAM_TRAFFIC = json.loads(_traffic_file.read_text())
    except Exception:
        pass


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return d, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 <= ts <= t1 + 0.1:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:   # timed region shorter than the sampling period: use every sample we have
            for ts, line in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def build_oracle_workload(sample_B: int, wl: Workload = HEADLINE, device: str = "cpu", dtype=None):
    import torch

    from oracle import fixtures, whisper as owhisper

    geo = wl.model.replace("whisper-", "")
    model = owhisper.build_whisper(geo)
    cfg = model.config
    weights = owhisper.make_adapter_weights(model, wl.rank, wl.adapters)
    sd = fixtures.make_router_state_dict(cfg.d_model, wl.adapters)
    if dtype is not None:
        model = model.to(dtype)
        weights = {k: (A.to(dtype), B.to(dtype)) for k, (A, B) in weights.items()}
    if device != "cpu":
        model = model.to(device)
        weights = {k: (A.to(device), B.to(device)) for k, (A, B) in weights.items()}
        sd = {k: v.to(device) for k, v in sd.items()}
    oracle = owhisper.RoutedWhisperOracle(model, weights, wl.rank, 2 * wl.rank, sd)
    protos = owhisper.make_input_features(wl.adapters, cfg.num_mel_bins, list(range(wl.adapters)), wl.adapters, seed=99)
    langs = fixtures.language_mix(sample_B, wl.adapters, "uniform")
    x = owhisper.make_input_features(sample_B, cfg.num_mel_bins, langs, wl.adapters)
    dec, _ = owhisper.make_decoder_inputs(sample_B, T_DEC, cfg.vocab_size, cfg.decoder_start_token_id)
    if device != "cpu" or dtype is not None:
        protos, x = protos.to(device, dtype or protos.dtype), x.to(device, dtype or x.dtype)
        dec = dec.to(device)
    feats = oracle.lid_features(protos)
    oracle.router_sd = {k: v.to(feats.device) for k, v in owhisper.fit_router_head(
        {k: v.float().cpu() for k, v in sd.items()}, feats.float().cpu()).items()}
    return oracle, x, dec


def run_reference_arm(args):
    """The reference's own CPU path (HF Whisper fp32 + PEFT-formula LoRA + per-utterance hard-routing loop —
    oracle/, since the reference has no native code to compile and `peft` is not installable offline), all host
    threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_B = args.ref_batch
    oracle, x, dec = build_oracle_workload(sample_B)
    times = []
    for i in range(args.warmup_ref + args.steps_ref):
        t0 = time.perf_counter()
        oracle.forward_hard(x, dec)
        dt = time.perf_counter() - t0
        if i >= args.warmup_ref:
            times.append(dt)
    sec = statistics.median(times)
    value = sample_B / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": HEADLINE.name, "sample_batch": sample_B, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample_B} clips per step (LID encoder pass + per-utterance routed forward), "
                                   f"median of {args.steps_ref} steps; oracle port of the reference path "
                                   "(HF Whisper fp32 + PEFT-formula LoRA); host has %d logical CPUs" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(max_seconds: float = 30.0):
    """Oracle timed on this box's host cores on a bounded sample (reported beside the GPU number, not a target)."""
    import torch

    cores = os.cpu_count() or 1
    prev = torch.get_num_threads()
    torch.set_num_threads(cores)
    sample_B = 2
    oracle, x, dec = build_oracle_workload(sample_B)
    oracle.forward_hard(x[:1], dec[:1])   # warm-up
    times = []
    t_start = time.perf_counter()
    for _ in range(3):
        t0 = time.perf_counter()
        oracle.forward_hard(x, dec)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > max_seconds:
            break
    torch.set_num_threads(prev)
    sec = statistics.median(times)
    return {"value": sample_B / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_B} clips per run (LID pass + per-utterance routed fwd, fp32), median of {len(times)} runs"}


def gpu_eager_baseline(dev, sample_B: int = 4):
    """The reference's path as torch would run it on THIS GPU: HF Whisper in bf16 (the reference's CUDA dtype policy,
    src/models/base.py:103-109) + eager PEFT-formula LoRA + the per-utterance batch-1 loop of adapter_router.py:610-622
    (the oracle port, moved to the device).  Bounded sample; informational — it is what the libsar kernels replace."""
    import torch

    oracle, x, dec = build_oracle_workload(sample_B, device=str(dev), dtype=torch.bfloat16)
    with torch.no_grad():
        oracle.forward_hard(x[:1], dec[:1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            oracle.forward_hard(x, dec)
        e1.record()
        torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / 2
    del oracle
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": sample_B / sec, "unit": UNIT, "kind": "port", "device": "cuda (same B200)", "dtype": "bf16",
            "sample": f"{sample_B} clips per run: LID encoder pass + per-utterance batch-1 routed forward, HF eager layers + "
                      "eager LoRA (torch / cuBLAS / cuDNN kernels), mean of 2 runs"}


# --------------------------------------------------------------------------------------------- B200 arm
def fit_head_gpu(clf, feats):
    """Closed-form nearest-centroid output layer so the synthetic languages route to distinct adapters."""
    import torch

    with torch.no_grad():
        f = clf.layer_norm(feats.float()).mean(1)
        z = f
        for i, layer in enumerate(clf.classifier):
            if i == len(clf.classifier) - 1:
                break
            z = layer(z)
        c = z - z.mean(0, keepdim=True)
        n = c.norm(dim=1, keepdim=True).clamp_min(1e-6)
        w = 8.0 * (c / n) / n.mean()
        clf.classifier[-1].weight.copy_(w)
        clf.classifier[-1].bias.copy_(-(w @ z.mean(0)))


def build_b200_workload(wl: Workload, dev, seed):
    import torch

    import speech_adapter_routing_b200 as sar

    model = sar.load_base_model(wl.model, device=dev, random_init=True)     # bf16 on CUDA (reference dtype policy)
    cfg = model.config
    for p in model.parameters():
        p.requires_grad = False
    lcfg = sar.LoraConfig(r=wl.rank, lora_alpha=2 * wl.rank, lora_dropout=0.0, target_modules=["q_proj", "v_proj"])
    for lang in wl.languages:
        sar.inject_lora(model, lcfg, adapter_name=lang)
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for m in sar.lora_modules(model).values():
            for lang in wl.languages:   # PEFT's zero-init lora_B would make the adapter path a no-op numerically
                m.lora_B[lang].weight.copy_((torch.randn(m.out_features, wl.rank, generator=g) * 0.02).to(dev))
    clf = sar.LanguageClassifier(input_dim=cfg.d_model, num_classes=wl.adapters, languages=wl.languages).to(dev).eval()
    router = sar.AdapterRouter.from_stacked(model, clf, wl.languages, strategy="hard").eval()

    gt = torch.Generator().manual_seed(4234)
    templates = torch.rand(wl.adapters, cfg.num_mel_bins, generator=gt) * 2 - 1

    def clips(B, langs, gen):
        return 0.5 * torch.randn(B, cfg.num_mel_bins, 3000, generator=gen) + templates[torch.tensor(langs)][:, :, None]

    protos = clips(wl.adapters, list(range(wl.adapters)), g).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        fit_head_gpu(clf, router.extract_encoder_features(protos))
    return router, cfg, clips, g


class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = dist
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch

        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_ms(self, ms: float) -> float:
        import torch

        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()


def qkv_roofline(timeline, M_big, d, rank_r, ms_total, peaks, peak_src):
    """Roofline of the dominant kernel from the per-launch CUDA events of the timed region: the tcgen05 pair kernel at
    the fused q|k|v + routed-LoRA call of the routed pass's encoder layers (M = B*1500 rows, N = 3d)."""
    fl, ms_k1, n_big, ms_gemm_all = 0.0, 0.0, 0, 0.0
    by_kind = {}
    for (kind, M, d_in, d_out, flops, a, b) in timeline:
        dt = a.elapsed_time(b)
        ms_gemm_all += dt
        lora = flops > 2.0 * M * d_in * d_out
        key = f"{kind} M={M} {d_in}->{d_out}" + (" +lora" if lora else "")
        k = by_kind.setdefault(key, [0, 0.0, 0.0])
        k[0] += 1; k[1] += dt; k[2] += flops
        if kind == "proj" and M == M_big and d_out == 3 * d and lora:
            fl += flops
            ms_k1 += dt
            n_big += 1
    achieved = fl / (ms_k1 * 1e-3) / 1e12 if ms_k1 > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
            "frac_of_burst_peak": achieved / float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])),
            "launches_timed": n_big, "avg_launch_us": 1e3 * ms_k1 / max(n_big, 1),
            "gemm_share_of_step": ms_gemm_all / ms_total if ms_total else None,
            "algorithmic_flops_per_launch": fl / max(n_big, 1),
            "all_tcgen05_gemms": {k: {"launches": v[0], "avg_us": 1e3 * v[1] / v[0], "tflops": v[2] / (v[1] * 1e-3) / 1e12}
                                  for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1][1])[:20]}}


def run_routed_leg(D: Dist, wl: Workload, steps: int, warmup: int, with_e2e: bool, sampler_on: bool):
    """Device-resident (and optionally end-to-end) timing of one routed-forward workload.  Returns a dict."""
    import torch

    from speech_adapter_routing_b200 import ops, whisper_blocks

    dev, world = D.dev, D.world
    router, cfg, clips, g = build_b200_workload(wl, dev, seed=1234 + D.rank)
    B = wl.batch
    n_in = 3   # rotate input buffers; every activation tensor (B*1500*d*2 B >= 147 MB) already exceeds the 126 MB L2
    langs = [[(i + j) % wl.adapters for i in range(B)] for j in range(n_in)]
    host_inputs = [clips(B, langs[j], g).pin_memory() for j in range(n_in)]          # fp32, as a data loader yields
    dev_inputs = [h.to(dev).to(torch.bfloat16) for h in host_inputs]
    dec = torch.randint(5, cfg.vocab_size, (B, T_DEC), generator=g)
    dec[:, 0] = cfg.decoder_start_token_id
    dec_dev = dec.to(dev)

    def step(x):
        with torch.no_grad():
            return router(x, decoder_input_ids=dec_dev)["logits"]

    for i in range(warmup):
        step(dev_inputs[i % n_in])
    D.barrier()
    sampler = ClockSampler(D.local_rank) if (sampler_on and D.rank == 0) else None
    ops.reset_counters()
    ops.K1_TIMELINE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    torch.cuda.nvtx.range_push("timed")
    e0.record()
    for i in range(steps):
        step(dev_inputs[i % n_in])
    e1.record()
    D.barrier()
    torch.cuda.nvtx.range_pop()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    timeline, ops.K1_TIMELINE = ops.K1_TIMELINE, None
    launches = sum(ops.LAUNCHES.values())
    ms_total = e0.elapsed_time(e1)
    ms_step = D.max_ms(ms_total) / steps
    peaks, peak_src = load_peaks()
    d = cfg.d_model
    roof = qkv_roofline(timeline, B * 1500, d, wl.rank, ms_total, peaks, peak_src)
    fused_lnu = whisper_blocks.FUSED_LN_U and ops.layernorm_lora_u_supported(d, wl.rank, 2)
    roof["kernel"] = ("sar_attn_proj_fwd q|k|v + routed LoRA (M=%d, %d->%d, r=%d): " % (B * 1500, d, 3 * d, wl.rank)) + (
        "ONE launch, k1v2<256,AUG> dense tiles with the low-rank term as one extra K block; U = scale·x·A_kᵀ comes from the "
        "fused LayerNorm kernel (its 2·M·r·d·2 flops are not credited here)" if fused_lnu else
        "two launches, k1v2 U pass + k1v2<256,AUG> dense tiles")
    out = {"value": world * B / (ms_step * 1e-3), "ms_per_step": ms_step, "clocks": clocks, "gpu_launches": launches,
           "roofline": roof, "d_model": d}

    if with_e2e:
        def e2e_step(x_host):
            x = x_host.to(dev, non_blocking=True).to(torch.bfloat16)
            logits = step(x)
            return logits.argmax(dim=-1).cpu()          # device->host read of the step's result (greedy token ids)

        for i in range(2):
            e2e_step(host_inputs[i % n_in])
        D.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(steps):
            toks = e2e_step(host_inputs[i % n_in])
        s1.record()
        D.barrier()
        e2e_ms = D.max_ms(s0.elapsed_time(s1)) / steps
        out["e2e"] = {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": host_inputs[0].numel() * 4, "d2h_bytes_per_step": toks.numel() * 8,
                      "ms_per_step": e2e_ms, "api": "AdapterRouter.forward(input_features, decoder_input_ids)"}

    # ---- HBM-bound kernels of the path, timed alone on tensors of the step's shapes (burst HBM peak)
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))

    def time_alone(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    hs = [torch.randn(B, 1500, d, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    it = [0]

    def nxt():
        it[0] += 1
        return hs[it[0] % 2]
    clf = router.classifier
    ms_k2 = time_alone(lambda: clf.route_batch(nxt()))
    k2_bytes = 2.0 * B * 1500 * d
    out["roofline_router"] = {"bound": "hbm", "kernel": "sar_router_fwd (k2_pool + k2_head): LayerNorm -> mean over T -> MLP -> softmax -> argmax -> segments",
                              "achieved": k2_bytes / (ms_k2 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": k2_bytes / (ms_k2 * 1e-3) / 1e9 / hbm_peak, "us": ms_k2 * 1e3,
                              "algorithmic_bytes": k2_bytes, "note": "both launches + output allocation, timed alone"}
    if fused_lnu:
        layer = router.whisper.model.encoder.layers[0]
        qkv = layer._sar_pack["self"].qkv.get()
        ln = layer._sar_pack["ln1"].get()
        idx = torch.tensor(langs[0], dtype=torch.int32, device=dev)
        ms_lnu = time_alone(lambda: ops.layernorm_lora_u_fwd(nxt(), ln.W, ln.b, qkv.A, idx, qkv.n_sets, qkv.scale))
        ms_ln = time_alone(lambda: ops.layernorm_fwd(nxt(), ln.W, ln.b))
        lnu_bytes = 4.0 * B * 1500 * d + 2.0 * B * 1500 * wl.rank * qkv.n_sets
        out["roofline_ln_u"] = {"bound": "hbm", "kernel": "sar_layernorm_lora_u_fwd: LayerNorm + U = scale·x·A_kᵀ for q and v in one pass over h",
                                "achieved": lnu_bytes / (ms_lnu * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": lnu_bytes / (ms_lnu * 1e-3) / 1e9 / hbm_peak, "us": ms_lnu * 1e3,
                                "algorithmic_bytes": lnu_bytes, "plain_layernorm_us": ms_ln * 1e3}
    del router, dev_inputs, host_inputs, hs
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_train_leg(D: Dist, steps: int, warmup: int):
    """BASELINE configs[4]: whisper-small LoRA r16 training step — forward + LoRA-only backward (K1 / K3 through autograd,
    bf16 base, fp32 adapters, gradient checkpointing as in WhisperLoRA's default), 16 clips per GPU, the flat fp32
    adapter-gradient bucket all-reduced over NCCL (src/training/trainer.py:251-277 semantics: clip after the reduce)."""
    import torch

    os.environ.setdefault("SAR_RANDOM_INIT", "1")
    import speech_adapter_routing_b200 as sar
    from speech_adapter_routing_b200 import ops
    from speech_adapter_routing_b200.dist import FlatGradBucket

    dev, world = D.dev, D.world
    w = sar.WhisperLoRA("whisper-small", lora_r=16, lora_alpha=32, lora_dropout=0.0, device=str(dev),
                        use_gradient_checkpointing=True)
    w.train()
    cfg = w.model.config
    params = [p for p in w.model.parameters() if p.requires_grad]
    bucket = FlatGradBucket(params)
    opt = torch.optim.AdamW(params, lr=1e-4)
    g = torch.Generator().manual_seed(100 + D.rank)
    B = 16
    xs = [torch.randn(B, cfg.num_mel_bins, 3000, generator=g).to(dev).to(torch.bfloat16) for _ in range(2)]
    labels = torch.randint(5, cfg.vocab_size, (B, T_DEC), generator=g).to(dev)
    ar = []

    bucket.enable_overlap(n_chunks=4)   # chunk all-reduces go out on a side stream as the gradients land in backward

    def step(i, timed=False):
        bucket.zero_()
        loss = w(input_features=xs[i % 2], labels=labels).loss
        loss.backward()
        if timed and world > 1:         # what is left of the collective after backward: the exposed (non-overlapped) part
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            bucket.finish_overlap()
            b.record()
            ar.append((a, b))
        else:
            bucket.finish_overlap()
        bucket.clip_grad_norm_(1.0)
        opt.step()
        return loss

    def timed_loop(fn, n_warm):
        for i in range(n_warm):
            fn(i)
        D.barrier()
        ops.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = fn(i, timed=True)
        e1.record()
        D.barrier()
        return D.max_ms(e0.elapsed_time(e1)) / steps, loss

    # (1) the reference trainer's loop as written (eager Python, one launch at a time)
    ms_eager, loss = timed_loop(step, warmup)
    eager_launches = {k: v for k, v in ops.LAUNCHES.items() if v}
    # (2) the same forward + backward (+ the chunked all-reduces) captured once as a CUDA graph and replayed per batch:
    #     the eager step is host-bound at 16 clips per GPU (torch.profiler: 64.6 ms of GPU work in 109 ms of host time)
    graph_err = None
    ms = ms_eager
    try:
        gstep = sar.GraphedTrainStep(w, bucket, xs[0], labels, warmup=2)

        def step_graphed(i, timed=False):
            loss = gstep(xs[i % 2], labels)
            bucket.clip_grad_norm_(1.0)
            opt.step()
            return loss

        ms_graph, loss = timed_loop(step_graphed, warmup)
        ms = ms_graph
    except Exception as e:                                   # report the eager number rather than nothing
        import traceback

        traceback.print_exc(file=sys.stderr)
        graph_err = f"{type(e).__name__}: {str(e)[:200]}"
        ms_graph = None
    out = {"metric": "whisper_small_lora_r16_train_step_clips_per_sec", "value": world * B / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "batch_per_gpu": B, "t_dec": T_DEC, "steps": steps, "gradient_checkpointing": True,
           "mode": "cuda_graph_replay" if ms_graph is not None else "eager",
           "eager_ms_per_step": ms_eager, "eager_value": world * B / (ms_eager * 1e-3),
           "loss": float(loss), "trainable_params": sum(p.numel() for p in params),
           "allreduce_bytes": bucket.buffer.numel() * 4, "libsar_launches_eager_loop": eager_launches,
           "scaling": "weak"}
    if graph_err:
        out["graph_error"] = graph_err
    if ar:
        ar_ms = statistics.mean(a.elapsed_time(b) for a, b in ar)
        # the same bucket all-reduced alone (no overlap): latency and bandwidth of the collective itself
        bucket.set_overlap_enabled(False)
        for _ in range(3):
            D.dist.all_reduce(bucket.buffer)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            D.dist.all_reduce(bucket.buffer)
        b.record()
        torch.cuda.synchronize()
        alone_ms = D.max_ms(a.elapsed_time(b) / 10)
        out["allreduce"] = {"backend": "nccl", "chunks": len(bucket._overlap.launched),
                            "exposed_ms_after_backward": ar_ms, "exposed_share_of_step": ar_ms / ms,
                            "alone_ms": alone_ms, "alone_busbw_gbs": 2.0 * (world - 1) / world * bucket.buffer.numel() * 4 / (alone_ms * 1e-3) / 1e9,
                            "note": "flat fp32 bucket in 4 chunks, each all-reduced in place on a side stream as soon as its "
                                    "gradients have landed (K3 writes them straight into the bucket); clip after the reduce"}
    del w, opt, bucket, xs
    gc.collect()
    torch.cuda.empty_cache()
    return out


def run_b200_arm(args):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    from speech_adapter_routing_b200 import _lib

    D = Dist()
    if _lib.lib().sar_device_ok() != 1:
        raise SystemExit("libsar: device is not sm_100")
    wl = Workload(args.model, args.adapters, args.rank, args.batch)
    head = run_routed_leg(D, wl, args.steps, args.warmup, with_e2e=True, sampler_on=True)
    default_shape = (wl.model, wl.rank, wl.batch) == (HEADLINE.model, HEADLINE.rank, HEADLINE.batch)
    head["roofline"]["traffic"] = K1_DRAM_TRAFFIC.get("bytes") if default_shape else None
    head["roofline"]["traffic_source"] = K1_DRAM_TRAFFIC.get("source") if default_shape else None

    extra = {}
    if args.extras != "none" and default_shape:
        legs = ["medium", "large_v3", "train", "gpu_eager_baseline"] if args.extras == "all" else args.extras.split(",")
        for name in legs:
            try:
                t0 = time.time()
                if name in EXTRA_WORKLOADS:
                    ewl = EXTRA_WORKLOADS[name]
                    r = run_routed_leg(D, ewl, args.extra_steps, 3, with_e2e=False, sampler_on=False)
                    extra[name] = {"workload": ewl.name, "batch_per_gpu": ewl.batch, "steps": args.extra_steps,
                                   "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "scaling": "weak",
                                   "gpu_launches": r["gpu_launches"], "roofline": {k: v for k, v in r["roofline"].items()
                                                                                   if k != "all_tcgen05_gemms"},
                                   "roofline_router": r["roofline_router"]}
                elif name == "train":
                    extra[name] = run_train_leg(D, args.extra_steps + 2, 3)
                elif name == "gpu_eager_baseline" and D.world == 1:
                    extra[name] = gpu_eager_baseline(D.dev)
                if name in extra:
                    extra[name]["leg_seconds"] = time.time() - t0
            except Exception as exc:   # an extra leg never takes the headline down
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                gc.collect()
                torch.cuda.empty_cache()

    # ---- informational: the log-mel front-end (waveform -> input_features on the GPU, sar_logmel_fwd); the metric
    # itself is quoted from log-mel inputs like the reference's model API takes them
    frontend = None
    try:
        from speech_adapter_routing_b200 import logmel

        n_mels = 128 if "large" in wl.model else 80
        wav = 0.1 * torch.randn(wl.batch, logmel.N_SAMPLES, device=D.dev)
        for _ in range(2):
            logmel.log_mel_spectrogram(wav, n_mels=n_mels)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(3):
            logmel.log_mel_spectrogram(wav, n_mels=n_mels)
        f1.record()
        torch.cuda.synchronize()
        lm_ms = f0.elapsed_time(f1) / 3
        frontend = {"logmel_ms_per_batch": lm_ms,
                    "clips_per_s_from_waveform": D.world * wl.batch / ((head["ms_per_step"] + lm_ms) * 1e-3),
                    "note": "sar_logmel_fwd on [B, 480000] fp32 waveforms resident in HBM, added to ms_per_step"}
        del wav
    except Exception as exc:   # never let the informational leg take the bench line down
        frontend = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    cpu = None
    if D.rank == 0 and D.world == 1 and not args.skip_cpu_baseline:
        cpu = cpu_baseline()

    if D.rank == 0:
        from speech_adapter_routing_b200 import whisper_blocks

        own_enc_attn = whisper_blocks.OWN_ATTN_MAX_TQ >= 1500
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl.name, "batch_per_gpu": wl.batch, "global_batch": D.world * wl.batch, "t_dec": T_DEC,
                       "parallelism": f"utterance-sharded x{D.world}, adapters replicated, no data-path collective",
                       "l2": "inputs rotate over 3 buffers; each activation tensor (>= 147 MB) exceeds the 126 MB L2",
                       "weights": "random-init",
                       "rest_of_model": "libsar kernels for every projection / FFN / LayerNorm / conv front-end / lm head and the decoder's attention; encoder 1500x1500 softmax(QK^T)V = "
                                        + ("libsar fa_fwd_kernel" if own_enc_attn else "torch SDPA (cuDNN)") + "; embedding gathers = torch"},
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "roofline_router": head.get("roofline_router"),
            "roofline_ln_u": head.get("roofline_ln_u"), "cpu_baseline": cpu, "frontend": frontend, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if D.world > 1:
        D.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=2, help="clips per step of the reference (CPU) arm")
    ap.add_argument("--extras", default="all", help="all | none | comma list of medium,large_v3,train,gpu_eager_baseline")
    ap.add_argument("--extra-steps", type=int, default=3, help="timed steps of each extra leg")
    ap.add_argument("--model", default=HEADLINE.model, help="profiling only: whisper-medium / whisper-large-v3 (default: config 2)")
    ap.add_argument("--adapters", type=int, default=HEADLINE.adapters)
    ap.add_argument("--rank", type=int, default=HEADLINE.rank)
    ap.add_argument("--batch", type=int, default=HEADLINE.batch, help="clips per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # the CPU arm is ~seconds per clip: bound its run to a few minutes whatever K/W the caller passes
    args.steps_ref = min(args.steps, 5)
    args.warmup_ref = min(args.warmup, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
