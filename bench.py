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

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
WORKLOAD = "whisper-small routed fwd: 4 adapters r16 on q_proj/v_proj, LID pass + router + routed enc/dec pass, T_dec=128"
MODEL, N_ADAPTERS, RANK_R, BATCH_PER_GPU, T_DEC = "whisper-small", 4, 16, 64, 128
LANGUAGES = ["hindi", "italian", "punjabi", "telugu"]
ALL_LANGUAGES = ["hindi", "italian", "punjabi", "telugu", "english", "german", "french", "spanish"]


def configure(model: str, adapters: int, rank: int, batch: int) -> None:
    """Non-default workloads (BASELINE configs 3 / 4: whisper-medium r32, whisper-large-v3 8 adapters r64) for
    profiling runs; the driver's bench line always uses the defaults (config 2, the one the metric is quoted on)."""
    global WORKLOAD, MODEL, N_ADAPTERS, RANK_R, BATCH_PER_GPU, LANGUAGES
    MODEL, N_ADAPTERS, RANK_R, BATCH_PER_GPU = model, adapters, rank, batch
    LANGUAGES = ALL_LANGUAGES[:adapters]
    WORKLOAD = (f"{model} routed fwd: {adapters} adapters r{rank} on q_proj/v_proj, LID pass + router + routed enc/dec "
                f"pass, T_dec={T_DEC}")
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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
def run_reference_arm(args):
    """The reference's own CPU path (HF Whisper fp32 + PEFT-formula LoRA + per-utterance hard-routing loop —
    oracle/, since the reference has no native code to compile and `peft` is not installable offline), all host
    threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import fixtures, whisper as owhisper

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
        "config": {"workload": WORKLOAD, "sample_batch": sample_B, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample_B} clips per step (LID encoder pass + per-utterance routed forward), "
                                   f"median of {args.steps_ref} steps; oracle port of the reference path "
                                   "(HF Whisper fp32 + PEFT-formula LoRA); host has %d logical CPUs" % cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_oracle_workload(sample_B: int):
    import torch

    from oracle import fixtures, whisper as owhisper

    model = owhisper.build_whisper("small")
    cfg = model.config
    weights = owhisper.make_adapter_weights(model, RANK_R, N_ADAPTERS)
    sd = fixtures.make_router_state_dict(cfg.d_model, N_ADAPTERS)
    oracle = owhisper.RoutedWhisperOracle(model, weights, RANK_R, 2 * RANK_R, sd)
    protos = owhisper.make_input_features(N_ADAPTERS, cfg.num_mel_bins, list(range(N_ADAPTERS)), N_ADAPTERS, seed=99)
    oracle.router_sd = owhisper.fit_router_head(sd, oracle.lid_features(protos))
    langs = fixtures.language_mix(sample_B, N_ADAPTERS, "uniform")
    x = owhisper.make_input_features(sample_B, cfg.num_mel_bins, langs, N_ADAPTERS)
    dec, _ = owhisper.make_decoder_inputs(sample_B, T_DEC, cfg.vocab_size, cfg.decoder_start_token_id)
    return oracle, x, dec


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


# --------------------------------------------------------------------------------------------- B200 arm
def fit_head_gpu(clf, feats):
    """Closed-form nearest-centroid output layer so the synthetic languages route to distinct adapters."""
    import torch
    import torch.nn.functional as F

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


def build_b200_workload(dev, seed):
    import torch

    import speech_adapter_routing_b200 as sar

    model = sar.load_base_model(MODEL, device=dev, random_init=True)     # bf16 on CUDA (reference dtype policy)
    cfg = model.config
    for p in model.parameters():
        p.requires_grad = False
    lcfg = sar.LoraConfig(r=RANK_R, lora_alpha=2 * RANK_R, lora_dropout=0.0, target_modules=["q_proj", "v_proj"])
    for lang in LANGUAGES:
        sar.inject_lora(model, lcfg, adapter_name=lang)
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for m in sar.lora_modules(model).values():
            for lang in LANGUAGES:   # PEFT's zero-init lora_B would make the adapter path a no-op numerically
                m.lora_B[lang].weight.copy_((torch.randn(m.out_features, RANK_R, generator=g) * 0.02).to(dev))
    clf = sar.LanguageClassifier(input_dim=cfg.d_model, num_classes=N_ADAPTERS, languages=LANGUAGES).to(dev).eval()
    router = sar.AdapterRouter.from_stacked(model, clf, LANGUAGES, strategy="hard").eval()

    gt = torch.Generator().manual_seed(4234)
    templates = torch.rand(N_ADAPTERS, cfg.num_mel_bins, generator=gt) * 2 - 1

    def clips(B, langs, gen):
        return 0.5 * torch.randn(B, cfg.num_mel_bins, 3000, generator=gen) + templates[torch.tensor(langs)][:, :, None]

    protos = clips(N_ADAPTERS, list(range(N_ADAPTERS)), g).to(dev).to(torch.bfloat16)
    with torch.no_grad():
        fit_head_gpu(clf, router.extract_encoder_features(protos))
    return router, cfg, clips, g


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from speech_adapter_routing_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if _lib.lib().sar_device_ok() != 1:
        raise SystemExit("libsar: device is not sm_100")

    router, cfg, clips, g = build_b200_workload(dev, seed=1234 + rank)
    B = BATCH_PER_GPU
    n_in = 3   # rotate input buffers; every activation tensor (B*1500*768*2 B = 147 MB) already exceeds the 126 MB L2
    langs = [[(i + j) % N_ADAPTERS for i in range(B)] for j in range(n_in)]
    host_inputs = [clips(B, langs[j], g).pin_memory() for j in range(n_in)]          # fp32, as a data loader yields
    dev_inputs = [h.to(dev).to(torch.bfloat16) for h in host_inputs]
    dec = torch.randint(5, cfg.vocab_size, (B, T_DEC), generator=g)
    dec[:, 0] = cfg.decoder_start_token_id
    dec_dev = dec.to(dev)

    def step(x):
        with torch.no_grad():
            return router(x, decoder_input_ids=dec_dev)["logits"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------
    for i in range(args.warmup):
        step(dev_inputs[i % n_in])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ops.reset_counters()
    ops.K1_TIMELINE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    torch.cuda.nvtx.range_push("timed")
    e0.record()
    for i in range(args.steps):
        out = step(dev_inputs[i % n_in])
    e1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    timeline, ops.K1_TIMELINE = ops.K1_TIMELINE, None
    launches = sum(ops.LAUNCHES.values())
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: the tcgen05 pair kernel at the fused q|k|v + routed-LoRA call of the routed
    # pass's encoder layers (M = B*1500 rows, N = 3d), timed by CUDA events around every launch inside the timed region
    peaks, peak_src = load_peaks()
    M_big = B * 1500
    d = cfg.d_model
    fl, ms_k1, n_big = 0.0, 0.0, 0
    by_kind = {}
    ms_gemm_all = 0.0
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
    default_shape = (MODEL, RANK_R, B) == ("whisper-small", 16, 64)
    roofline = {"bound": "tensor", "kernel": "sar_attn_proj_fwd q|k|v + routed LoRA (M=%d, %d->%d, r=%d): k1v2 U pass + k1v2<256,AUG> dense tiles" % (M_big, d, 3 * d, RANK_R),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": K1_DRAM_TRAFFIC_BYTES if default_shape else None, "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                "launches_timed": n_big, "avg_launch_us": 1e3 * ms_k1 / max(n_big, 1),
                "gemm_share_of_step": ms_gemm_all / ms_total if ms_total else None,
                "algorithmic_flops_per_launch": fl / max(n_big, 1),
                "all_tcgen05_gemms": {k: {"launches": v[0], "avg_us": 1e3 * v[1] / v[0], "tflops": v[2] / (v[1] * 1e-3) / 1e12}
                                      for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1][1])[:20]}}

    # ---- end to end through the public API with host buffers ------------------------------------------------
    def e2e_step(x_host):
        x = x_host.to(dev, non_blocking=True).to(torch.bfloat16)
        logits = step(x)
        return logits.argmax(dim=-1).cpu()          # device->host read of the step's result (greedy token ids)

    for i in range(2):
        e2e_step(host_inputs[i % n_in])
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        toks = e2e_step(host_inputs[i % n_in])
    s1.record()
    barrier()
    te = torch.tensor([s0.elapsed_time(s1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item() / args.steps
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": host_inputs[0].numel() * 4, "d2h_bytes_per_step": toks.numel() * 8,
           "ms_per_step": e2e_ms, "api": "AdapterRouter.forward(input_features, decoder_input_ids)"}

    # ---- informational: the log-mel front-end (waveform -> input_features on the GPU, sar_logmel_fwd); the metric
    # itself is quoted from log-mel inputs like the reference's model API takes them
    frontend = None
    try:
        from speech_adapter_routing_b200 import logmel

        wav = 0.1 * torch.randn(B, logmel.N_SAMPLES, device=dev)
        for _ in range(2):
            logmel.log_mel_spectrogram(wav, n_mels=cfg.num_mel_bins)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(3):
            logmel.log_mel_spectrogram(wav, n_mels=cfg.num_mel_bins)
        f1.record()
        torch.cuda.synchronize()
        lm_ms = f0.elapsed_time(f1) / 3
        frontend = {"logmel_ms_per_batch": lm_ms, "clips_per_s_from_waveform": world * B / ((ms_step + lm_ms) * 1e-3),
                    "note": "sar_logmel_fwd on [B, 480000] fp32 waveforms resident in HBM, added to ms_per_step"}
        del wav
    except Exception as exc:   # never let the informational leg take the bench line down
        frontend = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "t_dec": T_DEC,
                       "parallelism": f"utterance-sharded x{world}, adapters replicated, no data-path collective",
                       "l2": "inputs rotate over 3 buffers; each activation tensor (147 MB) exceeds the 126 MB L2",
                       "weights": "random-init",
                       "rest_of_model": "libsar kernels for every projection / FFN / LayerNorm / conv front-end / lm head and the decoder's attention; encoder 1500x1500 softmax(QK^T)V = torch SDPA (cuDNN); embedding gathers = torch"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "frontend": frontend,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE call of the roofline op (fused q|k|v + routed LoRA, M = 96000,
# 768 -> 2304, r = 16, split path = two launches), from the `ncu --set full` capture summarised in
# profiles/r01_v5_pair_kernel_ncu_full_summary.csv: U pass 147.7 MB read + 5.9 MB written, dense AUG kernel 178.2 MB
# read + 390.5 MB written.  Algorithmic bytes = x 147.5 + y 442.4 + W 3.5 + adapters 0.3 = 593.7 MB; the surplus is
# the second read of x by the U pass (the single-launch kernel that keeps U on chip moves 584 MB but runs at 52 %
# tensor-active instead of 74 %: DESIGN.md §4).
K1_DRAM_TRAFFIC_BYTES = 722.3e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=2, help="clips per step of the reference (CPU) arm")
    ap.add_argument("--model", default=MODEL, help="profiling only: whisper-medium / whisper-large-v3 (default: config 2)")
    ap.add_argument("--adapters", type=int, default=N_ADAPTERS)
    ap.add_argument("--rank", type=int, default=RANK_R)
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="clips per GPU")
    args = ap.parse_args()
    if (args.model, args.adapters, args.rank, args.batch) != (MODEL, N_ADAPTERS, RANK_R, BATCH_PER_GPU):
        configure(args.model, args.adapters, args.rank, args.batch)
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
