"""Greedy generation throughput: CUDA-graph'd native decode loop (decode.py) vs HF's generate loop over the same model.

    python tools/bench_decode.py [--batch 64] [--new 64]
whisper-small geometry, 4 routed adapters r16, random-init weights, synthetic clips; EOS disabled so every row decodes
`--new` tokens.  Prints tokens/s (batch * new tokens / wall time incl. the encoder pass) for both paths.
"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from speech_adapter_routing_b200 import decode  # noqa: E402
from speech_adapter_routing_b200.routing import route  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--new", type=int, default=64)
    ap.add_argument("--skip-hf", action="store_true")
    ap.add_argument("--model", default="whisper-small")
    ap.add_argument("--adapters", type=int, default=4)
    ap.add_argument("--rank", type=int, default=16)
    args = ap.parse_args()
    if (args.model, args.adapters, args.rank) != ("whisper-small", 4, 16):
        bench.configure(args.model, args.adapters, args.rank, args.batch)
    dev = torch.device("cuda", 0)
    router, cfg, clips, g = bench.build_b200_workload(dev, seed=1234)
    B = args.batch
    x = clips(B, [i % bench.N_ADAPTERS for i in range(B)], g).to(dev).to(torch.bfloat16)
    kw = {"max_new_tokens": args.new, "eos_token_id": cfg.vocab_size + 5}   # an id that never wins: no early stop

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n, out

    sec, ids = timed(lambda: router.generate(x, **kw))
    print(f"native: {sec*1e3:.1f} ms per batch, {B*args.new/sec:.0f} tokens/s, {B/sec:.1f} clips/s, out {tuple(ids.shape)}", flush=True)
    assert decode.LAST_FALLBACK_REASON == "" or ids is not None
    if not args.skip_hf:
        with torch.no_grad():
            idx = router.detect_indices(router.extract_encoder_features(x)).idx

            def hf():
                with route(idx):
                    return router.whisper.generate(input_features=x, **kw)
            sec2, ids2 = timed(hf, n=1)
        print(f"HF loop: {sec2*1e3:.1f} ms per batch, {B*args.new/sec2:.0f} tokens/s, {B/sec2:.1f} clips/s, out {tuple(ids2.shape)}")
        print(f"speed-up {sec2/sec:.1f}x; token agreement {(ids[:, :ids2.shape[1]] == ids2[:, :ids.shape[1]]).float().mean().item():.4f}")


if __name__ == "__main__":
    main()
