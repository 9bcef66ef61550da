#!/usr/bin/env python
"""Top-k parity of the precision modes of the AttFind sweep at BASELINE sizes.

Runs the SAME job (same seeded generator, calibrated classifier, latents, noise) through several arms and compares
the per-class greedy picks + the merged list (NB:731-814), with the margins that explain the result:

  fp32        parity mode: fp32 FFMA generator kernels + the fp32 eager PyTorch classifier, TF32 off
  bench       throughput mode, exactly what bench.py times: bf16 tcgen05 generator + bf16 fused classifier
              (BN folded, s2d stem, native max-pool, native preprocessing)
  verify      bench + attfind_verify_topk: the candidates of the bench sweep re-evaluated in the parity mode (must give
              exactly the fp32 arm's picks at a fraction of its cost); needs the bench arm before it
  bf16g_fp32c bf16 generator + fp32 eager classifier (which half of the bench mode moves the effects?)
  oracle      the oracle's own torch functions (oracle/stylex_oracle.py: literal per-sample-weight grouped convs,
              full forwards, no prefix reuse) executed on the GPU in fp32 with TF32 off, batched over coord_shift

    python profiles/topk_parity.py --image-size 64 --latents 256 --arms fp32,bench,bf16g_fp32c,oracle \
        --out gpurun_out/topk_parity_64.json --dump gpurun_out/topk_parity_64.npz

Test infrastructure (imports the oracle as the checker); nothing here is on the product path.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--image-size", type=int, default=64)
    ap.add_argument("--latents", type=int, default=256)
    ap.add_argument("--classifier", default="resnet")
    ap.add_argument("--arms", default="fp32,bench,verify,bf16g_fp32c,oracle")
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--oracle-batch", type=int, default=128)
    ap.add_argument("--oracle-max-seconds", type=float, default=1200.0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--latent-seed", type=int, default=None, help="seed of the latents (default: --seed); 4242 = bench.py's job leg")
    ap.add_argument("--out", default=None)
    ap.add_argument("--dump", default=None, help="npz with the effects / base logits of the --dump-arms (float32)")
    ap.add_argument("--dump-arms", default="fp32,bench,bf16g_fp32c")
    args = ap.parse_args()

    import stylex_b200 as sx
    from stylex_b200 import _native, synthetic
    from oracle import stylex_oracle as O
    import helpers

    torch.set_grad_enabled(False)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _native.device_check()
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    size, kind, n = args.image_size, args.classifier, args.latents
    sd = synthetic.make_generator_state(size, seed=args.seed)
    model = synthetic.make_classifier_model(kind, args.seed)
    noise = synthetic.make_noise(size, args.seed).to(dev)
    G = sx.Generator(size, 514).to(dev)
    G.load_state_dict(sd, strict=False)
    L, S = G.num_layers, G.num_style_coords
    lat = synthetic.make_latents(n, args.seed if args.latent_seed is None else args.latent_seed).to(dev)

    # calibration exactly like bench.py: fp32 generator images of 32 seeded latents
    clf0 = sx.make_classifier(kind, copy.deepcopy(model).to(dev), size)
    G.precision = "fp32"
    calib_lat = synthetic.make_latents(32, 7).to(dev)
    calib = torch.cat([G(sx.styles_def_to_tensor([(calib_lat[i:i + 8], L)]).contiguous(), noise) for i in range(0, 32, 8)])
    synthetic.calibrate_classifier(clf0.model, clf0.preprocess, calib, chunk=8)
    model_cal = copy.deepcopy(clf0.model).float()

    def classifier(mode):
        c = sx.make_classifier(kind, copy.deepcopy(model_cal).to(dev), size)
        if mode == "bench":
            info = c.configure_throughput(calib[:8], dtype=torch.bfloat16 if kind == "resnet" else torch.float32)
        else:   # the wrapper as the reference has it: eager fp32 module, torchvision resize + Normalize
            info = {"classifier_mode": "eager fp32 (TF32 off)", "preprocess": "torch (resize, Normalize)"}
        return c, info

    results, record = {}, {"image_size": size, "latents": n, "S": S, "classifier": kind, "coord_evals": 2 * S * n, "arms": {}}

    def run_ours(name, precision, clf_mode):
        c, info = classifier(clf_mode)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sx.attfind_sweep(G, c, lat, noise, precision=precision, max_batch=args.max_batch)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        picks, merged, scores = sx.attfind_select(res["style_change"], res["base_prob"], 5, 0.5)
        sweeps[name] = res
        results[name] = (res["style_change"].cpu().numpy(), res["base_prob"].cpu().numpy())
        record["arms"][name] = {"generator": precision, "classifier": info, "seconds": dt, "coord_evals_per_s": 2 * S * n / dt,
                                "picks": {str(k): [list(p) for p in v] for k, v in picks.items()},
                                "merged": [list(p) for p in merged], "scores": scores}
        print(f"[{name}] {dt:.1f} s, {2 * S * n / dt:.0f} coord-evals/s, picks {picks}", flush=True)

    sweeps = {}

    def run_verify(name):
        c, info = classifier("fp32")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        picks, merged, scores, vi = sx.attfind_verify_topk(G, c, lat, noise, sweeps["bench"], 5, 0.5, precision="fp32", max_batch=128)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        results[name] = (vi["style_change"].cpu().numpy(), vi["base_prob"].cpu().numpy())
        record["arms"][name] = {"generator": "bf16 sweep, candidates re-evaluated in fp32", "classifier": info, "seconds": dt,
                                "picks": {str(k): [list(p) for p in v] for k, v in picks.items()},
                                "merged": [list(p) for p in merged], "scores": scores,
                                "verify": {k: v for k, v in vi.items() if k not in ("style_change", "base_prob")},
                                "exact_fraction": vi["exact_evals"] / (2 * S * n)}
        print(f"[{name}] +{dt:.1f} s, {vi['candidates']} columns exact ({100 * vi['exact_evals'] / (2 * S * n):.2f} % of the coord-evals), "
              f"band {vi['band']:.3e}, verified {vi['verified']}, picks {picks}", flush=True)

    def run_oracle(name):
        """oracle functions on the GPU: full forwards with a functional coordinate shift, batch = oracle_batch."""
        params = {k: v.to(dev) for k, v in sd.items()}
        c, info = classifier("fp32")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        styles = O.styles_def_to_tensor([(lat, L)])
        sc = torch.empty(n, S, device=dev)
        base = torch.empty(n, 2, device=dev)
        for i in range(0, n, 32):
            img, s_ = O.generator_forward(params, styles[i:i + 32], noise, get_style_coords=True)
            sc[i:i + 32] = s_
            base[i:i + 32] = c.classify_images(img)
        minima, maxima = O.get_min_max_style_vectors(sc)
        eff = torch.zeros(n, 2, S, 2, device=dev)
        half = args.oracle_batch // 2
        done = 0
        for i in range(n):
            w = styles[i:i + 1]
            for s0 in range(0, S, half):
                cnt = min(half, S - s0)
                idx = torch.arange(s0, s0 + cnt, device=dev)
                shift = torch.zeros(2 * cnt, S, device=dev)
                shift[torch.arange(cnt, device=dev), idx] = minima[idx] - sc[i, idx]            # direction 0: towards the minimum
                shift[cnt + torch.arange(cnt, device=dev), idx] = maxima[idx] - sc[i, idx]      # direction 1: towards the maximum
                img = O.generator_forward(params, w.expand(2 * cnt, -1, -1), noise, coord_shift=shift)
                lg = c.classify_images(img) - base[i]
                eff[i, 0, s0:s0 + cnt] = lg[:cnt]
                eff[i, 1, s0:s0 + cnt] = lg[cnt:]
            done = i + 1
            if done % 8 == 0:
                torch.cuda.synchronize()
                el = time.perf_counter() - t0
                print(f"[{name}] {done}/{n} latents, {el:.0f} s", flush=True)
                if el > args.oracle_max_seconds and done < n:
                    print(f"[{name}] time budget reached, stopping after {done} latents", flush=True)
                    break
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e_np, b_np = eff.cpu().numpy(), base.cpu().numpy()
        rec = {"generator": "oracle.generator_forward (literal grouped conv), fp32, TF32 off, GPU", "classifier": info,
               "seconds": dt, "latents_done": done}
        if done == n:
            picks, merged, scores = O.attfind_select(e_np, b_np, 5, 0.5)
            rec.update({"picks": {str(k): [list(p) for p in v] for k, v in picks.items()}, "merged": [list(p) for p in merged],
                        "scores": scores})
            print(f"[{name}] {dt:.1f} s, picks {picks}", flush=True)
        results[name] = (e_np[:done], b_np)
        record["arms"][name] = rec

    for arm in args.arms.split(","):
        if arm == "fp32":
            run_ours("fp32", "fp32", "fp32")
        elif arm == "bench":
            run_ours("bench", "bf16", "bench")
        elif arm == "bf16g_fp32c":
            run_ours("bf16g_fp32c", "bf16", "fp32")
        elif arm == "verify":
            run_verify("verify")
        elif arm == "oracle":
            run_oracle("oracle")
        else:
            raise SystemExit(f"unknown arm {arm}")

    # ---- comparisons: every arm against the first one that ran in full
    ref_name = "fp32" if "fp32" in results else next(iter(results))
    e_ref, b_ref = results[ref_name]
    record["reference_arm"] = ref_name
    record["margins"] = {}
    for name, (e, b) in results.items():
        if name == ref_name:
            record["margins"][name] = helpers.selection_margin_report(e_ref, b_ref)
            continue
        m = e.shape[0]
        if m == e_ref.shape[0]:
            rep = helpers.selection_margin_report(e_ref, b_ref, e, b)
            rep["picks_equal"] = (record["arms"][name].get("picks") == record["arms"][ref_name]["picks"]
                                  and record["arms"][name].get("merged") == record["arms"][ref_name]["merged"])
        else:   # partial oracle run: element-wise comparison on the latents that were swept
            rep = {"latents_compared": m, "max_abs_effect_err": float(np.abs(e - e_ref[:m]).max()),
                   "base_err": float(np.abs(b - b_ref).max()), "picks_equal": None}
        record["margins"][name] = rep
        print(f"[{name} vs {ref_name}] picks_equal={rep.get('picks_equal')} max|eff err|={rep.get('max_abs_effect_err'):.3e} "
              f"base err={rep.get('base_err')} min gap={rep.get('min_gap')} worst 2err/gap={rep.get('worst_2err_over_gap')}", flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(record, open(args.out, "w"), indent=1)
    if args.dump:
        arrs = {}
        for name, (e, b) in results.items():
            if name not in args.dump_arms.split(","):
                continue
            arrs[name + ".effects"] = e.astype(np.float32)
            arrs[name + ".base"] = b.astype(np.float32)
        np.savez_compressed(args.dump, **arrs)
    print(json.dumps({k: v for k, v in record.items() if k != "margins"})[:2000])


if __name__ == "__main__":
    main()
