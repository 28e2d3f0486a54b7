#!/usr/bin/env python
"""bench.py — VQA forward questions/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workloads updown,regat]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path on the box's host cores

A step = one forward pass (question encoder → top-down attention → [ReGAT] → classifier →
answers) over one batch of 1024 synthetic questions per GPU (36x2048 region features, 14
tokens, 3129 answers), inputs resident in HBM as bf16.  The batch is sharded data-parallel
(weak scaling: 1024 questions per GPU, no forward collective).  Rank 0 prints ONE JSON line:

  * ``value`` ...            Up-Down forward (BASELINE configs[1], the configuration the metric is quoted on)
  * ``regat`` {...}          the same keys for Up-Down + ReGAT (configs[2], the north-star workload): value,
                             ms_per_step, e2e, roofline(s), parity, per-rank times — at EVERY N
  * ``roofline``             the LONGEST kernel of the Up-Down step, ``roofline_kernels`` every hot kernel timed alone
  * ``parity``               answers / logits of batch 0 against the CPU oracle at the full batch (N = 1)

Every timed step is one CUDA-graph launch (the whole forward incl. its side stream is captured once per resident
batch); 4 resident batches of 151 MB rotate, so no step finds its features in the 126 MB L2.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "VQA forward questions/sec (36x2048 regions, bs=1024)"
UNIT = "questions/s"
FLOPS_PER_Q = {"updown": 290_500_608, "regat": 1_214_553_216}       # SURVEY.md §8d (minimal algebra)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")                    # git-ignored copy of the reference (build())
WORKLOAD_NAME = {
    "updown": "Up-Down VQA forward bf16 batch 1024 on 1xB200 (tcgen05 projections + fused attention/softmax)",
    "regat": "ReGAT spatial-relation VQA forward (11 relation labels, KxK masked graph attention) batch 1024",
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch at B = 1024 from the `ncu --set full` captures committed
# under profiles/ (bytes, file); scaled by B / 1024 below.  None = no capture of this kernel in this round.
NCU_TRAFFIC = {
    "wv": (163.7e6, "profiles/r02_ncu_wv.md"),        # 159.47 read + 4.21 written; algorithmic 151.0 + 4.2 + 0.6
    "wide": (588.8e6, "profiles/r02_ncu_wide.md"),    # 178.8 read + 410.0 written; algorithmic 151 + 25 + 453 (part of Y is still in L2 at kernel end)
    "gat": (613.3e6, "profiles/r02_ncu_gat.md"),      # 606.3 read + 7.1 written; algorithmic 604 + 4.2
    "pool": (156.8e6, "profiles/r02_ncu_pool.md"),    # 151.6 read + 5.2 written; algorithmic 151.0 + 4.3
    "gru": (18.0e6, "profiles/r02_ncu_gru.md (gru_persistent_kernel, the single-CTA variant: ncu cannot launch the "
                    "cooperative cluster kernel; operands are L2-resident)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workloads", default="updown,regat,train",
                    help="comma list; the first one is the line's `value`; `train` = the config-4 training-step block")
    ap.add_argument("--workload", default=None, help="(compat) a single workload")
    ap.add_argument("--batch", type=int, default=1024, help="questions per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32tc"])
    ap.add_argument("--no-exact-block", action="store_true",
                    help="skip the fp32tc block (fp32-class arithmetic on the tensor cores: exact answers, 1e-5 logits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="direct launches instead of CUDA graph replays")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--overlap", type=int, default=None, help="force the encoder-beside-projection schedule on (1) / off (0)")
    ap.add_argument("--side-sms", type=int, default=0)
    ap.add_argument("--side-permille", type=int, default=0)
    ap.add_argument("--chase", type=int, default=None, help="SMs of the graph attention chasing the wide projection (0 = serial)")
    a = ap.parse_args()
    a.workloads = [a.workload] if a.workload else [w for w in a.workloads.split(",") if w]
    a.train_block = "train" in a.workloads
    a.workloads = [w for w in a.workloads if w != "train"]
    for w in a.workloads:
        if w not in WORKLOAD_NAME:
            ap.error(f"unknown workload {w}")
    if not a.workloads:
        ap.error("need at least one forward workload")
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons through NVML, sampled by the MAIN thread while it waits for the timed region's end
    event (no polling thread: 8 ranks x a 5 ms poller on a shared 32-vCPU box is measurable jitter in a 7 ms window)"""

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_mhz = None

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def wait_for(self, event):
        """sample until `event` has completed (at least once, during the region)"""
        self.sample()
        while not event.query():
            self.sample()
            time.sleep(0.0005)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def build_reference_model(workload):
    """the UNMODIFIED reference (baseline/_ref, copied there from /root/reference by __graft_entry__.build()):
    set_model(...) exactly as BASELINE.md §3 states, default initialisers, eval mode.  None when the copy is absent."""
    if not os.path.isdir(os.path.join(REF_DIR, "modules")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import contextlib
    import io
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):     # the reference prints while it builds
        warnings.simplefilter("ignore")
        from modules.wrapper import set_model              # the reference's own module tree
        torch.manual_seed(1111)
        m = set_model(encoder_type="relation" if workload == "regat" else "base", predictor_type="base",
                      decoder_type="none", ntoken=20000, v_dim=2048, embed_dim=300, hidden_dim=1024,
                      decoder_hidden_dim=512, rnn_layer=1, ans_dim=3129, cls_layer=2, c_len=20, device="cpu",
                      dropout=0.5, neg_slope=0.5, rnn_type="GRU", att_type="new", conv_layer=1, conv_type="corr",
                      pretrained_embed_path="")
    return m.eval()


def cpu_forward_qps(workload, B, iters, warmup=1, threads=None):
    """The reference's CPU path on the host cores: Wrapper.forward_vqa(batch) (wrapper.py:113-118) of the real
    reference when baseline/_ref is present (kind "reference"), else the oracle port (kind "port").  A step is the
    WHOLE batch of B questions; the ReGAT batch goes through forward_vqa in chunks of 128 questions (the reference
    gathers a [B,36,36,2048] f32 label-bias tensor, gcn.py:107: 11 GB at B = 1024)."""
    from oracle import vqa_oracle as O                       # synthetic inputs (+ the port when the reference is absent)
    cfg = O.FULL_REGAT if workload == "regat" else O.FULL
    cores = threads or host_cores()
    torch.set_num_threads(cores)
    batch = O.make_batch(cfg, B, 2000)
    batch = {k: v for k, v in batch.items() if k not in ("bbox", "wh")}
    chunk = 128 if workload == "regat" else B
    model = build_reference_model(workload)
    if model is None:
        W = O.make_weights(cfg, 1111)
        run = lambda b: O.forward_vqa(b, W, cfg)
        kind = "port"
    else:
        run = model.forward_vqa
        kind = "reference"
    times = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            for c0 in range(0, B, chunk):
                run({k: (v[c0:c0 + chunk] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B else v)
                     for k, v in batch.items()})
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sample = (f"{B} questions per step ({'in chunks of %d' % chunk if chunk < B else 'one call'}), fp32, "
              f"{'the reference Wrapper.forward_vqa from baseline/_ref' if kind == 'reference' else 'torch-CPU oracle port of Wrapper.forward_vqa'}, "
              f"{cores} threads, {len(times)} timed passes")
    return {"qps": B / (sum(times) / len(times)), "best": B / min(times), "times": times, "cores": cores,
            "kind": kind, "sample": sample, "B": B}


def workload_config(args, workload, extra=None):
    cfg = {"workload": WORKLOAD_NAME[workload], "batch_per_gpu": args.batch, "regions": 36, "v_dim": 2048, "hidden": 1024,
           "tokens": 14, "answers": 3129, "ntoken": 20000, "parallelism": f"dp{args.gpus}",
           "l2": "4 rotating resident batches (151 MB bf16 features each > 126 MB L2)"}
    cfg.update(extra or {})
    return cfg


def run_reference(args, rank):
    """--impl reference: the reference's CPU path, full batch per step, rank 0 only"""
    if rank != 0:
        return
    blocks = {}
    for wl in args.workloads:
        # bounded: the driver's K/W apply to the first workload; the others get 2 timed passes
        iters = max(args.steps, 1) if wl == args.workloads[0] else min(max(args.steps, 1), 2)
        warm = max(args.warmup, 1) if wl == args.workloads[0] else 1
        if wl == "regat":
            iters, warm = min(iters, 3), 1            # ~3-8 s per 1024-question step
        r = cpu_forward_qps(wl, args.batch, iters, warm)
        blocks[wl] = {"value": r["qps"], "unit": UNIT, "ms_per_step": 1e3 * sum(r["times"]) / len(r["times"]),
                      "steps": len(r["times"]), "best_of": r["best"],
                      "cpu_baseline": {"value": r["qps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                       "sample": r["sample"], "cpu_model": cpu_model()},
                      "e2e": {"value": r["qps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "config": workload_config(args, wl)}
    first = blocks[args.workloads[0]]
    line = {"impl": "reference", "metric": METRIC, "value": first["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": first["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": first["config"],
            "cpu_baseline": first["cpu_baseline"], "e2e": first["e2e"], "gpu_launches": 0, "same_config": True}
    for wl in args.workloads[1:]:
        line[wl] = blocks[wl]
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
class Ctx:
    pass


def all_max(ctx, ms):
    """max over ranks + the per-rank list (device-timed, NCCL all_gather of one double)"""
    if ctx.dist is None:
        return ms, [ms]
    t = torch.tensor([ms], device=ctx.dev, dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(ctx.world)]
    ctx.dist.all_gather(out, t)
    per = [float(x.item()) for x in out]
    return max(per), per


def barrier(ctx):
    torch.cuda.synchronize()
    if ctx.dist is not None:
        ctx.dist.barrier()
    torch.cuda.synchronize()


def parity_block(O, cfg, W, eng, batch_cpu, out, relation):
    """answers / logits / attention of one full batch against the CPU oracle (the checker, rank 0 only).
    margin-ok rows = rows whose reference top-2 margin exceeds 4x the largest logit error of the run (SURVEY H1:
    a bf16 path cannot be asked to reproduce an fp32 near-tie)."""
    B = batch_cpu["img"].shape[0]
    chunk = 128 if relation else 256
    ref_logits, ref_att = [], []
    with torch.no_grad():
        for c0 in range(0, B, chunk):
            sub = {k: (v[c0:c0 + chunk] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B else v)
                   for k, v in batch_cpu.items()}
            lg, enc = O.forward(sub, W, cfg)
            ref_logits.append(lg)
            ref_att.append(enc["v_att"][:, :, 0])
    ref_logits, ref_att = torch.cat(ref_logits), torch.cat(ref_att)
    logits, att, label = out["logits"].float().cpu(), out["att"].float().cpu(), out["label"].cpu()
    ref_label = ref_logits.max(1)[1]
    err = float((logits - ref_logits).abs().max())
    top2 = ref_logits.topk(2, dim=1)[0]
    margin_ok = (top2[:, 0] - top2[:, 1]) > 4.0 * err
    equal = label == ref_label
    return {"n": int(B), "n_equal": int(equal.sum()), "n_margin_ok": int(margin_ok.sum()),
            "n_equal_margin_ok": int((equal & margin_ok).sum()),
            "n_equal_outside_margin": int((equal & ~margin_ok).sum()),
            "max_rel_logit_err": err / float(ref_logits.abs().max()),
            "max_rel_att_err": float((att - ref_att).abs().max() / ref_att.abs().max()),
            "oracle": "oracle/vqa_oracle.py forward() fp32 on the same seeded batch (pinned to the real reference's goldens)"}


def run_workload(ctx, args, wl):
    O, ops = ctx.O, ctx.ops
    from vqa_collection_b200.engine import VQAEngine, host_cores_per_rank, host_pack_threads
    relation = wl == "regat"
    cfg = O.FULL_REGAT if relation else O.FULL
    W = O.make_weights(cfg, 1111)
    kw = {} if args.overlap is None else {"overlap": bool(args.overlap)}
    if args.chase is not None:
        kw["gat_chase_sms"] = args.chase
    if args.side_sms:
        kw["side_sms"] = args.side_sms
    if args.side_permille:
        kw["side_tile_permille"] = args.side_permille
    eng = VQAEngine(W, relation=relation, precision=args.precision, device=ctx.dev, **kw)
    B, NB, dev, rank = args.batch, 4, ctx.dev, ctx.rank
    batch0 = O.make_batch(cfg, B, 3000 + rank)                 # batch 0 is a full oracle batch (parity below)
    g = torch.Generator(device="cpu").manual_seed(1000 * 2 + rank)
    imgs, toks, labs = [], [], []
    for i in range(NB):
        if i == 0:
            img, tok = batch0["img"], batch0["q"]
        else:
            img = torch.rand((B, 36, 2048), generator=g, dtype=torch.float32)
            tok = torch.randint(0, cfg.ntoken, (B, 14), generator=g)
        imgs.append(eng.resident(img.to(dev)))
        toks.append(tok.to(dev))
        if relation:
            if i == 0:
                labs.append(batch0["graph"].to(torch.uint8).to(dev))
            else:
                boxes = torch.from_numpy(O.make_boxes(B, 36, 50 + i + 10 * rank)).to(dev)
                labs.append(ops.relation_labels(boxes, 640, 480))
    host_img = torch.rand((B, 36, 2048), generator=g, dtype=torch.float32).pin_memory()
    host_tok = torch.randint(0, cfg.ntoken, (B, 14), generator=g).pin_memory()
    host_lab = labs[1].cpu().pin_memory() if relation else None

    def direct(i):
        return eng.forward(imgs[i % NB], toks[i % NB], labels=labs[i % NB] if relation else None)

    # ---- one CUDA graph per resident batch: a step is ONE graph launch (no per-step tensor-map encodes, argument
    # structs, output allocations or 8-12 launches from Python)
    graphs, outs, mode = [], [], "direct"
    out0 = direct(0)
    launches_per_step = eng.last_launches
    if not args.no_graph:
        try:
            for i in range(NB):
                gph, o = eng.capture(imgs[i], toks[i], labels=labs[i] if relation else None)
                graphs.append(gph)
                outs.append(o)
            mode = "cuda_graph"
        except Exception as e:                              # keep measuring: the direct launches are the same kernels
            graphs, outs, mode = [], [], f"direct (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            torch.cuda.synchronize()

    def step(i):
        if graphs:
            graphs[i % NB].replay()
        else:
            direct(i)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier(ctx)
    sampler = ClockSampler(ctx.local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sampler.wait_for(e1)
    barrier(ctx)
    ms_rank = e0.elapsed_time(e1)
    ms_total, per_rank = all_max(ctx, ms_rank)
    value = B * ctx.world * args.steps / (ms_total / 1e3)
    res = {"value": value, "unit": UNIT, "ms_per_step": ms_total / args.steps, "steps": args.steps,
           "per_rank_ms_total": [round(x, 4) for x in per_rank], "launch_mode": mode,
           "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
           "schedule": (f"two streams: graph attention on {eng.gat_chase_sms} SMs chasing the wide projection (reads Y from L2)"
                        if relation and eng.gat_chase_sms > 0 and args.precision == "bf16" else
                        "two streams: question encoder beside the question-independent projection" if eng.overlap and
                        args.precision == "bf16" and B >= 512 else "one stream"),
           "clocks": sampler.summary(), "config": workload_config(args, wl)}

    # ---- parity of batch 0 at the full batch size (rank 0, N = 1: the oracle takes seconds)
    if ctx.world == 1 and not args.no_parity:
        o = outs[0] if outs else out0
        if graphs:
            graphs[0].replay()
        else:
            o = direct(0)
        torch.cuda.synchronize()
        res["parity"] = parity_block(O, cfg, W, eng, batch0, o, relation)

    # ---- e2e: host buffers in the reference wire format (pinned f32 features) through the C-ABI host entry
    # point: host cores pack f32→bf16 chunk by chunk while the previous chunk's DMA is in flight → H2D → forward →
    # answers D2H, all inside the timed region.  Hybrid staging: of every `period` chunks one crosses PCIe as raw f32
    # (device cast), the others are packed; which split wins depends on the box, so the candidates are timed for a few
    # steps first (all ranks agree on the one with the best worst-rank time), then the winner is measured.
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 20))

        def time_host(img_h, period, steps):
            pack = args.precision == "bf16" and period != 1
            kwh = dict(labels_h=host_lab, pack_on_host=pack, raw_chunk_period=period if pack else 0)
            for _ in range(2):
                eng.forward_host(img_h, host_tok, **kwh)
            barrier(ctx)
            e0.record()
            pending = None
            for _ in range(steps):                        # two batches in flight: stage n+1 while n computes
                nxt = eng.forward_host_async(img_h, host_tok, **kwh)
                if pending is not None:
                    pending.result()
                pending = nxt
            _, h2d_, d2h_ = pending.result()
            e1.record()
            barrier(ctx)
            ms, _ = all_max(ctx, e0.elapsed_time(e1))
            return ms, h2d_, d2h_

        candidates = [0, 4, 3, 2, 1] if args.precision == "bf16" else [1]
        trials = {c: time_host(host_img, c, 6)[0] / 6 for c in candidates}
        period = min(trials, key=trials.get)
        ms_e2e, h2d, d2h = time_host(host_img, period, e2e_steps)
        res["e2e"] = {"value": B * ctx.world * e2e_steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                      "host_pack_threads": host_pack_threads(), "host_cores_per_rank": host_cores_per_rank(),
                      "raw_chunk_period": period, "raw_chunk_period_trials_ms": {str(k): round(v, 3) for k, v in trials.items()},
                      "batches_in_flight": 2,
                      "note": "vqa_forward_host_submit/_wait (the call Wrapper.forward_vqa makes for a host batch): pinned f32 "
                              "host features (reference wire format) -> 64-image chunks, one in raw_chunk_period sent as f32 and "
                              "cast on the device, the others packed to bf16 by this rank's share of the host cores (0 = all "
                              "packed, 1 = all raw; picked by a short trial of every candidate), pipelined with H2D -> forward "
                              "-> answers D2H; batch n+1 is staged while batch n computes"}
        if args.precision == "bf16":
            ms_b, h2d_b, _ = time_host(host_img.to(torch.bfloat16).pin_memory(), 1, e2e_steps)
            res["e2e_bf16_host_cache"] = {"value": B * ctx.world * e2e_steps / (ms_b / 1e3), "unit": UNIT,
                                          "h2d_bytes_per_step": h2d_b, "ms_per_step": ms_b / e2e_steps,
                                          "note": "bf16 host feature cache -> H2D -> forward -> answers D2H"}

    # ---- rooflines: every hot kernel of the step timed ALONE with CUDA events on its launch stream (rotating inputs)
    if args.precision == "bf16":
        res["roofline_kernels"] = kernel_rooflines(ctx, args, eng, imgs, toks, labs, relation)
        res["roofline"] = max(res["roofline_kernels"], key=lambda r: r["launch_ms"])
    if args.precision == "fp32tc" and not relation:
        res["roofline_kernels"] = kernel_rooflines_split(ctx, args, eng, imgs, toks)
        res["roofline"] = max(res["roofline_kernels"], key=lambda r: r["launch_ms"])
    res["path_tflops_per_gpu"] = value / ctx.world * FLOPS_PER_Q[wl] / 1e12
    res["path_frac_of_sustained_bf16"] = res["path_tflops_per_gpu"] / peaks()["bf16_tflops_sustained"]
    del graphs, outs
    return res


def kernel_rooflines(ctx, args, eng, imgs, toks, labs, relation):
    ops, P, B, NB, dev = ctx.ops, eng.P, args.batch, len(imgs), ctx.dev
    pk = peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, min(args.steps, 50))

    def time_kernel(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def entry(key, name, bound, ms, work, unit_scale):
        peak = pk["bf16_tflops"] if bound == "tensor" else pk["hbm_gbs"]
        achieved = work / (ms / 1e3) / unit_scale
        traffic, src = NCU_TRAFFIC.get(key, (None, None))
        return {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": achieved / peak,
                "traffic": (traffic * B / 1024 if traffic else None), "traffic_source": src,
                "peak_source": pk["src"] + " burst (kernel timed alone)", "launch_ms": ms,
                ("flops_per_launch" if bound == "tensor" else "bytes_per_launch"): work}

    H, V, K, T, E = P["H"], P["V"], 36, toks[0].shape[1], P["E_pad"]
    qq = torch.rand((B, 2 * H), device=dev)
    out = []
    # question GRU: T steps of [B,E_pad+H] x [3H, E_pad+H] (the embedding gather and the counter memset are separate launches)
    ms = time_kernel(lambda i: ops.gru_last_state(toks[i % NB], P["emb"], P["w_ih"], P["b_ih"], P["w_hh"], P["b_hh"],
                                                  packed=(P["wx_packed"], P["wh_packed"], P["bias_packed"]),
                                                  gi_table=P.get("gi_table") if eng.use_gi_table else None), reps)
    out.append(entry("gru", "gru_pair_kernel (persistent GRU, 14 steps, tcgen05 cta_group::2, token-table form; + memset)"
                     if (eng.use_gi_table and "gi_table" in P) else
                     "gru_pair_kernel (persistent GRU, 14 steps, tcgen05 cta_group::2; + gather + memset launches)",
                     "tensor", ms, 2.0 * B * T * 3 * H * (E + H) - 2.0 * B * 3 * H * H, 1e12))
    ms = time_kernel(lambda i: ops.linear(imgs[i % NB].view(B * K, V), P["Wv"], P["sv"], P["bv"], relu=True, mul=qq,
                                          mul_row_div=K, logit_w=P["wlin"]), reps)
    out.append(entry("wv", "linear_tc_kernel<256,pair> (W_v projection + x Qp + logit reduction, tcgen05 cta_group::2)",
                     "tensor", ms, 2.0 * B * K * H * V, 1e12))
    parts = torch.rand((B * K, 4), device=dev)
    ms = time_kernel(lambda i: ops.attention_pool(parts, 0.0, imgs[i % NB], want_att=True, want_vsum=True), reps)
    out.append(entry("pool", "attention_pool_stream_kernel (softmax over 36 regions + attention-weighted feature sum)",
                     "hbm", ms, float(B * K * V * 2 + B * V * 2 + B * K * 4 * 5), 1e9))
    if relation and "Wg3" in P:
        Y = None

        def wide(i):
            nonlocal Y
            Y = ops.linear(imgs[i % NB].view(B * K, V), P["Wg3"])
        ms = time_kernel(wide, max(5, reps // 4))
        out.append(entry("wide", "linear_tc_kernel<256,pair> (wide ReGAT projection [B*36,2048]x[6144,2048]^T, tcgen05 cta_group::2)",
                         "tensor", ms, 2.0 * B * K * V * P["Wg3"].shape[0], 1e12))
        att = torch.rand((B, K), device=dev)
        ms = time_kernel(lambda i: ops.graph_attention_merged(Y, imgs[i % NB].view(B * K, V), att, labs[i % NB], P["wvec"],
                                                              P["gat_c0"], P["label_bias_lp"], P["num_labels"], K,
                                                              want_out=False, want_vsum=True), max(5, reps // 2))
        out.append(entry("gat", "graph_attention_tc_kernel (relation-masked KxK graph attention, tcgen05; reads Y and x)",
                         "hbm", ms, float(B * K * V * 2 * 4 + B * V * 2), 1e9))
    return out


def kernel_rooflines_split(ctx, args, eng, imgs, toks):
    """fp32tc (VQA_F16X2): the three hot kernels of the Up-Down step timed alone.  Tensor work = THREE tcgen05.mma per
    k-step (hi·hi, hi·lo', lo'·hi) = 3x the algorithmic FLOPs of the layer; both are reported."""
    ops, P, B, NB, dev = ctx.ops, eng.P, args.batch, len(imgs), ctx.dev
    pk = peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(5, min(args.steps, 20))

    def time_kernel(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    H, V, K, T = P["H"], P["V"], 36, toks[0].shape[1]
    qq = torch.rand((B, 2 * H), device=dev)
    out = []
    ms = time_kernel(lambda i: ops.linear_split(imgs[i % NB].view(2, B * K, V), P["Wv"], P["sv"], P["bv"], relu=True, mul=qq,
                                                mul_row_div=K, logit_w=P["wlin"]), reps)
    alg = 2.0 * B * K * H * V
    out.append({"kernel": "linear_tc_kernel<256,pair,split> (W_v projection on fp16 plane pairs + x Qp + logit reduction)",
                "bound": "tensor", "achieved": 3 * alg / (ms / 1e3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": 3 * alg / (ms / 1e3) / 1e12 / pk["bf16_tflops"], "traffic": 321.9e6 * B / 1024,
                "traffic_source": "profiles/r02b_ncu_wv_split.md (315.8 MB read + 6.1 MB written at B = 1024; algorithmic: 302 MB of "
                                  "feature planes + 8 MB of weight planes)",
                "peak_source": pk["src"] + " burst, f16 = bf16 rate (kernel timed alone)", "launch_ms": ms,
                "flops_per_launch": 3 * alg, "algorithmic_flops_per_launch": alg,
                "note": "achieved counts the three MMAs per k-step that the split arithmetic issues"})
    ms = time_kernel(lambda i: ops.gru_last_state_split(toks[i % NB], P["gi_table"], P["w_hh"], P["b_hh"], P.get("wh_packed")), reps)
    alg = 2.0 * B * (T - 1) * 3 * H * H
    out.append({"kernel": "gru_gate_table_kernel + linear_tc_kernel<96,pair,split,gru> (13 recurrent steps in one cooperative launch)",
                "bound": "tensor", "achieved": 3 * alg / (ms / 1e3) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": 3 * alg / (ms / 1e3) / 1e12 / pk["bf16_tflops"], "traffic": 13 * 33.5e6 * B / 1024,
                "traffic_source": "profiles/r02b_ncu_gru_split.md (33.5 MB per step: token-table rows, W_hh planes, state; 13 steps)",
                "peak_source": pk["src"] + " burst (kernels timed alone)", "launch_ms": ms,
                "flops_per_launch": 3 * alg, "algorithmic_flops_per_launch": alg})
    parts = torch.rand((B * K, 4), device=dev)
    ms = time_kernel(lambda i: ops.attention_pool_split(parts, 0.0, imgs[i % NB]), reps)
    byts = float(B * K * V * 4 + B * V * 4 + B * K * 4 * 5)
    out.append({"kernel": "attention_pool_split_kernel (softmax over 36 regions + attention-weighted sum of the plane pair)",
                "bound": "hbm", "achieved": byts / (ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": byts / (ms / 1e3) / 1e9 / pk["hbm_gbs"], "traffic": 308.6e6 * B / 1024,
                "traffic_source": "profiles/r02b_ncu_pool_split.md",
                "peak_source": pk["src"] + " burst (kernel timed alone)", "launch_ms": ms, "bytes_per_launch": byts})
    return out


def run_train_block(ctx, args, global_batch=512, steps=20, warmup=5):
    """BASELINE config 4: Up-Down training step at a FIXED global batch of 512 (strong scaling: 512 / N questions per GPU),
    one step = train.py:103-111 (get_loss -> backward -> clip_grad_norm_ -> Adamax.step -> zero_grad) with the NCCL gradient
    all-reduce in two buckets (training.py).  Also times the fused forward+backward call alone and the all-reduce of the
    flat gradient buffer alone, so the record says what bounds the step at every N."""
    import vqa_collection_b200 as pkg
    from vqa_collection_b200 import optim as fused, training
    from vqa_collection_b200.modules.wrapper import set_model
    from vqa_collection_b200.parallel import shard_batch, average_gradients_
    O, dev, world, rank = ctx.O, ctx.dev, ctx.world, ctx.rank
    cfg = O.FULL
    prev = pkg.get_precision()
    pkg.set_precision(args.precision)
    try:
        m = set_model(encoder_type="base", predictor_type="base", decoder_type="none", ntoken=cfg.ntoken, v_dim=cfg.v_dim,
                      embed_dim=cfg.embed_dim, hidden_dim=cfg.hidden_dim, rnn_layer=1, ans_dim=cfg.ans_dim, cls_layer=2, c_len=20,
                      device=str(dev), dropout=0.2, rnn_type="GRU", att_type="new", conv_layer=1, conv_type="corr")
        m.load_state_dict(O.make_weights(cfg, 1111), strict=True)
        opt = fused.Adamax([{'params': m.encoder.parameters()}, {'params': m.predictor.parameters(), 'lr': 0.002}], lr=0.002)
        b = shard_batch(O.make_batch(cfg, global_batch, 7), world, rank)
        dt = torch.bfloat16 if args.precision == "bf16" else torch.float32
        batch = {"img": b["img"].to(dt).to(dev), "q": b["q"].to(dev), "a": b["a"].float().to(dev)}
        m.train()

        def step():
            loss, _ = m.get_loss(batch)
            loss.backward()
            fused.clip_grad_norm_(m.parameters(), 0.25)
            opt.step()
            opt.zero_grad()
            return loss

        for _ in range(warmup):
            step()
        barrier(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        barrier(ctx)
        ms, per_rank = all_max(ctx, e0.elapsed_time(e1) / steps)
        # the fused forward + loss + backward call (+ its all-reduce buckets when N > 1), no optimizer, no host reads
        e0.record()
        for _ in range(steps):
            l, _ = training.updown_loss(m, batch["img"], batch["q"], batch["a"], seed=1)
            l.backward()
            opt.zero_grad()
        e1.record()
        barrier(ctx)
        ms_core, _ = all_max(ctx, e0.elapsed_time(e1) / steps)
        n_grad = sum(p.numel() for p in m.parameters())
        ar_ms = None
        if ctx.dist is not None:
            flat = torch.zeros((n_grad,), dtype=torch.float32, device=dev)
            for _ in range(3):
                average_gradients_(flat)
            barrier(ctx)
            e0.record()
            for _ in range(10):
                average_gradients_(flat)
            e1.record()
            barrier(ctx)
            ar_ms, _ = all_max(ctx, e0.elapsed_time(e1) / 10)
        return {"metric": "Up-Down VQA training questions/sec (global batch %d)" % global_batch,
                "value": global_batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "per_rank_ms_per_step": [round(x, 4) for x in per_rank],
                "ms_fwd_loss_bwd_allreduce": ms_core, "allreduce_alone_ms": ar_ms, "allreduce_bytes": 4 * n_grad,
                "allreduce": "ncclAllReduce AVG over NVLink, two buckets: the 7 weight-normed layers under the BPTT, then GRU + embedding",
                "scaling": "strong", "steps": steps, "loss": float(loss), "dtype": args.precision,
                "config": {"workload": "Up-Down VQA training step batch 512 with NCCL gradient allreduce",
                           "global_batch": global_batch, "per_gpu_batch": int(batch["img"].shape[0]), "parallelism": f"dp{world}",
                           "optimizer": "vqa_collection_b200.optim.Adamax + clip_grad_norm_(0.25), one launch each (train.py:108-111)"}}
    finally:
        pkg.set_precision(prev)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.rank, ctx.local_rank, ctx.world = rank, local_rank, world
    ctx.dev = torch.device("cuda", local_rank)
    ctx.dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=ctx.dev)
        ctx.dist = dist

    from oracle import vqa_oracle as O          # synthetic weights/inputs, the parity checker and the cpu_baseline leg only
    from vqa_collection_b200 import ops
    ctx.O, ctx.ops = O, ops

    results = {wl: run_workload(ctx, args, wl) for wl in args.workloads}
    # the same workloads in the fp32-class tensor-core mode (precision 'fp32tc'): the mode whose answers are bit-exact
    exact = None
    if args.precision == "bf16" and not args.no_exact_block:
        import copy
        a2 = copy.copy(args)
        a2.precision, a2.no_e2e, a2.steps = "fp32tc", True, max(3, min(args.steps, 20))
        exact = {}
        for wl in args.workloads:
            try:
                r2 = run_workload(ctx, a2, wl)
                exact[wl] = {k: r2[k] for k in ("value", "unit", "ms_per_step", "steps", "per_rank_ms_total", "launch_mode",
                                                "launches_per_step", "parity", "clocks", "roofline_kernels") if k in r2}
            except Exception as e:                          # never lose the bf16 line to this block
                exact[wl] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    train = None
    if args.train_block:
        try:
            train = run_train_block(ctx, args)
        except Exception as e:                              # never lose the forward line to the training block
            train = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return
    first = args.workloads[0]
    r = results[first]
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # the same function, model and full batch as `--impl reference`, bounded to ~10-30 s of CPU work
        cpu = {}
        for wl in args.workloads:
            c = cpu_forward_qps(wl, args.batch, 3 if wl == "updown" else 1, warmup=1)
            cpu[wl] = {"value": c["qps"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "best_of": c["best"],
                       "cpu_model": cpu_model(), "sample": c["sample"]}
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "fp32": "f32", "fp32tc": "f16x2 (fp16 plane pairs, fp32-class)"}[args.precision],
        "data": "synthetic", "config": r["config"],
    }
    for k in ("clocks", "e2e", "e2e_bf16_host_cache", "gpu_launches", "launches_per_step", "launch_mode", "schedule",
              "per_rank_ms_total", "parity", "roofline", "roofline_kernels", "path_tflops_per_gpu",
              "path_frac_of_sustained_bf16"):
        if k in r:
            line[k] = r[k]
    line["cpu_baseline"] = cpu[first] if cpu else None
    for wl in args.workloads[1:]:
        blk = dict(results[wl])
        blk["cpu_baseline"] = cpu[wl] if cpu else None
        line[wl] = blk
    if exact is not None:
        exact["note"] = ("precision 'fp32tc': every GEMM operand as an fp16 plane pair (hi + lo'·2^-11), three tcgen05.mma per "
                         "k-step, fp32 accumulators in TMEM; device-resident inputs (plane pairs), CUDA-graph steps, same "
                         "batches and timing rules as `value`; parity = the fp32 gates (1e-5 logits / attention, answers bit-exact)")
        line["fp32tc"] = exact
    if train is not None:
        line["train"] = train
    print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
