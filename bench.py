#!/usr/bin/env python
"""bench.py — VQA forward questions/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload updown|regat]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

A step = one forward pass (question encoder → top-down attention → [ReGAT] → classifier →
answers) over one batch of 1024 synthetic questions per GPU (36x2048 region features, 14
tokens, 3129 answers), inputs resident in HBM as bf16.  Batch is sharded data-parallel
(weak scaling: 1024 questions per GPU, no forward collective).  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "VQA forward questions/sec (36x2048 regions, bs=1024)"
UNIT = "questions/s"
FLOPS_PER_Q = {"updown": 290_500_608, "regat": 1_214_553_216}       # SURVEY.md §8d (minimal algebra)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="updown", choices=["updown", "regat"])
    ap.add_argument("--batch", type=int, default=1024, help="questions per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """samples SM clock / throttle reasons with NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self._stop, self.ok = [], set(), threading.Event(), False
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

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(0.005)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join()

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


def cpu_forward_qps(workload, sample_b, iters, warmup=1, threads=None):
    """the reference's CPU path (oracle port: same torch-CPU ops, op for op) on all host cores"""
    from oracle import vqa_oracle as O
    cfg = O.FULL_REGAT if workload == "regat" else O.FULL
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = O.make_weights(cfg, 1111)
    batch = O.make_batch(cfg, sample_b, 2000)
    times = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            O.forward_vqa(batch, W, cfg)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sample_b / (sum(times) / len(times)), times, cores


def run_reference(args, rank):
    if rank != 0:
        return
    sample_b = 128 if args.workload == "updown" else 32
    qps, times, cores = cpu_forward_qps(args.workload, sample_b, max(args.steps, 1), max(args.warmup, 1))
    ms = 1e3 * sum(times) / len(times)
    sample = (f"{sample_b} of the {args.batch} questions per step, fp32, torch-CPU oracle port of "
              f"Wrapper.forward_vqa, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_b),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_step_b=None):
    name = ("Up-Down VQA forward bf16 batch 1024 on 1xB200 (tcgen05 projections + fused attention/softmax)"
            if args.workload == "updown" else
            "ReGAT spatial-relation VQA forward (11 relation labels, KxK masked graph attention) batch 1024")
    return {"workload": name, "batch_per_gpu": args.batch, "regions": 36, "v_dim": 2048, "hidden": 1024,
            "tokens": 14, "answers": 3129, "ntoken": 20000, "parallelism": f"dp{args.gpus}",
            "l2": "4 rotating resident batches (151 MB bf16 features each > 126 MB L2)",
            **({"sample_per_step": per_step_b} if per_step_b else {})}


# ----------------------------------------------------------------------------- B200 arm
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
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from oracle import vqa_oracle as O          # only for synthetic weights/inputs + the cpu_baseline leg
    from vqa_collection_b200.engine import VQAEngine
    from vqa_collection_b200 import ops

    relation = args.workload == "regat"
    cfg = O.FULL_REGAT if relation else O.FULL
    W = O.make_weights(cfg, 1111)
    eng = VQAEngine(W, relation=relation, precision=args.precision, device=dev)
    B, NB = args.batch, 4
    g = torch.Generator(device="cpu").manual_seed(1000 * 2 + rank)
    imgs, toks, labs = [], [], []
    for i in range(NB):
        img = torch.rand((B, 36, 2048), generator=g, dtype=torch.float32)
        imgs.append(eng.resident(img.to(dev)))
        toks.append(torch.randint(0, cfg.ntoken, (B, 14), generator=g).to(dev))
        if relation:
            boxes = torch.from_numpy(O.make_boxes(B, 36, 50 + i + 10 * rank)).to(dev)
            labs.append(ops.relation_labels(boxes, 640, 480))
    host_img = torch.rand((B, 36, 2048), generator=g, dtype=torch.float32).pin_memory()
    host_tok = torch.randint(0, cfg.ntoken, (B, 14), generator=g).pin_memory()
    host_lab = labs[0].cpu().pin_memory() if relation else None

    def step(i):
        return eng.forward(imgs[i % NB], toks[i % NB], labels=labs[i % NB] if relation else None)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        sampler.sample()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.last_launches * args.steps
    if dist is not None:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = B * world * args.steps / (ms_total / 1e3)

    # ---- e2e: host buffers in the reference wire format (pinned f32 features) through the C-ABI host entry
    # point vqa_forward_host: host cores pack f32→bf16 chunk by chunk while the previous chunk's DMA is in
    # flight → H2D → forward → answers D2H, all inside the timed region
    e2e_steps = max(3, min(args.steps, 20))

    # hybrid staging: of every `period` chunks one crosses PCIe as raw f32 (device cast), the others are packed to bf16 by
    # the host cores this rank may count on (all cores / ranks on the box); 0 = pack everything, 1 = everything raw.
    # Which split wins depends on the box (cores and memory bandwidth per GPU, ranks sharing them), so the candidates are
    # timed for a few steps first (all ranks agree on the one with the best worst-rank time), then the winner is measured.
    from vqa_collection_b200.engine import host_cores_per_rank, host_pack_threads

    def time_host(period, steps):
        pack = args.precision == "bf16" and period != 1
        kw = dict(labels_h=host_lab, pack_on_host=pack, raw_chunk_period=period if pack else 0)
        for _ in range(2):
            eng.forward_host(host_img, host_tok, **kw)
        barrier()
        e0.record()
        pending = None
        for _ in range(steps):                        # two batches in flight: stage n+1 while n computes
            nxt = eng.forward_host_async(host_img, host_tok, **kw)
            if pending is not None:
                pending.result()
            pending = nxt
        _, h2d_, d2h_ = pending.result()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, h2d_, d2h_

    candidates = [0, 4, 3, 2, 1] if args.precision == "bf16" else [1]
    trials = {c: time_host(c, 6)[0] / 6 for c in candidates}
    RAW_PERIOD = min(trials, key=trials.get)
    ms_e2e, h2d, d2h = time_host(RAW_PERIOD, e2e_steps)
    e2e_value = B * world * e2e_steps / (ms_e2e / 1e3)
    e2e_alt = None
    if args.precision == "bf16":
        ms_alt, h2d_alt, _ = time_host(1, e2e_steps)
        e2e_alt = {"value": B * world * e2e_steps / (ms_alt / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_alt,
                   "ms_per_step": ms_alt / e2e_steps, "note": "f32 features over PCIe + device cast (no host packing)"}

    # a host-side feature cache kept in the resident bf16 format: what the same call does when the
    # caller stores its features as bf16 — reported next to the headline, which keeps the reference's f32 wire format
    e2e_bf16 = None
    if args.precision == "bf16":
        host_img_bf16 = host_img.to(torch.bfloat16).pin_memory()
        saved = host_img
        host_img = host_img_bf16
        ms_b, h2d_b, _ = time_host(1, e2e_steps)
        host_img = saved
        e2e_bf16 = {"value": B * world * e2e_steps / (ms_b / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_b,
                    "ms_per_step": ms_b / e2e_steps, "note": "bf16 host feature cache -> H2D -> forward -> answers D2H"}

    # ---- roofline of the dominant kernel (W_v projection fused with the attention logits),
    # timed alone with CUDA events on its launch stream
    P = eng.P
    pk = peaks()
    reps = max(10, min(args.steps, 100))
    qq = torch.rand((B, 2 * P["H"]), device=dev)

    def wv(i):
        return ops.linear(imgs[i % NB].view(B * 36, 2048), P["Wv"], P["sv"], P["bv"], relu=True, mul=qq,
                          mul_row_div=36, logit_w=P["wlin"])

    def wide(i):
        return ops.linear(imgs[i % NB].view(B * 36, 2048), P["Wg3"])

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

    roof = None
    if args.precision == "bf16":
        # dominant kernel of the workload: Up-Down = the W_v projection fused with the attention logits; ReGAT = the wide
        # projection x·[W0+W1; W2; WbᵀWa]ᵀ (55 % of the step).  traffic = dram__bytes_read.sum + dram__bytes_write.sum per
        # launch from `ncu --set full` (profiles/r01e_ncu_wv.md: 159.48 + 6.71 MB, algorithmic 159.0 MB;
        # profiles/r01e_ncu_wide.md: 181.17 + 412.90 MB, algorithmic 151 + 25 + 453 MB — part of Y is still in L2 at kernel end)
        if relation and "Wg3" in P:
            k_ms = time_kernel(wide, max(5, reps // 5))
            flops = 2.0 * B * 36 * P["V"] * P["Wg3"].shape[0]
            name, traffic, src = ("linear_tc_kernel<256,pair> (wide ReGAT projection [B*36,2048]x[6144,2048]^T, tcgen05 cta_group::2)",
                                  594.07e6, "profiles/r01e_ncu_wide.md")
        else:
            k_ms = time_kernel(wv, reps)
            flops = 2.0 * B * 36 * P["H"] * P["V"]
            name, traffic, src = ("linear_tc_kernel<256,pair> (W_v projection + logit reduction, tcgen05 cta_group::2)",
                                  166.2e6, "profiles/r01e_ncu_wv.md")
        achieved = flops / (k_ms / 1e3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"], "traffic": traffic * B / 1024, "traffic_source": src,
                "peak_source": pk["src"] + " burst (kernel timed alone)", "launch_ms": k_ms, "flops_per_launch": flops}
    path_tflops = value / world * FLOPS_PER_Q[args.workload] / 1e12

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample_b = 64 if not relation else 16
        qps, times, cores = cpu_forward_qps(args.workload, sample_b, 5 if not relation else 3)
        best = sample_b / min(times)
        qps1, _, _ = cpu_forward_qps(args.workload, sample_b if not relation else 8, 1, warmup=1, threads=1)
        torch.set_num_threads(cores)
        cpu = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "best_of": best, "value_1_thread": qps1,
               "cpu_model": cpu_model(),
               "sample": f"{sample_b} questions x {len(times)} passes (mean; best_of = fastest pass), fp32 torch-CPU oracle "
                         f"port of Wrapper.forward_vqa, {cores} threads"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision if args.precision == "bf16" else "f32",
        "data": "synthetic", "config": workload_config(args),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "host_pack_threads": host_pack_threads(), "host_cores_per_rank": host_cores_per_rank(),
                "raw_chunk_period": RAW_PERIOD, "raw_chunk_period_trials_ms": {str(k): round(v, 3) for k, v in trials.items()},
                "batches_in_flight": 2,
                "note": "vqa_forward_host_submit/_wait: pinned f32 host features (reference wire format) -> 64-image chunks, "
                        "one in raw_chunk_period sent as f32 and cast on the device, the others packed to bf16 by this rank's "
                        "share of the host cores (0 = all packed, 1 = all raw; the split is picked by a short trial run of every "
                        "candidate), all pipelined with H2D -> forward -> answers D2H; batch n+1 is staged while batch n computes"},
        "e2e_f32_over_pcie": e2e_alt,
        "e2e_bf16_host_cache": e2e_bf16,
        "gpu_launches": launches,
        "roofline": roof,
        "path_tflops_per_gpu": path_tflops,
        "path_frac_of_sustained_bf16": path_tflops / pk["bf16_tflops_sustained"],
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
