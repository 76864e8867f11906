#!/usr/bin/env python
"""bench.py -- encoder-forward utterances/s (30 s @ 16 kHz windows), whisper-large-v3-turbo + FDDT (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one DiCoWEncoder forward over one batch of B=32 synthetic utterances per GPU (BASELINE configs[1]):
log-mel-shaped input_features [B, 128, 3000] fp32 + soft STNO masks [B, 4, 1500], random-init weights of the
large-v3-turbo architecture with FDDT parameters perturbed off their identity init.  Utterances are independent, so
N GPUs run N replicas with no data-path collective ("weak" scaling: B per GPU fixed).

  value  : utt/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e    : utt/s through the public API with pinned HOST buffers: H2D of features + masks and D2H of
           last_hidden_state inside the timed region, every step
  roofline: the tcgen05 GEMM kernel family (dominant: ~60 % of the step), algorithmic FLOPs / CUDA-event time per
           launch measured live in extra instrumented steps right after the timed region (an event pair around every
           launch perturbs the step by up to 6 %), against the measured cuBLAS bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the REFERENCE's own DiCoWEncoder.forward (oracle/_ref: its src/models/dicow modules,
           copied there unmodified by oracle/make_ref.py in the build container; kind "reference") in fp32 on all host
           cores, on a bounded sample (B=1 forwards) of the same workload; where oracle/_ref is missing, the oracle port
           (oracle/dicow_oracle.py, kind "port").
  secondary: BASELINE configs[2], [3], [4] measured in the same process after the headline (tools/workloads.py): the
           fine-tune step with the gradient all-reduce (under torchrun: NCCL over NVLink, exposed communication time
           reported), the CTC pre-train step, SE-DiCoW greedy decode -- each with its own bracketed clock sample and
           roofline fraction -- and the long-form seek loop with / without the speculative next-window encoder pass
           (SURVEY 8(f).3).  --secondary none skips them.
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

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encoder-fwd utterances/sec (30s@16kHz) large-v3-turbo+FDDT"
GFLOP_PER_UTT = 2273.8  # conv stem 17.7 + 32 x (QKVO 19.661 + QK^T/PV 11.520 + MLP 39.322)


from tools.workloads import make_inputs, perturb_, turbo_config  # noqa: E402,F401  (re-exported for tools/)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self) -> int:
        """number of samples taken so far (brackets the timed region: the sampler is started before the warm-up so that
        nvidia-smi's own start-up -- process launch, NVML initialisation -- does not land inside it)"""
        return len(self.rows)

    def window(self, first: int = 0, last: int = None):
        """summary of the samples [first, last) (marks taken around a timed region)"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[first:max(last, first + 1) if last is not None else None] or self.rows[-3:]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self, first: int = 0, last: int = None):
        out = self.window(first, last)
        if self.proc is not None:
            self.proc.terminate()
        return out


def time_reference(steps, warmup, B=1, threads=None, device="cpu", force_port=False):
    """The reference's implementation of the path on the host cores: its own DiCoWEncoder.forward from oracle/_ref (kind
    "reference") or, where that copy is missing, the oracle port (kind "port"); fp32, SDPA, all host threads.
    ``--reference-device cuda`` (context only): the same eager code under bf16 autocast on the GPU (cuBLAS + SDPA flash)."""
    from oracle import ref_loader
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    feats, stno = make_inputs(B, 100)
    dev = torch.device(device)
    if ref_loader.available() and not force_port:
        kind = "reference"
        RefConfig, RefEncoder = ref_loader.load()
        cfg = RefConfig(**turbo_config().to_dict())
        cfg._attn_implementation = "sdpa"
        torch.manual_seed(7)
        enc = RefEncoder(cfg).eval()
        perturb_(enc, "cpu")
        enc = enc.to(dev)

        def fwd(f, s):
            return enc(f, stno_mask=s).last_hidden_state
    else:
        kind = "port"
        from oracle import dicow_oracle as orc
        from oracle import synth
        dm = synth.LARGE_V3_TURBO
        g = torch.Generator().manual_seed(7)
        p = {}
        for k, shp in synth.param_shapes(dm, decoder=False).items():
            if "lm_head" in k or "subsample" in k or "additional" in k:
                continue
            if k.endswith("weight") and len(shp) >= 2:
                p[k] = torch.randn(shp, generator=g) * (1.0 / (shp[1] * (shp[2] if len(shp) > 2 else 1)) ** 0.5)
            elif "fddt" in k and k.endswith("weight"):
                p[k] = torch.rand(shp, generator=g) + 0.5
            elif "layer_norm.weight" in k:
                p[k] = torch.rand(shp, generator=g) * 0.4 + 0.8
            else:
                p[k] = torch.randn(shp, generator=g) * 0.1
        p = {k: v.to(dev) for k, v in p.items()}

        def fwd(f, s):
            return orc.encoder_forward(p, dm, f, s)
    feats, stno = feats.to(dev), stno.to(dev)
    times = []
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=device != "cpu"):
        for i in range(warmup + steps):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fwd(feats, stno)
            if device != "cpu":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    total = sum(times)
    how = "fp32" if device == "cpu" else f"bf16 autocast on {device} (eager cuBLAS / SDPA, context only)"
    what = ("the reference's own DiCoWEncoder.forward (oracle/_ref)" if kind == "reference" else
            "oracle port of the reference's algorithm (oracle/dicow_oracle.py)")
    return {"utt_per_s": B * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": threads, "kind": kind,
            "sample": f"{len(times)} forward(s) of B={B} utterance(s) of the same workload, {what}, {how}, torch "
                      f"{torch.__version__} SDPA, {warmup} warm-up"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per step (BASELINE configs[1]: 32)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", default="all", help="comma list of finetune_step,ctc_pretrain_step,se_dicow_greedy; "
                    "'all' (default) or 'none'")
    ap.add_argument("--reference-device", default="cpu", help="--impl reference: 'cpu' (the baseline) or 'cuda' (the same "
                    "stock PyTorch code under bf16 autocast on the GPU, for context)")
    ap.add_argument("--reference-port", action="store_true", help="--impl reference: time the oracle port even if oracle/_ref exists")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "whisper-large-v3-turbo + FDDT encoder forward, batch=32 synthetic 30s per B200 "
                          "(BASELINE configs[1])", "batch_per_gpu": args.batch, "mel_bins": 128, "frames": 3000,
              "layers": 32, "d_model": 1280, "parallelism": f"replicas x{world}, no data-path collective",
              "l2": "per-step working set (2.5 GB weights+activations per layer pass) >> 126 MB L2; 3 input batches rotate"}

    if args.impl == "reference":
        if rank != 0:
            return
        on_gpu = args.reference_device != "cpu"
        # the caller's step / warm-up counts, bounded so that the run ends within a few minutes on the host cores
        # (one step = one B=1 forward of the same workload, ~1.2-1.5 s on the GPU box's 16 threads)
        K, W = max(1, min(args.steps, 40)), max(1, min(args.warmup, 5))
        r = time_reference(K, W, B=args.batch if on_gpu else 1, device=args.reference_device, force_port=args.reference_port)
        line = {"impl": "reference", "metric": METRIC, "value": r["utt_per_s"], "unit": "utt/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if on_gpu else "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["utt_per_s"], "unit": "utt/s", "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["utt_per_s"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist
    from ts_asr_whisper_b200 import ops, parallel
    from ts_asr_whisper_b200.modeling import DiCoWEncoder
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    parallel.init_process_group("nccl", dev)  # no-op for a single process

    with torch.device(dev):
        enc = DiCoWEncoder(turbo_config())
    perturb_(enc, dev)
    enc.eval()
    B = args.batch
    host = [make_inputs(B, 10 + i, pin=True) for i in range(3)]
    resident = [(f.to(dev), s.to(dev)) for f, s in host]
    out_host = torch.empty(B, 1500, 1280, dtype=torch.float32).pin_memory()
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4
    d2h = out_host.numel() * 4

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        enc(resident[i % 3][0], stno_mask=resident[i % 3][1])
    torch.cuda.synchronize()
    time.sleep(0.3)  # let nvidia-smi finish starting up before the timed region
    # ---- timed: device-resident inputs ----
    m0 = sampler.mark()
    ops.timing_log = None
    l0 = ops.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        enc(resident[i % 3][0], stno_mask=resident[i % 3][1])
    e1.record()
    barrier()
    launches = ops.launch_count - l0
    ms = e0.elapsed_time(e1)
    m1 = sampler.mark()
    # ---- per-kernel durations for the roofline: the same steps again with a CUDA event pair around every launch.  Kept
    # out of the region `value` is timed over: 458 event records per step open gaps between the kernels (measured on one
    # box: 79.6 ms/step without them, 84.7 ms with them)
    roofline_steps = min(args.steps, 3)
    ops.timing_log = []
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(roofline_steps):
        enc(resident[i % 3][0], stno_mask=resident[i % 3][1])
    e3.record()
    barrier()
    ms_instr = e2.elapsed_time(e3)
    log, ops.timing_log = ops.timing_log, None
    # ---- timed: end to end through the public API with host buffers ----
    # Every step copies its inputs from pinned host memory and its result back (inside the timed region); the copies run
    # on two side streams, double-buffered, so that step i + 1's H2D and step i - 1's D2H overlap step i's kernels -- what
    # a caller feeding the encoder from a DataLoader does (pin_memory + non_blocking).
    cur = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]

    def e2e_pass(n):
        ev_in = [None, None]
        ev_free = [None, None]  # compute that last read input buffer k has finished
        ev_out = [None, None]   # D2H out of host buffer k has finished

        def h2d_copy(i):
            k = i % 2
            with torch.cuda.stream(s_in):
                if ev_free[k] is not None:
                    s_in.wait_event(ev_free[k])
                dev_in[k][0].copy_(host[i % 3][0], non_blocking=True)
                dev_in[k][1].copy_(host[i % 3][1], non_blocking=True)
                ev_in[k] = s_in.record_event()

        h2d_copy(0)
        for i in range(n):
            k = i % 2
            cur.wait_event(ev_in[k])
            o = enc(dev_in[k][0], stno_mask=dev_in[k][1]).last_hidden_state
            ev_free[k] = cur.record_event()
            if i + 1 < n:
                h2d_copy(i + 1)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_free[k])
                if ev_out[k] is not None:
                    s_out.wait_event(ev_out[k])
                out_hosts[k].copy_(o, non_blocking=True)
                o.record_stream(s_out)
                ev_out[k] = s_out.record_event()
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)

    e2e_pass(2)
    barrier()
    e0.record()
    e2e_pass(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.window(m0, max(m1, m0 + 1)) if rank == 0 else None  # samples taken during the resident timed region
    ms, ms_e2e = parallel.max_over_ranks([ms, ms_e2e], dev)  # multi-GPU numbers are the slowest rank's

    # ---- secondary workloads (BASELINE configs[2], [4], [3]); every rank takes part, rank 0 reports ----
    del enc, resident, dev_in, out_hosts, out_host, host
    torch.cuda.empty_cache()
    names = {"all": ["finetune_step", "ctc_pretrain_step", "se_dicow_greedy", "longform_speculation"], "none": []}.get(
        args.secondary, [n for n in args.secondary.split(",") if n])
    secondary = {}
    from tools import workloads
    smp = sampler if rank == 0 else None
    for name in names:
        try:
            if name == "finetune_step":
                secondary[name] = workloads.train_step("finetune", dev, rank, world, steps=5, warmup=3, sampler=smp)
            elif name == "ctc_pretrain_step":
                secondary[name] = workloads.train_step("ctc_pretrain", dev, rank, world, steps=5, warmup=3, sampler=smp)
            elif name == "se_dicow_greedy":
                secondary[name] = workloads.se_dicow_greedy(dev, rank, world, sampler=smp)
            elif name == "longform_speculation":
                secondary[name] = workloads.longform_speculation(dev, rank, world, sampler=smp)
            else:
                raise SystemExit(f"unknown secondary workload {name}")
        except Exception as exc:  # a secondary failure must not take the headline line down; it is reported, not hidden
            if world > 1:
                raise
            secondary[name] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    from tools.workloads import peaks as read_peaks
    peaks = read_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    hbm = peaks.get("hbm_gbs") or 6650.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
        "fallback B200_PROFILING.md sustained"
    by_kind = {}
    for kind, fl, a, b in log:
        d = by_kind.setdefault(kind, [0.0, 0.0, 0])
        d[0] += fl
        d[1] += a.elapsed_time(b)
        d[2] += 1
    step_ms = ms / args.steps
    others = []
    for k, v in by_kind.items():
        if k == "gemm":
            continue
        o = {"kernel": k, "launches_per_step": v[2] // max(1, roofline_steps), "ms_per_step": v[1] / roofline_steps,
             "share_of_step": (v[1] / roofline_steps) / step_ms if step_ms else None}
        if v[0]:
            tf = v[0] / (v[1] * 1e-3) / 1e12
            o.update({"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf})
        elif k in ("fddt_ln", "layernorm"):
            # SURVEY 8d / DESIGN section 4: LN1 14 B/element + LN2 8 B/element per layer + final LN
            nbytes = B * 1500 * 1280 * (32 * (14 + 8) + 12) * roofline_steps
            gbs = nbytes / (v[1] * 1e-3) / 1e9
            o.update({"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm})
        others.append(o)
    gemm = by_kind.get("gemm", [0.0, 1.0, 1])
    achieved = gemm[0] / (gemm[1] * 1e-3) / 1e12
    traffic, traffic_src = None, None
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f)["dram_bytes_per_launch_avg"]
            traffic_src = f"committed ncu --set full capture (profiles/{name}); NOT measured by this run"
            break
        except (OSError, KeyError, ValueError):
            pass
    utts = world * B * args.steps
    value = utts / (ms * 1e-3)
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_2cta_kernel<256,*> (tcgen05 GEMM family: QKV/out/fc1/fc2/conv)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "launches_per_step": gemm[2] // max(1, roofline_steps),
                "gflop_per_launch_avg": gemm[0] / max(1, gemm[2]) / 1e9,
                "us_per_launch_avg": 1e3 * gemm[1] / max(1, gemm[2]),
                "share_of_step": (gemm[1] / roofline_steps) / step_ms if ms else None,
                "measured_over_steps": roofline_steps, "instrumented_ms_per_step": ms_instr / roofline_steps,
                "note": "per-launch CUDA-event times from extra instrumented steps after the timed region",
                "others": others,
                "whole_step": {"tflops_per_gpu": value / world * GFLOP_PER_UTT / 1e3,
                               "frac_of_sustained_peak": value / world * GFLOP_PER_UTT / 1e3 / peak_tf}}
    line = {"metric": METRIC, "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": utts / (ms_e2e * 1e-3), "unit": "utt/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "secondary": secondary}
    if world == 1 and not args.no_cpu_baseline:
        r = time_reference(2, 1, B=1)
        line["cpu_baseline"] = {"value": r["utt_per_s"], "unit": "utt/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
