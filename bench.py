#!/usr/bin/env python
"""Benchmark of record: frames/s of the 224x224 CNN + bi-GRU event detector forward (BASELINE.json configs[1]:
64 clips x 32 frames per GPU, DenseNet-121 features -> BiGRU(128) -> max over time -> Dense(11)).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)

ours      : one step = one forward of the whole batch through libtennis_b200.so.  `value` times the step with the
            clips already resident in HBM; `e2e` times the same model through the public host-facing call
            (pinned host clips -> H2D -> forward -> logits D2H).  N>1: weak scaling, 64 clips per GPU, frames sharded
            by rank, one NCCL all-gather of per-frame features before the temporal head.
reference : the reference's own implementation is MXNet/Gluon on the CPU (evaluate.py --num_gpus 0); MXNet cannot be
            installed offline, so this arm times the CPU oracle port (oracle/, torch fp32, all host threads) on a
            bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS_PER_GPU, T, SIZE, CLASSES, HIDDEN = 64, 32, 224, 11, 128
ARCH = "densenet121"
FLOP_PER_FRAME = 5.666e9  # conv layers of DenseNet-121 @224^2, SURVEY.md §8d / BASELINE.md §3
RESNET_FLOP_PER_FRAME = 3.627e9  # ResNet-18 v2 @224^2
TRAIN_GLOBAL_CLIPS = 256  # BASELINE.json configs[2]: 256 clips sharded over the GPUs of the run (strong scaling)


def kernel_source_hash():
    """sha256 over the CUDA sources of the inference CNN (everything a DenseNet-121 forward launches): profiles/conv_traffic.json
    is only valid for the kernels it was measured on."""
    import hashlib
    d = os.path.join(ROOT, "tennis_b200", "csrc")
    h = hashlib.sha256()
    path_sources = ("tn_backbone.cu", "tn_common.cu", "tn_common.h", "tn_conv1x1_ts.cu", "tn_conv1x1_ts.h", "tn_conv3x3.cu",
                    "tn_conv3x3.h", "tn_conv_gemm.cu", "tn_conv_gemm.h", "tn_dense_fused.cu", "tn_dense_fused.h", "tn_elementwise.cu",
                    "tn_elementwise.h", "tn_precise.cu", "tn_precise.h", "tn_ptx.cuh", "tn_stem.cu", "tn_stem.h")
    for f in sorted(os.listdir(d)):
        if f in path_sources:
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def conv_traffic_per_frame():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per frame of the conv-kernel family, measured by
    `python tools/measure_traffic.py` (an ncu launch list of one 2048-frame forward) and committed as profiles/conv_traffic.json
    together with the hash of the kernel sources it was taken on.  A stale file (sources changed since) yields None."""
    path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if not os.path.exists(path):
        return None, "profiles/conv_traffic.json missing: run tools/measure_traffic.py under ncu"
    with open(path) as f:
        d = json.load(f)
    if d.get("kernel_source_hash") != kernel_source_hash():
        return None, "profiles/conv_traffic.json is stale (kernel sources changed since it was measured): re-run tools/measure_traffic.py"
    return float(d["conv_dram_bytes_per_frame"]), "measured by ncu on %d frames (%s), tools/measure_traffic.py" % (d["frames"], d["when"])
WORKLOAD = "configs[1]: CNN+GRU event detector fwd, %d clips x %d frames @%dx%d per GPU" % (CLIPS_PER_GPU, T, SIZE, SIZE)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def build_oracle_model():
    import torch
    from oracle import vision as O
    p = O.synthetic_params(ARCH, seed=1234)
    rp = O.synthetic_rnn_params("gru", 1024, HIDDEN, seed=4321)
    g = torch.Generator().manual_seed(77)
    cw = (torch.rand(CLASSES, 2 * HIDDEN, generator=g) * 2 - 1) * 0.07
    cb = torch.zeros(CLASSES)
    return O, p, rp, cw, cb


def oracle_forward(O, p, rp, cw, cb, clips):
    import torch
    with torch.no_grad():
        return O.cnnrnn(clips, lambda x: O.FEATURES[ARCH](x, p), rp, "gru", HIDDEN, cw, cb)


def time_cpu_port(budget_s, clips_per_step=1, steps=None, warmup=1):
    """Times the CPU oracle port on a bounded sample: `clips_per_step` 32-frame clips per step."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, p, rp, cw, cb = build_oracle_model()
    g = torch.Generator().manual_seed(100)
    clips = torch.randn(clips_per_step, T, 3, SIZE, SIZE, generator=g)
    for _ in range(warmup):
        oracle_forward(O, p, rp, cw, cb, clips)
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        oracle_forward(O, p, rp, cw, cb, clips)
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() - t_start) >= budget_s:
            break
    total = sum(times)
    return {"frames_per_s": clips_per_step * T * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "steps": len(times), "cores": torch.get_num_threads(), "clips_per_step": clips_per_step}


def run_reference(args, rank):
    if rank != 0:
        return
    r = time_cpu_port(0.0, clips_per_step=2, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    sample = "%d steps x %d clips (%d frames) of the %s workload; CPU oracle port (torch fp32), MXNet not installable offline" % (
        r["steps"], r["clips_per_step"], r["clips_per_step"] * T, ARCH)
    line = {
        "impl": "reference", "metric": "frames/sec (224x224 CNN+GRU fwd)", "value": r["frames_per_s"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arch": ARCH, "sample": sample},
        "cpu_baseline": {"value": r["frames_per_s"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
def build_model(device):
    """CNNRNN(FrameModel(DenseNet121.features, 11), 11, hidden 128, 'gru') with the seeded synthetic weights the
    oracle uses (train.py:204-236 assembly)."""
    import torch
    from tennis_b200 import model_zoo
    from tennis_b200 import synthetic as O  # seeded weights shared with the parity tests (no oracle in this arm)
    from tennis_b200.models.vision.definitions import CNNRNN, FrameModel
    backbone = model_zoo.get_model("DenseNet121", pretrained=False).features
    model = CNNRNN(FrameModel(backbone, CLASSES), CLASSES, hidden_size=HIDDEN, type="gru")
    model.initialize(ctx=device)
    p = O.synthetic_params(ARCH, seed=1234)
    for k, v in p.items():
        model.td.model._reg_params[k].set_data(v)
    rp = O.synthetic_rnn_params("gru", 1024, HIDDEN, seed=4321)
    for k, v in rp.items():
        prm = model.rnn._reg_params[k]
        prm.shape = tuple(v.shape)
        prm._data = v.to(device).contiguous()
        prm._version += 1
    g = torch.Generator().manual_seed(77)
    model.classes.weight.shape = (CLASSES, 2 * HIDDEN)
    model.classes.weight._data = ((torch.rand(CLASSES, 2 * HIDDEN, generator=g) * 2 - 1) * 0.07).to(device)
    model.classes.bias._data = torch.zeros(CLASSES, device=device)
    model.collect_params().reset_ctx(device)
    model.hybridize()
    return model


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from tennis_b200 import _lib
    from tennis_b200.parallel import HostPipeline, ShardedCNNRNN
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: tennis_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    from tennis_b200.parallel import bind_to_gpu_numa_node
    full_affinity = os.sched_getaffinity(0)
    numa_node = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: host clips live next to their GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.check(_lib.lib().tn_device_check(local_rank))

    model = build_model(device)
    sharded = ShardedCNNRNN(model)
    pipe = HostPipeline(sharded, chunks=8)

    B = CLIPS_PER_GPU
    g = torch.Generator().manual_seed(100 + rank)
    clips_host = torch.empty((B, T, 3, SIZE, SIZE), dtype=torch.float32).pin_memory()
    for i in range(B):  # synthetic normalised pixels (config 1 recipe), generated per clip to bound host memory
        clips_host[i] = torch.randn(T, 3, SIZE, SIZE, generator=g)
    clips_dev = clips_host.to(device, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(max(3, args.warmup)):
        logits = sharded(clips_dev)
    barrier()
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        logits = sharded(clips_dev)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    frames_total = B * T * world * args.steps
    value = frames_total / (ms_total * 1e-3)
    finite = bool(torch.isfinite(logits).all().item())

    # ---- end to end from pinned host memory (e2e): the call a user of the scripts makes -- decoded uint8 NHWC frames in pinned
    # host memory -> HostPipeline.submit()/result() -> logits on the host.  Every step's H2D copy, the forward and the D2H read
    # of the logits are inside the timed region; one step is kept in flight (the copy of step s+1 runs under the kernels of s).
    def pipelined(pipe_, host):
        pipe_.result(pipe_.submit(host))
        pipe_.result(pipe_.submit(host))
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pending = pipe_.submit(host)
        out_ = None
        for s_i in range(args.steps):
            nxt = pipe_.submit(host) if s_i + 1 < args.steps else None
            out_ = pipe_.result(pending)
            pending = nxt
        p1.record()
        barrier()
        tt = torch.tensor([p0.elapsed_time(p1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return frames_total / (tt.item() * 1e-3), out_

    clips_u8 = torch.empty((B, T, SIZE, SIZE, 3), dtype=torch.uint8).pin_memory()
    clips_u8.random_(0, 256, generator=g)
    pipe_u8 = HostPipeline(sharded, chunks=3)  # with a step in flight the copy is already hidden: few, large chunks
    e2e_u8_value, out8 = pipelined(pipe_u8, clips_u8)
    d2h = out8.numel() * out8.element_size()
    # the same frames as normalised fp32 NCHW tensors (the reference's in-memory format after its transforms): 4x the bytes over
    # the host link, which then bounds the step (~55 GB/s per GPU, less when eight ranks share one host)
    pipe_f32 = HostPipeline(sharded, chunks=3)
    e2e_f32_value, _ = pipelined(pipe_f32, clips_host)
    h2d = clips_host.numel() * clips_host.element_size()
    # one blocking call per step (chunked H2D/compute overlap inside the call only)
    for _ in range(2):
        pipe(clips_u8)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        pipe(clips_u8)
    f1.record()
    barrier()
    t2 = torch.tensor([f0.elapsed_time(f1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_serial_value = frames_total / (t2.item() * 1e-3)
    plan_u8 = dict(pipe.last_plan)

    extra = {}
    # ---- fp32-grade mode (split-bf16, |logit - fp32 oracle| <= 1e-3: tests/test_gpu_precise.py), same workload, device-resident
    try:
        model.td.model.precision = "split_bf16"
        for _ in range(2):
            lp = sharded(clips_dev)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        psteps = max(2, min(args.steps, 5))
        q0.record()
        for _ in range(psteps):
            lp = sharded(clips_dev)
        q1.record()
        barrier()
        tp = torch.tensor([q0.elapsed_time(q1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        extra["precise"] = {
            "mode": "split_bf16: every tensor a (hi, lo) bf16 pair, three tcgen05 products per contraction, fp32 accumulation",
            "value": B * T * world * psteps / (tp.item() * 1e-3), "unit": "frames/s", "ms_per_step": tp.item() / psteps,
            "tolerance": "max |logit - fp32 oracle| 2.7e-4 on 64 x 32 clips (<= 1e-3), argmax bit-exact (tests/test_gpu_precise.py)",
            "max_abs_diff_vs_bf16_mode_logits": float((lp - logits).abs().max().item())}
    except Exception as exc:  # the headline metric must still be reported
        extra["precise"] = {"error": repr(exc)}
    finally:
        model.td.model.precision = "bf16"

    # ---- strong scaling: BASELINE.json configs[2]/[4] shape the work as 256 clips in total over the GPUs of the run
    try:
        per = TRAIN_GLOBAL_CLIPS // world
        strong_dev = clips_dev[:per] if per <= B else torch.cat([clips_dev] * (per // B), 0)
        for _ in range(2):
            sharded(strong_dev)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ssteps = max(2, min(args.steps, 5))
        s0.record()
        for _ in range(ssteps):
            sharded(strong_dev)
        s1.record()
        barrier()
        ts_ = torch.tensor([s0.elapsed_time(s1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        extra["strong_scaling"] = {"workload": "%d clips x %d frames in total, %d per GPU, forward" % (per * world, T, per),
                                   "value": per * world * T * ssteps / (ts_.item() * 1e-3), "unit": "frames/s",
                                   "ms_per_step": ts_.item() / ssteps, "scaling": "strong"}
        del strong_dev
    except Exception as exc:
        extra["strong_scaling"] = {"error": repr(exc)}

    # ---- configs[2]: CNN+GRU training step (fwd + bwd + SGD), 256 clips sharded over the GPUs, frozen backbone (the published
    # 0042 setting: tensor-core forward, fused BPTT head, gradient all-reduce over NCCL in Trainer.step)
    try:
        extra["train_step"] = bench_train_step(model, clips_dev, device, world, barrier, args)
    except Exception as exc:
        extra["train_step"] = {"error": repr(exc)}

    # ---- the same step with a TRAINABLE backbone (train.py without --freeze_backbone / --feats_model): training-mode BatchNorm,
    # forward / data-gradient / weight-gradient contractions on tcgen05 (csrc/tn_gemm_tc.cu), gradient all-reduce of every parameter
    try:
        extra["train_step_trainable"] = bench_train_step_trainable(model, clips_dev, device, world, barrier)
    except Exception as exc:
        extra["train_step_trainable"] = {"error": repr(exc)}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        conv_ms = prof["conv_ms"]
        conv_launches = prof["conv_launches"]
        flops = FLOP_PER_FRAME * B * T * args.steps  # algorithmic conv FLOPs executed by this rank's conv-GEMM launches
        achieved = flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        traffic_pf, traffic_note = conv_traffic_per_frame()
        traffic_step = traffic_pf * B * T if traffic_pf is not None else None
        os.sched_setaffinity(0, full_affinity)  # the CPU baseline gets every host core again
        cpu = time_cpu_port(budget_s=12.0, clips_per_step=1, warmup=1)
        line = {
            "metric": "frames/sec (224x224 CNN+GRU fwd)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arch": ARCH, "head": "BiGRU(128)+max+Dense(11)",
                       "global_clips": B * world, "frames_per_step": B * T * world,
                       "parallelism": "frame-sharded dp%d: NCCL all-gather of bf16 features, temporal head on each rank's own clips, "
                                      "logits all-gather" % world,
                       "precision": "value/e2e: bf16 operands + fp32 accumulation (stated tolerance: logits within 2e-2 x max|logit| "
                                    "of the fp32 reference arithmetic); the fp32-grade mode (<= 1e-3) is the `precise` object",
                       "l2": "inputs %.2f GB per GPU per step > 126 MB L2, no flush needed" % (h2d / 1e9),
                       "input_formats": "`value`: normalised fp32 NCHW clips resident in HBM (the reference's tensor format, 1.23 GB per "
                                        "step through the input-conversion kernel); `e2e`: uint8 NHWC frames from pinned host memory "
                                        "(0.31 GB per step, normalised on the device) -- the cheaper conversion is why `e2e` can "
                                        "come within 1-2 % of `value` or pass it; `e2e_f32_host` is the fp32 host format",
                       "outputs_finite": finite, "numa_node_of_rank0": numa_node},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic_step,
                         "traffic_note": "DRAM bytes per step of all conv-kernel launches (like `achieved`): " + traffic_note +
                                         "; algorithmic minimum with the bottleneck kept on chip 24.3 MB/frame = 49.8 GB/step",
                         "hbm_view": {"achieved_gbs": (traffic_step * args.steps / (conv_ms * 1e-3) / 1e9) if (traffic_step and conv_ms > 0) else None,
                                      "peak_gbs": peaks.get("hbm_gbs")},
                         "kernel": "conv kernel family (conv1x1_ts / conv3x3_halo / conv_gemm / stem_pool, all tcgen05), %d launches/step, %.2f ms/step summed over "
                                   "CUDA events on the launch stream" % (conv_launches // max(1, args.steps),
                                                                         conv_ms / max(1, args.steps)),
                         "peak_source": "bf16_tflops_sustained, " + peak_src,
                         "algorithmic": "5.666 GFLOP/frame x %d frames/step" % (B * T)},
            "cpu_baseline": {"value": cpu["frames_per_s"], "unit": "frames/s", "cores": cpu["cores"], "kind": "port",
                             "sample": "%d x 1 clip (32 frames) through the torch-fp32 CPU oracle of the same model" % cpu["steps"]},
            "e2e": {"value": e2e_u8_value, "unit": "frames/s", "h2d_bytes_per_step": clips_u8.numel(), "d2h_bytes_per_step": d2h,
                    "path": "pinned host uint8 NHWC frames (what the decoder / dataset delivers) -> HostPipeline.submit()/result(): "
                            "H2D copy, ToTensor+Normalize on the device, forward and logits D2H of every step are inside the timed "
                            "region; one step is kept in flight (the copy of step s+1 runs under the kernels of step s)"},
            "e2e_serial": {"value": e2e_serial_value, "unit": "frames/s", "h2d_bytes_per_step": clips_u8.numel(),
                           "d2h_bytes_per_step": d2h, "path": "same uint8 frames, one blocking call per step", "chunk_plan": plan_u8},
            "e2e_f32_host": {"value": e2e_f32_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "path": "normalised fp32 NCHW clips on the host (the reference's tensor format), submit()/result(); "
                                     "host-link-bandwidth-bound: 602 KB per frame"},
            "gpu_launches": int(prof["conv_launches"] + prof["other_launches"]),
            "clocks": clocks,
        }
        line.update(extra)
        if world == 1:
            try:
                line["resnet18_v2"] = bench_resnet18(clips_dev, device)
            except Exception as exc:
                line["resnet18_v2"] = {"error": repr(exc)}
            try:
                line["gru_head"] = bench_gru_head(model, device)
            except Exception as exc:
                line["gru_head"] = {"error": repr(exc)}
            try:
                line["training"] = bench_training(device)
            except Exception as exc:
                line["training"] = {"error": repr(exc)}
            try:
                line["captioner"] = bench_captioner(device)
            except Exception as exc:  # the headline metric must still be reported
                line["captioner"] = {"error": repr(exc)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _capture_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Route fd 1 to stderr for the whole run and keep a private handle for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def bench_captioner(device, steps=3):
    """BASELINE.json configs[3]: GNMT captioner, 1024-d features -> 2-layer LSTM enc/dec + attention, V=254, beam 5.
    Returns caption tokens/s (best beam, BOS/EOS stripped, train_gnmt.py:289-294) from host features to host token ids,
    next to the CPU oracle port timed on a bounded sample (4 sentences)."""
    import torch
    from tennis_b200 import synthetic as S
    from tennis_b200.gluon import Dropout, Embedding, HybridSequential
    from tennis_b200.models.captioning.gnmt import BeamSearchScorer, NMTModel, get_gnmt_encoder_decoder
    from tennis_b200.utils.translation import BeamSearchTranslator
    from tennis_b200.vocab import Vocab, count_tokens
    B, Tsrc, D, H, E, V, beam, max_len = 32, 224, 1024, 128, 100, 254, 5, 150
    p = S.synthetic_gnmt_params(seed=10000, scale=0.35, cell="lstm", H=H, D_src=D, E=E, V=V)
    vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
    src_embed = HybridSequential()
    src_embed.add(Dropout(0.0))
    enc, dec = get_gnmt_encoder_decoder(cell_type="lstm", hidden_size=H, dropout=0.0, num_layers=2, num_bi_layers=1)
    model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix="gnmt_",
                     src_embed=src_embed, tgt_embed=Embedding(V, E))
    params = model.collect_params()
    for k, v in p.items():
        params[k].shape, params[k]._data = tuple(v.shape), v.to(device)
        params[k]._version += 1
    x, vl = S.synthetic_sources(B, Tsrc, D, seed=100, min_len=64)
    xh, vlh = x.pin_memory(), vl.pin_memory()
    tr = BeamSearchTranslator(model, beam_size=beam, scorer=BeamSearchScorer(alpha=1.0, K=5), max_length=max_len)

    def run():
        s, _, v = tr.translate(xh.to(device, non_blocking=True), vlh.to(device, non_blocking=True))
        v = v.cpu()
        return s.cpu(), v, int((v[:, 0] - 2).clamp(min=0).sum())
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    toks = 0
    for _ in range(steps):
        s, v, n = run()
        toks += n
    dt = time.perf_counter() - t0
    # CPU port on a bounded sample (cpu_baseline leg: the only place this function touches the oracle), checked for token
    # equality on that sample
    from oracle import captioning as C
    nb = 4
    t1 = time.perf_counter()
    with torch.no_grad():
        s_ref, _, v_ref = C.translate(p, x[:nb], vl[:nb], cell="lstm", H=H, beam=beam, max_length=max_len)
    dt_cpu = time.perf_counter() - t1
    same = C.best_tokens(s_ref, v_ref) == C.best_tokens(s[:nb], v[:nb])
    return {"metric": "caption tokens/sec (GNMT LSTM, beam 5)", "value": toks / dt, "unit": "tokens/s",
            "config": {"workload": "configs[3]: B=32 sources x T_src<=224 x 1024-d, H=128, V=254, beam=5, max_length=150",
                       "decode_steps_per_call": int(s.shape[2]) - 1, "host_to_host": True},
            "cpu_baseline": {"value": float((v_ref[:, 0] - 2).clamp(min=0).sum()) / dt_cpu, "unit": "tokens/s",
                             "cores": torch.get_num_threads(), "kind": "port", "sample": "%d of the 32 sources" % nb},
            "token_ids_equal_to_oracle_on_sample": bool(same)}


def bench_training(device, steps=4):
    """Training rows of SURVEY.md 8a, timed as the scripts run them (one optimiser step = forward with saved activations,
    loss, backward, update), beside the CPU oracle + torch.autograd on a bounded sample:
      head      : CNNRNN(None, 11, 'gru') on pre-extracted features, 256 clips x 32 x 1024 (train.py:404-424, the 0042 setting)
      captioner : NMTModel LSTM H=128 on B=128 sources of <=224 x 1024-d features, targets <= 30 tokens, Adam (train_gnmt.py:330-337)
    """
    import torch
    from tennis_b200 import autograd
    from tennis_b200 import synthetic as S
    from tennis_b200.gluon import (Dropout, Embedding, HybridSequential, MaskedSoftmaxCELoss, SoftmaxCrossEntropyLoss,
                                   Trainer)
    from tennis_b200.models.captioning.gnmt import NMTModel, get_gnmt_encoder_decoder
    from tennis_b200.models.vision.definitions import CNNRNN
    from tennis_b200.vocab import Vocab, count_tokens
    out = {}
    # ---- temporal head on features
    B, Tn, D, H = 256, 32, 1024, HIDDEN
    g = torch.Generator().manual_seed(11)
    feats = torch.randn(B, Tn, D, generator=g).relu().to(device)
    labels = torch.randint(0, CLASSES, (B,), generator=g).to(device)
    head = CNNRNN(None, CLASSES, hidden_size=H, type="gru")
    head.initialize(ctx=device)
    for k, v in S.synthetic_rnn_params("gru", D, H, seed=4321).items():
        prm = head.rnn._reg_params[k]
        prm.shape, prm._data = tuple(v.shape), v.to(device)
        prm._version += 1
    head(feats)  # materialise the deferred classifier shape
    loss_fn = SoftmaxCrossEntropyLoss()
    tr = Trainer(head.collect_params(), 'sgd', {'learning_rate': 1e-3, 'momentum': 0.9, 'wd': 1e-4})

    def head_step():
        with autograd.record():
            loss = loss_fn(head(feats), labels)
        autograd.backward([loss])
        tr.step(B)
        return loss
    for _ in range(2):
        head_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = head_step()
    lh = float(loss.cpu().numpy().mean())
    dt = (time.perf_counter() - t0) / steps
    out["head"] = {"workload": "BiGRU(128)+max+Dense(11) training step on (256,32,1024) features, SGD momentum",
                   "ms_per_step": dt * 1e3, "clips_per_s": B / dt, "frames_per_s": B * Tn / dt, "loss_finite": lh == lh}
    # ---- captioner
    Bc, Ts, Dc, Hc, E, V, Tt = 128, 224, 1024, 128, 100, 254, 30
    p = S.synthetic_gnmt_params(seed=10000, scale=0.1, cell="lstm", H=Hc, D_src=Dc, E=E, V=V)
    vocab = Vocab(count_tokens(["w%03d" % i for i in range(V - 4)]))
    src_embed = HybridSequential()
    src_embed.add(Dropout(0.0))
    enc, dec = get_gnmt_encoder_decoder(cell_type="lstm", hidden_size=Hc, dropout=0.2, num_layers=2, num_bi_layers=1)
    model = NMTModel(src_vocab=None, tgt_vocab=vocab, encoder=enc, decoder=dec, embed_size=E, prefix="gnmt_",
                     src_embed=src_embed, tgt_embed=Embedding(V, E))
    params = model.collect_params()
    for k, v in p.items():
        params[k].shape, params[k]._data = tuple(v.shape), v.clone().to(device)
        params[k]._version += 1
    x, vl = S.synthetic_sources(Bc, Ts, Dc, seed=100, min_len=64)
    tgt = torch.randint(4, V, (Bc, Tt), generator=g).float()
    tvl = torch.randint(6, Tt + 1, (Bc,), generator=g).float()
    tvl[0] = Tt
    xs, vls, ts, tvls = x.to(device), vl.to(device), tgt.to(device), tvl.to(device)
    scale = float((Tt - 1) / (tvl - 1).mean())
    mce = MaskedSoftmaxCELoss()
    trc = Trainer(model.collect_params(), 'adam', {'learning_rate': 1e-3})

    def cap_step():
        with autograd.record():
            o, _ = model(xs, ts[:, :-1], vls, tvls - 1)
            lv = mce(o, ts[:, 1:], tvls - 1)
        autograd.backward([lv], [torch.full_like(lv, scale / Bc)])
        trc.step(1)
        return lv
    cap_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        lv = cap_step()
    lc = float(lv.cpu().numpy().mean()) * scale
    dt = (time.perf_counter() - t0) / steps
    words = float(vl.sum() + (tvl - 1).sum())  # the reference's "wps" numerator (train_gnmt.py:339-340)
    # CPU port (cpu_baseline leg): oracle forward + torch.autograd on 8 sentences
    from oracle import captioning as C
    nb = 8
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    t1 = time.perf_counter()
    o = C.nmt_forward(q, x[:nb], tgt[:nb, :-1], vl[:nb], tvl[:nb] - 1, cell="lstm", H=Hc)
    (C.masked_softmax_ce(o, tgt[:nb, 1:], tvl[:nb] - 1).mean() * scale).backward()
    dt_cpu = time.perf_counter() - t1
    words_cpu = float(vl[:nb].sum() + (tvl[:nb] - 1).sum())
    out["captioner"] = {"workload": "GNMT LSTM H=128 training step: B=128, T_src<=224 x 1024-d, T_tgt<=30, V=254, dropout 0.2, Adam",
                        "ms_per_step": dt * 1e3, "words_per_s": words / dt, "loss": lc,
                        "cpu_baseline": {"value": words_cpu / dt_cpu, "unit": "words/s", "cores": torch.get_num_threads(),
                                         "kind": "port", "sample": "forward + autograd backward of %d of the 128 sentences" % nb}}
    return out


def bench_train_step(model, clips_dev, device, world, barrier, args, steps=3):
    """BASELINE.json configs[2]: CNN+GRU event-detector TRAINING step (forward + backward + SGD) on 256 clips x 32 frames in total,
    sharded over the GPUs of the run (reference train.py:410-424: split_and_load, per-device forward/backward, trainer.step sums
    the gradients).  Frozen DenseNet-121 backbone -- the published CNN-RNN 0042 setting: tensor-core CNN forward, bi-GRU with saved
    activations, softmax-CE, fused BPTT, SGD with momentum; Trainer.step all-reduces the head gradients over NCCL."""
    import torch
    import torch.distributed as dist
    from tennis_b200 import autograd
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer
    per = TRAIN_GLOBAL_CLIPS // world
    B0 = clips_dev.shape[0]
    x = clips_dev[:per] if per <= B0 else torch.cat([clips_dev] * (per // B0), 0)
    for prm in model.td.model.collect_params().values():
        prm.grad_req = 'null'
    labels = torch.arange(per, device=device) % CLASSES
    loss_fn = SoftmaxCrossEntropyLoss()
    tr = Trainer(model.collect_params(), 'sgd', {'learning_rate': 1e-3, 'momentum': 0.9, 'wd': 1e-4})

    def step():
        with autograd.record():
            loss = loss_fn(model(x), labels)
        autograd.backward([loss])
        tr.step(TRAIN_GLOBAL_CLIPS)
        return loss
    for _ in range(2):
        loss = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    lv = float(loss.float().mean().item())
    return {"workload": "configs[2]: CNN+GRU training step (fwd+bwd+SGD), %d clips x %d frames in total, %d per GPU, frozen "
                        "DenseNet-121 backbone, NCCL gradient all-reduce" % (per * world, T, per),
            "ms_per_step": ms, "clips_per_s": per * world / (ms * 1e-3), "frames_per_s": per * world * T / (ms * 1e-3),
            "scaling": "strong", "loss_finite": lv == lv}


def bench_train_step_trainable(model, clips_dev, device, world, barrier, clips_per_gpu=8, steps=2):
    """CNN+GRU training step with the DenseNet-121 backbone TRAINED end to end (reference train.py:410-424 without
    --freeze_backbone): batch-statistics BatchNorm, convolution forward / dgrad / wgrad as split-bf16 tcgen05 GEMMs, bi-GRU BPTT,
    SGD with momentum on every parameter, NCCL all-reduce of all gradients.  Weak scaling: `clips_per_gpu` clips on every GPU."""
    import torch
    import torch.distributed as dist
    from tennis_b200 import _lib, autograd, tcgemm
    from tennis_b200.gluon import SoftmaxCrossEntropyLoss, Trainer
    x = clips_dev[:clips_per_gpu]
    per = x.shape[0]
    backbone_params = list(model.td.model.collect_params().values())
    saved = [prm.grad_req for prm in backbone_params]
    for prm in backbone_params:
        if not prm.name.endswith(("running_mean", "running_var")):
            prm.grad_req = 'write'
    labels = torch.arange(per, device=device) % CLASSES
    loss_fn = SoftmaxCrossEntropyLoss()
    tr = Trainer(model.collect_params(), 'sgd', {'learning_rate': 1e-4, 'momentum': 0.9, 'wd': 1e-4})

    def step():
        with autograd.record():
            loss = loss_fn(model(x), labels)
        autograd.backward([loss])
        tr.step(per * world)
        return loss
    try:
        loss = step()
        barrier()
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        barrier()
        prof = _lib.profile_read(reset=True)
        _lib.profile_enable(False)
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item() / steps
        lv = float(loss.float().mean().item())
    finally:
        for prm, g in zip(backbone_params, saved):
            prm.grad_req = g
    frames = per * world * T
    peaks, _ = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    algo_tflops = 3.0 * FLOP_PER_FRAME * per * T / (ms * 1e-3) / 1e12  # forward + dgrad + wgrad, per GPU
    return {"workload": "CNN+GRU training step (fwd+bwd+SGD) with a TRAINABLE DenseNet-121, %d clips x %d frames per GPU, "
                        "NCCL all-reduce of every gradient" % (per, T),
            "ms_per_step": ms, "frames_per_s": frames / (ms * 1e-3), "clips_per_s": per * world / (ms * 1e-3), "scaling": "weak",
            "gemm": "TN_TRAIN_GEMM=%s (x3: split-bf16, three tcgen05 products per contraction)" % tcgemm.mode(),
            "tensor_core_gemm_ms_per_step": prof["conv_ms"] / steps, "tensor_core_gemm_launches_per_step": prof["conv_launches"] // steps,
            "other_kernels_ms_per_step": prof["other_ms"] / steps,
            "roofline": {"bound": "tensor", "algorithmic": "3 x 5.666 GFLOP/frame (forward, data gradient, weight gradient)",
                         "achieved": algo_tflops, "peak": peak, "unit": "TFLOP/s", "frac": algo_tflops / peak,
                         "note": "whole step incl. BatchNorm, operand splits, BPTT and SGD; fp32 activations bound the batch at 8 "
                                 "clips x 32 frames per GPU (34 GB), 64 x 32 clips do not fit"},
            "round1_simt_fp32_path": "954 ms per 64 frames = 67 frames/s on the same GPU type (profiles/r2_cnn_train.md)",
            "loss_finite": lv == lv}


def bench_resnet18(clips_dev, device, steps=5):
    """ResNet-18 v2 `features` (the reference's --backbone default, train.py:32) on the same 2048 frames: frames/s and the fraction of
    the bf16 tensor roofline (3.627 GFLOP/frame)."""
    import torch
    from tennis_b200 import _lib, ops
    from tennis_b200 import synthetic as O
    p = O.synthetic_params("resnet18_v2", seed=1234)
    bb = ops.Backbone("resnet18_v2", O.flatten_params("resnet18_v2", p), device=device.index or 0)
    x = clips_dev.reshape((-1,) + tuple(clips_dev.shape[2:]))
    for _ in range(2):
        bb(x)
    torch.cuda.synchronize()
    _lib.profile_read(reset=True)
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = bb(x)
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    peaks, peak_src = measured_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    ach = RESNET_FLOP_PER_FRAME * x.shape[0] * steps / (prof["conv_ms"] * 1e-3) / 1e12 if prof["conv_ms"] > 0 else 0.0
    return {"workload": "ResNet-18 v2 features, %d frames @%dx%d, forward, device-resident" % (x.shape[0], SIZE, SIZE),
            "value": x.shape[0] / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "algorithmic": "3.627 GFLOP/frame", "peak_source": "bf16_tflops_sustained, " + peak_src},
            "outputs_finite": bool(torch.isfinite(out).all().item())}


def bench_gru_head(model, device, iters=20):
    """north_star secondary target: the Bi-GRU(128)+max+Dense head at 256 clips x 32 frames against the HBM roofline.
    Algorithmic bytes: features B*T*D*2 (bf16 as exchanged) + 3.56 MB weights + logits."""
    import torch
    B, Tn, D = 256, 32, 1024
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(B, Tn, D, generator=g).relu().to(device)
    twin = feats.to(torch.bfloat16)
    feats._tn_bf16 = twin

    def run():
        y = model.rnn.forward_max(feats)
        return model.classes(y)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    bytes_alg = B * Tn * D * 2 + 3.56e6 + B * 11 * 4
    peaks, _ = measured_peaks()
    gbs = bytes_alg / (us * 1e-6) / 1e9
    return {"workload": "BiGRU(128)+max+Dense(11) on (256,32,1024) bf16 features", "us_per_call": us,
            "algorithmic_bytes": bytes_alg, "achieved_gbs": gbs, "hbm_peak_gbs": peaks["hbm_gbs"],
            "frac_of_hbm_roofline": gbs / peaks["hbm_gbs"],
            "note": "32 serial recurrence steps: latency-bound (per-step breakdown: profiles/r2_summary.md section 5)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-launch ourselves with one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--impl", args.impl]
        raise SystemExit(subprocess.call(cmd))  # the ranks inherit the real stdout: rank 0 prints the JSON line there
    _capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
