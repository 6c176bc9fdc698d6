#!/usr/bin/env python
"""bench.py -- AVC training throughput (pairs/s) of the B200-native L3 path.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU restatement of the reference path

A "step" is one keras `train_on_batch` of `cnn_L3_melspec2` (BASELINE.json configs[1]): forward, backward, gradient
all-reduce (N>1) and the Adam update on a batch of synthetic AVC pairs (224x224x3 uint8 frame + 1 s 48 kHz int16
audio), bf16 activations / tcgen05 convolutions with fp32 accumulation, per-GPU batch 64 (weak scaling).
`value` is timed with the inputs resident in HBM; `e2e` goes through the host-buffer call
(`Engine.train_step_host` -> `l3_train_step_host`) with the H2D copies and the metric read-back inside the
timed region.  The reference's own Keras/TF path cannot run, so `--impl reference` and `cpu_baseline`
time the PyTorch-CPU restatement in oracle/ ("port") on the box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_TYPE = "cnn_L3_melspec2"
PER_GPU_BATCH = 64
METRIC = "AVC training pairs/sec (cnn_L3_melspec2, fwd+bwd+Adam)"
UNIT = "pairs/s"
# algorithmic FLOPs (2*MACs, conv + dense only), per pair
FWD_GFLOP = 40.925
TRAIN_GFLOP = 122.78


def conv_class_gflop_per_pair():
    """fwd / dgrad / wgrad algorithmic GFLOP per pair for cnn_L3_melspec2 (dgrad skips nothing: both towers have an
    input BN, so even the first layer's data gradient is needed)."""
    def tower(h, w, c0, same):
        chans = [(c0, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512), (512, 512)]
        tot = 0.0
        for i, (ci, co) in enumerate(chans):
            tot += 2.0 * h * w * 9 * ci * co
            if i in (1, 3, 5):
                h, w = ((h + 1) // 2, (w + 1) // 2) if same else (h // 2, w // 2)
        return tot / 1e9
    t = tower(224, 224, 3, True) + tower(256, 199, 1, False)
    return {"conv_fwd": t, "conv_dgrad": t, "conv_wgrad": t}


# DRAM bytes (read + write) per training step at B=64 and the time-weighted tensor-pipe activity, summed over the
# launches of each convolution class, from ONE `ncu --clock-control none` capture of a whole step
# (profiles/r1_ncu_full_conv_step.txt, reduced to profiles/r1_ncu_conv_classes.json by tools/ncu_conv_step.py)
def ncu_conv_classes():
    p = os.path.join(ROOT, "profiles", "r1_ncu_conv_classes.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        names = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80}
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def _sample_smi(self):
        import subprocess
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
            if v.strip().lower() == "active":
                self.reasons.add(k)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self._nvml else 0.5)

    def finish(self):
        self._stop_evt.set()
        self.join(5)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm
def run_cpu_port(batch, steps, warmup, threads=None):
    """Times oracle.train_step (PyTorch CPU fp32 restatement of the reference graph) -> pairs/s."""
    import torch
    from oracle import l3_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = O.to_torch(O.init_weights(MODEL_TYPE, seed=20180123), requires_grad=True)
    st = O.AdamState()
    video, audio, label = O.synthetic_batch(batch, seed=1)
    for _ in range(warmup):
        O.train_step(video, audio, label, w, st, MODEL_TYPE, 1e-5)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(video, audio, label, w, st, MODEL_TYPE, 1e-5)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads


def embedding_delta_vs_port():
    """The second half of BASELINE.json's metric ("embedding max|delta| vs the reference"): the fp32 parity mode's audio
    embedding and AVC logits on two synthetic pairs against the fp64 CPU restatement (the reference's Keras path cannot
    run).  Part of the cpu_baseline leg: the only place bench.py touches oracle/."""
    import numpy as np
    import torch
    from l3embedding_b200.engine import Engine
    from oracle import l3_oracle as O
    w_np = O.init_weights(MODEL_TYPE, seed=20180123, randomize_bn=True)
    video, audio, _ = O.synthetic_batch(2, seed=42)
    eng = Engine(MODEL_TYPE, 2, "f32", training=True, weights=w_np)   # the configuration smoke() validates
    try:
        cfg = O.OracleConfig(dtype=torch.float64)
        w = O.to_torch(w_np, dtype=torch.float64)
        af = torch.from_numpy(O.pcm2float(audio, "float64"))
        vf = torch.from_numpy(O.scale_video(video)).double()
        emb = eng.embed_audio(audio, "original").cpu().numpy()
        ref = O.audio_embedding(af, w, MODEL_TYPE, "original", cfg).numpy()
        _, logits = eng.predict(video, audio)
        ref_logits = O.avc_forward(vf, af, w, MODEL_TYPE, False, cfg).numpy()
        return {"embedding_max_abs_delta": float(np.abs(emb - ref).max()), "embedding_abs_max": float(np.abs(ref).max()),
                "logits_max_abs_delta": float(np.abs(logits - ref_logits).max()),
                "mode": "f32 parity mode vs the fp64 CPU restatement, 2 synthetic pairs, 6144-d audio embedding"}
    finally:
        eng.close()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 4
    value, sec, threads = run_cpu_port(batch, args.steps, args.warmup)
    sample = "%d pairs per step x %d steps (bounded sample of the batch-%d workload)" % (batch, args.steps, PER_GPU_BATCH)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cnn_L3_melspec2 train_on_batch (fwd+bwd+Adam), synthetic AVC pairs",
                       "per_gpu_batch": PER_GPU_BATCH, "note": "PyTorch-CPU restatement of the Keras/TF graph (the "
                       "reference itself needs keras 2.0.9 / TF 1.4 / kapre, not installable here)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from l3embedding_b200 import _lib, dp
    from l3embedding_b200.engine import Engine
    from l3embedding_b200.synthetic import synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version banner / debug output to STDOUT (fd 1) while the communicator comes up: point fd 1
        # at stderr for that phase so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    par = dp.TorchDistReplicas() if world > 1 else dp.SingleReplica()
    B = args.batch
    G = B * world
    lr = 1e-5   # jobs/l3embedding-train-melspec2-09192018.sbatch
    eng = Engine(MODEL_TYPE, B, args.dtype, training=True, device=dev, seed=20180123)
    lib = _lib.load()

    # a pool of distinct synthetic batches, resident in HBM (and pinned on the host for the e2e arm)
    pool_n = args.pool
    pool_dev, pool_host = [], []
    for i in range(pool_n):
        v, a, l = synthetic_batch(B, seed=20180123 + 1000 * rank + i)
        hv, ha, hl = (torch.from_numpy(x).pin_memory() for x in (v, a, l))
        pool_host.append((hv, ha, hl))
        pool_dev.append((hv.to(dev), ha.to(dev), hl.to(dev)))
    torch.cuda.synchronize()

    def step_resident(i):
        v, a, l = pool_dev[i % pool_n]
        eng.forward_backward(v, a, l, global_batch=G)
        par.allreduce_grads(eng)
        eng.adam_step(lr)

    def step_e2e(i):
        hv, ha, hl = pool_host[i % pool_n]
        if world == 1:
            return eng.train_step_host(hv.numpy(), ha.numpy(), hl.numpy(), lr)
        # N>1: H2D upload, forward/backward, NCCL all-reduce, Adam, metric read-back
        v, a, l = hv.to(dev, non_blocking=True), ha.to(dev, non_blocking=True), hl.to(dev, non_blocking=True)
        eng.forward_backward(v, a, l, global_batch=G)
        par.allreduce_grads(eng)
        m = eng.metrics()
        eng.adam_step(lr)
        return m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        if profile:
            eng.profile(True)
        l0 = lib.l3_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        clocks = sampler.finish()
        ms = e0.elapsed_time(e1)
        prof = eng.profile_read() if profile else None
        if profile:
            eng.profile(False)
        launches = lib.l3_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, prof, launches

    ms, clocks, _, launches = timed(step_resident, args.steps, args.warmup)
    value = G * args.steps / (ms / 1e3)
    # per-kernel-class durations: a separate pass with the two towers serialised on one stream (in the timed runs
    # they overlap on two streams, which makes per-class CUDA-event intervals overlap too)
    eng.set_two_streams(False)
    prof_steps = max(3, min(args.steps, 10))
    ms_serial, _, prof, _ = timed(step_resident, prof_steps, 2, profile=True)
    eng.set_two_streams(True)
    e2e_steps = max(3, min(args.steps, 10))
    ms2, _, _, _ = timed(step_e2e, e2e_steps, 2)
    e2e_value = G * e2e_steps / (ms2 / 1e3)

    if rank == 0:
        peak_tf, peak_hbm, peak_src = measured_peaks()
        gf = conv_class_gflop_per_pair()
        kernels = {}
        for k, (kms, n) in prof.items():
            if k in gf and kms > 0:
                kernels[k] = {"ms_per_step": kms / prof_steps, "launches_per_step": n / prof_steps,
                              "tflops": gf[k] * B * prof_steps / kms, "frac_of_serial_step": kms / ms_serial}
            elif kms > 0:
                kernels[k] = {"ms_per_step": kms / prof_steps, "launches_per_step": n / prof_steps,
                              "frac_of_serial_step": kms / ms_serial}
        dom = max((k for k in kernels if k in gf), key=lambda k: kernels[k]["ms_per_step"])
        ncu = ncu_conv_classes()
        achieved = kernels[dom]["tflops"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "cnn_L3_melspec2 train_on_batch (fwd+bwd+Adam), synthetic AVC pairs "
                                   "(224x224x3 u8 frame + 48000-sample i16 audio)",
                       "per_gpu_batch": B, "global_batch": G, "parallelism": "dp%d" % world,
                       "tensor_cores": bool(eng.uses_tensor_cores), "tower_streams": 2,
                       "l2": "inputs rotate over a %d-batch pool; each step streams >5 GB of activations through the "
                             "126 MB L2, so no step sees a warm L2" % pool_n},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * (224 * 224 * 3 + 48000 * 2 + 8),
                    "d2h_bytes_per_step": 16, "steps": e2e_steps, "ms_per_step": ms2 / e2e_steps,
                    "api": ("Engine.train_step_host -> l3_train_step_host (pinned host buffers)" if world == 1 else
                            "pinned host batch -> H2D -> Engine.forward_backward -> NCCL all-reduce -> metrics D2H -> adam")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "traffic": (ncu.get(dom, {}).get("dram_bytes") if (B == 64 and args.dtype == "bf16") else None),
                         "traffic_unit": "DRAM bytes per step for the class (ncu, profiles/r1_ncu_full_conv_step.txt)",
                         "tensor_pipe_active_pct_ncu": {k: round(v["tensor_pipe_active_pct_time_weighted"], 1)
                                                        for k, v in ncu.items()},
                         "peak_source": peak_src,
                         "whole_step_frac": value * TRAIN_GFLOP / 1e3 / world / peak_tf,
                         "serial_ms_per_step": ms_serial / prof_steps,
                         "note": "kernel classes timed with CUDA events on the library stream in a pass with the two "
                                 "towers serialised; the headline step overlaps them on two streams",
                         "kernels": kernels},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, csteps = 8, 1
            v, sec, threads = run_cpu_port(cb, csteps, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d-pair batch, 1 warm-up + %d timed train_step of the PyTorch-CPU "
                                              "restatement (oracle/), %.1f s/step" % (cb, csteps, sec)}
            try:
                line["cpu_baseline"]["parity"] = embedding_delta_vs_port()
            except Exception as e:   # a reporting extra must never cost the bench line
                line["cpu_baseline"]["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--pool", type=int, default=4, help="distinct input batches cycled through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
