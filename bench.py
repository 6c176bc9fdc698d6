#!/usr/bin/env python
"""bench.py -- AVC training throughput (pairs/s) of the B200-native L3 path.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU restatement of the reference path

A "step" is one keras `train_on_batch` of `cnn_L3_melspec2` (BASELINE.json configs[1]): forward, backward, gradient
all-reduce (N>1) and the Adam update on a batch of synthetic AVC pairs (224x224x3 uint8 frame + 1 s 48 kHz int16
audio), bf16 activations / tcgen05 convolutions with fp32 accumulation, per-GPU batch 64 (weak scaling).
`value` is timed with the inputs resident in HBM; `e2e` goes through the host-buffer calls a user's fit loop makes
(`Engine.upload_host` -> `l3_upload_batch_host` from pinned buffers, then `l3_train_step_staged` /
`l3_dp_train_step_staged`) with the H2D copies and the metric read-back inside the timed region; the upload of
batch k+1 is enqueued on the library's copy stream before step k is run, as `L3Model.fit_generator` does.
At N > 1 the gradient exchange is the library's own (l3_dp_*: NCCL all-reduce in buckets overlapped with backward).
The reference's own Keras/TF path cannot run, so `--impl reference` and `cpu_baseline` time the
PyTorch-CPU restatement in oracle/ ("port") on the box's host cores, at the same per-GPU batch.
Extra records under `configs`: BASELINE configs[2] (audio-embedding inference, 10k clips) at N=1, configs[3]
(cnn_L3_kapredbinputbn, 4x64) when --gpus 4, configs[4] (cnn_L3_melspec2, 8x128) when --gpus 8.
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
# launches of each convolution class, from ONE `ncu --clock-control none` capture of a whole step, reduced by
# tools/ncu_conv_step.py.  STATIC: read from the committed profile, not measured in the run that prints the line.
NCU_CLASSES_FILES = ["profiles/r2_ncu_conv_classes.json", "profiles/r1_ncu_conv_classes.json"]


def ncu_conv_classes():
    for rel in NCU_CLASSES_FILES:
        try:
            return json.load(open(os.path.join(ROOT, rel))), rel
        except Exception:
            continue
    return {}, None


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
def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return None


def run_cpu_port(batch, steps, warmup, threads=None, budget_s=None):
    """Times oracle.train_step (PyTorch CPU fp32 restatement of the reference graph) -> pairs/s.  With budget_s the
    number of steps (never the batch) is cut so that the whole run stays inside the budget; returns what was run."""
    import torch
    from oracle import l3_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    w = O.to_torch(O.init_weights(MODEL_TYPE, seed=20180123), requires_grad=True)
    st = O.AdamState()
    video, audio, label = O.synthetic_batch(batch, seed=1)
    t0 = time.perf_counter()
    done_warm = 0
    for _ in range(max(warmup, 1)):
        O.train_step(video, audio, label, w, st, MODEL_TYPE, 1e-5)
        done_warm += 1
        if budget_s is not None and (time.perf_counter() - t0) * (done_warm + 1) / done_warm > 0.3 * budget_s:
            break
    per = (time.perf_counter() - t0) / done_warm
    if budget_s is not None:
        steps = max(1, min(steps, int((budget_s - (time.perf_counter() - t0)) / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(video, audio, label, w, st, MODEL_TYPE, 1e-5)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, threads, steps, done_warm


def parity_vs_port():
    """The second half of BASELINE.json's metric ("embedding max|delta| vs the reference"): audio embedding (6144-d) and
    AVC logits on two synthetic pairs against the fp64 CPU restatement (the reference's Keras path cannot run) -- for the
    fp32 parity mode (north_star's 1e-3 bar) AND for the bf16 throughput mode that the line's throughput is measured in.
    Part of the cpu_baseline leg: the only place bench.py touches oracle/."""
    import numpy as np
    import torch
    from l3embedding_b200.engine import Engine
    from oracle import l3_oracle as O
    w_np = O.init_weights(MODEL_TYPE, seed=20180123, randomize_bn=True)
    video, audio, _ = O.synthetic_batch(2, seed=42)
    cfg = O.OracleConfig(dtype=torch.float64)
    w = O.to_torch(w_np, dtype=torch.float64)
    af = torch.from_numpy(O.pcm2float(audio, "float64"))
    vf = torch.from_numpy(O.scale_video(video)).double()
    ref = O.audio_embedding(af, w, MODEL_TYPE, "original", cfg).numpy()
    ref_logits = O.avc_forward(vf, af, w, MODEL_TYPE, False, cfg).numpy()
    out = {"reference": "fp64 CPU restatement (oracle/), 2 synthetic pairs, 6144-d audio embedding",
           "embedding_abs_max": float(np.abs(ref).max())}
    for mode in ("f32", "f32tc", "bf16"):
        eng = Engine(MODEL_TYPE, 2, mode, training=True, weights=w_np)
        try:
            emb = eng.embed_audio(audio, "original").cpu().numpy()
            _, logits = eng.predict(video, audio)
            out[mode] = {"embedding_max_abs_delta": float(np.abs(emb - ref).max()),
                         "logits_max_abs_delta": float(np.abs(logits - ref_logits).max())}
        finally:
            eng.close()
    out["note"] = ("f32 / f32tc = parity modes (bar 1e-3; SIMT fp32 / split 16-bit operands on tcgen05); bf16 = the throughput "
                   "mode this line's value / e2e are measured in")
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch
    avail = _mem_available_gb()
    note = ""
    if avail is not None and avail < 0.45 * batch + 8:          # ~0.33 GB of autograd state per pair (measured)
        batch = max(4, int((avail - 8) / 0.45))
        note = " (host memory %.0f GB: batch cut from %d)" % (avail, args.batch)
    value, sec, threads, steps, warm = run_cpu_port(batch, args.steps, args.warmup, budget_s=240.0)
    sample = "%d-pair batch (the per-GPU batch of the GPU arm)%s, %d warm-up + %d timed train_step, %.1f s/step" % (
        batch, note, warm, steps, sec)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cnn_L3_melspec2 train_on_batch (fwd+bwd+Adam), synthetic AVC pairs "
                                   "(224x224x3 u8 frame + 48000-sample i16 audio)",
                       "per_gpu_batch": batch, "global_batch": batch, "parallelism": "cpu",
                       "requested_steps": args.steps, "requested_warmup": args.warmup,
                       "note": "PyTorch-CPU restatement of the Keras/TF graph (the reference itself needs keras 2.0.9 / "
                               "TF 1.4 / kapre, not installable here); one process on all host cores; steps cut (never "
                               "the batch) to keep the run inside 4 minutes"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class TrainBench:
    """One training configuration (model type, per-GPU batch) on this rank: resident and end-to-end timings."""

    def __init__(self, model_type, B, dtype, world, rank, dev, par, pool_n, lr=1e-5):
        import torch
        from l3embedding_b200 import _lib
        from l3embedding_b200.engine import Engine
        from l3embedding_b200.synthetic import synthetic_batch
        self.torch, self.lib = torch, _lib.load()
        self.B, self.G, self.world, self.rank, self.dev, self.lr, self.pool_n = B, B * world, world, rank, dev, lr, pool_n
        self.eng = Engine(model_type, B, dtype, training=True, device=dev, seed=20180123)
        par.attach(self.eng)          # N > 1: joins the library's NCCL communicator (collective)
        self.exchange = par.world_size > 1
        # a pool of distinct synthetic batches, resident in HBM and pinned on the host for the e2e arm
        self.pool_dev, self.pool_host = [], []
        for i in range(pool_n):
            v, a, l = synthetic_batch(B, seed=20180123 + 1000 * rank + i)
            hv, ha, hl = (torch.from_numpy(x).pin_memory() for x in (v, a, l))
            self.pool_host.append((hv.numpy(), ha.numpy(), hl.numpy(), (hv, ha, hl)))
            self.pool_dev.append((hv.to(dev), ha.to(dev), hl.to(dev)))
        torch.cuda.synchronize()

    def step_resident(self, i):
        v, a, l = self.pool_dev[i % self.pool_n]
        self.eng.forward_backward(v, a, l, global_batch=self.G)     # N > 1: gradient buckets leave during backward
        self.eng.adam_step(self.lr)                                 # waits for them

    def _upload(self, i):
        hv, ha, hl, _ = self.pool_host[i % self.pool_n]
        self.eng.upload_host(hv, ha, hl)

    def step_e2e(self, i):
        # the fit loop's software pipeline: batch i was uploaded during step i-1; enqueue batch i+1, run step i
        self._upload(i + 1)
        if not self.exchange:
            return self.eng.train_step_staged(self.B, self.lr)
        return self.eng.dp_train_step_staged(self.B, self.G, self.lr)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, profile=False, local=0, prime=None):
        torch = self.torch
        if prime is not None:
            prime(0)
        for i in range(warmup):
            fn(i)
        self.barrier()
        if profile:
            self.eng.profile(True)
        l0 = self.lib.l3_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        self.barrier()
        clocks = sampler.finish()
        ms = e0.elapsed_time(e1)
        prof = self.eng.profile_read() if profile else None
        if profile:
            self.eng.profile(False)
        launches = self.lib.l3_launch_count() - l0
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, prof, launches

    def drain_staged(self):
        """the e2e loop leaves one uploaded batch pending: consume it so the next phase starts clean"""
        if not self.exchange:
            self.eng.train_step_staged(self.B, self.lr)
        else:
            self.eng.dp_train_step_staged(self.B, self.G, self.lr)

    def run(self, steps, warmup, local, with_profile):
        ms, clocks, _, launches = self.timed(self.step_resident, steps, warmup, local=local)
        out = {"value": self.G * steps / (ms / 1e3), "ms_per_step": ms / steps, "clocks": clocks, "launches": int(launches)}
        if with_profile:
            # per-kernel-class durations: a separate pass with the two towers serialised on one stream (in the timed
            # runs they overlap on two streams, which makes per-class CUDA-event intervals overlap too)
            self.eng.set_two_streams(False)
            ps = max(3, min(steps, 10))
            ms_serial, _, prof, _ = self.timed(self.step_resident, ps, 2, profile=True, local=local)
            self.eng.set_two_streams(True)
            out.update(prof=prof, prof_steps=ps, ms_serial=ms_serial)
        es = max(3, min(steps, 10))
        ms2, _, _, _ = self.timed(self.step_e2e, es, 2, local=local, prime=self._upload)
        self.drain_staged()
        out.update(e2e_value=self.G * es / (ms2 / 1e3), e2e_ms=ms2 / es, e2e_steps=es)
        return out

    def close(self):
        self.eng.close()


def embedding_inference_record(dev):
    """BASELINE configs[2]: audio-tower embedding inference (the 05_generate_embedding_samples.py path), 10 000 x 1 s
    clips on one B200: clips/s with the clips resident in HBM and host -> host (pinned int16 in, pinned float32 out)."""
    import torch
    from l3embedding_b200.engine import Engine
    from l3embedding_b200.synthetic import synthetic_batch
    N, BATCH = 10000, 500
    _, audio, _ = synthetic_batch(BATCH, seed=7)
    host = torch.from_numpy(audio).pin_memory()
    rec = {"workload": "cnn_L3_melspec2 audio embedding, %d clips of 1 s / 48 kHz int16, batch %d" % (N, BATCH)}
    for dtype, n_iter in (("bf16", N // BATCH), ("f32tc", 6), ("f32", 2)):
        eng = Engine(MODEL_TYPE, BATCH, dtype, training=False, towers=("audio",), host_staging=False, device=dev)
        try:
            d = host.to(dev)
            for pooling, dim in (("original", 6144), ("short", 512)):
                res = torch.empty(BATCH, dim, device=dev)
                for _ in range(2):
                    eng.embed_audio(d, pooling, out=res)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n_iter):
                    eng.embed_audio(d, pooling, out=res)
                e1.record()
                torch.cuda.synchronize()
                r = {"clips_per_s_resident": n_iter * BATCH / (e0.elapsed_time(e1) / 1e3), "clips": n_iter * BATCH,
                     "conv_tflops": n_iter * BATCH / (e0.elapsed_time(e1) / 1e3) * 20.405 / 1e3}
                rec["%s_%s" % (dtype, pooling)] = r
        finally:
            eng.close()
    # the call a user of the reference makes (05_generate_embedding_samples.py:154-157 -> features.py:304): the keras-
    # style embedding model's predict() on a pageable host array of all 10 000 clips, pageable float32 result
    try:
        import numpy as np
        from l3embedding_b200 import model as M
        m, _, _ = M.MODELS[MODEL_TYPE]()
        m.configure(dtype="bf16")
        x = np.tile(audio, (N // BATCH, 1, 1))
        for pooling in ("original", "short"):
            e, _, _ = M.convert_audio_model_to_embedding(m.get_layer("audio_model"), m.inputs[1], MODEL_TYPE, pooling)
            e.predict(x[:1024])                               # engine creation + warm-up
            t0 = time.perf_counter()
            y = e.predict(x)
            dt = time.perf_counter() - t0
            rec["bf16_" + pooling]["clips_per_s_host_to_host"] = N / dt
            rec["bf16_" + pooling]["host_to_host_api"] = ("EmbeddingModel.predict: pageable int16 (10000,1,48000) -> pageable "
                                                          "float32 (10000,%d), 3-stage pinned pipeline" % y.shape[1])
            e._engine.close()
    except Exception as ex:
        rec["host_to_host_error"] = "%s: %s" % (type(ex).__name__, ex)
    return rec


def main_gpu(args):
    import torch
    import torch.distributed as dist
    from l3embedding_b200 import dp

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
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)     # bootstrap + the timing barrier / max only
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        # --no-exchange (diagnostic): N independent replicas, same barrier / max-over-ranks timing, no gradient exchange
        par = dp.LibraryReplicas() if (world > 1 and not args.no_exchange) else dp.SingleReplica()
        B = args.batch
        tb = TrainBench(MODEL_TYPE, B, args.dtype, world, rank, dev, par, args.pool)
    finally:
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    G = B * world
    r = tb.run(args.steps, args.warmup, local, with_profile=True)
    uses_tc = bool(tb.eng.uses_tensor_cores)
    tb.close()

    # the other BASELINE configurations this launch can cover
    configs = {}
    if world == 4 or args.all_configs:
        t4 = TrainBench("cnn_L3_kapredbinputbn", 64, args.dtype, world, rank, dev, par, args.pool)
        r4 = t4.run(max(5, args.steps // 2), 3, local, with_profile=False)
        t4.close()
        configs["config4_kapredbinputbn_dp%d_x64" % world] = {
            "pairs_per_s": r4["value"], "ms_per_step": r4["ms_per_step"], "e2e_pairs_per_s": r4["e2e_value"],
            "global_batch": 64 * world, "whole_step_frac": r4["value"] * 122.54 / 1e3 / world / measured_peaks()[0]}
    if world == 8 or args.all_configs:
        t5 = TrainBench(MODEL_TYPE, 128, args.dtype, world, rank, dev, par, max(2, args.pool // 2))
        r5 = t5.run(max(5, args.steps // 2), 3, local, with_profile=False)
        t5.close()
        configs["config5_melspec2_dp%d_x128" % world] = {
            "pairs_per_s": r5["value"], "ms_per_step": r5["ms_per_step"], "e2e_pairs_per_s": r5["e2e_value"],
            "global_batch": 128 * world, "whole_step_frac": r5["value"] * TRAIN_GFLOP / 1e3 / world / measured_peaks()[0]}
    parity_mode = None
    if world == 1 and rank == 0 and not args.no_configs:
        # the modes that meet north_star's 1e-3 embedding bar, on the same training workload (few steps: they are slower)
        parity_mode = {"note": "same workload and batch as the headline; f32tc = fp32 storage, split 16-bit operands on "
                               "tcgen05 (fp16 parts forward, bf16 parts backward, fp32 accumulate); f32 = SIMT fp32"}
        for mode, st in (("f32tc", 6), ("f32", 3)):
            try:
                tp = TrainBench(MODEL_TYPE, B, mode, world, rank, dev, par, 2)
                ms_p, _, _, _ = tp.timed(tp.step_resident, st, 3, local=local)
                parity_mode[mode] = {"pairs_per_s": B * st / (ms_p / 1e3), "ms_per_step": ms_p / st, "steps": st}
                tp.close()
            except Exception as e:
                parity_mode[mode] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            configs["config3_embedding_inference"] = embedding_inference_record(dev)
        except Exception as e:   # a reporting extra must never cost the bench line
            configs["config3_embedding_inference"] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        peak_tf, peak_hbm, peak_src = measured_peaks()
        gf = conv_class_gflop_per_pair()
        prof, prof_steps, ms_serial = r["prof"], r["prof_steps"], r["ms_serial"]
        kernels = {}
        for k, (kms, n) in prof.items():
            if kms <= 0:
                continue
            kernels[k] = {"ms_per_step": kms / prof_steps, "launches_per_step": n / prof_steps,
                          "frac_of_serial_step": kms / ms_serial}
            if k in gf:
                kernels[k]["tflops"] = gf[k] * B * prof_steps / kms
        dom = max((k for k in kernels if k in gf), key=lambda k: kernels[k]["ms_per_step"])
        ncu, ncu_src = ncu_conv_classes()
        achieved = kernels[dom]["tflops"]
        static_ok = B == 64 and args.dtype == "bf16" and ncu_src is not None
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "cnn_L3_melspec2 train_on_batch (fwd+bwd+Adam), synthetic AVC pairs "
                                   "(224x224x3 u8 frame + 48000-sample i16 audio)",
                       "per_gpu_batch": B, "global_batch": G, "parallelism": "dp%d" % world,
                       "tensor_cores": uses_tc, "tower_streams": 2,
                       "gradient_exchange": ("none" if (world == 1 or args.no_exchange) else
                                             "libl3b200 l3_dp_*: NCCL all-reduce in buckets on a communication stream, "
                                             "overlapped with backward"),
                       "l2": "inputs rotate over a %d-batch pool; each step streams >5 GB of activations through the "
                             "126 MB L2, so no step sees a warm L2" % args.pool},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": B * (224 * 224 * 3 + 48000 * 2 + 8),
                    "d2h_bytes_per_step": 16, "steps": r["e2e_steps"], "ms_per_step": r["e2e_ms"],
                    "api": ("Engine.upload_host (pinned u8/i16 -> l3_upload_batch_host, copy stream) of batch k+1, then "
                            + ("l3_train_step_staged" if world == 1 else "l3_dp_train_step_staged")
                            + " of batch k (metrics read back every step)")},
            "gpu_launches": r["launches"],
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "traffic": (ncu.get(dom, {}).get("dram_bytes") if static_ok else None),
                         "traffic_unit": "DRAM bytes per step for the class",
                         "traffic_source": ("static: %s (one ncu capture of a B=64 step, not this run)" % ncu_src) if static_ok else None,
                         "tensor_pipe_active_pct_ncu": ({"source": "static: %s (not measured in this run)" % ncu_src,
                                                         **{k: round(v["tensor_pipe_active_pct_time_weighted"], 1)
                                                            for k, v in ncu.items()}} if ncu_src else None),
                         "peak_source": peak_src,
                         "whole_step_frac": r["value"] * TRAIN_GFLOP / 1e3 / world / peak_tf,
                         "serial_ms_per_step": ms_serial / prof_steps,
                         "note": "kernel classes timed live with CUDA events on the library stream in a pass with the two "
                                 "towers serialised; the headline step overlaps them on two streams",
                         "kernels": kernels},
        }
        if configs:
            line["configs"] = configs
        if parity_mode:
            line["parity_mode"] = parity_mode
        if world == 1 and not args.no_cpu_baseline:
            cb = 8
            v, sec, threads, csteps, cwarm = run_cpu_port(cb, 1, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d-pair batch, %d warm-up + %d timed train_step of the PyTorch-CPU "
                                              "restatement (oracle/), %.1f s/step; `--impl reference` runs the full "
                                              "%d-pair batch" % (cb, cwarm, csteps, sec, B)}
            try:
                line["cpu_baseline"]["parity"] = parity_vs_port()
            except Exception as e:   # a reporting extra must never cost the bench line
                line["cpu_baseline"]["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE-config records")
    ap.add_argument("--all-configs", action="store_true", help="run the config 4 / 5 records at any --gpus")
    ap.add_argument("--no-exchange", action="store_true", help="diagnostic: N independent replicas (no gradient exchange)")
    a = ap.parse_args()
    if a.warmup < 3 and a.impl == "b200":
        a.warmup = 3
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
