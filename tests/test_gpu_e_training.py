"""GPU parity tests of the training step in fp32 parity mode: loss, accuracy, every gradient tensor, BN moving
statistics and the keras-Adam update against fp64 autograd of the oracle; data-parallel replica semantics.

  gradients (f32)   <= 1e-2 relative L2 per tensor vs fp64 autograd (measured fp32 noise floor of the SAME graph in
                    PyTorch-CPU fp32 against its fp64 self: 5.8e-3 at B=2), or <= 1e-4 absolute for the analytically
                    zero ones (bias of a conv feeding training-mode BN)
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import l3_oracle as O
from _gpu_common import MODEL_TYPES, GOLDEN, F64, engine as _engine, rel_l2, pad as _pad, oracle_inputs as _oracle_inputs

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("mode", ["f32", "f32tc"])
@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2", "cnn_L3_orig"])
def test_training_step_gradients_f32(model_type, mode):
    """train_on_batch: loss, accuracy, every gradient tensor, BN moving statistics and the Adam update against
    fp64 autograd of the oracle."""
    B = 2
    w_np = O.init_weights(model_type, seed=7, randomize_bn=True)
    video, audio, label = O.synthetic_batch(B, seed=303)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
    grads, out, stats = O.compute_grads(vf, af, torch.from_numpy(label), w, model_type, F64)
    eng = _engine(model_type, B, mode, training=True, weights=w_np)   # f32tc: split 16-bit operands on tcgen05
    eng.forward_backward(video, audio, label)
    m = eng.metrics()
    assert abs(m["loss"] - float(out["loss"])) <= 1e-4 * max(1.0, abs(float(out["loss"])))
    assert abs(m["acc"] - float(out["acc"])) < 1e-6
    got = eng.get_grads()
    bad = []
    for name, g_ref in grads.items():
        g_ref = g_ref.numpy()
        if name.endswith("/kernel"):
            g_ref = g_ref - 2e-5 * w_np[name]          # the device applies the l2 term inside Adam
        err = rel_l2(got[name], g_ref)
        max_abs = float(np.abs(got[name] - g_ref).max())
        # fp32 noise floor, measured: the SAME graph in PyTorch-CPU fp32 deviates from its fp64 self by up to 5.8e-3
        # relative L2 on these gradients at B=2 (BN backward cancels large terms), and by up to 6e-5 absolute on the
        # analytically-zero ones (bias of a conv feeding training-mode BN).  Hence 1e-2 relative / 1e-4 absolute.
        tol = 1e-2
        if mode == "f32tc" and ("/bn0/" in name or name == "vision/bn1b/beta"):
            # f32tc splits the BACKWARD operands into two bf16 parts (fp32's range; 2^-17 per operand instead of 2^-24).
            # That is ~1e-5 on every well-conditioned gradient, but the input-BN gradients (and beta of the one BN that
            # feeds a convolution without a ReLU in between) are sums of data gradients that cancel down to border
            # effects -- the same tensors the bf16 test singles out (test_gpu_d_bf16.py) -- measured 1.1e-2 here
            tol = 3e-2
        if not (err <= tol or max_abs <= 1e-4):
            bad.append((name, err, max_abs, float(np.abs(g_ref).max())))
    assert not bad, bad
    # BN moving statistics (momentum 0.99, Bessel-corrected variance)
    O.update_moving_stats(w, stats, F64)
    w_after = eng.get_weights()
    for name in w_np:
        if name.endswith(("moving_mean", "moving_variance")):
            assert np.abs(w_after[name] - w[name].detach().numpy()).max() <= 1e-5, name
    # keras Adam step, isolated from gradient noise: apply the oracle's update rule to the DEVICE gradients (+ the l2
    # term the device folds into Adam) and compare the resulting weights.  (Feeding each side its own gradients is
    # meaningless for the analytically-zero ones: Adam normalises pure round-off noise to a +-lr step.)
    g_dev = {k: torch.from_numpy(v.astype(np.float64) + (2e-5 * w_np[k] if k.endswith("/kernel") else 0.0))
             for k, v in got.items()}
    w_ref = O.to_torch(w_np, dtype=torch.float64)
    O.adam_update(w_ref, g_dev, O.AdamState(), 1e-3, F64)
    eng.adam_step(1e-3)
    w_after = eng.get_weights()
    for name in g_dev:
        assert np.abs(w_after[name] - w_ref[name].numpy()).max() <= 2e-6, name


def test_parity_mode_gradients_are_run_to_run_stable():
    """The parity-mode reductions (weight gradients, BN statistics and BN-backward sums) are merged in fp64 and rounded
    once, so two runs of the same step agree to fp32 round-off of a single rounding -- not to "whatever order the
    atomics arrived in".  (Round 1's fp32 merges moved the input-BN gradient by up to 1.3e-2 between identical runs.)"""
    mt, B = "cnn_L3_melspec2", 4
    w_np = O.init_weights(mt, seed=11, randomize_bn=True)
    video, audio, label = O.synthetic_batch(B, seed=910)
    runs = []
    for _ in range(3):
        eng = _engine(mt, B, "f32", training=True, weights=w_np)
        eng.forward_backward(video, audio, label)
        runs.append(eng.get_grads())
        eng.close()
    worst = max(rel_l2(runs[i][k], runs[0][k]) for i in (1, 2) for k in runs[0]
                if np.abs(runs[0][k]).max() > 1e-4)
    print("run-to-run worst rel-L2 over gradient tensors: %.3g" % worst)
    assert worst <= 1e-5, worst


def test_input_bn_gradient_fallback_when_gamma_is_zero():
    """Input-BN gradients with a zero gamma.  (Parity mode computes sum(da), sum(da*xhat) directly from dz with fp64
    accumulation; the throughput mode derives them from the first layer's weight gradient divided by gamma and falls
    back to the direct kernel when a gamma is ~0 -- see the bf16 variant below.)"""
    mt, B = "cnn_L3_kapredbinputbn", 2
    w_np = O.init_weights(mt, seed=9, randomize_bn=True)
    # only one of the three vision channels: zeroing the single audio channel would make the whole audio tower's
    # input constant and every later BatchNorm amplify pure round-off (an ill-conditioned comparison)
    w_np["vision/bn0/gamma"][1] = 0.0
    video, audio, label = O.synthetic_batch(B, seed=404)
    vf, af = _oracle_inputs(video, audio)
    w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
    grads, _, _ = O.compute_grads(vf, af, torch.from_numpy(label), w, mt, F64)
    eng = _engine(mt, B, "f32", training=True, weights=w_np)
    eng.forward_backward(video, audio, label)
    got = eng.get_grads()
    for name in ("audio/bn0/gamma", "audio/bn0/beta", "vision/bn0/gamma", "vision/bn0/beta"):
        ref = grads[name].numpy()
        assert rel_l2(got[name], ref) <= 1e-2 or np.abs(got[name] - ref).max() <= 1e-4, (name, got[name], ref)


def test_data_parallel_replicas_match_oracle_replica_by_replica():
    """Row (e): N replicas x B/N samples with PER-REPLICA BatchNorm statistics (training_utils.py:121-170 semantics, no
    sync-BN).  Two replicas of four samples are emulated on one GPU: each slice runs forward/backward with
    global_batch = 8, the two gradient arenas are summed (what the NCCL all-reduce does) and compared with the oracle
    evaluated replica by replica and gradient-averaged over the global batch.  Same bar as the single-replica step:
    1e-2 relative L2 per tensor (the parity-mode reductions are merged in fp64, so the result does not depend on the
    arrival order of the partial sums)."""
    mt, G, R = "cnn_L3_melspec2", 8, 2
    w_np = O.init_weights(mt, seed=11, randomize_bn=True)
    video, audio, label = O.synthetic_batch(G, seed=909)
    eng = _engine(mt, G // R, "f32", training=True, weights=w_np)
    summed, ref = {}, {}
    loss_dev = loss_ref = 0.0
    for r in range(R):
        sl = slice(r * (G // R), (r + 1) * (G // R))          # contiguous slices, training_utils.py:121-133
        eng.forward_backward(video[sl], audio[sl], label[sl], global_batch=G)
        loss_dev += eng.metrics()["loss"] * (G // R) / G
        for k, v in eng.get_grads().items():
            summed[k] = summed.get(k, 0.0) + v.astype(np.float64)
        vf, af = _oracle_inputs(video[sl], audio[sl])
        w = O.to_torch(w_np, dtype=torch.float64, requires_grad=True)
        grads, out, _ = O.compute_grads(vf, af, torch.from_numpy(label[sl]), w, mt, F64)
        loss_ref += float(out["loss"]) * (G // R) / G
        for k, g in grads.items():
            g = g.numpy()
            if k.endswith("/kernel"):
                g = g - 2e-5 * w_np[k]                          # the l2 term is applied once, inside Adam
            ref[k] = ref.get(k, 0.0) + g * ((G // R) / G)       # mean over the slice -> share of the global mean
    assert abs(loss_dev - loss_ref) <= 1e-4 * max(1.0, abs(loss_ref)), (loss_dev, loss_ref)
    rows, bad = [], []
    for k in ref:
        err, max_abs = rel_l2(summed[k], ref[k]), float(np.abs(summed[k] - ref[k]).max())
        rows.append((k, round(err, 5), max_abs))
        if not (err <= 1e-2 or max_abs <= 1e-4):
            bad.append((k, round(err, 5), max_abs))
    print("replica test: loss (device, oracle)", loss_dev, loss_ref, "worst", sorted(rows, key=lambda r: -r[1])[:5])
    assert not bad, bad


def test_train_steps_from_host_decrease_loss():
    """BASELINE config 1 on the device path: cnn_L3_orig, batch 4, 8 steps from host buffers."""
    mt = "cnn_L3_orig"
    video, audio, label = O.synthetic_batch(4, seed=11)
    eng = _engine(mt, 4, "f32", training=True)
    losses = [eng.train_step_host(video, audio, label, 1e-4)["loss"] for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
