"""GPU tests of the data-parallel step INSIDE the library (l3_dp_*: NCCL bound at run time, gradient buckets all-reduced
on a communication stream while the backward pass runs) -- the replacement of l3embedding/training_utils.py:21-170.
Needs >= 2 GPUs in the box (`gpurun --gpus 2`); skipped on a single-GPU box.  One process per GPU, as in production.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MT = "cnn_L3_melspec2"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_exchange(rank, world, port, dtype, q):
    import torch.distributed as dist
    from oracle import l3_oracle as O
    from l3embedding_b200 import dp
    from l3embedding_b200.engine import Engine
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        per, G = 4, 4 * world
        w_np = O.init_weights(MT, seed=11, randomize_bn=True)
        video, audio, label = O.synthetic_batch(G, seed=909)
        sl = dp.replica_slice(G, rank, world)
        v, a, l = (np.ascontiguousarray(x[sl]) for x in (video, audio, label))
        # (1) the local gradient of this rank's slice, no exchange
        loc = Engine(MT, per, dtype, training=True, weights=w_np, device="cuda:%d" % rank)
        loc.forward_backward(v, a, l, global_batch=G)
        g_local = loc.grads.clone()
        m_local = loc.metrics()
        loc.close()
        # (2) the same step with the library's overlapped exchange
        eng = Engine(MT, per, dtype, training=True, weights=w_np, device="cuda:%d" % rank)
        par = dp.current(world)
        par.attach(eng)
        assert eng.dp_world == (rank, world)
        eng.upload_host(v, a, l)
        eng.forward_backward_staged(per, global_batch=G)
        m = eng.metrics()                       # waits for the collectives; sums over the GLOBAL batch
        g_dp = eng.grads.clone()
        # reference for the exchange: gather every rank's local gradient through gloo and add in rank order
        parts = [torch.empty_like(g_local, device="cpu") for _ in range(world)]
        dist.all_gather(parts, g_local.cpu())
        loc_ce = [None] * world
        dist.all_gather_object(loc_ce, (m_local["ce_sum"], m_local["correct"]))
        want = parts[0].clone()
        for p in parts[1:]:
            want += p
        # (3) one full data-parallel step: parameters must stay identical on all ranks
        eng.upload_host(v, a, l)
        m2 = eng.dp_train_step_staged(per, G, 1e-4)
        params = eng.params.cpu()
        eng.dp_average_bn_state()
        torch.cuda.synchronize()
        bn = eng.bn_state.cpu()
        q.put(dict(rank=rank, g_dp=g_dp.cpu().numpy(), want=want.numpy(), ce=m["ce_sum"], correct=m["correct"],
                   batch=m["batch"], loc_ce=loc_ce, params=params.numpy(), bn=bn.numpy(), loss2=m2["loss"]))
        eng.close()
    finally:
        dist.destroy_process_group()


def _run(target, world, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r["rank"])


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_library_gradient_exchange_equals_sum_of_local_gradients(dtype):
    """Two ranks x 4 samples, per-replica BN statistics.  The gradient arena after the library's bucketed, overlapped
    all-reduce equals the sum of the ranks' local gradient arenas: bit for bit in parity mode (two addends: fp32 a+b in
    any order; the parity-mode step is run-to-run deterministic), to fp32 round-off of the bf16 step's atomics
    otherwise.  Loss / accuracy sums are global, parameters stay identical across ranks after Adam, and the BN moving
    statistics are identical after l3_dp_average_bn_state."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _run(_worker_exchange, world, dtype)
    for r in res:
        if dtype == "f32":
            assert np.array_equal(r["g_dp"], r["want"])
        else:
            scale = np.abs(r["want"]).max()
            assert np.abs(r["g_dp"] - r["want"]).max() <= 2e-3 * scale
        assert abs(r["ce"] - sum(c for c, _ in r["loc_ce"])) <= 1e-4 * max(1.0, abs(r["ce"]))
        assert r["correct"] == sum(k for _, k in r["loc_ce"]) and r["batch"] == 4 * world
        assert np.isfinite(r["loss2"])
    assert np.array_equal(res[0]["g_dp"], res[1]["g_dp"])            # the all-reduce result is the same everywhere
    assert np.array_equal(res[0]["params"], res[1]["params"])
    assert np.array_equal(res[0]["bn"], res[1]["bn"])


def _worker_fit(rank, world, port, q):
    import torch.distributed as dist
    from l3embedding_b200 import model as M
    from l3embedding_b200.synthetic import synthetic_batch
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        m, _, _ = M.MODELS[MT](num_gpus=world)
        m.configure(dtype="bf16")
        m.compile(M.Adam(lr=1e-4), loss="categorical_crossentropy", metrics=["accuracy"])
        batches = [synthetic_batch(8, seed=50 + i) for i in range(3)]     # the same global batches on every rank

        def gen():
            i = 0
            while True:
                v, a, l = batches[i % 3]
                i += 1
                yield [v, a], l
        h = m.fit_generator(gen(), steps_per_epoch=3, epochs=2, validation_data=gen(), validation_steps=2, verbose=0)
        w = m.named_weights()
        q.put(dict(rank=rank, hist=h.history, digest={k: float(np.abs(v).sum()) for k, v in w.items()}))
    finally:
        dist.destroy_process_group()


def test_fit_generator_data_parallel_keeps_replicas_identical():
    """The keras-style loop with num_gpus = 2 (gpu_wrapper / multi_gpu_model, model.py:184-195): every rank slices the
    same global batches (training_utils.py:121-133), the library exchanges gradients, BN statistics are averaged before
    validation, validation is sharded -- so both ranks report the same epoch logs and hold the same weights."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _run(_worker_fit, world)
    h0, h1 = res[0]["hist"], res[1]["hist"]
    assert set(h0) == {"loss", "acc", "val_loss", "val_acc"} and len(h0["loss"]) == 2
    for k in h0:
        assert np.allclose(h0[k], h1[k], rtol=0, atol=1e-6), k
        assert np.isfinite(h0[k]).all()
    assert res[0]["digest"] == res[1]["digest"]
