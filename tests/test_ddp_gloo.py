"""Data-parallel path on CPU: world_size-2 gloo processes, bucketed gradient all-reduce through GradSink.

Each rank runs the engine (CUDA kernels replaced by tests/cpu_emulation.py) on its own half of a batch; after
`sink.finish()` every rank must hold the average of the two ranks' gradients, which must equal the gradients of
a single-process run over the full batch (the loss is a mean over the batch).  Also checks that several buckets
were actually flushed during backward (overlap plumbing), not one reduction at the end.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load_case


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _install_emulation():
    from bubbleformer_b200 import engine, ops
    from tests import cpu_emulation
    for n in cpu_emulation.ALL:
        setattr(ops, n, getattr(cpu_emulation, n))
    engine.BF16 = torch.float32
    engine.F16 = torch.float32


def _forward(model, x, cond, T):
    from bubbleformer_b200 import engine
    B = x.shape[0]
    _, _, C, H, W = x.shape
    p = model.patch_size
    geom = engine.Geom(B, T, H // p, W // p)
    gb = model.film_embed.gamma_beta(cond)
    X = model.embed.tokens(x.reshape(B * T, C, H, W), gb, T)
    for blk in model.blocks:
        X = blk.tokens(X, geom, lambda p_: p_.detach().reshape(p_.shape[0], -1), None)
    return model.debed.images(X, geom).view(B, T, -1, H, W)


def _loss(y, tgt):
    return ((y - tgt) ** 2).mean()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    _install_emulation()
    from bubbleformer_b200 import get_model
    from bubbleformer_b200.parallel import GradSink
    case = load_case("film_eval_e128", dtype=torch.float32)
    model = get_model("filmavit", time_window=case["T"], **case["cfg"])
    model.load_state_dict(case["sd"], strict=True)
    model.eval()
    sink = GradSink(model, bucket_bytes=64 << 10)
    flushes = []
    orig = sink._flush

    def counting_flush():
        if sink._lo is not None:
            flushes.append((sink._lo, sink._hi))
        orig()
    sink._flush = counting_flush
    B = case["x"].shape[0]
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    sink.begin_step()
    y = _forward(model, case["x"][sl], case["cond"][sl], case["T"])
    _loss(y, case["tgt"][sl]).backward()
    sink.finish()
    grads = {k: p.grad.detach().clone().numpy() for k, p in model.named_parameters()}
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), n_flush=len(flushes), **grads)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_gradsink_allreduce_world2_matches_full_batch(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single-process reference over the full batch (no process group: GradSink world size 1)
    _install_emulation()
    from bubbleformer_b200 import get_model
    case = load_case("film_eval_e128", dtype=torch.float32)
    assert case["x"].shape[0] % world == 0
    model = get_model("filmavit", time_window=case["T"], **case["cfg"])
    model.load_state_dict(case["sd"], strict=True)
    model.eval()
    y = _forward(model, case["x"], case["cond"], case["T"])
    _loss(y, case["tgt"]).backward()
    r0 = np.load(os.path.join(tmp_path, "rank0.npz"))
    r1 = np.load(os.path.join(tmp_path, "rank1.npz"))
    assert int(r0["n_flush"]) >= 2, "gradient buckets should be reduced progressively during backward"
    gn = np.sqrt(sum(float((p.grad.double() ** 2).sum()) for p in model.parameters()))
    for k, p in model.named_parameters():
        a, b = r0[k], r1[k]
        assert np.array_equal(a, b), f"ranks disagree on {k}"
        err = float(np.linalg.norm(a.astype(np.float64) - p.grad.double().numpy())) / gn
        assert err < 1e-4, (k, err)      # fp32 summation order differs between the split and the full batch


# ---------------------------------------------------------------------------------------------
# trajectory-sharded rollout (config 4): independent trajectories, round-robin over ranks, no data-path collective
# ---------------------------------------------------------------------------------------------
N_TRAJ, ROLL_STEPS = 3, 2


def _trajectories(case):
    x0 = case["x"][:1]
    return [x0 * (1.0 + 0.1 * j) for j in range(N_TRAJ)], case["cond"][:1]


def _free_rollout(model, x0, cond, T, steps):
    outs, inp = [], x0
    with torch.no_grad():
        for _ in range(steps):
            inp = _forward(model, inp, cond, T)
            outs.append(inp)
    return torch.stack(outs)


def _rollout_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    _install_emulation()
    from bubbleformer_b200 import get_model
    from bubbleformer_b200.rollout import gather_trajectories, shard_trajectories
    case = load_case("film_eval_e128", dtype=torch.float32)
    model = get_model("filmavit", time_window=case["T"], **case["cfg"])
    model.load_state_dict(case["sd"], strict=True)
    model.eval()
    trajs, cond = _trajectories(case)
    mine = list(shard_trajectories(N_TRAJ, rank, world))
    local = torch.stack([_free_rollout(model, trajs[j], cond, case["T"], ROLL_STEPS) for j in mine])
    full = gather_trajectories(local, N_TRAJ, rank, world)
    np.savez(os.path.join(out_dir, f"roll{rank}.npz"), mine=np.array(mine), full=full.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_rollout_world2_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_rollout_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    _install_emulation()
    from bubbleformer_b200 import get_model
    case = load_case("film_eval_e128", dtype=torch.float32)
    model = get_model("filmavit", time_window=case["T"], **case["cfg"])
    model.load_state_dict(case["sd"], strict=True)
    model.eval()
    trajs, cond = _trajectories(case)
    ref = torch.stack([_free_rollout(model, t, cond, case["T"], ROLL_STEPS) for t in trajs]).numpy()
    r0 = np.load(os.path.join(tmp_path, "roll0.npz"))
    r1 = np.load(os.path.join(tmp_path, "roll1.npz"))
    # the shards partition the trajectories (ragged: 2 + 1), and every rank ends up with all of them in order
    assert sorted(list(r0["mine"]) + list(r1["mine"])) == list(range(N_TRAJ))
    assert list(r0["mine"]) == [0, 2] and list(r1["mine"]) == [1]
    assert np.array_equal(r0["full"], r1["full"])
    assert r0["full"].shape == ref.shape
    err = np.linalg.norm(r0["full"].astype(np.float64) - ref) / np.linalg.norm(ref)
    assert err < 1e-5, err            # a trajectory's result does not depend on which rank computed it
