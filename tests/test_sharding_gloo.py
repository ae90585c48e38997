"""Host-side logic of the per-camera sharding on CPU: planning, slicing, masked camera-sum and the
collective, with world_size-2 gloo processes. The per-unit operator is injected (the CPU oracle here, the sm_100a
kernel in the product), so no CUDA is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevformer_tensorrt_b200.sharding import (ShardedSCASampler, group_cameras, plan_chunk_bounds, plan_chunked,
                                              plan_units)
from bevformer_tensorrt_b200.workloads import MSDAConfig, bev_reference_points_cam, camera_ring_lidar2img, make_msda_inputs

CFG = MSDAConfig("shard_case", 6, 20 * 20, 8, 32, ((12, 20), (6, 10)), 8, 4, (20, 20))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 6, 8])
def test_plan_covers_every_camera_query_pair_once_and_is_balanced(world):
    plan = plan_units(6, 40000, world)
    cover = np.zeros((6, 40000), np.int32)
    for units in plan:
        for u in units:
            cover[u.cam, u.q0 : u.q1] += 1
    assert (cover == 1).all()
    loads = [sum(u.q1 - u.q0 for u in units) for units in plan]
    assert max(loads) - min(loads) <= 8 * 6
    assert sum(len(group_cameras(u)) for u in plan) <= 2 * world + 6


@pytest.mark.parametrize("world,chunks", [(2, 4), (4, 4), (8, 4), (8, 1)])
def test_chunked_plan_is_balanced_per_chunk(world, chunks):
    plan = plan_chunked(6, 40000, world, chunks)
    cover = np.zeros((6, 40000), np.int32)
    for chunk in plan:
        loads = [sum(u.q1 - u.q0 for u in units) for units in chunk]
        assert max(loads) - min(loads) <= 8 * 6
        for units in chunk:
            for u in units:
                cover[u.cam, u.q0 : u.q1] += 1
    assert (cover == 1).all()
    bounds = plan_chunk_bounds(plan)
    assert bounds[0][0] == 0 and bounds[-1][1] == 40000
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(len(bounds) - 1))


def test_plan_rejects_nonsense():
    with pytest.raises(ValueError):
        plan_units(6, 100, 0)


def _oracle_op(value, shapes, ref, off, logits):
    from oracle import msda as omsda

    return torch.from_numpy(omsda.msda_f32(value.numpy(), shapes.numpy(), ref.numpy(), off.numpy(), logits.numpy()))


def _full_reference():
    value, shapes, ref, off, logits = make_msda_inputs(CFG, "G", 3, torch.float32)
    _, mask = bev_reference_points_cam(CFG.bev_hw, camera_ring_lidar2img(6))
    out = _oracle_op(value, shapes, ref, off, logits)  # [6, nq, 8, 32]
    slots = (out.reshape(6, CFG.num_query, -1) * mask).sum(0)  # spatial_cross_attention.py:270
    return (value, shapes, ref, off, logits, mask), slots


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    (value, shapes, ref, off, logits, mask), want = _full_reference()
    plan = plan_units(6, CFG.num_query, world)
    s = ShardedSCASampler(plan[rank], CFG.num_query, _oracle_op).load(value, shapes, ref, off, logits, mask, "cpu")
    got = s.step()
    err = float((got - want).abs().max())
    # chunked plan: one (async) all-reduce per query chunk, overlapped with the next chunk's work
    cplan = plan_chunked(6, CFG.num_query, world, 3)
    s2 = ShardedSCASampler([c[rank] for c in cplan], CFG.num_query, _oracle_op,
                           chunk_bounds=plan_chunk_bounds(cplan)).load(value, shapes, ref, off, logits, mask, "cpu")
    err = max(err, float((s2.step() - want).abs().max()))
    # fp16 on the wire: partial sums rounded once to fp16 for the collective
    s3 = ShardedSCASampler(plan[rank], CFG.num_query, _oracle_op, wire_dtype=torch.float16).load(
        value, shapes, ref, off, logits, mask, "cpu")
    assert float((s3.step() - want).abs().max()) < 2e-3
    q.put((rank, err, float(want.abs().max())))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])  # 4: cameras AND query tiles are split (two tiles per camera)
def test_world_size_2_gloo_matches_single_process(world):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, err, mag in res:
        assert mag > 0.01 and err < 1e-5, (rank, err, mag)


def test_single_process_sampler_equals_masked_camera_sum():
    (value, shapes, ref, off, logits, mask), want = _full_reference()
    for world in (1, 3, 4, 8):
        total = torch.zeros_like(want)
        for rank in range(world):
            s = ShardedSCASampler(plan_units(6, CFG.num_query, world)[rank], CFG.num_query, _oracle_op)
            total += s.load(value, shapes, ref, off, logits, mask, "cpu").step(reduce=False)
        assert (total - want).abs().max() < 1e-5
