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

from bevformer_tensorrt_b200.sharding import (GroupedSCASampler, ShardedSCASampler, choose_camera_groups, group_cameras,
                                              plan_chunk_bounds, plan_chunked, plan_grid, plan_units)
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


# ---------------------------------------------------------------------------------------------------------------
# round 2: camera-group x query-tile grid, owner-sliced reduce-scatter
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 6, 8, 12])
def test_grid_plan_covers_and_owns_every_query_once(world):
    plan = plan_grid(6, 40000, world)
    A = choose_camera_groups(6, world)
    assert len(plan) == world and all(s.groups == A and s.groups * s.tiles == world for s in plan)
    cover = np.zeros((6, 40000), np.int32)
    owned = np.zeros(40000, np.int32)
    for s in plan:
        cover[s.cam0 : s.cam1, s.tile_queries().numpy()] += 1
        owned[s.own_queries().numpy()] += 1
        assert s.q0 <= s.own0 <= s.own1 <= s.q1 and s.rank in s.peers and len(s.peers) == A
        assert (s.own0 - s.q0) % 4 == 0  # float4 granularity of the exchange kernel at width 256
        for r in s.peers:  # the ranks of a camera group share the query tile
            assert (plan[r].q0, plan[r].q1) == (s.q0, s.q1)
    assert (cover == 1).all() and (owned == 1).all()
    loads = [(s.cam1 - s.cam0) * (s.q1 - s.q0) for s in plan]
    assert max(loads) - min(loads) <= 8 * 6


def test_interleaved_tiles_balance_the_camera_ring():
    """Visible (camera, query) pairs per rank on the synthetic camera ring: contiguous query ranges are badly skewed
    (a camera sees a compact wedge of the BEV), interleaved 8-query blocks are balanced."""
    from bevformer_tensorrt_b200.workloads import bev_reference_points_cam, camera_ring_lidar2img

    _, mask = bev_reference_points_cam((200, 200), camera_ring_lidar2img(6))
    vis = mask[..., 0] > 0
    for world in (4, 8):
        for block, bound in ((0, 1.5), (8, 1.01)):
            plan = plan_grid(6, 40000, world, block=block)
            loads = [int(vis[s.cam0 : s.cam1][:, s.tile_queries()].sum()) for s in plan]
            ratio = max(loads) / (sum(loads) / world)
            assert (ratio > bound) if block == 0 else (ratio < bound), (world, block, ratio)


def test_grid_plan_traffic_is_minimal_for_even_worlds():
    for world in (2, 4, 8):
        assert choose_camera_groups(6, world) == 2  # pairs: a rank sends half of its tile, 1/world of the accumulator
    assert choose_camera_groups(6, 3) == 3 and choose_camera_groups(6, 5) == 1 and choose_camera_groups(6, 1) == 1
    with pytest.raises(ValueError):
        plan_grid(6, 40000, 8, groups=3)


def _oracle_fused(value, shapes, ref, off, logits, mask, accum):
    out = _oracle_op(value, shapes, ref, off, logits)
    accum += (out.reshape(value.shape[0], accum.shape[0], -1) * mask.reshape(value.shape[0], accum.shape[0], 1)).sum(0)
    return accum


def _grid_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    (value, shapes, ref, off, logits, mask), want = _full_reference()
    plan = plan_grid(6, CFG.num_query, world)
    s = plan[rank]
    smp = GroupedSCASampler(s, want.shape[1], _oracle_fused).load(value, shapes, ref, off, logits, mask, "cpu")
    smp.connect(plan, "cpu")
    errs = []
    for _ in range(3):  # repeated steps: the partial is re-zeroed, the result does not drift
        got = smp.step()
        errs.append(float((got - want[s.own_queries()]).abs().max()))
    q.put((rank, max(errs), float(want[s.own_queries()].abs().max()), got.shape[0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_grid_sampler_gloo_matches_single_process(world):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sum(r[3] for r in res) == CFG.num_query  # the owned slices tile the BEV
    for rank, err, mag, _ in res:
        assert err < 1e-5, (rank, err, mag)
    assert max(r[2] for r in res) > 0.01


def test_grid_sampler_single_rank_is_the_fused_op():
    (value, shapes, ref, off, logits, mask), want = _full_reference()
    plan = plan_grid(6, CFG.num_query, 1)
    smp = GroupedSCASampler(plan[0], want.shape[1], _oracle_fused).load(value, shapes, ref, off, logits, mask, "cpu")
    smp.connect(plan, "cpu")
    assert (smp.step() - want).abs().max() < 1e-5
