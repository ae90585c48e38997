"""Per-camera sharding of spatial cross-attention's sampling step across GPUs.

The SCA output is a sum over cameras of per-camera MSDA outputs weighted by ``bev_mask[cam, q]``
(reference: det2trt/models/modules/spatial_cross_attention.py:254-270 — ``slots = (queries * bev_mask).sum(0)``), and
each camera's MSDA touches only that camera's value stack, so the work splits into independent (camera, query-tile)
units. Every rank runs the sm_100a MSDA kernel on its units, folds them into a local BEV accumulator ``[nq, heads*ch]``
and ONE collective (all-reduce, NCCL over NVLink on GPUs / gloo in the CPU tests) produces the summed accumulator.
6 cameras do not divide 4 or 8 ranks, hence query tiles: the unit count is lcm-friendly (12 units for 4 ranks,
24 for 8) so every rank gets the same amount of work.
"""
import math
from dataclasses import dataclass
from typing import Callable, List, Sequence

import torch


@dataclass(frozen=True)
class Unit:
    cam: int
    q0: int
    q1: int


def plan_units(num_cams: int, num_query: int, world_size: int, align: int = 8) -> List[List[Unit]]:
    """Round-robin assignment of (camera, query-tile) units: returns per-rank unit lists, equal counts per rank.
    Tiles per camera = world_size / gcd(num_cams, world_size); tile edges are multiples of ``align`` queries."""
    if world_size < 1 or num_cams < 1 or num_query < 1:
        raise ValueError("world_size, num_cams and num_query must be positive")
    tiles = world_size // math.gcd(num_cams, world_size)
    edges = [min(num_query, ((num_query * t // tiles + align - 1) // align) * align) for t in range(tiles)] + [num_query]
    units = [Unit(c, edges[t], edges[t + 1]) for c in range(num_cams) for t in range(tiles) if edges[t + 1] > edges[t]]
    per_rank: List[List[Unit]] = [[] for _ in range(world_size)]
    # camera-major order keeps one camera's tiles on as few ranks as possible: rank r takes a contiguous run
    n = len(units)
    for r in range(world_size):
        per_rank[r] = units[n * r // world_size : n * (r + 1) // world_size]
    return per_rank


def plan_chunked(num_cams: int, num_query: int, world_size: int, chunks: int, align: int = 8) -> List[List[List[Unit]]]:
    """Splits the query range into ``chunks`` pieces and plans each piece like ``plan_units``: result[chunk][rank] is
    that rank's unit list for that query chunk. Every rank has the same amount of work in every chunk, so the
    all-reduce of chunk k can overlap the kernels of chunk k+1 without stragglers."""
    chunks = max(1, min(chunks, num_query // align if num_query >= align else 1))
    edges = [min(num_query, ((num_query * c // chunks + align - 1) // align) * align) for c in range(chunks)] + [num_query]
    out = []
    for c in range(chunks):
        q0, q1 = edges[c], edges[c + 1]
        if q1 <= q0:
            continue
        per_rank = plan_units(num_cams, q1 - q0, world_size, align)
        out.append([[Unit(u.cam, u.q0 + q0, u.q1 + q0) for u in units] for units in per_rank])
    return out


def plan_chunk_bounds(plan) -> List[tuple]:
    """(q_lo, q_hi) of every chunk of a ``plan_chunked`` result, over all ranks."""
    return [(min(u.q0 for units in chunk for u in units), max(u.q1 for units in chunk for u in units)) for chunk in plan]


def merge_units(units: Sequence[Unit]) -> List[Unit]:
    """Coalesces adjacent query tiles of the same camera (fewer, larger kernel launches)."""
    out: List[Unit] = []
    for u in sorted(units, key=lambda x: (x.cam, x.q0)):
        if out and out[-1].cam == u.cam and out[-1].q1 == u.q0:
            out[-1] = Unit(u.cam, out[-1].q0, u.q1)
        else:
            out.append(u)
    return out


def group_cameras(units: Sequence[Unit]):
    """Batches consecutive cameras that cover the same query range into one launch: (cam0, cam1, q0, q1) tuples."""
    groups = []
    for u in merge_units(units):
        if groups and groups[-1][1] == u.cam and groups[-1][2:] == (u.q0, u.q1):
            groups[-1] = (groups[-1][0], u.cam + 1, u.q0, u.q1)
        else:
            groups.append((u.cam, u.cam + 1, u.q0, u.q1))
    return groups


class ShardedSCASampler:
    """Holds one rank's slice of the SCA sampling inputs and produces the all-reduced BEV accumulator.

    ``units`` is either one unit list (single collective at the end) or a list of per-chunk unit lists from
    ``plan_chunked`` (the all-reduce of query chunk k is issued asynchronously right after chunk k's kernels and runs
    on the communication stream while chunk k+1 computes). ``msda`` is the operator to run per launch group (the
    product passes ``multi_scale_deformable_attn``; the CPU tests pass a checker so the host-side plumbing — planning,
    slicing, masking, the collective — is covered without a GPU). ``fused_sca`` (optional,
    ``multi_scale_deformable_attn_sca``) folds MSDA + bev_mask camera-sum into one kernel writing the accumulator.
    """

    def __init__(self, units, num_query: int, msda: Callable, group=None, accum_dtype: torch.dtype = torch.float32,
                 fused_sca: Callable = None, chunk_bounds=None, wire_dtype: torch.dtype = None):  # fmt: skip
        chunked = len(units) > 0 and isinstance(units[0], (list, tuple))
        self.chunk_units = [merge_units(u) for u in units] if chunked else [merge_units(units)]
        self.num_query = num_query
        self.msda = msda
        self.group = group
        self.accum_dtype = accum_dtype
        self.fused_sca = fused_sca if accum_dtype == torch.float32 else None
        # dtype on the wire: every rank accumulates in accum_dtype (fp32); with wire_dtype=float16 the partial sums are
        # rounded once to fp16 for the collective (half the NVLink bytes) and the reduced result is widened back
        self.wire_dtype = wire_dtype if wire_dtype not in (None, accum_dtype) else None
        self._wire = None
        # the accumulator slice every rank reduces for chunk k — must be identical on all ranks (plan_chunk_bounds)
        self.chunk_bounds = list(chunk_bounds) if chunk_bounds is not None else None
        if chunked and self.chunk_bounds is None:
            raise ValueError("chunked unit lists need chunk_bounds (use plan_chunk_bounds on the full plan)")
        self.chunks = []  # per chunk: (q_lo, q_hi, [(cam0, cam1, q0, q1)], [(value, ref, off, logits, mask)])
        self.shapes = None
        self.accum = None
        self._graph = None

    def load(self, value, shapes, ref, off, logits, bev_mask, device):
        """Copies this rank's units out of full-size host tensors (a real model produces them in place)."""
        self.shapes = shapes.to(device)
        self.chunks = []
        cams = {}
        for units in self.chunk_units:
            groups = group_cameras(units)  # (cam0, cam1, q0, q1): consecutive cameras sharing a query range
            local = []
            for c0, c1, q0, q1 in groups:
                sl = slice(q0, q1)
                if (c0, c1) not in cams:  # one device copy of a camera's value stack, shared by its query tiles
                    cams[(c0, c1)] = value[c0:c1].to(device).contiguous()
                local.append((cams[(c0, c1)],) + tuple(
                    t.to(device).contiguous() for t in (ref[c0:c1, sl], off[c0:c1, sl], logits[c0:c1, sl],
                                                        bev_mask[c0:c1, sl])))  # fmt: skip
            lo, hi = self.chunk_bounds[len(self.chunks)] if self.chunk_bounds is not None else (0, self.num_query)
            self.chunks.append((lo, hi, groups, local))
        M, C = value.shape[2], value.shape[3]
        self.accum = torch.zeros(self.num_query, M * C, dtype=self.accum_dtype, device=device)
        if self.wire_dtype is not None:
            self._wire = torch.empty(self.num_query, M * C, dtype=self.wire_dtype, device=device)
        return self

    def _compute(self, groups, local):
        for (c0, c1, q0, q1), (v, r, o, w, mask) in zip(groups, local):
            if self.fused_sca is not None:
                self.fused_sca(v, self.shapes, r, o, w, mask, self.accum[q0:q1])
                continue
            out = self.msda(v, self.shapes, r, o, w)  # [cams, q, M, C]
            weighted = out.reshape(c1 - c0, q1 - q0, -1).to(self.accum_dtype) * mask.to(self.accum_dtype)
            self.accum[q0:q1] += weighted.sum(0)

    def capture(self, warmup: int = 3):
        """Captures one full step (accumulator memset, this rank's kernels, wire conversion, the NCCL all-reduce) into a
        CUDA graph; ``step()`` then replays it. The step is 0.2-0.7 ms of device work issued as five to ten small
        launches plus a collective, so the host launch path and the inter-stream hand-offs around the collective are
        a visible share of it; the replay removes them. Every rank must call this (it runs collectives)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager(True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._step_eager(True)
        self._graph = graph
        return self

    def step(self, reduce: bool = True):
        """One SCA sampling step on this rank: kernels for the local units, masked camera-sum, all-reduce (one
        collective, or one per query chunk overlapped with the next chunk's kernels). Replays the captured graph when
        ``capture()`` has been called."""
        if reduce and self._graph is not None:
            self._graph.replay()
            return self.accum
        return self._step_eager(reduce)

    def _step_eager(self, reduce: bool = True):
        import torch.distributed as dist

        do_reduce = reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        self.accum.zero_()
        works = []
        for lo, hi, groups, local in self.chunks:
            self._compute(groups, local)
            if do_reduce and hi > lo:
                buf = self.accum[lo:hi]
                if self._wire is not None:
                    buf = self._wire[lo:hi]
                    buf.copy_(self.accum[lo:hi])
                works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group,
                                             async_op=len(self.chunks) > 1))  # fmt: skip
        for w in works:
            if w is not None:
                w.wait()
        if do_reduce and self._wire is not None:
            self.accum.copy_(self._wire)
        return self.accum
