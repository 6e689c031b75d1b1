"""Candidate-level data parallelism: every rank holds the full tables and
takes a contiguous slice of the candidate batch, balanced by work; the only exchange is one
all-gather of the per-candidate results (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(costs, world):
    """Split range(len(costs)) into `world` contiguous slices of near-equal total cost.

    costs[i] ~ work of candidate i (e.g. facts * epochs).  Returns world+1 boundaries."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n == 0:
        return [0] * (world + 1)
    csum = np.concatenate([[0.0], np.cumsum(np.maximum(costs, 1e-9))])
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left")
    bounds = [0] + [int(min(max(c, 0), n)) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def gather_results(local, bounds, group=None):
    """All-gather variable-length per-candidate results ([n_local, k] float tensor) into rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    k = local.shape[1] if local.dim() > 1 else 1
    width = max(bounds[i + 1] - bounds[i] for i in range(world))
    pad = torch.zeros((width, k), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local.view(-1, k)
    out = torch.empty((world, width, k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.unsqueeze(0), group=group)
    return torch.cat([out[r, : bounds[r + 1] - bounds[r]] for r in range(world)], 0)


class ShardedEngine:
    """Wraps a Necessary/SufficientPostTrainingEngine for one-process-per-GPU runs: `compute_relevances` post-trains the
    slice of rules owned by this rank and all-gathers the relevances, so every rank returns the full list.

    Every rank draws the random numbers of EVERY candidate (cheap, host side: `owned=` of the engine), so the torch /
    numpy / CUDA generators end exactly where the single-process run leaves them, the per-candidate generator snapshots
    the builder's early-stop rewind needs exist on every rank, and the selected explanations are those of one GPU.
    The homologous (base) mimic of a prediction is post-trained redundantly on every rank.
    Everything else (compute_relevance of a single candidate, select_entities_to_convert, caches) is the engine's."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group

    def __getattr__(self, name):  # dataset, model, hp, set_cache, select_entities_to_convert, entities_to_convert, ...
        return getattr(self.engine, name)

    def compute_relevance(self, pred, rule):
        return self.engine.compute_relevance(pred, rule)

    def compute_relevances(self, pred, rules, snapshots=False):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        bounds = shard_bounds([len(r) + 1 for r in rules], world)
        res = self.engine.compute_relevances(pred, rules, snapshots=snapshots, owned=(bounds[rank], bounds[rank + 1]))
        local, snaps = res if snapshots else (res, None)
        dev = "cuda" if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor(local, dtype=torch.float64, device=dev).view(-1, 1)
        rels = gather_results(t, bounds, self.group).view(-1).tolist()
        return (rels, snaps) if snapshots else rels
