"""Host-side logic of the multi-GPU paths: which rank owns which unit of work, how results are
exchanged (torch.distributed: NCCL on the GPUs, gloo in the CPU tests) and how per-rank lists are merged.

  * batched independent scan-pair ICP (config C5): pair p -> rank p % world, one all-gather of fixed-size result
    records at the end of a batch, never per iteration;
  * Scan Context database search (config C4): entry i -> rank i % world (sb_loop_create(rank, world)); each rank
    returns its local candidates, one all-gather, then the same (distance, entry) merge on every rank — the order of
    std::sort on pair<double,int> at loop_closure.hpp:92;
  * loop-closure verification: candidate j is verified by the rank that owns its cloud; acceptance walks the merged
    order until max_candidates successes (loop_closure.hpp:95-122).
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD = 20  # T[16], final_error, num_iterations, converged, status


def owner(unit, world):
    return int(unit) % int(world)


def shard_units(n_units, rank, world):
    """Indices of the units (pairs / database entries) owned by `rank`."""
    return np.arange(rank, n_units, world, dtype=np.int64)


def pack_results(results):
    """ICP results -> (n, RECORD) float64 array."""
    rec = np.zeros((len(results), RECORD))
    for i, r in enumerate(results):
        rec[i, :16] = np.asarray(r.transformation).reshape(16)
        rec[i, 16:] = (r.final_error, r.num_iterations, float(r.converged), r.status)
    return rec


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_records(local, n_units):
    """All-gathers the per-rank record blocks of a round-robin sharded batch and returns them in unit order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    width = local.shape[1] if local.ndim == 2 else RECORD
    per = (n_units + world - 1) // world
    pad = np.zeros((per, width))
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(_device())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    full = np.zeros((n_units, width))
    for r in range(world):
        ids = shard_units(n_units, r, world)
        full[ids] = out[r].cpu().numpy()[:len(ids)]
    return full


def merge_candidates(dist_lists, entry_lists):
    """Merges per-rank (distance, entry) lists into the single ascending list of loop_closure.hpp:92."""
    d = np.concatenate([np.asarray(x, dtype=np.float64) for x in dist_lists]) if dist_lists else np.zeros(0)
    e = np.concatenate([np.asarray(x, dtype=np.int64) for x in entry_lists]) if entry_lists else np.zeros(0, np.int64)
    order = np.lexsort((e, d))
    return d[order], e[order]


def all_gather_candidates(local_dist, local_entry, capacity):
    """Every rank contributes up to `capacity` local candidates; every rank gets the merged list."""
    world = dist.get_world_size()
    buf = np.full((capacity, 2), np.inf)
    m = min(len(local_dist), capacity)
    buf[:m, 0] = local_dist[:m]
    buf[:m, 1] = local_entry[:m]
    t = torch.from_numpy(buf).to(_device())
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    ds, es = [], []
    for o in out:
        a = o.cpu().numpy()
        keep = np.isfinite(a[:, 0])
        ds.append(a[keep, 0])
        es.append(a[keep, 1].astype(np.int64))
    return merge_candidates(ds, es)


def accept_in_order(merged_entries, converged, fitness, fitness_threshold, max_candidates):
    """loop_closure.hpp:95-122: walk the candidates in order, count only acceptances, stop at max_candidates."""
    accepted = []
    for j, e in enumerate(merged_entries):
        if len(accepted) >= max_candidates:
            break
        if converged[j] and fitness[j] < fitness_threshold:
            accepted.append(int(e))
    return accepted


def sharded_detect(det, rank, world, top_k=10):
    """One detect() of a loop-closure database sharded over `world` ranks (sb_loop_create(rank, world): entry i lives
    on rank i % world; every rank holds the newest entry, the query).  Every rank calls this collectively:
      1. local Scan Context search: the rank's best k candidates, selected on the device (the job's best k are
         inside the union of the local lists);
      2. all-gather of the local lists, the same (distance, entry) merge on every rank (loop_closure.hpp:92);
      3. each rank verifies by ICP the candidates it owns (their clouds are local) — the whole list in parallel;
      4. all-reduce of the verification records (each candidate is verified by exactly one rank); acceptance walks the
         merged order until max_candidates successes (loop_closure.hpp:95-122: `verified` counts successes only).
    The reference keeps walking down the candidate list until it has max_candidates acceptances; so if the best
    k = top_k candidates yield fewer and more candidates exist, k doubles and only the new ones are verified.
    Returns (merged distances, merged entries, accepted entries, records) with records[j] = [verified, converged,
    icp_fitness, match_frame, T(16)] for candidate j of the merged list."""
    k = max(int(top_k), 1)
    known = {}          # entry -> record, from earlier rounds
    while True:
        cd, ce = det.candidates_local(capacity=k)
        if world > 1:
            md, me = all_gather_candidates(cd, ce, k)
        else:
            md, me = merge_candidates([cd], [ce])
        md, me = md[:k], me[:k]
        rec = np.zeros((k, 20))
        new = [j for j, e in enumerate(me) if int(e) not in known]
        mine = np.array([j for j in new if owner(me[j], world) == rank], dtype=np.int64)
        if len(mine):
            res, conv = det.verify_entries(me[mine], md[mine])
            for i, j in enumerate(mine):
                rec[j, :4] = (1.0, float(conv[i]), res[i]["icp_fitness"], res[i]["match_frame"])
                rec[j, 4:] = np.asarray(res[i]["transform"]).reshape(16)
        if world > 1:
            t = torch.from_numpy(rec).to(_device())
            dist.all_reduce(t)           # every new candidate is verified by exactly one rank: the sum is a gather
            rec = t.cpu().numpy()
        for j, e in enumerate(me):
            if int(e) in known:
                rec[j] = known[int(e)]
            else:
                known[int(e)] = rec[j].copy()
        acc = accept_in_order(me, rec[:len(me), 1] > 0.5, rec[:len(me), 2], det.cfg.icp_fitness_threshold,
                              det.cfg.max_candidates)
        if len(acc) >= det.cfg.max_candidates or len(me) < k:   # enough acceptances, or no candidate left
            return md, me, acc, rec[:len(me)]
        k *= 2
