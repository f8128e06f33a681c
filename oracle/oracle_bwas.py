"""Sequential CPU restatement of the reference's batch weighted A* (BWAS).  TEST INFRASTRUCTURE ONLY.

Follows cpp/parallel_weighted_astar.cpp:138-346 (`parallelWeightedAStar`), the implementation that produced
the reference's shipped results:
  * root pushed to OPEN with cost 0 / heuristic 0 and inserted in CLOSED (:160-162);
  * each iteration pops up to `batch_size` nodes in cost order and STOPS at the first solved node (:177-204);
    the cheapest solved node seen so far is remembered, and the search ends when a goal was already known
    before this iteration and the best popped cost is >= the goal's cost (:205-208), or at once if
    batch_size == 1 (:191-193);
  * all popped nodes are expanded (:217-230), every child counts as generated (:266, root counts 1, :166);
  * CLOSED keeps a child iff its state is new or reached with strictly smaller depth (:243-265);
  * cost = heuristic * (!solved) + weight * depth in float32 (:298); kept children are pushed (:309-319).

Two deliberate, documented choices where the reference is silent or the GPU design differs:
  * heap ties: the reference's std::priority_queue order among equal costs is unspecified; here (and on
    the GPU) ties break towards the smaller node id (Python's heapq path is FIFO, astar.py:66, same effect);
  * `batch_dedup="min"` (GPU behaviour): among children of ONE iteration that are the same state, only the
    one with the smallest (depth, id) can be kept.  `"sequential"` is the reference's child-order loop, in
    which a later duplicate with strictly smaller depth is kept IN ADDITION to the earlier one.
Node ids follow the GPU layout (id = slot * A + move; slot 0 = root; each iteration's parents take
consecutive slots starting at a multiple of `slot_align`) so that tie-breaking is comparable bit for bit.
"""
from __future__ import annotations

import heapq
from typing import Callable, Dict, List, Optional

import numpy as np


def slot_align(state_dim: int, num_moves: int) -> int:
    a = 1
    while (state_dim * num_moves * a) % 16:
        a *= 2
    return a


def bwas(env, start: np.ndarray, heuristic: Callable[[np.ndarray], np.ndarray], weight: float, batch_size: int,
         batch_dedup: str = "sequential", max_iters: Optional[int] = None, keep_trace: bool = False,
         mutate_stored: bool = True) -> Dict:
    A, S = env.num_moves, env.state_dim
    align = slot_align(S, A)
    w32 = np.float32(weight)
    states: Dict[int, np.ndarray] = {0: np.asarray(start, dtype=np.uint8).copy()}
    depth: Dict[int, int] = {0: 0}
    solved: Dict[int, bool] = {0: bool(env.is_solved(states[0][None])[0])}
    parent: Dict[int, int] = {}
    closed: Dict[bytes, List[int]] = {states[0].tobytes(): [0, 0, 0]}   # state -> [best depth, its node id, STORED node id]
    # `(*found)->depth / parentMove / parent = node->...` (:255-257): the node object stored in CLOSED (the first one of its
    # state) takes over the improving node's depth and parent link.  `link[id]` = the node whose (parent, move) the stored node
    # `id` now carries; path reconstruction (:336-341) follows it.  The depth half of the mutation only matters if the stored
    # node is expanded AFTER being improved; its children then carry the improved depth + 1 (restated via `depth[stored]`).
    link: Dict[int, int] = {}
    open_heap = [(np.float32(0.0), 0)]                                   # (cost f32, id)
    next_slot = 1
    nodes_generated = 1
    goal = None            # (cost, id)
    done = False
    iters = 0
    trace = []
    while not done:
        if max_iters is not None and iters >= max_iters:
            break
        if not open_heap:
            break
        # ---- pop (:177-208) ----
        num_pop = min(len(open_heap), batch_size)
        goal_prev = goal is not None
        popped: List[int] = []
        popped_cost: List[np.float32] = []
        for _ in range(num_pop):
            c, nid = heapq.heappop(open_heap)
            popped.append(nid); popped_cost.append(c)
            if solved[nid]:
                if batch_size == 1:
                    goal = (c, nid); done = True
                elif goal is None or goal[0] > c:
                    goal = (c, nid)
                break
        if goal_prev and popped_cost[0] >= goal[0]:
            done = True
        iters += 1
        nodes_generated += len(popped) * A                                # :266 (also on the final iteration)
        if done:
            if keep_trace:
                trace.append({"popped": list(popped), "kept": []})
            break
        # ---- expand (:217-230) ----
        base_slot = -(-next_slot // align) * align
        par = np.stack([states[p] for p in popped])
        ch, _ = env.expand(par)                                           # [n, A, S]
        flat = ch.reshape(-1, S)
        sv = env.is_solved(flat)
        ids = [(base_slot + j) * A + a for j in range(len(popped)) for a in range(A)]
        dep = [depth[p] + 1 for p in popped for _ in range(A)]
        next_slot = base_slot + len(popped)
        # ---- CLOSED (:243-265) ----
        keep = [False] * len(ids)
        if batch_dedup == "sequential":
            for i, nid in enumerate(ids):
                key = flat[i].tobytes()
                e = closed.get(key)
                if e is None:
                    closed[key] = [dep[i], nid, nid]; keep[i] = True
                elif e[0] > dep[i]:
                    e[0] = dep[i]; e[1] = nid; keep[i] = True
                    if mutate_stored:
                        link[e[2]] = nid; depth[e[2]] = dep[i]
        else:
            best: Dict[bytes, int] = {}
            for i in range(len(ids)):
                key = flat[i].tobytes()
                j = best.get(key)
                if j is None or (dep[i], ids[i]) < (dep[j], ids[j]):
                    best[key] = i
            for key, i in best.items():
                e = closed.get(key)
                if e is None:
                    closed[key] = [dep[i], ids[i], ids[i]]; keep[i] = True
                elif e[0] > dep[i]:
                    e[0] = dep[i]; e[1] = ids[i]; keep[i] = True
        kept = [i for i in range(len(ids)) if keep[i]]
        # ---- heuristic + cost (:237, 275-300): values only matter for kept children ----
        for i, nid in enumerate(ids):
            states[nid] = flat[i]; depth[nid] = dep[i]; solved[nid] = bool(sv[i]); parent[nid] = popped[i // A]
        if kept:
            h = np.maximum(np.asarray(heuristic(flat[kept]), dtype=np.float32), np.float32(0.0))
            for k, i in enumerate(kept):
                ns = np.float32(0.0) if sv[i] else np.float32(1.0)
                cost = np.float32(np.float32(h[k] * ns) + np.float32(w32 * np.float32(dep[i])))
                heapq.heappush(open_heap, (cost, ids[i]))
        if keep_trace:
            trace.append({"popped": list(popped), "kept": [ids[i] for i in kept]})
    moves: Optional[List[int]] = None
    if goal is not None:
        moves = []
        nid = goal[1]
        while nid != 0:                                                   # :336-341
            src = link.get(nid, nid)                                      # a stored node that was improved carries the improver's link
            moves.append(src % A)
            nid = parent[src]
        moves.reverse()
    return {"moves": moves, "nodes_generated": nodes_generated, "iterations": iters, "done": done, "links": len(link),
            "goal_id": None if goal is None else goal[1], "trace": trace, "open_size": len(open_heap),
            "closed_size": len(closed)}


def bwas_python(env, start: np.ndarray, heuristic: Callable[[np.ndarray], np.ndarray], weight: float, batch_size: int,
                batch_dedup: str = "sequential", keep_trace: bool = False, max_steps: int = 100000,
                cost_dtype=np.float64) -> Dict:
    """The reference's PYTHON variant (search_methods/astar.py:232-340 `AStar`, :50-90 `Instance`, :99-209) for one instance:
      * root evaluated by the heuristic, pushed to OPEN, NOT put in CLOSED (:244-249, 50-62);
      * every step pops min(batch, |OPEN|) nodes in (cost, push order) order -- heapq FIFO ties (:64-76); solved pops are
        recorded as goal nodes (:73) and the caller stops after a step that found one (:421); ALL popped nodes are expanded;
      * cost = weight * g + h * (!solved) (:196, float64 in the reference; `cost_dtype=np.float32` mirrors the GPU engine,
        identical whenever weight * g is exact, e.g. weight 1.0 / 0.5);
      * CLOSED keeps a child iff unseen or strictly cheaper (:78-90); nodes generated = sum of children (:168);
      * answer = goal node with the smallest path cost (:327-333).
    Node ids as in `bwas` (push order == id order, so FIFO ties == smaller id first).
    Pinned by tests/golden/astar_python_traces.json (the reference's own AStar run in the build container)."""
    A, S = env.num_moves, env.state_dim
    align = slot_align(S, A)
    ct = cost_dtype
    states: Dict[int, np.ndarray] = {0: np.asarray(start, dtype=np.uint8).copy()}
    depth: Dict[int, int] = {0: 0}
    solved: Dict[int, bool] = {0: bool(env.is_solved(states[0][None])[0])}
    parent: Dict[int, int] = {}
    closed: Dict[bytes, List[int]] = {}
    h0 = np.maximum(np.asarray(heuristic(states[0][None]), dtype=np.float32), np.float32(0.0))[0]
    open_heap = [(ct(ct(weight) * ct(0)) + ct(h0) * ct(0.0 if solved[0] else 1.0), 0)]
    next_slot = 1
    nodes_generated = 0
    goals: List[int] = []
    steps = 0
    trace = []
    popped_per_step = []
    while not goals and open_heap and steps < max_steps:
        num_pop = min(len(open_heap), batch_size)
        popped = [heapq.heappop(open_heap)[1] for _ in range(num_pop)]
        goals.extend(p for p in popped if solved[p])
        popped_per_step.append(num_pop)
        steps += 1
        base_slot = -(-next_slot // align) * align
        par = np.stack([states[p] for p in popped])
        ch, _ = env.expand(par)
        flat = ch.reshape(-1, S)
        sv = env.is_solved(flat)
        ids = [(base_slot + j) * A + a for j in range(len(popped)) for a in range(A)]
        dep = [depth[p] + 1 for p in popped for _ in range(A)]
        next_slot = base_slot + len(popped)
        nodes_generated += len(ids)
        keep = [False] * len(ids)
        if batch_dedup == "sequential":
            for i in range(len(ids)):
                key = flat[i].tobytes()
                e = closed.get(key)
                if e is None or e[0] > dep[i]:
                    closed[key] = [dep[i], ids[i]]; keep[i] = True
        else:
            best: Dict[bytes, int] = {}
            for i in range(len(ids)):
                key = flat[i].tobytes()
                j = best.get(key)
                if j is None or (dep[i], ids[i]) < (dep[j], ids[j]):
                    best[key] = i
            for key, i in best.items():
                e = closed.get(key)
                if e is None or e[0] > dep[i]:
                    closed[key] = [dep[i], ids[i]]; keep[i] = True
        kept = [i for i in range(len(ids)) if keep[i]]
        for i, nid in enumerate(ids):
            states[nid] = flat[i]; depth[nid] = dep[i]; solved[nid] = bool(sv[i]); parent[nid] = popped[i // A]
        if kept:
            h = np.maximum(np.asarray(heuristic(flat[kept]), dtype=np.float32), np.float32(0.0))
            for k, i in enumerate(kept):
                cost = ct(ct(weight) * ct(dep[i])) + ct(h[k]) * ct(0.0 if sv[i] else 1.0)
                heapq.heappush(open_heap, (cost, ids[i]))
        if keep_trace:
            trace.append({"popped": list(popped), "kept": [ids[i] for i in kept]})
    moves = None
    goal_id = None
    if goals:
        goal_id = goals[int(np.argmin([depth[g] for g in goals]))]
        moves = []
        nid = goal_id
        while nid != 0:
            moves.append(nid % A); nid = parent[nid]
        moves.reverse()
    return {"moves": moves, "nodes_generated": nodes_generated, "steps": steps, "goal_id": goal_id, "trace": trace,
            "popped_per_step": popped_per_step, "open_size": len(open_heap), "closed_size": len(closed)}


def misplaced_heuristic(env):
    """An exactly-representable heuristic for bit-exact engine tests: (#positions != goal) / 8 on nnet input."""
    goal_in = env.nnet_input(env.goal[None])[0]

    def h_states(states: np.ndarray) -> np.ndarray:
        x = env.nnet_input(states)
        return ((x != goal_in[None]).sum(axis=1).astype(np.float32) / np.float32(8.0)).astype(np.float32)
    return h_states
