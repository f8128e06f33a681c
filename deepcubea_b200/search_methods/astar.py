"""Batch weighted A* search: the reference's Search API and CLI (search_methods/astar.py) on the B200 engine.

API kept: `Node`, `AStar(states, env, heuristic_fn, weights)`, `.step(heuristic_fn, batch_size, include_solved,
verbose)`, `.has_found_goal()`, `.get_goal_nodes(i)`, `.get_goal_node_smallest_path_cost(i)`,
`.get_num_nodes_generated(i)`, `.get_popped_nodes()`, module-level `get_path(node)` (astar.py:18-44, 213-340).
Nodes are views over the device-resident search: OPEN, CLOSED and the node arena never leave HBM; `Node`
objects are only materialised for goal / popped nodes when asked for.

CLI kept (astar.py:343-397): --states --model_dir --env --batch_size --weight --language --results_dir
--start_idx --nnet_batch_size --verbose --debug; writes <results_dir>/output.txt and results.pkl with keys
states, solutions, paths, times, num_nodes_generated.  --language values:
    cuda (default), cpp : GPU engine with the C++ program's semantics (cpp/parallel_weighted_astar.cpp:138-346),
                          the variant that produced the reference's shipped results;
    python              : GPU engine with the Python AStar semantics (astar.py:232-340, 400-454).
Extra flags: --nnet_precision {fp32,tf32,bf16,fp16x3,fp16}, --max_nodes N, --num_states N.
"""
from __future__ import annotations

import os
import pickle
import sys
import time
from argparse import ArgumentParser
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from ..environments.environment_abstract import Environment, State
from .._lib import DcbError
from ..search.engine import NONE, BWASGpu, SearchEngine
from ..utils import data_utils, env_utils, misc_utils, nnet_utils, search_utils  # noqa: F401


class Node:
    __slots__ = ["state", "path_cost", "heuristic", "cost", "is_solved", "parent_move", "parent", "transition_costs",
                 "children", "bellman"]

    def __init__(self, state: State, path_cost: float, is_solved: bool, parent_move: Optional[int], parent):
        self.state: State = state
        self.path_cost: float = path_cost
        self.heuristic: Optional[float] = None
        self.cost: Optional[float] = None
        self.is_solved: bool = is_solved
        self.parent_move: Optional[int] = parent_move
        self.parent: Optional[Node] = parent
        self.transition_costs: List[float] = []
        self.children: List[Node] = []
        self.bellman: float = np.inf


def get_path(node: Node) -> Tuple[List[State], List[int], float]:
    """Walk parent links back to the root (astar.py:213-229)."""
    path: List[State] = []
    moves: List[int] = []
    cur = node
    while cur.parent is not None:
        path.append(cur.state)
        moves.append(cur.parent_move)
        cur = cur.parent
    path.append(cur.state)
    return path[::-1], moves[::-1], node.path_cost


def _device_heuristic(heuristic_fn: Callable, env: Environment) -> Callable[[torch.Tensor], torch.Tensor]:
    """Zero-copy form when the heuristic came from nnet_utils; otherwise adapt a reference-style callable
    (nnet-format numpy in, numpy out) -- functional for any user heuristic, at PCIe speed."""
    dev_fn = getattr(heuristic_fn, "device_fn", None)
    if dev_fn is not None:
        return dev_fn

    def adapted(x: torch.Tensor) -> torch.Tensor:
        out = heuristic_fn([x.cpu().numpy()], is_nnet_format=True)
        return torch.as_tensor(np.asarray(out, dtype=np.float32), device=x.device)
    return adapted


def _env_name(env: Environment) -> str:
    kind = type(env).__name__
    if kind in ("Cube3", "Cube4"):
        return kind.lower()
    return "lightsout%d" % env.dim if kind == "LightsOut" else "puzzle%d" % (env.dim * env.dim - 1)


class AStar:
    """astar.py:232-340.  ALL instances live in one device-resident engine (search/engine.py): one node arena, one CLOSED table
    keyed per instance, segmented OPEN; a `step` pops every unsolved instance, expands the popped nodes of all of them in one
    launch, deduplicates, evaluates the survivors of all instances in ONE heuristic call and pushes -- the flattened batch of
    astar.py:107-113 / :186, without leaving HBM.  One host round trip per step (the instance records).
    Every instance keeps its own weight.  (The reference zips `self.weights` with the FILTERED instance list, astar.py:277: with unequal
    weights an instance takes over a neighbour's weight as soon as an earlier instance has found its goal.  Equal weights -- every caller in
    the reference -- are unaffected.)"""

    def __init__(self, states: List[State], env: Environment, heuristic_fn: Callable, weights: List[float],
                 max_nodes: Optional[int] = None):
        self.env: Environment = env
        self.weights: List[float] = list(weights)
        self.states: List[State] = list(states)
        self.heuristic_fn = heuristic_fn
        self.step_num: int = 0
        self.timings: Dict[str, float] = {"pop": 0.0, "expand": 0.0, "check": 0.0, "heur": 0.0, "add": 0.0, "itr": 0.0}
        # total node budget of the shared arena, split evenly over the instances (default 2^26 nodes)
        self.max_nodes = int(max_nodes or os.environ.get("DCB_MAX_NODES", 1 << 26))
        self.engine: Optional[SearchEngine] = None
        self.popped_ids: List[List[int]] = [[] for _ in states]
        self.heuristic_calls = 0            # engine-level evaluations: one per step whatever the number of instances

    def _engine(self, batch_size: int) -> SearchEngine:
        if self.engine is None:
            n = len(self.states)
            per_inst = max(self.max_nodes // max(n, 1), 1)
            self.engine = SearchEngine(_env_name(self.env), _device_heuristic(self.heuristic_fn, self.env), self.weights, batch_size,
                                       n_inst=n, max_nodes=per_inst * n, semantics="python")
            self.engine.raise_on_error = False          # reported per instance below
            self.engine.reset(self.env.pack(self.states))
            self.heuristic_calls += 1
        assert self.engine.B == batch_size, "batch_size must not change during a search"
        return self.engine

    def step(self, heuristic_fn: Callable, batch_size: int, include_solved: bool = False, verbose: bool = False):
        t_itr = time.time()
        eng = self._engine(batch_size)
        eng.profile = bool(verbose)
        eng.step_all(include_solved=include_solved)
        self.heuristic_calls += 1
        stuck = [i for i, rec in enumerate(eng.inst) if rec.done in (2, 3, 4) and rec.n_goals == 0]
        if stuck:
            raise DcbError("instance(s) %s stopped without a goal (2 OPEN exhausted, 3 node arena full, 4 OPEN full): %s -- raise max_nodes "
                           "(now %d nodes for %d instances)" % (stuck[:8], [int(eng.inst[i].done) for i in stuck[:8]], eng.max_nodes, eng.n_inst))
        popped = eng.popped_ids.cpu().numpy().view(np.uint32)
        for i, rec in enumerate(eng.inst):
            if rec.resting:
                continue
            self.popped_ids[i].extend(popped[i * eng.Bpad: i * eng.Bpad + rec.n_popped].tolist())
        for k in ("pop", "expand", "check", "heur", "add"):
            self.timings[k] = eng.timings[k]
        itr = time.time() - t_itr
        self.timings["itr"] += itr
        if verbose:
            print("Itr: %i, Times - pop: %.2f, expand: %.2f, check: %.2f, heur: %.2f, add: %.2f, itr: %.2f\n" % (
                self.step_num, self.timings["pop"], self.timings["expand"], self.timings["check"], self.timings["heur"],
                self.timings["add"], itr))
        self.step_num += 1

    # ---- Node materialisation (only for the nodes the caller asks about) ---------------------------------
    def node_chain(self, node_id: int) -> Node:
        eng = self.engine
        A, npi = eng.A, eng.nodes_per_inst
        ids = [int(node_id)]
        while ids[-1] % npi != 0:
            ids.append(int(eng.slot_parent[ids[-1] // A].item()) & 0xFFFFFFFF)
        ids.reverse()
        states = self.env.unpack(eng.node_states(ids))
        sel = torch.tensor(ids, dtype=torch.int64, device=eng.dev)
        g = eng.node_g[sel].cpu().numpy()
        sv = eng.node_solved[sel].cpu().numpy()
        parent: Optional[Node] = None
        for k, nid in enumerate(ids):
            parent = Node(states[k], float(g[k]), bool(sv[k]), None if k == 0 else nid % A, parent)
        return parent

    def has_found_goal(self) -> List[bool]:
        if self.engine is None:
            return [False] * len(self.states)
        return [rec.n_goals > 0 for rec in self.engine.inst]

    def _goal_ids(self, inst_idx: int) -> List[int]:
        ids = self.popped_ids[inst_idx]
        if not ids or self.engine is None:
            return []
        sv = self.engine.node_solved[torch.tensor(ids, dtype=torch.int64, device=self.engine.dev)].cpu().numpy()
        return [i for i, s in zip(ids, sv) if s]

    def get_goal_nodes(self, inst_idx) -> List[Node]:
        return [self.node_chain(g) for g in self._goal_ids(inst_idx)]

    def get_goal_node_smallest_path_cost(self, inst_idx) -> Node:
        return self.node_chain(self.engine.inst[inst_idx].goal_id)

    def get_num_nodes_generated(self, inst_idx: int) -> int:
        return int(self.engine.inst[inst_idx].nodes_generated) if self.engine else 0

    def get_popped_nodes(self) -> List[List[Node]]:
        return [[self.node_chain(i) for i in ids] for ids in self.popped_ids]


# =====================================================================================================
# CLI
# =====================================================================================================
def main(argv: Optional[List[str]] = None):
    parser = ArgumentParser()
    parser.add_argument("--states", type=str, required=True, help="File containing states to solve")
    parser.add_argument("--model_dir", type=str, required=True, help="Directory of nnet model")
    parser.add_argument("--env", type=str, required=True, help="Environment: cube3, cube4, puzzle15, puzzle24, puzzle35, puzzle48, lightsout7")
    parser.add_argument("--batch_size", type=int, default=1, help="Batch size for BWAS")
    parser.add_argument("--weight", type=float, default=1.0, help="Weight of path cost")
    parser.add_argument("--language", type=str, default="cuda", help="cuda (=cpp semantics on the GPU), cpp, or python")
    parser.add_argument("--results_dir", type=str, required=True, help="Directory to save results")
    parser.add_argument("--start_idx", type=int, default=0, help="")
    parser.add_argument("--nnet_batch_size", type=int, default=None,
                        help="States evaluated by the neural network at a time; does not affect results")
    parser.add_argument("--verbose", action="store_true", default=False, help="Set for verbose")
    parser.add_argument("--debug", action="store_true", default=False, help="Set when debugging")
    parser.add_argument("--nnet_precision", type=str, default=None, help="fp16x3 (default: hand-written tcgen05 layers, fp32-parity) | fp32 | tf32 | bf16 (cuBLAS) | fp16 (tcgen05, reduced)")
    parser.add_argument("--max_nodes", type=int, default=1 << 26, help="Node arena capacity per search")
    parser.add_argument("--num_states", type=int, default=None, help="Solve only the first N states (after --start_idx)")
    args = parser.parse_args(argv)

    if not os.path.exists(args.results_dir):
        os.makedirs(args.results_dir)
    results_file = "%s/results.pkl" % args.results_dir
    output_file = "%s/output.txt" % args.results_dir
    stdout_prev = sys.stdout
    rank = int(os.environ.get("RANK", 0))
    if not args.debug:
        # under torchrun only rank 0 owns output.txt (it prints every state's line after the gather); the others keep a private log
        sys.stdout = data_utils.Logger(output_file if rank == 0 else "%s/output.rank%d.txt" % (args.results_dir, rank), "w")
    try:
        input_data = pickle.load(open(args.states, "rb"))
        states: List[State] = input_data["states"][args.start_idx:]
        if args.num_states is not None:
            states = states[:args.num_states]
        env: Environment = env_utils.get_environment(args.env)
        results: Dict[str, Any] = {"states": states}
        lang = args.language.lower()
        if lang == "python":
            solns, paths, times, num_nodes_gen = bwas_python(args, env, states)
        elif lang in ("cuda", "cpp"):
            solns, paths, times, num_nodes_gen = bwas_cuda(args, env, states)
        else:
            raise ValueError("Unknown language %s" % args.language)
        results["solutions"] = solns
        results["paths"] = paths
        results["times"] = times
        results["num_nodes_generated"] = num_nodes_gen
        tc = getattr(getattr(args, "_heuristic_fn", None), "device_fn", None)
        if hasattr(tc, "gemm_launches"):
            print("nnet: precision=%s kernel=dcb_resnet_gemm (tcgen05) launches=%d" % (tc.mode, tc.gemm_launches))
        else:
            print("nnet: precision=%s (cuBLAS via PyTorch)" % (args.nnet_precision or os.environ.get("DCB_NNET_PRECISION", "fp32")))
        if rank == 0:
            pickle.dump(results, open(results_file, "wb"), protocol=-1)
    finally:
        sys.stdout = stdout_prev


def _load_heuristic(args, env: Environment):
    device, devices, on_gpu = nnet_utils.get_device()
    print("device: %s, devices: %s, on_gpu: %s" % (device, devices, on_gpu))
    if not on_gpu:
        raise RuntimeError("the BWAS engine is CUDA-only; no GPU is visible and there is no CPU fallback")
    fn = nnet_utils.load_heuristic_fn(args.model_dir, device, on_gpu, env.get_nnet_model(), env, clip_zero=True,
                                      batch_size=args.nnet_batch_size, precision=args.nnet_precision)
    args._heuristic_fn = fn
    return fn


def bwas_python(args, env: Environment, states: List[State]):
    """astar.py:400-454 on the GPU engine (Python semantics)."""
    heuristic_fn = _load_heuristic(args, env)
    solns, paths, times, num_nodes_gen = [], [], [], []
    for state_idx, state in enumerate(states):
        start_time = time.time()
        num_itrs = 0
        astar = AStar([state], env, heuristic_fn, [args.weight], max_nodes=args.max_nodes)
        while not min(astar.has_found_goal()):
            astar.step(heuristic_fn, args.batch_size, verbose=args.verbose)
            num_itrs += 1
        goal_node = astar.get_goal_node_smallest_path_cost(0)
        path, soln, path_cost = get_path(goal_node)
        n_gen = astar.get_num_nodes_generated(0)
        solve_time = time.time() - start_time
        solns.append(soln); paths.append(path); times.append(solve_time); num_nodes_gen.append(n_gen)
        assert search_utils.is_valid_soln(state, soln, env)
        timing_str = ", ".join(["%s: %.2f" % (k, v) for k, v in astar.timings.items()])
        print("Times - %s, num_itrs: %i" % (timing_str, num_itrs))
        print("State: %i, SolnCost: %.2f, # Moves: %i, # Nodes Gen: %s, Time: %.2f" % (
            state_idx, path_cost, len(soln), format(n_gen, ","), solve_time))
        astar.engine = None                     # release HBM before the next state
    return solns, paths, times, num_nodes_gen


def bwas_cuda(args, env: Environment, states: List[State]):
    """The `--language cpp` flow of astar.py:457-568 with the child process, socket and stdout protocol
    replaced by in-process calls into the C ABI: one engine per GPU, states solved one after the other.
    Under torchrun (one process per GPU) every rank draws whole start states from a dynamic queue (a fetch-add on the job's c10d
    store: searches differ in length by orders of magnitude) and rank 0 merges the results; nothing else crosses GPUs."""
    import torch.distributed as dist

    from ..search import sharding
    rank, world_size, local = sharding.world()
    if world_size > 1:
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    heuristic_fn = _load_heuristic(args, env)
    engine = BWASGpu(args.env, heuristic_fn.device_fn, args.weight, args.batch_size, max_nodes=args.max_nodes, semantics="cpp")
    packed = env.pack(states)
    if world_size > 1:
        local_res = []
        for state_idx in sharding.InstanceQueue(len(states)):
            res = engine.solve(packed[state_idx])
            if res.moves is None:
                raise RuntimeError("OPEN exhausted without reaching the goal for state %d" % state_idx)
            local_res.append((state_idx, ([int(m) for m in res.moves], res.solve_time, res.nodes_generated)))
        sharding.completion_barrier()
        merged = sharding.gather_results(local_res, len(states))
        if rank != 0:
            return [], [], [], []
        solved = {i: r for i, r in enumerate(merged)}
    else:
        solved = None
    solns, paths, times, num_nodes_gen = [], [], [], []
    for state_idx, state in enumerate(states):
        if solved is not None:
            from types import SimpleNamespace
            mv, tm, ng = solved[state_idx]
            res = SimpleNamespace(moves=mv, solve_time=tm, nodes_generated=ng, timings={}, iterations=0)
            _finish_state(args, env, state_idx, state, res, solns, paths, times, num_nodes_gen)
            continue
        res = engine.solve(packed[state_idx])
        if res.moves is None:
            raise RuntimeError("OPEN exhausted without reaching the goal for state %d" % state_idx)
        _finish_state(args, env, state_idx, state, res, solns, paths, times, num_nodes_gen)
    return solns, paths, times, num_nodes_gen


def _finish_state(args, env, state_idx, state, res, solns, paths, times, num_nodes_gen):
    """Replay the moves into a path, validate, record and print (astar.py:539-562)."""
    soln = [int(m) for m in res.moves]
    path: List[State] = [state]
    cur = state
    tcs: List[float] = []
    for move in soln:
        nxt, tc = env.next_state([cur], move)
        cur = nxt[0]
        path.append(cur); tcs.append(tc[0])
    solns.append(soln); paths.append(path); times.append(res.solve_time); num_nodes_gen.append(res.nodes_generated)
    assert search_utils.is_valid_soln(state, soln, env)
    if args.verbose and res.timings:
        print("Times - %s, num_itrs: %i" % (", ".join("%s: %.2f" % kv for kv in res.timings.items()), res.iterations))
    print("State: %i, SolnCost: %.2f, # Moves: %i, # Nodes Gen: %s, Time: %.2f" % (
        state_idx, sum(tcs), len(soln), format(res.nodes_generated, ","), res.solve_time))


if __name__ == "__main__":
    main()
