"""Batch weighted A* (BWAS) for ONE start state with every data structure resident in HBM.

Algorithm = cpp/parallel_weighted_astar.cpp:138-346 (the reference's `--language cpp` path, which produced
its shipped results); see oracle/oracle_bwas.py for the line-by-line restatement this engine is tested
against.  What runs where:

  OPEN   pop B cheapest / push      dcb_open_pop / dcb_open_push      (radix-select bucket queue, open_set.cu)
  expand + is_solved + hash         dcb_expand_indexed                 (expand_kernels.cu, TMA stores into the arena)
  depth / parent bookkeeping        dcb_child_meta                     (node_ops.cu)
  CLOSED insert-or-improve          dcb_closed_insert                  (closed_table.cu)
  cost-to-go                        `heuristic` on the same stream     (PyTorch / tcgen05 path, nnet/)
  cost, push                        dcb_compute_cost, dcb_open_push

Differences from the reference that do not change results: duplicates are removed BEFORE the heuristic is
evaluated (the reference sends every child to the NN, parallel_weighted_astar.cpp:237, then discards the
values of dropped nodes, :285-287); children never leave the device (no socket, :121-136, 275-279); the
final iteration's children are counted in `nodes_generated` (:266) but not materialised.

Node ids: id = slot * A + move, state at arena + id * S.  Slot 0 is the root's record.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .. import _lib
from .._lib import ENV_IDS, OpenState, check, ptr

NONE = 0xFFFFFFFF


@dataclass
class BWASResult:
    moves: Optional[List[int]]
    nodes_generated: int
    iterations: int
    solve_time: float
    path_cost: float
    done: int
    open_size: int = 0
    closed_size: int = 0
    timings: Dict[str, float] = field(default_factory=dict)
    trace: Optional[List[Dict]] = None


def _next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p *= 2
    return p


class BWASGpu:
    """Reusable engine: buffers are allocated once for `max_nodes` nodes and recycled per start state."""

    def __init__(self, env_name: str, heuristic: Callable[[torch.Tensor], torch.Tensor], weight: float,
                 batch_size: int, max_nodes: int = 1 << 24, device: Optional[torch.device] = None,
                 semantics: str = "cpp"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DcbError("BWASGpu needs a CUDA device; there is no CPU fallback")
        self.env_name = env_name.lower()
        if self.env_name not in ENV_IDS:
            raise ValueError("No known environment %s" % env_name)
        self.env = ENV_IDS[self.env_name]
        self.S = self.lib.dcb_env_state_bytes(self.env)
        self.A = self.lib.dcb_env_num_moves(self.env)
        self.align = self.lib.dcb_env_slot_align(self.env)
        if semantics not in ("cpp", "python"):
            raise ValueError("semantics must be 'cpp' or 'python'")
        self.semantics = semantics
        self.heuristic = heuristic
        self.weight = float(weight)
        self.B = int(batch_size)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        S, A, B = self.S, self.A, self.B
        # node capacity in whole slots, leaving room for the alignment pads
        self.max_slots = max(int(max_nodes) // A, 4 * self.align + B + 8)
        self.max_nodes = self.max_slots * A
        if self.max_nodes >= (1 << 31):
            raise ValueError("max_nodes must stay below 2^31")
        dev = self.dev
        u8, u32, i64, f32 = torch.uint8, torch.int32, torch.int64, torch.float32   # int32 storage, u32 bits
        self.arena = torch.empty(self.max_nodes * S + 64, dtype=u8, device=dev)
        self.node_g = torch.empty(self.max_nodes, dtype=u32, device=dev)
        self.node_solved = torch.zeros(self.max_nodes, dtype=u8, device=dev)
        self.slot_parent = torch.empty(self.max_slots + 1, dtype=u32, device=dev)
        self.closed_cap = _next_pow2(2 * self.max_nodes)
        self.closed = torch.empty(self.closed_cap * 2, dtype=i64, device=dev)
        self.open_cap = self.max_nodes
        self.open_key = torch.empty(self.open_cap, dtype=u32, device=dev)
        self.open_id = torch.empty(self.open_cap, dtype=u32, device=dev)
        self.open_state = torch.zeros(16, dtype=u32, device=dev)
        self.open_scratch = torch.empty(int(self.lib.dcb_open_scratch_bytes(self.open_cap, B)) + 16, dtype=u8, device=dev)
        m = B * A
        self.popped_ids = torch.empty(B, dtype=u32, device=dev)
        self.hash_tmp = torch.empty(max(m, 2), dtype=i64, device=dev)
        self.slot_tmp = torch.empty(max(m, 1), dtype=u32, device=dev)
        self.keep_tmp = torch.empty(max(m, 1), dtype=u8, device=dev)
        self.kept_ids = torch.empty(max(m, 1), dtype=u32, device=dev)
        self.nn_in = torch.empty((max(m, 1), S), dtype=u8, device=dev)
        self.cost_tmp = torch.empty(max(m, 1), dtype=f32, device=dev)
        self.counters = torch.zeros(4, dtype=u32, device=dev)        # [0] n_kept, [1] closed entries
        self.path_moves = torch.empty(4096, dtype=u8, device=dev)
        self.path_len = torch.zeros(1, dtype=u32, device=dev)
        self.h_state = torch.zeros(16, dtype=u32).pin_memory()
        self.h_counters = torch.zeros(4, dtype=u32).pin_memory()
        self.kernel_launches = 0            # hand-written kernels launched (for bench.py's gpu_launches)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.expand_events = None           # list of (start, end, n_parents) CUDA events when profiling is on
        self.total_kept = 0                 # children that survived CLOSED (= rows sent to the heuristic), all searches

    # ------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _read_state(self) -> OpenState:
        self.h_state.copy_(self.open_state, non_blocking=True)
        self.d2h_bytes += 64
        torch.cuda.current_stream(self.dev).synchronize()
        return OpenState.from_buffer_copy(self.h_state.numpy().tobytes())

    # ------------------------------------------------------------------------------------------------
    def reset(self, start: np.ndarray) -> None:
        """Root node: OPEN and CLOSED as the chosen semantics prescribe."""
        lib, st = self.lib, self._stream()
        S, env = self.S, self.env
        start_t = torch.from_numpy(np.ascontiguousarray(start, dtype=np.uint8).reshape(1, S))
        with torch.cuda.device(self.dev):
            self.arena[:S].copy_(start_t.reshape(-1), non_blocking=True)
            self.h2d_bytes += S
            self.kernel_launches += 6
            check(lib.dcb_closed_clear(ptr(self.closed), self.closed_cap, st), "closed_clear")
            check(lib.dcb_open_clear(ptr(self.open_state), st), "open_clear")
            self.counters.zero_()
            self.node_g[:1].zero_()
            check(lib.dcb_is_solved(env, ptr(self.arena), 1, ptr(self.node_solved), st), "is_solved(root)")
            if self.semantics == "cpp":
                # root in CLOSED, cost 0 / heuristic 0 (parallel_weighted_astar.cpp:160-162)
                check(lib.dcb_hash_states(env, ptr(self.arena), 1, ptr(self.hash_tmp), st), "hash(root)")
                check(lib.dcb_closed_insert(env, ptr(self.closed), self.closed_cap, ptr(self.arena), ptr(self.hash_tmp),
                                            ptr(self.node_g), None, 0, 1, ptr(self.slot_tmp), ptr(self.keep_tmp),
                                            self.counters[1:].data_ptr(), st), "closed_insert(root)")
                self.cost_tmp[:1].zero_()
            else:
                # Python path: root NOT in CLOSED, root cost = w*0 + h(root) (search_methods/astar.py:244-249, 50-62)
                self.kept_ids[:1].zero_()
                check(lib.dcb_gather_nnet_input(env, ptr(self.arena), ptr(self.kept_ids), 1, ptr(self.nn_in), st), "gather(root)")
                h = self.heuristic(self.nn_in[:1]).float().contiguous()
                check(lib.dcb_compute_cost(ptr(h), ptr(self.kept_ids), ptr(self.node_g), ptr(self.node_solved), self.weight, 1,
                                           ptr(self.cost_tmp), st), "cost(root)")
            check(lib.dcb_open_push(ptr(self.open_state), ptr(self.open_key), ptr(self.open_id), self.open_cap,
                                    ptr(self.cost_tmp), None, 0, None, 1, st), "open_push(root)")
        self.next_slot = 1
        self.nodes_generated = 1 if self.semantics == "cpp" else 0       # :166 vs astar.py:168
        self.nodes_expanded = 0          # children actually materialised (the terminating iteration's are only counted above)
        self.iterations = 0
        self.done = 0
        self.goal_id = NONE
        self.goal_ids: List[int] = []          # python semantics: every solved node popped so far
        self.timings = {"pop": 0.0, "expand": 0.0, "check": 0.0, "heur": 0.0, "add": 0.0}
        self.last_popped = 0
        self.last_kept = 0

    def step(self, keep_trace: bool = False) -> Optional[Dict]:
        """One BWAS iteration.  Returns the trace record when asked."""
        lib, st = self.lib, self._stream()
        S, A, B, env = self.S, self.A, self.B, self.env
        tm = self.timings
        rec: Optional[Dict] = None
        with torch.cuda.device(self.dev):
            t0 = time.perf_counter()
            cpp = self.semantics == "cpp"
            check(lib.dcb_open_pop(ptr(self.open_state), ptr(self.open_key), ptr(self.open_id), self.open_cap, B, 1 if cpp else 0,
                                   ptr(self.node_solved), ptr(self.popped_ids), ptr(self.open_scratch), st), "open_pop")
            self.kernel_launches += 14
            os_ = self._read_state()
            if os_.overflow:
                raise _lib.DcbError("OPEN overflow: raise max_nodes (capacity %d)" % self.open_cap)
            n_pop = int(os_.n_popped)
            self.last_popped = n_pop
            self.iterations += 1
            self.nodes_generated += n_pop * A                       # :266, counted on the final iteration too
            if cpp:
                self.done, self.goal_id = int(os_.done), int(os_.goal_id)
            else:
                if n_pop == 0:
                    self.done = 2
                else:
                    pid = self.popped_ids[:n_pop].long()
                    sv = self.node_solved[pid]
                    if bool(sv.any()):                               # goal recorded at pop (astar.py:73)
                        gids = pid[sv.bool()]
                        self.goal_ids.extend(gids.cpu().tolist())
                        g = self.node_g[torch.tensor(self.goal_ids, device=self.dev)].cpu().numpy().view(np.uint32)
                        self.goal_id = self.goal_ids[int(np.argmin(g))]     # smallest path cost (astar.py:327-333)
                        self.done = 1
            t1 = time.perf_counter(); tm["pop"] += t1 - t0
            if (cpp and self.done) or n_pop == 0:
                self.last_kept = 0
                if keep_trace:
                    rec = {"popped": self.popped_ids[:n_pop].cpu().numpy().view(np.uint32).tolist(), "kept": []}
                return rec
            base_slot = -(-self.next_slot // self.align) * self.align
            if base_slot + n_pop + 1 > self.max_slots:
                raise _lib.DcbError("node arena full (%d nodes): raise max_nodes" % self.max_nodes)
            first_id = base_slot * A
            m = n_pop * A
            self.nodes_expanded += m
            # ---- expand (:217-230): children land directly in the arena ----
            if self.expand_events is not None:
                ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
            check(lib.dcb_expand_indexed(env, ptr(self.arena), ptr(self.popped_ids), n_pop,
                                         self.arena.data_ptr() + first_id * S, self.node_solved.data_ptr() + first_id,
                                         ptr(self.hash_tmp), st), "expand_indexed")
            if self.expand_events is not None:
                ev1 = torch.cuda.Event(enable_timing=True); ev1.record()
                self.expand_events.append((ev0, ev1, n_pop))
            self.kernel_launches += 8      # expand, child_meta, insert, resolve, compact, gather, cost, push
            check(lib.dcb_child_meta(env, ptr(self.popped_ids), n_pop, first_id, ptr(self.node_g), ptr(self.slot_parent), st),
                  "child_meta")
            t2 = time.perf_counter(); tm["expand"] += t2 - t1
            # ---- CLOSED (:243-265) ----
            check(lib.dcb_closed_insert(env, ptr(self.closed), self.closed_cap, ptr(self.arena), ptr(self.hash_tmp),
                                        self.node_g.data_ptr() + 4 * first_id, None, first_id, m, ptr(self.slot_tmp),
                                        ptr(self.keep_tmp), self.counters[1:].data_ptr(), st), "closed_insert")
            self.counters[:1].zero_()
            check(lib.dcb_compact_kept(ptr(self.keep_tmp), first_id, m, ptr(self.kept_ids), ptr(self.counters), st), "compact_kept")
            self.h_counters.copy_(self.counters, non_blocking=True)
            self.d2h_bytes += 16
            torch.cuda.current_stream(self.dev).synchronize()
            n_kept = int(self.h_counters[0]) & 0xFFFFFFFF
            self.last_kept = n_kept
            self.total_kept += n_kept
            t3 = time.perf_counter(); tm["check"] += t3 - t2
            # ---- heuristic on the kept children only, same stream ----
            if n_kept:
                if hasattr(self.heuristic, "eval_nodes"):       # tcgen05 path: one-hot input built straight from the arena
                    h = self.heuristic.eval_nodes(env, self.arena, self.kept_ids, n_kept)
                else:
                    check(lib.dcb_gather_nnet_input(env, ptr(self.arena), ptr(self.kept_ids), n_kept, ptr(self.nn_in), st), "gather_nnet_input")
                    h = self.heuristic(self.nn_in[:n_kept])
                if h.dtype != torch.float32 or not h.is_contiguous():
                    h = h.float().contiguous()
                check(lib.dcb_compute_cost(ptr(h), ptr(self.kept_ids), ptr(self.node_g), ptr(self.node_solved), self.weight,
                                           n_kept, ptr(self.cost_tmp), st), "compute_cost")
                t4 = time.perf_counter(); tm["heur"] += t4 - t3
                check(lib.dcb_open_push(ptr(self.open_state), ptr(self.open_key), ptr(self.open_id), self.open_cap,
                                        ptr(self.cost_tmp), ptr(self.kept_ids), 0, None, n_kept, st), "open_push")
                tm["add"] += time.perf_counter() - t4
            self.next_slot = base_slot + n_pop
            if keep_trace:
                rec = {"popped": self.popped_ids[:n_pop].cpu().numpy().view(np.uint32).tolist(),
                       "kept": sorted(self.kept_ids[:n_kept].cpu().numpy().view(np.uint32).tolist())}
        return rec

    def path_to(self, node_id: int) -> List[int]:
        """Moves root -> node (parallel_weighted_astar.cpp:336-341 / astar.py:213-229)."""
        lib, st = self.lib, self._stream()
        with torch.cuda.device(self.dev):
            check(lib.dcb_reconstruct_path(self.env, ptr(self.slot_parent), int(node_id), self.path_moves.numel(), ptr(self.path_moves),
                                           ptr(self.path_len), st), "reconstruct_path")
            n = int(self.path_len.cpu().numpy().view(np.int32)[0])
            if n < 0:
                raise _lib.DcbError("solution longer than %d moves" % self.path_moves.numel())
            return self.path_moves[:n].cpu().numpy().tolist()

    def node_states(self, ids: List[int]) -> np.ndarray:
        """States of the given node ids, u8 [len(ids), S] on the host."""
        idx = torch.tensor(ids, dtype=torch.int64, device=self.dev)
        offs = idx[:, None] * self.S + torch.arange(self.S, device=self.dev)[None, :]
        return self.arena[offs].cpu().numpy()

    def solve(self, start: np.ndarray, max_iters: Optional[int] = None, keep_trace: bool = False) -> BWASResult:
        t_begin = time.perf_counter()
        self.reset(start)
        trace: List[Dict] = []
        while not self.done:
            if max_iters is not None and self.iterations >= max_iters:
                break
            rec = self.step(keep_trace)
            if keep_trace:
                trace.append(rec)
        moves = self.path_to(self.goal_id) if self.goal_id != NONE else None
        final = self._read_state()
        self.h_counters.copy_(self.counters)
        return BWASResult(moves=moves, nodes_generated=self.nodes_generated, iterations=self.iterations,
                          solve_time=time.perf_counter() - t_begin, path_cost=float(len(moves)) if moves is not None else float("nan"),
                          done=self.done, open_size=int(final.size), closed_size=int(self.h_counters[1]) & 0xFFFFFFFF,
                          timings=dict(self.timings), trace=trace if keep_trace else None)
