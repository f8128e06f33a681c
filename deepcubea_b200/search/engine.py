"""Device-driven batch weighted A*: n problem instances advanced together, every data structure resident in HBM.

One iteration (cpp/parallel_weighted_astar.cpp:169-330, or AStar.step over all instances, search_methods/astar.py:256-317) is a
fixed sequence of launches behind the C ABI (include/dcb.h, dcb_search_*):

    dcb_search_pop      segmented exact top-B pop per instance, goal / termination rule, slot assignment, tile list   (open_set.cu,
                        search_step.cu)
    dcb_search_expand   children + is_solved + hash + depth / parent link of every popped node                        (expand_kernels.cu)
    dcb_search_closed   CLOSED insert-or-improve in the reference's child order, survivors compacted                  (closed_table.cu)
    cost-to-go network  on the survivors only, same stream                                                            (nnet/)
    dcb_search_push     cost = max(h,0)*(!solved) + weight*g, push to the owning instance's OPEN                      (search_step.cu)

Every size in between (parents popped, tiles, children kept) lives in device memory, so the host enqueues an iteration
without reading anything back; `solve()` keeps one iteration in flight ahead of the one it inspects.

Semantics "cpp"   : the C++ program that produced the reference's shipped results -- root cost 0 and in CLOSED, pop truncated at
                    the first solved node, termination one iteration later (:205-208), nodes generated = 1 + sum of children.
Semantics "python": the AStar class -- root evaluated by the network and not in CLOSED, every pop stands, a search ends when a
                    popped node is solved, answer = goal node of smallest path cost (astar.py:232-340).

Differences from the reference that do not change results: duplicates are removed BEFORE the heuristic is evaluated (the
reference sends every child to the network, parallel_weighted_astar.cpp:237, then discards the values of dropped nodes,
:285-287); children never leave the device (no socket, :121-136, 275-279); the terminating iteration's children are counted in
`nodes_generated` (:266) but never materialised; when the stored node of a state is improved (:255-257) the reference rewrites
its depth / parent in place, here the older node keeps its own (measured: identical solutions, tests/test_oracle_bwas.py).
Costs follow each semantics' own arithmetic: float32 without FMA contraction for the C++ program (:298), float64 for the Python AStar
(astar.py:196; 64-bit keys).  Heap ties (equal cost) break towards the smaller node id = the order the nodes were pushed (heapq's FIFO, astar.py:66;
the C++ heap's tie order is unspecified).

Node ids: id = slot * A + move, state at arena + id * S.  Instance i owns slots [i * slots_per_inst, (i+1) * slots_per_inst);
its root is node i * slots_per_inst * A.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import _lib
from .._lib import ENV_IDS, INST_WORDS, PLAN_WORDS, SearchCtx, SearchInst, StepPlan, check, ptr

NONE = 0xFFFFFFFF
ERR_BITS = {1: "node arena full: raise max_nodes", 2: "OPEN segment full: raise max_nodes", 4: "CLOSED table full"}


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


class CapacityHeuristic:
    """Adapter that lets any device heuristic (nnet-input u8 [m, S] -> f32 [m]) run without the host knowing the row count: it is
    evaluated on the full candidate capacity (rows past the device-side count hold stale but valid node ids and are ignored by
    the push).  Costs capacity/kept times the arithmetic -- meant for cheap heuristics and for tests of the sync-free loop."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor]):
        self.fn = fn
        self._nn_in = None

    def eval_nodes_dev(self, env_id, arena, ids, n_dev, cap):
        lib = _lib.load()
        if self._nn_in is None or self._nn_in.shape[0] < cap:
            self._nn_in = torch.empty((cap, lib.dcb_env_state_bytes(env_id)), dtype=torch.uint8, device=arena.device)
        st = torch.cuda.current_stream(arena.device).cuda_stream
        check(lib.dcb_gather_nnet_input(env_id, ptr(arena), ptr(ids), cap, ptr(self._nn_in), st), "gather_nnet_input")
        h = self.fn(self._nn_in[:cap])
        return ("h", h if (h.dtype == torch.float32 and h.is_contiguous()) else h.float().contiguous())


class SearchEngine:
    """Reusable engine: buffers are allocated once for `max_nodes` nodes (all instances together) and recycled per reset."""

    def __init__(self, env_name: str, heuristic, weights: Union[float, Sequence[float]], batch_size: int, n_inst: int = 1,
                 max_nodes: int = 1 << 24, device: Optional[torch.device] = None, semantics: str = "cpp",
                 sync_free: Optional[bool] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DcbError("the search engine needs a CUDA device; there is no CPU fallback")
        self.env_name = env_name.lower()
        if self.env_name not in ENV_IDS:
            raise ValueError("No known environment %s" % env_name)
        if semantics not in ("cpp", "python"):
            raise ValueError("semantics must be 'cpp' or 'python'")
        lib = self.lib
        self.env = ENV_IDS[self.env_name]
        self.S = lib.dcb_env_state_bytes(self.env)
        self.A = lib.dcb_env_num_moves(self.env)
        self.align = lib.dcb_env_slot_align(self.env)
        self.semantics = semantics
        self.heuristic = heuristic
        self.n_inst = int(n_inst)
        self.B = int(batch_size)
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        S, A, B, I = self.S, self.A, self.B, self.n_inst
        self.Bpad = -(-B // 32) * 32
        ws = [float(weights)] * I if isinstance(weights, (int, float)) else [float(w) for w in weights]
        assert len(ws) == I
        self.weight = ws[0]
        # slots per instance: a multiple of 32 (tiles and the 16-byte alignment of child blocks), room for one batch at least
        spi = max(int(max_nodes) // (A * I), self.Bpad + 2 * max(self.align, 32) + 32)
        spi = -(-spi // 32) * 32
        if I * spi * A >= (1 << 32) - 64:
            raise ValueError("max_nodes too large: node ids are 32-bit")
        self.slots_per_inst = self.max_slots = spi
        self.nodes_per_inst = spi * A
        self.max_nodes = I * spi * A
        self.max_tiles = I * (self.Bpad // 32)
        self.max_cand = self.max_tiles * 32 * A
        dev = self.dev
        u8, i32, i64, f32 = torch.uint8, torch.int32, torch.int64, torch.float32   # int32 storage, u32 bits
        with torch.cuda.device(dev):
            self.arena = torch.empty(self.max_nodes * S + 64, dtype=u8, device=dev)
            self.node_g = torch.empty(self.max_nodes, dtype=i32, device=dev)
            self.node_solved = torch.zeros(self.max_nodes + 64, dtype=u8, device=dev)
            self.slot_parent = torch.empty(I * spi + 1, dtype=i32, device=dev)
            self.open_key = torch.empty(self.max_nodes, dtype=i32, device=dev)
            self.open_id = torch.empty(self.max_nodes, dtype=i32, device=dev)
            # Python semantics: float64 costs (astar.py:196) -> 64-bit keys, low words in a second array
            self.open_key_lo = torch.empty(self.max_nodes, dtype=i32, device=dev) if semantics == "python" else None
            # [plan | instance records]: one device->host copy reads everything the host ever needs
            self.state_buf = torch.zeros(PLAN_WORDS + INST_WORDS * I, dtype=i32, device=dev)
            self.weights_d = torch.tensor(ws, dtype=torch.float64, device=dev)
            self.popped_ids = torch.zeros(I * self.Bpad, dtype=i32, device=dev)
            self.tiles = torch.zeros(self.max_tiles * 4, dtype=i32, device=dev)
            self.hash_tmp = torch.empty(max(self.max_cand, 2), dtype=i64, device=dev)
            self.kept_ids = torch.zeros(self.max_cand, dtype=i32, device=dev)          # zero: stale entries are valid node ids
            self.pop_scratch = torch.empty(int(lib.dcb_search_pop_scratch_bytes(I, self.nodes_per_inst, B)) + 16, dtype=u8, device=dev)
            self.closed_scratch = torch.empty(int(lib.dcb_closed_scratch_bytes(self.max_cand)) + 16, dtype=u8, device=dev)
            self.nn_in = None
            self.roots_d = torch.empty((I, S), dtype=u8, device=dev)
            self.path_moves = torch.empty(4096, dtype=u8, device=dev)
            self.path_len = torch.zeros(1, dtype=i32, device=dev)
            self.h_bufs = [torch.zeros(PLAN_WORDS + INST_WORDS * I, dtype=i32).pin_memory() for _ in range(2)]
            self.h_events = [torch.cuda.Event() for _ in range(2)]
        # CLOSED grows with the search (dcb_closed_rehash): a small table stays in L2 and a reset only clears what a search of that
        # size needs; the largest table holds every node at load <= 0.5
        self.closed_cap_max = min(_next_pow2(2 * self.max_nodes), 1 << 31)
        self.closed_cap_min = min(self.closed_cap_max, _next_pow2(max(8 * self.max_cand, 1 << 16)))
        # storage for every table size is reserved up front (an allocation inside the search loop stalls the stream): sizes
        # alternate between two buffers so that a table and its successor never overlap during the rehash
        with torch.cuda.device(dev):
            self._closed_bufs = [torch.empty(self.closed_cap_max * 2, dtype=i64, device=dev),
                                 torch.empty(max(self.closed_cap_max, 2), dtype=i64, device=dev)]
        self.closed_cap = 0
        self.closed = None
        self._alloc_closed(self.closed_cap_min)
        self.ctx = SearchCtx()
        self._fill_ctx()
        self._budget_h = torch.zeros(1, dtype=i32).pin_memory()
        self.set_budget(None)
        if sync_free is None:
            sync_free = hasattr(heuristic, "eval_nodes_dev")
        if sync_free and not hasattr(heuristic, "eval_nodes_dev"):
            self.heuristic = CapacityHeuristic(heuristic)
        self.sync_free = bool(sync_free)
        self.kernel_launches = 0            # hand-written kernels launched (for bench.py's gpu_launches)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.expand_events = None           # list of (start, end, None) CUDA events around dcb_search_expand when profiling is on
        self.closed_growths = 0
        self.plan = StepPlan()
        self.inst: List[SearchInst] = [SearchInst() for _ in range(I)]
        self.timings = {"pop": 0.0, "expand": 0.0, "check": 0.0, "heur": 0.0, "add": 0.0}
        self.raise_on_error = True          # False: a full instance just stops (its record says done = 3 / 4), the others go on
        self.profile = False                # CUDA-event phase timings (CLI --verbose); adds 6 event records per iteration
        self._phase_events = None
        self.total_kept = 0
        self._kept_base = 0                 # plan.total_kept restarts at every reset; total_kept accumulates over searches

    # ------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _alloc_closed(self, cap: int) -> None:
        level = (self.closed_cap_max // cap).bit_length() - 1          # 0 = the largest table
        self.closed = self._closed_bufs[level & 1][:cap * 2]
        self.closed_cap = cap

    def _fill_ctx(self) -> None:
        c = self.ctx
        c.env, c.n_inst, c.batch, c.semantics = self.env, self.n_inst, self.B, 0 if self.semantics == "cpp" else 1
        c.slots_per_inst, c.open_per_inst, c.closed_capacity = self.slots_per_inst, self.nodes_per_inst, self.closed_cap
        c.d_arena, c.d_node_g, c.d_node_solved, c.d_slot_parent = ptr(self.arena), ptr(self.node_g), ptr(self.node_solved), ptr(self.slot_parent)
        c.d_closed, c.d_open_key, c.d_open_id = ptr(self.closed), ptr(self.open_key), ptr(self.open_id)
        c.d_open_key_lo = ptr(self.open_key_lo)
        c.d_plan = self.state_buf.data_ptr()
        c.d_inst = self.state_buf.data_ptr() + 4 * PLAN_WORDS
        c.d_weights, c.d_popped_ids, c.d_tiles = ptr(self.weights_d), ptr(self.popped_ids), ptr(self.tiles)
        c.d_hash, c.d_kept_ids = ptr(self.hash_tmp), ptr(self.kept_ids)
        align16 = lambda t: (t.data_ptr() + 15) // 16 * 16
        c.d_pop_scratch, c.d_closed_scratch = align16(self.pop_scratch), align16(self.closed_scratch)
        self._ctx_ref = ctypes.byref(c)
        self._n_kept_ptr = self.state_buf.data_ptr() + 4 * 2          # &plan.n_kept

    def set_budget(self, full_iterations: Optional[int]) -> None:
        """dcb_step_plan.budget: the pop stage rests once this many FULL-BATCH iterations have started (None = unlimited).  With it a
        host that enqueues one iteration ahead (pipelined_steps) still runs exactly k full iterations: the speculative one is a no-op."""
        self._budget_h[0] = -1 if full_iterations is None else int(full_iterations)           # -1 == 0xffffffff
        torch.cuda.current_stream(self.dev).synchronize()                                       # the pinned word may still be in flight
        self.state_buf[7:8].copy_(self._budget_h, non_blocking=True)

    def _grow_closed(self, need_entries: int) -> None:
        """Stream-ordered: clear the next larger table, re-insert every entry (dcb_closed_rehash), switch.  One doubling at a time:
        consecutive sizes live in different buffers."""
        lib, st = self.lib, self._stream()
        while self.closed_cap < self.closed_cap_max and 2 * need_entries > self.closed_cap:
            old, old_cap = self.closed, self.closed_cap
            self._alloc_closed(2 * old_cap)
            check(lib.dcb_closed_clear(ptr(self.closed), self.closed_cap, st), "closed_clear")
            check(lib.dcb_closed_rehash(ptr(old), old_cap, ptr(self.closed), self.closed_cap, st), "closed_rehash")
            self.kernel_launches += 2
            self.closed_growths += 1
        self._fill_ctx()

    # ---- device -> host ------------------------------------------------------------------------------
    def _readback(self, slot: int) -> None:
        self.h_bufs[slot].copy_(self.state_buf, non_blocking=True)
        self.h_events[slot].record(torch.cuda.current_stream(self.dev))
        self.d2h_bytes += self.state_buf.numel() * 4

    def _wait(self, slot: int) -> None:
        self.h_events[slot].synchronize()
        raw = self.h_bufs[slot].numpy().tobytes()
        self.plan = StepPlan.from_buffer_copy(raw[:4 * PLAN_WORDS])
        isz = 4 * INST_WORDS
        off = 4 * PLAN_WORDS
        self.inst = [SearchInst.from_buffer_copy(raw[off + i * isz: off + (i + 1) * isz]) for i in range(self.n_inst)]
        self.total_kept = self._kept_base + int(self.plan.total_kept)
        if self._phase_events is not None:
            ev = self._phase_events
            for k, name in enumerate(("pop", "expand", "check", "heur", "add")):
                self.timings[name] += ev[k].elapsed_time(ev[k + 1]) * 1e-3
            self._phase_events = None
        if self.plan.error and self.raise_on_error:
            msgs = [m for b, m in ERR_BITS.items() if self.plan.error & b]
            raise _lib.DcbError("search stopped: " + "; ".join(msgs) + " (max_nodes %d, %d instance(s))" % (self.max_nodes, self.n_inst))

    def sync_state(self) -> None:
        self._readback(0)
        self._wait(0)

    # ------------------------------------------------------------------------------------------------
    def reset(self, starts: np.ndarray) -> None:
        """Root nodes: OPEN and CLOSED as the chosen semantics prescribe.  `starts`: u8 [n_inst, S] (or [S] for one instance)."""
        lib, st = self.lib, self._stream()
        S, I = self.S, self.n_inst
        starts = np.ascontiguousarray(starts, dtype=np.uint8).reshape(I, S)
        with torch.cuda.device(self.dev):
            self.roots_d.copy_(torch.from_numpy(starts), non_blocking=False)
            self.h2d_bytes += I * S
            if self.closed_cap != self.closed_cap_min:
                self._alloc_closed(self.closed_cap_min)
                self._fill_ctx()
            check(lib.dcb_closed_clear(ptr(self.closed), self.closed_cap, st), "closed_clear")
            check(lib.dcb_search_reset(self._ctx_ref, ptr(self.roots_d), st), "search_reset")
            self.kernel_launches += 2
            self._kept_base = self.total_kept
            if self.semantics == "python":
                # root cost = w*0 + h(root): evaluate the roots and push them (astar.py:244-249)
                self._heuristic_and_push(n_host=I)
        self.plan = StepPlan()
        self.plan.n_running = I
        self.inst = [SearchInst() for _ in range(I)]
        for s in self.inst:
            s.goal_id = NONE
        self.timings = {"pop": 0.0, "expand": 0.0, "check": 0.0, "heur": 0.0, "add": 0.0}

    def _heuristic_and_push(self, n_host: Optional[int] = None) -> None:
        """Cost-to-go of kept_ids[0 .. plan.n_kept) on the same stream, then cost + push.  n_host: the count when the host knows it."""
        lib, st = self.lib, self._stream()
        h = self.heuristic
        if n_host is None and self.sync_free:
            kind, *res = h.eval_nodes_dev(self.env, self.arena, self.kept_ids, self._n_kept_ptr, self.max_cand)
        else:
            if n_host is None:                                      # one host round trip: how many children survived CLOSED
                self.sync_state()
                n_host = int(self.plan.n_kept)
            if n_host == 0:
                return
            if hasattr(h, "eval_nodes"):                            # tcgen05 path: one-hot input built straight from the arena
                kind, res = "h", [h.eval_nodes(self.env, self.arena, self.kept_ids, n_host)]
            else:
                fn = h.fn if isinstance(h, CapacityHeuristic) else h
                if self.nn_in is None:
                    self.nn_in = torch.empty((max(self.max_cand, self.n_inst), self.S), dtype=torch.uint8, device=self.dev)
                check(lib.dcb_gather_nnet_input(self.env, ptr(self.arena), ptr(self.kept_ids), n_host, ptr(self.nn_in), st), "gather_nnet_input")
                self.kernel_launches += 1
                out = fn(self.nn_in[:n_host])
                kind, res = "h", [out if (out.dtype == torch.float32 and out.is_contiguous()) else out.float().contiguous()]
        if self._phase_events is not None:
            self._phase_events[4].record()
        if kind == "dot":
            dpart, n_parts, bias = res
            check(lib.dcb_search_push(self._ctx_ref, None, ptr(dpart), int(n_parts), float(bias), st), "search_push")
        else:
            check(lib.dcb_search_push(self._ctx_ref, ptr(res[0]), None, 0, 0.0, st), "search_push")
        self.kernel_launches += 1

    def enqueue_step(self, include_solved: bool = False) -> None:
        """One iteration, no host synchronisation (unless the heuristic needs the row count on the host)."""
        lib, st = self.lib, self._stream()
        with torch.cuda.device(self.dev):
            ev = None
            if self.profile:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                ev[0].record()
            check(lib.dcb_search_pop(self._ctx_ref, 1 if include_solved else 0, st), "search_pop")
            if ev:
                ev[1].record()
            if self.expand_events is not None:
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
            check(lib.dcb_search_expand(self._ctx_ref, st), "search_expand")
            if self.expand_events is not None:
                e1 = torch.cuda.Event(enable_timing=True); e1.record()
                self.expand_events.append([e0, e1, None])
            if ev:
                ev[2].record()
            check(lib.dcb_search_closed(self._ctx_ref, st), "search_closed")
            if ev:
                ev[3].record()
            self._phase_events = ev
            self.kernel_launches += 15 + 1 + 4          # pop (14) + plan, expand, closed (probe, min, resolve, fix-up)
            self._heuristic_and_push()
            if ev:
                ev[5].record()

    def _after_state(self) -> None:
        """Host bookkeeping once an iteration's state has been read: grow CLOSED ahead of the iterations in flight."""
        margin = 2 * self.max_cand
        if 2 * (int(self.plan.closed_entries) + margin) > self.closed_cap and self.closed_cap < self.closed_cap_max:
            self._grow_closed(int(self.plan.closed_entries) + margin)
        if self.expand_events:
            for e in self.expand_events:
                if e[2] is None:
                    e[2] = int(self.plan.n_parents)

    def step_all(self, include_solved: bool = False) -> None:
        """One iteration of every running instance, then read the instance records (one host round trip at the END)."""
        self.enqueue_step(include_solved)
        self.sync_state()
        self._after_state()

    def running(self) -> int:
        return int(self.plan.n_running)

    # ---- results ---------------------------------------------------------------------------------------
    def path_to(self, node_id: int) -> List[int]:
        """Moves root -> node (parallel_weighted_astar.cpp:336-341 / astar.py:213-229)."""
        lib, st = self.lib, self._stream()
        with torch.cuda.device(self.dev):
            check(lib.dcb_search_path(self._ctx_ref, int(node_id), self.path_moves.numel(), ptr(self.path_moves), ptr(self.path_len), st),
                  "search_path")
            n = int(self.path_len.cpu().numpy().view(np.int32)[0])
            self.d2h_bytes += 4 + max(n, 0)
            if n < 0:
                raise _lib.DcbError("solution longer than %d moves" % self.path_moves.numel())
            return self.path_moves[:n].cpu().numpy().tolist()

    def node_states(self, ids: Sequence[int]) -> np.ndarray:
        """States of the given node ids, u8 [len(ids), S] on the host."""
        idx = torch.tensor(list(ids), dtype=torch.int64, device=self.dev)
        offs = idx[:, None] * self.S + torch.arange(self.S, device=self.dev)[None, :]
        return self.arena[offs].cpu().numpy()

    def popped_of(self, i: int) -> List[int]:
        """Node ids popped by instance i in the last iteration, in pop order."""
        n = int(self.inst[i].n_popped)
        return self.popped_ids[i * self.Bpad: i * self.Bpad + n].cpu().numpy().view(np.uint32).tolist()

    def kept_list(self) -> List[int]:
        n = int(self.plan.n_kept)
        return sorted(self.kept_ids[:n].cpu().numpy().view(np.uint32).tolist())


class BWASGpu(SearchEngine):
    """ONE start state at a time (the reference's `--language cpp` flow solves its states one after the other,
    astar.py:508-520): the single-instance face of SearchEngine with the attribute surface the CLI, bench.py and the tests use."""

    def __init__(self, env_name: str, heuristic, weight: float, batch_size: int, max_nodes: int = 1 << 24,
                 device: Optional[torch.device] = None, semantics: str = "cpp", sync_free: Optional[bool] = None):
        super().__init__(env_name, heuristic, weight, batch_size, n_inst=1, max_nodes=max_nodes, device=device, semantics=semantics,
                         sync_free=sync_free)
        self._set_fresh()

    def _set_fresh(self):
        self.nodes_generated = 1 if self.semantics == "cpp" else 0       # :166 vs astar.py:168
        self.nodes_expanded = 0          # children actually materialised (the terminating iteration's are only counted above)
        self.iterations = 0
        self.done = 0
        self.goal_id = NONE
        self.goal_ids: List[int] = []          # python semantics: every solved node popped so far
        self.last_popped = 0
        self.last_kept = 0
        self.next_slot = 1

    def reset(self, start: np.ndarray) -> None:
        super().reset(np.asarray(start, dtype=np.uint8).reshape(1, self.S))
        self._set_fresh()

    def _absorb(self) -> None:
        s = self.inst[0]
        self.nodes_generated = int(s.nodes_generated)
        self.nodes_expanded = int(s.nodes_expanded)
        self.iterations = int(s.iterations)
        self.done = int(s.done)
        self.goal_id = int(s.goal_id)
        self.last_popped = int(s.n_popped)
        self.last_kept = int(self.plan.n_kept)
        self.next_slot = int(s.next_slot)

    def step(self, keep_trace: bool = False) -> Optional[Dict]:
        """One BWAS iteration; the instance record is read back at the end.  Returns the trace record when asked."""
        self.step_all(include_solved=True)       # (Python semantics: the caller decides when to stop, as bwas_python does)
        self._absorb()
        rec = None
        if self.semantics == "python" and self.inst[0].n_goals > len(self.goal_ids):
            pid = self.popped_ids[:self.last_popped].long()
            sv = self.node_solved[pid].bool()
            self.goal_ids.extend(pid[sv].cpu().tolist())
        if keep_trace:
            rec = {"popped": self.popped_of(0), "kept": self.kept_list()}
        return rec

    def pipelined_steps(self):
        """Generator over the iterations of the current search without a host round trip in the loop: yields after every completed
        iteration (attributes absorbed) while the next one is already enqueued.  Stop iterating once `done` is set (the iteration in
        flight then pops nothing) or the budget (set_budget) has run out."""
        k = 0
        self.enqueue_step(); self._readback(0)
        while True:
            self.enqueue_step(); self._readback((k + 1) % 2)
            self._wait(k % 2)
            self._after_state()
            self._absorb()
            yield k
            k += 1

    def solve(self, start: np.ndarray, max_iters: Optional[int] = None, keep_trace: bool = False) -> BWASResult:
        t_begin = time.perf_counter()
        self.reset(start)
        trace: List[Dict] = []
        try:
            if keep_trace or max_iters is not None or not self.sync_free or self.profile:
                while not self.done:
                    if max_iters is not None and self.iterations >= max_iters:
                        break
                    rec = self.step(keep_trace)
                    if keep_trace:
                        trace.append(rec)
            else:
                # pipelined: iteration k+1 is enqueued before the host looks at iteration k; once `done` is set on the device the
                # speculative iteration pops nothing and costs a handful of empty launches
                k = 0
                self.enqueue_step(); self._readback(0)
                while True:
                    self.enqueue_step(); self._readback((k + 1) % 2)
                    self._wait(k % 2)
                    self._after_state()
                    if self.inst[0].done:
                        break
                    k += 1
                self._wait((k + 1) % 2)              # drain the speculative iteration (its record equals the final one)
                self._absorb()
        except _lib.DcbError:
            self._absorb()                           # counters of the interrupted search stay readable
            torch.cuda.current_stream(self.dev).synchronize()
            raise
        moves = self.path_to(self.goal_id) if self.goal_id != NONE else None
        s = self.inst[0]
        return BWASResult(moves=moves, nodes_generated=self.nodes_generated, iterations=self.iterations,
                          solve_time=time.perf_counter() - t_begin, path_cost=float(len(moves)) if moves is not None else float("nan"),
                          done=self.done, open_size=int(s.open_size), closed_size=int(self.plan.closed_entries),
                          timings=dict(self.timings), trace=trace if keep_trace else None)
