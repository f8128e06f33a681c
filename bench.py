#!/usr/bin/env python
"""bench.py -- batch-weighted-A* node expansions / second on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's C++ BWAS on the host cores
    python bench.py --full --num_states 1000                         # whole searches to completion (BASELINE configs[1] / [3])
    python bench.py --workload puzzle15|puzzle48                     # BASELINE configs[2] / [4]

Workload (BASELINE.json configs[1]): cube3 A*, weight 0.8, batch_size 20000, start states from the reference's scramble
generator generate_states(n, (0, 26)) under fixed seeds.  BOTH arms run the SAME start states under the SAME counting rule:

  * the bounded sample of the workload = the scrambles of depth >= 20 of that seeded list, in order (shallower scrambles are
    solved within a few dozen children and never reach a full batch); rank r takes every world-th state starting at r;
  * a STEP is one FULL-BATCH BWAS iteration -- pop 20000 nodes from OPEN, expand them (240000 children), CLOSED
    insert-or-improve, cost-to-go network on the surviving children, push.  The ramp-up iterations of a search (1, 12, 132 ...
    children) ride along inside the window when a search ends and the next state starts: their time and their nodes are counted,
    they are not steps;
  * `value` = nodes generated AND materialised (expanded, deduplicated, evaluated) inside the window / time.  The children of
    an iteration that fires the termination rule are part of the reference's `num_nodes_generated` but this engine never
    materialises them: they are NOT in the numerator.
  * warm-up = the iterations up to and including the W-th full-batch one.

Timed region: barrier + synchronize, K steps, synchronize + barrier; device time from CUDA events, max over ranks.  The per-step
working set (arena + CLOSED + OPEN, hundreds of MB) is larger than L2.  `full_search` in the line = whole searches run to
completion right after the window (the steady-state number the window slightly overstates).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "nodes/s"
BATCH = 20000
MIN_SCRAMBLE = 20
SEED = 1234
POOL = 1024                                            # size of the seeded scramble list both arms draw from
NCU_TRAFFIC_BYTES = 116.28e6 + 1527.19e6               # ncu --set full, expand_kernel<cube3>, 2^21 parents (profiles/expand_r01_ncu.txt)
# ncu --set full, resnet_gemm_pair_kernel, the 10 launches of one cube3 forward pass at 131072 rows (profiles/resnet_gemm_r02_ncu.txt):
# dram__bytes_read.sum 10.72 GB + dram__bytes_write.sum 7.46 GB
NCU_GEMM_BYTES_PER_ROW = (10.72e9 + 7.46e9) / 131072

# per workload: env name, weight, state bytes, moves, MFLOP per state of the cost-to-go net, BASELINE config
WORKLOADS = {
    "cube3": dict(env="cube3", weight=0.8, S=54, A=12, mflop=29.24, config="cube3 A* weight=0.8 batch_size=20000, scrambles depth<=26 (BASELINE configs[1])"),
    "puzzle15": dict(env="puzzle15", weight=0.8, S=16, A=4, mflop=28.56, config="puzzle15 A* weight=0.8 batch_size=20000 on data/puzzle15/test (BASELINE configs[2])"),
    "puzzle48": dict(env="puzzle48", weight=0.8, S=49, A=4, mflop=50.01, config="puzzle48 A* weight=0.8 batch_size=20000 on data/puzzle48/test (BASELINE configs[4])"),
}


def metric_name(wl):
    return "%s_astar_node_expansions_per_sec" % wl


def weights_path(env):
    return os.path.join(ROOT, "assets", "saved_models", env, "current", "model_state_dict.pt")


def measured_peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, bf16 TFLOP/s burst, source)"""
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", 1400.0)), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================
# the workload's start states -- ONE definition for both arms
# =====================================================================================================
def workload_states(wl: str, n: int, full: bool = False):
    """u8 [n, S] start states + a description.  cube3: the reference generator's semantics (environments/cube3.py:96-127)
    restated by the oracle's numpy port (identical to the GPU-backed environment, tests/test_gpu_cli.py); the puzzles use the
    reference's own test files (assets/, copied by tools/fetch_assets.py)."""
    w = WORKLOADS[wl]
    if wl == "cube3":
        import torch
        if torch.cuda.is_available():           # the product's GPU-backed environment (same RNG call order as the reference)
            from deepcubea_b200.utils.env_utils import get_environment
            env = get_environment("cube3")
            gen = lambda k: (lambda st, d: (env.pack(st), d))(*env.generate_states(k, (0, 26)))
        else:                                   # reference arm on a GPU-less host: the checker's numpy port of the same generator
            from oracle import oracle_env as O
            gen = lambda k: O.OracleCube3().generate_states(k, (0, 26))
        pool = n if full else POOL
        while True:
            np.random.seed(SEED); random.seed(SEED)
            states, depths = gen(pool)
            if full:
                return states[:n], "generate_states(%d,(0,26)), seed %d, every state" % (n, SEED)
            keep = np.nonzero(np.asarray(depths) >= MIN_SCRAMBLE)[0]
            if len(keep) >= n:
                return states[keep[:n]], "generate_states(%d,(0,26)), seed %d, scramble depth >= %d kept, first %d in order" % (pool, SEED, MIN_SCRAMBLE, n)
            pool *= 8                            # fixed tiers, so both arms always draw from the same list
    path = os.path.join(ROOT, "assets", "data", w["env"], "test", "data_0.pkl")
    if not os.path.exists(path):
        raise RuntimeError("%s missing: run tools/fetch_assets.py %s where /root/reference exists" % (path, w["env"]))
    sys.path.insert(0, ROOT)
    data = pickle.load(open(path, "rb"))
    attr = "tiles"
    arr = np.stack([np.asarray(getattr(s, attr), dtype=np.uint8) for s in data["states"][:n]])
    return arr, "first %d states of data/%s/test/data_0.pkl" % (len(arr), w["env"])


def build_heuristic(wl, device, precision: str):
    import torch
    from deepcubea_b200.nnet.folded import DeviceHeuristic, FoldedResnet
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.nnet_utils import load_nnet
    env = get_environment(WORKLOADS[wl]["env"])
    model = env.get_nnet_model()
    wp = weights_path(WORKLOADS[wl]["env"])
    if os.path.exists(wp):
        load_nnet(wp, model, device=torch.device("cpu"))
        src = "trained weights (assets/)"
    else:
        torch.manual_seed(0)
        for m in model.modules():                       # non-trivial BN statistics so folding is exercised
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
        src = "random-init weights (seed 0)"
    model.eval()
    if precision in ("fp16x3", "fp16"):                 # hand-written tcgen05 dense layers (csrc/resnet_kernels.cu)
        from deepcubea_b200.nnet.tc_resnet import TcResnet
        return TcResnet(model, device, mode=precision), src
    return DeviceHeuristic(FoldedResnet(model, mode=precision).to(device), chunk=1 << 17), src


def interval_union(spans):
    """Total length covered by a list of (start, end) intervals (overlaps counted once)."""
    total, cur_s, cur_e = 0.0, None, None
    for a, b in sorted(spans):
        if cur_e is None or a > cur_e:
            total += (cur_e - cur_s) if cur_e is not None else 0.0
            cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    return total + ((cur_e - cur_s) if cur_e is not None else 0.0)


# =====================================================================================================
class Lane:
    """One search engine on its own CUDA stream.  A GPU runs `--lanes` of them side by side on independent problem instances:
    the small kernels of one search (pop, expand, CLOSED, push) run while the tensor cores work on the other's network
    (tools/exp_two_streams.py: +3 % over one search at a time).  The host enqueues iteration k+1 of a lane before it reads the
    record of iteration k (BWASGpu.pipelined_steps)."""

    def __init__(self, idx, wl, dev, precision, max_nodes, states):
        import torch
        from deepcubea_b200.search.bwas_gpu import BWASGpu
        W = WORKLOADS[wl]
        self.idx, self.A, self.states = idx, W["A"], states
        self.heur, self.weights_src = build_heuristic(wl, dev, precision)
        self.eng = BWASGpu(W["env"], self.heur, W["weight"], BATCH, max_nodes=max_nodes, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.i, self.gen, self.seen, self.job = 0, None, 0, None
        self.target = self.full = self.nodes = self.iters = self.solved = 0
        self.lens, self.rows0 = [], 0

    def _ctx(self):
        import torch
        return torch.cuda.stream(self.stream)

    # ---- window mode: consecutive start states until `k_full` FULL-BATCH iterations were materialised ---------------------
    def begin(self, k_full):
        with self._ctx():
            self.eng.set_budget(k_full)       # device-side: the iteration in flight after the k_full-th full one is a no-op
        self.target, self.full, self.nodes, self.iters, self.solved, self.lens = k_full, 0, 0, 0, 0, []
        self.rows0 = self.eng.total_kept

    def finished(self):
        return self.full >= self.target

    def advance(self):
        from deepcubea_b200 import _lib
        eng = self.eng
        with self._ctx():
            if self.iters > 400 * max(1, self.target):
                raise _lib.DcbError("bench window: the searches never reach full batches")
            if self.gen is None:
                eng.reset(self.states[self.i % len(self.states)])
                self.gen, self.seen = eng.pipelined_steps(), 0
            next(self.gen); self.iters += 1
            got = eng.nodes_expanded - self.seen
            self.seen = eng.nodes_expanded
            self.nodes += got
            self.full += got == BATCH * self.A
            if not eng.done and not self.finished() and eng.next_slot + 3 * BATCH + 64 > eng.max_slots:
                eng.set_budget(0)                     # arena nearly full: nothing new may start ...
                next(self.gen); self.iters += 1       # ... but the iteration already in flight is real work: count it, then move on
                got = eng.nodes_expanded - self.seen; self.nodes += got; self.full += got == BATCH * self.A
                eng.set_budget(max(self.target - self.full, 0))
                eng.done = eng.done or 3
            if eng.done:
                if eng.done == 1:
                    self.solved += 1; self.lens.append(len(eng.path_to(eng.goal_id)))
                self.i += 1; self.gen = None

    def end(self):
        with self._ctx():
            self.eng.set_budget(None)
        return self.eng.total_kept - self.rows0

    # ---- whole-search mode -------------------------------------------------------------------------------------------
    def start(self, index, state):
        with self._ctx():
            self.eng.set_budget(None)
            self.eng.reset(state)
            self.gen = self.eng.pipelined_steps()
        self.job = (index, time.perf_counter())

    def poll(self):
        """Advance the running search by one iteration; returns its result tuple once it has ended, else None."""
        import torch
        from deepcubea_b200 import _lib
        eng = self.eng
        index, t0 = self.job
        with self._ctx():
            try:
                next(self.gen)
                if not eng.done:
                    return None
                n_moves = len(eng.path_to(eng.goal_id)) if eng.done == 1 else -1
            except _lib.DcbError:                 # node arena full: the search is abandoned, its work still counts (time AND nodes)
                eng._absorb()
                torch.cuda.current_stream().synchronize()
                n_moves = -2
        self.job = self.gen = None
        return (index, n_moves, int(eng.nodes_generated), time.perf_counter() - t0, int(eng.iterations))


def run_searches(lanes, next_index, states):
    """Whole searches to completion, one per lane at a time, instances drawn from `next_index()` (None = exhausted)."""
    results = []
    while True:
        busy = False
        for lane in lanes:
            if lane.job is None:
                i = next_index()
                if i is None:
                    continue
                lane.start(i, states[i])
            busy = True
            r = lane.poll()
            if r is not None:
                results.append(r)
        if not busy:
            return results


def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepcubea_b200 import _lib, ops
    from deepcubea_b200.search import sharding
    wl = args.workload
    W = WORKLOADS[wl]
    A = W["A"]
    rank, world, local = sharding.world()
    if not torch.cuda.is_available():
        raise _lib.DcbError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    max_nodes = args.max_nodes or ((1 << 28) if args.full else (1 << 27))
    n_lanes = max(1, args.lanes)

    def barrier():
        torch.cuda.synchronize()
        sharding.completion_barrier()

    if args.full:
        lanes = [Lane(j, wl, dev, args.nnet_precision, max_nodes, None) for j in range(n_lanes)]
        return run_full(args, lanes, rank, world, dev, barrier)

    n_inst = max(8, (args.steps + args.warmup) // 2)
    all_states, states_desc = workload_states(wl, n_inst * world)
    states = all_states[sharding.shard_indices(len(all_states), rank, world)]
    lanes = [Lane(j, wl, dev, args.nnet_precision, max_nodes, states[j::n_lanes]) for j in range(n_lanes)]
    heur, weights_src = lanes[0].heur, lanes[0].weights_src

    def run_window(k_full):
        """k_full FULL-BATCH iterations in total, split over the lanes; the lanes are advanced in turn.
        Returns (nodes materialised, rows evaluated, iterations, solved, solution lengths)."""
        for j, lane in enumerate(lanes):
            lane.begin(k_full // n_lanes + (1 if j < k_full % n_lanes else 0))
        while not all(l.finished() for l in lanes):
            for lane in lanes:
                if not lane.finished():
                    lane.advance()
        kept = sum(lane.end() for lane in lanes)
        return (sum(l.nodes for l in lanes), kept, sum(l.iters for l in lanes), sum(l.solved for l in lanes), [x for l in lanes for x in l.lens])

    run_window(args.warmup)
    # ---- device-resident timed region ----------------------------------------------------------------
    tc_heur = hasattr(heur, "gemm_events")
    for lane in lanes:
        lane.eng.expand_events = []
        if tc_heur:
            lane.heur.gemm_events = []
    gemm0 = sum(l.heur.gemm_launches for l in lanes) if tc_heur else 0
    launches0 = sum(l.eng.kernel_launches for l in lanes)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    prof = os.environ.get("DCB_CUDA_PROFILER") == "1"     # `ncu --profile-from-start off`: capture the timed region only
    if prof:
        torch.cuda.profiler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()                                          # (device idle after the barrier: every lane's work starts after this point)
    nodes, kept, iters, solved, lens = run_window(args.steps)
    for lane in lanes:
        torch.cuda.current_stream().wait_stream(lane.stream)
    ev1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    lane_steps = [int(l.full) for l in lanes]
    launches = sum(l.eng.kernel_launches for l in lanes) - launches0
    gemm_ev, gemm_busy_ms = [], 0.0
    if tc_heur:
        # With several lanes a GEMM launch may wait for the other lane's GEMM (both want every SM): its own event pair then spans
        # the wait too.  The tensor cores' busy time is the UNION of the launches' [start, end] intervals on the device clock.
        spans = sorted((ev0.elapsed_time(a), ev0.elapsed_time(b)) for l in lanes for a, b, _ in l.heur.gemm_events)
        gemm_ev = spans
        gemm_busy_ms = interval_union(spans)
        gemm_flops = heur.flops_per_row * float(kept)       # algorithmic: every surviving child passes through every layer once
        n_nn = sum(l.heur.gemm_launches for l in lanes) - gemm0
        launches += n_nn + max(1, n_nn // 10)              # + the one-hot kernel of each forward pass
    for lane in lanes:
        lane.eng.expand_events = None
        if tc_heur:
            lane.heur.gemm_events = None
    # ---- end-to-end through the public API: host start state in, host solution out ---------------------
    for lane in lanes:                                     # fresh searches; their ramp-up is untimed, as in the device window
        lane.i += 1 if lane.gen is not None else 0
        lane.gen = None
    run_window(args.warmup)
    h2d0, d2h0 = sum(l.eng.h2d_bytes for l in lanes), sum(l.eng.d2h_bytes for l in lanes)
    barrier()
    t0 = time.perf_counter()
    e_nodes, _, _, _, _ = run_window(args.steps)
    torch.cuda.synchronize()
    e_sec = time.perf_counter() - t0
    barrier()
    h2d, d2h = sum(l.eng.h2d_bytes for l in lanes) - h2d0, sum(l.eng.d2h_bytes for l in lanes) - d2h0
    # ---- the gather kernel at the A* loop's launch size: one lane alone, so that a launch never waits for the other lane's GEMM ----
    one = lanes[0]
    one.eng.expand_events = []
    one.begin(min(10, args.steps))
    while not one.finished():
        one.advance()
    one.end()
    torch.cuda.synchronize()
    in_loop = [(a.elapsed_time(b), n) for a, b, n in one.eng.expand_events if n and n == BATCH]
    one.eng.expand_events = None
    # ---- whole searches to completion (steady state incl. ramp-up, large OPEN / CLOSED and the final iteration) ----------
    for lane in lanes:
        lane.gen = None
    todo = [(max(l.i for l in lanes) * n_lanes + n_lanes + j) % len(states) for j in range(args.full_states)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = run_searches(lanes, lambda it=iter(todo): next(it, None), states)
    torch.cuda.synchronize()
    f_sec = (time.perf_counter() - t0) if res else 0.0
    f_nodes = float(sum(r[2] for r in res))
    f_lens = [r[1] for r in res if r[1] >= 0]
    f_solved = len(f_lens)
    # ---- rooflines ----------------------------------------------------------------------------------------------
    peak, tc_sus, tc_burst, peak_src = measured_peaks()
    roof = roof_dom = None
    if rank == 0 and gemm_ev:
        t_s = gemm_busy_ms * 1e-3
        fl = gemm_flops
        ach = fl / t_s / 1e12
        exec_ratio = (89.7 / 29.24) if args.nnet_precision == "fp16x3" else (29.9 / 29.24)
        roof_dom = {"kernel": "resnet_gemm_pair_kernel (tcgen05 cta_group::2 dense layers of the cost-to-go ResNet)", "bound": "tensor", "achieved": round(ach, 1),
                    "peak": tc_sus, "unit": "TFLOP/s", "frac": round(ach / tc_sus, 4),
                    "traffic": (round(NCU_GEMM_BYTES_PER_ROW * float(kept) / len(gemm_ev)) if wl == "cube3" else None),
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the ten launches of a forward pass at 131072 rows "
                                      "(profiles/resnet_gemm_r02_ncu.txt: 138.7 KB per row), scaled to this run's average rows per launch",
                    "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step); burst %.1f" % tc_burst,
                    "launches": len(gemm_ev), "avg_us": round(t_s / len(gemm_ev) * 1e6, 1), "share_of_timed_region": round(t_s * 1e3 / ms, 4),
                    "timing": "CUDA events around every launch on its lane's stream; busy time = union of the [start, end] intervals of all lanes",
                    "algorithmic_flops": "2*rows*N*K of the unpadded layer (%.2f MFLOP per %s state, SURVEY 8d), one product" % (W["mflop"], wl),
                    "note": "precision mode %s executes %s MMAs per algorithmic product (fp16 hi/lo operand pairs, fp32-parity: max |err| 3e-5 vs fp64) "
                            "on padded tiles; executed tensor work ~ %.0f TFLOP/s" % (args.nnet_precision, "3" if args.nnet_precision == "fp16x3" else "1", ach * exec_ratio)}
    if rank == 0 and wl == "cube3":
        alg = 54.0 / 12 + 54 + 1 + 8          # expand + is_solved + hash, unpadded
        n_par = 1 << 21
        g = torch.Generator(device=dev); g.manual_seed(0)
        par = torch.arange(54, dtype=torch.uint8, device=dev).repeat(n_par, 1)
        for a in torch.randint(0, 12, (12,), generator=g, device=dev).tolist():
            par = ops.next_state(0, par, a)
        ch = torch.empty((n_par, 12, 54), dtype=torch.uint8, device=dev)
        times = []
        for it in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.expand(0, par, out=ch); b.record(); torch.cuda.synchronize()
            if it >= 3:
                times.append(a.elapsed_time(b))
        t = float(np.mean(times)) * 1e-3
        ach = alg * n_par * 12 / t / 1e9
        roof = {"kernel": "expand_kernel<cube3> (expand+is_solved+hash)", "bound": "hbm", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": NCU_TRAFFIC_BYTES, "peak_source": peak_src,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this launch shape, profiles/expand_r01_ncu.txt "
                                  "(algorithmic bytes per launch: %.4g)" % (alg * n_par * 12),
                "launch": "%d parents -> %d children, outputs 1.7 GB > L2" % (n_par, n_par * 12),
                "alg_bytes_per_child": alg, "children_per_sec": round(n_par * 12 / t, 1)}
        if in_loop:
            tl = float(np.mean([x[0] for x in in_loop])) * 1e-3
            nl = float(np.mean([x[1] for x in in_loop])) * 12
            roof["in_loop"] = {"launches": len(in_loop), "avg_children": nl, "avg_us": round(tl * 1e6, 2),
                               "achieved": round(alg * nl / tl / 1e9, 1),
                               "note": "A* launches move ~16 MB each: launch-latency bound, not HBM bound"}
        del par, ch
    # ---- reduce over ranks -------------------------------------------------------------------------------------
    per_rank = [{"rank": rank, "ms": round(ms, 3), "nodes": int(nodes), "heuristic_rows": int(kept), "iterations": int(iters), "solved": int(solved),
                 "e2e_s": round(e_sec, 4), "e2e_nodes": int(e_nodes), "full_steps_per_lane": lane_steps}]
    len_sum = sum(lens)
    if world > 1:
        bucket = [None] * world
        dist.all_gather_object(bucket, per_rank[0])
        per_rank = bucket
        nodes, ms = sharding.reduce_throughput(nodes, ms, dev)            # nodes SUM over ranks, device time MAX over ranks
        e_nodes, e_sec = sharding.reduce_throughput(e_nodes, e_sec, dev)
        # whole searches differ in length from rank to rank: the job's rate is the sum of the ranks' own rates
        f_rate = torch.tensor([f_nodes / f_sec if f_sec else 0.0, f_nodes], dtype=torch.float64, device=dev)
        dist.all_reduce(f_rate, op=dist.ReduceOp.SUM)
        f_nodes = float(f_rate[1].item())
        f_sec = (f_nodes / float(f_rate[0].item())) if float(f_rate[0].item()) else 0.0
        cnt = torch.tensor([launches, solved, len_sum, h2d, d2h, kept, iters, f_solved, sum(f_lens)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches, solved, len_sum, h2d, d2h, kept, iters, f_solved, f_len_sum = cnt.tolist()
    else:
        f_len_sum = sum(f_lens)
    extra = multi = None
    arena_nodes = lanes[0].eng.max_nodes
    if rank == 0 and world == 1 and not args.no_extras:
        for lane in lanes:
            lane.eng = None
        torch.cuda.empty_cache()
        extra = {}
        for other in ("puzzle15", "puzzle48"):
            try:
                extra[other] = workload_summary(other, dev, args.nnet_precision, peak, tc_sus)
            except Exception as e:   # missing assets must never take the bench down
                extra[other] = {"unavailable": str(e)[:160]}
        try:
            multi = multi_instance_summary(dev, heur)
        except Exception as e:
            multi = {"unavailable": str(e)[:160]}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = reference_sample(wl, steps=3, warmup=1, use_gpu_heuristic=True)
            except Exception as e:  # the baseline must never take the bench down
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
        line = {"metric": metric_name(wl), "value": nodes / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 states / u64 hashes / f32 costs; heuristic GEMMs %s" % {
                    "fp32": "fp32 (cuBLAS SGEMM)", "tf32": "tf32 (cuBLAS)", "bf16": "bf16 (cuBLAS)",
                    "fp16x3": "fp16 hi/lo x3 products, fp32 accumulate (tcgen05, fp32-parity mode)",
                    "fp16": "fp16, fp32 accumulate (tcgen05)"}[args.nnet_precision],
                "data": "synthetic: " + states_desc + "; " + weights_src,
                "config": {"workload": W["config"], "states": states_desc,
                           "step": "one FULL-BATCH BWAS iteration (pop 20000, expand %dx, CLOSED, heuristic on survivors, push); ramp-up "
                                   "iterations of a new search ride along (time and nodes counted, not steps); the children of a "
                                   "terminating iteration are not materialised and not counted" % A,
                           "instances_per_gpu": len(states), "max_nodes": arena_nodes,
                           "parallelism": "instances sharded over %d GPU(s), no data-path collective; %d concurrent searches (CUDA streams) per GPU: the "
                                          "small kernels of one overlap the network of the other" % (world, n_lanes),
                           "lanes_per_gpu": n_lanes,
                           "l2": "working set (arena+CLOSED+OPEN) >> L2; roofline launches write 1.7 GB each",
                           "solved_in_timed_region": int(solved), "iterations_in_timed_region": int(iters),
                           "avg_children_per_step": nodes / args.steps / world, "avg_heuristic_rows_per_step": kept / args.steps / world,
                           "mean_solution_len": (len_sum / solved) if solved else None, "per_rank": per_rank},
                "e2e": {"value": e_nodes / e_sec, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps / world, "d2h_bytes_per_step": d2h / args.steps / world,
                        "note": "BWASGpu.reset(host state) / pipelined_steps() / path_to() on every lane by wall clock, same window rule; the search never "
                                "leaves HBM: the start state goes in, one 192-byte record per iteration and the solution come out"},
                "full_search": {"value": (f_nodes / f_sec) if f_sec else None, "unit": UNIT, "states": int(args.full_states * world), "solved": int(f_solved),
                                "nodes_generated": int(f_nodes), "mean_solution_len": (f_len_sum / f_solved) if f_solved else None,
                                "note": "whole searches to completion right after the window (reference counting: every generated child, "
                                        "wall clock incl. reset / ramp-up / path); python bench.py --full runs BASELINE configs[1]/[3] this way"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roof_dom if roof_dom is not None else roof,      # dominant kernel of the step
                "roofline_gather": roof,                                      # BASELINE.json: "gather-kernel HBM GB/s vs roofline"
                "other_workloads": extra,                                     # BASELINE configs[2] / [4]: whole searches on the reference's test states
                "multi_instance": multi,                                      # AStar(states, ...) API: many instances per step in one engine
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def workload_summary(wl, dev, precision, hbm_peak, tc_peak, n_states=3):
    """BASELINE configs[2] / [4] in short: whole searches on the first states of the reference's test file (weight 0.8, batch 20000),
    nodes/s by the reference's counting, the GEMM's algorithmic TFLOP/s inside the searches, the gather kernel at streaming size."""
    import torch
    from deepcubea_b200 import ops
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    W = WORKLOADS[wl]
    heur, src = build_heuristic(wl, dev, precision)
    eng = BWASGpu(W["env"], heur, W["weight"], BATCH, max_nodes=1 << 26, device=dev)
    states, desc = workload_states(wl, n_states + 1)
    eng.solve(states[n_states], max_iters=12)                        # warm the kernels / buffers
    torch.cuda.synchronize()
    if hasattr(heur, "gemm_events"):
        heur.gemm_events = []
    kept0 = eng.total_kept
    nodes = secs = 0.0
    lens = []
    for s in states[:n_states]:
        t0 = time.perf_counter()
        r = eng.solve(s)
        secs += time.perf_counter() - t0
        nodes += r.nodes_generated
        lens.append(len(r.moves) if r.moves is not None else -1)
    out = {"config": W["config"], "states": desc.replace("first %d" % (n_states + 1), "first %d" % n_states), "value": nodes / secs, "unit": UNIT,
           "nodes_generated": int(nodes), "solution_lens": lens, "weights": src}
    if hasattr(heur, "gemm_events") and heur.gemm_events:
        t_s = sum(a.elapsed_time(b) for a, b, _ in heur.gemm_events) * 1e-3
        ach = heur.flops_per_row * float(eng.total_kept - kept0) / t_s / 1e12
        out["roofline"] = {"kernel": "resnet_gemm_pair_kernel", "bound": "tensor", "achieved": round(ach, 1), "peak": tc_peak, "unit": "TFLOP/s",
                           "frac": round(ach / tc_peak, 4), "mflop_per_state": W["mflop"], "launches": len(heur.gemm_events),
                           "share_of_search_time": round(t_s / secs, 4)}
        heur.gemm_events = None
    S, A = W["S"], W["A"]
    alg = S / A + S + 1 + 8
    n_par = 1 << 22
    eid = eng.env
    par = torch.from_numpy(states[:1].repeat(n_par, 0)).to(dev)
    g = torch.Generator(device=dev); g.manual_seed(0)
    for a in torch.randint(0, A, (8,), generator=g, device=dev).tolist():
        par = ops.next_state(eid, par, a)
    ch = torch.empty((n_par, A, S), dtype=torch.uint8, device=dev)
    times = []
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.expand(eid, par, out=ch); b.record(); torch.cuda.synchronize()
        if it >= 2:
            times.append(a.elapsed_time(b))
    t = float(np.mean(times)) * 1e-3
    out["roofline_gather"] = {"kernel": "expand_kernel<%s>" % wl, "bound": "hbm", "achieved": round(alg * n_par * A / t / 1e9, 1), "peak": hbm_peak,
                              "unit": "GB/s", "frac": round(alg * n_par * A / t / 1e9 / hbm_peak, 4), "alg_bytes_per_child": alg,
                              "children_per_sec": round(n_par * A / t, 1)}
    del eng, heur, par, ch
    torch.cuda.empty_cache()
    return out


def multi_instance_summary(dev, heur, n_inst=64, batch=100, weight=0.6):
    """The reference's multi-instance API (AStar(states, env, heuristic_fn, weights); astar.py:232-317 -- what its GBFS / AVI updaters
    drive): 64 cube3 instances, batch 100 each, advanced together by ONE engine (one pop / expand / CLOSED / network call per step)
    against the same instances solved one after the other by a single-instance engine."""
    import torch
    from deepcubea_b200.search.engine import BWASGpu, SearchEngine
    from deepcubea_b200.utils.env_utils import get_environment
    env = get_environment("cube3")
    np.random.seed(SEED + 1); random.seed(SEED + 1)
    st, depths = env.generate_states(n_inst, (8, 14))
    starts = env.pack(st)
    eng = SearchEngine("cube3", heur, [weight] * n_inst, batch, n_inst=n_inst, max_nodes=1 << 25, device=dev, semantics="python")
    eng.raise_on_error = False

    def run_multi():
        eng.reset(starts)
        steps = 0
        while eng.running() and steps < 2000:
            eng.step_all(); steps += 1
        return steps, sum(int(r.nodes_generated) for r in eng.inst), sum(1 for r in eng.inst if r.n_goals)
    run_multi()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    steps, nodes, solved = run_multi()
    torch.cuda.synchronize(); t_multi = time.perf_counter() - t0
    del eng
    one = BWASGpu("cube3", heur, weight, batch, max_nodes=1 << 22, device=dev, semantics="python")

    def run_seq():
        it = nd = 0
        for s in starts:
            one.reset(s)
            while not one.goal_ids and not one.done and one.iterations < 2000:
                one.step()
            it += one.iterations; nd += one.nodes_generated
        return it, nd
    run_seq()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    it_seq, nodes_seq = run_seq()
    torch.cuda.synchronize(); t_seq = time.perf_counter() - t0
    del one
    torch.cuda.empty_cache()
    return {"workload": "%d cube3 instances (scramble depth 8-14), weight %.1f, batch_size %d per instance, Python AStar semantics" % (n_inst, weight, batch),
            "one_engine": {"steps": steps, "seconds": round(t_multi, 4), "nodes_generated": nodes, "nodes_per_sec": nodes / t_multi,
                           "heuristic_calls_per_step": 1, "instances_with_goal": solved},
            "one_instance_at_a_time": {"iterations": it_seq, "seconds": round(t_seq, 4), "nodes_generated": nodes_seq, "nodes_per_sec": nodes_seq / t_seq},
            "speedup": t_seq / t_multi, "same_nodes_generated": nodes == nodes_seq}


def run_full(args, lanes, rank, world, dev, barrier):
    """BASELINE configs[1] (1 GPU) / configs[3] (8 GPUs): num_states scrambles PER GPU solved to completion, whole instances
    handed out by a dynamic queue (a lane that draws easy instances simply takes more of them).  value = sum nodes generated
    (the reference's counting, scripts/compare_solutions.py:27-28) / wall time of the slowest rank."""
    import torch
    import torch.distributed as dist
    from deepcubea_b200.search import sharding
    wl = args.workload
    W = WORKLOADS[wl]
    weights_src = lanes[0].weights_src
    n_total = args.num_states * world
    states, desc = workload_states(wl, n_total, full=True)
    q = sharding.InstanceQueue(len(states))
    for lane in lanes:                                # warm the kernels / allocator on every lane of every rank
        with lane._ctx():
            lane.eng.solve(states[0], max_iters=6)
    barrier()
    t0 = time.perf_counter()
    mine = run_searches(lanes, q.next, states)
    torch.cuda.synchronize()
    my_sec = time.perf_counter() - t0
    barrier()
    wall = time.perf_counter() - t0
    rows = [(rank, my_sec, mine)]
    if world > 1:
        bucket = [None] * world
        dist.all_gather_object(bucket, rows[0])
        rows = bucket
    if rank == 0:
        allr = sorted(x for _, _, m in rows for x in m)
        nodes = sum(x[2] for x in allr)
        solved = [x for x in allr if x[1] >= 0]
        max_sec = max(s for _, s, _ in rows)
        line = {"metric": metric_name(wl), "mode": "full", "value": nodes / max_sec, "unit": UNIT, "n_gpus": world, "higher_is_better": True,
                "scaling": "weak", "data": "synthetic: " + desc + "; " + weights_src,
                "config": {"workload": W["config"], "states": desc, "instances": len(allr), "instances_per_gpu": args.num_states,
                           "queue": "dynamic whole-instance queue (c10d store fetch-add)" if world > 1 else "in order",
                           "lanes_per_gpu": len(lanes),
                           "per_rank": [{"rank": r, "seconds": round(s, 3), "instances": len(m), "nodes": sum(x[2] for x in m)} for r, s, m in rows]},
                "solved": len(solved), "unsolved": len(allr) - len(solved), "nodes_generated": int(nodes), "wall_s": round(wall, 3), "max_rank_s": round(max_sec, 3),
                "mean_solution_len": float(np.mean([x[1] for x in solved])) if solved else None,
                "mean_nodes_per_instance": nodes / max(1, len(allr)), "sum_instance_seconds": round(sum(x[3] for x in allr), 3),
                "balance": {"min_rank_s": round(min(s for _, s, _ in rows), 3), "max_rank_s": round(max_sec, 3)}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================
def reference_sample(wl: str, steps: int, warmup: int, use_gpu_heuristic: bool = True):
    """The reference's own C++ BWAS (oracle/_ref/parallel_weighted_astar, compiled from /root/reference/cpp) on the host
    cores (OpenMP, all threads), heuristic served over its AF_UNIX protocol by a plain PyTorch fp32 ResnetModel -- on the GPU
    when there is one, exactly how the reference's --language cpp path runs.  Same start states, same step / window / counting
    rule as run_ours: the window opens when the `warmup`-th full-batch request has been answered and closes when `steps` more
    have; every child of every request answered inside it counts (the reference materialises all of them)."""
    import torch
    from oracle.ref_runner import REF_BINARY, HeuristicServer, have_reference_binary
    if not have_reference_binary():
        raise RuntimeError("oracle/_ref/parallel_weighted_astar missing")
    from deepcubea_b200.utils.pytorch_models import ResnetModel
    W = WORKLOADS[wl]
    S, A = W["S"], W["A"]
    dev = torch.device("cuda:0") if (use_gpu_heuristic and torch.cuda.is_available()) else torch.device("cpu")
    depth = 6 if wl == "cube3" else S
    model = ResnetModel(S, depth, 5000, 1000, 4, 1, True)
    wp = weights_path(W["env"])
    if os.path.exists(wp):
        sd = torch.load(wp, map_location="cpu")
        model.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()})
    else:
        torch.manual_seed(0)
    model.eval().to(dev)
    stamps = []

    def heur(states: np.ndarray) -> np.ndarray:
        x = torch.from_numpy((states // 9).astype(np.uint8) if wl == "cube3" else states).to(dev)     # state_to_nnet_input (cube3.py:77-85)
        outs = []
        with torch.no_grad():
            for i in range(0, x.shape[0], 10000):            # --nnet_batch_size 10000 (train.sh)
                outs.append(model(x[i:i + 10000])[:, 0])
        out = torch.cat(outs).float().cpu().numpy()
        stamps.append((time.perf_counter(), states.shape[0]))
        return out

    n_inst = max(8, (steps + warmup) // 2)
    states, desc = workload_states(wl, n_inst)             # == rank 0's shard of run_ours at N=1
    srv = HeuristicServer(S, heur)
    nodes, t_first, t_last, full = 0, None, None, 0
    target = steps + warmup
    try:
        for s in states:
            # torchrun exports OMP_NUM_THREADS=1; the reference's OpenMP loops get every host thread, as in its own runs
            child_env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            p = subprocess.Popen([REF_BINARY, " ".join(str(int(v)) for v in s), str(W["weight"]), str(BATCH), srv.path, W["env"], "0"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=child_env)
            seen = 0
            while full < target:
                alive = p.poll() is None
                while seen < len(stamps) and full < target:
                    ts, n = stamps[seen]; seen += 1
                    if t_first is not None:
                        nodes += n; t_last = ts
                    if n == BATCH * A:
                        full += 1
                        if full == warmup:
                            t_first = ts
                if not alive and seen >= len(stamps):
                    break
                time.sleep(0.002)
            if p.poll() is None:
                p.kill()
            p.wait()
            stamps.clear()
            if full >= target:
                break
    finally:
        srv.close()
    if not nodes or t_last is None or t_last <= t_first:
        raise RuntimeError("reference sample produced no full-batch iteration")
    return {"value": nodes / (t_last - t_first), "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": "%d full-batch BWAS iterations (20000 pops, %d children each; ramp-up iterations in between counted) of "
                      "oracle/_ref/parallel_weighted_astar on %s, OpenMP on %d host threads, heuristic = PyTorch fp32 ResnetModel on %s over "
                      "the reference's AF_UNIX protocol" % (steps, BATCH * A, desc, os.cpu_count(), "cuda:0" if dev.type == "cuda" else "cpu"),
            "states": desc}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    wl = args.workload
    W = WORKLOADS[wl]
    try:
        r = reference_sample(wl, args.steps, max(args.warmup, 1))
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e).replace("\n", " ")[:200]}))
        return
    line = {"impl": "reference", "metric": metric_name(wl), "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": BATCH * W["A"] / r["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 states; f32 costs; heuristic fp32", "data": "synthetic: " + r["states"],
            "config": {"workload": W["config"], "states": r["states"],
                       "step": "one FULL-BATCH BWAS iteration of the reference C++ program (bounded sample; same states, window and counting rule as the CUDA arm)"},
            "cpu_baseline": r, "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=str, default="cube3", choices=sorted(WORKLOADS))
    ap.add_argument("--nnet_precision", type=str, default=os.environ.get("DCB_NNET_PRECISION", "fp16x3"), choices=["fp32", "tf32", "bf16", "fp16x3", "fp16"],
                    help="heuristic arithmetic: fp16x3 = hand-written tcgen05, fp32-parity (max err 3e-5 vs fp64; default); fp32 = cuBLAS SGEMM")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=2, help="concurrent searches (engines on their own CUDA streams) per GPU")
    ap.add_argument("--no_extras", action="store_true", help="skip the puzzle15 / puzzle48 / multi-instance summaries (N=1 only)")
    ap.add_argument("--max_nodes", type=int, default=0, help="node arena capacity per GPU (default 2^27; 2^28 with --full: at weight 0.8 / batch 20000 "
                    "the hardest of 1000 scrambles generate more than 1.3e8 nodes; the reference's own weight-0.6 runs needed up to 6.1e7)")
    ap.add_argument("--full_states", type=int, default=2, help="whole searches per GPU run to completion after the window (full_search in the line)")
    ap.add_argument("--full", action="store_true", help="run --num_states start states per GPU to completion (BASELINE configs[1]/[3]) instead of the window")
    ap.add_argument("--num_states", type=int, default=1000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
