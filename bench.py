#!/usr/bin/env python
"""bench.py -- cube3 batch-weighted-A* node expansions / second on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W                  # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's C++ BWAS on the host cores

Workload (BASELINE.json configs[1]): cube3 A*, weight 0.8, batch_size 20000, start states from the
reference's scramble generator generate_states(n, (0, 26)) under fixed seeds.  A STEP is one BWAS iteration:
pop <= 20000 nodes from OPEN, expand them (240k children), CLOSED insert-or-improve, cost-to-go network on the
surviving children, push.  Steps run back to back over consecutive start states (a solved state is followed by
the next one).  `value` = nodes generated (every child, the reference's Nodes/Sec numerator) / time.

Timed region: barrier + synchronize, K steps, synchronize + barrier; device time from CUDA events, max over ranks.
The per-step working set (arena + CLOSED + OPEN, hundreds of MB) is larger than L2.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cube3_astar_node_expansions_per_sec"
UNIT = "nodes/s"
WEIGHT, BATCH = 0.8, 20000
ALG_BYTES_PER_CHILD = 54.0 / 12 + 54 + 1 + 8          # expand + is_solved + hash, unpadded
WEIGHTS = os.path.join(ROOT, "assets", "saved_models", "cube3", "current", "model_state_dict.pt")
NCU_TRAFFIC_BYTES = 116.28e6 + 1527.19e6               # ncu --set full, expand_kernel<cube3>, 2^21 parents (profiles/expand_r01_ncu.txt)


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
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


def make_states(n: int, seed: int) -> np.ndarray:
    """Synthetic start states: the reference generator's semantics (environments/cube3.py:96-127), run by this
    repo's GPU-backed environment."""
    from deepcubea_b200.utils.env_utils import get_environment
    env = get_environment("cube3")
    np.random.seed(seed); random.seed(seed)
    states, _ = env.generate_states(n, (0, 26))
    return env.pack(states)


def build_heuristic(device, precision: str):
    import torch
    from deepcubea_b200.nnet.folded import DeviceHeuristic, FoldedResnet
    from deepcubea_b200.utils.env_utils import get_environment
    from deepcubea_b200.utils.nnet_utils import load_nnet
    env = get_environment("cube3")
    model = env.get_nnet_model()
    if os.path.exists(WEIGHTS):
        load_nnet(WEIGHTS, model, device=torch.device("cpu"))
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


# =====================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from deepcubea_b200 import _lib, ops
    from deepcubea_b200.search.bwas_gpu import BWASGpu
    from deepcubea_b200.search import sharding
    rank, world, local = sharding.world()
    if not torch.cuda.is_available():
        raise _lib.DcbError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    heur, weights_src = build_heuristic(dev, args.nnet_precision)
    steps_total = args.steps + args.warmup
    max_nodes = int(min(1 << 27, max(1 << 24, 2 * steps_total * BATCH * 12)))
    eng = BWASGpu("cube3", heur, WEIGHT, BATCH, max_nodes=max_nodes, device=dev)
    # instances shard by rank (instance i -> rank i % world): weak scaling, no data-path collective
    n_inst = max(4, steps_total // 8)
    all_states = make_states(n_inst * world, seed=1234)
    states = all_states[sharding.shard_indices(len(all_states), rank, world)]

    def barrier():
        torch.cuda.synchronize()
        sharding.completion_barrier()

    def run_steps(k, cursor):
        """k BWAS iterations over consecutive start states; returns (nodes, solved, lens, cursor)."""
        nodes, solved, lens = 0, 0, []
        while k > 0:
            if cursor["fresh"]:
                eng.reset(states[cursor["i"] % len(states)]); cursor["fresh"] = False
            before = eng.nodes_generated
            eng.step(); k -= 1
            nodes += eng.nodes_generated - before
            if not eng.done and eng.next_slot + 2 * BATCH + 64 > eng.max_slots:
                eng.done = 3                      # arena nearly full (never with trained weights): move on to the next instance
            if eng.done:
                if eng.done == 1:
                    solved += 1; lens.append(len(eng.path_to(eng.goal_id)))
                cursor["i"] += 1; cursor["fresh"] = True
        return nodes, solved, lens

    cursor = {"i": 0, "fresh": True}
    run_steps(args.warmup, cursor)
    # ---- device-resident timed region ----------------------------------------------------------------
    eng.expand_events = []
    tc_heur = hasattr(heur, "gemm_events")
    if tc_heur:
        heur.gemm_events = []
        gemm0 = heur.gemm_launches
    launches0 = eng.kernel_launches
    kept0 = eng.total_kept
    sampler = ClockSampler(local); sampler.start()
    barrier()
    prof = os.environ.get("DCB_CUDA_PROFILER") == "1"     # `ncu --profile-from-start off`: capture the timed region only
    if prof:
        torch.cuda.profiler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    nodes, solved, lens = run_steps(args.steps, cursor)
    ev1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.kernel_launches - launches0
    kept = eng.total_kept - kept0
    gemm_ev = []
    if tc_heur:
        gemm_ev = [(a.elapsed_time(b), f) for a, b, f in heur.gemm_events]
        heur.gemm_events = None
        n_nn = heur.gemm_launches - gemm0
        launches += n_nn + 2 * max(1, n_nn // 14)          # + one-hot and fc_out kernels of each forward pass
    in_loop = [(a.elapsed_time(b), n) for a, b, n in eng.expand_events]
    eng.expand_events = None
    # ---- end-to-end through the public API: host start state in, host solution out ---------------------
    cursor2 = {"i": cursor["i"] + 1, "fresh": True}
    h2d0, d2h0 = eng.h2d_bytes, eng.d2h_bytes
    barrier()
    t0 = time.perf_counter()
    e_nodes, _, _ = run_steps(args.steps, cursor2)
    torch.cuda.synchronize()
    e_sec = time.perf_counter() - t0
    barrier()
    h2d, d2h = eng.h2d_bytes - h2d0, eng.d2h_bytes - d2h0
    # ---- gather-kernel roofline: streaming-size launches of the same kernel, CUDA events ------------------
    peak, tc_sus, tc_burst, peak_src = measured_peaks()
    roof = None
    roof_dom = None
    if rank == 0 and gemm_ev:
        # the dominant kernel of the step (95% of device time, profiles/launches_r01_summary.txt): the tcgen05 dense layers
        t_s = sum(x[0] for x in gemm_ev) * 1e-3
        fl = sum(x[1] for x in gemm_ev)
        ach = fl / t_s / 1e12
        roof_dom = {"kernel": "resnet_gemm_kernel (tcgen05 dense layers of the cost-to-go ResNet)", "bound": "tensor", "achieved": round(ach, 1),
                    "peak": tc_sus, "unit": "TFLOP/s", "frac": round(ach / tc_sus, 4), "traffic": None, "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step); burst %.1f" % tc_burst,
                    "launches": len(gemm_ev), "avg_us": round(t_s / len(gemm_ev) * 1e6, 1), "share_of_timed_region": round(t_s * 1e3 / ms, 4),
                    "algorithmic_flops": "2*rows*N*K of the unpadded layer (29.24 MFLOP per cube3 state, SURVEY 8d), one product",
                    "note": "precision mode %s executes %s MMAs per algorithmic product (fp16 hi/lo operand pairs, fp32-parity: max |err| 2e-5 vs fp64) "
                            "on tiles padded to 256x64; executed tensor work = %.0f TFLOP/s; ncu: tensor pipe active 77-89%% (profiles/resnet_gemm_r01_ncu.txt); "
                            "DRAM traffic per K=N=1024 launch at 131072 rows (ncu --set full): 1.03 GB without / 1.60 GB with a residual input vs "
                            "1.08 / 1.61 GB algorithmic (A hi+lo read, out hi+lo written, residual hi+lo read) -- no re-reads; traffic is null above "
                            "because the in-loop launches differ in row count"
                            % (args.nnet_precision, "3" if args.nnet_precision == "fp16x3" else "1",
                               ach * (89.7 / 29.24 if args.nnet_precision == "fp16x3" else 29.9 / 29.24))}
    if rank == 0:
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
        ach = ALG_BYTES_PER_CHILD * n_par * 12 / t / 1e9
        roof = {"kernel": "expand_kernel<cube3> (expand+is_solved+hash)", "bound": "hbm", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": NCU_TRAFFIC_BYTES, "peak_source": peak_src,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this launch shape, profiles/expand_r01_ncu.txt "
                                  "(algorithmic bytes per launch: %.4g)" % (ALG_BYTES_PER_CHILD * n_par * 12),
                "launch": "%d parents -> %d children, outputs 1.7 GB > L2" % (n_par, n_par * 12),
                "alg_bytes_per_child": ALG_BYTES_PER_CHILD, "children_per_sec": round(n_par * 12 / t, 1)}
        if in_loop:
            tl = float(np.mean([x[0] for x in in_loop])) * 1e-3
            nl = float(np.mean([x[1] for x in in_loop])) * 12
            roof["in_loop"] = {"launches": len(in_loop), "avg_children": nl, "avg_us": round(tl * 1e6, 2),
                               "achieved": round(ALG_BYTES_PER_CHILD * nl / tl / 1e9, 1),
                               "note": "A* launches move ~16 MB each: launch-latency bound, not HBM bound"}
        del par, ch
    # ---- reduce over ranks -------------------------------------------------------------------------------------
    len_sum = sum(lens)
    if world > 1:
        nodes, ms = sharding.reduce_throughput(nodes, ms, dev)            # nodes SUM over ranks, device time MAX over ranks
        e_nodes, e_sec = sharding.reduce_throughput(e_nodes, e_sec, dev)
        cnt = torch.tensor([launches, solved, len_sum, h2d, d2h], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches, solved, len_sum, h2d, d2h = cnt.tolist()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = reference_sample(steps=3, warmup=1, use_gpu_heuristic=True)
            except Exception as e:  # the baseline must never take the bench down
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
        line = {"metric": METRIC, "value": nodes / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 states / u64 hashes / f32 costs; heuristic GEMMs %s" % {
                    "fp32": "fp32 (cuBLAS SGEMM)", "tf32": "tf32 (cuBLAS)", "bf16": "bf16 (cuBLAS)",
                    "fp16x3": "fp16 hi/lo x3 products, fp32 accumulate (tcgen05, fp32-parity mode)",
                    "fp16": "fp16, fp32 accumulate (tcgen05)"}[args.nnet_precision],
                "data": "synthetic cube3 scrambles (generate_states(n,(0,26)), seed 1234); " + weights_src,
                "config": {"workload": "cube3 A* weight=0.8 batch_size=20000, scrambles depth<=26 (BASELINE configs[1])",
                           "step": "one BWAS iteration (pop<=20000, expand 12x, CLOSED, heuristic on survivors, push)",
                           "instances_per_gpu": len(states), "max_nodes": max_nodes, "parallelism": "instances sharded over %d GPU(s)" % world,
                           "l2": "working set (arena+CLOSED+OPEN) >> L2; roofline launches write 1.7 GB each",
                           "solved_in_timed_region": int(solved), "avg_children_per_step": nodes / args.steps / world,
                           "avg_heuristic_rows_per_step": kept / args.steps, "mean_solution_len": (len_sum / solved) if solved else None},
                "e2e": {"value": e_nodes / e_sec, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                        "note": "BWASGpu.reset(host state)/step()/path_to() wall clock; the search never leaves HBM, only the start "
                                "state goes in and counters/solution come out"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": roof_dom if roof_dom is not None else roof,      # dominant kernel of the step
                "roofline_gather": roof,                                      # BASELINE.json: "gather-kernel HBM GB/s vs roofline"
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================
def reference_sample(steps: int, warmup: int, use_gpu_heuristic: bool = True):
    """The reference's own C++ BWAS (oracle/_ref/parallel_weighted_astar, compiled from /root/reference/cpp) on
    the host cores (OpenMP, all threads), heuristic served over its AF_UNIX protocol by a plain PyTorch fp32
    ResnetModel -- on the GPU when there is one, exactly how the reference's --language cpp path runs."""
    import torch
    from oracle import oracle_env as O
    from oracle.ref_runner import REF_BINARY, HeuristicServer, have_reference_binary
    if not have_reference_binary():
        raise RuntimeError("oracle/_ref/parallel_weighted_astar missing")
    from deepcubea_b200.utils.pytorch_models import ResnetModel
    dev = torch.device("cuda:0") if (use_gpu_heuristic and torch.cuda.is_available()) else torch.device("cpu")
    model = ResnetModel(54, 6, 5000, 1000, 4, 1, True)
    if os.path.exists(WEIGHTS):
        sd = torch.load(WEIGHTS, map_location="cpu")
        model.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()})
    else:
        torch.manual_seed(0)
    model.eval().to(dev)
    env = O.OracleCube3()
    stamps = []

    def heur(states: np.ndarray) -> np.ndarray:
        x = torch.from_numpy(env.nnet_input(states)).to(dev)
        outs = []
        with torch.no_grad():
            for i in range(0, x.shape[0], 10000):            # --nnet_batch_size 10000 (train.sh)
                outs.append(model(x[i:i + 10000])[:, 0])
        out = torch.cat(outs).float().cpu().numpy()
        stamps.append((time.perf_counter(), states.shape[0]))
        return out

    np.random.seed(1234); random.seed(1234)
    states, _ = env.generate_states(8, (20, 26))
    srv = HeuristicServer(54, heur)
    nodes, t_first, t_last = 0, None, None
    target = steps + warmup
    try:
        for s in states:
            stamps.clear()
            # torchrun exports OMP_NUM_THREADS=1; the reference's OpenMP loops get every host thread, as in its own runs
            child_env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
            p = subprocess.Popen([REF_BINARY, " ".join(str(int(v)) for v in s), str(WEIGHT), str(BATCH), srv.path, "cube3", "0"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=child_env)
            # only full-batch iterations count as steps (the first ~6 requests of a search are the ramp-up)
            full = 0
            seen = 0
            while p.poll() is None and full < target:
                time.sleep(0.005)
                while seen < len(stamps):
                    ts, n = stamps[seen]; seen += 1
                    if n == BATCH * 12:
                        full += 1
                        if full == warmup:
                            t_first = ts
                        elif full > warmup and t_first is not None:
                            nodes += n; t_last = ts
                    if full >= target:
                        break
            p.kill(); p.wait()
            if full >= target:
                break
    finally:
        srv.close()
    if not nodes or t_last is None or t_last <= t_first:
        raise RuntimeError("reference sample produced no full-batch iteration")
    return {"value": nodes / (t_last - t_first), "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": "%d full BWAS iterations (20000 pops, 240000 children each) of oracle/_ref/parallel_weighted_astar, OpenMP on %d "
                      "host threads, heuristic = PyTorch fp32 ResnetModel on %s over the reference's AF_UNIX protocol"
                      % (steps, os.cpu_count(), "cuda:0" if dev.type == "cuda" else "cpu")}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    try:
        r = reference_sample(args.steps, max(args.warmup, 1))
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": str(e).replace("\n", " ")[:200]}))
        return
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 240000.0 / r["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 states; f32 costs; heuristic fp32", "data": "synthetic cube3 scrambles (depth 20-26, seed 1234)",
            "config": {"workload": "cube3 A* weight=0.8 batch_size=20000, scrambles depth<=26 (BASELINE configs[1])",
                       "step": "one BWAS iteration of the reference C++ program (bounded sample)"},
            "cpu_baseline": r, "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--nnet_precision", type=str, default=os.environ.get("DCB_NNET_PRECISION", "fp16x3"), choices=["fp32", "tf32", "bf16", "fp16x3", "fp16"],
                    help="heuristic arithmetic: fp16x3 = hand-written tcgen05, fp32-parity (max err 2e-5 vs fp64; default); fp32 = cuBLAS SGEMM")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
