"""Side-by-side statistics of two results.pkl files (the reference's scripts/compare_solutions.py): solve times, solution
lengths, nodes generated, nodes per second -- the metric BASELINE.json quotes -- and how often the two agree on length."""
import os
import pickle
import sys
from argparse import ArgumentParser

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # results hold environments.* State objects


def summary(values) -> str:
    v = np.asarray(values, dtype=np.float64)
    return "Min/Max/Median/Mean(Std) %f/%f/%f/%f(%f)" % (v.min(), v.max(), np.median(v), v.mean(), v.std())


def describe(results) -> None:
    seconds = np.asarray(results["times"], dtype=np.float64)
    nodes = np.asarray(results["num_nodes_generated"], dtype=np.float64)
    lengths = [len(s) for s in results["solutions"]]
    for title, column in (("Times", seconds), ("Lengths", lengths), ("Nodes Generated", nodes), ("Nodes/Sec", nodes / seconds)):
        print("-%s-" % title)
        print(summary(column))


def main():
    ap = ArgumentParser()
    ap.add_argument("--soln1", type=str, required=True)
    ap.add_argument("--soln2", type=str, required=True)
    opt = ap.parse_args()
    first, second = (pickle.load(open(path, "rb")) for path in (opt.soln1, opt.soln2))
    print("%i states" % len(first["states"]))
    for tag, res in (("SOLUTION 1", first), ("SOLUTION 2", second)):
        print("\n--%s---" % tag)
        describe(res)
    delta = np.array([len(b) - len(a) for a, b in zip(first["solutions"], second["solutions"])])
    print("\n\n------Solution 2 - Solution 1 Lengths-----")
    print(summary(delta))
    print("%.2f%% soln2 equal to soln1" % (100.0 * np.mean(delta == 0)))


if __name__ == "__main__":
    main()
